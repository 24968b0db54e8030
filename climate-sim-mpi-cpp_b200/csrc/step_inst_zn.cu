// dispatch-target instantiations for upwind selectors vx: 0, vy: -1 (see step_tb_inst.cuh)
#include "step_tb_inst.cuh"
namespace csim {
cudaError_t tb_launch_zn(bool staged, int T, int mode, const TbArgs& a, cudaStream_t stream) {
    return tb_launch_signed<0, -1>(staged, T, mode, a, stream);
}
}  // namespace csim
