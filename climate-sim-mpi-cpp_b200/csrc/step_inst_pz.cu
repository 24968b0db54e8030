// dispatch-target instantiations for upwind selectors vx: 1, vy: 0 (see step_tb_inst.cuh)
#include "step_tb_inst.cuh"
namespace csim {
cudaError_t tb_launch_pz(bool staged, int T, int mode, const TbArgs& a, cudaStream_t stream) {
    return tb_launch_signed<1, 0>(staged, T, mode, a, stream);
}
}  // namespace csim
