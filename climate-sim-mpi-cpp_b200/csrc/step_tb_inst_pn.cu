// k_step_tb instantiations for vx>=0: true, vy>=0: false (see step_tb_inst.cuh)
#include "step_tb_inst.cuh"
namespace csim {
cudaError_t tb_launch_pn(int T, int mode, const TbArgs& a, cudaStream_t stream) {
    return tb_launch_signed<true, false>(T, mode, a, stream);
}
}  // namespace csim
