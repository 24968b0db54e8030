// k_step_tb instantiations for vx>=0: false, vy>=0: false (see step_tb_inst.cuh)
#include "step_tb_inst.cuh"
namespace csim {
cudaError_t tb_launch_nn(int T, int mode, const TbArgs& a, cudaStream_t stream) {
    return tb_launch_signed<false, false>(T, mode, a, stream);
}
}  // namespace csim
