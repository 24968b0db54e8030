// k_step_tb instantiations for vx>=0: true, vy>=0: true (see step_tb_inst.cuh)
#include "step_tb_inst.cuh"
namespace csim {
cudaError_t tb_launch_pp(int T, int mode, const TbArgs& a, cudaStream_t stream) {
    return tb_launch_signed<true, true>(T, mode, a, stream);
}
}  // namespace csim
