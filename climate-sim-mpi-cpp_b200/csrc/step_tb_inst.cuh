// step_tb_inst.cuh — instantiation helper: each step_tb_inst_<sign>.cu defines one launcher for a
// fixed pair of upwind directions and switches over (T, MODE) at run time.  Splitting by sign keeps
// the four translation units compiling in parallel.
#pragma once
#include <cuda_runtime.h>

#include "step_tb.cuh"

namespace csim {

template <bool VXP, bool VYP>
cudaError_t tb_launch_signed(int T, int mode, const TbArgs& a, cudaStream_t stream) {
    const dim3 block(32 * kTbWarpsPerBlock);
    const dim3 grid((a.n_items + kTbWarpsPerBlock - 1) / kTbWarpsPerBlock);
#define CSIM_TB_CASE(TT, MM)                                             \
    if (T == TT && mode == MM) {                                         \
        k_step_tb<TT, MM, VXP, VYP><<<grid, block, 0, stream>>>(a);      \
        return cudaGetLastError();                                       \
    }
    CSIM_TB_CASE(1, MODE_UNIT)
    CSIM_TB_CASE(2, MODE_UNIT)
    CSIM_TB_CASE(3, MODE_UNIT)
    CSIM_TB_CASE(4, MODE_UNIT)
    CSIM_TB_CASE(1, MODE_RECIP)
    CSIM_TB_CASE(2, MODE_RECIP)
    CSIM_TB_CASE(3, MODE_RECIP)
    CSIM_TB_CASE(4, MODE_RECIP)
    CSIM_TB_CASE(1, MODE_DIV)
#undef CSIM_TB_CASE
    return cudaErrorInvalidValue;
}

cudaError_t tb_launch_pp(int T, int mode, const TbArgs& a, cudaStream_t stream);
cudaError_t tb_launch_pn(int T, int mode, const TbArgs& a, cudaStream_t stream);
cudaError_t tb_launch_np(int T, int mode, const TbArgs& a, cudaStream_t stream);
cudaError_t tb_launch_nn(int T, int mode, const TbArgs& a, cudaStream_t stream);

}  // namespace csim
