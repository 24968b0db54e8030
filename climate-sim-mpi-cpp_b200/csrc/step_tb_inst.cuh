// step_tb_inst.cuh — instantiation helper: each step_inst_<xy>.cu defines one launcher for a fixed
// pair of upwind selectors (p: v >= 0, n: v < 0, z: v == +0.0 with the term dropped, see tb_update)
// and switches over (kernel, T, MODE) at run time.  Splitting by selector keeps the translation units
// compiling in parallel.
#pragma once
#include <cuda_runtime.h>

#include "step_tb.cuh"
#include "step_tbs.cuh"

namespace csim {

// staged = false: k_step_tb  (level-0 rows loaded into registers one tick ahead),
// staged = true : k_step_tbs (level-0 rows landed in a shared-memory ring by TMA; exists for T >= 3 in the
//                 multiplication modes, which is where the loop spends its time)
int tb_carveout_env();  // kernels.cu: CSIM_CARVEOUT (percent) or -1

template <int VXS, int VYS>
cudaError_t tb_launch_signed(bool staged, int T, int mode, const TbArgs& a, cudaStream_t stream) {
    const dim3 block(32 * kTbWarpsPerBlock);
    const dim3 grid((a.n_items + kTbWarpsPerBlock - 1) / kTbWarpsPerBlock);
#define CSIM_TB_CASE(TT, MM)                                                  \
    if (!staged && T == TT && mode == MM) {                                   \
        k_step_tb<TT, MM, VXS, VYS><<<grid, block, 0, stream>>>(a);           \
        return cudaGetLastError();                                            \
    }
#define CSIM_TBS_CASE(TT, MM)                                                 \
    if (staged && T == TT && mode == MM) {                                    \
        static const bool prepared = [] {                                     \
            const int pct = tb_carveout_env();                                \
            if (pct >= 0)                                                     \
                cudaFuncSetAttribute(k_step_tbs<TT, MM, VXS, VYS>,            \
                                     cudaFuncAttributePreferredSharedMemoryCarveout, pct); \
            return true;                                                      \
        }();                                                                  \
        (void)prepared;                                                       \
        k_step_tbs<TT, MM, VXS, VYS><<<grid, block, 0, stream>>>(a);          \
        return cudaGetLastError();                                            \
    }
    CSIM_TBS_CASE(3, MODE_UNIT)
    CSIM_TBS_CASE(4, MODE_UNIT)
    CSIM_TBS_CASE(3, MODE_RECIP)
    CSIM_TBS_CASE(4, MODE_RECIP)
    if (VXS != 0 && VYS != 0) {
        CSIM_TB_CASE(1, MODE_UNIT)
        CSIM_TB_CASE(2, MODE_UNIT)
        CSIM_TB_CASE(3, MODE_UNIT)
        CSIM_TB_CASE(4, MODE_UNIT)
        CSIM_TB_CASE(1, MODE_RECIP)
        CSIM_TB_CASE(2, MODE_RECIP)
        CSIM_TB_CASE(3, MODE_RECIP)
        CSIM_TB_CASE(4, MODE_RECIP)
        CSIM_TB_CASE(1, MODE_DIV)
        CSIM_TB_CASE(2, MODE_DIV)
        CSIM_TB_CASE(3, MODE_DIV)
    } else {
        // dropped-term variants exist for the blocking depths the loop spends its time in; the few
        // remainder sweeps (nsteps % T) take the full-arithmetic kernels, which give the same bits
        CSIM_TB_CASE(3, MODE_UNIT)
        CSIM_TB_CASE(4, MODE_UNIT)
        CSIM_TB_CASE(3, MODE_RECIP)
        CSIM_TB_CASE(4, MODE_RECIP)
    }
#undef CSIM_TB_CASE
#undef CSIM_TBS_CASE
    return cudaErrorInvalidValue;
}

// vxs, vys in {-1, 0, +1}
cudaError_t tb_launch(int vxs, int vys, bool staged, int T, int mode, const TbArgs& a, cudaStream_t stream);
bool tb_has_zero_variant(int T, int mode);
bool tb_has_staged_variant(int T, int mode);

#define CSIM_DECLARE_LAUNCH(name) \
    cudaError_t tb_launch_##name(bool staged, int T, int mode, const TbArgs& a, cudaStream_t stream);
CSIM_DECLARE_LAUNCH(pp) CSIM_DECLARE_LAUNCH(pn) CSIM_DECLARE_LAUNCH(np) CSIM_DECLARE_LAUNCH(nn)
CSIM_DECLARE_LAUNCH(pz) CSIM_DECLARE_LAUNCH(nz) CSIM_DECLARE_LAUNCH(zp) CSIM_DECLARE_LAUNCH(zn)
CSIM_DECLARE_LAUNCH(zz)
#undef CSIM_DECLARE_LAUNCH

}  // namespace csim
