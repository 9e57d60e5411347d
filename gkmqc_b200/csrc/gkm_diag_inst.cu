/* gkm_diag_inst.cu -- instantiates the bit-sliced kernel for ONE word length.
 * Compiled once per L = 2..16 with -DGKM_INST_L=<L> (see Makefile) so that the
 * 90 specialisations (15 L x {4,8,16} bins x {plain, weighted}) build in parallel. */
#include "gkm_diag_kernel.cuh"

#ifndef GKM_INST_L
#error "compile with -DGKM_INST_L=<word length>"
#endif

#define GKM_CAT2(a, b) a##b
#define GKM_CAT(a, b) GKM_CAT2(a, b)

extern "C" const void *GKM_CAT(gkm_diag_fn_L, GKM_INST_L)(int nb, int weighted)
{
    constexpr int L = GKM_INST_L;
    if (nb == 4) return weighted ? (const void *) gkm_diag_kernel<L, 4, true> : (const void *) gkm_diag_kernel<L, 4, false>;
    if (nb == 8) return weighted ? (const void *) gkm_diag_kernel<L, 8, true> : (const void *) gkm_diag_kernel<L, 8, false>;
    if (nb == 16) return weighted ? (const void *) gkm_diag_kernel<L, 16, true> : (const void *) gkm_diag_kernel<L, 16, false>;
    return nullptr;
}
