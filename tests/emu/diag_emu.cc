/* tests/emu/diag_emu.cc -- CPU lane emulator for the bit-sliced diagonal kernel.
 *
 * TEST INFRASTRUCTURE: runs gkm_bitslice.h (the exact arithmetic the sm_100a kernel
 * executes per lane) serially on the host, over the packed image produced by the
 * product's own packer (gkm_seq.c), so that packing, circular indexing, the
 * sliding-window adders and the binning can be checked against the oracle in the
 * CPU-only test tier.  It is not a fallback: nothing in the product links it.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../gkmqc_b200/csrc/gkm_internal.h"
#include "../../gkmqc_b200/csrc/gkm_bitslice.h"

/* the device layer is absent in the emulator build */
extern "C" {
void gkm_dev_release(gkmb200_problem *) {}
}

template <int L, int NB, bool WEIGHTED>
static void emu_pair(const gkmb200_problem *p, int a, int b, int32_t *H)
{
    const int W = p->Wmax;
    const int la = p->len[a], lb = p->len[b];
    const int Wa = (la + 31) / 32, Wb = (lb + 31) / 32;
    std::vector<gkm_apos> apos((size_t) 32 * Wa);
    const uint32_t *pa = p->planes + (size_t) a * 4 * W;
    const uint8_t *wea = p->weighted ? p->wend + (size_t) a * 2 * 32 * W : NULL;
    for (int e = 0; e < 32 * Wa; e++) {
        gkm_apos r;
        r.a0 = ((pa[0 * W + (e >> 5)] >> (e & 31)) & 1u) ? ~0u : 0u;
        r.a1 = ((pa[1 * W + (e >> 5)] >> (e & 31)) & 1u) ? ~0u : 0u;
        r.va = (e >= L - 1 && e < la) ? ~0u : 0u;
        r.wa = wea ? wea[e] : 1u;
        apos[(size_t) e] = r;
    }
    std::vector<uint32_t> E((size_t) Wb, 0u);
    for (int j = L - 1; j < lb; j++) E[(size_t) (j >> 5)] |= 1u << (j & 31);
    int32_t acc[NB];
    for (int m = 0; m < NB; m++) acc[m] = 0;
    const uint32_t *pb = p->planes + (size_t) b * 4 * W;
    for (int strand = 0; strand < 2; strand++) {
        const uint32_t *S0 = pb + (size_t) (2 * strand) * W;
        const uint32_t *S1 = pb + (size_t) (2 * strand + 1) * W;
        const uint8_t *we = p->weighted ? p->wend + ((size_t) b * 2 + strand) * 32 * W : NULL;
        for (int q = 0; q < Wb; q++)
            gkm_diag_lane<L, NB, WEIGHTED>(apos.data(), la, S0, S1, E.data(), Wb, q, we, acc);
    }
    for (int m = 0; m <= p->param.d; m++) H[m] = acc[m];
}

template <int L>
static int emu_dispatch_nb(const gkmb200_problem *p, int a, int b, int32_t *H)
{
    const int d = p->param.d;
    if (d < 4) { if (p->weighted) emu_pair<L, 4, true>(p, a, b, H); else emu_pair<L, 4, false>(p, a, b, H); }
    else if (d < 8) { if (p->weighted) emu_pair<L, 8, true>(p, a, b, H); else emu_pair<L, 8, false>(p, a, b, H); }
    else { if (p->weighted) emu_pair<L, 16, true>(p, a, b, H); else emu_pair<L, 16, false>(p, a, b, H); }
    return 0;
}

extern "C" int gkm_emu_hist(gkmb200_problem *p, int a, int b, int32_t *H)
{
    if (gkm_pack_problem(p)) return 1;
    switch (p->param.L) {
#define CASE(x) case x: return emu_dispatch_nb<x>(p, a, b, H);
        CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9)
        CASE(10) CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16)
#undef CASE
        default: return 1;
    }
}

/* all pairs j < a of the problem: H[(a*n + j)*(d+1) + m] */
extern "C" int gkm_emu_hist_lower(gkmb200_problem *p, int32_t *H)
{
    const int n = p->n, nb = p->param.d + 1;
    for (int a = 0; a < n; a++)
        for (int b = 0; b <= a; b++)
            if (gkm_emu_hist(p, a, b, H + ((size_t) a * n + b) * nb)) return 1;
    return 0;
}
