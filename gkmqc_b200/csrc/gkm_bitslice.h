/* gkm_bitslice.h -- the bit-sliced "diagonal" formulation of the truncated
 * mismatch histogram, shared verbatim by the sm_100a kernel (gkm_diag_kernel.cuh)
 * and by the CPU lane emulator in tests/emu (which exists so that this arithmetic
 * can be checked against the oracle on a machine without a GPU).
 *
 * Replaces, by a different algorithm, what the reference computes with its k-mer
 * tree DFS (libgkm.c:315-387) and, for the diagonal, with an XOR + byte-table
 * loop (libgkm.c:738-751):  H_m(a,b) = sum over L-mer pairs at Hamming distance
 * m <= d of wt_a * wt_b.
 *
 * Formulation.  Fix query a, target strand s (forward or reverse complement of
 * b, circularly extended to P = 32*Wb positions) and a diagonal delta.  Let
 *     mm_delta[e] = [ a[e] != s[(e + delta) mod P] ]          (one bit)
 * The L-mer pair (window of a ENDING at e, window of s ending at j = e + delta)
 * has Hamming distance  sum_{t<L} mm_delta[e - t].  One 32-bit word holds 32
 * consecutive diagonals delta = 32q .. 32q+31 (one word per lane), and the kernel
 * walks e = 0, 1, 2, ...: the window sum is a sliding sum ACROSS successive words,
 * so it needs no shifts at all -- only bit-sliced adders (LOP3 xor3 / majority)
 * over a register ring of the last 16 words.  Per 32 L-mer pairs: 3 funnel shifts
 * (target planes and validity mask seen through a bit window that slides with e),
 * 2 LOP3 for the mismatch word, ~12 LOP3 for the sliding count (L = 11), ~6 for
 * thresholding into the d+1 bins, one POPC per bin.  The canonical XOR/POPC
 * formulation needs ~5 integer operations per PAIR.
 *
 * Validity: windows that run off either sequence, wrap around the circular
 * extension or end before position L-1 are removed by AND-ing the hit word with
 *   Va[e]  (query side, warp-uniform word: all-ones iff L-1 <= e < len_a) and
 *   E[j]   (target side bit plane: 1 iff L-1 <= j < len_b),
 * so garbage in padding bits can never be counted.
 */
#ifndef GKM_BITSLICE_H_INCLUDED
#define GKM_BITSLICE_H_INCLUDED

#include <stdint.h>

#if defined(__CUDACC__)
#define GKM_HD __host__ __device__ __forceinline__
#define GKM_HDM __host__ __device__ __forceinline__
#else
#define GKM_HD static inline
#define GKM_HDM inline
#endif

/* query-side record for one position e, read as one 16-byte broadcast load */
struct alignas(16) gkm_apos {
    uint32_t a0; /* all-ones iff low  bit of base code a[e] is 1 */
    uint32_t a1; /* all-ones iff high bit of base code a[e] is 1 */
    uint32_t va; /* all-ones iff a window may end at e           */
    uint32_t wa; /* positional weight of the window ending at e  */
};

GKM_HD uint32_t gkm_funnel_r(uint32_t lo, uint32_t hi, uint32_t s)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return s ? ((lo >> s) | (hi << (32u - s))) : lo;
#endif
}

GKM_HD int gkm_popc(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

GKM_HD int gkm_ffs0(uint32_t x) /* index of lowest set bit, x != 0 */
{
#if defined(__CUDA_ARCH__)
    return __ffs((int) x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

GKM_HD uint32_t gkm_xor3(uint32_t a, uint32_t b, uint32_t c) { return a ^ b ^ c; }
GKM_HD uint32_t gkm_maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (a & c) | (b & c); }

/* ---- bit-sliced unsigned numbers: plane i holds bit i of 32 independent counters ---- */
constexpr int gkm_nplanes(int maxv) { return maxv >= 16 ? 5 : maxv >= 8 ? 4 : maxv >= 4 ? 3 : maxv >= 2 ? 2 : 1; }

template <int MAXV>
struct gkm_bs {
    static constexpr int NP = gkm_nplanes(MAXV);
    uint32_t p[NP];
    GKM_HDM uint32_t plane(int i) const { return i < NP ? p[i] : 0u; }
};

/* X + Y (+ carry-in plane) by ripple carry: one xor3 and one majority per plane */
template <int MX, int MY, bool CIN>
GKM_HD gkm_bs<MX + MY + (CIN ? 1 : 0)> gkm_bs_add(const gkm_bs<MX> &x, const gkm_bs<MY> &y, uint32_t cin)
{
    gkm_bs<MX + MY + (CIN ? 1 : 0)> r;
    constexpr int NR = gkm_bs<MX + MY + (CIN ? 1 : 0)>::NP;
    uint32_t c = CIN ? cin : 0u;
#pragma unroll
    for (int i = 0; i < NR; i++) {
        const uint32_t xi = x.plane(i), yi = y.plane(i);
        r.p[i] = gkm_xor3(xi, yi, c);
        c = gkm_maj3(xi, yi, c);
    }
    return r;
}

/* X + one bit plane */
template <int MX>
GKM_HD gkm_bs<MX + 1> gkm_bs_inc(const gkm_bs<MX> &x, uint32_t bit)
{
    gkm_bs<MX + 1> r;
    constexpr int NR = gkm_bs<MX + 1>::NP;
    uint32_t c = bit;
#pragma unroll
    for (int i = 0; i < NR; i++) {
        const uint32_t xi = x.plane(i);
        r.p[i] = xi ^ c;
        c = xi & c;
    }
    return r;
}

/* ---- per-lane sliding state: the last 16 mismatch words and the last RW 3-sums ---- */
template <int L>
struct gkm_win_state {
    static constexpr int Q = L / 3;                   /* number of 3-sums in the window  */
    static constexpr int R = L % 3;                   /* left-over single positions      */
    static constexpr int RW = (Q >= 4) ? 16 : 8;      /* ring length for the 3-sums      */
    uint32_t m[16];
    uint32_t t0[RW], t1[RW];                          /* planes of w3[e] = m[e]+m[e-1]+m[e-2] */
    GKM_HDM void clear()
    {
#pragma unroll
        for (int i = 0; i < 16; i++) m[i] = 0u;
#pragma unroll
        for (int i = 0; i < RW; i++) { t0[i] = 0u; t1[i] = 0u; }
    }
};

/* accumulate the remaining 3-sums (index I..Q-1) and single bits (index BI..R-1) */
template <int L, int S, int I, int BI, int MV>
GKM_HD auto gkm_win_accum(const gkm_bs<MV> &acc, const gkm_win_state<L> &st)
{
    constexpr int Q = gkm_win_state<L>::Q, R = gkm_win_state<L>::R, RW = gkm_win_state<L>::RW;
    if constexpr (I < Q) {
        gkm_bs<3> t;
        t.p[0] = st.t0[(S - 3 * I) & (RW - 1)];
        t.p[1] = st.t1[(S - 3 * I) & (RW - 1)];
        if constexpr (BI < R) {
            auto n = gkm_bs_add<MV, 3, true>(acc, t, st.m[(S - (3 * Q + BI)) & 15]);
            return gkm_win_accum<L, S, I + 1, BI + 1>(n, st);
        } else {
            auto n = gkm_bs_add<MV, 3, false>(acc, t, 0u);
            return gkm_win_accum<L, S, I + 1, BI>(n, st);
        }
    } else if constexpr (BI < R) {
        auto n = gkm_bs_inc<MV>(acc, st.m[(S - (3 * Q + BI)) & 15]);
        return gkm_win_accum<L, S, I, BI + 1>(n, st);
    } else {
        return acc;
    }
}

/* push mismatch word `mm` of step S (compile-time ring slot) and return the bit-sliced
 * number of mismatches of the L-window ending here (value 0..L per bit lane) */
template <int L, int S>
GKM_HD gkm_bs<L> gkm_win_push(gkm_win_state<L> &st, uint32_t mm)
{
    constexpr int Q = gkm_win_state<L>::Q, RW = gkm_win_state<L>::RW;
    st.m[S & 15] = mm;
    if constexpr (Q == 0) {
        gkm_bs<1> b; b.p[0] = mm;
        if constexpr (L == 1) return b;
        else return gkm_bs_inc<1>(b, st.m[(S - 1) & 15]);
    } else {
        const uint32_t m1 = st.m[(S - 1) & 15], m2 = st.m[(S - 2) & 15];
        gkm_bs<3> t;
        t.p[0] = gkm_xor3(mm, m1, m2);
        t.p[1] = gkm_maj3(mm, m1, m2);
        st.t0[S & (RW - 1)] = t.p[0];
        st.t1[S & (RW - 1)] = t.p[1];
        return gkm_win_accum<L, S, 1, 0>(t, st);
    }
}

/* ---- binning: NB = 4, 8 or 16 histogram bins (count values 0..NB-1) ---- */
template <int NB> struct gkm_log2nb;
template <> struct gkm_log2nb<4> { static constexpr int v = 2; };
template <> struct gkm_log2nb<8> { static constexpr int v = 3; };
template <> struct gkm_log2nb<16> { static constexpr int v = 4; };

/* word of bit lanes whose count is < NB, restricted to `valid` */
template <int L, int NB>
GKM_HD uint32_t gkm_hit_word(const gkm_bs<L> &cnt, uint32_t valid)
{
    constexpr int LB = gkm_log2nb<NB>::v;
    uint32_t high = 0u;
#pragma unroll
    for (int i = LB; i < gkm_bs<L>::NP; i++) high |= cnt.p[i];
    return valid & ~high;
}

/* mask of bit lanes (within hit) whose count equals V */
template <int L, int NB, int V>
GKM_HD uint32_t gkm_bin_mask(const gkm_bs<L> &cnt, uint32_t hit)
{
    constexpr int LB = gkm_log2nb<NB>::v;
    uint32_t r = hit;
#pragma unroll
    for (int i = 0; i < LB; i++) r &= ((V >> i) & 1) ? cnt.plane(i) : ~cnt.plane(i);
    return r;
}

template <int L, int NB, int V>
GKM_HD void gkm_bins_popc(const gkm_bs<L> &cnt, uint32_t hit, int32_t *acc)
{
    if constexpr (V < NB) {
        if constexpr (V <= L) acc[V] += gkm_popc(gkm_bin_mask<L, NB, V>(cnt, hit));
        gkm_bins_popc<L, NB, V + 1>(cnt, hit, acc);
    }
}

/* weighted bins: rare path, one (e, j) pair per set bit of `hit` */
template <int L, int NB>
GKM_HD void gkm_bins_weighted(const gkm_bs<L> &cnt, uint32_t hit, uint32_t wa, const uint8_t *wend,
                              int jbase_lo, int jbase_hi, int s, int32_t *acc)
{
    constexpr int LB = gkm_log2nb<NB>::v;
    while (hit) {
        const int bit = gkm_ffs0(hit);
        hit &= hit - 1u;
        int v = 0;
#pragma unroll
        for (int i = 0; i < LB; i++) v |= (int) ((cnt.plane(i) >> bit) & 1u) << i;
        const int off = s + bit;
        const int j = (off < 32) ? (jbase_lo + off) : (jbase_hi + off - 32);
        const int w = (int) wa * (int) wend[j];
#pragma unroll
        for (int b = 0; b < NB; b++) acc[b] += (v == b) ? w : 0;
    }
}

/* one step e = 32*c + S of one lane */
template <int L, int NB, bool WEIGHTED, int S>
GKM_HD void gkm_diag_step(gkm_win_state<L> &st, const gkm_apos &ap,
                          uint32_t lo0, uint32_t hi0, uint32_t lo1, uint32_t hi1, uint32_t loE, uint32_t hiE,
                          const uint8_t *wend, int jbase_lo, int jbase_hi, int32_t *acc)
{
    const uint32_t s0 = gkm_funnel_r(lo0, hi0, S);
    const uint32_t s1 = gkm_funnel_r(lo1, hi1, S);
    const uint32_t ev = gkm_funnel_r(loE, hiE, S) & ap.va;
    const uint32_t mm = (s0 ^ ap.a0) | (s1 ^ ap.a1);
    const gkm_bs<L> cnt = gkm_win_push<L, S>(st, mm);
    const uint32_t hit = gkm_hit_word<L, NB>(cnt, ev);
    if constexpr (WEIGHTED) {
        if (hit) gkm_bins_weighted<L, NB>(cnt, hit, ap.wa, wend, jbase_lo, jbase_hi, S, acc);
    } else {
        gkm_bins_popc<L, NB, 0>(cnt, hit, acc);
    }
}

template <int L, int NB, bool WEIGHTED, int S0, int S1>
GKM_HD void gkm_diag_steps(gkm_win_state<L> &st, const gkm_apos *ap,
                           uint32_t lo0, uint32_t hi0, uint32_t lo1, uint32_t hi1, uint32_t loE, uint32_t hiE,
                           const uint8_t *wend, int jbase_lo, int jbase_hi, int32_t *acc)
{
    if constexpr (S0 < S1) {
        gkm_diag_step<L, NB, WEIGHTED, S0>(st, ap[S0], lo0, hi0, lo1, hi1, loE, hiE, wend, jbase_lo, jbase_hi, acc);
        gkm_diag_steps<L, NB, WEIGHTED, S0 + 1, S1>(st, ap, lo0, hi0, lo1, hi1, loE, hiE, wend, jbase_lo, jbase_hi, acc);
    }
}

/* One lane = one (target strand, block of 32 diagonals q) against one query.
 *   apos      query records, 32*Wa of them (positions >= len_a have va = 0)
 *   len_a     query length
 *   S0,S1,E   target strand bit planes and valid-window-end plane, Wb words each (circular)
 *   wend      WEIGHTED: target weights by window end position, 32*Wb bytes
 *   acc       NB counters, added to                                                     */
template <int L, int NB, bool WEIGHTED>
GKM_HD void gkm_diag_lane(const gkm_apos *apos, int len_a,
                          const uint32_t *S0, const uint32_t *S1, const uint32_t *E, int Wb, int q,
                          const uint8_t *wend, int32_t *acc)
{
    gkm_win_state<L> st;
    st.clear();
    int k = q;
    uint32_t lo0 = S0[k], lo1 = S1[k], loE = E[k];
    const int Wa = (len_a + 31) >> 5;
    for (int c = 0; c < Wa; c++) {
        const int kn = (k + 1 == Wb) ? 0 : k + 1;
        const uint32_t hi0 = S0[kn], hi1 = S1[kn], hiE = E[kn];
        const gkm_apos *ap = apos + 32 * c;
        gkm_diag_steps<L, NB, WEIGHTED, 0, 16>(st, ap, lo0, hi0, lo1, hi1, loE, hiE, wend, 32 * k, 32 * kn, acc);
        if (32 * c + 16 < len_a)
            gkm_diag_steps<L, NB, WEIGHTED, 16, 32>(st, ap, lo0, hi0, lo1, hi1, loE, hiE, wend, 32 * k, 32 * kn, acc);
        lo0 = hi0; lo1 = hi1; loE = hiE;
        k = kn;
    }
}

#endif /* GKM_BITSLICE_H_INCLUDED */
