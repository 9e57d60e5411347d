/* gkm_lmer_kernel.cuh -- sm_100a kernel "lmer": candidate (a) in its canonical form.
 *
 * One L-mer pair = XOR of two packed words, fold the two bit planes, POPC, compare
 * with d (the scalar form of the reference's own sqnorm loop, libgkm.c:738-751,
 * with the 64 Ki byte table replaced by LOP3 + POPC).  Any L <= 16 and any d <= 15
 * at run time, weighted or not: this is the general fallback and the cross-check
 * for the bit-sliced kernel; ~5 integer operations per pair, so it is the slower
 * of the two by design.
 *
 * L-mer word: bits 0..L-1 = low code bit of the L bases, bits 16..16+L-1 = high
 * code bit.  mismatches(x,y) = popc(((x^y) | (x^y)>>16) & 0xFFFF).
 * Words are cut out of the same circular both-strand bit planes the diag kernel
 * reads: forward L-mer j starts at bit j, reverse-complement L-mer j at bit len+j.
 *
 * CTA = 256 threads, tile TA x TB sequence pairs.  A warp owns one (a,b) pair at a
 * time: lanes stride over the 2*nk_b target L-mers (both strands) holding 4 of them
 * in registers, the query L-mers are broadcast from shared memory.  Hits are rare
 * (0.12 % of pairs at L=11, d=3) and go to the tile histogram with shared atomics.
 */
#ifndef GKM_LMER_KERNEL_CUH_INCLUDED
#define GKM_LMER_KERNEL_CUH_INCLUDED

#include "gkm_diag_kernel.cuh" /* gkm_emit_entry */

#define GKM_LMER_THREADS 256
#define GKM_LMER_RJ 4

/* WA = 32-position chunks of the longest sequence: at most 32*WA L-mers per strand */
__host__ __device__ inline unsigned gkm_lmer_smem_bytes(int WA, int TA, int TB, int nbins, int weighted)
{
    unsigned nk = 32u * (unsigned) WA;
    unsigned o = ((unsigned) TA + 2u * (unsigned) TB) * nk * 4u;     /* L-mer words */
    if (weighted) o += ((unsigned) TA + 2u * (unsigned) TB) * nk;    /* weights by L-mer start */
    o += (unsigned) TA * (unsigned) TB * (unsigned) nbins * 4u;      /* histograms */
    o += ((unsigned) TA + (unsigned) TB) * 4u;
    return (o + 15u) & ~15u;
}

/* L bases starting at bit position `o` of the circular string */
__device__ __forceinline__ uint32_t gkm_lmer_word(const uint32_t *pl0, const uint32_t *pl1, int W, int o, uint32_t maskL)
{
    const int wi = o >> 5, sh = o & 31;
    const uint32_t n0 = (wi + 1 < W) ? pl0[wi + 1] : 0u, n1 = (wi + 1 < W) ? pl1[wi + 1] : 0u;
    const uint32_t p0 = __funnelshift_r(pl0[wi], n0, sh) & maskL;
    const uint32_t p1 = __funnelshift_r(pl1[wi], n1, sh) & maskL;
    return p0 | (p1 << 16);
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(GKM_LMER_THREADS)
gkm_lmer_kernel(const __grid_constant__ gkm_kparams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int W = p.W, TA = p.TA, TB = p.TB, L = p.L, d = p.d, NBN = p.nbins;
    const int NK = 32 * p.WA;
    uint32_t *sIdA = reinterpret_cast<uint32_t *>(smem);           /* [TA][NK] */
    uint32_t *sIdB = sIdA + (size_t) TA * NK;                       /* [TB][2][NK] */
    uint8_t *sWtA = reinterpret_cast<uint8_t *>(sIdB + (size_t) TB * 2 * NK);
    uint8_t *sWtB = sWtA + (WEIGHTED ? (size_t) TA * NK : 0);
    int32_t *sH = reinterpret_cast<int32_t *>(sWtB + (WEIGHTED ? (size_t) TB * 2 * NK : 0));
    int *sLenA = reinterpret_cast<int *>(sH + (size_t) TA * TB * NBN);
    int *sLenB = sLenA + TA;

    const int tid = threadIdx.x;
    const int row0 = p.row_begin + (int) blockIdx.y * TA;
    const int col0 = p.col_begin + (int) blockIdx.x * TB;
    const int row_last = min(row0 + TA, p.row_end) - 1;
    const int col_last = min(col0 + TB, p.col_end) - 1;
    if (p.mode == GKM_MODE_LOWER && col0 > row_last) return;
    if (p.mode == GKM_MODE_DIAG && (col0 > row_last || col_last < row0)) return;

    if (tid < TA) sLenA[tid] = (row0 + tid < p.row_end) ? p.lens[row0 + tid] : 0;
    if (tid < TB) sLenB[tid] = (col0 + tid < p.col_end) ? p.lens[col0 + tid] : 0;
    for (int i = tid; i < TA * TB * NBN; i += GKM_LMER_THREADS) sH[i] = 0;
    __syncthreads();

    const uint32_t maskL = (L >= 16) ? 0xFFFFu : ((1u << L) - 1u);
    /* query L-mers: forward strand only */
    for (int i = tid; i < TA * NK; i += GKM_LMER_THREADS) {
        const int a = i / NK, j = i - a * NK;
        uint32_t v = 0u; uint8_t wt = 0;
        if (j + L <= sLenA[a]) {
            const uint32_t *pl = p.planes + (size_t) (row0 + a) * 3 * W;
            v = gkm_lmer_word(pl, pl + W, W, j, maskL);
            if (WEIGHTED) wt = p.wend[(size_t) (row0 + a) * 32 * W + j + L - 1];
        }
        sIdA[i] = v;
        if (WEIGHTED) sWtA[i] = wt;
    }
    /* target L-mers: forward and reverse-complement strands */
    for (int i = tid; i < TB * 2 * NK; i += GKM_LMER_THREADS) {
        const int b = i / (2 * NK), r = i - b * 2 * NK, strand = r / NK, j = r - strand * NK;
        uint32_t v = 0u; uint8_t wt = 0;
        if (j + L <= sLenB[b]) {
            const uint32_t *pl = p.planes + (size_t) (col0 + b) * 3 * W;
            const int o = strand * sLenB[b] + j;
            v = gkm_lmer_word(pl, pl + W, W, o, maskL);
            if (WEIGHTED) wt = p.wend[(size_t) (col0 + b) * 32 * W + o + L - 1];
        }
        sIdB[i] = v;
        if (WEIGHTED) sWtB[i] = wt;
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;
    for (int pair = warp; pair < TA * TB; pair += GKM_LMER_THREADS / 32) {
        const int a_l = pair / TB, b_l = pair - a_l * TB;
        const int a_g = row0 + a_l, b_g = col0 + b_l;
        if (a_g >= p.row_end || b_g >= p.col_end) continue;
        if (p.mode == GKM_MODE_LOWER && b_g >= a_g) continue;
        if (p.mode == GKM_MODE_DIAG && b_g != a_g) continue;
        const int nkA = sLenA[a_l] - L + 1, nkB = sLenB[b_l] - L + 1;
        const uint32_t *xa = sIdA + (size_t) a_l * NK;
        const uint32_t *yb = sIdB + (size_t) b_l * 2 * NK;
        int32_t *h = sH + (size_t) pair * NBN;
        /* target index jj in [0, 2*nkB): strand = jj / nkB */
        for (int j0 = 0; j0 < 2 * nkB; j0 += 32 * GKM_LMER_RJ) {
            uint32_t y[GKM_LMER_RJ];
            int wy[GKM_LMER_RJ];
            bool ok[GKM_LMER_RJ];
#pragma unroll
            for (int r = 0; r < GKM_LMER_RJ; r++) {
                const int jj = j0 + 32 * r + lane;
                ok[r] = jj < 2 * nkB;
                const int strand = (jj >= nkB) ? 1 : 0;
                const int idx = strand * NK + (jj - strand * nkB);
                y[r] = ok[r] ? yb[idx] : 0u;
                wy[r] = (WEIGHTED && ok[r]) ? (int) sWtB[(size_t) b_l * 2 * NK + idx] : 1;
            }
            for (int i = 0; i < nkA; i++) {
                const uint32_t x = xa[i];
#pragma unroll
                for (int r = 0; r < GKM_LMER_RJ; r++) {
                    const uint32_t v = x ^ y[r];
                    const int mm = __popc((v | (v >> 16)) & 0xFFFFu);
                    if (mm <= d && ok[r]) {
                        const int wgt = WEIGHTED ? (int) sWtA[(size_t) a_l * NK + i] * wy[r] : 1;
                        atomicAdd(h + mm, wgt);
                    }
                }
            }
        }
    }
    __syncthreads();

    for (int i = tid; i < TA * TB; i += GKM_LMER_THREADS) {
        const int a_l = i / TB, b_l = i - a_l * TB;
        const int a_g = row0 + a_l, b_g = col0 + b_l;
        if (a_g >= p.row_end || b_g >= p.col_end) continue;
        if (p.mode == GKM_MODE_LOWER && b_g >= a_g) continue;
        if (p.mode == GKM_MODE_DIAG && b_g != a_g) continue;
        gkm_emit_entry(p, a_g, b_g, sH + (size_t) i * NBN);
    }
}

#endif /* GKM_LMER_KERNEL_CUH_INCLUDED */
