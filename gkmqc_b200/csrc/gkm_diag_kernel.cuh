/* gkm_diag_kernel.cuh -- sm_100a kernel "diag": bit-sliced diagonals (see gkm_bitslice.h).
 *
 * Replaces kmertree_dfs + gkmkernel_kernelfunc_batch_single + the normalisation of
 * gkmkernel_kernelfunc_batch_all (libgkm.c:315-387, :553-589, :1156-1185) for a
 * TA x TB tile of (query, target) sequence pairs per CTA.
 *
 * CTA = 256 threads.  Shared memory holds, for the tile,
 *   sA    [TA][32W]  gkm_apos   query records (broadcast-read, one LDS.128 per step)
 *   sS    [TB][2 strands][2 planes][W] target bit planes, sE [TB][W] valid-window-end plane
 *   sW    [TB][2][32W] bytes    target weights by window end (weighted types)
 *   sTask [<= TB*2*W]           lane tasks (target, strand, block of 32 diagonals)
 *   sH    [TA][TB][NB] int32    the tile's histograms
 * A warp takes (query, group of 32 tasks) combos round-robin; every lane runs
 * gkm_diag_lane over the whole query, then adds its NB counters into sH with shared
 * atomics (a few per ~10^4 instructions).  The epilogue turns each histogram into
 * the normalised double exactly like the reference does: ascending-m sum from 0.0,
 * one division by sqnorm_a*sqnorm_b, no FMA contraction (libgkm.c:576-582,:1169-1179).
 * Nothing but the 8-byte results (or raw histograms when asked) reaches HBM.
 */
#ifndef GKM_DIAG_KERNEL_CUH_INCLUDED
#define GKM_DIAG_KERNEL_CUH_INCLUDED

#include "gkm_bitslice.h"
#include "gkm_kparams.h"

#define GKM_DIAG_THREADS 256

struct gkm_diag_layout_t {
    unsigned offA, offS, offE, offZ, offW, offTask, offH, offLenA, offLenB, offPre, total;
};

__host__ __device__ inline gkm_diag_layout_t gkm_diag_layout(int W, int TA, int TB, int NB, int weighted)
{
    gkm_diag_layout_t l;
    unsigned o = 0;
    l.offA = o;    o += (unsigned) TA * 32u * (unsigned) W * 16u;
    l.offS = o;    o += (unsigned) TB * 4u * (unsigned) W * 4u;
    l.offE = o;    o += (unsigned) TB * (unsigned) W * 4u;
    l.offZ = o;    o += (unsigned) W * 4u;
    l.offW = o;    o += weighted ? (unsigned) TB * 2u * 32u * (unsigned) W : 0u;
    l.offTask = o; o += (unsigned) TB * 2u * (unsigned) W * 4u;
    l.offH = o;    o += (unsigned) TA * (unsigned) TB * (unsigned) NB * 4u;
    l.offLenA = o; o += (unsigned) TA * 4u;
    l.offLenB = o; o += (unsigned) TB * 4u;
    l.offPre = o;  o += ((unsigned) TB + 1u) * 4u;
    l.total = (o + 15u) & ~15u;
    return l;
}

/* histogram -> kernel value; shared by all kernel variants */
__device__ __forceinline__ double gkm_kraw(const gkm_kparams &p, const int32_t *H)
{
    double sum = 0.0;
    for (int m = 0; m < p.nbins; m++) sum = __dadd_rn(sum, __dmul_rn(p.w[m], (double) H[m]));
    return sum;
}

__device__ __forceinline__ void gkm_emit_entry(const gkm_kparams &p, int a_g, int b_g, const int32_t *H)
{
    if (p.hist) {
        int32_t *dst = p.hist + ((size_t) (a_g - p.row_base) * (size_t) p.hist_cols + (size_t) (b_g - p.col_base)) * (size_t) p.nbins;
        for (int m = 0; m < p.nbins; m++) dst[m] = H[m];
    }
    const double kraw = gkm_kraw(p, H);
    if (p.mode == GKM_MODE_DIAG) {
        if (p.sqnorm_out) p.sqnorm_out[a_g] = __dsqrt_rn(kraw);
        return;
    }
    if (!p.out && !p.decision) return;
    double v = __ddiv_rn(kraw, __dmul_rn(p.sqnorm[a_g], p.sqnorm[b_g]));
    if (p.kernel_type == 3 || p.kernel_type == 5) v = exp(__dmul_rn(p.gamma, __dadd_rn(v, -1.0)));
    if (p.out) p.out[(size_t) (a_g - p.row_base) * (size_t) p.ld + (size_t) (b_g - p.col_base)] = v;
    if (p.decision) atomicAdd(p.decision + (a_g - p.row_base), p.alpha[b_g - p.col_base] * v);
}

template <int L, int NB, bool WEIGHTED>
__global__ void __launch_bounds__(GKM_DIAG_THREADS)
gkm_diag_kernel(const __grid_constant__ gkm_kparams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int W = p.W, TA = p.TA, TB = p.TB;
    const gkm_diag_layout_t lay = gkm_diag_layout(W, TA, TB, NB, WEIGHTED ? 1 : 0);
    gkm_apos *sA = reinterpret_cast<gkm_apos *>(smem + lay.offA);
    uint32_t *sS = reinterpret_cast<uint32_t *>(smem + lay.offS);
    uint32_t *sE = reinterpret_cast<uint32_t *>(smem + lay.offE);
    uint32_t *sZ = reinterpret_cast<uint32_t *>(smem + lay.offZ);
    uint8_t *sW = smem + lay.offW;
    uint32_t *sTask = reinterpret_cast<uint32_t *>(smem + lay.offTask);
    int32_t *sH = reinterpret_cast<int32_t *>(smem + lay.offH);
    int *sLenA = reinterpret_cast<int *>(smem + lay.offLenA);
    int *sLenB = reinterpret_cast<int *>(smem + lay.offLenB);
    int *sPre = reinterpret_cast<int *>(smem + lay.offPre);

    const int tid = threadIdx.x;
    const int row0 = p.row_begin + (int) blockIdx.y * TA;
    const int col0 = p.col_begin + (int) blockIdx.x * TB;
    const int row_last = min(row0 + TA, p.row_end) - 1;
    const int col_last = min(col0 + TB, p.col_end) - 1;
    if (p.mode == GKM_MODE_LOWER && col0 > row_last) return;           /* tile entirely above the diagonal */
    if (p.mode == GKM_MODE_DIAG && (col0 > row_last || col_last < row0)) return;

    /* ---- stage the tile ---- */
    if (tid < TA) sLenA[tid] = (row0 + tid < p.row_end) ? p.lens[row0 + tid] : 0;
    if (tid < TB) sLenB[tid] = (col0 + tid < p.col_end) ? p.lens[col0 + tid] : 0;
    for (int i = tid; i < W; i += GKM_DIAG_THREADS) sZ[i] = 0u;
    for (int i = tid; i < TA * TB * NB; i += GKM_DIAG_THREADS) sH[i] = 0;
    for (int i = tid; i < TB * 4 * W; i += GKM_DIAG_THREADS) {
        const int b = i / (4 * W);
        sS[i] = (col0 + b < p.col_end) ? p.planes[(size_t) (col0 + b) * 4 * W + (i - b * 4 * W)] : 0u;
    }
    if (WEIGHTED) {
        const int wordsPerB = 16 * W; /* 2 strands * 32W bytes */
        const uint32_t *src = reinterpret_cast<const uint32_t *>(p.wend);
        uint32_t *dst = reinterpret_cast<uint32_t *>(sW);
        for (int i = tid; i < TB * wordsPerB; i += GKM_DIAG_THREADS) {
            const int b = i / wordsPerB;
            dst[i] = (col0 + b < p.col_end) ? src[(size_t) (col0 + b) * wordsPerB + (i - b * wordsPerB)] : 0u;
        }
    }
    __syncthreads();
    /* valid-window-end plane of each target: bits L-1 .. len-1 */
    for (int i = tid; i < TB * W; i += GKM_DIAG_THREADS) {
        const int b = i / W, wi = i - b * W;
        const int lo = max(L - 1, 32 * wi) - 32 * wi, hi = min(sLenB[b], 32 * wi + 32) - 32 * wi;
        uint32_t m = 0u;
        if (hi > lo) m = ((hi >= 32) ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
        sE[i] = m;
    }
    /* query records */
    for (int i = tid; i < TA * 32 * W; i += GKM_DIAG_THREADS) {
        const int a = i / (32 * W), e = i - a * 32 * W;
        gkm_apos r;
        r.a0 = 0u; r.a1 = 0u; r.va = 0u; r.wa = 1u;
        if (row0 + a < p.row_end) {
            const uint32_t *pl = p.planes + (size_t) (row0 + a) * 4 * W;
            r.a0 = ((pl[e >> 5] >> (e & 31)) & 1u) ? 0xFFFFFFFFu : 0u;
            r.a1 = ((pl[W + (e >> 5)] >> (e & 31)) & 1u) ? 0xFFFFFFFFu : 0u;
            r.va = (e >= L - 1 && e < sLenA[a]) ? 0xFFFFFFFFu : 0u;
            if (WEIGHTED) r.wa = p.wend[(size_t) (row0 + a) * 64 * W + e];
        }
        sA[i] = r;
    }
    /* lane tasks: target b contributes 2 strands x ceil(len_b/32) blocks of 32 diagonals */
    if (tid == 0) {
        int acc = 0;
        for (int b = 0; b < TB; b++) { sPre[b] = acc; acc += 2 * ((sLenB[b] + 31) >> 5); }
        sPre[TB] = acc;
    }
    __syncthreads();
    const int ntasks = sPre[TB];
    for (int t = tid; t < ntasks; t += GKM_DIAG_THREADS) {
        int b = 0;
        while (sPre[b + 1] <= t) b++;
        const int wb = (sLenB[b] + 31) >> 5;
        const int r = t - sPre[b];
        const int strand = r / wb, q = r - strand * wb;
        sTask[t] = ((uint32_t) b << 16) | ((uint32_t) strand << 15) | (uint32_t) q;
    }
    __syncthreads();

    /* ---- main loop: (query, task group) combos ---- */
    const int lane = tid & 31, warp = tid >> 5;
    const int ngroups = (ntasks + 31) >> 5;
    const int ncombos = TA * ngroups;
    for (int combo = warp; combo < ncombos; combo += GKM_DIAG_THREADS / 32) {
        const int a_l = combo / ngroups, g = combo - a_l * ngroups;
        const int a_g = row0 + a_l;
        if (a_g >= p.row_end) break;
        const int t = 32 * g + lane;
        bool active = t < ntasks;
        const uint32_t task = active ? sTask[t] : 0u;
        const int b_l = (int) (task >> 16), strand = (int) ((task >> 15) & 1u), q = (int) (task & 0x7FFFu);
        const int b_g = col0 + b_l;
        if (p.mode == GKM_MODE_LOWER) active = active && (b_g < a_g);
        if (p.mode == GKM_MODE_DIAG) active = active && (b_g == a_g);
        if (!__any_sync(0xFFFFFFFFu, active)) continue;

        const uint32_t *S0 = sS + ((b_l * 2 + strand) * 2 + 0) * W;
        const uint32_t *S1 = sS + ((b_l * 2 + strand) * 2 + 1) * W;
        const uint32_t *E = active ? (sE + b_l * W) : sZ;
        const int Wb = active ? ((sLenB[b_l] + 31) >> 5) : 1;
        const uint8_t *wendp = WEIGHTED ? (sW + (size_t) (b_l * 2 + strand) * 32 * W) : nullptr;

        int32_t acc[NB];
#pragma unroll
        for (int m = 0; m < NB; m++) acc[m] = 0;
        gkm_diag_lane<L, NB, WEIGHTED>(sA + (size_t) a_l * 32 * W, sLenA[a_l], S0, S1, E, Wb, active ? q : 0, wendp, acc);
        if (active) {
            int32_t *h = sH + (a_l * TB + b_l) * NB;
#pragma unroll
            for (int m = 0; m < NB; m++)
                if (acc[m]) atomicAdd(h + m, acc[m]);
        }
    }
    __syncthreads();

    /* ---- epilogue: histogram -> normalised double ---- */
    for (int i = tid; i < TA * TB; i += GKM_DIAG_THREADS) {
        const int a_l = i / TB, b_l = i - a_l * TB;
        const int a_g = row0 + a_l, b_g = col0 + b_l;
        if (a_g >= p.row_end || b_g >= p.col_end) continue;
        if (p.mode == GKM_MODE_LOWER && b_g >= a_g) continue;
        if (p.mode == GKM_MODE_DIAG && b_g != a_g) continue;
        gkm_emit_entry(p, a_g, b_g, sH + (a_l * TB + b_l) * NB);
    }
}

#endif /* GKM_DIAG_KERNEL_CUH_INCLUDED */
