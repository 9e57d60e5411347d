/* gkm_mma_kernel.cuh -- sm_100a kernel "mma": candidate (b) of the north star.
 *
 * One-hot L-mer GEMM on the 5th-generation tensor cores: every L-mer becomes a row of
 * K = 64 bytes (byte 4t + code(t) = 1 for t < L, zero padding behind 4L), so that
 *     D[i][j] = sum_k A[i][k] B[j][k] = number of MATCHING positions of query L-mer i
 * and target L-mer j; mismatches = L - D.  `tcgen05.mma.kind::i8` (M = 128, N = 256,
 * two K = 32 steps, int32 accumulators in TMEM) is issued by one thread; the operands
 * are built in shared memory straight from the 2-bit planes (never from HBM one-hot
 * data), in the canonical K-major no-swizzle core-matrix layout; the accumulators are
 * read back with `tcgen05.ld.32x32b.x32` and binned in a fused epilogue, so the
 * L-mer x L-mer product never reaches HBM.
 *
 * It exists for the measured comparison with the bit-sliced kernel: every
 * accumulator -- one per L-mer PAIR -- has to be read from TMEM and compared on the CUDA
 * cores, i.e. >= 1 ALU-pipe instruction per pair, while gkm_diag_kernel needs ~0.5.
 * The tensor pipe idles; the epilogue is the bound.  Selected with GKM_KERNEL=mma.
 *
 * CTA = 256 threads = 8 warps; warps w and w+4 share the TMEM lanes 32(w%4).. and split
 * the 256 columns.  One query per CTA (all its L-mers, MT tiles of 128 rows), TB targets.
 */
#ifndef GKM_MMA_KERNEL_CUH_INCLUDED
#define GKM_MMA_KERNEL_CUH_INCLUDED

#include "gkm_diag_kernel.cuh" /* gkm_emit_entry */

#define GKM_MMA_THREADS 256
#define GKM_MMA_M 128
#define GKM_MMA_N 256
#define GKM_MMA_TB 8

__host__ __device__ inline unsigned gkm_mma_smem_bytes(int WA, int nbins, int weighted)
{
    const unsigned MT = (32u * (unsigned) WA + GKM_MMA_M - 1) / GKM_MMA_M;
    unsigned o = 1024;                                 /* barrier, TMEM slot, lengths */
    o += MT * GKM_MMA_M * 64u;                         /* A operand: all query L-mers */
    o += GKM_MMA_N * 64u;                              /* B operand: one tile of target L-mers */
    o += GKM_MMA_TB * (unsigned) nbins * 4u;           /* histograms */
    if (weighted) o += MT * GKM_MMA_M + GKM_MMA_N;     /* weights by L-mer */
    return (o + 127u) & ~127u;
}

__device__ __forceinline__ uint32_t gkm_smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

/* K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
 * core matrix = 8 rows x 16 bytes, rows 16 bytes apart; LBO = distance of the two 16-byte K chunks,
 * SBO = distance of consecutive 8-row groups, both in 16-byte units; version 1 (Blackwell). */
__device__ __forceinline__ uint64_t gkm_umma_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16)
{
    return (uint64_t) ((saddr >> 4) & 0x3FFFu) | ((uint64_t) (lbo16 & 0x3FFFu) << 16) |
           ((uint64_t) (sbo16 & 0x3FFFu) << 32) | ((uint64_t) 1 << 46);
}

/* one operand row (64 bytes = 16 words, word t = one-hot byte of base t) as 4 chunks of 16 bytes:
 * chunk kc of row r lives at base + kc*rows*16 + r*16 */
__device__ __forceinline__ void gkm_mma_store_row(unsigned char *base, int rows, int r, uint32_t p0, uint32_t p1, int L, bool valid)
{
#pragma unroll
    for (int kc = 0; kc < 4; kc++) {
        uint32_t w[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int t = 4 * kc + u;
            const uint32_t code = ((p0 >> t) & 1u) | (((p1 >> t) & 1u) << 1);
            w[u] = (valid && t < L) ? (1u << (8u * code)) : 0u;
        }
        *reinterpret_cast<uint4 *>(base + (size_t) kc * rows * 16 + (size_t) r * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

__device__ __forceinline__ void gkm_lmer_planes(const uint32_t *pl, int W, int o, int L, uint32_t &p0, uint32_t &p1)
{
    const int wi = o >> 5, sh = o & 31;
    const uint32_t mask = (L >= 32) ? 0xFFFFFFFFu : ((1u << L) - 1u);
    const uint32_t n0 = (wi + 1 < W) ? pl[wi + 1] : 0u, n1 = (wi + 1 < W) ? pl[W + wi + 1] : 0u;
    p0 = __funnelshift_r(pl[wi], n0, sh) & mask;
    p1 = __funnelshift_r(pl[W + wi], n1, sh) & mask;
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(GKM_MMA_THREADS)
gkm_mma_kernel(const __grid_constant__ gkm_kparams p)
{
    extern __shared__ __align__(128) unsigned char smem_mma[];
    unsigned char *smem = smem_mma;
    const int W = p.W, L = p.L, d = p.d, NBN = p.nbins;
    const int MT = (32 * p.WA + GKM_MMA_M - 1) / GKM_MMA_M;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 16);
    int *sLenB = reinterpret_cast<int *>(smem + 32);                        /* GKM_MMA_TB ints */
    unsigned char *sAop = smem + 1024;
    unsigned char *sBop = sAop + (size_t) MT * GKM_MMA_M * 64;
    int32_t *sH = reinterpret_cast<int32_t *>(sBop + GKM_MMA_N * 64);
    uint8_t *sWa = reinterpret_cast<uint8_t *>(sH + GKM_MMA_TB * NBN);
    uint8_t *sWb = sWa + (WEIGHTED ? MT * GKM_MMA_M : 0);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int a_g = p.row_begin + (int) blockIdx.y;
    const int col0 = p.col_begin + (int) blockIdx.x * GKM_MMA_TB;
    if (a_g >= p.row_end) return;
    const int col_last = min(col0 + GKM_MMA_TB, p.col_end) - 1;
    if (p.mode == GKM_MODE_LOWER && col0 >= a_g) return;
    if (p.mode == GKM_MODE_DIAG && (col0 > a_g || col_last < a_g)) return;

    /* ---- set-up: TMEM allocation (one warp), barrier, lengths ---- */
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(gkm_smem_u32(tmem_slot)), "n"(GKM_MMA_N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(gkm_smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (tid < GKM_MMA_TB) sLenB[tid] = (col0 + tid < p.col_end) ? p.lens[col0 + tid] : 0;
    for (int i = tid; i < GKM_MMA_TB * NBN; i += GKM_MMA_THREADS) sH[i] = 0;
    const int lenA = p.lens[a_g], nkA = lenA - L + 1;
    /* A operand: query L-mers, forward strand only */
    {
        const uint32_t *pl = p.planes + (size_t) a_g * 3 * W;
        for (int i = tid; i < MT * GKM_MMA_M; i += GKM_MMA_THREADS) {
            const int mt = i / GKM_MMA_M, r = i - mt * GKM_MMA_M;
            uint32_t p0 = 0, p1 = 0;
            const bool valid = i < nkA;
            if (valid) gkm_lmer_planes(pl, W, i, L, p0, p1);
            gkm_mma_store_row(sAop + (size_t) mt * GKM_MMA_M * 64, GKM_MMA_M, r, p0, p1, L, valid);
            if (WEIGHTED) sWa[i] = valid ? p.wend[(size_t) a_g * 32 * W + i + L - 1] : 0;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;"); /* A operand written with generic stores, read by the tensor core */
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;

    /* instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 << 4), A = B = unsigned 8 bit (0),
     * both K-major, N >> 3 at bit 17, M >> 4 at bit 24 */
    const uint32_t idesc = (2u << 4) | ((uint32_t) (GKM_MMA_N >> 3) << 17) | ((uint32_t) (GKM_MMA_M >> 4) << 24);
    const int thr = L - d; /* matches needed for a hit; the launcher guarantees thr >= 1 */
    uint32_t parity = 0;
    /* TMEM lanes of this warp and its half of the columns */
    const int lane_base = 32 * (warp & 3), col_half = (warp >> 2) * (GKM_MMA_N / 2);

    for (int b_l = 0; b_l < GKM_MMA_TB; b_l++) {
        const int b_g = col0 + b_l;
        if (b_g >= p.col_end) break;
        if (p.mode == GKM_MODE_LOWER && b_g >= a_g) break;
        if (p.mode == GKM_MODE_DIAG && b_g != a_g) continue;
        const int lenB = sLenB[b_l], nkB = lenB - L + 1;
        const uint32_t *plb = p.planes + (size_t) b_g * 3 * W;
        const int NT = (2 * nkB + GKM_MMA_N - 1) / GKM_MMA_N;
        for (int nt = 0; nt < NT; nt++) {
            /* B operand: target L-mers jj = 256 nt .. +255 over both strands (jj >= nkB: reverse complement) */
            for (int r = tid; r < GKM_MMA_N; r += GKM_MMA_THREADS) {
                const int jj = GKM_MMA_N * nt + r;
                const bool valid = jj < 2 * nkB;
                const int strand = (jj >= nkB) ? 1 : 0;
                const int o = strand * lenB + (jj - strand * nkB);
                uint32_t p0 = 0, p1 = 0;
                if (valid) gkm_lmer_planes(plb, W, o, L, p0, p1);
                gkm_mma_store_row(sBop, GKM_MMA_N, r, p0, p1, L, valid);
                if (WEIGHTED) sWb[r] = valid ? p.wend[(size_t) b_g * 32 * W + o + L - 1] : 0;
            }
            asm volatile("fence.proxy.async.shared::cta;"); /* generic-proxy stores -> visible to the tensor core */
            __syncthreads();
            for (int mt = 0; mt * GKM_MMA_M < nkA; mt++) {
                if (tid == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
                    for (int ks = 0; ks < 2; ks++) { /* K = 64 bytes = two K = 32 instructions = chunks (2ks, 2ks+1) */
                        const uint64_t adesc = gkm_umma_desc(gkm_smem_u32(sAop + (size_t) mt * GKM_MMA_M * 64 + (size_t) ks * 2 * GKM_MMA_M * 16), GKM_MMA_M, 8);
                        const uint64_t bdesc = gkm_umma_desc(gkm_smem_u32(sBop + (size_t) ks * 2 * GKM_MMA_N * 16), GKM_MMA_N, 8);
                        const uint32_t acc = ks ? 1u : 0u;
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                                     :: "r"(tmem_base), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u));
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(gkm_smem_u32(mbar)));
                }
                /* everyone waits for the accumulators */
                {
                    uint32_t done = 0, spins = 0;
                    while (!done) {
                        if (++spins > (1u << 22)) __trap(); /* never hang the device: a lost commit is a hard error */
                        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                                     : "=r"(done) : "r"(gkm_smem_u32(mbar)), "r"(parity));
                    }
                    parity ^= 1u;
                }
                asm volatile("tcgen05.fence::after_thread_sync;");
                /* fused epilogue: thread = one query L-mer (TMEM lane), 128 target L-mers in 4 loads of 32 columns */
                const int i = mt * GKM_MMA_M + lane_base + lane;
                const int wa = WEIGHTED ? (int) sWa[i] : 1;
#pragma unroll 1
                for (int cc = 0; cc < GKM_MMA_N / 2; cc += 32) {
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t) lane_base << 16) + (uint32_t) (col_half + cc);
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                 : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;");
                    /* cheap reject: the maximum of 8 accumulators against the threshold, then the exact bins */
#pragma unroll
                    for (int g8 = 0; g8 < 32; g8 += 8) {
                        int mx = max(max((int) v[g8], (int) v[g8 + 1]), max((int) v[g8 + 2], (int) v[g8 + 3]));
                        mx = max(mx, max(max((int) v[g8 + 4], (int) v[g8 + 5]), max((int) v[g8 + 6], (int) v[g8 + 7])));
                        if (mx >= thr) {
#pragma unroll
                            for (int u = 0; u < 8; u++) {
                                const int m = (int) v[g8 + u];
                                if (m >= thr) {
                                    const int wgt = WEIGHTED ? wa * (int) sWb[col_half + cc + g8 + u] : 1;
                                    atomicAdd(sH + b_l * NBN + (L - m), wgt);
                                }
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncthreads(); /* TMEM and the B operand may be overwritten now */
            }
        }
    }
    __syncthreads();
    if (tid < GKM_MMA_TB) {
        const int b_g = col0 + tid;
        const bool skip = b_g >= p.col_end || (p.mode == GKM_MODE_LOWER && b_g >= a_g) || (p.mode == GKM_MODE_DIAG && b_g != a_g);
        if (!skip) gkm_emit_entry(p, a_g, b_g, sH + tid * NBN);
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(GKM_MMA_N));
}

#endif /* GKM_MMA_KERNEL_CUH_INCLUDED */
