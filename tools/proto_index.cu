// proto_index.cu -- feasibility prototype of candidate (c): inverted L-mer index + neighbour enumeration.
// One CTA per query row a; its histogram row H[m][b] lives in shared memory; every forward L-mer x of a is
// XOR-ed with every delta of Hamming weight <= d (base-4), the target postings of y = x ^ delta are
// fetched from an L2-resident direct-addressed table and scattered into the row with shared atomics.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o build/proto_index tools/proto_index.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

static const int L = 11, D = 3, NB = 4, LEN = 300;
static const int NQ = LEN - L + 1;

template <int UNROLL>
__global__ void __launch_bounds__(1024, 1)
row_kernel(const uint32_t *__restrict__ qlmers, // [n][NQ]
           const uint32_t *__restrict__ deltas, int ndelta, // delta | m << 28
           const uint32_t *__restrict__ idx,    // [4^L + 1]
           const uint32_t *__restrict__ post,   // b | wt << 24, sorted by b within an L-mer
           int n, int row0, int32_t *__restrict__ Hout /* [rows][NB][n] */, int triangle)
{
    extern __shared__ int32_t H[]; // [NB][n]
    const int a = row0 + blockIdx.x;
    const int nthr = blockDim.x;
    for (int i = threadIdx.x; i < NB * n; i += nthr) H[i] = 0;
    __syncthreads();
    const int bmax = triangle ? a : n;
    const uint32_t *xq = qlmers + (size_t) a * NQ;
    // thread keeps one delta, walks the query L-mers UNROLL at a time (independent lookups in flight)
    for (int t = threadIdx.x; t < ndelta; t += nthr) {
        const uint32_t dl = deltas[t];
        const uint32_t dx = dl & 0x0FFFFFFFu;
        int32_t *Hm = H + (dl >> 28) * n;
        for (int xi = 0; xi < NQ; xi += UNROLL) {
            uint32_t lo[UNROLL], hi[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                if (xi + u < NQ) {
                    const uint32_t y = xq[xi + u] ^ dx;
                    lo[u] = idx[y]; hi[u] = idx[y + 1];
                } else { lo[u] = hi[u] = 0; }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                for (uint32_t p = lo[u]; p < hi[u]; p++) {
                    const uint32_t e = post[p];
                    const int b = (int) (e & 0xFFFFFFu);
                    if (b >= bmax) break;
                    atomicAdd(&Hm[b], (int) (e >> 24));
                }
            }
        }
    }
    __syncthreads();
    int32_t *out = Hout + (size_t) blockIdx.x * NB * n;
    for (int i = threadIdx.x; i < NB * n; i += nthr) out[i] = H[i];
}


// v2: 8-byte slots {first posting, overflow offset}; overflow lists end with a 0xFFFFFFFF sentinel; deltas ordered so
// that the 4 neighbours that differ only in the last base sit in adjacent lanes (one 32-byte sector)
template <int UNROLL>
__global__ void __launch_bounds__(1024, 1)
row_kernel2(const uint32_t *__restrict__ qlmers, const uint32_t *__restrict__ deltas, int ndelta,
            const uint2 *__restrict__ tab, const uint32_t *__restrict__ ovf,
            int n, int row0, int32_t *__restrict__ Hout, int triangle)
{
    extern __shared__ int32_t H[]; // [NB][n]
    const int a = row0 + blockIdx.x;
    const int nthr = blockDim.x;
    for (int i = threadIdx.x; i < NB * n; i += nthr) H[i] = 0;
    __syncthreads();
    const uint32_t bmax = triangle ? a : n;
    const uint32_t *xq = qlmers + (size_t) a * NQ;
    for (int t = threadIdx.x; t < ndelta; t += nthr) {
        const uint32_t dl = deltas[t];
        const uint32_t dx = dl & 0x0FFFFFFFu;
        int32_t *Hm = H + (dl >> 28) * n;
        for (int xi = 0; xi < NQ; xi += UNROLL) {
            uint2 sl[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                if (xi + u < NQ) sl[u] = tab[xq[xi + u] ^ dx]; else sl[u] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                uint32_t e = sl[u].x;
                uint32_t b = e & 0xFFFFFFu;
                if (b < bmax) {
                    atomicAdd(&Hm[b], (int) (e >> 24));
                    uint32_t p = sl[u].y;
                    if (p != 0xFFFFFFFFu) {
                        while (true) {
                            e = ovf[p++];
                            b = e & 0xFFFFFFu;
                            if (b >= bmax) break;
                            atomicAdd(&Hm[b], (int) (e >> 24));
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    int32_t *out = Hout + (size_t) blockIdx.x * NB * n;
    for (int i = threadIdx.x; i < NB * n; i += nthr) out[i] = H[i];
}

// smem atomic throughput: random addresses
__global__ void atoms_bench(int iters, int nbins, int *sink)
{
    extern __shared__ int32_t H[];
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) H[i] = 0;
    __syncthreads();
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x * 97u + 12345u;
    for (int i = 0; i < iters; i++) {
        s = s * 1664525u + 1013904223u;
        atomicAdd(&H[(s >> 8) & (uint32_t) (nbins - 1)], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) sink[blockIdx.x] = H[0];
}

static uint32_t rng_state = 12345u;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 17; rng_state ^= rng_state << 5; return rng_state; }

int main(int argc, char **argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 10000;
    int rows = argc > 2 ? atoi(argv[2]) : 1480;
    const int only_v2 = argc > 3 ? atoi(argv[3]) : 0;
    if (rows > n) rows = n;
    std::vector<uint8_t> seq((size_t) n * LEN);
    for (auto &c : seq) c = rnd() & 3;
    const uint32_t mask = (1u << (2 * L)) - 1;
    const size_t NL = (size_t) 1 << (2 * L);
    std::vector<uint32_t> q((size_t) n * NQ), trg((size_t) n * 2 * NQ);
    for (int s = 0; s < n; s++) {
        const uint8_t *p = &seq[(size_t) s * LEN];
        uint32_t f = 0, r = 0;
        for (int i = 0; i < LEN; i++) {
            f = ((f << 2) | p[i]) & mask;
            r = (r >> 2) | ((uint32_t) (3 - p[i]) << (2 * (L - 1)));
            if (i >= L - 1) {
                q[(size_t) s * NQ + i - (L - 1)] = f;
                trg[(size_t) s * 2 * NQ + i - (L - 1)] = f;
                trg[(size_t) s * 2 * NQ + NQ + i - (L - 1)] = r;
            }
        }
    }
    // index: counting sort by L-mer, b ascending inside
    std::vector<uint32_t> idx(NL + 1, 0), post((size_t) n * 2 * NQ);
    for (uint32_t y : trg) idx[y + 1]++;
    for (size_t i = 0; i < NL; i++) idx[i + 1] += idx[i];
    {
        std::vector<uint32_t> cur(idx.begin(), idx.end() - 1);
        for (int s = 0; s < n; s++)
            for (int i = 0; i < 2 * NQ; i++) post[cur[trg[(size_t) s * 2 * NQ + i]]++] = (uint32_t) s | (1u << 24);
    }
    // deltas
    std::vector<uint32_t> deltas;
    for (int m = 0; m <= D; m++) {
        // choose m positions, 3^m substitutions
        std::vector<int> pos(m);
        std::vector<int> st(m);
        // iterate combinations
        std::vector<int> c(m);
        for (int i = 0; i < m; i++) c[i] = i;
        while (true) {
            int nsub = 1; for (int i = 0; i < m; i++) nsub *= 3;
            for (int sidx = 0; sidx < nsub; sidx++) {
                uint32_t dlt = 0; int v = sidx;
                for (int i = 0; i < m; i++) { dlt |= (uint32_t) (1 + v % 3) << (2 * c[i]); v /= 3; }
                deltas.push_back(dlt | ((uint32_t) m << 28));
            }
            int i = m - 1;
            while (i >= 0 && c[i] == L - m + i) i--;
            if (i < 0) break;
            c[i]++;
            for (int j = i + 1; j < m; j++) c[j] = c[j - 1] + 1;
        }
    }
    const int nd = (int) deltas.size();
    printf("n=%d rows=%d ndelta=%d postings=%zu\n", n, rows, nd, post.size());

    uint32_t *dq, *dd, *didx, *dpost; int32_t *dH;
    CK(cudaMalloc(&dq, q.size() * 4)); CK(cudaMalloc(&dd, nd * 4)); CK(cudaMalloc(&didx, idx.size() * 4));
    CK(cudaMalloc(&dpost, post.size() * 4)); CK(cudaMalloc(&dH, (size_t) rows * NB * n * 4));
    CK(cudaMemcpy(dq, q.data(), q.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dd, deltas.data(), nd * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(didx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dpost, post.data(), post.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = (size_t) NB * n * 4;
    CK(cudaFuncSetAttribute(row_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    CK(cudaFuncSetAttribute(row_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    CK(cudaFuncSetAttribute(row_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int row0 = n - rows; // the most expensive rows of the triangle
    for (int tri = 0; tri < 2 && !only_v2; tri++)
        for (int un = 1; un <= 4; un *= 2)
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaEventRecord(e0));
                if (un == 1) row_kernel<1><<<rows, 1024, smem>>>(dq, dd, nd, didx, dpost, n, row0, dH, tri);
                if (un == 2) row_kernel<2><<<rows, 1024, smem>>>(dq, dd, nd, didx, dpost, n, row0, dH, tri);
                if (un == 4) row_kernel<4><<<rows, 1024, smem>>>(dq, dd, nd, didx, dpost, n, row0, dH, tri);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                double entries = tri ? ((double) rows * (row0 + row0 + rows - 1) / 2) : (double) rows * n;
                printf("tri=%d unroll=%d rep=%d: %.3f ms for %d rows -> %.1f M entries/s (%.1f ms per 10k rows-equivalent)\n",
                       tri, un, rep, ms, rows, entries / ms / 1e3, ms * 10000.0 / rows);
            }

    // ---- v2 structures ----
    std::vector<uint2> tab(NL);
    std::vector<uint32_t> ovf;
    for (size_t y = 0; y < NL; y++) {
        uint32_t c = idx[y + 1] - idx[y];
        if (c == 0) { tab[y] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu); continue; }
        tab[y].x = post[idx[y]];
        if (c == 1) { tab[y].y = 0xFFFFFFFFu; continue; }
        tab[y].y = (uint32_t) ovf.size();
        for (uint32_t i = 1; i < c; i++) ovf.push_back(post[idx[y] + i]);
        ovf.push_back(0xFFFFFFFFu);
    }
    std::vector<uint32_t> d2, singles;
    for (uint32_t dl : deltas) {
        uint32_t dx = dl & 0x0FFFFFFFu, m = dl >> 28;
        if (dx & 3u) continue;          // enumerate upper deltas only
        if ((int) m < D) for (uint32_t last = 0; last < 4; last++) d2.push_back((dx | last) | ((m + (last != 0)) << 28));
        else singles.push_back(dl);
    }
    printf("v2: %zu grouped + %zu single deltas, overflow entries %zu\n", d2.size(), singles.size(), ovf.size());
    d2.insert(d2.end(), singles.begin(), singles.end());
    uint2 *dtab; uint32_t *dovf, *dd2;
    CK(cudaMalloc(&dtab, tab.size() * 8)); CK(cudaMalloc(&dovf, ovf.size() * 4 + 4)); CK(cudaMalloc(&dd2, d2.size() * 4));
    CK(cudaMemcpy(dtab, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dovf, ovf.data(), ovf.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dd2, d2.data(), d2.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(row_kernel2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    CK(cudaFuncSetAttribute(row_kernel2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    CK(cudaFuncSetAttribute(row_kernel2<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    for (int tri = only_v2 ? 1 : 0; tri < 2; tri++)
        for (int un = only_v2 ? 4 : 1; un <= 4; un *= 2)
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaEventRecord(e0));
                if (un == 1) row_kernel2<1><<<rows, 1024, smem>>>(dq, dd2, (int) d2.size(), dtab, dovf, n, row0, dH, tri);
                if (un == 2) row_kernel2<2><<<rows, 1024, smem>>>(dq, dd2, (int) d2.size(), dtab, dovf, n, row0, dH, tri);
                if (un == 4) row_kernel2<4><<<rows, 1024, smem>>>(dq, dd2, (int) d2.size(), dtab, dovf, n, row0, dH, tri);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                double entries = tri ? ((double) rows * (row0 + row0 + rows - 1) / 2) : (double) rows * n;
                printf("v2 tri=%d unroll=%d rep=%d: %.3f ms for %d rows -> %.1f M entries/s (%.1f ms per 10k rows-equivalent)\n",
                       tri, un, rep, ms, rows, entries / ms / 1e3, ms * 10000.0 / rows);
            }
    // verify a few entries against brute force (tri=1 was last: b < a only)
    std::vector<int32_t> Hh((size_t) NB * n);
    int bad = 0;
    for (int r = 0; r < 3; r++) {
        int blk = r * (rows - 1) / 2, a = row0 + blk;
        CK(cudaMemcpy(Hh.data(), dH + (size_t) blk * NB * n, Hh.size() * 4, cudaMemcpyDeviceToHost));
        for (int bi = 0; bi < 40; bi++) {
            int b = (int) (rnd() % (uint32_t) a);
            int ref[NB] = {0, 0, 0, 0};
            for (int i = 0; i < NQ; i++)
                for (int j = 0; j < 2 * NQ; j++) {
                    uint32_t x = q[(size_t) a * NQ + i] ^ trg[(size_t) b * 2 * NQ + j];
                    x = (x | (x >> 1)) & 0x55555555u;
                    int m = __builtin_popcount(x);
                    if (m <= D) ref[m]++;
                }
            for (int m = 0; m < NB; m++) if (ref[m] != Hh[(size_t) m * n + b]) { bad++; if (bad < 5) printf("MISMATCH a=%d b=%d m=%d ref=%d got=%d\n", a, b, m, ref[m], Hh[(size_t) m * n + b]); }
        }
    }
    printf("verify: %s\n", bad ? "FAILED" : "ok (120 entries bit-exact vs brute force)");
    // smem atomics micro-benchmark
    int *sink; CK(cudaMalloc(&sink, 4096 * 4));
    CK(cudaFuncSetAttribute(atoms_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 160000));
    for (int nb = 1024; nb <= 32768; nb *= 32) {
        atoms_bench<<<148, 1024, 160000>>>(10, nb, sink);
        CK(cudaEventRecord(e0));
        atoms_bench<<<148, 1024, 160000>>>(2000, nb, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("ATOMS random over %d bins: %.2f G atomics/s (%.3f per clk per SM @1.965GHz)\n", nb,
               148.0 * 1024 * 2000 / ms / 1e6, 148.0 * 1024 * 2000 / (ms * 1e-3) / 148 / 1.965e9);
    }
    return bad != 0;
}
