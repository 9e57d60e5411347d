/* Copy-out of a finished chunk into the caller's fresh (never touched) rows: which way is faster on this box?
 *   (a) product path: D2H into pinned staging, then T host threads memcpy into the pageable rows (first-touch faults)
 *   (b) cudaHostRegister of the chunk's row span + one cudaMemcpy2DAsync straight into it + cudaHostUnregister
 *       (VERDICT r1, item 2 ii: moves the faults into the driver's get_user_pages)
 *   (c) like (a) with MADV_HUGEPAGE on the destination span first
 *   (d) like (a) into a destination that has been touched before (no page faults): what the host's memory system allows
 * nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/hostreg_probe.cu -o build/hostreg_probe -lpthread
 * usage: hostreg_probe [rows=592] [cols=50000] [chunks=8] [threads=16] */
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
struct job { char *dst; const char *src; size_t stride, width; int r0, r1; };
static void *worker(void *a) { job *j = (job *) a; for (int r = j->r0; r < j->r1; r++) memcpy(j->dst + (size_t) r * j->stride, j->src + (size_t) r * j->width, j->width); return NULL; }
static void scatter(char *dst, const char *src, size_t stride, size_t width, int rows, int T)
{
    pthread_t th[64]; job jb[64];
    const int per = (rows + T - 1) / T;
    for (int t = 0; t < T; t++) { jb[t] = { dst, src, stride, width, t * per, (t + 1) * per < rows ? (t + 1) * per : rows }; if (t) pthread_create(&th[t], NULL, worker, &jb[t]); }
    worker(&jb[0]);
    for (int t = 1; t < T; t++) pthread_join(th[t], NULL);
}
int main(int argc, char **argv)
{
    const int rows = argc > 1 ? atoi(argv[1]) : 592, cols = argc > 2 ? atoi(argv[2]) : 50000, chunks = argc > 3 ? atoi(argv[3]) : 8;
    int T = argc > 4 ? atoi(argv[4]) : 16; if (T > 64) T = 64;
    const size_t width = (size_t) cols * 8, stride = width, chunk_bytes = (size_t) rows * stride, total = chunk_bytes * chunks;
    char *d_src, *h_stage;
    CK(cudaMalloc(&d_src, (size_t) rows * width));
    CK(cudaMemset(d_src, 1, (size_t) rows * width));
    CK(cudaMallocHost(&h_stage, (size_t) rows * width));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    for (int mode = 0; mode < 4; mode++) {
        char *m = (char *) mmap(NULL, total + (2u << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        char *base = (char *) (((size_t) m + (2u << 20) - 1) & ~((size_t) (2u << 20) - 1));
        if (mode == 3) memset(base, 1, total); /* fault everything in first */
        double t_reg = 0, t_copy = 0, t_unreg = 0, t0 = now();
        for (int c = 0; c < chunks; c++) {
            char *dst = base + (size_t) c * chunk_bytes;
            if (mode == 1) {
                double a = now();
                CK(cudaHostRegister(dst, chunk_bytes, cudaHostRegisterDefault));
                double b = now();
                CK(cudaMemcpy2DAsync(dst, stride, d_src, width, width, rows, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                double e = now();
                CK(cudaHostUnregister(dst));
                t_reg += b - a; t_copy += e - b; t_unreg += now() - e;
            } else {
                if (mode == 2) madvise(dst, chunk_bytes, MADV_HUGEPAGE);
                double a = now();
                CK(cudaMemcpyAsync(h_stage, d_src, (size_t) rows * width, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                double b = now();
                scatter(dst, h_stage, stride, width, rows, T);
                t_copy += b - a; t_reg += now() - b;
            }
        }
        const double dt = now() - t0;
        if (mode == 1) printf("register + Memcpy2DAsync + unregister: %.1f ms for %.2f GB = %.2f GB/s (register %.1f, copy %.1f, unregister %.1f ms)\n",
                              1e3 * dt, total / 1e9, total / 1e9 / dt, 1e3 * t_reg, 1e3 * t_copy, 1e3 * t_unreg);
        else printf("staging + %d-thread scatter%s: %.1f ms for %.2f GB = %.2f GB/s (D2H %.1f ms, scatter %.1f ms = %.2f GB/s)\n", T, mode == 2 ? " + MADV_HUGEPAGE" : mode == 3 ? " into touched memory" : "",
                    1e3 * dt, total / 1e9, total / 1e9 / dt, 1e3 * t_copy, 1e3 * t_reg, total / 1e9 / t_reg);
        munmap(m, total + (2u << 20));
    }
    return 0;
}
