/* How fast can this host first-touch fresh anonymous memory?  (The caller of gkm_main_pywrapper hands over a
 * never-touched np.zeros matrix, gkmsvm.py:75; at 10k sequences the scatter of 400 MB into it is what bounds the
 * end-to-end call.)   gcc -O2 -pthread tools/pagefault_probe.c -o build/pagefault_probe */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static char *base; static size_t stride = 120000, nrows = 10000; static int T, mode; static int cursor;
static char src[120000];
static void *worker(void *arg)
{
    (void) arg;
    for (;;) {
        int k = __atomic_fetch_add(&cursor, 32, __ATOMIC_RELAXED);
        if (k >= (int) nrows) break;
        for (int r = k; r < k + 32 && r < (int) nrows; r++) {
            char *dst = base + (size_t) r * stride; size_t len = (size_t) r * 8;
            if (mode == 0) memcpy(dst, src, len);
            else if (mode == 1) { size_t a = (size_t) dst & ~4095ul, b = ((size_t) dst + len + 4095) & ~4095ul; if (b > a) madvise((void *) a, b - a, MADV_POPULATE_WRITE); }
            else { for (size_t o = 0; o < len; o += 4096) dst[o] = 1; }
        }
    }
    return NULL;
}
int main(void)
{
    memset(src, 1, sizeof src);
    const char *names[] = { "memcpy triangle", "MADV_POPULATE_WRITE triangle", "one byte per page" };
    for (int thp = 0; thp < 2; thp++)
    for (mode = 0; mode < 3; mode++)
    for (T = 1; T <= 16; T *= 2) {
        double best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            size_t bytes = stride * nrows + (4u << 20);
            char *m = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (m == MAP_FAILED) { perror("mmap"); return 1; }
            base = (char *) (((size_t) m + (2u << 20) - 1) & ~((size_t) (2u << 20) - 1));
            if (thp) madvise(base, stride * nrows, MADV_HUGEPAGE);
            cursor = 0;
            pthread_t th[16];
            double t0 = now();
            for (int t = 1; t < T; t++) pthread_create(&th[t], NULL, worker, NULL);
            worker(NULL);
            for (int t = 1; t < T; t++) pthread_join(th[t], NULL);
            double dt = now() - t0;
            if (dt < best) best = dt;
            munmap(m, bytes);
        }
        printf("thp=%d %-30s T=%2d: %.1f ms  (%.1f GB/s of 400 MB)\n", thp, names[mode], T, 1e3 * best, 0.4 / best);
        fflush(stdout);
    }
    return 0;
}
