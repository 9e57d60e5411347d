/* gkmkern_cli.c -- stand-alone `gkmkern` on the GPU engine.
 *
 * Same job and default output as the reference's CLI (src/gkmkern_main.c: `gkmkern pos neg out`,
 * lower triangle as TSV, "%e\t" per entry and "1.0\t" on the diagonal, :221-228), with the
 * parameters the reference hard-codes (:99-107: L=10 k=6 d=3, EST_TRUNC) exposed as options and
 * two of its defects fixed: it silently drops the last N mod 4 rows (:58,:221) and overruns its
 * 10 000-double row buffers for larger problems (:187).
 *
 *   gkmkern [-t type] [-l L] [-k k] [-d d] [-M M] [-H H] [-g gamma] [-T threads] [-v verbosity]
 *           [-p digits | -b] posfile negfile outfile
 *     -p digits   print with %.<digits>g instead of %e (17 round-trips a double)
 *     -b          binary output: int32 n, then rows 0..n-1, row a = a doubles K(a,0..a-1)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/gkm_b200.h"

int main(int argc, char **argv)
{
    gkm_parameter param;
    memset(&param, 0, sizeof(param));
    param.kernel_type = EST_TRUNC; param.L = 10; param.k = 6; param.d = 3; /* gkmkern_main.c:99-107 */
    param.M = 50; param.H = 50; param.gamma = 1.0; param.nthreads = 1;
    int threads = 4, verbosity = 2, digits = 0, binary = 0, c;
    while ((c = getopt(argc, argv, "t:l:k:d:M:H:g:T:v:p:b")) != -1) {
        switch (c) {
            case 't': param.kernel_type = atoi(optarg); break;
            case 'l': param.L = atoi(optarg); break;
            case 'k': param.k = atoi(optarg); break;
            case 'd': param.d = atoi(optarg); break;
            case 'M': param.M = (u_int8_t) atoi(optarg); break;
            case 'H': param.H = atof(optarg); break;
            case 'g': param.gamma = atof(optarg); break;
            case 'T': threads = atoi(optarg); break;
            case 'v': verbosity = atoi(optarg); break;
            case 'p': digits = atoi(optarg); break;
            case 'b': binary = 1; break;
            default: argc = 0; break;
        }
    }
    if (argc - optind != 3) {
        fprintf(stderr, "usage: gkmkern [-t type] [-l L] [-k k] [-d d] [-M M] [-H H] [-g gamma] [-T threads] "
                        "[-v verbosity] [-p digits | -b] posfile negfile outfile\n");
        return 1;
    }
    const char *posfile = argv[optind], *negfile = argv[optind + 1], *outfile = argv[optind + 2];
    gkmb200_set_verbosity(verbosity);
    const char *bad = gkmb200_check_parameter(&param);
    if (bad && !(strcmp(bad, "L > 12") == 0 && param.L <= 16)) { fprintf(stderr, "gkmkern: %s\n", bad); return 1; }

    gkmb200_problem *p = gkmb200_problem_new(&param);
    if (!p) { fprintf(stderr, "gkmkern: %s\n", gkmb200_last_error()); return 1; }
    if (gkmb200_problem_read(p, posfile, negfile) < 0) { fprintf(stderr, "gkmkern: %s\n", gkmb200_last_error()); return 1; }
    const int n = gkmb200_problem_size(p);
    FILE *fo = fopen(outfile, binary ? "wb" : "w");
    if (!fo) { perror("error occurred while opening a file"); return 1; }
    if (binary) { int32_t n32 = n; fwrite(&n32, sizeof(n32), 1, fo); }
    char fmt[16];
    if (digits > 0) snprintf(fmt, sizeof(fmt), "%%.%dg\t", digits); else snprintf(fmt, sizeof(fmt), "%%e\t");

    /* bands of rows keep the host buffer small whatever n is */
    int band = 1024;
    if ((long long) band * n * 8 > (512LL << 20)) band = (int) ((512LL << 20) / ((long long) n * 8));
    if (band < 1) band = 1;
    double *buf = (double *) malloc(sizeof(double) * (size_t) band * (size_t) n);
    if (!buf) { fprintf(stderr, "gkmkern: out of memory\n"); return 1; }
    (void) threads;
    for (int r0 = 0; r0 < n; r0 += band) {
        const int r1 = (r0 + band < n) ? r0 + band : n;
        /* always the full column range: the index blocks partition the columns of the first call, and a range that
         * grew from band to band made every band rebuild all of them */
        if (gkmb200_kernel_block(p, r0, r1 - r0, 0, n, 1, buf, n)) {
            fprintf(stderr, "gkmkern: %s\n", gkmb200_last_error());
            return 1;
        }
        for (int a = r0; a < r1; a++) {
            const double *row = buf + (size_t) (a - r0) * (size_t) n;
            if (binary) {
                fwrite(row, sizeof(double), (size_t) a, fo);
            } else {
                for (int j = 0; j < a; j++) fprintf(fo, fmt, row[j]);
                fprintf(fo, "1.0\t\n");
            }
        }
    }
    free(buf);
    fclose(fo);
    gkmb200_problem_free(p);
    return 0;
}
