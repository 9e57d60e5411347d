/* oracle/ref_hook.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin probe around the UNMODIFIED reference engine.  The reference sources are
 * compiled from where they lie under /root/reference/src (see oracle/Makefile);
 * this file textually includes libgkm.c so that the probe can reach the
 * file-static DFS (libgkm.c:315) and read the integer mismatch profile that the
 * reference allocates and frees inside gkmkernel_kernelfunc_batch_single
 * (libgkm.c:568-588) and never exports.  Nothing here re-implements reference
 * arithmetic: every number returned is produced by reference code.
 *
 * Built twice by oracle/Makefile:
 *   _ref/gkmref_hook.so      -- stock limits (MAX_MM 12)
 *   _ref/gkmref_hook_L16.so  -- header copy with MAX_MM raised to 16 so that the
 *                               weight table (libgkm.c:190 writes index 0..L) does
 *                               not overrun for L = 13..16; the L<=12 gate lives in
 *                               gkmkern_pylib.c:54 which this probe does not go through.
 */
#define CLOG_MAIN
#include "libgkm.c" /* found through -I<reference>/src */

#include <pthread.h>
#include <time.h>

static gkm_parameter g_param;
static gkm_kernel *g_kernel = NULL;
static svm_problem g_prob;
static int g_npos = 0;

extern "C" {

/* open a problem exactly the way gkm_main_pywrapper does (gkmkern_pylib.c:109-166),
 * minus the parameter gate.  Returns the number of sequences, or -1. */
int gkmref_open(int kernel_type, int L, int k, int d, int M, double H, double gamma,
                const char *posfile, const char *negfile)
{
    if (g_kernel) return -1;
    if (_clog_loggers[LOGGER_ID] == NULL) {
        if (clog_init_fd(LOGGER_ID, 1) != 0) return -1;
        clog_set_fmt(LOGGER_ID, LOGGER_FORMAT);
        clog_set_level(LOGGER_ID, CLOG_ERROR);
    }
    g_param.kernel_type = kernel_type;
    g_param.L = L;
    g_param.k = k;
    g_param.d = d;
    g_param.M = (u_int8_t) M;
    g_param.H = H;
    g_param.gamma = gamma;
    g_param.nthreads = 1;
    g_kernel = gkmkernel_init(&g_param);
    g_npos = gkmkernel_read_problems(g_kernel, &g_prob, posfile, negfile);
    gkmkernel_build_tree(g_kernel, g_prob.x, g_prob.l);
    return g_prob.l;
}

int gkmref_npos(void) { return g_npos; }

void gkmref_weights(double *out, int n)
{
    for (int i = 0; i < n; i++) out[i] = g_kernel->weights[i];
}

/* w[m] alone, without allocating a tree: the reference's own static weight
 * routines behind the dispatch of libgkm.c:997-1019 */
void gkmref_weights_only(int kernel_type, int L, int k, double *out, int n)
{
    gkm_parameter p;
    gkm_kernel kern;
    memset(&kern, 0, sizeof(kern));
    p.kernel_type = kernel_type; p.L = L; p.k = k; p.d = 0; p.M = 50; p.H = 50; p.gamma = 1; p.nthreads = 1;
    kern.param = &p;
    if (kernel_type == GKM) calc_gkm_kernel_wt(&kern);
    else if (kernel_type == EST_FULL) calc_gkm_kernel_lmerest_wt(&kern, 0);
    else calc_gkm_kernel_lmerest_wt(&kern, 1);
    for (int i = 0; i < n; i++) out[i] = kern.weights[i];
}

double gkmref_sqnorm(int i) { return g_prob.x[i]->sqnorm; }
int gkmref_seqlen(int i) { return g_prob.x[i]->seqlen; }
/* record id and spelling as the reference's reader stored them (libgkm.c:849-860,1287-1292) */
const char *gkmref_sid(int i) { return g_prob.x[i]->sid; }
const char *gkmref_seq_string(int i) { return g_prob.x[i]->seq_string; }

void gkmref_poswt(int i, unsigned char *wt, unsigned char *wt_rc)
{
    int n = g_prob.x[i]->seqlen - g_param.L + 1;
    for (int j = 0; j < n; j++) { wt[j] = g_prob.x[i]->wt[j]; wt_rc[j] = g_prob.x[i]->wt_rc[j]; }
}

void gkmref_codes(int i, unsigned char *fwd, unsigned char *rc)
{
    for (int j = 0; j < g_prob.x[i]->seqlen; j++) { fwd[j] = g_prob.x[i]->seq[j]; rc[j] = g_prob.x[i]->seq_rc[j]; }
}

/* integer mismatch profile of row a against ids [0,end): out[k*end + j], k = 0..d.
 * Seeds the live list like libgkm.c:556-565 and runs the reference DFS. */
void gkmref_mmprofile(int a, int end, int *out)
{
    const gkm_data *da = g_kernel->prob_svm_data[a];
    const int d = g_param.d;
    int n = da->seqlen - g_param.L + 1;
    BaseMismatchCount *live = (BaseMismatchCount *) malloc(sizeof(BaseMismatchCount) * MAX_SEQ_LENGTH);
    for (int i = 0; i < n; i++) {
        live[i].bid = da->seq + i;
        live[i].wt = da->wt[i];
        live[i].mmcnt = 0;
    }
    int **prof = (int **) malloc(sizeof(int *) * (size_t) (d + 1));
    for (int k = 0; k <= d; k++) {
        prof[k] = out + (size_t) k * (size_t) end;
        for (int j = 0; j < end; j++) prof[k][j] = 0;
    }
    kmertree_dfs(g_kernel->prob_kmertree, end, 0, 0, live, n, prof);
    free(prof);
    free(live);
}

/* normalised kernel row through the public entry point (libgkm.c:1156) */
void gkmref_row(int a, int start, int end, double *res)
{
    gkmkernel_kernelfunc_batch_all(g_kernel, a, start, end, res);
}

/* ---- CPU baseline driver: a stated subsample of rows of the open problem,
 * row-interleaved over nthreads like gkmkern_pylib.c:70-90 ---- */
typedef struct {
    const int *rows; int nrows; int tid; int nthreads; int endcap; double checksum;
} rows_job_t;

static void *rows_worker(void *p)
{
    rows_job_t *job = (rows_job_t *) p;
    double *res = (double *) malloc(sizeof(double) * (size_t) (g_prob.l + 1));
    double cs = 0.0;
    for (int i = job->tid; i < job->nrows; i += job->nthreads) {
        int a = job->rows[i];
        int end = (job->endcap > 0) ? job->endcap : a;
        if (end <= 0) continue;
        gkmkernel_kernelfunc_batch_all(g_kernel, a, 0, end, res);
        for (int j = 0; j < end; j++) cs += res[j];
    }
    free(res);
    job->checksum = cs;
    return 0;
}

/* returns wall seconds; endcap>0 selects the rectangular (test x SV) shape
 * batch_all(a, 0, endcap) instead of the triangular batch_all(a, 0, a). */
double gkmref_rows_timed(const int *rows, int nrows, int nthreads, int endcap, double *checksum)
{
    struct timespec t0, t1;
    if (nthreads < 1) nthreads = 1;
    rows_job_t *jobs = (rows_job_t *) malloc(sizeof(rows_job_t) * (size_t) nthreads);
    pthread_t *th = (pthread_t *) malloc(sizeof(pthread_t) * (size_t) nthreads);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nthreads; t++) {
        jobs[t].rows = rows; jobs[t].nrows = nrows; jobs[t].tid = t;
        jobs[t].nthreads = nthreads; jobs[t].endcap = endcap; jobs[t].checksum = 0.0;
        if (t > 0) pthread_create(&th[t], NULL, rows_worker, &jobs[t]);
    }
    rows_worker(&jobs[0]);
    double cs = jobs[0].checksum;
    for (int t = 1; t < nthreads; t++) { pthread_join(th[t], NULL); cs += jobs[t].checksum; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (checksum) *checksum = cs;
    free(jobs); free(th);
    return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}

/* ---- parity driver: many rows at once, the reference's own numbers ----
 * For every i: kout[i*ld + j] = K(rows[i], j) for j < end_i through the public entry point (libgkm.c:1156) and, when
 * hout is given, hout[(i*(d+1) + m)*ld + j] = H_m(rows[i], j) from the reference's own DFS (gkmref_mmprofile above);
 * end_i = endcap > 0 ? endcap : rows[i].  Rows are interleaved over nthreads like gkmkern_pylib.c:70-90. */
typedef struct { const int *rows; int nrows, tid, nthreads, endcap; double *kout; int *hout; long ld; } vals_job_t;

static void *vals_worker(void *p)
{
    vals_job_t *job = (vals_job_t *) p;
    const int nb = g_param.d + 1;
    int *tmp = job->hout ? (int *) malloc(sizeof(int) * (size_t) nb * (size_t) (g_prob.l + 1)) : NULL;
    for (int i = job->tid; i < job->nrows; i += job->nthreads) {
        const int a = job->rows[i];
        const int end = (job->endcap > 0) ? job->endcap : a;
        if (end <= 0) continue;
        gkmkernel_kernelfunc_batch_all(g_kernel, a, 0, end, job->kout + (size_t) i * (size_t) job->ld);
        if (tmp) {
            gkmref_mmprofile(a, end, tmp);
            for (int m = 0; m < nb; m++)
                memcpy(job->hout + ((size_t) i * (size_t) nb + (size_t) m) * (size_t) job->ld, tmp + (size_t) m * (size_t) end, sizeof(int) * (size_t) end);
        }
    }
    free(tmp);
    return 0;
}

double gkmref_rows_values(const int *rows, int nrows, int nthreads, int endcap, double *kout, int *hout, long ld)
{
    struct timespec t0, t1;
    if (nthreads < 1) nthreads = 1;
    vals_job_t *jobs = (vals_job_t *) malloc(sizeof(vals_job_t) * (size_t) nthreads);
    pthread_t *th = (pthread_t *) malloc(sizeof(pthread_t) * (size_t) nthreads);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < nthreads; t++) {
        vals_job_t j = { rows, nrows, t, nthreads, endcap, kout, hout, ld };
        jobs[t] = j;
        if (t > 0) pthread_create(&th[t], NULL, vals_worker, &jobs[t]);
    }
    vals_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(jobs); free(th);
    return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}

void gkmref_close(void)
{
    if (!g_kernel) return;
    for (int i = 0; i < g_prob.l; i++) gkmkernel_delete_object(g_prob.x[i]);
    free(g_prob.y);
    free(g_prob.x);
    gkmkernel_destroy(g_kernel);
    g_kernel = NULL;
}

} /* extern "C" */
