/* gkm_capi.c -- the exported entry points: the extended ABI (include/gkm_b200.h) and
 * the reference's own ABI (include/gkm_abi.h == src/libgkm.h of the reference),
 * including gkm_main_pywrapper, the operator gkmQC calls (gkmkern_pylib.c:92-246).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "gkm_internal.h"
#include "gkm_options.h"

/* ================================================================== */
/* extended ABI                                                        */
/* ================================================================== */
int gkmb200_problem_upload(gkmb200_problem *p)
{
    if (!p) { gkm_set_error("null problem"); return 1; }
    return gkm_dev_upload(p);
}

int gkmb200_problem_sqnorm(gkmb200_problem *p, double *out)
{
    if (!p || !out) { gkm_set_error("null argument"); return 1; }
    if (gkm_dev_upload(p)) return 1;
    memcpy(out, p->sqnorm, sizeof(double) * (size_t) p->n);
    return 0;
}

int gkmb200_kernel_lower(gkmb200_problem *p, double **rows, int copy_threads)
{
    if (!p || !rows) { gkm_set_error("null argument"); return 1; }
    return gkm_dev_compute(p, 0, p->n, 0, p->n, 1, NULL, 0, rows, NULL, copy_threads);
}

int gkmb200_kernel_block(gkmb200_problem *p, int row0, int nrows, int col0, int ncols, int lower, double *out, long ld)
{
    if (!p || !out) { gkm_set_error("null argument"); return 1; }
    if (ld < ncols) { gkm_set_error("ld < ncols"); return 1; }
    return gkm_dev_compute(p, row0, nrows, col0, ncols, lower, out, ld, NULL, NULL, 4);
}

int gkmb200_hist_block(gkmb200_problem *p, int row0, int nrows, int col0, int ncols, int lower, int32_t *hist)
{
    if (!p || !hist) { gkm_set_error("null argument"); return 1; }
    return gkm_dev_compute(p, row0, nrows, col0, ncols, lower, NULL, 0, NULL, hist, 1);
}

int gkmb200_decision_values(gkmb200_problem *p, int row0, int nrows, int col0, int ncols,
                            const double *alpha, double bias, double *out)
{
    return gkm_dev_decision(p, row0, nrows, col0, ncols, alpha, bias, out);
}

static gkmb200_stats g_last_stats; /* of the most recent gkm_main_pywrapper call */
static long long g_last_nk = 0;

int gkmb200_get_stats(const gkmb200_problem *p, gkmb200_stats *out)
{
    if (!out) return 1;
    if (!p) { /* the pywrapper's problem does not outlive the call (ownership) */
        *out = g_last_stats;
        out->lmer_pairs = out->entries * g_last_nk * 2 * g_last_nk;
        return 0;
    }
    *out = p->stats;
    /* L-mer pair comparisons behind the entries of the last call, for uniform-length problems:
     * n_a * 2 n_b per entry; ragged problems report the maximum-length figure */
    int maxlen = 0;
    for (int i = 0; i < p->n; i++) if (p->len[i] > maxlen) maxlen = p->len[i];
    long long nk = maxlen - p->param.L + 1;
    out->lmer_pairs = out->entries * nk * 2 * nk;
    return 0;
}

int gkmb200_bench_lower_resident(gkmb200_problem *p, int steps, int warmup, int flush_l2, double *ms_each)
{
    return gkm_dev_bench_lower(p, steps, warmup, flush_l2, ms_each);
}

int gkmb200_microbench(const char *what, double *result) { return gkm_dev_microbench(what, result); }

int gkmb200_svm_cv(gkmb200_problem *p, const double *kmat, long ld, int n, int ntasks, const gkmb200_svm_task *tasks,
                   const int *train_idx, const signed char *train_y, const int *test_idx,
                   double C, double eps, int max_iter, double *scores, gkmb200_svm_fit *fits, double *alpha)
{
    return gkm_dev_svm_cv(p, kmat, ld, n, ntasks, tasks, train_idx, train_y, test_idx, C, eps, max_iter, scores, fits, alpha);
}

/* ================================================================== */
/* the operator: gkm_main_pywrapper (gkmkern_pylib.c:92-246)           */
/* ================================================================== */
int gkm_main_pywrapper(gkmOpt *opts, double **kmat, int *kmat_size)
{
    if (!opts || !kmat || !kmat_size) { fprintf(stderr, "gkm_main_pywrapper: null argument\n"); return 1; }
    if (opts->verbosity < 0 || opts->verbosity > 4) {
        /* the reference prints this and calls exit(0), taking the Python host down with it
         * (gkmkern_pylib.c:135-137); an error return reaches the same sys.exit() in gkmsvm.py:90-92 */
        fprintf(stderr, "Unknown verbosity: %d\n", opts->verbosity);
        return 1;
    }
    gkmb200_set_verbosity(opts->verbosity);

    gkm_parameter param;
    memset(&param, 0, sizeof(param));
    param.kernel_type = opts->kernel_type;
    param.L = opts->L;
    param.k = opts->k;
    param.d = opts->d;
    param.M = opts->M;
    param.H = opts->H;
    param.gamma = opts->gamma;
    param.nthreads = 1;

    gkm_log(GKM_LOG_INFO, "Arguments:");
    gkm_log(GKM_LOG_INFO, "  posfile = %s", opts->posfile);
    gkm_log(GKM_LOG_INFO, "  negfile = %s", opts->negfile);
    gkm_log(GKM_LOG_INFO, "Parameters:");
    gkm_log(GKM_LOG_INFO, "  kernel-type = %d", param.kernel_type);
    gkm_log(GKM_LOG_INFO, "  L = %d", param.L);
    gkm_log(GKM_LOG_INFO, "  k = %d", param.k);
    gkm_log(GKM_LOG_INFO, "  d = %d", param.d);
    if (param.kernel_type == EST_TRUNC_RBF || param.kernel_type == EST_TRUNC_PW_RBF)
        gkm_log(GKM_LOG_INFO, "  gamma = %g", param.gamma);
    if (param.kernel_type == EST_TRUNC_PW || param.kernel_type == EST_TRUNC_PW_RBF) {
        gkm_log(GKM_LOG_INFO, "  M = %d", param.M);
        gkm_log(GKM_LOG_INFO, "  H = %g", param.H);
    }

    const char *bad = gkm_param_problem(&param, gkm_opt_max_L()); /* L <= 12 unless GKM_MAX_L=16 */
    if (bad) { gkm_set_error("%s", bad); return 1; }
    if (!opts->posfile || !opts->negfile) { gkm_set_error("can't open file"); return 1; }

    gkmb200_problem *p = gkmb200_problem_new(&param);
    if (!p) return 1;
    gkm_problem_shard_from_env(p); /* one process per GPU: this call fills only the chunks of its rank */
    if (p->shard_world > 1)
        gkm_log(GKM_LOG_WARN, "GKM_SHARD=%d/%d: this process computes only its own chunks of the matrix", p->shard_rank, p->shard_world);
    int rc = 1;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int npos = gkmb200_problem_read(p, opts->posfile, opts->negfile);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double read_ms = 1e3 * (double) (t1.tv_sec - t0.tv_sec) + 1e-6 * (double) (t1.tv_nsec - t0.tv_nsec);
    if (npos >= 0 && p->n > 0) {
        int nt = opts->nthreads;
        if (nt < 1) nt = 1;
        rc = gkmb200_kernel_lower(p, kmat, nt);
        if (!rc) {
            kmat_size[0] = npos;
            kmat_size[1] = p->n - npos;
            g_last_stats = p->stats;
            g_last_nk = 0;
            for (int i = 0; i < p->n; i++) if (p->len[i] - param.L + 1 > g_last_nk) g_last_nk = p->len[i] - param.L + 1;
            gkm_log(GKM_LOG_INFO, "kernel matrix: %d sequences, %lld entries, %d GPU(s): read %.1f ms, pack+upload+sqnorm %.1f ms, "
                    "kernels %.1f ms (device), compute+copy-out %.1f ms (wall)",
                    p->n, p->stats.entries, p->stats.devices, read_ms, p->stats.upload_ms, p->stats.kernel_ms, p->stats.wall_ms);
        }
    } else if (npos >= 0) {
        gkm_set_error("no sequences in %s / %s", opts->posfile, opts->negfile);
    }
    gkmb200_problem_free(p);
    return rc;
}

/* ================================================================== */
/* the rest of libgkm.h, on top of the same engine                     */
/* ================================================================== */
/* per-kernel engine state hangs off gkm_kernel.prob_kmertree (opaque to callers) */
typedef struct gkm_shadow {
    gkmb200_problem *prob; /* device-resident image of prob_svm_data */
    int dirty;
    gkm_data **read_x;     /* the objects gkmkernel_read_problems made `prob` from, in image order */
    int read_n;
} gkm_shadow;

static gkm_shadow *shadow_of(gkm_kernel *kernel) { return (gkm_shadow *) (void *) kernel->prob_kmertree; }

void gkmkernel_set_num_threads(gkm_parameter *param)
{
    /* the reference picks a 1/4/16-thread DFS here (libgkm.c:1187-1205); the GPU engine has no
     * such mode, but the observable side effect on the parameter is kept */
    if (param->nthreads != 1 && param->nthreads != 4 && param->nthreads != 16) {
        gkm_log(GKM_LOG_WARN, "Supported number of threads are 1, 4 and 16. nthread is set to 1");
        param->nthreads = 1;
    }
}

gkm_kernel *gkmkernel_init(gkm_parameter *param)
{
    if (!param) return NULL;
    gkmkernel_set_num_threads(param);
    const char *bad = gkm_param_problem(param, GKM_MAX_L);
    if (bad) { gkm_set_error("%s", bad); return NULL; }
    gkm_kernel *kernel = (gkm_kernel *) calloc(1, sizeof(gkm_kernel));
    if (!kernel) return NULL;
    kernel->param = param;
    double w[GKM_MAX_L + 1];
    gkm_calc_weights(param->kernel_type, param->L, param->k, w);
    for (int i = 0; i <= MAX_MM && i <= param->L; i++) kernel->weights[i] = w[i];
    gkm_log(GKM_LOG_DEBUG, "gkm-kernel weights:");
    for (int i = 0; i <= param->d; i++) gkm_log(GKM_LOG_DEBUG, "  c[%d] = %.6f", i, w[i]);
    gkm_shadow *sh = (gkm_shadow *) calloc(1, sizeof(gkm_shadow));
    kernel->prob_kmertree = (KmerTree *) (void *) sh;
    kernel->mmcnt_nlookups = (param->L <= MMCNT_LOOKUPTAB_WIDTH) ? 1 : 2;
    kernel->mmcnt_lookuptab_mask = 0xFFFF;
    return kernel;
}

void gkmkernel_destroy(gkm_kernel *kernel)
{
    if (!kernel) return;
    gkm_shadow *sh = shadow_of(kernel);
    if (sh) { gkmb200_problem_free(sh->prob); free(sh->read_x); free(sh); }
    free(kernel->prob_svm_data);
    free(kernel->prob_gkmkernel_index);
    free(kernel->prob_libsvm_index);
    free(kernel);
}

static char *dup_string(const char *s)
{
    size_t n = strlen(s) + 1;
    char *r = (char *) malloc(n);
    if (r) memcpy(r, s, n);
    return r;
}

/* the arrays of a gkm_data from the letters of a sequence (libgkm.c:864-932): seq, seq_rc (codes 1..4), seq_string, wt, wt_rc
 * and the leaf index of every L-mer in the reference's 4-ary tree (libgkm.c:891-908: base-4 digits 0..3).
 * Returns 1 when an allocation failed (the caller deletes the object). */
static int fill_object(gkm_data *d, const gkm_parameter *pa, const uint8_t *text, int len, const char *sid, int seqid)
{
    const int nk = len - pa->L + 1;
    d->sid = sid ? dup_string(sid) : NULL;
    d->seqid = seqid;
    d->seqlen = len;
    d->seq_string = (char *) malloc((size_t) len + 1);
    d->seq = (u_int8_t *) malloc((size_t) len);
    d->seq_rc = (u_int8_t *) malloc((size_t) len);
    d->wt = (u_int8_t *) malloc((size_t) nk);
    d->wt_rc = (u_int8_t *) malloc((size_t) nk);
    d->kmerids = (int *) malloc(sizeof(int) * (size_t) nk);
    d->kmerids_rc = (int *) malloc(sizeof(int) * (size_t) nk);
    if ((sid && !d->sid) || !d->seq_string || !d->seq || !d->seq_rc || !d->wt || !d->wt_rc || !d->kmerids || !d->kmerids_rc) {
        gkm_set_error("out of memory");
        return 1;
    }
    for (int j = 0; j < len; j++) {
        d->seq[j] = (u_int8_t) (gkm_base_code(text[j]) + 1);
        d->seq_rc[j] = (u_int8_t) (4 - gkm_base_code(text[len - 1 - j]));
        d->seq_string[j] = (char) text[j]; /* the reference keeps the spelling it was given (strcpy, libgkm.c:859-860) */
    }
    d->seq_string[len] = '\0';
    gkm_calc_posweights(nk, pa->kernel_type, pa->M, pa->H, d->wt, d->wt_rc);
    u_int8_t *strands[2] = { d->seq, d->seq_rc };
    int *ids[2] = { d->kmerids, d->kmerids_rc };
    for (int s = 0; s < 2; s++)
        for (int j = 0; j < nk; j++) {
            int v = 0;
            for (int i = 0; i < pa->L; i++) v = v * 4 + (strands[s][i + j] - 1);
            ids[s][j] = v;
        }
    return 0;
}

/* a gkm_data laid out like the reference's (libgkm.c:841-938); sqnorm comes from the GPU.
 * The reference bounds the length only in read_fasta_file (2047 bases, libgkm.c:1294-1299); the engine holds at
 * most GKM_MAX_BASES per sequence, so a longer string handed to this function directly is cut there as well
 * (with the warning the reader would have given) and seqlen / seq_string describe what is actually used. */
gkm_data *gkmkernel_new_object(gkm_kernel *kernel, char *seq, char *sid, int seqid)
{
    if (!kernel || !seq) return NULL;
    const gkm_parameter *pa = kernel->param;
    size_t full = strlen(seq);
    if (full > GKM_MAX_BASES) {
        gkm_log(GKM_LOG_WARN, "maximum sequence length allowed is %d. The first %d nucleotides of %s will only be used",
                GKM_MAX_BASES, GKM_MAX_BASES, sid ? sid : "?");
        full = GKM_MAX_BASES;
    }
    const int len = (int) full;
    if (len - pa->L + 1 < 1) { gkm_set_error("sequence %s is shorter than L", sid ? sid : "?"); return NULL; }
    gkm_data *d = (gkm_data *) calloc(1, sizeof(gkm_data));
    if (!d) return NULL;
    /* one-sequence problem: coding, and sqnorm = sqrt(Kraw(x,x)) on the device */
    gkmb200_problem *one = gkmb200_problem_new(pa);
    if (!one || gkmb200_problem_add(one, seq, len) < 0 || fill_object(d, pa, gkm_letters(one, 0), len, sid, seqid) || gkm_dev_upload(one)) {
        gkmb200_problem_free(one);
        gkmkernel_delete_object(d);
        return NULL;
    }
    d->sqnorm = one->sqnorm[0];
    gkmb200_problem_free(one);
    gkm_log(GKM_LOG_TRACE, "%d's sqnorm is %f", seqid, d->sqnorm);
    return d;
}

void gkmkernel_free_object(gkm_data *d)
{
    if (!d) return;
    free(d->kmerids); free(d->kmerids_rc); free(d->seq_string); free(d->wt); free(d->wt_rc);
    free(d->seq); free(d->seq_rc); free(d->sid);
    d->kmerids = d->kmerids_rc = NULL;
    d->seq_string = NULL; d->sid = NULL;
    d->wt = d->wt_rc = d->seq = d->seq_rc = NULL;
}

void gkmkernel_delete_object(gkm_data *d)
{
    if (!d) return;
    gkmkernel_free_object(d);
    free(d);
}

/* one caller-held object -> one sequence of an engine problem.  The object may come from anywhere (the ABI lets
 * callers build gkm_data themselves): its length is checked against what the engine holds, its arrays against NULL. */
static int add_object(gkmb200_problem *prob, const gkm_data *x, int i)
{
    static const char letters[4] = { 'A', 'C', 'G', 'T' };
    char buf[GKM_MAX_BASES + 1];
    if (!x || !x->seq) { gkm_set_error("object %d is null or was freed with gkmkernel_free_object", i); return 1; }
    if (x->seqlen < 1 || x->seqlen > GKM_MAX_BASES) { gkm_set_error("object %d has %d bases; the engine holds 1..%d", i, x->seqlen, GKM_MAX_BASES); return 1; }
    for (int j = 0; j < x->seqlen; j++) buf[j] = letters[(x->seq[j] - 1) & 3];
    return gkmb200_problem_add(prob, buf, x->seqlen) < 0;
}

/* rebuild the device image from prob_svm_data (order = current gkmkernel index order) */
static int shadow_sync(gkm_kernel *kernel)
{
    gkm_shadow *sh = shadow_of(kernel);
    if (!sh) return 1;
    if (sh->prob && !sh->dirty) return 0;
    gkmb200_problem_free(sh->prob);
    sh->prob = gkmb200_problem_new(kernel->param);
    if (!sh->prob) return 1;
    for (int i = 0; i < kernel->prob_num; i++)
        if (add_object(sh->prob, kernel->prob_svm_data[i], i)) return 1;
    sh->dirty = 0;
    return gkm_dev_upload(sh->prob);
}

/* "build the tree" (libgkm.c:1035-1056): here, make the sequence set resident on the GPUs */
void gkmkernel_build_tree(gkm_kernel *kernel, gkm_data **x, int n)
{
    if (!kernel || !x || n < 0) return;
    free(kernel->prob_svm_data);
    free(kernel->prob_gkmkernel_index);
    free(kernel->prob_libsvm_index);
    kernel->prob_svm_data = (gkm_data **) malloc(sizeof(gkm_data *) * (size_t) (n ? n : 1));
    kernel->prob_gkmkernel_index = (int *) malloc(sizeof(int) * (size_t) (n ? n : 1));
    kernel->prob_libsvm_index = (int *) malloc(sizeof(int) * (size_t) (n ? n : 1));
    if (!kernel->prob_svm_data || !kernel->prob_gkmkernel_index || !kernel->prob_libsvm_index) {
        free(kernel->prob_svm_data); free(kernel->prob_gkmkernel_index); free(kernel->prob_libsvm_index);
        kernel->prob_svm_data = NULL; kernel->prob_gkmkernel_index = NULL; kernel->prob_libsvm_index = NULL;
        kernel->prob_num = 0;
        gkm_set_error("out of memory");
        gkm_log(GKM_LOG_ERROR, "gkmkernel_build_tree: out of memory");
        return;
    }
    memcpy(kernel->prob_svm_data, x, sizeof(gkm_data *) * (size_t) n);
    kernel->prob_num = n;
    for (int i = 0; i < n; i++) { kernel->prob_gkmkernel_index[i] = i; kernel->prob_libsvm_index[i] = i; }
    gkm_shadow *sh = shadow_of(kernel);
    if (sh) {
        /* the image gkmkernel_read_problems uploaded serves as it is when the tree is built over exactly its objects */
        const int same = sh->prob && sh->read_x && sh->read_n == n && n > 0 && memcmp(sh->read_x, x, sizeof(gkm_data *) * (size_t) n) == 0;
        sh->dirty = !same;
        free(sh->read_x); sh->read_x = NULL; sh->read_n = 0;
    }
    if (n > 0 && shadow_sync(kernel)) gkm_log(GKM_LOG_ERROR, "gkmkernel_build_tree: %s", gkmb200_last_error());
}

/* K(prob[a], prob[j]) for j in [start,end) (libgkm.c:1156-1185) */
double *gkmkernel_kernelfunc_batch_all(gkm_kernel *kernel, const int a, const int start, const int end, double *res)
{
    if (!kernel || !res || end < start) return res;
    for (int j = 0; j < end - start; j++) res[j] = 0;
    if (end == start) return res;
    if (shadow_sync(kernel) ||
        gkm_dev_compute(shadow_of(kernel)->prob, a, 1, start, end - start, 0, res, end - start, NULL, NULL, 1))
        gkm_log(GKM_LOG_ERROR, "gkmkernel_kernelfunc_batch_all: %s", gkmb200_last_error());
    return res;
}

/* the intended meaning of libgkm.c:1115-1153 -- K(prob[a], db_array[i]), normalised (+RBF);
 * see SURVEY.md 3.4 for why the reference's own body does not compute that */
double *gkmkernel_kernelfunc_batch(gkm_kernel *kernel, int a, const gkm_data **db_array, const int n, double *res)
{
    if (!kernel || !res || n < 0) return res;
    for (int i = 0; i < n; i++) res[i] = 0;
    if (n == 0 || a < 0 || a >= kernel->prob_num) return res;
    if (!db_array) return res;
    gkmb200_problem *tmp = gkmb200_problem_new(kernel->param);
    int ok = tmp != NULL;
    for (int i = 0; ok && i <= n; i++) /* db_array[0..n-1] then the query */
        ok = !add_object(tmp, (i < n) ? db_array[i] : kernel->prob_svm_data[a], i);
    if (!ok || gkm_dev_compute(tmp, n, 1, 0, n, 0, res, n, NULL, NULL, 1))
        gkm_log(GKM_LOG_ERROR, "gkmkernel_kernelfunc_batch: %s", gkmb200_last_error());
    gkmb200_problem_free(tmp);
    return res;
}

/* single pair; declared in libgkm.h:140 but defined nowhere upstream */
double gkmkernel_kernelfunc(const gkm_data *da, const gkm_data *db)
{
    (void) da; (void) db;
    gkm_set_error("gkmkernel_kernelfunc needs a kernel handle; use gkmkernel_kernelfunc_batch");
    return NAN;
}

/* libgkm.c:1316-1333 + read_fasta_file: every record becomes a gkm_data with the fields the reference fills (sid,
 * seq, seq_rc, seq_string, kmerids, kmerids_rc, wt, wt_rc, sqnorm) and, beyond the reference, its label. */
int gkmkernel_read_problems(gkm_kernel *kernel, svm_problem *prob, const char *posfile, const char *negfile)
{
    if (!kernel || !prob) return -1;
    gkmb200_problem *p = gkmb200_problem_new(kernel->param);
    if (!p) return -1;
    int npos = gkmb200_problem_read(p, posfile, negfile);
    if (npos < 0 || gkm_dev_upload(p)) { gkmb200_problem_free(p); return -1; }
    const int n = p->n;
    prob->l = n;
    prob->y = (double *) malloc(sizeof(double) * (size_t) (n ? n : 1));
    prob->x = (gkm_data **) calloc((size_t) (n ? n : 1), sizeof(gkm_data *));
    int bad = !prob->y || !prob->x;
    for (int i = 0; !bad && i < n; i++) {
        gkm_data *d = (gkm_data *) calloc(1, sizeof(gkm_data));
        prob->x[i] = d;
        if (!d || fill_object(d, kernel->param, gkm_letters(p, i), p->len[i], p->sid[i] ? p->sid[i] : "", i)) { bad = 1; break; }
        d->label = (i < npos) ? 1 : -1;
        d->sqnorm = p->sqnorm[i];
        prob->y[i] = d->label;
    }
    if (bad) {
        gkm_set_error("out of memory reading %s / %s", posfile, negfile);
        if (prob->x) for (int i = 0; i < n; i++) gkmkernel_delete_object(prob->x[i]);
        free(prob->x); free(prob->y);
        prob->x = NULL; prob->y = NULL; prob->l = 0;
        gkmb200_problem_free(p);
        return -1;
    }
    /* keep the uploaded image: gkmkernel_build_tree on exactly these objects, in this order, adopts it
     * instead of packing and uploading the same sequences again */
    gkm_shadow *sh = shadow_of(kernel);
    if (sh) {
        gkmb200_problem_free(sh->prob);
        sh->prob = p; sh->dirty = 1;
        free(sh->read_x);
        sh->read_x = (gkm_data **) malloc(sizeof(gkm_data *) * (size_t) (n ? n : 1));
        sh->read_n = sh->read_x ? n : 0;
        if (sh->read_x) memcpy(sh->read_x, prob->x, sizeof(gkm_data *) * (size_t) n);
    } else gkmb200_problem_free(p);
    return npos;
}

void gkmkernel_swap_index(gkm_kernel *kernel, int i, int j)
{
    if (!kernel || !kernel->prob_gkmkernel_index || i < 0 || j < 0 || i >= kernel->prob_num || j >= kernel->prob_num) return;
    int *gi = kernel->prob_gkmkernel_index, *li = kernel->prob_libsvm_index;
    int t = li[gi[i]]; li[gi[i]] = li[gi[j]]; li[gi[j]] = t;
    t = gi[i]; gi[i] = gi[j]; gi[j] = t;
}

/* apply the accumulated permutation (libgkm.c:1084-1109): afterwards id i means what
 * libsvm calls i; the device image is rebuilt in the new order on next use */
void gkmkernel_update_index(gkm_kernel *kernel)
{
    if (!kernel || !kernel->prob_svm_data) return;
    const int n = kernel->prob_num;
    gkm_data **fresh = (gkm_data **) malloc(sizeof(gkm_data *) * (size_t) (n ? n : 1));
    if (!fresh) { gkm_set_error("out of memory"); gkm_log(GKM_LOG_ERROR, "gkmkernel_update_index: out of memory"); return; }
    for (int i = 0; i < n; i++) fresh[i] = kernel->prob_svm_data[kernel->prob_gkmkernel_index[i]];
    free(kernel->prob_svm_data);
    kernel->prob_svm_data = fresh;
    for (int i = 0; i < n; i++) { kernel->prob_gkmkernel_index[i] = i; kernel->prob_libsvm_index[i] = i; }
    gkm_shadow *sh = shadow_of(kernel);
    if (sh) sh->dirty = 1;
}
