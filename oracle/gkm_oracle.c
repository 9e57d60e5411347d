/* oracle/gkm_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, CPU, brute-force restatement of the gkm-kernel hot path of
 * Dongwon-Lee/gkmQC.  It exists to CHECK the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * it.  The product (gkmqc_b200/bin/gkmkern_pylib.so) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * below against vectors produced by the unmodified reference compiled from
 * /root/reference/src (oracle/Makefile target `ref`, generator
 * oracle/gen_golden.py, fixtures in tests/golden/).  The reference itself ships
 * no tests or golden vectors.
 *
 * Each function cites the reference lines whose arithmetic it restates.  The
 * reference walks a k-mer tree (libgkm.c:315-387); this file does NOT: it uses
 * the dense double loop that the reference itself uses for the diagonal
 * (libgkm.c:738-751) for every sequence pair.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define GKMO_MAX_SEQ 2048 /* libgkm.h:32 */
#define GKMO_MAX_L 16

/* ------------------------------------------------------------------ */
/* w[m]: libgkm.c:73-217                                               */
/* ------------------------------------------------------------------ */

/* libgkm.c:73-105 -- binomial by Pascal additions, extended to n<0 */
static double o_comb(int n, int r)
{
    if (r < 0) return 0;
    if (n < 0) return o_comb(r - n - 1, r) * ((r % 2 == 0) ? 1 : -1);
    if (n < r) return 0;
    if (n == 0 && r == 0) return 1.0;
    double cur[64], old[64];
    for (int i = 0; i <= r; i++) cur[i] = old[i] = 0;
    cur[0] = old[0] = 1;
    for (int i = 1; i <= n; i++) {
        for (int j = 0; j <= r; j++) old[j] = cur[j];
        for (int j = 1; j <= r; j++) cur[j] = old[j] + old[j - 1];
    }
    return cur[r];
}

/* libgkm.c:204-217 -- type 0: w[m] = C(L-m, k) while L-m >= k; other slots untouched */
static void o_weights_gkm(int L, int K, double *w)
{
    for (int i = 0; i <= L; i++)
        if ((L - i) >= K) w[i] = o_comb(L - i, K);
}

/* libgkm.c:107-202 -- types 1..5: estimated l-mer weights, full or truncated filter */
static void o_weights_est(int L, int K, int truncated, double *w)
{
    const int b = 4;
    double A[GKMO_MAX_L + 2][GKMO_MAX_L + 2], B[GKMO_MAX_L + 2][GKMO_MAX_L + 2];
    double (*wL)[GKMO_MAX_L + 2] = A, (*wLp)[GKMO_MAX_L + 2] = B, (*sw)[GKMO_MAX_L + 2];
    double wm[GKMO_MAX_L + 2], kern[GKMO_MAX_L + 2], kernTr[GKMO_MAX_L + 2];

    for (int i = 0; i <= K; i++)
        for (int j = 0; j <= K; j++) wL[i][j] = wLp[i][j] = 1.0; /* :124-131 */

    for (int iL = 1; iL <= L; iL++) { /* :133-143 */
        for (int iK = 1; iK <= K; iK++) {
            wL[iK][0] = wLp[iK][0] + (b - 1) * wLp[iK - 1][0];
            for (int jM = 1; jM <= iK; jM++) wL[iK][jM] = (wL[iK - 1][jM - 1] * (iK - iL)) / iK;
        }
        sw = wLp; wLp = wL; wL = sw;
    }

    double nnorm = o_comb(L, K) * pow(b, 1.0 * L); /* :145 */
    for (int i = 0; i <= K; i++) wm[i] = wLp[K][i] / nnorm;

    for (int m = 0; m <= L; m++) { /* :152-158 */
        int ub = (m < K) ? m : K;
        kern[m] = 0;
        for (int i = 0; i <= ub; i++) kern[m] += wm[i] * o_comb(L - m, K - i) * o_comb(m, i);
    }

    int keep = 1; /* :160-168 */
    for (int i = 0; i <= L; i++) {
        if (kern[i] < 1e-50) keep = 0;
        kernTr[i] = keep ? kern[i] : 0.0;
    }

    for (int m = 0; m <= L; m++) { /* :171-191 */
        double acc = 0;
        for (int m1 = 0; m1 <= L; m1++)
            for (int m2 = 0; m2 <= L; m2++)
                for (int t = 0; t <= L; t++) {
                    int r = m1 + m2 - 2 * t - L + m;
                    if ((t <= m) && ((m1 - t) <= (L - m)) && (r <= (m1 - t)) && (r >= 0)) {
                        double cc = o_comb(m, t) * o_comb(L - m, m1 - t) * o_comb(m1 - t, r) *
                                    pow(b - 1, 1.0 * t) * pow(b - 2, 1.0 * r);
                        if (truncated) acc += cc * kernTr[m1] * kernTr[m2];
                        else acc += cc * kern[m1] * kern[m2];
                    }
                }
        w[L - m] = acc;
    }
}

/* dispatch libgkm.c:997-1019.  w must hold L+1 doubles; slots the reference
 * leaves unwritten (type 0, m > L-k) are set to 0 here and never read (d <= L-k). */
int gkmo_weights(int kernel_type, int L, int k, double *w)
{
    if (L < 1 || L > GKMO_MAX_L || k < 0 || k > L) return 1;
    for (int i = 0; i <= L; i++) w[i] = 0.0;
    if (kernel_type == 0) o_weights_gkm(L, k, w);
    else if (kernel_type == 1) o_weights_est(L, k, 0, w);
    else o_weights_est(L, k, 1, w);
    return 0;
}

/* ------------------------------------------------------------------ */
/* per-sequence preparation: libgkm.c:841-938                          */
/* ------------------------------------------------------------------ */

/* libgkm.c:864-888: A,C,G,T -> 1,2,3,4 (case-insensitive), anything else -> 1;
 * reverse complement = 5 - code, reversed */
void gkmo_encode(const char *s, int len, uint8_t *fwd, uint8_t *rc)
{
    for (int i = 0; i < len; i++) {
        switch (toupper((unsigned char) s[i])) {
            case 'A': fwd[i] = 1; break;
            case 'C': fwd[i] = 2; break;
            case 'G': fwd[i] = 3; break;
            case 'T': fwd[i] = 4; break;
            default: fwd[i] = 1; break;
        }
    }
    for (int i = 0; i < len; i++) rc[i] = (uint8_t) (5 - fwd[len - i - 1]);
}

/* libgkm.c:910-932: positional weights; uniform for types 0..3 */
void gkmo_poswt(int nk, int kernel_type, int M, double H, uint8_t *wt, uint8_t *wt_rc)
{
    int center = nk / 2;
    if (kernel_type == 4 || kernel_type == 5) {
        for (int i = 0; i < nk; i++) {
            double v = floor(M * exp((-1) * log(2) * abs(center - i) / H) + 1);
            uint8_t u = (uint8_t) (int) v; /* u_int8_t cast of the reference, mod 256 on x86 */
            if (u > M) u = (uint8_t) M;
            wt[i] = u;
            wt_rc[nk - i - 1] = u;
        }
    } else {
        for (int i = 0; i < nk; i++) { wt[i] = 1; wt_rc[nk - i - 1] = 1; }
    }
}

/* ------------------------------------------------------------------ */
/* sequence set                                                        */
/* ------------------------------------------------------------------ */
typedef struct {
    int len, nk;
    uint8_t *fwd, *rc;    /* codes 1..4 */
    uint8_t *wt, *wt_rc;  /* per L-mer start */
    uint32_t *id, *id_rc; /* 2-bit packed L-mers (code-1, first base most significant: libgkm.c:656-692) */
    double sqnorm;
    char *sid;
} o_seq;

typedef struct gkmo_set {
    int kernel_type, L, k, d, M;
    double H, gamma;
    double w[GKMO_MAX_L + 1];
    int n, cap, npos;
    o_seq *s;
} gkmo_set;

static void o_pack(const uint8_t *codes, int len, int L, uint32_t *ids)
{
    uint32_t mask = (L == 16) ? 0xFFFFFFFFu : ((1u << (2 * L)) - 1u);
    uint32_t v = 0;
    for (int i = 0; i < len; i++) {
        v = ((v << 2) | (uint32_t) (codes[i] - 1)) & mask;
        if (i >= L - 1) ids[i - L + 1] = v;
    }
}

/* number of non-zero 2-bit fields of x: what the 64 Ki byte table of
 * libgkm.c:619-654 returns for (id ^ id') */
static inline int o_mm(uint32_t x)
{
    return __builtin_popcount((x | (x >> 1)) & 0x55555555u);
}

/* truncated mismatch histogram: libgkm.c:738-751 applied to a pair.
 * a: forward strand only; b: forward and reverse-complement strands. */
static void o_hist(const gkmo_set *S, const o_seq *a, const o_seq *b, int *H)
{
    const int d = S->d;
    for (int m = 0; m <= d; m++) H[m] = 0;
    for (int i = 0; i < a->nk; i++) {
        const uint32_t x = a->id[i];
        const int wi = a->wt[i];
        for (int j = 0; j < b->nk; j++) {
            int mm = o_mm(x ^ b->id[j]);
            if (mm <= d) H[mm] += wi * (int) b->wt[j];
        }
        for (int j = 0; j < b->nk; j++) {
            int mm = o_mm(x ^ b->id_rc[j]);
            if (mm <= d) H[mm] += wi * (int) b->wt_rc[j];
        }
    }
}

/* libgkm.c:576-582 / :753-756 -- ascending m from 0.0, no contraction */
static double o_kraw(const gkmo_set *S, const int *H)
{
    double sum = 0;
    for (int m = 0; m <= S->d; m++) sum += (S->w[m] * H[m]);
    return sum;
}

gkmo_set *gkmo_set_new(int kernel_type, int L, int k, int d, int M, double H, double gamma)
{
    gkmo_set *S = (gkmo_set *) calloc(1, sizeof(gkmo_set));
    S->kernel_type = kernel_type; S->L = L; S->k = k; S->d = d; S->M = M; S->H = H; S->gamma = gamma;
    if (gkmo_weights(kernel_type, L, k, S->w)) { free(S); return NULL; }
    return S;
}

void gkmo_set_free(gkmo_set *S)
{
    if (!S) return;
    for (int i = 0; i < S->n; i++) {
        o_seq *q = &S->s[i];
        free(q->fwd); free(q->rc); free(q->wt); free(q->wt_rc); free(q->id); free(q->id_rc); free(q->sid);
    }
    free(S->s);
    free(S);
}

/* gkmkernel_new_object, libgkm.c:841-938 (sqnorm: :723-759).  Returns the id or -1. */
int gkmo_set_add(gkmo_set *S, const char *seq, const char *sid)
{
    int len = (int) strlen(seq);
    if (len < S->L) return -1;
    if (S->n == S->cap) {
        S->cap = S->cap ? 2 * S->cap : 64;
        S->s = (o_seq *) realloc(S->s, sizeof(o_seq) * (size_t) S->cap);
    }
    o_seq *q = &S->s[S->n];
    q->len = len; q->nk = len - S->L + 1;
    q->fwd = (uint8_t *) malloc((size_t) len); q->rc = (uint8_t *) malloc((size_t) len);
    q->wt = (uint8_t *) malloc((size_t) q->nk); q->wt_rc = (uint8_t *) malloc((size_t) q->nk);
    q->id = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) q->nk);
    q->id_rc = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) q->nk);
    q->sid = sid ? strdup(sid) : NULL;
    gkmo_encode(seq, len, q->fwd, q->rc);
    gkmo_poswt(q->nk, S->kernel_type, S->M, S->H, q->wt, q->wt_rc);
    o_pack(q->fwd, len, S->L, q->id);
    o_pack(q->rc, len, S->L, q->id_rc);
    int H[GKMO_MAX_L + 1];
    o_hist(S, q, q, H);
    q->sqnorm = sqrt(o_kraw(S, H));
    return S->n++;
}

/* read_fasta_file, libgkm.c:1251-1314: '>' starts a record, id = first blank-delimited
 * token, sequence lines are concatenated, a line ends at the first CR or LF,
 * at most 2047 bases are kept.  Returns the number of records appended, -1 on I/O error. */
int gkmo_set_read_fasta(gkmo_set *S, const char *path)
{
    FILE *fp = fopen(path, "r");
    if (!fp) return -1;
    size_t cap = 1 << 16;
    char *line = (char *) malloc(cap);
    char *seq = (char *) malloc(GKMO_MAX_SEQ);
    char sid[GKMO_MAX_SEQ];
    int have = 0, added = 0, seqlen = 0;
    seq[0] = 0; sid[0] = 0;
    for (;;) {
        size_t n = 0; int c;
        while ((c = fgetc(fp)) != EOF && c != '\n') {
            if (n + 2 > cap) { cap *= 2; line = (char *) realloc(line, cap); }
            line[n++] = (char) c;
        }
        if (c == EOF && n == 0) break;
        line[n] = 0;
        line[strcspn(line, "\r\n")] = 0;
        if (line[0] == '>') {
            if (have) { if (gkmo_set_add(S, seq, sid) < 0) { added = -1; break; } added++; }
            have = 1; seq[0] = 0; seqlen = 0;
            char *tok = strtok(line, " \t\r\n");
            strncpy(sid, tok + 1, sizeof(sid) - 1); sid[sizeof(sid) - 1] = 0;
        } else if (seqlen < GKMO_MAX_SEQ - 1) {
            size_t ll = strlen(line);
            if ((size_t) seqlen + ll >= GKMO_MAX_SEQ) { ll = (size_t) (GKMO_MAX_SEQ - seqlen - 1); line[ll] = 0; }
            memcpy(seq + seqlen, line, ll + 1);
            seqlen += (int) ll;
        }
        if (c == EOF) break;
    }
    if (added >= 0 && have) { if (gkmo_set_add(S, seq, sid) < 0) added = -1; else added++; }
    free(line); free(seq); fclose(fp);
    return added;
}

/* gkmkernel_read_problems, libgkm.c:1316-1333: positives first, then negatives */
int gkmo_set_read_problem(gkmo_set *S, const char *posfile, const char *negfile)
{
    int np = gkmo_set_read_fasta(S, posfile);
    if (np < 0) return -1;
    int nn = gkmo_set_read_fasta(S, negfile);
    if (nn < 0) return -1;
    S->npos = np;
    return np + nn;
}

int gkmo_set_size(const gkmo_set *S) { return S->n; }
int gkmo_set_npos(const gkmo_set *S) { return S->npos; }
double gkmo_sqnorm(const gkmo_set *S, int i) { return S->s[i].sqnorm; }
int gkmo_seqlen(const gkmo_set *S, int i) { return S->s[i].len; }
void gkmo_get_weights(const gkmo_set *S, double *w) { for (int i = 0; i <= S->L; i++) w[i] = S->w[i]; }
void gkmo_get_poswt(const gkmo_set *S, int i, uint8_t *wt, uint8_t *wt_rc)
{
    memcpy(wt, S->s[i].wt, (size_t) S->s[i].nk);
    memcpy(wt_rc, S->s[i].wt_rc, (size_t) S->s[i].nk);
}
void gkmo_get_codes(const gkmo_set *S, int i, uint8_t *fwd, uint8_t *rc)
{
    memcpy(fwd, S->s[i].fwd, (size_t) S->s[i].len);
    memcpy(rc, S->s[i].rc, (size_t) S->s[i].len);
}

void gkmo_hist(const gkmo_set *S, int a, int b, int *H) { o_hist(S, &S->s[a], &S->s[b], H); }

/* normalise + optional RBF: libgkm.c:1169-1179 */
static double o_finish(const gkmo_set *S, double kraw, double sa, double sb)
{
    double v = kraw / (sa * sb);
    if (S->kernel_type == 3 || S->kernel_type == 5) v = exp(S->gamma * (v - 1));
    return v;
}

double gkmo_kernel(const gkmo_set *S, int a, int b)
{
    int H[GKMO_MAX_L + 1];
    o_hist(S, &S->s[a], &S->s[b], H);
    return o_finish(S, o_kraw(S, H), S->s[a].sqnorm, S->s[b].sqnorm);
}

/* ---- tiny pthread parallel-for (rows handed out by an atomic counter) ---- */
typedef struct {
    const gkmo_set *S; const int *rows; int nrows, ncols, tri; double *K; long ld; int *Hout; int next;
} o_job;

static void *o_worker(void *p)
{
    o_job *J = (o_job *) p;
    const gkmo_set *S = J->S;
    const int nb = S->d + 1;
    int H[GKMO_MAX_L + 1];
    for (;;) {
        int r = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
        if (r >= J->nrows) break;
        const int a = J->rows ? J->rows[r] : r;
        const int end = J->tri ? a : J->ncols;
        const long hstride = J->tri ? S->n : J->ncols;
        for (int j = 0; j < end; j++) {
            o_hist(S, &S->s[a], &S->s[j], H);
            J->K[(long) r * J->ld + j] = o_finish(S, o_kraw(S, H), S->s[a].sqnorm, S->s[j].sqnorm);
            if (J->Hout) for (int m = 0; m < nb; m++) J->Hout[((long) r * hstride + j) * nb + m] = H[m];
        }
        if (J->tri) J->K[(long) r * J->ld + a] = 1.0;
    }
    return 0;
}

static void o_run(o_job *J)
{
    long nt = sysconf(_SC_NPROCESSORS_ONLN);
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    pthread_t th[64];
    J->next = 0;
    for (long t = 1; t < nt; t++) pthread_create(&th[t], NULL, o_worker, J);
    o_worker(J);
    for (long t = 1; t < nt; t++) pthread_join(th[t], NULL);
}

/* what gkm_main_pywrapper leaves in kmat (gkmkern_pylib.c:169-221): K[a*ld + j] for j < a,
 * 1.0 on the diagonal, upper triangle untouched.  Hout (may be NULL): [(a*n + j)*(d+1) + m]. */
void gkmo_matrix_lower(const gkmo_set *S, double *K, long ld, int *Hout)
{
    o_job J = { S, NULL, S->n, 0, 1, K, ld, Hout, 0 };
    o_run(&J);
}

/* rectangular shape, gkmkernel_kernelfunc_batch_all(kernel, a, 0, ncols, res) (libgkm.c:1156):
 * K[r*ld + j] = K(rows[r], j), j in [0,ncols).  Hout: [(r*ncols + j)*(d+1) + m]. */
void gkmo_rect(const gkmo_set *S, const int *rows, int nrows, int ncols, double *K, long ld, int *Hout)
{
    o_job J = { S, rows, nrows, ncols, 0, K, ld, Hout, 0 };
    o_run(&J);
}
