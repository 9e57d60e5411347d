/* gkm_weights.c -- host-side scalars of the path: parameter gate, w[m], positional weights.
 *
 * w[m] (the factor each mismatch bin is multiplied with) is computed once per
 * problem in double precision.  To reproduce the reference's doubles bit for bit
 * the floating-point operations below are issued in the same order as
 * libgkm.c:107-217 (binomials and integer powers are exact in double, so a lookup
 * table is interchangeable with the reference's Pascal recurrences).
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>

#include "gkm_internal.h"

/* ---- parameter gate: gkmkern_pylib.c:38-64, with the L ceiling configurable ---- */
const char *gkm_param_problem(const gkm_parameter *p, int max_L)
{
    if (p->kernel_type < GKM || p->kernel_type > EST_TRUNC_PW_RBF) return "unknown kernel type";
    if (p->L < 2) return "L < 2";
    if (p->L > max_L) return (max_L == 12) ? "L > 12" : "L > 16";
    if (p->k > p->L) return "k > L";
    if (p->d > (p->L - p->k)) return "d > L - k";
    if (p->k < 0) return "k < 0";
    if (p->d < 0) return "d < 0";
    if (p->d >= GKM_MAX_BINS) return "d > 15";
    return NULL;
}

/* exact binomials up to n = 2*GKM_MAX_L */
#define NB_MAX (2 * GKM_MAX_L + 2)
static double g_binom[NB_MAX][NB_MAX];
static int g_binom_ready = 0;

static void binom_init(void)
{
    if (g_binom_ready) return;
    for (int n = 0; n < NB_MAX; n++) {
        for (int r = 0; r < NB_MAX; r++) g_binom[n][r] = 0.0;
        g_binom[n][0] = 1.0;
        for (int r = 1; r <= n; r++) g_binom[n][r] = g_binom[n - 1][r - 1] + (r <= n - 1 ? g_binom[n - 1][r] : 0.0);
    }
    g_binom_ready = 1;
}

static double binom(int n, int r)
{
    if (r < 0 || n < 0 || n < r) return 0.0; /* the call sites below never pass n < 0 */
    return g_binom[n][r];
}

static double ipow(int base, int e) /* exact for the ranges used: 4^16, 3^16, 2^16 */
{
    double v = 1.0;
    while (e-- > 0) v *= (double) base;
    return v;
}

/* gapped k-mer counting weights, type 0 (libgkm.c:204-217) */
static void weights_plain(int L, int K, double *w)
{
    for (int m = 0; m + K <= L; m++) w[m] = binom(L - m, K);
}

/* estimated l-mer weights, full or truncated filter, types 1..5 (libgkm.c:107-202) */
static void weights_estimated(int L, int K, int truncate, double *w)
{
    enum { S = GKM_MAX_L + 1 };
    const int alpha = 4;
    double bufA[S * S], bufB[S * S];
    double *nxt = bufA, *cur = bufB;
    double wm[S], filt[S], filt_tr[S];

    for (int i = 0; i < S * S; i++) bufA[i] = bufB[i] = 1.0;

    /* stage 1 (:133-143): two-array recurrence over the word length */
    for (int len = 1; len <= L; len++) {
        for (int kk = 1; kk <= K; kk++) {
            nxt[kk * S + 0] = cur[kk * S + 0] + (alpha - 1) * cur[(kk - 1) * S + 0];
            for (int j = 1; j <= kk; j++) nxt[kk * S + j] = (nxt[(kk - 1) * S + (j - 1)] * (kk - len)) / kk;
        }
        double *t = cur; cur = nxt; nxt = t;
    }
    const double norm = binom(L, K) * ipow(alpha, L); /* :145 */
    for (int i = 0; i <= K; i++) wm[i] = cur[K * S + i] / norm;

    /* stage 2 (:152-168): the mismatch filter and its truncation at the first tiny tap */
    for (int m = 0; m <= L; m++) {
        const int top = (m < K) ? m : K;
        double s = 0;
        for (int i = 0; i <= top; i++) s += wm[i] * binom(L - m, K - i) * binom(m, i);
        filt[m] = s;
    }
    int alive = 1;
    for (int m = 0; m <= L; m++) {
        if (filt[m] < 1e-50) alive = 0;
        filt_tr[m] = alive ? filt[m] : 0.0;
    }
    const double *f = truncate ? filt_tr : filt;

    /* stage 3 (:171-191): self-convolution of the filter over l-mer space */
    for (int m = 0; m <= L; m++) {
        double s = 0;
        for (int m1 = 0; m1 <= L; m1++)
            for (int m2 = 0; m2 <= L; m2++)
                for (int t = 0; t <= L; t++) {
                    const int r = m1 + m2 - 2 * t - L + m;
                    if (t > m || (m1 - t) > (L - m) || r > (m1 - t) || r < 0) continue;
                    const double cc = binom(m, t) * binom(L - m, m1 - t) * binom(m1 - t, r) *
                                      ipow(alpha - 1, t) * ipow(alpha - 2, r);
                    s += cc * f[m1] * f[m2];
                }
        w[L - m] = s;
    }
}

int gkm_calc_weights(int kernel_type, int L, int k, double *w)
{
    if (L < 1 || L > GKM_MAX_L || k < 0 || k > L) return 1;
    binom_init();
    for (int m = 0; m <= L; m++) w[m] = 0.0;
    if (kernel_type == GKM) weights_plain(L, k, w);
    else weights_estimated(L, k, kernel_type != EST_FULL, w); /* dispatch libgkm.c:997-1019 */
    return 0;
}

/* positional weights per L-mer start (libgkm.c:910-932): exponential decay from the centre
 * for the wgkm types, 1 otherwise; the reverse-complement strand gets the mirrored vector */
/* weight of the L-mer `dist` starts away from the centre one (libgkm.c:922-925, the reference's own expression) */
static uint8_t posweight_at(int dist, int M, double H)
{
    const double x = floor(M * exp((-1) * log(2) * dist / H) + 1);
    uint8_t v = (uint8_t) (int) x; /* the reference stores into u_int8_t: 256 wraps to 0 */
    if (v > M) v = (uint8_t) M;
    return v;
}

void gkm_calc_posweights(int nk, int kernel_type, int M, double H, uint8_t *wt, uint8_t *wt_rc)
{
    const int centre = nk / 2;
    const int decays = (kernel_type == EST_TRUNC_PW || kernel_type == EST_TRUNC_PW_RBF);
    for (int i = 0; i < nk; i++) {
        const uint8_t v = decays ? posweight_at(abs(centre - i), M, H) : 1;
        wt[i] = v;
        wt_rc[nk - 1 - i] = v;
    }
}

/* The weight depends on the distance from the centre only: one table serves every sequence length.  The device
 * packer (gkm_device.cu: gkm_pack_kernel) indexes it; exp() never runs on the GPU, so the bytes are the host's. */
void gkm_posweight_table(int kernel_type, int M, double H, uint8_t *tab)
{
    const int decays = (kernel_type == EST_TRUNC_PW || kernel_type == EST_TRUNC_PW_RBF);
    for (int dist = 0; dist <= GKM_MAX_BASES; dist++) tab[dist] = decays ? posweight_at(dist, M, H) : 1;
}

const char *gkmb200_check_parameter(const gkm_parameter *param) { return gkm_param_problem(param, 12); }
int gkmb200_weights(int kernel_type, int L, int k, double *w) { return gkm_calc_weights(kernel_type, L, k, w); }
int gkmb200_posweights(int nk, int kernel_type, int M, double H, uint8_t *wt, uint8_t *wt_rc)
{
    if (nk < 0) return 1;
    gkm_calc_posweights(nk, kernel_type, M, H, wt, wt_rc);
    return 0;
}
