/* gkm_abi.h -- the C-ABI of the gkm kernel engine, binary compatible with the
 * reference's src/libgkm.h (Dongwon-Lee/gkmQC).
 *
 * This header is written from the ABI facts, not from the reference text:
 * struct layouts, sizes and offsets were probed on x86-64 and
 * are pinned by the static assertions at the bottom; the thirteen prototypes are
 * the ones libgkm.h:132-147,164 declares.  A caller compiled against the
 * reference's own libgkm.h links against gkmkern_pylib.so of this repo unchanged,
 * and scripts/gkmsvm.py:48-61,85-88 binds to it unchanged.
 *
 * What is different behind the ABI: there is no k-mer tree.  The three tree
 * pointers of gkm_kernel are kept for layout only; `prob_kmertree` carries an
 * opaque handle of the device-resident packed sequence set and `kmertree` is NULL.
 */
#ifndef GKM_ABI_H_INCLUDED
#define GKM_ABI_H_INCLUDED

#include <stdint.h>
#include <sys/types.h> /* u_int8_t, like the reference's users get it */

/* the reference header's include guard: whoever includes both gets one set of types */
#ifndef LIBSVM_GKM_H_INCLUDED
#define LIBSVM_GKM_H_INCLUDED

#ifdef __cplusplus
extern "C" {
#endif

/* limits (libgkm.h:29-34) */
#define MAX_ALPHABET_SIZE 4
#define MAX_ALPHABET_SIZE_SQ 16
#define MAX_MM 12          /* weights[] has MAX_MM+1 slots: w[0..12] */
#define MAX_SEQ_LENGTH 2048 /* at most 2047 bases of a record are used */
#define MMCNT_LOOKUPTAB_WIDTH 8
#define LOGGER_ID 0
#define LOGGER_FORMAT "%l %d %t: %m\n"

/* kernel_type (libgkm.h:51) */
enum { GKM = 0, EST_FULL = 1, EST_TRUNC = 2, EST_TRUNC_RBF = 3, EST_TRUNC_PW = 4, EST_TRUNC_PW_RBF = 5 };

typedef struct _KmerTree KmerTree;
typedef struct _KmerTreeLeaf KmerTreeLeaf;
typedef struct _KmerTreeLeafData KmerTreeLeafData;
typedef struct _KmerTreeCoef KmerTreeCoef;
typedef struct _NodeMismatchCount NodeMismatchCount;
typedef struct _gkm_parameter gkm_parameter;
typedef struct _gkm_data gkm_data;
typedef struct _svm_problem svm_problem;
typedef struct _gkm_kernel gkm_kernel;
typedef struct _gkmOpt gkmOpt;

struct _gkm_parameter { /* 48 bytes */
    int kernel_type;
    int L;        /* word length */
    int k;        /* informative columns; enters w[m] only */
    int d;        /* max mismatches kept */
    u_int8_t M;   /* wgkm: peak positional weight */
    double H;     /* wgkm: half-life in positions */
    double gamma; /* RBF types */
    int nthreads;
};

struct _gkm_data { /* 88 bytes */
    char *sid;
    int seqid;
    int label;
    int seqlen;
    u_int8_t *seq;     /* base codes A,C,G,T = 1,2,3,4 */
    u_int8_t *seq_rc;  /* reverse complement, same coding */
    u_int8_t *wt;      /* positional weight per L-mer start, forward strand */
    u_int8_t *wt_rc;   /* same for the reverse-complement strand */
    int *kmerids;
    int *kmerids_rc;
    char *seq_string;
    double sqnorm;     /* sqrt(Kraw(x,x)) */
};

struct _svm_problem { /* 24 bytes */
    int l;
    double *y;
    gkm_data **x;
};

struct _gkm_kernel { /* 176 bytes */
    gkm_parameter *param;
    double weights[MAX_MM + 1];
    KmerTree *kmertree;      /* reference: dynamic tree.   here: NULL */
    KmerTree *prob_kmertree; /* reference: static tree.    here: opaque device problem handle */
    gkm_data **prob_svm_data;
    int prob_num;
    int *prob_gkmkernel_index;
    int *prob_libsvm_index;
    u_int8_t *mmcnt_lookuptab; /* reference: 64 KiB XOR table. here: NULL */
    int mmcnt_lookuptab_mask;
    int mmcnt_nlookups;
};

struct _KmerTreeLeafData { int seqid; int wt; };
struct _KmerTreeLeaf { int count; int capacity; KmerTreeLeafData *data; };
struct _KmerTree { int L; int k; int d; int node_count; int leaf_count; int *node; KmerTreeLeaf *leaf; };
struct _KmerTreeCoef { int depth; int node_count; int leaf_count; double *coef_sum; };

struct _gkmOpt { /* 64 bytes; mirrored by ctypes in scripts/gkmsvm.py:48-61 */
    int kernel_type;
    int L;
    int k;
    int d;
    u_int8_t M;
    double H;
    double gamma;
    char *posfile;
    char *negfile;
    int nthreads;  /* reference: row threads.  here: host copy-out threads (the rows run on the GPUs) */
    int verbosity; /* 0..4 = ERROR, WARN, INFO, DEBUG, TRACE */
};

/* ---- lifecycle (libgkm.h:132-134; libgkm.c:978, :1058, :1187) ---- */
gkm_kernel *gkmkernel_init(gkm_parameter *param);
void gkmkernel_destroy(gkm_kernel *kernel);
void gkmkernel_set_num_threads(gkm_parameter *param);

/* ---- sequence objects (libgkm.h:136-138; libgkm.c:841, :941, :956) ---- */
gkm_data *gkmkernel_new_object(gkm_kernel *kernel, char *seq, char *sid, int seqid);
void gkmkernel_delete_object(gkm_data *d);
void gkmkernel_free_object(gkm_data *d);

/* ---- kernel evaluation (libgkm.h:140-141,147; libgkm.c:1115, :1156) ---- */
double gkmkernel_kernelfunc(const gkm_data *da, const gkm_data *db); /* declared, never defined upstream */
double *gkmkernel_kernelfunc_batch(gkm_kernel *kernel, int a, const gkm_data **db_array, const int n, double *res);
double *gkmkernel_kernelfunc_batch_all(gkm_kernel *kernel, const int a, const int start, const int end, double *res);

/* ---- problem set-up (libgkm.h:143-146; libgkm.c:1035, :1316, :1071, :1084) ---- */
void gkmkernel_build_tree(gkm_kernel *kernel, gkm_data **x, int n);
int gkmkernel_read_problems(gkm_kernel *kernel, svm_problem *prob, const char *posfile, const char *negfile);
void gkmkernel_swap_index(gkm_kernel *kernel, int i, int j);
void gkmkernel_update_index(gkm_kernel *kernel);

/* ---- the operator gkmQC calls (libgkm.h:164; gkmkern_pylib.c:92) ----
 * kmat: >= N row pointers, each to >= N doubles.  On return rows 0..N-1 hold the
 * strict lower triangle K(a,j), j<a, and 1.0 on the diagonal; everything else is
 * untouched.  kmat_size = {n_pos, n_neg}.  Returns 0, or 1 on error. */
int gkm_main_pywrapper(gkmOpt *opts, double **kmat, int *kmat_size);

#ifdef __cplusplus
}
#endif

#endif /* LIBSVM_GKM_H_INCLUDED */

/* layout pins (x86-64 SysV), SURVEY.md 8(a) a14 / 8(b) */
#if defined(__x86_64__) && !defined(GKM_ABI_NO_STATIC_ASSERT)
#include <stddef.h>
#ifdef __cplusplus
#define GKM_SA(c, m) static_assert(c, m)
#else
#define GKM_SA(c, m) _Static_assert(c, m)
#endif
GKM_SA(sizeof(struct _gkmOpt) == 64, "gkmOpt size");
GKM_SA(offsetof(struct _gkmOpt, M) == 16 && offsetof(struct _gkmOpt, H) == 24, "gkmOpt M/H");
GKM_SA(offsetof(struct _gkmOpt, posfile) == 40 && offsetof(struct _gkmOpt, nthreads) == 56, "gkmOpt tail");
GKM_SA(sizeof(struct _gkm_parameter) == 48, "gkm_parameter size");
GKM_SA(sizeof(struct _gkm_data) == 88 && offsetof(struct _gkm_data, sqnorm) == 80, "gkm_data");
GKM_SA(sizeof(struct _svm_problem) == 24, "svm_problem size");
GKM_SA(sizeof(struct _gkm_kernel) == 176, "gkm_kernel size");
GKM_SA(offsetof(struct _gkm_kernel, weights) == 8 && offsetof(struct _gkm_kernel, kmertree) == 112, "gkm_kernel head");
GKM_SA(offsetof(struct _gkm_kernel, prob_svm_data) == 128 && offsetof(struct _gkm_kernel, prob_num) == 136, "gkm_kernel prob");
GKM_SA(sizeof(struct _KmerTree) == 40, "KmerTree size");
#undef GKM_SA
#endif

#endif /* GKM_ABI_H_INCLUDED */
