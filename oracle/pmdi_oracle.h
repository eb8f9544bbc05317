/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (C++17, single thread) of the conditional-SMC allocation sweep of
 * nathancunn/ParticleMDI.jl, used ONLY by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the checker / baseline.  The product
 * path (libpmdi_cuda.so) never links, imports or calls anything in this directory.
 *
 * Parity status: the reference is Julia and Julia is not installed here or on the GPU box,
 * and the reference ships no golden vectors.  The oracle is therefore
 * pinned against the closed-form identities the reference's own tests assert
 * (test/runtests.jl:13-54, 136-162) re-evaluated with scipy, and against scipy closed forms
 * for the parts the reference never tests (NegBinom, calc_logmarginal, ESS, resampling):
 * for those parts parity is UNPINNED by the reference itself.
 */
#ifndef PMDI_ORACLE_H
#define PMDI_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { OR_GAUSSIAN = 0, OR_CATEGORICAL = 1, OR_NEGBINOM = 2 };

/* sweep mode bits */
enum {
  OR_MODE_DEDUP         = 1, /* reference data structures: copy-on-write cluster pool (src/pmdi.jl:131-146) */
  OR_MODE_LITERAL_NEWID = 2, /* keep the stale new_id cache of src/pmdi.jl:167; dedup only    */
  OR_MODE_SSTAR_COMPAT  = 4  /* pmdi() does not permute sstar on resample (src/pmdi.jl:324)     */
};

typedef struct or_ctx or_ctx;
typedef struct or_cluster or_cluster;

typedef struct {
  int32_t mode;
  int32_t _pad;
  const int64_t* s_in;       /* n_obs x K, column-major, labels 1..N (src/pmdi.jl:63)   */
  const int64_t* order_obs;  /* n_obs, 1-based permutation (src/pmdi.jl:172)            */
  int64_t n1;                /* floor(rho*n_obs) (src/pmdi.jl:161)                      */
  const double* Pi;          /* N x K column-major (src/pmdi.jl:179)                    */
  const double* phi;         /* K(K-1)/2, pair order of src/misc.jl:1-13                */
  double logweight_init;     /* 0.0 on the first iteration, 1.0 after (src/pmdi.jl:99,372) */
  uint64_t seed;
  uint32_t iter;
  uint32_t _pad2;
  const double* tape_alloc;   /* optional [steps][K][P]; NULL -> Philox */
  const double* tape_resamp;  /* optional [steps]                      */
  const double* tape_shuffle; /* optional [steps][P] (entry i-1 is the pick for position i) */
  const double* tape_select;  /* optional [1]                          */
  /* outputs */
  int64_t* s_out;        /* n_obs x K column-major                                  */
  int64_t* p_star;       /* 1-based                                                 */
  double*  logweight;    /* P, before the final reset (src/pmdi.jl:345)             */
  int64_t* n_ops;        /* number of calc_logprob calls (src/__pmdi.jl:187)        */
  int64_t* n_resamples;
  /* optional debug capture (NULL to skip) */
  double*  dbg_lp;       /* [steps][K][P][N] log predictive per particle/label      */
  double*  dbg_lw;       /* [steps][P] log-weights after coupling, before resample  */
  int32_t* dbg_alloc;    /* [steps][K][P] chosen label, 1-based                     */
  int32_t* dbg_anc;      /* [steps][P] ancestors, 1-based; all 0 when no resample   */
  int64_t* cluster_n;    /* [K][P][N] occupancy of each particle's clusters at end  */
} or_sweep_args;

or_ctx* or_create(int K, int n_obs, int N, int P);
void    or_destroy(or_ctx*);
/* data: column-major n_obs x D; f64 for OR_GAUSSIAN, int64 for the others */
int     or_set_dataset(or_ctx*, int k, int type, const void* data, int D);
int     or_set_flags(or_ctx*, int k, const uint8_t* flags);
int     or_sweep(or_ctx*, or_sweep_args*);

/* feature selection (src/pmdi.jl:120-128, 354-370) */
int     or_feature_null(or_ctx*, int k, double* out_D);
int     or_feature_select(or_ctx*, int k, const int64_t* labels_n, const double* feature_null,
                          uint64_t seed, uint32_t iter, const double* tape_f,
                          double* out_prob_D, uint8_t* out_flags_D);

/* single-cluster handles, for pinning the plugin maths against closed forms */
or_cluster* or_cl_new(or_ctx*, int k);
void        or_cl_free(or_cluster*);
void        or_cl_add(or_ctx*, int k, or_cluster*, int64_t obs_1based);
double      or_cl_logprob(or_ctx*, int k, or_cluster*, int64_t obs_1based);
void        or_cl_logmarginal(or_ctx*, int k, or_cluster*, double* out_D);
int64_t     or_cl_n(or_cluster*);
/* field: 0 mu, 1 sum, 2 lambda, 3 beta (Gaussian, D doubles); 4 counts (Lmax*D, as double);
   5 isum (NegBinom, D as double) */
int         or_cl_get(or_cluster*, int field, double* out);

/* helpers restated from src/misc.jl */
double  or_calc_ess(const double* logweight, int P);
void    or_draw_partstar(const double* logweight, int P, double r, const double* shuffle_u,
                         int64_t* partstar_out);
double  or_uniform_c(uint64_t seed, uint32_t iter, uint32_t kind, uint32_t step, uint32_t k,
                     uint32_t index);

#ifdef __cplusplus
}
#endif
#endif
