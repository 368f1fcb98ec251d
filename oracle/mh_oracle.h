/* oracle/mh_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C, CPU restatement of the reference MH engine (rplzzz/mcpar) used as the
 * parity checker for the B200 engine.  Every function cites the reference lines
 * it follows.  Parity of this restatement is PINNED against the reference itself:
 * tests/test_oracle.py runs the reference's own unmodified sources
 * (oracle/_ref, built by oracle/Makefile) on the same replay streams and demands
 * bit-identical traces; tests/golden/ holds fixtures generated that way for boxes
 * where oracle/_ref is absent.  (The reference ships no tests or golden vectors of
 * its own -- SURVEY.md section 4.)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 */
#ifndef MH_ORACLE_H_
#define MH_ORACLE_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_ROSENBROCK1 = 0, ORC_ROSENBROCK2 = 1, ORC_GAUSSIAN = 2, ORC_DUALGAUSSIAN = 3, ORC_GAUSSMIX = 4 };

typedef struct orc_config {
  int nparam;            /* d                                            */
  int nchain;            /* C: chains per rank (replay) / total N (counter mode) */
  int nranks;            /* R (replay mode; counter mode ignores it)     */
  int nsamp, nburn;
  int lik;               /* ORC_*                                        */
  int n_lik_par;
  double pl, armin, armax, dfac, ifac;   /* mcpar.hh:32-33 defaults 0.9,0.2,0.5,0.2,1.5 */
  int sync;              /* SYNCSTEP, default 10                         */
  int pinit_per_rank;    /* 0: every rank starts from the same C*d block */
  int trace_steps;       /* per-rank trace capacity                      */
  int trace_musig;
  /* counter (Philox) mode only */
  uint64_t seed;
  int coin_group;        /* chains sharing one local/remote coin (power of two <= 32); 0 = whole job */
  int pool_m;            /* remote-mixture pool size, 0 => all chains    */
  int thin;              /* keep every thin-th main step in rows         */
  int remote_mode;       /* 0: the reference's max-mixture rejection loop (mcpar.cc:315-451);
                            1: sum-mixture independence proposal with NORMALISED components and
                               no rejection loop (Murray 2010)            */
  int pool_lag;          /* 0: window w reads the pool published at the end of window w-1;
                            1: one window older (takes the exchange off the critical path)       */
} orc_config;

/* ---- primitives ---- */
void   orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_u53(uint32_t hi, uint32_t lo);                 /* [0,1) with 53 bits */
int    orc_cholesky_lower(int d, double *a);              /* mcpar.cc:454-484 (spotrf 'U' col-major == row-major lower) */
void   orc_covar_setup(int d, const double *incov, double *cov);
int    orc_loglik(int lik, int d, const double *par, int npset, const double *x, double *y);
int    orc_sobol_points(int dimen, uint64_t first_scalar, size_t nscalar, double *out); /* mcutil.cc:16-25 */
void   orc_qriguess(int rank, int npset, int d, const double *plo, const double *phi, double *pout);

/* ---- the reference engine on R ranks, REPLAY streams (verification mode) ----
 * Layouts are those of oracle/ref_harness.cc::ref_run (same argument meaning).
 * extra: tr_accept [R][T][C] (0/1), tr_remote [R][T] (0/1), tr_iters [R][T].  */
int orc_run_replay(const orc_config *cfg, const double *pinit, const double *incov, const double *par,
                   const double *Z, size_t nz, const double *U, size_t nu, const int *I, size_t ni,
                   double *rows, double *st_p, double *st_ly, double *st_mu, double *st_sig,
                   double *st_psum2, double *st_cov, double *st_musig, long long *used, double *maxl,
                   double *tr_pre_p, double *tr_pre_ly, double *tr_trial_p, double *tr_trial_ly,
                   double *tr_cfac, double *tr_cov, double *tr_musig, long long *tr_cursors,
                   int *tr_accept, int *tr_remote, int *tr_iters);

/* ---- the same algorithm with counter-based Philox draws per global chain ----
 * (the B200 engine's normal mode; see DESIGN.md "normal-mode semantics").
 *   pinit [N][d]; rows [ceil(nsamp/thin)][N][d+1]; st_* [N][...]; pool_out [M][d][2]
 *   acc_counts[2] = accepted, tried over the main phase; cov_out [d][d] = tuned factor;
 *   tr_accept [nburn+nsamp][N] or NULL. */
int orc_run_counter(const orc_config *cfg, const double *pinit, const double *incov, const double *par,
                    double *rows, double *st_p, double *st_ly, double *st_mu, double *st_psum2,
                    double *pool_out, long long *acc_counts, double *cov_out,
                    unsigned char *tr_accept, long long *remote_iters);

#ifdef __cplusplus
}
#endif
#endif
