/* oracle/mh_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the rplzzz/mcpar Metropolis-Hastings engine, in fp64,
 * operation by operation (no FMA contraction: build with -ffp-contract=off).
 * It is the CPU checker for the B200 engine; nothing in the product links it.
 * Pinned against the reference's own sources by tests/test_oracle.py and
 * the fixtures in tests/golden/.  Citations are file:line into /root/reference/.
 */
#include "mh_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define FPEPS 1.0e-14            /* src/mcpar.cc:15 */

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11) -- the counter RNG  */
/* of the B200 engine's normal mode.  Not part of the reference (its    */
/* MT2203/BoxMuller2 live inside MKL, unpinned); KATs in tests/.        */
/* ------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double orc_u53(uint32_t hi, uint32_t lo)
{
  /* top 52 bits -> mantissa of a double in [1,2), minus 1: uniform on [0,1) with 52 bits */
  uint64_t b = ((uint64_t)hi << 32) | lo;
  uint64_t m = 0x3FF0000000000000ull | (b >> 12);
  double x; memcpy(&x, &m, 8);
  return x - 1.0;
}

/* ------------------------------------------------------------------ */
/* covar_setup: src/mcpar.cc:454-484                                   */
/* ------------------------------------------------------------------ */
/* spotrf('U') on the column-major view == Cholesky LOWER factor of the row-major
 * view, in place; entries above the diagonal keep the input covariance. */
int orc_cholesky_lower(int d, double *a)
{
  for (int i = 0; i < d; ++i)
    for (int j = 0; j <= i; ++j) {
      double sum = a[i*d + j];
      for (int k = 0; k < j; ++k) sum -= a[i*d + k] * a[j*d + k];
      if (i == j) { if (!(sum > 0)) return i + 1; a[i*d + i] = sqrt(sum); }
      else a[i*d + j] = sum / a[j*d + j];
    }
  return 0;
}

void orc_covar_setup(int d, const double *incov, double *cov)
{
  if (incov) for (int i = 0; i < d*d; ++i) cov[i] = incov[i];        /* mcpar.cc:457-459 */
  else {                                                              /* :460-467 */
    for (int i = 0; i < d*d; ++i) cov[i] = 0.0;
    for (int i = 0; i < d; ++i) cov[i*(d+1)] = 1.0;
  }
  (void)orc_cholesky_lower(d, cov);                                   /* :480, info unchecked */
}

/* ------------------------------------------------------------------ */
/* likelihoods: src/rosenbrock.cc                                      */
/* ------------------------------------------------------------------ */
int orc_loglik(int lik, int d, const double *par, int npset, const double *x, double *y)
{
  const int ntot = npset * d;
  switch (lik) {
  case ORC_ROSENBROCK1:                       /* rosenbrock.cc:4-21 */
    if (d < 2 || d % 2) return -2;            /* rosenbrock.hh:13-16 */
    for (int j = 0; j < npset; ++j) y[j] = 0.0;
    for (int i = 0; i < ntot - 1; i += 2) {
      int j = i / d;
      double t1 = 1 - x[i];
      double t2 = x[i+1] - x[i]*x[i];
      y[j] -= t1*t1 + 100.0*t2*t2;
    }
    return 0;
  case ORC_ROSENBROCK2:                       /* rosenbrock.cc:25-41: flat loop, so the last
                                                 parameter of set j pairs with the first of set
                                                 j+1; sign is t1^2 - 100 t2^2 */
    if (d < 2) return -2;
    for (int j = 0; j < npset; ++j) y[j] = 0.0;
    for (int i = 0; i < ntot - 1; ++i) {
      int j = i / d;
      double t1 = 1 - x[i];
      double t2 = x[i+1] - x[i]*x[i];
      y[j] -= t1*t1 - 100.0*t2*t2;
    }
    return 0;
  case ORC_GAUSSIAN: {                        /* rosenbrock.cc:44-61, ctor rosenbrock.hh:42-48 */
    if (d != 2) return -2;
    double mu[2], s2i[2];
    for (int k = 0; k < 2; ++k) { mu[k] = par ? par[k] : 0.0; s2i[k] = par ? 1.0 / par[2+k] : 1.0; }
    for (int j = 0; j < npset; ++j) y[j] = 0.0;
    for (int i = 0; i < ntot; ++i) {
      int j = i / d, k = i % d;
      double arg = x[i] - mu[k];
      y[j] -= 0.5*arg*arg*s2i[k];
    }
    return 0;
  }
  case ORC_DUALGAUSSIAN: {                    /* rosenbrock.cc:63-78 (no log-sum-exp guard) */
    const double w = par ? par[0] : 5.0;
    for (int j = 0; j < npset; ++j) {
      int ix = j * 2;
      double arg1 = 0.5*(x[ix]*x[ix] + x[ix+1]*x[ix+1]);
      double t2a = x[ix] - 5.0, t2b = x[ix+1] - 5.0;
      double arg2 = 0.5*(t2a*t2a + t2b*t2b);
      y[j] = log(w*exp(-arg1) + exp(-arg2));
    }
    return 0;
  }
  case ORC_GAUSSMIX: {                        /* NEW (not in the reference):
                                                 log sum_k w_k exp(-1/2 sum_i (x_i-mu_ki)^2/s2_ki),
                                                 DualGaussian's form generalised, WITH log-sum-exp.
                                                 par = [K, mu[K][d], s2[K][d], w[K]] */
    const int K = (int)par[0];
    const double *mu = par + 1, *s2 = mu + (size_t)K*d, *w = s2 + (size_t)K*d;
    double *a = (double*)malloc(sizeof(double) * (size_t)K);
    for (int j = 0; j < npset; ++j) {
      double m = -INFINITY;
      for (int k = 0; k < K; ++k) {
        double s = 0.0;
        for (int i = 0; i < d; ++i) { double xm = x[(size_t)j*d+i] - mu[(size_t)k*d+i]; s += xm*xm / s2[(size_t)k*d+i]; }
        a[k] = log(w[k]) - 0.5*s;
        if (a[k] > m) m = a[k];
      }
      double sum = 0.0;
      for (int k = 0; k < K; ++k) sum += exp(a[k] - m);
      y[j] = m + log(sum);
    }
    free(a);
    return 0;
  }
  }
  return -1;
}

/* ------------------------------------------------------------------ */
/* Sobol + qriguess: src/mcutil.cc:3-34.  MKL's direction numbers are   */
/* unobtainable (unpinned); Joe-Kuo (2008) numbers for dims 1..64,      */
/* gray-code order, first point = origin, value = integer * 2^-32.      */
/* ------------------------------------------------------------------ */
#include "sobol_jk64.h"     /* generated by tools/gen_sobol_table.py from the public Joe-Kuo file */

static void sobol_dirs(int dim, uint32_t v[32])
{
  if (dim == 0) { for (int i = 0; i < 32; ++i) v[i] = 1u << (31 - i); return; }
  const sobol_jk_t *p = &SOBOL_JK[dim]; const int s = p->s;
  for (int i = 0; i < 32; ++i) {
    if (i < s) v[i] = p->m[i] << (31 - i);
    else {
      v[i] = v[i-s] ^ (v[i-s] >> s);
      for (int k = 1; k < s; ++k) v[i] ^= (((p->a >> (s-1-k)) & 1u) * v[i-k]);
    }
  }
}

/* scalars [first_scalar, first_scalar+nscalar) of the point-major interleaved stream
 * (what vslSkipAheadStream + vsRngUniform deliver, mcutil.cc:23-25) */
int orc_sobol_points(int dimen, uint64_t first_scalar, size_t nscalar, double *out)
{
  if (dimen < 1 || dimen > SOBOL_JK_NDIM) return -1;
  uint32_t (*v)[32] = (uint32_t (*)[32])malloc(sizeof(uint32_t) * 32 * (size_t)dimen);
  for (int k = 0; k < dimen; ++k) sobol_dirs(k, v[k]);
  for (size_t q = 0; q < nscalar; ++q) {
    uint64_t sc = first_scalar + q;
    uint64_t n = sc / (uint64_t)dimen; int k = (int)(sc % (uint64_t)dimen);
    uint64_t gray = n ^ (n >> 1);            /* x_n = XOR of v[b] over set bits b of gray(n) */
    uint32_t xv = 0;
    for (int b = 0; gray; ++b, gray >>= 1) if (gray & 1ull) xv ^= v[k][b];
    out[q] = (double)xv * (1.0 / 4294967296.0);
  }
  free(v);
  return 0;
}

void orc_qriguess(int rank, int npset, int d, const double *plo, const double *phi, double *pout)
{
  const size_t ntot = (size_t)npset * (size_t)d;                /* mcutil.cc:19 (int there; size_t: 2^20 x 64 fits, 2^26 x 64 would not) */
  double *q = (double*)malloc(sizeof(double) * ntot);
  orc_sobol_points(d, rank > 0 ? (uint64_t)rank * (uint64_t)ntot : 0, ntot, q);  /* :22-25 */
  for (size_t j = 0; j < (size_t)npset; ++j)                     /* :28-32 */
    for (int i = 0; i < d; ++i) { size_t ix = j*d + i; pout[ix] = plo[i] + q[ix]*(phi[i]-plo[i]); }
  free(q);
}

/* ------------------------------------------------------------------ */
/* shared step pieces                                                  */
/* ------------------------------------------------------------------ */
/* un-normalised diagonal Gaussian Q_qi(x) of mcpar.cc:367-387 / :424-436;
 * ms = (mu,sig2) pairs of component qi, interleaved */
static double q_value(int d, const double *ms, const double *x)
{
  double arg = 0.0;
  for (int i = 0; i < d; ++i) {
    double xm = ms[2*i] - x[i];
    double sig2 = ms[2*i+1];
    arg += xm*xm/sig2;
  }
  return exp(-0.5*arg);
}

/* log of the NORMALISED sum-mixture density of the pool at x (up to the constant
 * -log M - d/2 log 2 pi that cancels in the Hastings ratio):
 *   L(x) = log sum_s exp(n_s - 1/2 sum_i (mu_si - x_i)^2 / sig2_si),  n_s = -1/2 sum_i log sig2_si,
 * by log-sum-exp.  Remote mode 1 only; the reference's Q_i (mcpar.cc:367-390) lacks n_s and is
 * combined by max, not sum. */
static double pool_lse(int d, int M, const double *pool, const double *x)
{
  double m = -INFINITY;
  for (int s = 0; s < M; ++s) {
    const double *ms = pool + (size_t)s*d*2;
    double arg = 0.0, n = 0.0;
    for (int i = 0; i < d; ++i) { double xm = ms[2*i] - x[i]; arg += xm*xm/ms[2*i+1]; n += log(ms[2*i+1]); }
    double a = -0.5*n - 0.5*arg;
    if (a > m) m = a;
  }
  if (!(m > -INFINITY)) return m;              /* every term underflowed to -inf (or NaN input) */
  double sum = 0.0;
  for (int s = 0; s < M; ++s) {
    const double *ms = pool + (size_t)s*d*2;
    double arg = 0.0, n = 0.0;
    for (int i = 0; i < d; ++i) { double xm = ms[2*i] - x[i]; arg += xm*xm/ms[2*i+1]; n += log(ms[2*i+1]); }
    sum += exp(-0.5*n - 0.5*arg - m);
  }
  return m + log(sum);
}

/* Welford update with remote adoption, mcpar.cc:186-209, for one parameter */
static void welford(double p, double pwgt, double winv, int adopt, double mut, double sigt,
                    double *mu, double *psum2, double *sig)
{
  if (adopt) { *mu = mut; *sig = sigt; *psum2 = (*sig)*(pwgt - 1.0); }   /* :190-197 */
  double delta = p - *mu;                                                  /* :199 */
  *mu += delta * winv;                                                     /* :200 */
  *psum2 += delta*(p - *mu);                                               /* :201 */
  *sig = *psum2 * winv;                                                    /* :202 */
}

/* ------------------------------------------------------------------ */
/* REPLAY engine: the reference on R ranks with supplied streams        */
/* ------------------------------------------------------------------ */
typedef struct {
  double *pvals, *ptrial, *mu, *sig, *mutrial, *sigtrial, *psum2, *musigall;
  double *lylast, *lytrial, *cfac, *pacpt, *acpt, *qisum, *qimax, *cov;
  int *rjct, *chnsel;
  double ntrial, naccept; int irate;
  const double *Z, *U; const int *I; size_t nz, nu, ni, iz, iu, ii; int overrun;
  double maxlval; double *maxlparams;
  int iters_last;
} rank_t;

static double take_z(rank_t *k) { if (k->iz >= k->nz) { k->overrun = 1; ++k->iz; return 0.0; } return k->Z[k->iz++]; }
static double take_u(rank_t *k) { if (k->iu >= k->nu) { k->overrun = 1; ++k->iu; return 0.0; } return k->U[k->iu++]; }
static int    take_i(rank_t *k) { if (k->ii >= k->ni) { k->overrun = 1; ++k->ii; return 0; } return k->I[k->ii++]; }

/* MCPar::genLocal, mcpar.cc:302-312; transform of vsRngGaussianMV(FULL): row-major
 * lower factor, r_i = a_i + sum_{k<=i} T[i*d+k] z_k accumulated left to right */
static void gen_local(rank_t *k, int C, int d)
{
  double z[256];
  for (int j = 0; j < C; ++j) {
    for (int i = 0; i < d; ++i) z[i] = take_z(k);
    for (int i = 0; i < d; ++i) {
      double acc = k->pvals[j*d+i];
      for (int q = 0; q <= i; ++q) acc += k->cov[i*d+q] * z[q];
      k->ptrial[j*d+i] = acc;
    }
    k->cfac[j] = 1.0;
  }
}

/* MCPar::genRemote, mcpar.cc:315-451: lock-step rejection sampling of max_i Q_i */
static void gen_remote(rank_t *k, int C, int d, int N)
{
  for (int j = 0; j < C; ++j) k->rjct[j] = 1;                        /* :329-330 */
  int anyrjct = 1, iters = 0;
  do {
    ++iters;
    for (int j = 0; j < C; ++j) k->chnsel[j] = take_i(k);             /* :337 ints for ALL chains */
    for (int j = 0; j < C; ++j) if (k->rjct[j]) {                     /* :339-352 */
      for (int i = 0; i < d; ++i) {
        int t = 2*(d*k->chnsel[j] + i);
        k->mutrial[j*d+i]  = k->musigall[t];
        k->sigtrial[j*d+i] = sqrt(k->musigall[t+1]);
      }
      for (int i = 0; i < d; ++i) {                                  /* DIAGONAL storage */
        double z = take_z(k);
        k->ptrial[j*d+i] = k->mutrial[j*d+i] + k->sigtrial[j*d+i]*z;
      }
    }
    for (int j = 0; j < C; ++j) {                                     /* :355-365 */
      if (k->rjct[j]) { k->qimax[j] = FPEPS; k->qisum[j] = FPEPS; }
      else            { k->qimax[j] = 0.0;   k->qisum[j] = 1.0; }
    }
    for (int qi = 0; qi < N; ++qi)                                    /* :367-395 all pairs */
      for (int j = 0; j < C; ++j) {
        double gv = q_value(d, k->musigall + 2*(size_t)qi*d, k->ptrial + j*d);
        if (k->rjct[j]) { k->qisum[j] += gv; k->qimax[j] = gv > k->qimax[j] ? gv : k->qimax[j]; }
      }
    for (int j = 0; j < C; ++j) k->pacpt[j] = k->qimax[j] / k->qisum[j];   /* :397-398 */
    for (int j = 0; j < C; ++j) k->acpt[j] = take_u(k);                    /* :401 */
    anyrjct = 0;
    for (int j = 0; j < C; ++j) {                                          /* :405-442 */
      if (k->acpt[j] < k->pacpt[j]) {
        k->rjct[j] = 0;
        k->cfac[j] = 0.0;
        for (int qi = 0; qi < N; ++qi) {
          double gv = q_value(d, k->musigall + 2*(size_t)qi*d, k->pvals + j*d);
          k->cfac[j] = gv > k->cfac[j] ? gv : k->cfac[j];
        }
        k->cfac[j] /= k->qimax[j];
      }
      anyrjct += k->rjct[j];
    }
  } while (anyrjct);
  for (int i = 0; i < C*d; ++i) k->sigtrial[i] *= k->sigtrial[i];           /* :447-448 */
  k->iters_last = iters;
}

int orc_run_replay(const orc_config *cfg, const double *pinit, const double *incov, const double *par,
                   const double *Z, size_t nz, const double *U, size_t nu, const int *I, size_t ni,
                   double *rows, double *st_p, double *st_ly, double *st_mu, double *st_sig,
                   double *st_psum2, double *st_cov, double *st_musig, long long *used, double *maxl,
                   double *tr_pre_p, double *tr_pre_ly, double *tr_trial_p, double *tr_trial_ly,
                   double *tr_cfac, double *tr_cov, double *tr_musig, long long *tr_cursors,
                   int *tr_accept, int *tr_remote, int *tr_iters)
{
  const int R = cfg->nranks, C = cfg->nchain, d = cfg->nparam, N = R*C;
  const int nt = C*d, nc = d*d, T = cfg->trace_steps;
  const size_t nm = (size_t)2*N*d;
  if (d > 256) return -1;
  rank_t *rk = (rank_t*)calloc((size_t)R, sizeof(rank_t));
  int rc = 0;
  for (int r = 0; r < R; ++r) {                                   /* ctor, mcpar.cc:216-272 */
    rank_t *k = &rk[r];
    k->pvals = calloc(nt, 8); k->ptrial = calloc(nt, 8); k->mu = calloc(nt, 8); k->sig = calloc(nt, 8);
    k->mutrial = calloc(nt, 8); k->sigtrial = calloc(nt, 8); k->psum2 = calloc(nt, 8);
    k->musigall = calloc(nm, 8);
    k->lylast = calloc(C, 8); k->lytrial = calloc(C, 8); k->cfac = calloc(C, 8); k->pacpt = calloc(C, 8);
    k->acpt = calloc(C, 8); k->qisum = calloc(C, 8); k->qimax = calloc(C, 8); k->cov = calloc(nc, 8);
    k->rjct = calloc(C, sizeof(int)); k->chnsel = calloc(C, sizeof(int));
    k->maxlparams = calloc(d, 8); k->maxlval = -INFINITY;            /* mcout.cc:19-20 */
    k->Z = Z ? Z + (size_t)r*nz : 0; k->nz = Z ? nz : 0;
    k->U = U ? U + (size_t)r*nu : 0; k->nu = U ? nu : 0;
    k->I = I ? I + (size_t)r*ni : 0; k->ni = I ? ni : 0;
    orc_covar_setup(d, incov, k->cov);                               /* run(): mcpar.cc:20 */
    const double *pi = pinit + (cfg->pinit_per_rank ? (size_t)r*nt : 0);
    for (int i = 0; i < nt; ++i) k->pvals[i] = pi[i];                /* :47-50 */
    rc = orc_loglik(cfg->lik, d, par, C, k->pvals, k->lylast);       /* :53 */
    if (rc) goto done;
    k->irate = 50;                                                   /* :57 */
  }

  int step = 0;     /* trace index */
  /* ---- burn-in, mcpar.cc:56-97 ---- */
  for (int isamp = 0; isamp < cfg->nburn; ++isamp, ++step) {
    for (int r = 0; r < R; ++r) {
      rank_t *k = &rk[r];
      gen_local(k, C, d);                                            /* :59 */
      orc_loglik(cfg->lik, d, par, C, k->ptrial, k->lytrial);        /* :60 */
      if (step < T && tr_pre_p) {
        size_t o = (size_t)r*T + step;
        memcpy(tr_pre_p + o*nt, k->pvals, 8*nt); memcpy(tr_pre_ly + o*C, k->lylast, 8*C);
        memcpy(tr_trial_p + o*nt, k->ptrial, 8*nt); memcpy(tr_trial_ly + o*C, k->lytrial, 8*C);
        memcpy(tr_cfac + o*C, k->cfac, 8*C); memcpy(tr_cov + o*nc, k->cov, 8*nc);
        if (cfg->trace_musig && tr_musig) memcpy(tr_musig + o*nm, k->musigall, 8*nm);
        tr_cursors[3*o] = (long long)k->iz; tr_cursors[3*o+1] = (long long)k->iu; tr_cursors[3*o+2] = (long long)k->ii;
        if (tr_remote) tr_remote[o] = 0;
        if (tr_iters) tr_iters[o] = 0;
      }
      for (int j = 0; j < C; ++j) k->acpt[j] = take_u(k);            /* :63 */
      k->ntrial += C;                                                /* :65 */
      for (int j = 0; j < C; ++j) {                                  /* :66-70 */
        k->pacpt[j] = exp(k->lytrial[j] - k->lylast[j]);
        int a = k->acpt[j] < k->pacpt[j];
        k->lylast[j] = a ? k->lytrial[j] : k->lylast[j];
        k->naccept += a;
        if (step < T && tr_accept) tr_accept[((size_t)r*T + step)*C + j] = a;
      }
      for (int i = 0; i < nt; ++i) { int j = i/d; if (k->acpt[j] < k->pacpt[j]) k->pvals[i] = k->ptrial[i]; }  /* :71-75 */
      if (isamp > k->irate) {                                        /* :78-96 */
        double arate = k->naccept / k->ntrial;
        if (arate < cfg->armin)      { k->naccept = k->ntrial = 0.0; for (int i = 0; i < nc; ++i) k->cov[i] *= cfg->dfac; }
        else if (arate > cfg->armax) { k->naccept = k->ntrial = 0.0; for (int i = 0; i < nc; ++i) k->cov[i] *= cfg->ifac; }
        k->irate += 50;
      }
    }
  }

  for (int r = 0; r < R; ++r)                                        /* :100-103 */
    for (int i = 0; i < nt; ++i) { rk[r].mu[i] = 0.0; rk[r].psum2[i] = FPEPS; }
  double pwgt = 0.0;                                                 /* :104 (same on every rank) */

  /* ---- main loop, mcpar.cc:113-210 ---- */
  for (int isamp = 0; isamp < cfg->nsamp; ++isamp, ++step) {
    if (isamp % cfg->sync == 0 && R > 1)                             /* :127-140 in-place all-gather */
      for (int r = 0; r < R; ++r)
        for (int s = 0; s < R; ++s) if (s != r)
          memcpy(rk[r].musigall + (size_t)s*2*nt, rk[s].musigall + (size_t)s*2*nt, 8*(size_t)2*nt);
    pwgt += 1.0;                                                     /* :186 (hoisted: identical on all ranks) */
    const double winv = 1.0 / pwgt;                                  /* :187 */
    for (int r = 0; r < R; ++r) {
      rank_t *k = &rk[r];
      double rndlocal = isamp < cfg->sync ? 0.0 : take_u(k);          /* :142-146 */
      int remotep;
      if (rndlocal <= cfg->pl) { gen_local(k, C, d); remotep = 0; k->iters_last = 0; }   /* :152-155 */
      else { gen_remote(k, C, d, N); remotep = 1; }                   /* :156-159 */
      orc_loglik(cfg->lik, d, par, C, k->ptrial, k->lytrial);         /* :160 */
      if (step < T && tr_pre_p) {
        size_t o = (size_t)r*T + step;
        memcpy(tr_pre_p + o*nt, k->pvals, 8*nt); memcpy(tr_pre_ly + o*C, k->lylast, 8*C);
        memcpy(tr_trial_p + o*nt, k->ptrial, 8*nt); memcpy(tr_trial_ly + o*C, k->lytrial, 8*C);
        memcpy(tr_cfac + o*C, k->cfac, 8*C); memcpy(tr_cov + o*nc, k->cov, 8*nc);
        if (cfg->trace_musig && tr_musig) memcpy(tr_musig + o*nm, k->musigall, 8*nm);
        tr_cursors[3*o] = (long long)k->iz; tr_cursors[3*o+1] = (long long)k->iu; tr_cursors[3*o+2] = (long long)k->ii;
        if (tr_remote) tr_remote[o] = remotep;
        if (tr_iters) tr_iters[o] = k->iters_last;
      }
      for (int j = 0; j < C; ++j) k->acpt[j] = take_u(k);             /* :163 */
      k->ntrial += C;
      for (int j = 0; j < C; ++j) {                                   /* :166-170 */
        k->pacpt[j] = exp(k->lytrial[j] - k->lylast[j]) * k->cfac[j];
        int a = k->acpt[j] < k->pacpt[j];
        k->lylast[j] = a ? k->lytrial[j] : k->lylast[j];
        k->naccept += a;
        if (step < T && tr_accept) tr_accept[((size_t)r*T + step)*C + j] = a;
      }
      for (int i = 0; i < nt; ++i) { int j = i/d; if (k->acpt[j] < k->pacpt[j]) k->pvals[i] = k->ptrial[i]; }  /* :171-175 */
      for (int j = 0; j < C; ++j) {                                   /* :177-182 MCout::add, mcout.cc:129-145 */
        if (rows) {
          double *row = rows + (((size_t)r*cfg->nsamp + isamp)*C + j)*(d+1);
          for (int i = 0; i < d; ++i) row[i] = k->pvals[j*d+i];
          row[d] = k->lylast[j];
        }
        if (k->lylast[j] > k->maxlval) { k->maxlval = k->lylast[j]; for (int i = 0; i < d; ++i) k->maxlparams[i] = k->pvals[j*d+i]; }
      }
      for (int i = 0; i < nt; ++i) {                                  /* :188-209 */
        int j = i/d;
        int adopt = remotep && (k->acpt[j] < k->pacpt[j]);
        welford(k->pvals[i], pwgt, winv, adopt, k->mutrial[i], k->sigtrial[i], &k->mu[i], &k->psum2[i], &k->sig[i]);
        size_t islot = 2*((size_t)r*nt + i);
        k->musigall[islot] = k->mu[i]; k->musigall[islot+1] = k->sig[i];
      }
    }
  }

  for (int r = 0; r < R; ++r) {
    rank_t *k = &rk[r];
    if (st_p) memcpy(st_p + (size_t)r*nt, k->pvals, 8*nt);
    if (st_ly) memcpy(st_ly + (size_t)r*C, k->lylast, 8*C);
    if (st_mu) memcpy(st_mu + (size_t)r*nt, k->mu, 8*nt);
    if (st_sig) memcpy(st_sig + (size_t)r*nt, k->sig, 8*nt);
    if (st_psum2) memcpy(st_psum2 + (size_t)r*nt, k->psum2, 8*nt);
    if (st_cov) memcpy(st_cov + (size_t)r*nc, k->cov, 8*nc);
    if (st_musig) memcpy(st_musig + (size_t)r*nm, k->musigall, 8*nm);
    if (used) { used[4*r] = (long long)k->iz; used[4*r+1] = (long long)k->iu; used[4*r+2] = (long long)k->ii; used[4*r+3] = k->overrun; }
  }
  if (maxl) {                                                         /* MCout::maxlike, mcout.cc:96-127 */
    int best = 0;
    for (int r = 1; r < R; ++r) if (rk[r].maxlval > rk[best].maxlval) best = r;
    for (int i = 0; i < d; ++i) maxl[i] = rk[best].maxlparams[i];
    maxl[d] = rk[best].maxlval;
  }
done:
  for (int r = 0; r < R; ++r) {
    rank_t *k = &rk[r];
    free(k->pvals); free(k->ptrial); free(k->mu); free(k->sig); free(k->mutrial); free(k->sigtrial);
    free(k->psum2); free(k->musigall); free(k->lylast); free(k->lytrial); free(k->cfac); free(k->pacpt);
    free(k->acpt); free(k->qisum); free(k->qimax); free(k->cov); free(k->rjct); free(k->chnsel); free(k->maxlparams);
  }
  free(rk);
  return rc;
}

/* ------------------------------------------------------------------ */
/* COUNTER engine: same algorithm, draws addressed by (chain, step, slot) */
/* ------------------------------------------------------------------ */
/* Philox counter = (g_lo, g_hi, step, slot); key = seed.  Every draw is ONE 32-bit word;
 * words are addressed as a stream: word idx of (chain, step, base) is word idx%4 of the block
 * at slot base + idx/4.
 *   local step (base 0):   pair q -> words (2q, 2q+1); accept uniform -> word 2*NP;
 *                          local/remote coin (of the coin group's leader chain) -> word 2*NP+1
 *   remote candidate it (base 0x40000000 | it<<6): word 0 -> component pick, word 1 ->
 *                          rejection uniform, pair q -> words (2+2q, 3+2q)
 * NP = ceil(d/2).  Box-Muller (MKL BOXMULLER2 convention): z0 = r sin(2 pi u2),
 * z1 = r cos(2 pi u2), r = sqrt(-2 ln v), v = (wa+1) 2^-32, u2 = wb 2^-32.             */
#define SLOT_REMOTE 0x40000000u
#define W32 (1.0 / 4294967296.0)

static uint32_t draw_word(const orc_config *cfg, uint64_t g, uint32_t step, uint32_t base, int idx)
{
  uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), step, base + (uint32_t)(idx / 4)};
  uint32_t key[2] = {(uint32_t)cfg->seed, (uint32_t)(cfg->seed >> 32)}, w[4];
  orc_philox4x32_10(ctr, key, w);
  return w[idx % 4];
}

static void normal_pair(uint32_t wa, uint32_t wb, double *z0, double *z1)
{
  double v = ((double)wa + 1.0) * W32, u2 = (double)wb * W32;
  double r = sqrt(fmax(-2.0 * log(v), 0.0));
  double a = 6.283185307179586476925 * u2;
  *z0 = r * sin(a); *z1 = r * cos(a);
}

int orc_run_counter(const orc_config *cfg, const double *pinit, const double *incov, const double *par,
                    double *rows, double *st_p, double *st_ly, double *st_mu, double *st_psum2,
                    double *pool_out, long long *acc_counts, double *cov_out,
                    unsigned char *tr_accept, long long *remote_iters)
{
  const int N = cfg->nchain, d = cfg->nparam;
  const int G = cfg->coin_group > 0 ? cfg->coin_group : N;   /* 0: one coin per step for the whole job */
  const int M = (cfg->pool_m > 0 && cfg->pool_m < N) ? cfg->pool_m : N;
  const int stride = N / M;
  const int thin = cfg->thin > 0 ? cfg->thin : 1;
  if (d > 256) return -1;
  double *x = malloc(8*(size_t)N*d), *ly = malloc(8*(size_t)N), *mu = calloc((size_t)N*d, 8), *ps = malloc(8*(size_t)N*d);
  const int lag = cfg->pool_lag > 0 ? 1 : 0;
  const int first_remote = cfg->sync * (1 + lag);     /* no pool to read before that */
  double *pool = calloc((size_t)M*d*2, 8), *pool_new = calloc((size_t)M*d*2, 8), *T0 = malloc(8*(size_t)d*d);
  double xt[256], z[257], mut[256], sigt[256];
  memset(mut, 0, sizeof mut); memset(sigt, 0, sizeof sigt);
  memcpy(x, pinit, 8*(size_t)N*d);
  orc_covar_setup(d, incov, T0);
  int rc = orc_loglik(cfg->lik, d, par, N, x, ly);
  if (rc) goto done;
  long long nacc = 0, ntry = 0; int irate = 50;
  long long riters = 0;
  uint32_t step = 0;
  const int NP = (d + 1) / 2;

  for (int isamp = 0; isamp < cfg->nburn + cfg->nsamp; ++isamp, ++step) {
    const int burn = isamp < cfg->nburn;
    const int t = isamp - cfg->nburn;                 /* main-phase step */
    if (!burn && t == 0) { for (size_t i = 0; i < (size_t)N*d; ++i) { mu[i] = 0.0; ps[i] = FPEPS; } nacc = ntry = 0; }
    const double pwgt = burn ? 0.0 : (double)(t + 1), winv = burn ? 0.0 : 1.0 / pwgt;
    if (!burn && t % cfg->sync == 0) {                /* exchange: refresh the pool (mcpar.cc:127-140) */
      if (lag) memcpy(pool, pool_new, 8*(size_t)M*d*2);   /* pool_lag 1: read what the PREVIOUS exchange delivered */
      double *dst = lag ? pool_new : pool;
      for (int s = 0; s < M; ++s) { size_t g = (size_t)s*stride;
        for (int i = 0; i < d; ++i) { dst[((size_t)s*d+i)*2] = mu[g*d+i]; dst[((size_t)s*d+i)*2+1] = ps[g*d+i] * (t ? 1.0/(double)t : 0.0); } }
    }
    for (int g = 0; g < N; ++g) {
      double *xg = x + (size_t)g*d;
      int remotep = 0; double cfac = 1.0, lcfac = 0.0;
      if (!burn && t >= first_remote) {                   /* mcpar.cc:142-159, one coin per group */
        remotep = !((double)draw_word(cfg, (uint64_t)(g / G) * G, step, 0, 2*NP + 1) * W32 <= cfg->pl);
      }
      if (!remotep) {                                  /* genLocal with the scaled factor */
        for (int p = 0; p < NP; ++p) normal_pair(draw_word(cfg, g, step, 0, 2*p), draw_word(cfg, g, step, 0, 2*p + 1), &z[2*p], &z[2*p+1]);
        for (int i = 0; i < d; ++i) { double acc = xg[i]; for (int q = 0; q <= i; ++q) acc += T0[i*d+q] * z[q]; xt[i] = acc; }
      } else if (cfg->remote_mode == 1) {              /* sum-mixture independence proposal, no rejection loop:
                                                          x' ~ (1/M) sum_s N(mu_s, diag sig2_s), drawn as candidate 0 of
                                                          the remote stream (word 0 -> component, word 1 unused);
                                                          Hastings factor q(x)/q(x') with normalised components */
        ++riters;
        const int c = (int)(((uint64_t)draw_word(cfg, g, step, SLOT_REMOTE, 0) * (uint64_t)M) >> 32);
        for (int p = 0; p < NP; ++p) normal_pair(draw_word(cfg, g, step, SLOT_REMOTE, 2 + 2*p), draw_word(cfg, g, step, SLOT_REMOTE, 3 + 2*p), &z[2*p], &z[2*p+1]);
        for (int i = 0; i < d; ++i) {
          mut[i] = pool[((size_t)c*d+i)*2]; sigt[i] = sqrt(pool[((size_t)c*d+i)*2+1]);
          xt[i] = mut[i] + sigt[i]*z[i];
        }
        lcfac = pool_lse(d, M, pool, xg) - pool_lse(d, M, pool, xt);
        for (int i = 0; i < d; ++i) sigt[i] *= sigt[i];
      } else {                                         /* genRemote, per-chain rejection loop over the pool */
        double qmax = 0, qsum = 0; int c = 0;
        for (uint32_t it = 0;; ++it) {
          ++riters;
          const uint32_t base = SLOT_REMOTE | (it << 6);
          c = (int)(((uint64_t)draw_word(cfg, g, step, base, 0) * (uint64_t)M) >> 32);
          double u = ((double)draw_word(cfg, g, step, base, 1) + 0.5) * W32;
          for (int p = 0; p < NP; ++p) normal_pair(draw_word(cfg, g, step, base, 2 + 2*p), draw_word(cfg, g, step, base, 3 + 2*p), &z[2*p], &z[2*p+1]);
          for (int i = 0; i < d; ++i) {
            mut[i] = pool[((size_t)c*d+i)*2]; sigt[i] = sqrt(pool[((size_t)c*d+i)*2+1]);
            xt[i] = mut[i] + sigt[i]*z[i];
          }
          qmax = FPEPS; qsum = FPEPS;
          for (int s = 0; s < M; ++s) { double gv = q_value(d, pool + (size_t)s*d*2, xt); qsum += gv; qmax = gv > qmax ? gv : qmax; }
          if (u < qmax/qsum) break;
          if (it >= (1u << 24) - 1) break;             /* slot space exhausted: give up (never in practice) */
        }
        double qold = 0.0;
        for (int s = 0; s < M; ++s) { double gv = q_value(d, pool + (size_t)s*d*2, xg); qold = gv > qold ? gv : qold; }
        cfac = qold / qmax;
        for (int i = 0; i < d; ++i) sigt[i] *= sigt[i];
      }
      double lyt;
      orc_loglik(cfg->lik, d, par, 1, xt, &lyt);
      double u = ((double)draw_word(cfg, g, step, 0, 2*NP) + 0.5) * W32;
      double pac = exp(lyt - ly[g]); if (!burn) pac *= cfac;
      if (remotep && cfg->remote_mode == 1) pac = exp((lyt - ly[g]) + lcfac);   /* one exponential of the summed logs */
      int a = u < pac;
      ++ntry; nacc += a;
      if (tr_accept) tr_accept[(size_t)isamp*N + g] = (unsigned char)(a | (remotep << 1));
      if (a) { ly[g] = lyt; for (int i = 0; i < d; ++i) xg[i] = xt[i]; }
      if (!burn) {
        if (rows && t % thin == 0) { double *row = rows + (((size_t)(t/thin))*N + g)*(d+1); for (int i = 0; i < d; ++i) row[i] = xg[i]; row[d] = ly[g]; }
        for (int i = 0; i < d; ++i) { double sg; welford(xg[i], pwgt, winv, remotep && a, mut[i], sigt[i], &mu[(size_t)g*d+i], &ps[(size_t)g*d+i], &sg); }
      }
    }
    if (burn && isamp > irate) {                       /* tuning on the GLOBAL acceptance rate */
      double arate = (double)nacc / (double)ntry;
      if (arate < cfg->armin)      { nacc = ntry = 0; for (int i = 0; i < d*d; ++i) T0[i] *= cfg->dfac; }
      else if (arate > cfg->armax) { nacc = ntry = 0; for (int i = 0; i < d*d; ++i) T0[i] *= cfg->ifac; }
      irate += 50;
    }
  }
  if (st_p) memcpy(st_p, x, 8*(size_t)N*d);
  if (st_ly) memcpy(st_ly, ly, 8*(size_t)N);
  if (st_mu) memcpy(st_mu, mu, 8*(size_t)N*d);
  if (st_psum2) memcpy(st_psum2, ps, 8*(size_t)N*d);
  if (pool_out) {                                    /* the latest publication: the engine publishes at the END of every
                                                        window, so a run that ends on a boundary has one more than the loop
                                                        above has read */
    double *last = lag ? pool_new : pool;
    if (cfg->nsamp > 0 && cfg->nsamp % cfg->sync == 0)
      for (int s = 0; s < M; ++s) { size_t g = (size_t)s*stride;
        for (int i = 0; i < d; ++i) { last[((size_t)s*d+i)*2] = mu[g*d+i]; last[((size_t)s*d+i)*2+1] = ps[g*d+i] * (1.0/(double)cfg->nsamp); } }
    memcpy(pool_out, last, 8*(size_t)M*d*2);
  }
  if (acc_counts) { acc_counts[0] = nacc; acc_counts[1] = ntry; }
  if (cov_out) memcpy(cov_out, T0, 8*(size_t)d*d);
  if (remote_iters) *remote_iters = riters;
done:
  free(x); free(ly); free(mu); free(ps); free(pool); free(pool_new); free(T0);
  return rc;
}
