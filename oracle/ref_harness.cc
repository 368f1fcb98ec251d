/* oracle/ref_harness.cc -- TEST INFRASTRUCTURE, not product code.
 *
 * C-ABI harness around the reference's OWN, UNMODIFIED sources
 * (/root/reference/src/{mcpar,mcout,rosenbrock,mcutil}.cc, compiled where they lie
 * by oracle/Makefile against oracle/shim/).  It constructs the reference's MCPar /
 * MCout / likelihood objects exactly as its mains do (mcpar-rosen1.cc:26-45,
 * mcpar-dgauss.cc:15-35), runs R ranks as threads over the mini-MPI, and records a
 * per-step trace through a pass-through VLFunc so the CUDA engine and the C
 * restatement (oracle/mh_oracle.c) can be compared with the real thing.
 *
 * Nothing here is shipped or timed as the product.  Only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load the library built
 * from this file.
 */
#include <iostream>
#include <fstream>
#include <sstream>
#include <iomanip>
#include <vector>
#include <limits>
#include <string>
#include <thread>
#include <chrono>
#include <streambuf>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include <unistd.h>

/* the trace needs MCPar's working arrays; the class layout is untouched */
#define private public
#include "mcpar.hh"
#undef private
#include "mcout.hh"
#include "mcutil.hh"
#include "rosenbrock.hh"
#include "vlfunc.hh"
#include "mpi.h"
#include "mkl_vsl.h"

typedef float real_t;   /* the build's real type: double under prelude64.h */

namespace {

struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };

VLFunc *make_lik(int lik, int nparam, const real_t *par) {
  switch (lik) {
    case 0: return new Rosenbrock1(nparam);
    case 1: return new Rosenbrock2(nparam);
    case 2: return new Gaussian(nparam, par, par ? par + 2 : 0);
    case 3: return new DualGaussian(par ? par[0] : (real_t)5);
  }
  return 0;
}

struct Trace {          /* one rank's slice of the caller's trace buffers */
  int max_steps, nsteps;
  real_t *pre_p, *pre_ly, *trial_p, *trial_ly, *cfac, *cov, *musig;
  long long *cursors;
};

/* pass-through likelihood that snapshots the engine at every trial evaluation
 * (the L(nchain,ptrial,lytrial) calls at mcpar.cc:60 and :160) */
struct Recorder : public VLFunc {
  VLFunc &inner; MCPar *mc; Trace *tr; shim_vsl::Source *src;
  Recorder(VLFunc &in, Trace *t, shim_vsl::Source *s) : inner(in), mc(0), tr(t), src(s) {}
  int operator()(int npset, const real_t *x, real_t *restrict y) {
    int rc = inner(npset, x, y);
    if (tr && mc && x == mc->ptrial && tr->nsteps < tr->max_steps) {
      const int k = tr->nsteps++;
      const int C = mc->nchain, nt = mc->ntot, nc = mc->ncov;
      memcpy(tr->pre_p   + (size_t)k*nt, mc->pvals,  sizeof(real_t)*nt);
      memcpy(tr->pre_ly  + (size_t)k*C,  mc->lylast, sizeof(real_t)*C);
      memcpy(tr->trial_p + (size_t)k*nt, x,          sizeof(real_t)*nt);
      memcpy(tr->trial_ly+ (size_t)k*C,  y,          sizeof(real_t)*C);
      memcpy(tr->cfac    + (size_t)k*C,  mc->cfac,   sizeof(real_t)*C);
      memcpy(tr->cov     + (size_t)k*nc, mc->cov,    sizeof(real_t)*nc);
      if (tr->musig) {
        const size_t nm = (size_t)2 * mc->tchains * mc->nparam;
        memcpy(tr->musig + (size_t)k*nm, mc->musigall, sizeof(real_t)*nm);
      }
      tr->cursors[3*k+0] = (long long)src->iz;
      tr->cursors[3*k+1] = (long long)src->iu;
      tr->cursors[3*k+2] = (long long)src->ii;
    }
    return rc;
  }
};

std::string g_text;     /* rank-0 text output of the last ref_run */

}  // namespace

extern "C" {

struct ref_config {
  int nparam, nchain, nranks;
  int nsamp, nburn;
  int lik;              /* 0 Rosenbrock1, 1 Rosenbrock2, 2 Gaussian, 3 DualGaussian */
  real_t pl, armin, armax, dfac, ifac;
  int sync;
  int rng_mode;         /* 0 replay, 1 philox */
  unsigned long long seed;
  int text_sink;        /* 0 discard formatted text, 1 keep it (ref_last_text) */
  int trace_steps;      /* per-rank trace capacity (0 = no trace) */
  int trace_musig;      /* also snapshot musigall per traced step */
  int pinit_per_rank;   /* 0: every rank gets the same nchain*nparam block (as the mains do) */
};

int ref_real_bytes(void) { return (int)sizeof(real_t); }

/* batched log-likelihood through the reference's own VLFunc (rosenbrock.cc) */
int ref_loglik(int lik, int nparam, const real_t *par, int npset, const real_t *x, real_t *y) {
  VLFunc *L = 0;
  try { L = make_lik(lik, nparam, par); } catch (const char *) { return -2; }
  if (!L) return -1;
  int rc = (*L)(npset, x, y);
  delete L;
  return rc;
}

/* MCPar::covar_setup (mcpar.cc:454-484) on a d x d input (NULL => identity) */
int ref_covar_setup(int nparam, const real_t *incov, real_t *cov_out) {
  shim_mpi::world_begin(1); shim_mpi::thread_enter(0);
  shim_vsl::Source s; memset(&s, 0, sizeof s); shim_vsl::bind_thread_source(&s);
  MCPar mc(nparam, 1, 1, 0);
  mc.covar_setup(incov, mc.cov);
  memcpy(cov_out, mc.cov, sizeof(real_t)*nparam*nparam);
  return 0;
}

/* mcutil::qriguess (mcutil.cc:3-34) */
int ref_qriguess(int rank, int npset, int nparam, const real_t *plo, const real_t *phi, real_t *pout) {
  mcutil u;
  u.qriguess(rank, npset, nparam, plo, phi, pout);
  return 0;
}

/* Run the reference engine on cfg->nranks thread-ranks.
 *   pinit   [nranks or 1][nchain*nparam]
 *   incov   [nparam*nparam] or NULL
 *   par     likelihood parameters or NULL
 *   Z,U,I   per-rank replay streams, row r at Z + r*nz etc. (rng_mode 0)
 * Outputs (any may be NULL):
 *   rows        [nranks][nsamp*nchain][nparam+1]   each rank's MCout contents
 *   st_p,st_mu,st_sig,st_psum2 [nranks][nchain*nparam];  st_ly [nranks][nchain]
 *   st_cov      [nranks][nparam*nparam];  st_musig [nranks][2*tchains*nparam]
 *   used        [nranks][4] = consumed Z, U, I counts and the overrun flag
 *   maxl        [nparam+1]: MCout::maxlike parameters then value (collective)
 *   trace buffers: see Trace; each [nranks][trace_steps][...]
 * Returns elapsed seconds of the slowest rank's MCPar::run, or <0 on error.
 */
double ref_run(const ref_config *cfg, const real_t *pinit, const real_t *incov, const real_t *par,
               const double *Z, size_t nz, const double *U, size_t nu, const int *I, size_t ni,
               real_t *rows, real_t *st_p, real_t *st_ly, real_t *st_mu, real_t *st_sig,
               real_t *st_psum2, real_t *st_cov, real_t *st_musig, long long *used, real_t *maxl,
               real_t *tr_pre_p, real_t *tr_pre_ly, real_t *tr_trial_p, real_t *tr_trial_ly,
               real_t *tr_cfac, real_t *tr_cov, real_t *tr_musig, long long *tr_cursors,
               int *tr_nsteps)
{
  const int R = cfg->nranks, C = cfg->nchain, d = cfg->nparam;
  const int nt = C*d, nc = d*d, T = cfg->trace_steps;
  const size_t nm = (size_t)2*R*C*d;
  shim_mpi::world_begin(R);
  std::vector<double> secs(R, 0.0);
  std::vector<int> fail(R, 0);
  std::ostringstream text;
  NullBuf nullbuf; std::ostream nullstream(&nullbuf);
  std::ostream *sink = cfg->text_sink ? (std::ostream*)&text : &nullstream;

  auto body = [&](int r) {
    shim_mpi::thread_enter(r);
    shim_vsl::Source src; memset(&src, 0, sizeof src);
    src.mode = cfg->rng_mode;
    if (cfg->rng_mode == 0) {
      src.Z = Z ? Z + (size_t)r*nz : 0; src.nz = Z ? nz : 0;
      src.U = U ? U + (size_t)r*nu : 0; src.nu = U ? nu : 0;
      src.I = I ? I + (size_t)r*ni : 0; src.ni = I ? ni : 0;
    } else { src.seed = cfg->seed; src.stream_id = (unsigned long long)r; }
    shim_vsl::bind_thread_source(&src);

    VLFunc *L = 0;
    try { L = make_lik(cfg->lik, d, par); } catch (const char *) { fail[r] = 1; }
    if (!L) { fail[r] = 1; return; }

    Trace tr; memset(&tr, 0, sizeof tr);
    const bool tracing = T > 0 && tr_pre_p;
    if (tracing) {
      tr.max_steps = T;
      tr.pre_p    = tr_pre_p    + (size_t)r*T*nt;
      tr.pre_ly   = tr_pre_ly   + (size_t)r*T*C;
      tr.trial_p  = tr_trial_p  + (size_t)r*T*nt;
      tr.trial_ly = tr_trial_ly + (size_t)r*T*C;
      tr.cfac     = tr_cfac     + (size_t)r*T*C;
      tr.cov      = tr_cov      + (size_t)r*T*nc;
      tr.musig    = (cfg->trace_musig && tr_musig) ? tr_musig + (size_t)r*T*nm : 0;
      tr.cursors  = tr_cursors  + (size_t)r*T*3;
    }
    Recorder rec(*L, tracing ? &tr : 0, &src);

    /* same construction order as the mains: MCout, then MCPar, then run */
    MCout rslts(d, sink, MPI_COMM_WORLD);
    MCPar mc(d, C, R, r, cfg->pl, cfg->armin, cfg->armax, cfg->dfac, cfg->ifac, cfg->sync);
    rec.mc = &mc;
    /* musigall is new[]'d uninitialised in the reference; zero it so traces of
       never-written slots are deterministic (those slots are never consumed
       before the first all-gather that follows a full SYNCSTEP of writes) */
    memset(mc.musigall, 0, sizeof(real_t)*nm);

    const real_t *pi = pinit + (cfg->pinit_per_rank ? (size_t)r*nt : 0);
    std::vector<real_t> cov_in;
    real_t *icv = 0;
    if (incov) { cov_in.assign(incov, incov + nc); icv = &cov_in[0]; }

    auto t0 = std::chrono::steady_clock::now();
    mc.run(cfg->nsamp, cfg->nburn, pi, rec, rslts, icv);
    auto t1 = std::chrono::steady_clock::now();
    secs[r] = std::chrono::duration<double>(t1 - t0).count();

    if (rows && rslts.size() > 0)
      memcpy(rows + (size_t)r*cfg->nsamp*C*(d+1), rslts.getpset(0),
             sizeof(real_t)*(size_t)rslts.size()*(d+1));
    if (st_p)     memcpy(st_p     + (size_t)r*nt, mc.pvals,  sizeof(real_t)*nt);
    if (st_ly)    memcpy(st_ly    + (size_t)r*C,  mc.lylast, sizeof(real_t)*C);
    if (st_mu)    memcpy(st_mu    + (size_t)r*nt, mc.mu,     sizeof(real_t)*nt);
    if (st_sig)   memcpy(st_sig   + (size_t)r*nt, mc.sig,    sizeof(real_t)*nt);
    if (st_psum2) memcpy(st_psum2 + (size_t)r*nt, mc.psum2,  sizeof(real_t)*nt);
    if (st_cov)   memcpy(st_cov   + (size_t)r*nc, mc.cov,    sizeof(real_t)*nc);
    if (st_musig) memcpy(st_musig + (size_t)r*nm, mc.musigall, sizeof(real_t)*nm);
    if (used) {
      used[4*r+0] = (long long)src.iz; used[4*r+1] = (long long)src.iu;
      used[4*r+2] = (long long)src.ii; used[4*r+3] = src.overrun;
    }
    if (tr_nsteps) tr_nsteps[r] = tr.nsteps;
    if (maxl) {                 /* collective: every rank calls it (mcout.hh:47-49) */
      real_t lmax;
      const std::vector<real_t> &pm = rslts.maxlike(&lmax);
      if (r == 0) { for (int i = 0; i < d; ++i) maxl[i] = pm[i]; maxl[d] = lmax; }
    }
    delete L;
  };

  std::vector<std::thread> th;
  for (int r = 0; r < R; ++r) th.emplace_back(body, r);
  for (auto &t : th) t.join();
  g_text = text.str();
  double worst = 0;
  for (int r = 0; r < R; ++r) { if (fail[r]) return -1.0; if (secs[r] > worst) worst = secs[r]; }
  return worst;
}

size_t ref_last_text(char *buf, size_t cap) {
  if (buf && cap) { size_t n = g_text.size() < cap ? g_text.size() : cap; memcpy(buf, g_text.data(), n); }
  return g_text.size();
}

}  // extern "C"
