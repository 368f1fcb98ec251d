// mcpar.cc -- MCPar on the B200 engine.  Same constructor, run() signature, log file and
// output cadence as the reference (src/mcpar.cc:17-214), but the chain loop is a handful
// of calls into the C ABI: the device keeps the state and runs SYNCSTEP-step launches.
#include "mcpar.hh"
#include "../../include/mcgpu.h"
#include <fstream>
#include <sstream>
#include <iomanip>
#include <vector>
#include <new>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <typeinfo>
#include <algorithm>

const Real MCPar::FPEPS = 1.0e-14;

namespace {
void die(mcgpu_engine *e, const char *what, int rc)
{
  fprintf(stderr, "MCPar: %s failed (%d): %s\n", what, rc, mcgpu_last_error(e));
  exit(3);                                   // the reference aborts on MKL/MPI errors (mcpar.hh:93)
}
#define CHECK(call) do { int _rc = (call); if (_rc != MCGPU_OK) die(eng, #call, _rc); } while (0)
}

MCPar::MCPar(int np, int nc, int mpisiz, int mpirank, Real pl, Real armin, Real armax, Real dfac, Real ifac, int sync)
  : TGT_ARATE_MIN(armin), TGT_ARATE_MAX(armax), SCALE_DEC(dfac), SCALE_INC(ifac), PLOCAL(pl), SYNCSTEP(sync),
    logging(false), logstep(1000), device(0), ngpu(1), pool_m(0), thin(1), seed(8675309ull),
    nparam(np), nchain(nc), size(mpisiz < 1 ? 1 : mpisiz), rank(mpirank), mdevice_ms(0), maccept(0)
{
  tchains = size * nchain;
  if (rank != 0) {
    fprintf(stderr, "rank = %d:  this build hosts every rank in one process; construct MCPar with mpirank 0\n", rank);
    throw("Invalid rank: the B200 engine is single-process");
  }
}

void MCPar::destroy_engines()
{
  for (size_t g = 0; g < engs.size(); ++g) if (engs[g]) mcgpu_destroy(engs[g]);
  engs.clear();
}

MCPar::~MCPar() { destroy_engines(); }

int MCPar::run(int nsamp, int nburn, const Real *pinit, VLFunc &L, MCout &outsamples, Real *incov)
{
  DeviceVLFunc *dl = dynamic_cast<DeviceVLFunc *>(&L);
  if (!dl) {
    fprintf(stderr, "MCPar::run: this likelihood has no device functor; the B200 engine has no host path\n");
    return ERROR;
  }
  // log file, as the reference: rank 0 writes mcpar-log.000.txt in the working directory
  std::stringstream logname;
  logname << "mcpar-log." << std::setfill('0') << std::setw(3) << rank << ".txt";
  std::ofstream logfile(logname.str().c_str());

  try {
    outsamples.newsamps((size_t)((nsamp + thin - 1) / thin) * (size_t)tchains);
  } catch (std::bad_alloc &) {
    logfile << "Unable to allocate space for output samples.  Exiting.\n";
    exit(2);
  }

  // chains sharded over G engines by contiguous blocks of ranks (the musigall slot rule, mcpar.cc:206)
  destroy_engines();
  const int G = ngpu < 1 ? 1 : ngpu;
  const int ndev = mcgpu_device_count();
  if (size % G != 0 || (G > 1 && (((long long)(size / G) * nchain) % 32) != 0)) {
    fprintf(stderr, "MCPar::run: ngpu = %d needs mpisiz %% ngpu == 0 and (mpisiz / ngpu) * nc a multiple of 32\n", G);
    return INVALID;
  }
  const long long cg_chains = (long long)(size / G) * nchain;          // chains per engine
  const int rpg = size / G;                                           // ranks per engine
  // one local/remote coin per rank-sized group of chains (mcpar.cc:106-109,142-159)
  int cg = 1; while (cg * 2 <= nchain && cg < 32) cg *= 2;
  mcgpu_engine *eng = 0;                                              // the engine CHECK reports on
  for (int g = 0; g < G; ++g) {
    mcgpu_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = MCGPU_ABI_VERSION; cfg.device = ndev > 0 ? (device + g) % ndev : device; cfg.mode = MCGPU_MODE_NORMAL;
    cfg.nparam = nparam; cfg.nchain = cg_chains; cfg.chain0 = g * cg_chains; cfg.nchain_total = tchains;
    cfg.sync = SYNCSTEP; cfg.pl = PLOCAL; cfg.armin = TGT_ARATE_MIN; cfg.armax = TGT_ARATE_MAX;
    cfg.dfac = SCALE_DEC; cfg.ifac = SCALE_INC; cfg.seed = seed;
    cfg.coin_group = cg;
    cfg.pool_m = pool_m; cfg.thin = thin;
    cfg.history_steps = (nsamp + thin - 1) / thin;
    mcgpu_engine *e = 0;
    const int rc = mcgpu_create(&cfg, &e);
    if (rc != MCGPU_OK) die(0, "mcgpu_create", rc);
    engs.push_back(e);
  }
  eng = engs[0];

  const std::vector<double> &par = dl->params();
  // pinit holds np*nc values; every rank starts from the same block (the mains pass the
  // same array on every rank, mcpar-rosen1.cc:43)
  std::vector<Real> p0((size_t)cg_chains * nparam);
  for (int r = 0; r < rpg; ++r) memcpy(&p0[(size_t)r * nchain * nparam], pinit, sizeof(Real) * (size_t)nchain * nparam);
  for (int g = 0; g < G; ++g) {
    eng = engs[g];
    CHECK(mcgpu_set_likelihood(eng, dl->lik_id(), par.empty() ? 0 : &par[0], (int)par.size()));
    CHECK(mcgpu_set_covariance(eng, incov));
    CHECK(mcgpu_set_state(eng, &p0[0]));
  }
  eng = engs[0];
  if (G > 1) CHECK(mcgpu_p2p_attach_local(&engs[0], G));

  logfile << "Starting burn-in.  Samples = " << nburn << std::endl;
  if (G > 1) CHECK(mcgpu_burnin_group(&engs[0], G, nburn));            // tuning counters summed over the engines
  else CHECK(mcgpu_burnin(eng, nburn));

  const int outstep = nsamp > 50 ? nsamp / 10 : 5;       // mcpar.cc:110
  logfile << "Starting main sample loop:  nsamp = " << nsamp << std::endl;
  logfile << "Output after each " << outstep << " steps." << std::endl;
  for (int g = 0; g < G; ++g) { eng = engs[g]; CHECK(mcgpu_sample_begin(eng, nsamp)); }

  // Rows reach MCout in the reference's order: batches of `outstep` steps; inside a batch
  // rank-major blocks; inside a rank block step-major, then chain (mcout.cc:52-94 gathers
  // rank blocks; src/anly/mcpar-analysis.R:80-120 relies on it).
  const int ncol = nparam + 1;
  std::vector<std::vector<Real> > block(G);
  int done = 0;
  while (done < nsamp) {
    const int n = (nsamp - done < outstep) ? nsamp - done : outstep;
    if (done > 0) {
      logfile << "Beginning output at step " << done << std::endl;
      outsamples.output();
      logfile << "Output finished\n" << std::endl;
    }
    if (logging && done % logstep == 0)
      logfile << "sample step " << done << ":\toutsamples size= " << outsamples.size() << "  maxsize = "
              << outsamples.maxsize() << "  ncol= " << outsamples.ncol() << std::endl;
    if (G == 1) CHECK(mcgpu_sample(eng, n));
    else {
      // engines are fed one exchange window at a time, in turn: a window kernel that waits for its peers'
      // publications is then never queued in front of the kernels that make them (engines may share a device)
      for (int k = 0; k < n;) {
        const int m = std::min(n - k, SYNCSTEP - (done + k) % SYNCSTEP);
        for (int g = 0; g < G; ++g) { eng = engs[g]; CHECK(mcgpu_sample(eng, m)); }
        k += m;
      }
    }
    const long long k0 = (done + thin - 1) / thin, k1 = (done + n + thin - 1) / thin;   // kept steps of this batch
    if (k1 > k0) {
      for (int g = 0; g < G; ++g) {
        eng = engs[g];
        block[g].resize((size_t)(k1 - k0) * cg_chains * ncol);
        CHECK(mcgpu_history_read(eng, k0, k1 - k0, &block[g][0]));
      }
      for (int r = 0; r < size; ++r)
        for (long long k = 0; k < k1 - k0; ++k)
          outsamples.addrows(&block[r / rpg][((size_t)k * cg_chains + (size_t)(r % rpg) * nchain) * ncol], (size_t)nchain);
    }
    done += n;
  }
  outsamples.output();                                   // remaining samples (mcpar.cc:212)

  long long acc = 0, tried = 0;
  mdevice_ms = 0;
  for (int g = 0; g < G; ++g) {
    eng = engs[g];
    CHECK(mcgpu_synchronize(eng));                         // also reports a peer that failed to publish
    mcgpu_stats st;
    CHECK(mcgpu_get_stats(eng, &st));
    acc += st.accepted; tried += st.tried;
    if (st.device_ms > mdevice_ms) mdevice_ms = st.device_ms;
  }
  maccept = tried ? (double)acc / (double)tried : 0.0;
  logfile << "Acceptance rate (main loop) = " << maccept << "  device time = " << mdevice_ms << " ms  ("
          << (double)tchains * (nburn + nsamp) / (mdevice_ms * 1e-3) << " chain-steps/s)" << std::endl;
  return OK;
}
