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
    logging(false), logstep(1000), device(0), pool_m(0), thin(1), seed(8675309ull),
    nparam(np), nchain(nc), size(mpisiz < 1 ? 1 : mpisiz), rank(mpirank), eng(0), mdevice_ms(0), maccept(0)
{
  tchains = size * nchain;
  if (rank != 0) {
    fprintf(stderr, "rank = %d:  this build hosts every rank in one process; construct MCPar with mpirank 0\n", rank);
    throw("Invalid rank: the B200 engine is single-process");
  }
}

MCPar::~MCPar()
{
  if (eng) mcgpu_destroy(eng);
}

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

  if (eng) { mcgpu_destroy(eng); eng = 0; }
  mcgpu_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.abi_version = MCGPU_ABI_VERSION; cfg.device = device; cfg.mode = MCGPU_MODE_NORMAL;
  cfg.nparam = nparam; cfg.nchain = tchains; cfg.chain0 = 0; cfg.nchain_total = tchains;
  cfg.sync = SYNCSTEP; cfg.pl = PLOCAL; cfg.armin = TGT_ARATE_MIN; cfg.armax = TGT_ARATE_MAX;
  cfg.dfac = SCALE_DEC; cfg.ifac = SCALE_INC; cfg.seed = seed;
  // one local/remote coin per rank-sized group of chains (mcpar.cc:106-109,142-159)
  int cg = 1; while (cg * 2 <= nchain && cg < 32) cg *= 2;
  cfg.coin_group = cg;
  cfg.pool_m = pool_m; cfg.thin = thin;
  cfg.history_steps = (nsamp + thin - 1) / thin;
  const int rc = mcgpu_create(&cfg, &eng);
  if (rc != MCGPU_OK) die(0, "mcgpu_create", rc);

  const std::vector<double> &par = dl->params();
  CHECK(mcgpu_set_likelihood(eng, dl->lik_id(), par.empty() ? 0 : &par[0], (int)par.size()));
  CHECK(mcgpu_set_covariance(eng, incov));
  // pinit holds np*nc values; every rank starts from the same block (the mains pass the
  // same array on every rank, mcpar-rosen1.cc:43)
  std::vector<Real> p0((size_t)tchains * nparam);
  for (int r = 0; r < size; ++r) memcpy(&p0[(size_t)r * nchain * nparam], pinit, sizeof(Real) * (size_t)nchain * nparam);
  CHECK(mcgpu_set_state(eng, &p0[0]));

  logfile << "Starting burn-in.  Samples = " << nburn << std::endl;
  CHECK(mcgpu_burnin(eng, nburn));

  const int outstep = nsamp > 50 ? nsamp / 10 : 5;       // mcpar.cc:110
  logfile << "Starting main sample loop:  nsamp = " << nsamp << std::endl;
  logfile << "Output after each " << outstep << " steps." << std::endl;
  CHECK(mcgpu_sample_begin(eng, nsamp));

  // Rows reach MCout in the reference's order: batches of `outstep` steps; inside a batch
  // rank-major blocks; inside a rank block step-major, then chain (mcout.cc:52-94 gathers
  // rank blocks; src/anly/mcpar-analysis.R:80-120 relies on it).
  const int ncol = nparam + 1;
  std::vector<Real> block;
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
    CHECK(mcgpu_sample(eng, n));
    const long long k0 = (done + thin - 1) / thin, k1 = (done + n + thin - 1) / thin;   // kept steps of this batch
    if (k1 > k0) {
      block.resize((size_t)(k1 - k0) * tchains * ncol);
      CHECK(mcgpu_history_read(eng, k0, k1 - k0, &block[0]));
      for (int r = 0; r < size; ++r)
        for (long long k = 0; k < k1 - k0; ++k)
          outsamples.addrows(&block[((size_t)k * tchains + (size_t)r * nchain) * ncol], (size_t)nchain);
    }
    done += n;
  }
  outsamples.output();                                   // remaining samples (mcpar.cc:212)

  mcgpu_stats st;
  CHECK(mcgpu_get_stats(eng, &st));
  mdevice_ms = st.device_ms;
  maccept = st.tried ? (double)st.accepted / (double)st.tried : 0.0;
  logfile << "Acceptance rate (main loop) = " << maccept << "  device time = " << mdevice_ms << " ms  ("
          << (double)tchains * (nburn + nsamp) / (mdevice_ms * 1e-3) << " chain-steps/s)" << std::endl;
  return OK;
}
