// mcpar.hh -- the MH driver with the reference's interface (src/mcpar.hh:10-91), hosted on
// the B200 engine through the C ABI (include/mcgpu.h).  One process owns all "ranks":
// mpisiz is the number of rank-sized chain groups (nc chains each) and mpirank must be 0.
#ifndef MCPAR_B200_MCPAR_HH_
#define MCPAR_B200_MCPAR_HH_
#include <vector>
#include "vlfunc.hh"
#include "mcout.hh"

struct mcgpu_engine;

class MCPar {
public:
  enum { OK, INVALID, ERROR };
  const Real TGT_ARATE_MIN, TGT_ARATE_MAX, SCALE_DEC, SCALE_INC, PLOCAL;
  const int SYNCSTEP;
  static const Real FPEPS;

  bool logging;                 // user switches, as in the reference (src/mcpar.hh:28-29)
  int logstep;

  // engine options beyond the reference's constructor (set before run())
  int device;                   // CUDA ordinal (of the first GPU)
  int ngpu;                     // GPUs to shard the rank-sized chain groups over (mpisiz % ngpu == 0 and
                                // (mpisiz / ngpu) * nc a multiple of 32); the engines exchange their
                                // (mu, sigma^2) slots peer to peer (mcgpu_p2p_attach_local), replacing
                                // MPI_Allgather (src/mcpar.cc:127-140).  More engines than devices share devices.
  int pool_m;                   // remote-mixture pool size; 0 = every chain, as the reference
  int thin;                     // keep every thin-th step
  int coin_group;               // -1: one local/remote coin per rank (nc chains), as the reference (mcpar.cc:106-109);
                                // 0: one coin per step for the whole job (needed by the d >= 32 kernels)
  int remote_mode;              // 0: the reference's genRemote (mcpar.cc:315-451); 1: normalised sum-mixture proposal
  int pool_lag;                 // 0 / 1: read the pool one exchange later (takes the GPU exchange off the critical path)
  long long history_bytes;      // device budget for the sample history ring (default 1 GiB per engine): rows are
                                // read back into MCout in pieces of at most that size
  unsigned long long seed;      // Philox key; reference seed by default (mcpar.cc:271)

  MCPar(int np, int nc = 1, int mpisiz = 1, int mpirank = 0, Real pl = 0.9, Real armin = 0.2,
        Real armax = 0.5, Real dfac = 0.2, Real ifac = 1.5, int sync = 10);
  ~MCPar();

  int run(int nsamp, int nburn, const Real *pinit, VLFunc &L, MCout &outsamples, Real *incov = 0);

  double last_device_ms() const { return mdevice_ms; }
  double last_accept_rate() const { return maccept; }
  double last_exchange_wait_ms() const { return mxwait_ms; }

private:
  int nparam, nchain, size, rank, tchains;
  std::vector<mcgpu_engine *> engs;
  void destroy_engines();
  double mdevice_ms, maccept, mxwait_ms;
  MCPar(const MCPar &); MCPar &operator=(const MCPar &);
};

#endif
