// driver_opts.hh -- the optional flags shared by the driver mains.  The reference's mains take positional
// arguments only (`mcpar-rosen1 [nsamp]`, src/mcpar-rosen1.cc:33-34; `mpirun -np R` sets the rank count);
// here `--ranks=R` replaces mpirun and every other flag is optional, so the positional form is undisturbed.
#ifndef MCPAR_B200_DRIVER_OPTS_HH_
#define MCPAR_B200_DRIVER_OPTS_HH_
#include <stdlib.h>
#include <string.h>
#include "mcpar.hh"

struct DriverOpts {
  int nsamp, ranks, ngpu, pool, thin, remote_mode, lag, chains, npos;
  bool job_coin;
  const char *binfile;
  unsigned long long seed;
  explicit DriverOpts(int nsamp_default)
    : nsamp(nsamp_default), ranks(1), ngpu(1), pool(0), thin(1), remote_mode(0), lag(0), chains(4), npos(0),
      job_coin(false), binfile(0), seed(8675309ull) {}
  void parse(int argc, char *argv[]) {
    for (int i = 1; i < argc; ++i) {
      const char *a = argv[i];
      if (!strncmp(a, "--ranks=", 8)) ranks = atoi(a + 8);                 // mpirun -np R
      else if (!strncmp(a, "--ngpu=", 7)) ngpu = atoi(a + 7);              // GPUs the ranks are sharded over
      else if (!strncmp(a, "--chains=", 9)) chains = atoi(a + 9);          // chains per rank (the mains fix 4)
      else if (!strncmp(a, "--pool=", 7)) pool = atoi(a + 7);              // remote-mixture pool size, 0 = all chains
      else if (!strncmp(a, "--thin=", 7)) thin = atoi(a + 7);
      else if (!strncmp(a, "--remote-mode=", 14)) remote_mode = atoi(a + 14);   // 0 reference genRemote, 1 sum-mixture
      else if (!strncmp(a, "--lag=", 6)) lag = atoi(a + 6);                // read the pool one exchange later
      else if (!strcmp(a, "--job-coin")) job_coin = true;                  // one local/remote coin per step for the whole job
      else if (!strncmp(a, "--seed=", 7)) seed = strtoull(a + 7, 0, 10);
      else if (!strncmp(a, "--binary=", 9)) binfile = a + 9;               // rows to FILE in MCout's binary format
      else if (a[0] != '-' && npos++ == 0) nsamp = atoi(a);
    }
  }
  void apply(MCPar &m) const {
    m.pool_m = pool; m.thin = thin; m.ngpu = ngpu; m.remote_mode = remote_mode; m.pool_lag = lag; m.seed = seed;
    if (job_coin) m.coin_group = 0;
  }
};
#endif
