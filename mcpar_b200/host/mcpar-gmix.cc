// mcpar-gmix -- the sum-of-Gaussians mixture BASELINE.json names (config 4: d = 64, 64 components, chains
// partitioned over the GPUs, exchange every sweep).  The reference has no such likelihood (only the fixed 2-D
// DualGaussian, src/rosenbrock.cc:63-78) and no such main; this one follows the shape of its mains
// (src/mcpar-dgauss.cc): `mcpar-gmix [nsamp]`, "nsamp = N", 4 chains per rank, nburn 200, rows on stdout or
// --binary=FILE, then the maximum-likelihood row.  SURVEY.md 8(d) C4 fixes the rest: mu_ki = 10 (u - 1/2),
// sig2_ki = 1/2 + 3/2 u', w_k = 1 (u, u' from a splitmix64 stream of the seed), chain g starts at mu_{g mod K},
// diagonal incov (2.38^2/64) I, SYNCSTEP 1, one local/remote coin per step for the whole job.
#include <iostream>
#include <fstream>
#include <vector>
#include "mcpar.hh"
#include "rosenbrock.hh"
#include "mcout.hh"
#include "driver_opts.hh"

static double next_u(unsigned long long &st)
{
  unsigned long long z = (st += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char *argv[])
{
  const int nparam = 64, K = 64;
  DriverOpts o(500);
  o.pool = 256; o.thin = 100; o.job_coin = true;
  o.parse(argc, argv);
  try {
    std::vector<Real> mu(K * nparam), s2(K * nparam), w(K, 1.0);
    unsigned long long st = o.seed;
    for (int i = 0; i < K * nparam; ++i) mu[i] = 10.0 * (next_u(st) - 0.5);
    for (int i = 0; i < K * nparam; ++i) s2[i] = 0.5 + 1.5 * next_u(st);
    GaussMix L(nparam, K, &mu[0], &s2[0], &w[0]);
    std::ofstream bin;
    if (o.binfile) {
      bin.open(o.binfile, std::ios::binary);
      if (!bin) { std::cerr << "cannot open " << o.binfile << "\n"; return 1; }
    }
    MCout rslts(nparam, o.binfile ? static_cast<std::ostream *>(&bin) : &std::cout, 0);
    if (o.binfile) rslts.set_format(MCout::BINARY);
    std::cout << "nsamp = " << o.nsamp << "\n";
    MCPar mcpar(nparam, o.chains, o.ranks, 0, 0.9, 0.2, 0.5, 0.2, 1.5, /* sync = */ 1);
    o.apply(mcpar);
    if (o.pool > o.chains * o.ranks) mcpar.pool_m = 0;
    // every rank starts from the same nc points (as the reference's mains): chain j at mu_{j mod K}
    std::vector<Real> pinit((size_t)o.chains * nparam);
    for (int c = 0; c < o.chains; ++c)
      for (int i = 0; i < nparam; ++i) pinit[(size_t)c * nparam + i] = mu[(size_t)(c % K) * nparam + i];
    std::vector<Real> incov((size_t)nparam * nparam, 0.0);
    for (int i = 0; i < nparam; ++i) incov[(size_t)i * nparam + i] = 2.38 * 2.38 / nparam;
    if (mcpar.run(o.nsamp, 200, &pinit[0], L, rslts, &incov[0]) != MCPar::OK) return 2;
    rslts.output();
    Real lmax;
    const std::vector<Real> &pmax = rslts.maxlike(&lmax);
    std::cerr << "max likelihood value: " << lmax << "  (p0 = " << pmax[0] << ")  acceptance " << mcpar.last_accept_rate()
              << "  device " << mcpar.last_device_ms() << " ms\n";
  } catch (const char *msg) {
    std::cerr << msg << "\n";
    return 1;
  }
  return 0;
}
