// mcpar-bench -- chain-count sweep through the C++ MCPar mirror (BASELINE.json config 5: 2-D Rosenbrock,
// 1K-16M chains; SURVEY.md section 7 step 6 names the driver, the reference has none).  For each total chain
// count N it runs MCPar(2, 32 chains per rank, N/32 ranks) with nburn 500 + nsamp 1000, thin 10, PLOCAL 0.9, pool
// M = min(N, 256), into an MCout without an output stream, and prints one line
//     N  chain-steps/s(device)  acceptance  exchange-wait-ms
// `mcpar-bench [--ngpu=G] [--remote-mode=1] [--lag=1] [--lik=rosen1|dgauss] [--nsamp=S] [N ...]`
#include <iostream>
#include <vector>
#include <stdio.h>
#include "mcpar.hh"
#include "rosenbrock.hh"
#include "mcout.hh"
#include "driver_opts.hh"

int main(int argc, char *argv[])
{
  DriverOpts o(1000);
  o.thin = 10; o.pool = 256;
  o.parse(argc, argv);
  o.nsamp = 1000;                                   // here the positional arguments are chain counts, not nsamp
  bool dgauss = false;
  std::vector<long long> sizes;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--lik=dgauss")) dgauss = true;
    else if (!strncmp(argv[i], "--nsamp=", 8)) o.nsamp = atoi(argv[i] + 8);
    else if (argv[i][0] != '-') sizes.push_back(atoll(argv[i]));
  }
  if (sizes.empty()) for (int e = 10; e <= 20; e += 2) sizes.push_back(1ll << e);
  try {
    Rosenbrock1 R(2); DualGaussian D(5.0);
    const Real p4[8] = {0.0, 0.0, 2.0, 2.0, 0.0, 1.5, 0.0, -2.0};
    std::vector<Real> pinit(64);
    for (int c = 0; c < 32; ++c) { pinit[2 * c] = p4[2 * (c % 4)]; pinit[2 * c + 1] = p4[2 * (c % 4) + 1]; }
    printf("# chains  chain-steps/s  accept  exchange_wait_ms   (nburn 500, nsamp %d, thin %d, remote mode %d, lag %d, %d GPU)\n",
           o.nsamp, o.thin, o.remote_mode, o.lag, o.ngpu);
    for (size_t k = 0; k < sizes.size(); ++k) {
      const long long N = sizes[k];
      if (N < 32 || N % 32) { fprintf(stderr, "chain count %lld must be a multiple of 32\n", N); continue; }
      MCout rs(2, 0, 0);                            // no output stream: rows reach MCout's store and are dropped with it
      MCPar m(2, 32, (int)(N / 32), 0);
      o.apply(m);
      m.pool_m = (int)(N < o.pool ? 0 : o.pool);
      m.coin_group = 0;
      if (m.run(o.nsamp, 500, &pinit[0], dgauss ? (VLFunc &)D : (VLFunc &)R, rs) != MCPar::OK) return 2;
      printf("%lld  %.4g  %.3f  %.3f\n", N, (double)N * (500 + o.nsamp) / (m.last_device_ms() * 1e-3), m.last_accept_rate(),
             m.last_exchange_wait_ms());
      fflush(stdout);
    }
  } catch (const char *msg) {
    std::cerr << msg << "\n";
    return 1;
  }
  return 0;
}
