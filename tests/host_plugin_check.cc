// tests/host_plugin_check.cc -- a USER-WRITTEN VLFunc through MCPar::run (built and run by tests/test_gpu_drivers.py).
// The plugin below is plain host C++ (the reference's Rosenbrock1 loop, src/rosenbrock.cc:4-21, typed in again as a
// user would); MCPar::run must call it on the host once per step and keep everything else on the GPU.  The same run
// with the built-in device functor is the reference: both consume the same counter-based draws, so the two
// histories agree row for row up to the rounding of the Box-Muller math (table-driven in the fused kernels, CUDA
// libm in the split path).  Prints "ok <fraction of equal rows> <plugin calls>".
#include <iostream>
#include <cmath>
#include <cstdlib>
#include "mcpar.hh"
#include "rosenbrock.hh"
#include "mcout.hh"

class UserRosenbrock : public VLFunc {
public:
  int n; long calls;
  explicit UserRosenbrock(int nc) : n(nc), calls(0) {}
  int operator()(int npset, const Real *x, Real *restrict fx) {
    ++calls;
    const int ntot = npset * n;
    for (int j = 0; j < npset; ++j) fx[j] = 0.0;
    for (int i = 0; i < ntot - 1; i += 2) {
      const int j = i / n;
      const Real t1 = 1 - x[i];
      const Real t2 = x[i + 1] - x[i] * x[i];
      fx[j] -= t1 * t1 + 100.0 * t2 * t2;
    }
    return 0;
  }
};

int main(int argc, char *argv[])
{
  const int ranks = argc > 1 ? atoi(argv[1]) : 16, nsamp = argc > 2 ? atoi(argv[2]) : 60, nburn = 120;
  const int remote_mode = argc > 3 ? atoi(argv[3]) : 0;
  const Real pinit[8] = {0.0, 0.0, 2.0, 2.0, 0.0, 1.5, 0.0, -2.0};
  MCout a(2, 0, 0), b(2, 0, 0);
  UserRosenbrock U(2);
  Rosenbrock1 D(2);
  {
    MCPar m(2, 4, ranks, 0, 0.7);
    m.pool_m = 8; m.remote_mode = remote_mode;
    if (m.run(nsamp, nburn, pinit, U, a) != MCPar::OK) { std::cout << "host run failed\n"; return 1; }
  }
  {
    MCPar m(2, 4, ranks, 0, 0.7);
    m.pool_m = 8; m.remote_mode = remote_mode;
    if (m.run(nsamp, nburn, pinit, D, b) != MCPar::OK) { std::cout << "device run failed\n"; return 1; }
  }
  if (a.size() != b.size() || a.size() != nsamp * 4 * ranks) { std::cout << "size mismatch " << a.size() << " " << b.size() << "\n"; return 1; }
  long same = 0;
  for (int r = 0; r < a.size(); ++r) {
    const Real *p = a.getpset(r), *q = b.getpset(r);
    bool eq = std::fabs(a.getlval(r) - b.getlval(r)) <= 1e-7 * (1 + std::fabs(b.getlval(r)));
    for (int i = 0; i < 2; ++i) eq = eq && std::fabs(p[i] - q[i]) <= 1e-9 + 1e-7 * std::fabs(q[i]);
    same += eq;
  }
  // nburn + nsamp steps + the initial L(nchain, pvals, lylast), one call per rank each (mcpar.cc:53, :60, :160)
  const long expect = (long)(nburn + nsamp + 1) * ranks;
  std::cout << ((double)same / a.size() > 0.995 && U.calls == expect ? "ok " : "MISMATCH ") << (double)same / a.size() << " " << U.calls << "\n";
  return 0;
}
