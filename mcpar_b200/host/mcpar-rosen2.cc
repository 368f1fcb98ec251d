// mcpar-rosen2 -- the higher-dimensional Rosenbrock run BASELINE.json names (d = 16,
// Rosenbrock1(16), tuned proposal covariance).  The file of this name in the reference is
// a stale copy of the 2-D demo that no longer compiles (src/mcpar-rosen2.cc:16,44); this
// driver keeps its shape: `mcpar-rosen2 [nsamp]`, "nsamp = N", rows on stdout.
#include <iostream>
#include <fstream>
#include <vector>
#include <stdlib.h>
#include <string.h>
#include "mcpar.hh"
#include "rosenbrock.hh"
#include "mcout.hh"
#include "driver_opts.hh"

int main(int argc, char *argv[])
{
  const int nparam = 16;
  DriverOpts o(10000);
  o.parse(argc, argv);
  const int nsamp = o.nsamp, ranks = o.ranks;
  const char *binfile = o.binfile;
  try {
    Rosenbrock1 L(nparam);
    std::ofstream bin;
    if (binfile) {
      bin.open(binfile, std::ios::binary);
      if (!bin) { std::cerr << "cannot open " << binfile << "\n"; return 1; }
    }
    MCout rslts(nparam, binfile ? static_cast<std::ostream *>(&bin) : &std::cout, 0);
    if (binfile) rslts.set_format(MCout::BINARY);
    std::cout << "nsamp = " << nsamp << "\n";
    MCPar mcpar(nparam, 4, ranks, 0);
    o.apply(mcpar);
    // the 2-D demo's four starting points, repeated over the 8 coordinate pairs
    const Real p4[8] = {0.0, 0.0, 2.0, 2.0, 0.0, 1.5, 0.0, -2.0};
    std::vector<Real> pinit(4 * nparam);
    for (int c = 0; c < 4; ++c)
      for (int i = 0; i < nparam; ++i) pinit[c * nparam + i] = p4[2 * c + (i & 1)];
    // proposal covariance: block-diagonal copies of the analytic 2-D target covariance
    // [[1/2, 1], [1, 2.505]] at the Roberts-Rosenthal scale 2.38^2/d
    std::vector<Real> incov(nparam * nparam, 0.0);
    const Real s = 2.38 * 2.38 / nparam;
    for (int b = 0; b < nparam / 2; ++b) {
      incov[(2 * b) * nparam + 2 * b] = s * 0.5;      incov[(2 * b) * nparam + 2 * b + 1] = s * 1.0;
      incov[(2 * b + 1) * nparam + 2 * b] = s * 1.0;  incov[(2 * b + 1) * nparam + 2 * b + 1] = s * 2.505;
    }
    if (mcpar.run(nsamp, 500, &pinit[0], L, rslts, &incov[0]) != MCPar::OK) return 2;
    rslts.output();
  } catch (const char *msg) {
    std::cerr << msg << "\n";
    return 1;
  }
  return 0;
}
