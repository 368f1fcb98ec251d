// mcpar-rosen1 -- 2-D Rosenbrock demo with the reference's command line and output
// (src/mcpar-rosen1.cc): `mcpar-rosen1 [nsamp]` prints "nsamp = N" and then every sample
// row; 4 chains per rank from the four fixed starting points.  `mpirun -np R` becomes
// --ranks=R (all ranks live on the GPU); extra flags never disturb the positional form.
// --binary=FILE sends the rows to FILE in MCout's binary format instead of stdout (text for 10^6 chains
// does not fit anywhere); stdout then carries only the banner.
#include <iostream>
#include <fstream>
#include <stdlib.h>
#include <string.h>
#include "mcpar.hh"
#include "rosenbrock.hh"
#include "mcout.hh"
#include "driver_opts.hh"

int main(int argc, char *argv[])
{
  const int nparam = 2;
  DriverOpts o(100000);
  o.parse(argc, argv);
  const int nsamp = o.nsamp, ranks = o.ranks;
  const char *binfile = o.binfile;
  try {
    Rosenbrock1 L(2);
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
    Real pinit[8] = {0.0, 0.0, 2.0, 2.0, 0.0, 1.5, 0.0, -2.0};
    if (mcpar.run(nsamp, 500, pinit, L, rslts) != MCPar::OK) return 2;
    rslts.output();
  } catch (const char *msg) {
    std::cerr << msg << "\n";
    return 1;
  }
  return 0;
}
