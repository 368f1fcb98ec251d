// mcpar-dgauss -- double-Gaussian demo with the reference's contract (src/mcpar-dgauss.cc):
// no arguments; DualGaussian(5), 4 chains per rank, run(8, 500); rows on stdout, the
// parameter columns tab-separated in mcpar-dgauss.RRR.txt, then the maximum-likelihood row.
#include <iostream>
#include <fstream>
#include <sstream>
#include <iomanip>
#include <stdlib.h>
#include <string.h>
#include "mcpar.hh"
#include "rosenbrock.hh"
#include "mcout.hh"
#include "driver_opts.hh"

int main(int argc, char *argv[])
{
  const int nparam = 2;
  DriverOpts o(8);
  o.parse(argc, argv);
  for (int i = 1; i < argc; ++i) if (!strncmp(argv[i], "--nsamp=", 8)) o.nsamp = atoi(argv[i] + 8);
  const int nsamp = o.nsamp, ranks = o.ranks;
  try {
    DualGaussian L(5.0);
    MCout rslts(nparam, &std::cout, 0);
    MCPar mcpar(nparam, 4, ranks, 0);
    o.apply(mcpar);
    Real pinit[8] = {0.0, 0.0, 2.0, 2.0, 0.0, 1.5, 0.0, -2.0};
    if (mcpar.run(nsamp, 500, pinit, L, rslts) != MCPar::OK) return 2;

    // per-rank files, as each MPI rank wrote its own (mcpar-dgauss.cc:38-47): rank r's rows
    const int per_rank = rslts.size() / ranks;
    for (int r = 0; r < ranks; ++r) {
      std::stringstream ofname;
      ofname << "mcpar-dgauss." << std::setfill('0') << std::setw(3) << r << ".txt";
      std::ofstream outfile(ofname.str().c_str());
      // MCout order is batch -> rank -> step -> chain; pick rank r's rows out of every batch
      const int outstep = nsamp > 50 ? nsamp / 10 : 5;
      int base = 0;
      for (int done = 0; done < nsamp; done += outstep) {
        const int n = (nsamp - done < outstep) ? nsamp - done : outstep;
        const int blk = n * 4;
        for (int i = 0; i < blk; ++i) {
          const Real *pset = rslts.getpset(base + r * blk + i);
          for (int j = 0; j < rslts.ncol() - 1; ++j) outfile << pset[j] << "\t";
          outfile << "\n";
        }
        base += blk * ranks;
      }
      (void)per_rank;
    }
    Real lmax;
    const std::vector<Real> &pmax = rslts.maxlike(&lmax);
    std::cout << "max likelihood value: " << lmax << "\n";
    for (size_t i = 0; i < pmax.size(); ++i) std::cout << pmax[i] << "  ";
    std::cout << "\n";
  } catch (const char *msg) {
    std::cerr << msg << "\n";
    return 1;
  }
  return 0;
}
