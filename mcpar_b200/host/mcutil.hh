// mcutil.hh -- utility class with the reference's interface (src/mcutil.hh:11-31).
#ifndef MCPAR_B200_MCUTIL_HH_
#define MCPAR_B200_MCUTIL_HH_
#include "vlfunc.hh"

class mcutil {
public:
  int device;
  mcutil() : device(0) {}
  // quasi-random (Sobol) initial guesses in the box [plo, phi], generated on the GPU;
  // rank r skips the r*npset*nparam scalars of the ranks before it (src/mcutil.cc:16-31)
  void qriguess(int rank, int npset, int nparam, const Real plo[], const Real phi[], Real *restrict pout);
};

#endif
