// mcout.hh -- the sample sink with the reference's interface (src/mcout.hh:13-51).
// One process drives every GPU, so the MPI communicator argument is an ignored handle and
// the collective calls (collect, maxlike) are local.
#ifndef MCPAR_B200_MCOUT_HH_
#define MCPAR_B200_MCOUT_HH_
#include <vector>
#include <iostream>
#include <stddef.h>
#include "vlfunc.hh"

typedef int MCComm;                                    // stands in for MPI_Comm

class MCout {
  std::vector<Real> pvals;                             // rows (p_0..p_{d-1}, logL)
  std::vector<Real> maxlparams;
  Real maxlval;
  const int mnparam, mncol;
  size_t next;                                         // elements stored
  size_t npset, maxsamps;
  size_t nextout;                                      // first element not yet output
  std::ostream *outstream;
  int mformat;
  bool wrote_header;
public:
  // Output formats: TEXT is the reference's ("v  v  v  \n", 6 significant digits; src/mcout.cc:37-47).
  // BINARY is for runs whose text would not fit anywhere (10^6 chains x 10^3 steps): a 16-byte header
  // "MCOUTB01", int32 ncol, int32 sizeof(Real), then the rows as raw reals in the same order.
  enum Format { TEXT = 0, BINARY = 1 };
  void set_format(Format f) { mformat = f; }
  MCout(int np, std::ostream *aoutstream = 0, MCComm acomm = 0);
  void newsamps(size_t nsamp);                         // reserve room for nsamp more parameter sets
  void add(const Real *pv, Real lval);                 // append one row, track the max-likelihood row
  void addrows(const Real *rows, size_t nrows);        // bulk append of (p..., logL) rows
  int size(void) const { return (int)npset; }
  int maxsize(void) const { return (int)maxsamps; }
  int ncol(void) { return mncol; }
  int nparam(void) const { return mnparam; }
  int vsize(void) const { return (int)pvals.size(); }
  const Real *getpset(int i) const { return &pvals[(size_t)i * mncol]; }
  Real getlval(int i) const { return pvals[(size_t)(i + 1) * mncol - 1]; }
  void output();                                       // print rows added since the last output
  Real *collect(size_t *ntot);                         // the same rows in a new[] buffer (caller deletes)
  void rewind(void) { nextout = 0; }
  const std::vector<Real> &maxlike(Real *lmax);
};

#endif
