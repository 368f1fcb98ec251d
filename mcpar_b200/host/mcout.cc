// mcout.cc -- sample sink (interface of src/mcout.hh; behaviour of src/mcout.cc:8-145 with
// the MPI gather replaced by rows that arrive from the GPU history already merged).
#include "mcout.hh"
#include <limits>
#include <assert.h>

MCout::MCout(int np, std::ostream *aoutstream, MCComm)
  : maxlparams(np), maxlval(-std::numeric_limits<Real>::infinity()), mnparam(np), mncol(np + 1),
    next(0), npset(0), maxsamps(0), nextout(0), outstream(aoutstream), mformat(TEXT), wrote_header(false)
{
}

void MCout::newsamps(size_t nsamp)
{
  maxsamps += nsamp;
  pvals.resize(pvals.size() + nsamp * (size_t)mncol);
}

void MCout::add(const Real *pv, Real lval)
{
  assert(next + (size_t)mncol <= pvals.size());          // the caller sizes the store (newsamps)
  Real *row = &pvals[next];
  for (int i = 0; i < mnparam; ++i) row[i] = pv[i];
  row[mnparam] = lval;
  next += (size_t)mncol;
  ++npset;
  if (lval > maxlval) {                                  // first maximum wins
    maxlval = lval;
    for (int i = 0; i < mnparam; ++i) maxlparams[i] = pv[i];
  }
}

void MCout::addrows(const Real *rows, size_t nrows)
{
  for (size_t r = 0; r < nrows; ++r) add(rows + r * (size_t)mncol, rows[r * (size_t)mncol + mnparam]);
}

// Text format of the reference: every column followed by two blanks, one row per line,
// default ostream precision (6 significant digits).
void MCout::output()
{
  size_t ntot = 0;
  Real *buf = collect(&ntot);
  if (!buf) return;
  if (outstream && mformat == BINARY) {
    std::ostream &os = *outstream;
    if (!wrote_header) {
      const int hdr[2] = {mncol, (int)sizeof(Real)};
      os.write("MCOUTB01", 8); os.write(reinterpret_cast<const char *>(hdr), sizeof hdr);
      wrote_header = true;
    }
    os.write(reinterpret_cast<const char *>(buf), (std::streamsize)(ntot * sizeof(Real)));
  } else if (outstream) {
    std::ostream &os = *outstream;
    const size_t nrow = ntot / (size_t)mncol;
    for (size_t r = 0; r < nrow; ++r) {
      for (int j = 0; j < mncol; ++j) os << buf[r * mncol + j] << "  ";
      os << "\n";
    }
  }
  delete[] buf;
}

Real *MCout::collect(size_t *ntot)
{
  *ntot = 0;
  if (next <= nextout) return 0;
  const size_t n = next - nextout;
  Real *buf = new Real[n];
  for (size_t i = 0; i < n; ++i) buf[i] = pvals[nextout + i];
  nextout = next;
  *ntot = n;
  return buf;
}

const std::vector<Real> &MCout::maxlike(Real *lmax)
{
  *lmax = maxlval;
  return maxlparams;
}
