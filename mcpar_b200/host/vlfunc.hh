// vlfunc.hh -- the likelihood plugin surface, kept from the reference (src/vlfunc.hh:9-12):
//   virtual int operator()(int npset, const Real *x, Real *restrict y) = 0;
// x holds npset parameter sets, chain-major {a1,b1,...,a2,b2,...}; y receives npset
// LOG-likelihoods; the functor knows nparam itself; the return code is reserved.
// Real is double here (the reference is float; see DESIGN.md "precision").
#ifndef MCPAR_B200_VLFUNC_HH_
#define MCPAR_B200_VLFUNC_HH_
#include <vector>

#ifndef restrict
#define restrict __restrict__            /* the reference gets this from -Drestrict=__restrict__ (src/Makefile:12) */
#endif

typedef double Real;

class VLFunc {
public:
  virtual int operator()(int npset, const Real *x, Real *restrict y) = 0;
  virtual ~VLFunc() {}
};

// A likelihood the B200 engine can run inside its fused step kernel.  The batched
// operator() is also served by the GPU (mcgpu_loglik); there is no host evaluation.
class DeviceVLFunc : public VLFunc {
public:
  virtual int lik_id() const = 0;                       // MCGPU_* id of the device functor
  virtual int nparam() const = 0;
  virtual const std::vector<double> &params() const { return mpar; }
  int operator()(int npset, const Real *x, Real *restrict y);     // evaluates on the GPU
  int device;                                           // CUDA ordinal used by operator()
protected:
  DeviceVLFunc() : device(0) {}
  std::vector<double> mpar;
};

#endif
