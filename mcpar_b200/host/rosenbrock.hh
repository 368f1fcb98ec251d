// rosenbrock.hh -- the built-in likelihoods with the reference's constructors
// (src/rosenbrock.hh:7-62).  Arithmetic lives in the device functors
// (csrc/mh_kernels.cuh: Lik<...>, loglik_aos); these classes only carry parameters.
#ifndef MCPAR_B200_ROSENBROCK_HH_
#define MCPAR_B200_ROSENBROCK_HH_
#include "vlfunc.hh"

class Rosenbrock1 : public DeviceVLFunc {             // non-overlapping pairs, rosenbrock.cc:4-21
  const int n;
public:
  explicit Rosenbrock1(int nc);                        // throws const char* unless n is even and >= 2
  int lik_id() const; int nparam() const { return n; }
};

class Rosenbrock2 : public DeviceVLFunc {             // overlapping pairs, rosenbrock.cc:25-41 (quirks kept)
  const int n;
public:
  explicit Rosenbrock2(int nc);                        // throws const char* unless n >= 2
  int lik_id() const; int nparam() const { return n; }
};

class Gaussian : public DeviceVLFunc {                // 2-D diagonal Gaussian, rosenbrock.cc:44-61
  const int n;
public:
  Gaussian(int nc, const Real muin[] = 0, const Real sig2[] = 0);   // throws const char* unless nc == 2
  int lik_id() const; int nparam() const { return n; }
};

class DualGaussian : public DeviceVLFunc {            // w N((0,0),I) + N((5,5),I), rosenbrock.cc:63-78
public:
  explicit DualGaussian(Real win);
  int lik_id() const; int nparam() const { return 2; }
};

// New: K-component diagonal Gaussian mixture in d dimensions, log-sum-exp.
class GaussMix : public DeviceVLFunc {
  const int n, K;
public:
  GaussMix(int nparam, int ncomp, const Real *mu, const Real *sig2, const Real *w);
  int lik_id() const; int nparam() const { return n; }
};

#endif
