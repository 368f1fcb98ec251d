// mh_fast.cu -- production instantiations of the fused step kernels (FMA contraction on).
#define MCGPU_NS fast
#include <stdlib.h>
#include "mh_kernels.cuh"
#include "mh_wide.cuh"
namespace mcgpu { namespace fast {
#include "mh_dispatch.inl"

// ---- wide (d >= 8, D/2 lanes per chain) kernels ------------------------------------------
bool wide_supported(int lik, int d)
{
  return (lik == MCGPU_ROSENBROCK1 || lik == MCGPU_GAUSSMIX) && (d == 8 || d == 16 || d == 32 || d == 64);
}

// chains per lane group: register-level reuse of the streamed GaussMix / pool parameters
#ifndef MCGPU_WIDE_NCH64
#define MCGPU_WIDE_NCH64 4
#endif
template <int D> struct WideNch { static constexpr int value = D >= 64 ? MCGPU_WIDE_NCH64 : (D >= 32 ? 2 : 1); };

template <int LIK, int D>
static cudaError_t launch_wide_d(int phase, const WideParams &p, cudaStream_t st)
{
  constexpr int L = D / 2, NCH = WideNch<D>::value;
  const int block = 128, gpb = block / L, cpb = gpb * NCH;
  const unsigned grid = (unsigned)((p.C + cpb - 1) / cpb);
  const size_t smem = sizeof(double) * ((size_t)MCGPU_MATH_SMEM + (size_t)gpb * 2 * D * NCH);
  switch (phase) {
    case PH_BURN:   mh_wide_kernel<LIK, D, NCH, PH_BURN><<<grid, block, smem, st>>>(p); break;
    case PH_LOCAL:  mh_wide_kernel<LIK, D, NCH, PH_LOCAL><<<grid, block, smem, st>>>(p); break;
    case PH_REMOTE: mh_wide_kernel<LIK, D, NCH, PH_REMOTE><<<grid, block, smem, st>>>(p); break;
    case PH_REMOTE_SUM: mh_wide_kernel<LIK, D, NCH, PH_REMOTE_SUM><<<grid, block, smem, st>>>(p); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

template <int LIK>
static cudaError_t launch_wide_lik(int d, int phase, const WideParams &p, cudaStream_t st)
{
  switch (d) {
    case 8:  return launch_wide_d<LIK, 8>(phase, p, st);
    case 16: return launch_wide_d<LIK, 16>(phase, p, st);
    case 32: return launch_wide_d<LIK, 32>(phase, p, st);
    case 64: return launch_wide_d<LIK, 64>(phase, p, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_wide(int lik, int d, int phase, const WideParams &p, cudaStream_t st)
{
  if (lik == MCGPU_ROSENBROCK1) return launch_wide_lik<MCGPU_ROSENBROCK1>(d, phase, p, st);
  if (lik == MCGPU_GAUSSMIX) return launch_wide_lik<MCGPU_GAUSSMIX>(d, phase, p, st);
  return cudaErrorInvalidValue;
}

cudaError_t launch_pool_prep(const double *pool, int M, int mpad, int D, double2 *pmh, double *psd, double *pnb,
                             float2 *pf, float *pnbf, float *pscal,
                             const unsigned long long *arrivals, unsigned long long wait_target, int *xflag,
                             unsigned long long *xstat, cudaStream_t st)
{
  cudaMemsetAsync(pscal, 0, 4 * sizeof(float), st);
  pool_prep_kernel<<<(D * mpad + 127) / 128, 128, 0, st>>>(pool, M, mpad, D, pmh, psd, pnb, pf, pnbf, pscal, arrivals, wait_target, xflag, xstat);
  return cudaGetLastError();
}

cudaError_t launch_factor_prep(const double *rm, double *cm, int D, int *diag, cudaStream_t st)
{
  factor_prep_kernel<<<1, 256, 0, st>>>(rm, cm, D, diag);
  return cudaGetLastError();
}
}}
