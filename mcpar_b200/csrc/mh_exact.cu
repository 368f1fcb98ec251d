// mh_exact.cu -- exact-arithmetic instantiations (compile with -fmad=false): the
// replay-stream production kernel, the verification-mode kernel and the generic
// batched likelihood evaluation.  Bit-comparable with the CPU oracle.
#define MCGPU_NS exact
#define MCGPU_EXACT_TU 1
#include <stdlib.h>
#include "mh_kernels.cuh"
#include "mh_hostlik.cuh"
namespace mcgpu { namespace exact {
#include "mh_dispatch.inl"

cudaError_t launch_verify(const VerifyParams &p, int nranks_local, cudaStream_t st)
{
  const int block = ((p.C + 31) / 32) * 32;
  mh_verify_kernel<<<nranks_local, block, 0, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_hostlik_propose(const HostLikParams &p, cudaStream_t st)
{
  hostlik_propose_kernel<<<(unsigned)((p.C + 127) / 128), 128, 0, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_hostlik_accept(const HostLikParams &p, int publish, double pub_winv, cudaStream_t st)
{
  hostlik_accept_kernel<<<(unsigned)((p.C + 127) / 128), 128, 0, st>>>(p, publish, pub_winv);
  return cudaGetLastError();
}

cudaError_t launch_loglik_aos(const LikSpec &L, const double *x, double *y, int npset, cudaStream_t st)
{
  loglik_aos_kernel<<<(npset + 255) / 256, 256, 0, st>>>(L, x, y, npset);
  return cudaGetLastError();
}
}}
