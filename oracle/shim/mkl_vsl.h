/* oracle/shim/mkl_vsl.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Stand-in for Intel MKL VSL.  The real generators (MT2203, Sobol, BoxMuller2) are
 * third-party arithmetic that is NOT under /root/reference (MKL 15.0.1 per
 * scripts/mcpar-rosen1.sh:8), so their bit streams are unpinned.  The shim gives
 * every stream one of two back ends, chosen by the harness per rank:
 *   REPLAY : normals / uniforms / ints are consumed sequentially from supplied
 *            arrays (the north star's "verification mode");
 *   PHILOX : host Philox4x32-10 + Box-Muller (for timing / statistical runs).
 * Call sites served: mcpar.cc:63,146,163,270-271,277,306,337,348,401;
 * mcutil.cc:16,23,25.
 */
#ifndef ORACLE_SHIM_MKL_VSL_H_
#define ORACLE_SHIM_MKL_VSL_H_
#include <stddef.h>

typedef void *VSLStreamStatePtr;

#define VSL_STATUS_OK                        0
#define VSL_BRNG_MT2203                      0x00100000
#define VSL_BRNG_SOBOL                       0x00200000
#define VSL_RNG_METHOD_UNIFORM_STD           0
#define VSL_RNG_METHOD_GAUSSIAN_BOXMULLER2   1
#define VSL_MATRIX_STORAGE_FULL              0
#define VSL_MATRIX_STORAGE_DIAGONAL          2

#ifdef __cplusplus
extern "C" {
#endif
int vslNewStream(VSLStreamStatePtr *stream, int brng, unsigned int seed);
int vslDeleteStream(VSLStreamStatePtr *stream);
int vslSkipAheadStream(VSLStreamStatePtr stream, long long nskip);
int vsRngUniform(int method, VSLStreamStatePtr stream, int n, float *r, float a, float b);
int viRngUniform(int method, VSLStreamStatePtr stream, int n, int *r, int a, int b);
int vsRngGaussianMV(int method, VSLStreamStatePtr stream, int n, float *r, int dimen,
                    int mstorage, const float *a, const float *t);
#ifdef __cplusplus
}
#endif

#ifdef __cplusplus
/* harness-side control (not part of VSL) */
namespace shim_vsl {
struct Source {
  int mode;                    /* 0 = REPLAY, 1 = PHILOX */
  /* REPLAY */
  const double *Z; size_t nz, iz;
  const double *U; size_t nu, iu;
  const int    *I; size_t ni, ii;
  int overrun;                 /* set if a stream ran dry (values past the end read as 0) */
  /* PHILOX */
  unsigned long long seed, stream_id, ctr;
  double spare; int has_spare;
};
/* the Source the calling thread's next vslNewStream(MT2203+rank) binds to */
void bind_thread_source(Source *s);
Source *thread_source();
}
#endif
#endif
