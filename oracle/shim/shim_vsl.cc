/* oracle/shim/shim_vsl.cc -- TEST INFRASTRUCTURE, not product code.
 *
 * Stand-in for the MKL VSL / LAPACK entry points the reference calls.  MKL is a
 * third-party dependency absent from /root/reference (version named only in
 * scripts/mcpar-rosen1.sh:8: mkl/15.0.1), so generator bit streams are UNPINNED;
 * what IS pinned here is the documented transform of each call:
 *   vsRngUniform(a,b)            r = a + (b-a)*u,            u in [0,1)
 *   viRngUniform(a,b)            r in [a,b)
 *   vsRngGaussianMV FULL         r_i = a_i + sum_{k<=i} T[i*d+k] z_k   (T row-major lower)
 *   vsRngGaussianMV DIAGONAL     r_i = a_i + t_i z_i
 *   spotrf('U') column-major  == row-major lower Cholesky factor in place
 * In REPLAY mode u, z and the ints come from caller-supplied arrays.
 */
#include "mkl.h"
#include "mkl_vsl.h"
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

namespace {
thread_local shim_vsl::Source *t_src = nullptr;

struct Stream {
  int brng;
  shim_vsl::Source *src;      /* MT2203 family streams */
  /* Sobol state */
  int dimen; unsigned long long index;
  std::vector<uint32_t> dirs; /* dimen x 32 direction numbers */
  std::vector<uint32_t> x;    /* current gray-code state per dimension */
  int cur;                    /* next dimension to emit */
};

/* ---- Philox4x32-10 (Salmon et al. 2011), host side of the shim RNG ---- */
inline void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
/* sequential 53-bit uniforms in [0,1): two per Philox block */
inline double next_u(shim_vsl::Source *s) {
  uint32_t c[4] = {(uint32_t)(s->ctr >> 1), (uint32_t)((s->ctr >> 1) >> 32),
                   (uint32_t)s->stream_id, (uint32_t)(s->stream_id >> 32)};
  philox(c, (uint32_t)s->seed, (uint32_t)(s->seed >> 32));
  int h = (int)(s->ctr & 1ull);
  ++s->ctr;
  uint64_t bits = ((uint64_t)c[2*h] << 32) | c[2*h+1];
  return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}
inline double next_z(shim_vsl::Source *s) {
  if (s->mode == 0) {
    if (s->iz >= s->nz) { s->overrun = 1; ++s->iz; return 0.0; }
    return s->Z[s->iz++];
  }
  if (s->has_spare) { s->has_spare = 0; return s->spare; }
  double u1 = 1.0 - next_u(s);          /* (0,1] */
  double u2 = next_u(s);
  double r = sqrt(-2.0 * log(u1));
  double a = 6.283185307179586476925 * u2;
  s->spare = r * cos(a); s->has_spare = 1;
  return r * sin(a);
}
inline double next_uniform(shim_vsl::Source *s) {
  if (s->mode == 0) {
    if (s->iu >= s->nu) { s->overrun = 1; ++s->iu; return 0.0; }
    return s->U[s->iu++];
  }
  return next_u(s);
}

/* Sobol direction numbers: Joe & Kuo (2008) new-joe-kuo-6 primitive polynomials
 * and initial m_i for the first 16 dimensions (MKL's own table is unobtainable;
 * dimension 1..n of any Sobol sequence share the radical-inverse first axis). */
struct JK { int s; uint32_t a; uint32_t m[7]; };
const JK jk[] = {
  {0,0,{0}},                               /* dim 1: van der Corput */
  {1,0,{1}}, {2,1,{1,3}}, {3,1,{1,3,1}}, {3,2,{1,1,1}}, {4,1,{1,1,3,3}},
  {4,4,{1,3,5,13}}, {5,2,{1,1,5,5,17}}, {5,4,{1,1,5,5,5}}, {5,7,{1,1,7,11,19}},
  {5,11,{1,1,5,1,1}}, {5,13,{1,1,1,3,11}}, {5,14,{1,3,5,5,31}}, {6,1,{1,3,3,9,7,49}},
  {6,13,{1,1,1,15,21,21}}, {6,16,{1,3,1,13,27,49}},
};
void sobol_init(Stream *st) {
  const int D = st->dimen;
  st->dirs.assign((size_t)D * 32, 0); st->x.assign(D, 0); st->index = 0; st->cur = 0;
  for (int d = 0; d < D; ++d) {
    uint32_t *v = &st->dirs[(size_t)d * 32];
    if (d == 0) { for (int i = 0; i < 32; ++i) v[i] = 1u << (31 - i); continue; }
    if (d >= (int)(sizeof(jk)/sizeof(jk[0]))) { fprintf(stderr, "shim Sobol: dimen > 16 unsupported\n"); abort(); }
    const JK &p = jk[d]; const int s = p.s;
    for (int i = 0; i < 32; ++i) {
      if (i < s) v[i] = p.m[i] << (31 - i);
      else {
        v[i] = v[i-s] ^ (v[i-s] >> s);
        for (int k = 1; k < s; ++k) v[i] ^= (((p.a >> (s-1-k)) & 1u) * v[i-k]);
      }
    }
  }
}
/* emit the next scalar of the interleaved (point-major) Sobol stream */
double sobol_next(Stream *st) {
  if (st->cur == 0 && st->index > 0) {     /* advance to the next point (gray code) */
    unsigned long long n = st->index - 1; int c = 0;
    while (n & 1ull) { n >>= 1; ++c; }
    for (int d = 0; d < st->dimen; ++d) st->x[d] ^= st->dirs[(size_t)d * 32 + c];
  }
  double r = (double)st->x[st->cur] * (1.0 / 4294967296.0);
  if (++st->cur == st->dimen) { st->cur = 0; ++st->index; }
  return r;
}
}

namespace shim_vsl {
void bind_thread_source(Source *s) { t_src = s; }
Source *thread_source() { return t_src; }
}

extern "C" {

int vslNewStream(VSLStreamStatePtr *stream, int brng, unsigned int seed) {
  Stream *st = new Stream();
  st->brng = brng; st->src = nullptr; st->dimen = 0; st->index = 0; st->cur = 0;
  if (brng >= VSL_BRNG_SOBOL) {
    /* for Sobol the "seed" argument is the dimension (mcutil.cc:16) */
    st->brng = VSL_BRNG_SOBOL; st->dimen = (int)seed; sobol_init(st);
  } else {
    st->src = t_src;          /* MT2203+rank family: bind to the rank thread's source */
  }
  *stream = st;
  return VSL_STATUS_OK;
}
int vslDeleteStream(VSLStreamStatePtr *stream) {
  delete (Stream*)*stream; *stream = nullptr; return VSL_STATUS_OK;
}
}

namespace {
inline Stream *ready(VSLStreamStatePtr p) { return (Stream*)p; }
}

extern "C" {

int vslSkipAheadStream(VSLStreamStatePtr stream, long long nskip) {
  Stream *st = ready(stream);
  if (st->brng == VSL_BRNG_SOBOL) { for (long long i = 0; i < nskip; ++i) (void)sobol_next(st); return VSL_STATUS_OK; }
  for (long long i = 0; i < nskip; ++i) (void)next_uniform(st->src);
  return VSL_STATUS_OK;
}

int vsRngUniform(int, VSLStreamStatePtr stream, int n, float *r, float a, float b) {
  Stream *st = ready(stream);
  if (st->brng == VSL_BRNG_SOBOL) {
    for (int i = 0; i < n; ++i) r[i] = a + (b - a) * (float)sobol_next(st);
    return VSL_STATUS_OK;
  }
  if (!st->src) return 1;
  for (int i = 0; i < n; ++i) r[i] = a + (b - a) * (float)next_uniform(st->src);
  return VSL_STATUS_OK;
}

int viRngUniform(int, VSLStreamStatePtr stream, int n, int *r, int a, int b) {
  Stream *st = ready(stream);
  shim_vsl::Source *s = st->src;
  if (!s) return 1;
  for (int i = 0; i < n; ++i) {
    if (s->mode == 0) {
      if (s->ii >= s->ni) { s->overrun = 1; ++s->ii; r[i] = a; }
      else r[i] = s->I[s->ii++];
    } else {
      r[i] = a + (int)(next_u(s) * (double)(b - a));
    }
  }
  return VSL_STATUS_OK;
}

int vsRngGaussianMV(int, VSLStreamStatePtr stream, int n, float *r, int dimen, int mstorage,
                    const float *a, const float *t) {
  Stream *st = ready(stream);
  shim_vsl::Source *s = st->src;
  if (!s) return 1;
  float z[256];
  if (dimen > 256) return 1;
  for (int k = 0; k < n; ++k) {
    for (int i = 0; i < dimen; ++i) z[i] = (float)next_z(s);
    for (int i = 0; i < dimen; ++i) {
      float acc = a[i];
      if (mstorage == VSL_MATRIX_STORAGE_FULL)
        for (int j = 0; j <= i; ++j) acc += t[i*dimen + j] * z[j];
      else
        acc += t[i] * z[i];
      r[k*dimen + i] = acc;
    }
  }
  return VSL_STATUS_OK;
}

/* Column-major Cholesky; only uplo='U' is used (mcpar.cc:475-480): A = U^T U with
 * U stored in the column-major upper triangle == row-major lower triangle L. */
void spotrf(const char *uplo, const int *np, float *a, const int *ldap, int *info) {
  const int n = *np, lda = *ldap;
  *info = 0;
  if (*uplo != 'U' && *uplo != 'u') { *info = -1; return; }
  /* row-major view: L[i][j] (j<=i) lives at a[i*lda + j] == column-major U(j,i) */
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j <= i; ++j) {
      float sum = a[i*lda + j];
      for (int k = 0; k < j; ++k) sum -= a[i*lda + k] * a[j*lda + k];
      if (i == j) {
        if (!(sum > 0)) { *info = i + 1; return; }
        a[i*lda + i] = sqrt(sum);
      } else {
        a[i*lda + j] = sum / a[j*lda + j];
      }
    }
  }
}
}
