// mh_hostlik.cuh -- the MH step split around a HOST likelihood call.
//
// A user-written VLFunc (src/vlfunc.hh:9-12; e.g. the R-backed RFunc, src/rfunc.cc:48-67) cannot be inlined
// into the fused step kernels: its operator() runs on the host.  For such a likelihood the engine does what
// MCPar::run does around the call (src/mcpar.cc:151-175), one step at a time:
//     propose kernel  -> ptrial (+ Hastings factor) on the device -> copied to the host
//     host:  L(nchain, ptrial, lytrial)                                       (the plugin call, mcpar.cc:160)
//     accept kernel   <- lytrial copied to the device: accept test, state update, running moments,
//                        sample store, pool publication
// Draws are the normal mode's counter-based Philox words (same addressing as mh_kernels.cuh), arithmetic is
// plain fp64 operation by operation (this unit is compiled with -fmad=false and CUDA libm), so a run with a
// host likelihood equals the fused run with the same likelihood on the device up to the likelihood's own
// rounding.  Runtime d <= 64, one thread per chain, state chain-major (AoS) in global memory: the host call
// and the two PCIe copies per step bound the speed of this path, not these kernels.
#pragma once

namespace mcgpu {
namespace MCGPU_NS {

static __device__ __forceinline__ uint32_t hl_word(const HostLikParams &p, unsigned long long g, uint32_t base, int idx)
{
  const Words w = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), p.step, base + (uint32_t)(idx / 4), p.key0, p.key1);
  return word_of(w, idx % 4);
}

// un-normalised diagonal Gaussian exponent of pool slot s at y (mcpar.cc:367-387), and the normalisation n_s
static __device__ double hl_exponent(const HostLikParams &p, int s, const double *y, bool normalised)
{
  const double *ms = p.pool + (size_t)s * p.d * 2;
  double arg = 0.0, n = 0.0;
  for (int i = 0; i < p.d; ++i) { const double xm = ms[2 * i] - y[i]; arg += xm * xm / ms[2 * i + 1]; if (normalised) n += log(ms[2 * i + 1]); }
  return -0.5 * n - 0.5 * arg;
}
static __device__ double hl_pool_lse(const HostLikParams &p, const double *y)
{
  double m = -INFINITY;
  for (int s = 0; s < p.pool_m; ++s) { const double a = hl_exponent(p, s, y, true); if (a > m) m = a; }
  if (!(m > -INFINITY)) return m;
  double sum = 0.0;
  for (int s = 0; s < p.pool_m; ++s) sum += exp(hl_exponent(p, s, y, true) - m);
  return m + log(sum);
}

// genLocal (mcpar.cc:302-312) / genRemote (:315-451, or the sum-mixture proposal of remote mode 1) for one step
static __global__ void hostlik_propose_kernel(const HostLikParams p)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p.C) return;
  const int d = p.d, NP = (d + 1) / 2;
  const unsigned long long g = (unsigned long long)(p.chain0 + j);
  const double *xg = p.x + (size_t)j * d;
  double *xt = p.ptrial + (size_t)j * d;
  double z[MCGPU_MAX_D + 1];
  bool remote = false;
  if (p.main_phase && p.t >= p.first_remote_t) {       // one coin per group (or per job), mcpar.cc:142-159
    const unsigned long long leader = p.coin_group > 0 ? g / (unsigned long long)p.coin_group * (unsigned long long)p.coin_group : 0ull;
    remote = !(u32_half(hl_word(p, leader, 0u, 2 * NP + 1)) <= p.pl);
  }
  double aux = remote && p.remote_mode == 1 ? 0.0 : 1.0;
  int c = 0;
  if (!remote) {
    for (int q = 0; q < NP; ++q) normal_pair(hl_word(p, g, 0u, 2 * q), hl_word(p, g, 0u, 2 * q + 1), z[2 * q], z[2 * q + 1]);
    for (int i = 0; i < d; ++i) { double acc = xg[i]; for (int q = 0; q <= i; ++q) acc += p.factor[i * d + q] * z[q]; xt[i] = acc; }
  } else if (p.remote_mode == 1) {                      // x' ~ uniform mixture of the pool, log cfac = log q(x) - log q(x')
    c = (int)__umulhi(hl_word(p, g, MCGPU_SLOT_REMOTE, 0), (uint32_t)p.pool_m);
    for (int q = 0; q < NP; ++q) normal_pair(hl_word(p, g, MCGPU_SLOT_REMOTE, 2 + 2 * q), hl_word(p, g, MCGPU_SLOT_REMOTE, 3 + 2 * q), z[2 * q], z[2 * q + 1]);
    for (int i = 0; i < d; ++i) xt[i] = p.pool[((size_t)c * d + i) * 2] + sqrt(p.pool[((size_t)c * d + i) * 2 + 1]) * z[i];
    aux = hl_pool_lse(p, xg) - hl_pool_lse(p, xt);
    atomicAdd(p.mcounts + 3, 1ull);
  } else {                                              // the reference's rejection loop, one chain at a time
    double qmax = MCGPU_FPEPS;
    for (uint32_t it = 0;; ++it) {
      atomicAdd(p.mcounts + 3, 1ull);
      const uint32_t base = MCGPU_SLOT_REMOTE | (it << 6);
      c = (int)__umulhi(hl_word(p, g, base, 0), (uint32_t)p.pool_m);
      const double u = u32_mid(hl_word(p, g, base, 1));
      for (int q = 0; q < NP; ++q) normal_pair(hl_word(p, g, base, 2 + 2 * q), hl_word(p, g, base, 3 + 2 * q), z[2 * q], z[2 * q + 1]);
      for (int i = 0; i < d; ++i) xt[i] = p.pool[((size_t)c * d + i) * 2] + sqrt(p.pool[((size_t)c * d + i) * 2 + 1]) * z[i];
      qmax = MCGPU_FPEPS; double qsum = MCGPU_FPEPS;
      for (int s = 0; s < p.pool_m; ++s) { const double gv = exp(hl_exponent(p, s, xt, false)); qsum += gv; qmax = gv > qmax ? gv : qmax; }
      if (u < qmax / qsum || it >= (1u << 24) - 1u) break;
    }
    double qold = 0.0;
    for (int s = 0; s < p.pool_m; ++s) { const double gv = exp(hl_exponent(p, s, xg, false)); qold = gv > qold ? gv : qold; }
    aux = qold / qmax;
  }
  p.aux[j] = aux;
  p.flags[j] = (remote ? 1 : 0) | (c << 8);
  if (remote) atomicAdd(p.mcounts + 2, 1ull);
}

// accept test, state update (mcpar.cc:65-75 / :165-175), MCout::add (mcout.cc:129-145), running moments with
// remote adoption (:186-209) and the publication of the pool slot at the end of a window (:205-208)
static __global__ void hostlik_accept_kernel(const HostLikParams p, int publish, double pub_winv)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p.C) return;
  const int d = p.d, NP = (d + 1) / 2;
  const unsigned long long g = (unsigned long long)(p.chain0 + j);
  double *xg = p.x + (size_t)j * d;
  const double *xt = p.ptrial + (size_t)j * d;
  const int fl = p.flags[j];
  const bool remote = fl & 1; const int c = fl >> 8;
  const double u = u32_mid(hl_word(p, g, 0u, 2 * NP));
  const double lyt = p.lytrial[j], delta = lyt - p.ly[j];
  const double pac = (remote && p.remote_mode == 1) ? exp(delta + p.aux[j]) : (p.main_phase ? exp(delta) * p.aux[j] : exp(delta));
  const bool a = u < pac;
  if (a) { p.ly[j] = lyt; for (int i = 0; i < d; ++i) xg[i] = xt[i]; }
  atomicAdd(p.counts, a ? 1ull : 0ull); atomicAdd(p.counts + 1, 1ull);
  if (!p.main_phase) return;
  atomicAdd(p.mcounts, a ? 1ull : 0ull); atomicAdd(p.mcounts + 1, 1ull);
  if (p.hist && p.hist_row >= 0) {
    double *row = p.hist + ((size_t)p.hist_row * p.C + j) * (d + 1);
    for (int i = 0; i < d; ++i) row[i] = xg[i];
    row[d] = p.ly[j];
  }
  const double pwgt = (double)(p.t + 1), winv = 1.0 / pwgt;
  for (int i = 0; i < d; ++i) {
    double mu = p.mu[(size_t)j * d + i], ps = p.ps[(size_t)j * d + i];
    if (remote && a) {                                // sigma -> sigma^2 round trip of :346, :447-448
      const double sd = sqrt(p.pool[((size_t)c * d + i) * 2 + 1]);
      mu = p.pool[((size_t)c * d + i) * 2]; ps = (sd * sd) * (pwgt - 1.0);
    }
    const double dl = xg[i] - mu;
    mu += dl * winv; ps += dl * (xg[i] - mu);
    p.mu[(size_t)j * d + i] = mu; p.ps[(size_t)j * d + i] = ps;
    if (publish && (long long)g % p.pool_stride == 0 && (long long)g / p.pool_stride < p.pool_m) {
      const long long s = (long long)g / p.pool_stride;
      p.pool_next[(s * d + i) * 2] = mu; p.pool_next[(s * d + i) * 2 + 1] = ps * pub_winv;
    }
  }
}

}  // namespace MCGPU_NS
}  // namespace mcgpu
