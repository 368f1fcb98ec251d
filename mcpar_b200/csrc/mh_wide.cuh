// mh_wide.cuh -- fused MH step kernel for d >= 8: D/2 lanes per chain, two parameters per lane.
//
// At d = 16..64 a thread-per-chain kernel needs 150-255 registers (x, x', mu, psum2 are 4d
// doubles) and runs at 2 warps per scheduler.  Here a chain is spread over L = D/2 lanes
// (d = 64: one chain per warp; d = 16: four chains per warp); each lane owns two consecutive
// parameters, draws its own Box-Muller pair (one Philox call), and the chain's lanes meet
// only where the algorithm couples parameters:
//   * x' = x + T z for a non-diagonal factor (z staged in shared memory, T column-major),
//   * the likelihood (lane-local pair terms + a shuffle reduction for Rosenbrock1; one
//     mixture component per lane over the staged x' for GaussMix),
//   * the remote proposal's pool test (one pool component per lane over the staged x').
// Same reference lines as mh_steps_kernel: genLocal mcpar.cc:302-312, genRemote :315-451,
// accept :165-175, moments :186-209, MCout::add mcout.cc:129-145.  Job-wide coin only
// (phases PH_BURN / PH_LOCAL / PH_REMOTE are launch-uniform).  Production (Philox) unit only.
#pragma once

namespace mcgpu {
namespace MCGPU_NS {


template <int L>
__device__ __forceinline__ double group_sum(double v)
{
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int L>
__device__ __forceinline__ double group_max(double v)
{
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) { const double w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
  return v;
}

// log-likelihood of the chain whose point is staged in sx[0..D) (and held as (x0,x1) per lane)
template <int LIK, int D>
__device__ __forceinline__ double wide_loglik(double x0, double x1, const double *sx, int r, const WideParams &p,
                                              const MathTables &T)
{
  constexpr int L = D / 2;
  if (LIK == MCGPU_ROSENBROCK1) {               // rosenbrock.cc:4-21: the pair (2r, 2r+1) lives in lane r
    const double t1 = 1 - x0;
    const double t2 = x1 - x0 * x0;
    return -group_sum<L>(t1 * t1 + 100.0 * t2 * t2);
  } else {                                      // GaussMix: lane r evaluates components r, r+L, ...
    double m = -INFINITY, s = 0.0;
    for (int k = r; k < p.kpad; k += L) {
      double q = 0.0;
#pragma unroll 8
      for (int i = 0; i < D; ++i) {
        const double xm = sx[i] - __ldg(p.gm_mu + (size_t)i * p.kpad + k);
        q += xm * xm * __ldg(p.gm_is2 + (size_t)i * p.kpad + k);
      }
      const double a = __ldg(p.gm_lw + k) - 0.5 * q;      // padding components carry log w = -inf
      if (a > m) { s = s * mc_exp(m - a, T) + 1.0; m = a; } else if (a > -INFINITY) s += mc_exp(a - m, T);
    }
    const double gm = group_max<L>(m);
    s = group_sum<L>(s * mc_exp(m - gm, T));
    return gm + mc_log(s, T);
  }
}

// max_s and (fp32-bounded) sum_s of Q_s over the pool for the point staged in sx; exact fp64 sum
// on demand.  Lane r evaluates pool slots r, r+L, ...; result identical in every lane of the chain.
template <int D>
__device__ __forceinline__ void wide_pool_eval(const double *sx, int r, const WideParams &p, double &amax, float &S)
{
  constexpr int L = D / 2;
  constexpr float L2E = 1.4426950408889634f;
  double m = -INFINITY; float sl = 0.0f;
  for (int s = r; s < p.mpad; s += L) {
    double a = 0.0;
#pragma unroll 8
    for (int i = 0; i < D; ++i) {
      const double xm = __ldg(p.pm + (size_t)i * p.mpad + s) - sx[i];
      a += xm * xm * __ldg(p.ph + (size_t)i * p.mpad + s);
    }
    const bool gt = a > m;
    const float e = ex2_approx(-fabsf((float)(a - m)) * L2E);
    sl = gt ? fmaf(sl, e, 1.0f) : sl + e;
    m = gt ? a : m;
  }
  amax = group_max<L>(m);
  const float scaled = sl * ex2_approx((float)(m - amax) * L2E);    // m = -inf (no slot): 0
  float tot = scaled;
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  S = tot;
}

template <int D>
__device__ __noinline__ double wide_pool_exact_sum(const double *sx, int r, const WideParams &p, double &qmax, const MathTables &T)
{
  constexpr int L = D / 2;
  double qs = 0.0, qm = 0.0;
  for (int s = r; s < p.pool_m; s += L) {
    double a = 0.0;
    for (int i = 0; i < D; ++i) {
      const double xm = __ldg(p.pm + (size_t)i * p.mpad + s) - sx[i];
      a += xm * xm * __ldg(p.ph + (size_t)i * p.mpad + s);
    }
    const double gv = mc_exp(a, T);
    qs += gv; qm = gv > qm ? gv : qm;
  }
  qmax = group_max<L>(qm);
  return group_sum<L>(qs);
}

template <int LIK, int D, int PHASE>
__global__ void __launch_bounds__(128, 4)
mh_wide_kernel(const WideParams p)
{
  constexpr int L = D / 2;                      // lanes per chain
  constexpr int CPW = 32 / L;                   // chains per warp
  constexpr bool MAIN = PHASE != PH_BURN;
  extern __shared__ double smem[];
  // smem: math tables | per chain of the CTA: staged point sx[D] and staged normals sz[D]
  MathTables T;
  T.exp_tab = smem; T.log_tab = smem + MCGPU_EXP_TAB; T.trig_tab = T.log_tab + 2 * MCGPU_LOG_TAB;
  for (int i = threadIdx.x; i < MCGPU_EXP_TAB; i += blockDim.x) smem[i] = MCGPU_EXP_TABLE[i];
  for (int i = threadIdx.x; i < 2 * MCGPU_LOG_TAB; i += blockDim.x) smem[MCGPU_EXP_TAB + i] = MCGPU_LOG_TABLE[i];
  for (int i = threadIdx.x; i < 2 * MCGPU_TRIG_TAB; i += blockDim.x) smem[MCGPU_EXP_TAB + 2 * MCGPU_LOG_TAB + i] = MCGPU_TRIG_TABLE[i];
  const int cib = threadIdx.x / L;              // chain within the CTA
  double *sx = smem + MCGPU_MATH_SMEM + (size_t)cib * 2 * D;
  double *sz = sx + D;
  __syncthreads();

  const int r = threadIdx.x % L;                // lane within the chain: owns parameters 2r, 2r+1
  const long long j = (long long)blockIdx.x * (blockDim.x / L) + cib;
  const bool live = j < p.C;
  const long long jc = live ? j : p.C - 1;      // idle groups shadow the last chain, never store
  const unsigned long long g = (unsigned long long)(p.chain0 + jc);
  const uint32_t glo = (uint32_t)g, ghi = (uint32_t)(g >> 32);
  const int i0 = 2 * r;

  double x0 = p.x[jc * D + i0], x1 = p.x[jc * D + i0 + 1];
  double ly = p.ly[jc];
  double mu0 = 0, mu1 = 0, ps0 = 0, ps1 = 0;
  if (MAIN) { mu0 = p.mu[jc * D + i0]; mu1 = p.mu[jc * D + i0 + 1]; ps0 = p.ps[jc * D + i0]; ps1 = p.ps[jc * D + i0 + 1]; }
  const double tdiag0 = p.factor_rm[i0 * D + i0], tdiag1 = p.factor_rm[(i0 + 1) * D + i0 + 1];
  const bool diag = *p.diagonal != 0;
  unsigned int nacc = 0;
  int tmod = MAIN ? p.t0 % p.thin : 0;
  long long tkeep = MAIN ? (long long)(p.t0 / p.thin) - p.hist_step0 : 0;

  for (int k = 0; k < p.nsteps; ++k) {
    const uint32_t step = p.step0 + (uint32_t)k;
    const int t = p.t0 + k;
    constexpr int ABLK = (2 * L) / 4, AW = (2 * L) % 4;     // accept uniform: word 2*NP (NP = L pairs) of the local stream
    const Words wacc = philox4x32_10(glo, ghi, step, (uint32_t)ABLK, p.key0, p.key1);
    const double u_acc = u32_mid(word_of(wacc, AW));
    double xt0, xt1, cfac = 1.0;
    int cpick = 0;

    if (PHASE != PH_REMOTE) {
      // genLocal: lane r's Box-Muller pair = words (2r, 2r+1) of the local stream
      double za, zb;
      {
        const Words b = philox4x32_10(glo, ghi, step, (uint32_t)(r >> 1), p.key0, p.key1);
        normal_pair_t((r & 1) ? b.w2 : b.w0, (r & 1) ? b.w3 : b.w1, za, zb, T);
      }
      if (diag) {
        xt0 = x0 + tdiag0 * za; xt1 = x1 + tdiag1 * zb;
      } else {
        sz[i0] = za; sz[i0 + 1] = zb;
        __syncwarp();
        double a0 = x0, a1 = x1;                 // rows 2r and 2r+1, terms added in q = 0,1,.. order
        for (int q = 0; q <= i0; ++q) { const double zq = sz[q]; a0 += __ldg(p.factor_cm + q * D + i0) * zq; a1 += __ldg(p.factor_cm + q * D + i0 + 1) * zq; }
        a1 += __ldg(p.factor_cm + (i0 + 1) * D + i0 + 1) * zb;
        xt0 = a0; xt1 = a1;
        __syncwarp();
      }
    } else {
      // genRemote: the chain's lanes run the reference's rejection loop together; lane r
      // evaluates pool slots r, r+L, ...
      // The loop is warp-uniform: when a warp holds several chains (d < 64) the groups that
      // have accepted keep evaluating (discarded) candidates until the warp's last group is
      // done, so every warp-wide primitive below is reached by all 32 lanes.
      double amax = 0.0;
      bool done = false;
      xt0 = x0; xt1 = x1;
      for (uint32_t it = 0;; ++it) {
        const uint32_t slot = MCGPU_SLOT_REMOTE | (it << 6);
        const Words w0 = philox4x32_10(glo, ghi, step, slot, p.key0, p.key1);      // same block in every lane of the chain
        const int c = (int)__umulhi(w0.w0, (uint32_t)p.pool_m);                     // viRngUniform, mcpar.cc:337
        const double u = u32_mid(w0.w1);                                            // vsRngUniform, mcpar.cc:401
        double za, zb;
        {                                                // lane r's pair = words (2+2r, 3+2r) of the candidate's stream
          const int qq = r + 1;
          const Words b = philox4x32_10(glo, ghi, step, slot + (uint32_t)(qq >> 1), p.key0, p.key1);
          normal_pair_t((qq & 1) ? b.w2 : b.w0, (qq & 1) ? b.w3 : b.w1, za, zb, T);
        }
        const double c0 = __ldg(p.pm + (size_t)i0 * p.mpad + c) + __ldg(p.psd + (size_t)i0 * p.mpad + c) * za;
        const double c1 = __ldg(p.pm + (size_t)(i0 + 1) * p.mpad + c) + __ldg(p.psd + (size_t)(i0 + 1) * p.mpad + c) * zb;
        sx[i0] = c0; sx[i0 + 1] = c1;
        __syncwarp();
        double am; float S;
        wide_pool_eval<D>(sx, r, p, am, S);
        bool acc = false, decided = false;
        if (am > -10.0) {                                // then FPEPS/qmax < 2.3e-10 (mcpar.cc:357-358 offsets)
          const double eps = 1.0e-4 + 2.0e-5 * (double)p.pool_m, Sd = (double)S;
          if (u * (Sd * (1.0 + eps) + 3.0e-10) < 1.0) { acc = true; decided = true; }
          else if (u * (Sd * (1.0 - eps)) >= 1.0) { acc = false; decided = true; }
        }
        if (!__all_sync(0xffffffffu, decided || done)) {  // rare: exact pacpt = qimax / qisum for the whole warp
          double qmax;
          const double qsum = wide_pool_exact_sum<D>(sx, r, p, qmax, T) + MCGPU_FPEPS;
          qmax = qmax > MCGPU_FPEPS ? qmax : MCGPU_FPEPS;
          if (!decided) acc = u < qmax / qsum;
        }
        __syncwarp();
        if (!done && (acc || it >= (1u << 24) - 2u)) { xt0 = c0; xt1 = c1; cpick = c; amax = am; done = true; }
        if (__all_sync(0xffffffffu, done)) break;
      }
      // cfac = max_i Q_i(x_old) / max_i Q_i(x'), mcpar.cc:412-439
      sx[i0] = x0; sx[i0 + 1] = x1;
      __syncwarp();
      double aold; float dummy;
      wide_pool_eval<D>(sx, r, p, aold, dummy);
      __syncwarp();
      double qmax = mc_exp(amax, T);
      qmax = qmax > MCGPU_FPEPS ? qmax : MCGPU_FPEPS;
      cfac = mc_exp(aold, T) / qmax;
    }

    if (LIK != MCGPU_ROSENBROCK1) { sx[i0] = xt0; sx[i0 + 1] = xt1; __syncwarp(); }
    const double lyt = wide_loglik<LIK, D>(xt0, xt1, sx, r, p, T);
    if (LIK != MCGPU_ROSENBROCK1) __syncwarp();
    const bool a = accept_test(u_acc, lyt - ly, MAIN ? cfac : 1.0, T);   // same inputs in every lane of the chain
    if (a) { ly = lyt; x0 = xt0; x1 = xt1; }
    nacc += a ? 1u : 0u;

    if (MAIN) {
      if (p.hist && live && tmod == 0) {               // MCout::add: one row per chain, coalesced over the lanes
        double *row = p.hist + (tkeep * p.C + j) * (D + 1);
        row[i0] = x0; row[i0 + 1] = x1;
        if (r == 0) row[D] = ly;
      }
      if (++tmod == p.thin) { tmod = 0; ++tkeep; }
      const double pwgt = (double)(t + 1), winv = 1.0 / pwgt;
      if (PHASE == PH_REMOTE && a) {                   // adopt the component's moments, mcpar.cc:190-197
        const double sd0 = __ldg(p.psd + (size_t)i0 * p.mpad + cpick), sd1 = __ldg(p.psd + (size_t)(i0 + 1) * p.mpad + cpick);
        mu0 = __ldg(p.pm + (size_t)i0 * p.mpad + cpick); mu1 = __ldg(p.pm + (size_t)(i0 + 1) * p.mpad + cpick);
        ps0 = (sd0 * sd0) * (pwgt - 1.0); ps1 = (sd1 * sd1) * (pwgt - 1.0);
      }
      double dl = x0 - mu0; mu0 += dl * winv; ps0 += dl * (x0 - mu0);
      dl = x1 - mu1; mu1 += dl * winv; ps1 += dl * (x1 - mu1);
    }
  }

  if (live) {
    p.x[j * D + i0] = x0; p.x[j * D + i0 + 1] = x1;
    if (r == 0) p.ly[j] = ly;
    if (MAIN) {
      p.mu[j * D + i0] = mu0; p.mu[j * D + i0 + 1] = mu1; p.ps[j * D + i0] = ps0; p.ps[j * D + i0 + 1] = ps1;
      const long long gg = p.chain0 + j;
      if (p.pool_next && gg % p.pool_stride == 0 && gg / p.pool_stride < p.pool_m) {
        const long long s = gg / p.pool_stride;
        const double winv = 1.0 / (double)(p.t0 + p.nsteps);
        p.pool_next[(s * D + i0) * 2] = mu0;     p.pool_next[(s * D + i0) * 2 + 1] = ps0 * winv;
        p.pool_next[(s * D + i0 + 1) * 2] = mu1; p.pool_next[(s * D + i0 + 1) * 2 + 1] = ps1 * winv;
      }
    }
  }
  // acceptance counters: one count per chain (lane 0 of each group)
  unsigned int wacc = (live && r == 0) ? nacc : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wacc += __shfl_xor_sync(0xffffffffu, wacc, o);
  const unsigned int nlive = __popc(__ballot_sync(0xffffffffu, live && r == 0));
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(p.counts, (unsigned long long)wacc);
    atomicAdd(p.counts + 1, (unsigned long long)nlive * (unsigned long long)p.nsteps);
  }
}

// pool [M][D][2] (mu, sigma^2) -> pm / ph / psd [D][Mpad]; padding slots get Q = 0
static __global__ void pool_prep_kernel(const double *pool, int M, int mpad, int D, double *pm, double *ph, double *psd)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= D * mpad) return;
  const int i = idx / mpad, s = idx % mpad;
  if (s < M) {
    const double s2 = pool[((size_t)s * D + i) * 2 + 1];
    pm[idx] = pool[((size_t)s * D + i) * 2]; ph[idx] = -0.5 / s2; psd[idx] = sqrt(s2);
  } else { pm[idx] = 1.0e300; ph[idx] = -1.0; psd[idx] = 0.0; }
}

// row-major lower factor -> column-major copy + "is diagonal" flag
static __global__ void factor_prep_kernel(const double *rm, double *cm, int D, int *diag)
{
  __shared__ int offd;
  if (threadIdx.x == 0) offd = 0;
  __syncthreads();
  for (int idx = threadIdx.x; idx < D * D; idx += blockDim.x) {
    const int i = idx / D, q = idx % D;
    const double v = q <= i ? rm[idx] : 0.0;          // strict upper triangle holds stale input (covar_setup)
    cm[q * D + i] = v;
    if (q < i && v != 0.0) offd = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) *diag = !offd;
}

}  // namespace MCGPU_NS
}  // namespace mcgpu
