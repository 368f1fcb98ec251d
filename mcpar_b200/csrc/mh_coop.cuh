// mh_coop.cuh -- CTA-cooperative fused MH step kernel for the Gaussian-mixture likelihood at d = 64,
// K <= 64 components (BASELINE config 4: SURVEY.md 8d C4).
//
// The wide kernel (mh_wide.cuh) streams the mixture's 64 KB of (a, b) pairs through L1 for every group of
// four chains and the remote pool (131 KB in fp32 at M = 256) through L1/L2: at 12 warps per SM it is bound
// by the latency of those loads (FP64 pipe 19 % busy, a sum-mixture remote step at M = 256 ten times a
// local step).  Here the data that every chain needs stays put and the chains stream past it:
//   * MIXTURE IN REGISTERS.  A CTA of 256 threads holds the whole mixture: thread t owns component
//     k = t mod 64 and the quarter h = t / 64 of the parameters (16 (a, b) pairs = 64 registers).  For a
//     batch of NB chains (NCW per warp) each thread accumulates its partial exponent
//     sum_{i in quarter} x_i (a_ki x_i + b_ki) for every chain of the batch -- the points come from shared
//     memory as warp-wide broadcasts, 8 LDS.128 per 32 DFMA -- and leaves it in shared memory.
//   * POOL IN SHARED MEMORY (remote mode 1).  The fp32 copy of the pool (g mu, g) is staged once per CTA;
//     thread t owns pool slot t mod SL and a slice of the parameters and accumulates the 2 NB exponents
//     (x' and x of every chain of the batch) in registers: 1 + NB/2 LDS per 4 NB FFMA.
//   * ONE WARP OWNS A CHAIN for everything else (lane r = parameters 2r, 2r+1, exactly the wide kernel's
//     layout): Philox + Box-Muller proposal, the log-sum-exp over the 64 exponents, the accept test, the
//     running moments, the history row and the publication to the exchange pool.
//   * PERSISTENT CTAs: the grid is the number of SMs (x 2 for the kernels without the pool), each CTA walks
//     batches blockIdx.x, blockIdx.x + gridDim.x, ...; registers and shared memory are loaded once per launch.
// Phases of a step:  P1 owner warps propose and stage x'  | barrier |  P2 all threads: mixture partials
// (+ pool partials)  | barrier |  (pool: reduce the partials | barrier |)  P3 owner warps: likelihood, accept, update.
//
// Same reference lines as the other step kernels: genLocal mcpar.cc:302-312, genRemote :315-451 (remote
// mode 1: the sum-mixture proposal of mh_kernels.cuh), accept :165-175, moments :186-209, MCout::add
// mcout.cc:129-145.  Draw addressing, arithmetic of the proposal, decisions by bounds and statistics are
// those of mh_wide.cuh; only the order of the floating-point sums of the likelihood and of the pool test
// differs (quarters, chunks).
#pragma once

namespace mcgpu {
namespace MCGPU_NS {

constexpr int kCoopThreads = 256, kCoopWarps = 8, kCoopKP = 64;

// dynamic shared memory of one CTA; SL = pool slots held side by side (power of two, 16..256; 0: no pool)
template <int D, int NCW>
__host__ __device__ constexpr size_t coop_smem_bytes(int SL)
{
  constexpr int NB = kCoopWarps * NCW, CP = 2 * NB;
  size_t b = sizeof(double) * ((size_t)MCGPU_MATH_SMEM + (size_t)NB * D + (size_t)kCoopWarps * D * NCW + (size_t)NB * 4 * kCoopKP);
  if (SL > 0) b += sizeof(float) * ((size_t)D * CP + (size_t)kCoopThreads * CP + (size_t)SL) + sizeof(float2) * (size_t)D * SL;
  return b;
}

template <int D, int NCW, int PHASE>
__global__ void __launch_bounds__(kCoopThreads, (PHASE == PH_REMOTE_SUM ? 1 : 2))
mh_coop_kernel(const WideParams p, const int nbatch, const int SL, const int lsl)
{
  constexpr int L = D / 2, NW = kCoopWarps, NB = NW * NCW, KP = kCoopKP, DQ = D / 4, CP = 2 * NB;
  static_assert(L == 32, "one warp owns a chain: d = 64");
  constexpr bool MAIN = PHASE != PH_BURN, REMOTE = PHASE == PH_REMOTE_SUM;
  constexpr int ABLK = (2 * L) / 4, AW = (2 * L) % 4;   // accept uniform: word 2*NP of the local stream
  extern __shared__ __align__(16) double smem[];
  MathTables T;
  T.exp_tab = smem; T.log_tab = smem + MCGPU_EXP_TAB; T.trig_tab = T.log_tab + 2 * MCGPU_LOG_TAB;
  stage_math_tables(smem);
  double *sx = smem + MCGPU_MATH_SMEM;                  // [NB][D]      proposals of the batch
  double *szw = sx + NB * D;                            // [NW][D*NCW]  per-warp scratch (normals of a dense factor; exact-path points)
  double *sq = szw + NW * D * NCW;                      // [NB][4][KP]  partial exponents
  float *sxf = reinterpret_cast<float *>(sq + NB * 4 * KP);   // [D][CP]   fp32 points: column 2c = x' of chain c, 2c+1 = x
  float *part = sxf + D * CP;                           // [DC][CP][SL] partial pool exponents, DC * SL = 256; [0] ends up holding the totals
  float *snb = part + kCoopThreads * CP;                // [SL]         n_s log2 e
  float2 *spool = reinterpret_cast<float2 *>(snb + SL); // [D][SL]      (g mu, g)
  if (p.npeers > 0 && *reinterpret_cast<volatile int *>(p.xflag)) return;   // a peer-to-peer wait timed out earlier: stop stepping

  const int tid = threadIdx.x, w = tid >> 5, r = tid & 31, i0 = 2 * r;
  const int kc = tid & (KP - 1), h = tid >> 6;          // this thread's mixture component and parameter quarter
  double A[DQ], B[DQ];                                  // exponent of component k: c_k + sum_i x_i (a_ki x_i + b_ki)
#pragma unroll
  for (int j = 0; j < DQ; ++j) { const double2 ab = __ldg(p.gm2 + (size_t)(h * DQ + j) * p.kpad + kc); A[j] = ab.x; B[j] = ab.y; }
  const double lw0 = __ldg(p.gm_lw + r), lw1 = __ldg(p.gm_lw + r + 32);   // owner-warp lane r sums components r and r + 32
  if (REMOTE) {
    for (int idx = tid; idx < D * SL; idx += kCoopThreads) {
      const int i = idx >> lsl, s = idx & (SL - 1);
      spool[idx] = s < p.mpad ? __ldg(p.pf + (size_t)i * p.mpad + s) : make_float2(1.0e18f, 0.0f);   // padding: Q = 0
    }
    for (int s = tid; s < SL; s += kCoopThreads) snb[s] = s < p.mpad ? __ldg(p.pnbf + s) : 0.0f;
  }
  const double tdiag0 = p.factor_rm[i0 * D + i0], tdiag1 = p.factor_rm[(i0 + 1) * D + i0 + 1];
  const bool diag = *p.diagonal != 0;
  double *sz = szw + (size_t)w * D * NCW;
  unsigned int wacc = 0, nlive = 0, nfb = 0;            // per lane 0: accepted steps, live chains walked, exact-path fallbacks
  __syncthreads();

  for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
    const long long jb = ((long long)batch * NW + w) * NCW;
    long long jc[NCW]; bool live[NCW]; uint32_t glo[NCW], ghi[NCW];
    double x0[NCW], x1[NCW], ly[NCW], mu0[NCW], mu1[NCW], ps0[NCW], ps1[NCW];
    unsigned int nacc[NCW];
#pragma unroll
    for (int c = 0; c < NCW; ++c) {
      live[c] = jb + c < p.C;
      jc[c] = live[c] ? jb + c : p.C - 1;              // idle slots shadow the last chain, never store
      const unsigned long long g = (unsigned long long)(p.chain0 + jc[c]);
      glo[c] = (uint32_t)g; ghi[c] = (uint32_t)(g >> 32);
      const double2 xv = *reinterpret_cast<const double2 *>(p.x + jc[c] * D + i0);
      x0[c] = xv.x; x1[c] = xv.y;
      ly[c] = p.ly[jc[c]];
      mu0[c] = mu1[c] = ps0[c] = ps1[c] = 0.0;
      nacc[c] = 0;
    }
    int tmod = MAIN ? p.t0 % p.thin : 0;
    int tring = p.hist_ring0;                           // ring row of the next kept step

    for (int k = 0; k < p.nsteps; ++k) {
      const uint32_t step = p.step0 + (uint32_t)k;
      const int t = p.t0 + k;
      double u_acc[NCW], xt0[NCW], xt1[NCW];
      int cpick[NCW];
      // ---- P1: the owner warp proposes (genLocal / sum-mixture genRemote) and stages x'
#pragma unroll
      for (int c = 0; c < NCW; ++c) {
        const Words wa = philox4x32_10_rk(glo[c], ghi[c], step, (uint32_t)ABLK, p.rk);
        u_acc[c] = u32_mid(word_of(wa, AW));
        cpick[c] = 0;
      }
      if (!REMOTE) {
        double za[NCW], zb[NCW];                        // lane r's Box-Muller pair = words (2r, 2r+1) of the chain's local stream
#pragma unroll
        for (int c = 0; c < NCW; ++c) {
          const Words b = philox4x32_10_rk(glo[c], ghi[c], step, (uint32_t)(r >> 1), p.rk);
          normal_pair_t((r & 1) ? b.w2 : b.w0, (r & 1) ? b.w3 : b.w1, za[c], zb[c], T);
        }
        if (diag) {
#pragma unroll
          for (int c = 0; c < NCW; ++c) { xt0[c] = x0[c] + tdiag0 * za[c]; xt1[c] = x1[c] + tdiag1 * zb[c]; }
        } else {                                        // x' = x + T z, rows 2r and 2r+1, terms added in q = 0,1,.. order
#pragma unroll
          for (int c = 0; c < NCW; ++c) { sz[i0 * NCW + c] = za[c]; sz[(i0 + 1) * NCW + c] = zb[c]; }
          __syncwarp();
          double a0[NCW], a1[NCW];
#pragma unroll
          for (int c = 0; c < NCW; ++c) { a0[c] = x0[c]; a1[c] = x1[c]; }
          for (int q = 0; q <= i0; ++q) {
            const double t0 = __ldg(p.factor_cm + q * D + i0), t1 = __ldg(p.factor_cm + q * D + i0 + 1);
#pragma unroll
            for (int c = 0; c < NCW; ++c) { const double zq = sz[q * NCW + c]; a0[c] += t0 * zq; a1[c] += t1 * zq; }
          }
          const double tl = __ldg(p.factor_cm + (i0 + 1) * D + i0 + 1);
#pragma unroll
          for (int c = 0; c < NCW; ++c) { xt0[c] = a0[c]; xt1[c] = a1[c] + tl * zb[c]; }
          __syncwarp();
        }
      } else {
        // remote mode 1: candidate 0 of the remote stream is THE proposal (one pick, one set of normals)
#pragma unroll
        for (int c = 0; c < NCW; ++c) {
          const Words w0 = philox4x32_10_rk(glo[c], ghi[c], step, MCGPU_SLOT_REMOTE, p.rk);
          cpick[c] = (int)__umulhi(w0.w0, (uint32_t)p.pool_m);
          const int qq = r + 1;                         // lane r's pair = words (2+2r, 3+2r)
          Words b = w0;
          if ((qq >> 1) != 0) b = philox4x32_10_rk(glo[c], ghi[c], step, MCGPU_SLOT_REMOTE + (uint32_t)(qq >> 1), p.rk);
          double za, zb;
          normal_pair_t((qq & 1) ? b.w2 : b.w0, (qq & 1) ? b.w3 : b.w1, za, zb, T);
          xt0[c] = __ldg(p.pmh + (size_t)i0 * p.mpad + cpick[c]).x + __ldg(p.psd + (size_t)i0 * p.mpad + cpick[c]) * za;
          xt1[c] = __ldg(p.pmh + (size_t)(i0 + 1) * p.mpad + cpick[c]).x + __ldg(p.psd + (size_t)(i0 + 1) * p.mpad + cpick[c]) * zb;
          const int cb = w * NCW + c;
          *reinterpret_cast<float2 *>(sxf + i0 * CP + 2 * cb) = make_float2((float)xt0[c], (float)x0[c]);
          *reinterpret_cast<float2 *>(sxf + (i0 + 1) * CP + 2 * cb) = make_float2((float)xt1[c], (float)x1[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < NCW; ++c) *reinterpret_cast<double2 *>(sx + (w * NCW + c) * D + i0) = make_double2(xt0[c], xt1[c]);
      __syncthreads();

      // ---- P2: every thread adds its quarter of its component's exponent for every chain of the batch
#pragma unroll 1
      for (int c0 = 0; c0 < NB; c0 += 4) {
        double q[4] = {0.0, 0.0, 0.0, 0.0};
        const double *xq = sx + c0 * D + h * DQ;
#pragma unroll
        for (int j = 0; j < DQ; j += 2) {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const double2 xv = *reinterpret_cast<const double2 *>(xq + cc * D + j);     // the same address in every lane: broadcast
            q[cc] = fma(fma(A[j], xv.x, B[j]), xv.x, q[cc]);
            q[cc] = fma(fma(A[j + 1], xv.y, B[j + 1]), xv.y, q[cc]);
          }
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) sq[((c0 + cc) * 4 + h) * KP + kc] = q[cc];
      }
      if (REMOTE) {
        // pool slot s, parameter chunk ch: -sum_{i in chunk} (g mu - g y_i)^2 for the 2 NB points of the batch
        const int s = tid & (SL - 1), ch = tid >> lsl, dpc = (D * SL) >> 8;
        float acc[CP];
#pragma unroll
        for (int v = 0; v < CP; ++v) acc[v] = 0.0f;
        const float2 *pp = spool + (size_t)ch * dpc * SL + s;
        const float4 *xv = reinterpret_cast<const float4 *>(sxf + (size_t)ch * dpc * CP);
#pragma unroll 2
        for (int ii = 0; ii < dpc; ++ii) {
          const float2 f = pp[(size_t)ii * SL];
#pragma unroll
          for (int v = 0; v < CP / 4; ++v) {
            const float4 t4 = xv[ii * (CP / 4) + v];
            float y;
            y = fmaf(-f.y, t4.x, f.x); acc[4 * v] = fmaf(-y, y, acc[4 * v]);
            y = fmaf(-f.y, t4.y, f.x); acc[4 * v + 1] = fmaf(-y, y, acc[4 * v + 1]);
            y = fmaf(-f.y, t4.z, f.x); acc[4 * v + 2] = fmaf(-y, y, acc[4 * v + 2]);
            y = fmaf(-f.y, t4.w, f.x); acc[4 * v + 3] = fmaf(-y, y, acc[4 * v + 3]);
          }
        }
#pragma unroll
        for (int v = 0; v < CP; ++v) part[((size_t)ch * CP + v) * SL + s] = acc[v];
      }
      __syncthreads();
      if (REMOTE) {
        // totals over the chunks: A_s(y) log2 e = nb_s - sum_i (..)^2, left in part[0][point][slot]
        const int s = tid & (SL - 1), ch = tid >> lsl, dc = kCoopThreads >> lsl;
        for (int v = ch; v < CP; v += dc) {
          float e = snb[s];
          for (int c2 = 0; c2 < dc; ++c2) e += part[((size_t)c2 * CP + v) * SL + s];
          part[(size_t)v * SL + s] = e;                 // only this thread reads part[0][v][s]
        }
        __syncthreads();
      }

      // ---- P3: the owner warp finishes the likelihood, decides and updates
      const double pwgt = (double)(t + 1), winv = 1.0 / pwgt;
#pragma unroll
      for (int c = 0; c < NCW; ++c) {
        const int cb = w * NCW + c;
        double a0 = lw0, a1 = lw1;
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) { a0 += sq[(cb * 4 + hh) * KP + r]; a1 += sq[(cb * 4 + hh) * KP + r + 32]; }
        const double gm = group_max<L>(a0 > a1 ? a0 : a1);
        double se = (a0 > -INFINITY ? mc_exp(a0 - gm, T) : 0.0) + (a1 > -INFINITY ? mc_exp(a1 - gm, T) : 0.0);
        se = group_sum<L>(se);
        const double lyt = gm + mc_log(se, T);          // no finite exponent: -inf + log 0 = -inf
        bool a;
        if (REMOTE) {
          // u < exp(lyt - ly) q(x)/q(x') from fp32 bounds on the Hastings factor (summix_bounds of mh_kernels.cuh;
          // here the exponents are summed chunk by chunk and taken relative to the picked component afterwards:
          // D + 256/SL + 2 roundings instead of D, covered by the larger rounding term cu)
          const float *En = part + (size_t)(2 * cb) * SL, *Eo = En + SL;
          const float ref = En[cpick[c]];
          float sn = 0.0f, so = 0.0f;
          for (int s = r; s < SL; s += 32) { sn += ex2_approx(En[s] - ref); so += ex2_approx(Eo[s] - ref); }
          sn = group_sumf<L>(sn); so = group_sumf<L>(so);
          const float xabs_n = group_maxf<L>(fmaxf(fabsf((float)xt0[c]), fabsf((float)xt1[c])));
          const float xabs_o = group_maxf<L>(fmaxf(fabsf((float)x0[c]), fabsf((float)x1[c])));
          const float mumax = __ldg(p.pscal), isig = __ldg(p.pscal + 1), nbmax = __ldg(p.pscal + 2);
          const float l2m = lg2_approx((float)p.pool_m), cD = (float)D, cu = (24.0f + cD) * 6.0e-8f;
          const float cn = 4.8e-7f * nbmax + 1.0e-5f + 1.0e-8f * (float)p.pool_m;
          bool ok = (sn < 1.0e30f) && (so < 1.0e30f) && (sn > 0.5f);
          const float th_n = 1.6e-7f * (mumax + xabs_n) * isig, th_o = 1.6e-7f * (mumax + xabs_o) * isig;
          const float lvl = nbmax - ref + l2m + 30.0f;
          const float Een = fmaxf(lvl - lg2_approx(sn), 30.0f), Eeo = fmaxf(lvl - lg2_approx(fmaxf(so, 1.0e-37f)), 30.0f);
          const float eps_n = 0.75f * (2.0f * th_n * sqrtf(cD * Een) + cu * Een + cD * th_n * th_n) + cn;
          const float eps_o = 0.75f * (2.0f * th_o * sqrtf(cD * Eeo) + cu * Eeo + cD * th_o * th_o) + cn;
          ok = ok && th_n < 1.0e-3f && th_o < 1.0e-3f && eps_n < 0.02f && eps_o < 0.02f;
          const float cf_lo = __fdividef(so * (1.0f - eps_o), sn * (1.0f + eps_n)) * (1.0f - 1.0e-6f);
          const float cf_hi = __fdividef(fmaf(so, 1.0f + eps_o, 2.0e-38f * (float)p.pool_m), sn * (1.0f - eps_n)) * (1.0f + 1.0e-6f);
          int dec = ok ? accept_test_bounded(u_acc[c], lyt - ly[c], cf_lo, cf_hi) : -1;
          if (dec < 0) {                                // rare, warp-uniform: exact log q(x) - log q(x') in fp64
            nfb += live[c] ? 1u : 0u;
            __syncwarp();
            sz[i0 * NCW + c] = xt0[c]; sz[(i0 + 1) * NCW + c] = xt1[c];
            __syncwarp();
            const double ln_ = wide_pool_lse_exact<D, NCW>(sz, c, r, p, T);
            __syncwarp();
            sz[i0 * NCW + c] = x0[c]; sz[(i0 + 1) * NCW + c] = x1[c];
            __syncwarp();
            const double lo_ = wide_pool_lse_exact<D, NCW>(sz, c, r, p, T);
            __syncwarp();
            dec = u_acc[c] < mc_exp((lyt - ly[c]) + (lo_ - ln_), T) ? 1 : 0;
          }
          a = dec != 0;
        } else {
          a = accept_test_local(u_acc[c], lyt - ly[c], T, 0);      // mcpar.cc:67-69 / :167-169 with cfac = 1
        }
        if (a) { ly[c] = lyt; x0[c] = xt0[c]; x1[c] = xt1[c]; }
        nacc[c] += a ? 1u : 0u;
        if (MAIN) {
          if (p.hist && live[c] && tmod == 0) {          // MCout::add: one row per chain, coalesced over the lanes
            double *row = p.hist + ((long long)tring * p.C + jb + c) * (D + 1);
            row[i0] = x0[c]; row[i0 + 1] = x1[c];
            if (r == 0) row[D] = ly[c];
          }
          if (k == 0) {                                  // the running moments join here (after the likelihood: fewer live registers before)
            const double2 mv = *reinterpret_cast<const double2 *>(p.mu + jc[c] * D + i0), pv = *reinterpret_cast<const double2 *>(p.ps + jc[c] * D + i0);
            mu0[c] = mv.x; mu1[c] = mv.y; ps0[c] = pv.x; ps1[c] = pv.y;
          }
          if (REMOTE && a) {                             // adopt the component's moments, mcpar.cc:190-197
            const double sd0 = __ldg(p.psd + (size_t)i0 * p.mpad + cpick[c]), sd1 = __ldg(p.psd + (size_t)(i0 + 1) * p.mpad + cpick[c]);
            mu0[c] = __ldg(p.pmh + (size_t)i0 * p.mpad + cpick[c]).x; mu1[c] = __ldg(p.pmh + (size_t)(i0 + 1) * p.mpad + cpick[c]).x;
            ps0[c] = (sd0 * sd0) * (pwgt - 1.0); ps1[c] = (sd1 * sd1) * (pwgt - 1.0);
          }
          double dl = x0[c] - mu0[c]; mu0[c] += dl * winv; ps0[c] += dl * (x0[c] - mu0[c]);
          dl = x1[c] - mu1[c]; mu1[c] += dl * winv; ps1[c] += dl * (x1[c] - mu1[c]);
        }
      }
      if (MAIN) { if (++tmod == p.thin) { tmod = 0; if (++tring == p.hist_cap) tring = 0; } }
    }

#pragma unroll
    for (int c = 0; c < NCW; ++c) {
      if (live[c]) {
        const long long j = jb + c;
        *reinterpret_cast<double2 *>(p.x + j * D + i0) = make_double2(x0[c], x1[c]);
        if (r == 0) p.ly[j] = ly[c];
        if (MAIN) {
          *reinterpret_cast<double2 *>(p.mu + j * D + i0) = make_double2(mu0[c], mu1[c]);
          *reinterpret_cast<double2 *>(p.ps + j * D + i0) = make_double2(ps0[c], ps1[c]);
          const long long gg = p.chain0 + j;
          if (p.pool_next && gg % p.pool_stride == 0 && gg / p.pool_stride < p.pool_m) {     // musigall slot rule, mcpar.cc:205-208
            const long long s = gg / p.pool_stride;
            const double wi = 1.0 / (double)(p.t0 + p.nsteps);
            if (p.npeers > 0) {                          // sharded: store into every GPU's next pool over NVLink
              wait_arrivals_thread(p.arrivals, p.pub_wait_target, p.xflag);   // never more than one publication ahead (mh_kernels.cuh)
              for (int q = 0; q < p.npeers; ++q) {
                double *dst = reinterpret_cast<double *>(p.peers[q] + p.next_off);
                dst[(s * D + i0) * 2] = mu0[c];     dst[(s * D + i0) * 2 + 1] = ps0[c] * wi;
                dst[(s * D + i0 + 1) * 2] = mu1[c]; dst[(s * D + i0 + 1) * 2 + 1] = ps1[c] * wi;
              }
              __threadfence_system();                    // each lane's stores are visible before its arrival
              for (int q = 0; q < p.npeers; ++q)
                atomicAdd_system(reinterpret_cast<unsigned long long *>(p.peers[q] + p.arr_off), 1ull);
            } else {
              p.pool_next[(s * D + i0) * 2] = mu0[c];     p.pool_next[(s * D + i0) * 2 + 1] = ps0[c] * wi;
              p.pool_next[(s * D + i0 + 1) * 2] = mu1[c]; p.pool_next[(s * D + i0 + 1) * 2 + 1] = ps1[c] * wi;
            }
          }
        }
        if (r == 0) { wacc += nacc[c]; ++nlive; }
      }
    }
  }

  // acceptance counters and main-phase statistics: one count per chain (lane 0 of its owner warp)
  if (r == 0 && nlive) {
    atomicAdd(p.counts, (unsigned long long)wacc);
    atomicAdd(p.counts + 1, (unsigned long long)nlive * (unsigned long long)p.nsteps);
    if (REMOTE) {                                       // counts[2] remote chain-steps, [3] candidates (one each), [6] exact-path fallbacks
      atomicAdd(p.counts + 2, (unsigned long long)nlive * (unsigned long long)p.nsteps);
      atomicAdd(p.counts + 3, (unsigned long long)nlive * (unsigned long long)p.nsteps);
      if (nfb) atomicAdd(p.counts + 6, (unsigned long long)nfb);
    }
  }
}

}  // namespace MCGPU_NS
}  // namespace mcgpu
