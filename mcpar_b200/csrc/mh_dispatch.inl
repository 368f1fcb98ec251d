// mh_dispatch.inl -- (likelihood, d, rng, phase) -> kernel instantiation table.
// Included inside namespace mcgpu::MCGPU_NS by mh_fast.cu and mh_exact.cu.

// CTA size of the step kernels (<= 128, the kernels' launch bound); MCGPU_BLOCK env overrides for tuning
static int step_block()
{
  static int b = 0;
  if (!b) { const char *e = getenv("MCGPU_BLOCK"); b = e ? atoi(e) : 128; if (b != 32 && b != 64 && b != 128) b = 128; }
  return b;
}

template <int LIK, int D>
static cudaError_t launch_lik_d(int rngk, int phase, const StepParams &p, size_t smem, cudaStream_t st)
{
  const int MCGPU_BLOCK = step_block();
  const unsigned grid = (unsigned)((p.C + MCGPU_BLOCK - 1) / MCGPU_BLOCK);
#define MCGPU_GO(R, M)                                                                          \
  do {                                                                                          \
    if (smem > 48 * 1024)                                                                       \
      cudaFuncSetAttribute(mh_steps_kernel<LIK, D, R, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    mh_steps_kernel<LIK, D, R, M><<<grid, MCGPU_BLOCK, smem, st>>>(p);                          \
  } while (0)
#ifdef MCGPU_EXACT_TU
  if (rngk == RNG_REPLAY && phase == PH_BURN) MCGPU_GO(RNG_REPLAY, PH_BURN);
  else if (rngk == RNG_REPLAY && phase == PH_LOCAL) MCGPU_GO(RNG_REPLAY, PH_LOCAL);
#else
  if (rngk == RNG_PHILOX && phase == PH_BURN) MCGPU_GO(RNG_PHILOX, PH_BURN);
  else if (rngk == RNG_PHILOX && phase == PH_MIXED) MCGPU_GO(RNG_PHILOX, PH_MIXED);
  else if (rngk == RNG_PHILOX && phase == PH_LOCAL) MCGPU_GO(RNG_PHILOX, PH_LOCAL);
  else if (rngk == RNG_PHILOX && phase == PH_REMOTE) MCGPU_GO(RNG_PHILOX, PH_REMOTE);
  else if (rngk == RNG_PHILOX && phase == PH_MIXED_SUM) MCGPU_GO(RNG_PHILOX, PH_MIXED_SUM);
  else if (rngk == RNG_PHILOX && phase == PH_REMOTE_SUM) MCGPU_GO(RNG_PHILOX, PH_REMOTE_SUM);
#endif
  else return cudaErrorInvalidValue;
#undef MCGPU_GO
  return cudaGetLastError();
}

template <int LIK>
static cudaError_t launch_lik(int d, int rngk, int phase, const StepParams &p, size_t smem, cudaStream_t st)
{
  switch (d) {
    case 2:  return launch_lik_d<LIK, 2>(rngk, phase, p, smem, st);
    case 4:  return launch_lik_d<LIK, 4>(rngk, phase, p, smem, st);
    case 8:  return launch_lik_d<LIK, 8>(rngk, phase, p, smem, st);
    case 16: return launch_lik_d<LIK, 16>(rngk, phase, p, smem, st);
  }
  return cudaErrorInvalidValue;
}

bool steps_supported(int lik, int d)
{
  if (lik == MCGPU_ROSENBROCK1 || lik == MCGPU_GAUSSMIX) return d == 2 || d == 4 || d == 8 || d == 16;
  if (lik == MCGPU_GAUSSIAN || lik == MCGPU_DUALGAUSSIAN) return d == 2;
  return false;
}

// pool (only the phases that can take a remote step stage it): (mu, h) + sigma in fp64, (g mu, g) + (sigma, mu)
// in fp32 = 40 bytes per slot and parameter, + n_s (fp64) and nb_s (fp32) per slot
size_t steps_smem_bytes(int d, int nsteps, int pool_m, bool with_pool)
{
  const size_t mpad = (size_t)((pool_m + 7) & ~7);
  return sizeof(double) * ((size_t)MCGPU_MATH_SMEM + (size_t)d * d + (size_t)((nsteps + 1) & ~1) + (with_pool ? mpad * d * 5 + mpad * 2 : 0));
}

cudaError_t launch_steps(int lik, int d, int rngk, int phase, const StepParams &p, cudaStream_t st)
{
  const size_t smem = steps_smem_bytes(d, p.nsteps, p.pool_m, phase != PH_BURN && phase != PH_LOCAL);
  switch (lik) {
    case MCGPU_ROSENBROCK1:  return launch_lik<MCGPU_ROSENBROCK1>(d, rngk, phase, p, smem, st);
    case MCGPU_GAUSSMIX:     return launch_lik<MCGPU_GAUSSMIX>(d, rngk, phase, p, smem, st);
    case MCGPU_GAUSSIAN:     return d == 2 ? launch_lik_d<MCGPU_GAUSSIAN, 2>(rngk, phase, p, smem, st) : cudaErrorInvalidValue;
    case MCGPU_DUALGAUSSIAN: return d == 2 ? launch_lik_d<MCGPU_DUALGAUSSIAN, 2>(rngk, phase, p, smem, st) : cudaErrorInvalidValue;
  }
  return cudaErrorInvalidValue;
}

template <int LIK>
static cudaError_t init_lik(int d, const StepParams &p, cudaStream_t st)
{
  const unsigned grid = (unsigned)((p.C + 255) / 256);
  switch (d) {
    case 2:  init_loglik_kernel<LIK, 2><<<grid, 256, 0, st>>>(p); break;
    case 4:  init_loglik_kernel<LIK, 4><<<grid, 256, 0, st>>>(p); break;
    case 8:  init_loglik_kernel<LIK, 8><<<grid, 256, 0, st>>>(p); break;
    case 16: init_loglik_kernel<LIK, 16><<<grid, 256, 0, st>>>(p); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_init_loglik(int lik, int d, const StepParams &p, cudaStream_t st)
{
  const unsigned grid = (unsigned)((p.C + 255) / 256);
  switch (lik) {
    case MCGPU_ROSENBROCK1:  return init_lik<MCGPU_ROSENBROCK1>(d, p, st);
    case MCGPU_GAUSSMIX:     return init_lik<MCGPU_GAUSSMIX>(d, p, st);
    case MCGPU_GAUSSIAN:     if (d != 2) return cudaErrorInvalidValue;
                             init_loglik_kernel<MCGPU_GAUSSIAN, 2><<<grid, 256, 0, st>>>(p); return cudaGetLastError();
    case MCGPU_DUALGAUSSIAN: if (d != 2) return cudaErrorInvalidValue;
                             init_loglik_kernel<MCGPU_DUALGAUSSIAN, 2><<<grid, 256, 0, st>>>(p); return cudaGetLastError();
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_tune(unsigned long long *counts, unsigned long long *cum, double *factor, int dd,
                        double armin, double armax, double dfac, double ifac, cudaStream_t st)
{
  tune_kernel<<<1, 64, 0, st>>>(counts, cum, factor, dd, armin, armax, dfac, ifac);
  return cudaGetLastError();
}
