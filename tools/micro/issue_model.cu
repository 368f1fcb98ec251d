// issue_model.cu -- does an FP64 warp instruction hold the SMSP dispatch port for two cycles on sm_100a?
// Loops of NF independent DFMA and NI independent IMAD (or LOP3 / FFMA) per iteration, 16 warps per SMSP
// (latency hidden), one CTA of 512 threads per SM x 2.  Prints cycles per iteration per SMSP-warp-slot:
//    additive model   (port held):      2 NF + NI
//    overlapped model (port released):  max(2 NF, NF + NI)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_model issue_model.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NI, int KIND>
__global__ void __launch_bounds__(512, 2) k(double *out, unsigned *outi, int iters, double a, unsigned m, long long *cyc)
{
  double f[8]; unsigned u[8]; float g[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x * 1e-3 + i; u[i] = threadIdx.x + i; g[i] = threadIdx.x * 1e-3f + i; }
  const double b = a * 1e-9; const float af = (float)a, bf = (float)b; const unsigned m2 = m * 3u + 1u;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[(r * NF + i) & 7]) : "d"(a), "d"(b));
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int q = (r * NI + i) & 7;
        if (KIND == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[q]) : "r"(m), "r"(m2));            // IMAD
        else if (KIND == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[q]) : "r"(m), "r"(m2));  // LOP3
        else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(g[q]) : "f"(af), "f"(bf));                    // FFMA
      }
    }
  }
  const long long t1 = clock64();
  double s = 0; unsigned v = 0; float h = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += f[i]; v ^= u[i]; h += g[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + h; outi[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NF, int NI, int KIND>
void run(const char *name)
{
  const int nb = 148 * 2, iters = 2000;
  double *o; unsigned *oi; long long *c;
  cudaMalloc(&o, nb * 512 * 8); cudaMalloc(&oi, nb * 512 * 4); cudaMalloc(&c, nb * 8);
  k<NF, NI, KIND><<<nb, 512>>>(o, oi, 10, 0.999999, 1664525u, c);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<NF, NI, KIND><<<nb, 512>>>(o, oi, iters, 0.999999, 1664525u, c);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[296]; cudaMemcpy(h, c, nb * 8, cudaMemcpyDeviceToHost);
  double mc = 0; for (int i = 0; i < nb; ++i) mc += h[i]; mc /= nb;
  // per SMSP: 2 CTAs x 16 warps / 4 = 8 warps; instructions issued per SMSP per iteration = 8 warps x 8 x (NF + NI)
  const double per = mc / iters / (8.0 * 8.0);       // cycles per (NF DFMA + NI other) group per SMSP
  printf("%-6s NF=%d NI=%d  cycles/group %.3f   additive 2NF+NI = %d   overlapped max(2NF, NF+NI) = %d   (%.3f ms)\n",
         name, NF, NI, per, 2 * NF + NI, (2 * NF > NF + NI ? 2 * NF : NF + NI), ms);
  cudaFree(o); cudaFree(oi); cudaFree(c);
}

int main()
{
  run<4, 0, 0>("dfma");
  run<0, 4, 0>("imad"); run<0, 4, 1>("lop3"); run<0, 4, 2>("ffma");
  run<4, 2, 0>("imad"); run<4, 4, 0>("imad"); run<4, 8, 0>("imad"); run<2, 8, 0>("imad"); run<1, 8, 0>("imad");
  run<4, 2, 1>("lop3"); run<4, 4, 1>("lop3"); run<4, 8, 1>("lop3"); run<2, 8, 1>("lop3");
  run<4, 2, 2>("ffma"); run<4, 4, 2>("ffma"); run<4, 8, 2>("ffma"); run<2, 8, 2>("ffma");
  return 0;
}
