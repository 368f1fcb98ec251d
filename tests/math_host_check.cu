// tests/math_host_check.cu -- CPU accuracy check of mcpar_b200/csrc/mcgpu_math.cuh against
// long-double libm; prints max ulp errors as JSON.  Built and run by tests/test_math_host.py.
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#define MCGPU_TABLE_QUAL static
#include "../mcpar_b200/csrc/mcgpu_tables.h"
#include "../mcpar_b200/csrc/mcgpu_math.cuh"

static double ulp_err(double got, long double want)
{
  if (want == 0.0L) return fabs(got) / 4.9e-324;
  int e; frexpl(want, &e);
  const long double ulp = ldexpl(1.0L, e - 53);
  return (double)(fabsl((long double)got - want) / ulp);
}
static double rnd() { return (double)rand() / ((double)RAND_MAX + 1.0); }

int main()
{
  mcgpu::MathTables T = {MCGPU_EXP_TABLE, MCGPU_LOG_TABLE, MCGPU_TRIG_TABLE};
  srand(12345);
  double e_exp = 0, e_log = 0, e_log01 = 0, e_sin = 0, e_cos = 0, abs_sc = 0;
  for (int i = 0; i < 2000000; ++i) {
    double x = (rnd() - 0.5) * (i & 1 ? 1400.0 : 40.0);
    if (x < -707.9 || x > 709) continue;
    double er = ulp_err(mcgpu::mc_exp(x, T), expl((long double)x));
    if (er > e_exp) e_exp = er;
  }
  for (int i = 0; i < 2000000; ++i) {
    double x = exp((rnd() - 0.5) * 200.0);
    double er = ulp_err(mcgpu::mc_log(x, T), logl((long double)x));
    if (fabs(x - 1.0) > 0.02 && er > e_log) e_log = er;
    double v = 1.0 - rnd() * (i & 1 ? 1.0 : 1e-3);            // Box-Muller argument 1-u, including v ~ 1
    if (v <= 0) continue;
    double ab = fabs((double)((long double)mcgpu::mc_log(v, T) - logl((long double)v)));
    if (ab > e_log01) e_log01 = ab;                           // absolute error near log = 0
  }
  for (int i = 0; i < 2000000; ++i) {
    double u = rnd();
    double s, c; mcgpu::mc_sincos2pi(u, s, c, T);
    long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)u;
    double es = fabs((double)((long double)s - sinl(a))), ec = fabs((double)((long double)c - cosl(a)));
    if (es > abs_sc) abs_sc = es; if (ec > abs_sc) abs_sc = ec;
    if (fabsl(sinl(a)) > 0.1L) { double er = ulp_err(s, sinl(a)); if (er > e_sin) e_sin = er; }
    if (fabsl(cosl(a)) > 0.1L) { double er = ulp_err(c, cosl(a)); if (er > e_cos) e_cos = er; }
  }
  // mc_sqrt_pos with the SFU seed off by +-2^-21 (twice the documented bound), mc_log_pos on its two call sites
  double e_sqrt = 0, e_logpos = 0;
  for (int i = 0; i < 2000000; ++i) {
    const double L = (i & 3) == 0 ? rnd() * 1e-9 : rnd() * 44.5;
    const double er = ulp_err(mcgpu::mc_sqrt_pos(L, (i & 4) ? 4.76837158203125e-7 : -4.76837158203125e-7), sqrtl((long double)L));
    if (er > e_sqrt) e_sqrt = er;
    const double v = (i & 1) ? ((double)(unsigned)(rnd() * 4294967295.0) + 1.0) * (1.0 / 4294967296.0) : 1.0 + rnd();
    if (fabs(v - 1.0) > 0.02) { const double el = ulp_err(mcgpu::mc_log_pos(v, T), logl((long double)v)); if (el > e_logpos) e_logpos = el; }
    if (mcgpu::mc_log_pos(v, T) != mcgpu::mc_log(v, T)) e_logpos = 1e9;       // same arithmetic as the checked routine
  }
  // mc_exp_nonpos: the branch-free variant must agree with mc_exp bit for bit on x <= 0 and flush below -708
  int nonpos_bad = 0;
  for (int i = 0; i < 2000000; ++i) {
    const double x = -rnd() * (i & 1 ? 900.0 : 40.0);
    if (mcgpu::mc_exp_nonpos(x, T) != mcgpu::mc_exp(x, T)) ++nonpos_bad;
  }
  if (mcgpu::mc_exp_nonpos(-INFINITY, T) != 0.0 || mcgpu::mc_exp_nonpos(-0.0, T) != 1.0 || mcgpu::mc_exp_nonpos(-1.0e300, T) != 0.0) ++nonpos_bad;
  const double sq0 = mcgpu::mc_sqrt_pos(0.0);
  double sp[4] = {mcgpu::mc_exp(-800.0, T), mcgpu::mc_exp(800.0, T), mcgpu::mc_exp(NAN, T), mcgpu::mc_log(0.0, T)};
  printf("{\"exp_ulp\": %.3f, \"log_ulp\": %.3f, \"log_near1_abs\": %.3e, \"sin_ulp\": %.3f, \"cos_ulp\": %.3f, "
         "\"sqrt_ulp\": %.3f, \"sqrt0\": %g, \"logpos_ulp\": %.3f, "
         "\"sincos_abs\": %.3e, \"exp_m800\": %g, \"exp_p800_inf\": %d, \"exp_nan\": %d, \"log0_minf\": %d, \"exp_nonpos_mismatches\": %d}\n",
         e_exp, e_log, e_log01, e_sin, e_cos, e_sqrt, sq0, e_logpos, abs_sc, sp[0], (int)isinf(sp[1]), (int)isnan(sp[2]), (int)(isinf(sp[3]) && sp[3] < 0), nonpos_bad);
  return 0;
}
