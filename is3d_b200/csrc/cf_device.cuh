// cf_device.cuh -- device-side helpers for the Cooper-Frye kernels (sm_100a only).
//
//  * mbarrier + 1-D TMA bulk copy wrappers (cp.async.bulk, SASS: UBLKCP / SYNCS) used to stream cell tiles
//    from HBM/L2 into shared memory;
//  * exp_neg(): exp(-x) with an explicit reduction + degree-11 polynomial, ~15 FP64-pipe instructions, integer
//    exponent insertion on the ALU pipe, and the IEEE overflow semantics of the reference's
//    `1.0 / (exp(x) + sign)` (smooth_kernels.cpp:289): exactly 0 once exp(x) would overflow, gradual underflow below;
//  * rcp_fast(): MUFU.RCP64H seed + one cubically convergent correction (3 DFMA), relative error < 1e-15.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace is3d {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }

__device__ __forceinline__ void mbar_fence_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{ while (!mbar_try_wait(bar, parity)) { } }

// global -> shared bulk copy (TMA, non-tensor form); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rcp_fast(double b)
{
  double y0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(b));
  double e = fma(-b, y0, 1.0);
  e = fma(e, e, e);
  return fma(y0, e, y0);
}

// sqrt(v), v > 0: MUFU.RSQ64H seed + two Goldschmidt steps (7 FP64-pipe instructions), relative error < 1e-15
__device__ __forceinline__ double sqrt_fast(double v)
{
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(v));
  double g = v * y0, h = 0.5 * y0;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  return fma(g, r, g);
}

// exp(x) overflows to +inf in the reference for x > ln(DBL_MAX); there 1/(inf + sign) = 0 exactly
#define IS3D_EXP_OVERFLOW_X 709.782712893384

// Taylor coefficients 1/k!, k = 11 .. 2, in constant memory so that DFMA reads them as c[bank][offset] operands
// (a literal costs two UMOV issue slots per use, which makes the loop issue-bound instead of FP64-bound).
__constant__ double kExpC[10] = {
  2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04,
  1.388888888888889e-03, 8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5};
__constant__ double kExpR[4] = {-1.4426950408889634, 6755399441055744.0, -6.93147180369123816490e-01, -1.90821492927058770002e-10};

// exp(-x) for x <= ln(DBL_MAX) (callers skip larger x: the reference's exp(x) overflows there and the term is 0).
__device__ __forceinline__ double exp_neg(double x)
{
  const double MAGIC = kExpR[1];                       // 1.5 * 2^52: round-to-nearest integer lands in the low word
  double fn = fma(x, kExpR[0], MAGIC);
  int n = __double2loint(fn);
  double nf = fn - MAGIC;
  double r = fma(nf, kExpR[2], -x);                    // -x - n ln2 (hi, lo split)
  r = fma(nf, kExpR[3], r);
  // e^r on [-ln2/2, ln2/2], Taylor to r^11: truncation < 1e-14 relative
  double p = kExpC[0];
#pragma unroll
  for (int k = 1; k < 10; k++) p = fma(p, r, kExpC[k]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  // normal result -> add n to the exponent field (ALU pipe, no FP64 slot)
  double a = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  if (__builtin_expect((unsigned)(n + 1021) > 2044u, 0)) {
    // rare: subnormal result (708.4 < x <= 709.78).  Two-step scaling keeps the gradual-underflow rounding.
    const int n1 = n >> 1, n2 = n - n1;
    const double s1 = __hiloint2double((n1 + 1023) << 20, 0);
    const double s2 = __hiloint2double((n2 + 1023) << 20, 0);
    a = (p * s1) * s2;
  }
  return a;
}

// Staged form for groups of evaluations (lets the scheduler interleave the members' dependency chains):
//   exp_neg_poly: mantissa polynomial p and binary exponent n with e^{-x} = p 2^n;  exp_neg_fast: exponent insertion;
//   exp_neg_is_rare / exp_neg_rare: the sub-normal case, tested once per group.
__device__ __forceinline__ void exp_neg_poly(double x, double &p_out, int &n_out)
{
  const double MAGIC = kExpR[1];
  double fn = fma(x, kExpR[0], MAGIC);
  n_out = __double2loint(fn);
  double nf = fn - MAGIC;
  double r = fma(nf, kExpR[2], -x);
  r = fma(nf, kExpR[3], r);
#ifndef IS3D_EXP_ESTRIN
  // Horner: measured 16 % faster than the Estrin form below on B200 (fewer FP64 issues and registers win over depth)
  double p = kExpC[0];
#pragma unroll
  for (int k = 1; k < 10; k++) p = fma(p, r, kExpC[k]);
  p = fma(p, r, 1.0);
  p_out = fma(p, r, 1.0);
#else
  // Estrin evaluation of sum_{k<=11} r^k / k!: dependency depth 5 instead of 11, 3 extra multiplies.  kExpC[i] = 1/(11-i)!
  const double r2 = r * r;
  const double p01 = 1.0 + r;                               // 1 + r
  const double p23 = fma(kExpC[8], r, kExpC[9]);            // 1/2 + r/6
  const double p45 = fma(kExpC[6], r, kExpC[7]);            // 1/4! + r/5!
  const double p67 = fma(kExpC[4], r, kExpC[5]);
  const double p89 = fma(kExpC[2], r, kExpC[3]);
  const double pAB = fma(kExpC[0], r, kExpC[1]);
  const double r4 = r2 * r2;
  const double q0 = fma(p23, r2, p01);
  const double q1 = fma(p67, r2, p45);
  const double q2 = fma(pAB, r2, p89);
  const double r8 = r4 * r4;
  const double s0 = fma(q1, r4, q0);
  p_out = fma(q2, r8, s0);
#endif
}
__device__ __forceinline__ double exp_neg_fast(double p, int n)
{ return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p)); }
__device__ __forceinline__ bool exp_neg_is_rare(int n) { return (unsigned)(n + 1021) > 2044u; }
__device__ __forceinline__ double exp_neg_rare(double p, int n)
{
  const int n1 = n >> 1, n2 = n - n1;
  return (p * __hiloint2double((n1 + 1023) << 20, 0)) * __hiloint2double((n2 + 1023) << 20, 0);
}

// true when the reference's exp(x) stays finite, i.e. x <= ln(DBL_MAX) = 0x40862E42FEFA39EF; integer compare on
// the ALU pipe (negative x has the sign bit set and passes).
__device__ __forceinline__ bool exp_finite(double x)
{ return __double_as_longlong(x) <= 0x40862E42FEFA39EFLL; }

// acc += pds * f when pds > thr (thr = +0 with outflow on: the reference skips p.dsigma <= 0, smooth_kernels.cpp:285;
// thr = LLONG_MIN with outflow off).  Sign/zero test on the integer pipe, predicated DFMA -- no select.
__device__ __forceinline__ void accumulate_outflow(double &acc, double pds, double f, long long thr)
{
  asm("{\n\t.reg .pred p;\n\tsetp.gt.s64 p, %1, %3;\n\t@p fma.rn.f64 %0, %2, %4, %0;\n\t}"
      : "+d"(acc) : "l"(__double_as_longlong(pds)), "d"(pds), "l"(thr), "d"(f));
}

// |v| >= 1 -> copysign(1, v), done on the integer pipe (regulate_deltaf, smooth_kernels.cpp:328).
// thr_hi = 0x3ff00000 when regulation is on, 0x7ff80000 (never reached by finite values) when off.
__device__ __forceinline__ double clamp_unit(double v, int thr_hi)
{
  const int hi = __double2hiint(v);
  const bool big = (hi & 0x7fffffff) >= thr_hi;
  const int hi2 = big ? ((hi & 0x80000000) | 0x3ff00000) : hi;
  const int lo2 = big ? 0 : __double2loint(v);
  return __hiloint2double(hi2, lo2);
}

}  // namespace is3d
