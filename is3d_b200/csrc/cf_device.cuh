// cf_device.cuh -- device-side helpers for the Cooper-Frye kernels (sm_100a only).
//
//  * mbarrier + 1-D TMA bulk copy wrappers (cp.async.bulk, SASS: UBLKCP / SYNCS) used to stream cell tiles
//    from HBM/L2 into shared memory;
//  * exp_neg(): exp(-x) = 2^n T[j] e^r with -x = (64 n + j) ln2/64 + r: Cody-Waite reduction, a 64-entry table of 2^{j/64}
//    in shared memory, degree-5 polynomial -- 10 FP64-pipe instructions (the table-free degree-11 form needs 15), integer
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

// Constants live in constant memory so that DFMA reads them as c[bank][offset] operands (a literal costs two UMOV issue
// slots per use, which makes the loop issue-bound instead of FP64-bound).
//   kExpC: 1/5!, 1/4!, 1/3!, 1/2!          kExpR: -64/ln2, 1.5 * 2^52, -ln2_hi/64, -ln2_lo/64 (ln2_hi has 20 trailing zero bits:
//                                                 k * ln2_hi/64 is exact for |k| < 2^20; here |k| <= 709.8 * 64 / ln2 < 2^17)
__constant__ double kExpC[4] = {8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5};
__constant__ double kExpR[4] = {-92.33248261689366, 6755399441055744.0, -0.01083042469326756, -2.9815858269852933e-12};
// 2^{j/64}, j = 0..63, correctly rounded
__constant__ double kExpT[64] = {
  0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
  0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
  0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
  0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
  0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
  0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
  0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
  0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
  0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
  0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
  0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
  0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
  0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
  0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
  0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
  0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};

// Per-block copy of kExpT in shared memory: the index differs from lane to lane, which constant memory would serialise.
// Every kernel that calls exp_neg*() must run exp_table_init() and a __syncthreads() first.
__shared__ double g_exp_tab[64];
__device__ __forceinline__ void exp_table_init()
{ for (int i = threadIdx.x; i < 64; i += blockDim.x) g_exp_tab[i] = kExpT[i]; }

// Mantissa p in [0.99, 1.99] and binary exponent n with e^{-x} = p 2^n, for |x| < 2^17 ln2/64.
//   k = round(-64 x / ln2) = 64 n + j,  r = -x - k ln2/64 (|r| <= ln2/128),  e^r = 1 + r Q(r),  p = T_j + (T_j r) Q
// truncation r^6/6! < 3.5e-17; total error ~1 ulp (table entry, T r, final fma).
__device__ __forceinline__ void exp_neg_poly(double x, double &p_out, int &n_out)
{
  const double MAGIC = kExpR[1];                       // 1.5 * 2^52: round-to-nearest integer lands in the low word
  const double fk = fma(x, kExpR[0], MAGIC);
  const int k = __double2loint(fk);
  const double kf = fk - MAGIC;
  double r = fma(kf, kExpR[2], -x);                    // hi / lo split of ln2/64
  r = fma(kf, kExpR[3], r);
  const double T = g_exp_tab[k & 63];
  n_out = k >> 6;
  double q = fma(kExpC[0], r, kExpC[1]);
  q = fma(q, r, kExpC[2]);
  q = fma(q, r, kExpC[3]);
  q = fma(q, r, 1.0);
  p_out = fma(T * r, q, T);
}
// normal result: add n to the exponent field (ALU pipe, no FP64 slot)
__device__ __forceinline__ double exp_neg_fast(double p, int n)
{ return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p)); }
__device__ __forceinline__ bool exp_neg_is_rare(int n) { return (unsigned)(n + 1021) > 2044u; }
// rare: subnormal result (708.4 < x <= 709.78).  Two-step scaling keeps the gradual-underflow rounding.
__device__ __forceinline__ double exp_neg_rare(double p, int n)
{
  const int n1 = n >> 1, n2 = n - n1;
  return (p * __hiloint2double((n1 + 1023) << 20, 0)) * __hiloint2double((n2 + 1023) << 20, 0);
}

// exp(-x) for x <= ln(DBL_MAX) (callers skip larger x: the reference's exp(x) overflows there and the term is 0).
// The staged pieces above let groups of evaluations interleave their chains and test the sub-normal case once per group.
__device__ __forceinline__ double exp_neg(double x)
{
  double p; int n;
  exp_neg_poly(x, p, n);
  return __builtin_expect(exp_neg_is_rare(n), 0) ? exp_neg_rare(p, n) : exp_neg_fast(p, n);
}

// true when the reference's exp(x) stays finite, i.e. x <= ln(DBL_MAX) = 0x40862E42FEFA39EF; integer compare on
// the ALU pipe (negative x has the sign bit set and passes).
__device__ __forceinline__ bool exp_finite(double x)
{ return __double_as_longlong(x) <= 0x40862E42FEFA39EFLL; }

// Grouped evaluation paths: cheap aliveness test on the high word only (one ISETP).  It lets through the sliver
// ln(DBL_MAX) < x < 709.78271484375 (same high word as the threshold); those arguments always take the sub-normal ("rare")
// branch, which applies the exact test and returns 0.
__device__ __forceinline__ bool exp_alive_hi(double x) { return __double2hiint(x) <= 0x40862E42; }

// acc += pds * f for an alive group member whose p.dsigma passes the outflow test; the test looks at the high word only
// (thr_hi = 0: p.dsigma > 0, values below 2^-1042 count as 0;  thr_hi = INT_MIN: outflow off).  Compiles to one
// ISETP.GT.AND and a predicated DFMA -- no selects on the chain.
// (x is the exponent argument of the member: its aliveness is re-derived here so that the predicate never leaves the
// predicate registers; written in C++ the compiler turns the two conditions into four FSELs per evaluation.)
__device__ __forceinline__ void accumulate_alive(double &acc, double pds, double f, int thr_hi, double x)
{
  asm("{\n\t.reg .pred p, q;\n\t"
      "setp.le.s32 q, %1, 0x40862E42;\n\t"
      "setp.gt.and.s32 p, %2, %3, q;\n\t"
      "@p fma.rn.f64 %0, %4, %5, %0;\n\t}"
      : "+d"(acc) : "r"(__double2hiint(x)), "r"(__double2hiint(pds)), "r"(thr_hi), "d"(pds), "d"(f));
}

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
