// cf_shift.cu -- "shifted-factor" hot kernel for the linear-delta-f models on 3+1D tiles (M_LIN14, M_LINCE, M_JONAHLIN, M_IDEAL).
//
// Same decomposition as cf_kernel (cf_kernels.cu): lane <-> (species, pT), thread = register tile of NYT slots x NPT phi points,
// block = 4 warps on one (y tile, phi tile, cell chunk), records streamed by TMA through an mbarrier pipeline, pair table for the
// lane-independent shear cross term.  What changes is how the NPT evaluations of a (cell, slot) get their exponentials and how the
// per-evaluation integer work (regulation, outflow) is paid:
//
//  * e^{-x_jk} with x_jk = mT A_j - pT B_k is split as e^{-(mT A_j - pT Bmax)} . e^{-pT (Bmax - B_k)}, Bmax = max_k B_k over the
//    tile's phi points.  Both factors lie in (0, 1]: neither can overflow or underflow before the product does, so no exponent
//    bookkeeping is needed (compare the mantissa / exponent pairs of cf_factored.cu).  The first factor is ONE exponential per
//    (cell, slot) and thread -- of the group's smallest argument xm, which also classifies the whole group (dead / dilute /
//    possibly sub-normal) -- the second does not depend on the slot or the species: the block tabulates it once per tile for
//    (cell, phi, pT) in shared memory, 12 exponentials per thread and 16-cell tile instead of 3 per thread and cell.
//    Per evaluation that leaves a DMUL where cf_kernel spends 9 FP64 instructions.
//  * regulation and outflow cost 5 integer instructions per evaluation instead of 8: g = 1 + df is clamped to [0, 2] (the
//    reference's max(-1, min(df, 1)), smooth_kernels.cpp:328) on its high word and zeroed where p.dsigma <= 0 (:285) by the same
//    selects (clamp_mask_g), followed by an unconditional accumulate.
//  * groups whose largest argument may reach the sub-normal range (xm + pT (Bmax - Bmin) >= 707.7) are evaluated member by
//    member with the exact routines of cf_kernel (distribution_group), so the zero pattern and gradual underflow are unchanged.
//
// On B200 a warp-wide FP64 instruction holds the scheduler for ~2 cycles and every other instruction for ~1.7 more without
// overlap (profiles/r2_ubench_mix_issue.txt): the kernel time follows the TOTAL instruction count, which is what this kernel cuts.
#include "cf_internal.h"
#include <algorithm>
#include "cf_device.cuh"
#include "cf_epilogue.cuh"

namespace is3d {

namespace {

constexpr int kShiftMaxPT = 64;           // pT points the per-tile e^{-pT dB} table is laid out for
constexpr int kUltraHi = 0x4042C000;      // high word of 37.5: e^{-x} < 2^-54

// g = 1 + df clamped to [0, 2] (regulate_deltaf, smooth_kernels.cpp:328) and zeroed where p.dsigma fails the outflow test (:285),
// all on the integer pipe from the high words: 6 instructions (2 VIMNMX, 2 ISETP, 2 SEL) for both features.
//   reg_lo / reg_hi / reg_chk = 0 / 0x40000000 / 0x3fffffff with regulation on (g keeps its low word iff its high word read as
//   unsigned is <= reg_chk), INT_MIN / INT_MAX / 0xffffffff off;  thr_hi = 0 with outflow on, INT_MIN off.
// A masked member contributes p.dsigma f_eq 0 = +-0 to the accumulator, i.e. nothing.
__device__ __forceinline__ double clamp_mask_g(double g, double pds, int reg_lo, int reg_hi, unsigned reg_chk, int thr_hi)
{
  int hi = __double2hiint(g), lo = __double2loint(g), hc, ho, lw;
  asm("{\n\t.reg .pred pass, keep;\n\t"
      "setp.gt.s32 pass, %3, %4;\n\t"                 // p.dsigma passes the outflow test
      "setp.le.and.u32 keep, %5, %6, pass;\n\t"       // ... and g lies in [0, 2): keeps its low word
      "max.s32 %0, %5, %7;\n\t"
      "min.s32 %0, %0, %8;\n\t"                       // (VIMNMX3)
      "selp.b32 %1, %0, 0, pass;\n\t"
      "selp.b32 %2, %9, 0, keep;\n\t}"
      : "=&r"(hc), "=r"(ho), "=r"(lw) : "r"(__double2hiint(pds)), "r"(thr_hi), "r"(hi), "r"(reg_chk), "r"(reg_lo), "r"(reg_hi), "r"(lo));
  return __hiloint2double(ho, lw);
}

// e^{-x} = p 2^n like exp_neg_poly(); the table is read through a precomputed shared-memory address (the generic form rebuilds
// the shared window base with uniform-datapath instructions per use) and the constants come from the kernel parameters, which
// an FP64 instruction can take as a direct operand (from __constant__ arrays they are re-loaded with LDC under register pressure)
__device__ __forceinline__ void exp_neg_poly_at(double x, uint32_t tab_addr, const double (&ec)[8], double &p_out, int &n_out)
{
  const double fk = fma(x, ec[0], ec[1]);
  const int k = __double2loint(fk);
  const double kf = fk - ec[1];
  double r = fma(kf, ec[2], -x);
  r = fma(kf, ec[3], r);
  double T;
  asm("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(tab_addr + ((k & (kExpTabSize - 1)) << 3)));
  n_out = k >> kExpTabBits;
  double q = fma(ec[4], r, ec[5]);
  q = fma(q, r, ec[6]);
  q = fma(q, r, 1.0);
  p_out = fma(T * r, q, T);
}

template <int MODEL, int NYT, int NPT, int MINB>
__global__ void __launch_bounds__(128, MINB)
cf_shift_kernel(const HotParams hp)
{
  constexpr int RY = kRec;
  constexpr bool DF = MODEL != M_IDEAL;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Layout &L = hp.L;
  const int CT = L.ct;
  const int y_doubles = CT * NYT * RY, p_doubles = CT * NPT * kRec, s_doubles = CT * kScal;
  const int stage_doubles = y_doubles + p_doubles + s_doubles;
  double *stage_base = reinterpret_cast<double *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(stage_base + (size_t)kStages * stage_doubles);
  constexpr int NPP = (NPT + 1) & ~1;                                            // pair-table row padded to whole 16-byte loads
  double *pair_tab = reinterpret_cast<double *>(full + kStages);                 // [CT][NYT][NPP]
  double *eb_tab = pair_tab + CT * NYT * NPP;                                    // [CT][NPT][n_pT]  e^{-pT (Bmax - B_k)}
  double *cellv = eb_tab + CT * NPT * L.n_pT;                                    // [CT][2]          Bmax, Bmax - Bmin
  double *pT_s = cellv + CT * 2;                                                 // [n_pT]

  const int n_bintiles = hp.n_groupblocks * L.n_ytiles * L.n_ptiles;
  const int chunk = blockIdx.x / n_bintiles;
  int bt = blockIdx.x - chunk * n_bintiles;
  const int tp = bt % L.n_ptiles; bt /= L.n_ptiles;
  const int ty = bt % L.n_ytiles; bt /= L.n_ytiles;
  const int gb = bt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int idx = (gb * hp.n_warps + warp) * 32 + lane;
  const bool lane_valid = idx < L.n_species * L.n_pT;
  const int ipart = lane_valid ? idx / L.n_pT : 0;
  const int ipT = lane_valid ? idx - ipart * L.n_pT : 0;
  const double mass = hp.mass[ipart], sign = hp.sign[ipart], pT = hp.pT[ipT];
  const double m2 = mass * mass, pT2 = pT * pT;
  const double mT2 = m2 + pT2;
  const double mT = sqrt(mT2);
  const double mTpT = mT * pT;
  const unsigned reg_chk = hp.reg_chk;
  const int reg_lo = hp.reg_lo, reg_hi = hp.reg_hi;
  const int thr_hi = (int)(hp.outflow_thr >> 32);
  const int reg_thr = hp.regulate_thr, one_hi = hp.one_hi;

  const int64_t t_begin = (L.n_tiles * (int64_t)chunk) / hp.n_chunks;
  const int64_t t_end = (L.n_tiles * (int64_t)(chunk + 1)) / hp.n_chunks;
  const int n_my_tiles = (int)(t_end - t_begin);

  const double *Yg = hp.Y + ((int64_t)ty * L.n_cells_pad) * NYT * RY;
  const double *Pg = hp.P + ((int64_t)tp * L.n_cells_pad) * NPT * kRec;
  const double *Sg = hp.S;
  const uint32_t stage_bytes = (uint32_t)stage_doubles * 8u;

  exp_table_init();
  for (int i = threadIdx.x; i < L.n_pT; i += blockDim.x) pT_s[i] = hp.pT[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; s++) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int t_local) {
    const int st = t_local % kStages;
    const int64_t cell = (t_begin + t_local) * CT;
    double *dst = stage_base + (size_t)st * stage_doubles;
    mbar_arrive_expect_tx(&full[st], stage_bytes);
    bulk_g2s(dst, Yg + cell * NYT * RY, (uint32_t)y_doubles * 8u, &full[st]);
    bulk_g2s(dst + y_doubles, Pg + cell * NPT * kRec, (uint32_t)p_doubles * 8u, &full[st]);
    bulk_g2s(dst + y_doubles + p_doubles, Sg + cell * kScal, (uint32_t)s_doubles * 8u, &full[st]);
  };
  if (threadIdx.x == 0)
    for (int t = 0; t < kStages && t < n_my_tiles; t++) issue(t);

  double acc[NYT * NPT];
#pragma unroll
  for (int i = 0; i < NYT * NPT; i++) acc[i] = 0.0;
  const uint32_t tab_addr = smem_u32(g_exp_tab);
  const double rare_x = __hiloint2double(kRareHi, 0);                            // 707.6997...: below it e^{-x} is a normal number

  for (int t = 0; t < n_my_tiles; t++) {
    const int st = t % kStages;
    mbar_wait(&full[st], (uint32_t)((t / kStages) & 1));
    const double *Ys = stage_base + (size_t)st * stage_doubles;
    const double *Ps = Ys + y_doubles;
    const double *Ss = Ps + p_doubles;

    // ---- block phase: tables that do not depend on the species
    if (DF) {
      for (int w = threadIdx.x; w < CT * NYT * NPT; w += blockDim.x) {             // pair table, as in cf_kernel
        const int c = w / (NYT * NPT), r = w - c * (NYT * NPT), j = r / NPT, k = r - j * NPT;
        const double *yr = Ys + (c * NYT + j) * RY, *pr = Ps + (c * NPT + k) * kRec;
        pair_tab[(c * NYT + j) * NPP + k] = fma(pr[4], yr[4], -(pr[3] * yr[3]));
      }
    }
    for (int c = warp; c < CT; c += 4) {                                         // e^{-pT (Bmax - B_k)}: one cell per warp-iteration
      double b[NPT];
#pragma unroll
      for (int k = 0; k < NPT; k++) b[k] = Ps[(c * NPT + k) * kRec];
      double bmax = b[0], bmin = b[0];
#pragma unroll
      for (int k = 1; k < NPT; k++) { bmax = b[k] > bmax ? b[k] : bmax; bmin = b[k] < bmin ? b[k] : bmin; }
      if (lane == 0) { cellv[2 * c] = bmax; cellv[2 * c + 1] = bmax - bmin; }
      for (int ip = lane; ip < L.n_pT; ip += 32) {
        const double pTi = pT_s[ip];
#pragma unroll
        for (int k = 0; k < NPT; k++) {
          double pp; int nn;
          const double dd = pTi * (bmax - b[k]);
          exp_neg_poly_at(dd, tab_addr, hp.ec, pp, nn);
          // beyond 707.7 the factor would leave the normal range: such members are only reached on the per-member path; 0 keeps
          // the lanes that merely ride along finite
          eb_tab[(c * NPT + k) * L.n_pT + ip] = __double2hiint(dd) < kRareHi ? exp_neg_fast(pp, nn) : 0.0;
        }
      }
    }
    __syncthreads();

    for (int c = 0; c < CT; c++) {
      const double2 k01 = *reinterpret_cast<const double2 *>(Ss + c * kScal);
      const double K2 = k01.y;
      const double K3 = (MODEL == M_JONAHLIN) ? Ss[c * kScal + 2] : 0.0;
      const double K0m = k01.x * m2;
      const double2 bb = *reinterpret_cast<const double2 *>(cellv + 2 * c);
      const double qm = pT * bb.x;                                               // largest pT B_k of the tile
      const int rare_hi = __double2hiint(fma(-pT, bb.y, rare_x));                 // xm >= 707.7 - pT (Bmax - Bmin): per-member path
      double q[NPT], pd[NPT], g0[NPT], eB[NPT];
#pragma unroll
      for (int k = 0; k < NPT; k++) {
        const double2 *pr = reinterpret_cast<const double2 *>(Ps + (c * NPT + k) * kRec);
        const double2 v0 = pr[0];
        q[k] = pT * v0.x;
        pd[k] = pT * v0.y;
        g0[k] = DF ? fma(pT2, pr[1].x, K0m) : 0.0;
        eB[k] = eb_tab[(c * NPT + k) * L.n_pT + ipT];
      }
#pragma unroll
      for (int j = 0; j < NYT; j++) {
        const double2 *yr = reinterpret_cast<const double2 *>(Ys + (c * NYT + j) * RY);
        const double2 v0 = yr[0];
        const double a = mT * v0.x;
        const double xm = a - qm;                                                // smallest argument of the group
        const int xh = __double2hiint(xm);
        // Control flow is WARP-uniform (votes): a lane whose group is dead (every member's exp(x) overflows: exact zeros) rides
        // along with e^{-xm} := 0, which makes its f_eq and therefore its contribution an exact 0.
        const bool alive = xh <= kAliveHi;
        if (!__any_sync(0xffffffffu, alive)) continue;
        const double cpm = mT * v0.y;
        const double h0 = DF ? mT2 * yr[1].x : 0.0;
        // (the eta weight of a 3+1D slot record is 1: p.dsigma = pT Dp + mT Cp)
        double pr2[NPP];
        if (DF) {
#pragma unroll
          for (int k = 0; k < NPP; k += 2) {
            const double2 t2 = *reinterpret_cast<const double2 *>(pair_tab + (c * NYT + j) * NPP + k);
            pr2[k] = t2.x; pr2[k + 1] = t2.y;
          }
        }
        double xs[NPT], pv[NPT], sv[NPT];
#pragma unroll
        for (int k = 0; k < NPT; k++) {
          xs[k] = a - q[k];
          pv[k] = pd[k] + cpm;
          sv[k] = DF ? fma(mTpT, pr2[k], h0 + g0[k]) : 0.0;
        }
        double *accj = acc + j * NPT;
        if (__builtin_expect(__any_sync(0xffffffffu, alive && xh >= rare_hi), 0)) {
          // some member of some lane may be sub-normal, or dead inside an alive group: member by member, exactly as cf_kernel
          double fv[NPT];
          distribution_group<MODEL, NPT>(xs, true, false, sv, K2, K3, sign, reg_thr, one_hi, fv);
#pragma unroll
          for (int k = 0; k < NPT; k++) accumulate_pos(accj[k], pv[k], fv[k], thr_hi);
          continue;
        }
        double pe; int ne;
        exp_neg_poly_at(xm, tab_addr, hp.ec, pe, ne);
        const double eA = alive ? exp_neg_fast(pe, ne) : 0.0;
        double av[NPT], dfs[NPT], feq[NPT], fb[NPT];
#pragma unroll
        for (int k = 0; k < NPT; k++) {
          av[k] = eA * eB[k];
          if (MODEL == M_LIN14) dfs[k] = fma(K2 * xs[k], xs[k], sv[k]);
          else if (DF) dfs[k] = fma(sv[k], rcp_fast(xs[k]), K2 * xs[k]);
        }
        // a_k <= e^{-xm} for every member: 1 / (1 + Theta a) is 1 to half an ulp once xm >= 37.5 (a < 2^-54; 57 % of the alive
        // groups of a whole warp on the cfg3 surface), 1 - Theta a + a^2 to 5e-17 once xm >= 12.5 (a < 2^-18; 35 %), else MUFU + 3 DFMA
        // g = 1 + df is formed inside the branch (no select between the two forms afterwards)
        const double K31 = (MODEL == M_JONAHLIN) ? K3 + 1.0 : 1.0;
        double g[NPT];
        if (__all_sync(0xffffffffu, xh >= kUltraHi)) {                           // (dead lanes have xh > kAliveHi)
#pragma unroll
          for (int k = 0; k < NPT; k++) { feq[k] = av[k]; g[k] = DF ? dfs[k] + K31 : 1.0; }
        } else if (__all_sync(0xffffffffu, xh >= kDiluteHi)) {
#pragma unroll
          for (int k = 0; k < NPT; k++) { feq[k] = occupation_bar_dilute(av[k], sign, fb[k]); g[k] = DF ? fma(fb[k], dfs[k], K31) : 1.0; }
        } else {
#pragma unroll
          for (int k = 0; k < NPT; k++) { feq[k] = occupation_bar(av[k], sign, fb[k]); g[k] = DF ? fma(fb[k], dfs[k], K31) : 1.0; }
        }
        if (!DF) {
#pragma unroll
          for (int k = 0; k < NPT; k++) accumulate_pos(accj[k], pv[k], feq[k], thr_hi);
          continue;
        }
        // f = f_eq g with g clamped to [0, 2]; members failing the outflow test get g = 0.  (A warp holds all pT values of a
        // species and the largest ones are always clamped: a "nobody needs it" shortcut would never fire.)
#pragma unroll
        for (int k = 0; k < NPT; k++)
          accj[k] = fma(pv[k], feq[k] * clamp_mask_g(g[k], pv[k], reg_lo, reg_hi, reg_chk, thr_hi), accj[k]);
      }
    }
    __syncthreads();                                   // every warp is done with stage st and with the tables
    if (threadIdx.x == 0 && t + kStages < n_my_tiles) issue(t + kStages);
  }

  hot_epilogue<NYT, NPT, false>(hp, acc, stage_base, chunk, gb, ty, tp, lane_valid, ipart, ipT);
}

struct ShiftShape { int nyt, npt, ct, minb; };
const ShiftShape kShiftShapes[kNumShiftVariants] = {{7, 3, 16, 4}, {7, 3, 16, 3}, {7, 4, 16, 3}, {7, 6, 16, 2}};

template <int MODEL, int NYT, int NPT, int MINB>
cudaError_t launch_shift_one(const HotParams &hp, cudaStream_t st, size_t *smem_out)
{
  const Layout &L = hp.L;
  const int stage_doubles = L.ct * (NYT * kRec + NPT * kRec + kScal);
  const size_t pipe = (size_t)kStages * stage_doubles * 8 + kStages * 8
                    + ((size_t)L.ct * NYT * ((NPT + 1) & ~1) + (size_t)L.ct * NPT * L.n_pT + (size_t)L.ct * 2 + L.n_pT) * 8;
  const size_t smem = std::max(pipe, (hot_epilogue_scratch_bytes(hp, NYT, false, 128) + 15) & ~(size_t)15);
  if (smem_out) *smem_out = smem;
  auto kern = cf_shift_kernel<MODEL, NYT, NPT, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t grid = (int64_t)hp.n_groupblocks * L.n_ytiles * L.n_ptiles * hp.n_chunks;
  if (grid == 0) return cudaSuccess;
  kern<<<(unsigned)grid, 128, smem, st>>>(hp);
  return cudaGetLastError();
}

template <int MODEL>
cudaError_t launch_shift_model(const HotParams &hp, int v, cudaStream_t st, size_t *smem_out)
{
  switch (v) {
    case 0: return launch_shift_one<MODEL, 7, 3, 4>(hp, st, smem_out);
    case 1: return launch_shift_one<MODEL, 7, 3, 3>(hp, st, smem_out);
    case 2: return launch_shift_one<MODEL, 7, 4, 3>(hp, st, smem_out);
    case 3: return launch_shift_one<MODEL, 7, 6, 2>(hp, st, smem_out);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

bool shift_supported(int model, const Layout &L)
{
  return !L.dim2 && L.n_pT <= kShiftMaxPT && (model == M_LIN14 || model == M_LINCE || model == M_JONAHLIN || model == M_IDEAL);
}

void shift_variant_shape(int v, int *nyt, int *npt, int *ct, int *max_warps)
{
  if (v < 0 || v >= kNumShiftVariants) v = 0;
  *nyt = kShiftShapes[v].nyt; *npt = kShiftShapes[v].npt; *ct = kShiftShapes[v].ct; *max_warps = 4;
}

// the block is always 4 warps: hp.n_warps must be 4 and hp.n_groupblocks = ceil(n_species n_pT / 128)
cudaError_t launch_shift(int model, const HotParams &hp, int v, cudaStream_t st, size_t *smem_out)
{
  switch (model) {
    case M_LIN14: return launch_shift_model<M_LIN14>(hp, v, st, smem_out);
    case M_LINCE: return launch_shift_model<M_LINCE>(hp, v, st, smem_out);
    case M_JONAHLIN: return launch_shift_model<M_JONAHLIN>(hp, v, st, smem_out);
    case M_IDEAL: return launch_shift_model<M_IDEAL>(hp, v, st, smem_out);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace is3d
