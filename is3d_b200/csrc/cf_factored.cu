// cf_factored.cu -- second-generation hot kernel for the linear-delta-f models in 3+1D (sm_100a): the loop nest of
// EmissionFunctionArray::calculate_dN_pTdpTdphidy (reference src/cpp/emissionfunction_smooth_kernels.cpp:246-347) for
// df_mode 1 (14 moment), 2 (Chapman-Enskog), Jonah's linearised df and the ideal f_eq, for species lists of >= 16 hadrons.
//
// Work decomposition
//   lane   <-> one SPECIES: the 32 lanes of a warp are 32 consecutive species of the chosen list (consecutive entries of a PDG
//              list have similar masses).  (cf_kernels.cu puts the 32 pT points of one species in a warp; pT spans 0.001-40 GeV
//              there, so a warp straddles dead / dilute / clamped evaluations and pays for every branch.)
//   block      one pT point x one group of 32 species x one y tile (NYT slots) x up to 4 phi tiles (NPT points each), ONE WARP PER
//              PHI TILE, x one contiguous cell chunk.  All warps of a block therefore share u.p's slot part mT A[slot]: its
//              exponential is evaluated once per (cell, slot, species) by the block and read back from shared memory.
//   thread     register tile of NYT x NPT accumulators, walks the cells of the chunk.
//   grid       pT points x phi blocks x y tiles x species groups x cell chunks; partial[chunk][bin] + reduce_kernel as in
//              cf_kernels.cu.  Cell tiles (kCTF cells: slot records, the block's phi records, scalars) are streamed global -> shared
//              with cp.async.bulk (TMA) through an mbarrier pipeline.
//
// Arithmetic per evaluation
//   1. Factored exponential: x = u.p/T = mT A[slot] - pT B[phi], so e^{-x} = e^{-mT A[slot]} e^{+pT B[phi]}, each factor kept as
//      mantissa x 2^n (exp_neg_poly).  Per tile the block fills shared-memory tables: the slot factor per species lane (4 warps
//      share it), the phi factor per phi point (pT is the block's, so it does not depend on the lane), pT D, pT^2 Qpp,
//      pT (R2 U2 - R1 U1).  An evaluation multiplies two mantissas (1 DMUL) and adds two binary exponents.
//   2. Classification without x: together with the slot table the block stores min / max of the slot exponents over its lanes;
//      min / max exponent of a (cell, slot) group of a warp is that plus min / max of the phi exponents.  A warp-UNIFORM compare
//      skips dead groups (every exp(x) overflows in the reference: f = 0 exactly), selects the occupation-factor form
//      (a = e^{-x} < 2^-18: 1 / (1 + Theta a) = 1 - Theta a + a^2; else MUFU seed + Newton), and tells whether any member is
//      near the overflow / sub-normal boundary.  Only such groups look at per-thread exponents; members within ~3 units of the
//      boundary go through late_member(), which decides from x itself exactly as cf_kernels.cu does.
//   3. g = 1 + df is formed directly (the reference multiplies f_eq (1 + df), :330); regulate_deltaf clamps g to [0, 2] on its
//      high word.  Apart from the dead-group skip and the choice of the occupation-factor form the inner loop has no branches:
//      profiling showed branch resolution and reconvergence (BSSY / BSYNC), not arithmetic, to be what the warps wait for.
//   4. The tables are double buffered: a tile costs ONE __syncthreads and its TMA stage is released before the inner loop.
#include "cf_internal.h"
#include <algorithm>
#include "cf_device.cuh"
#include "cf_epilogue.cuh"

namespace is3d {

namespace {

// exponent classes of a = mantissa (in [0.997, 3.99)) x 2^n
constexpr int kNDead = -1027;      // n <= kNDead: a < 2^-1025, x > 710.4 > ln(DBL_MAX): the term is exactly 0 in the reference
constexpr int kNNormal = -1021;    // n >= kNNormal: a >= 0.997 x 2^-1021, a normal number: exponent insertion is exact
constexpr int kNDilute = -20;      // n <= kNDilute: a < 2^-18
constexpr int kNForcedDead = -200000;

constexpr int kFW = 4;             // warps (= phi tiles) per block
constexpr int kCTF = 4;            // cells per TMA tile of this kernel (divides Layout::ct, which pads the record arrays)
constexpr int kStagesF = 3;        // TMA pipeline depth
constexpr int kDerP = 6;           // derived phi record (block's pT folded in): q = pT B, pd = pT D, G0 = pT^2 Qpp, fq, int fm, -
constexpr int kDerC = 4;           // derived cell record: K0, K2, 1 + K3, int2 {general-path flag, dead-cell flag}
constexpr double kQLimit = 2000.0; // |pT B| beyond this leaves the exact range of the Cody-Waite reduction: late_member() for the cell
constexpr unsigned kFull = 0xffffffffu;

template <int NPT> struct PairPitch { static constexpr int v = (NPT + 1) & ~1; };

// shared-memory map (offsets in doubles) for W warps: kStagesF stages {Y[kCTF][NYT][kRec], P[W][kCTF][NPT][kRec], S[kCTF][kScal]},
// then two table buffers {E[kCTF][NYT][32] double2 (mT A, mantissa), N[kCTF][NYT][32] int, CLS[kCTF][NYT] int2 (min, max exponent),
// DS[kCTF][NYT] double2 (Cp, Qyy), DP[W][kCTF][NPT][kDerP], DX[W][kCTF][NYT][PP], DC[kCTF][kDerC]}, then the mbarriers
template <int NYT, int NPT>
struct FMap {
  static constexpr int PP = PairPitch<NPT>::v;
  int W, y_doubles, p_doubles, stage_doubles, oE, oN, oCLS, oDS, oDP, oDX, oDC, der_doubles;
  __host__ __device__ explicit FMap(int w) : W(w)
  {
    y_doubles = kCTF * NYT * kRec; p_doubles = kCTF * NPT * kRec;
    stage_doubles = y_doubles + W * p_doubles + kCTF * kScal;
    oE = 0; oN = oE + kCTF * NYT * 32 * 2; oCLS = oN + kCTF * NYT * 32 / 2; oDS = oCLS + kCTF * NYT; oDP = oDS + kCTF * NYT * 2;
    oDX = oDP + W * kCTF * NPT * kDerP; oDC = oDX + W * kCTF * NYT * PP; der_doubles = oDC + kCTF * kDerC;
  }
  __host__ __device__ size_t bytes() const { return ((size_t)kStagesF * stage_doubles + 2 * (size_t)der_doubles) * 8 + kStagesF * sizeof(uint64_t); }
};

// clamp g = 1 + df to [0, 2] (regulate_deltaf: df in [-1, 1], smooth_kernels.cpp:328) on the high word: negative (sign bit) -> +0,
// >= 2 -> exactly 2.  With regulation on reg_lo / reg_hi / reg_chk are 0 / 0x40000000 / 0x3fffffff: g is in range, and keeps its
// low word, iff its high word read as unsigned is <= reg_chk.  Off: INT_MIN / INT_MAX / 0xffffffff (never clamps).
__device__ __forceinline__ bool needs_clamp(double g, unsigned reg_chk) { return (unsigned)__double2hiint(g) > reg_chk; }
__device__ __forceinline__ double clamp_g(double g, int reg_lo, int reg_hi, unsigned reg_chk)
{
  const int hi = __double2hiint(g);
  return __hiloint2double(min(max(hi, reg_lo), reg_hi), (unsigned)hi > reg_chk ? 0 : __double2loint(g));
}

// e^{-x} = p 2^n like exp_neg_poly(), with the table read through a precomputed shared-memory address (the generic form costs three
// uniform-datapath instructions per use to rebuild the window base)
__device__ __forceinline__ void exp_neg_poly_at(double x, uint32_t tab_addr, double &p_out, int &n_out)
{
  const double MAGIC = kExpR[1];
  const double fk = fma(x, kExpR[0], MAGIC);
  const int k = __double2loint(fk);
  const double kf = fk - MAGIC;
  double r = fma(kf, kExpR[2], -x);
  r = fma(kf, kExpR[3], r);
  double T;
  asm("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(tab_addr + ((k & (kExpTabSize - 1)) << 3)));
  n_out = k >> kExpTabBits;
  double q = fma(kExpC[0], r, kExpC[1]);
  q = fma(q, r, kExpC[2]);
  q = fma(q, r, 1.0);
  p_out = fma(T * r, q, T);
}

// delta-f polynomial without feqbar: s = bilinear part (mT^2 Qyy + pT^2 Qpp + K0 m^2 + mT pT pair), x = u.p / T
template <int MODEL>
__device__ __forceinline__ double df_poly(double s, double x, double K2)
{
  if (MODEL == M_LIN14) return fma(K2 * x, x, s);                 // + Pi bulk2 (u.p)^2
  return fma(s, rcp_fast(x), K2 * x);                              // [..] / (u.p) + (..) (u.p)
}

// One evaluation decided from x itself, for members near the overflow / sub-normal boundary of e^{-x} (and for cells whose pT B
// leaves the range of the factored form): returns p.dsigma f, exactly 0 where the reference's exp(x) overflows.  Out of line: rare.
template <int MODEL>
__device__ __noinline__ double late_member(double x, double s, double pv, double K2, double K31, double sign, int reg_lo, int reg_hi, int thr_hi)
{
  if (!exp_finite(x) || __double2hiint(pv) <= thr_hi) return 0.0;
  const double av = exp_neg(x);
  const double fb = rcp_fast(fma(sign, av, 1.0));
  double g = 1.0;
  if (MODEL != M_IDEAL) g = clamp_g(fma(fb, df_poly<MODEL>(s, x, K2), K31), reg_lo, reg_hi, reg_hi == 0x40000000 ? 0x3fffffffu : 0xffffffffu);
  return pv * ((av * fb) * g);
}

}  // namespace

template <int MODEL, int NYT, int NPT, int MINB>
__global__ void __launch_bounds__(kFW * 32, MINB)
cf_factored_kernel(const HotParams hp)
{
  constexpr bool POLY = (MODEL != M_IDEAL);
  constexpr int PP = PairPitch<NPT>::v;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Layout &L = hp.L;
  const int W = hp.n_warps;                                     // warps = phi tiles of this block
  const FMap<NYT, NPT> M(W);
  const int nthreads = blockDim.x;
  double *stage_base = reinterpret_cast<double *>(smem_raw);
  double *der_base = stage_base + (size_t)kStagesF * M.stage_doubles;
  uint64_t *full = reinterpret_cast<uint64_t *>(der_base + 2 * (size_t)M.der_doubles);

  // ---- task decode: blockIdx -> (pT point, phi block, y tile, species group, cell chunk); pT runs fastest, so the blocks that
  //      are resident together stream the same cells
  const int n_pblocks = (L.n_ptiles + W - 1) / W;
  const int per_chunk = hp.n_groupblocks * L.n_ytiles;          // n_groupblocks = species groups x pT points x phi blocks
  const int chunk = blockIdx.x / per_chunk;
  int bt = blockIdx.x - chunk * per_chunk;
  const int ipT = bt % L.n_pT; bt /= L.n_pT;
  const int pb = bt % n_pblocks; bt /= n_pblocks;
  const int ty = bt % L.n_ytiles; bt /= L.n_ytiles;
  const int sg = bt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tp = pb * W + warp;                                 // this warp's phi tile
  const bool tile_valid = tp < L.n_ptiles;
  const int n_ptiles_blk = min(W, L.n_ptiles - pb * W);         // phi tiles that exist in this block

  // ---- this lane's species; the block's pT
  const int isp = sg * 32 + lane;
  const bool lane_valid = isp < L.n_species;
  const int ipart = lane_valid ? isp : L.n_species - 1;
  const double mass = hp.mass[ipart], sign = hp.sign[ipart], pT = hp.pT[ipT];
  const double nsign = -sign;
  const double m2 = mass * mass, pT2 = pT * pT;
  const double mT2 = m2 + pT2;
  const double mT = sqrt(mT2);
  const int reg_lo = hp.reg_lo, reg_hi = hp.reg_hi;
  const uint32_t tab_addr = smem_u32(g_exp_tab);
  const int thr_hi = (int)(hp.outflow_thr >> 32);

  // cell tiles of this chunk, in units of kCTF cells (chunks are defined on Layout::ct tiles; ct is a multiple of kCTF)
  const int sub = L.ct / kCTF;
  const int64_t t_begin = (hp.chunk_tiles ? hp.chunk_tiles[chunk] : (L.n_tiles * (int64_t)chunk) / hp.n_chunks) * sub;
  const int64_t t_end = (hp.chunk_tiles ? hp.chunk_tiles[chunk + 1] : (L.n_tiles * (int64_t)(chunk + 1)) / hp.n_chunks) * sub;
  const int n_my_tiles = (int)(t_end - t_begin);

  const double *Yg = hp.Y + ((int64_t)ty * L.n_cells_pad) * NYT * kRec;
  const double *Sg = hp.S;
  const uint32_t stage_bytes = (uint32_t)(M.y_doubles + n_ptiles_blk * M.p_doubles + kCTF * kScal) * 8u;

  exp_table_init();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStagesF; s++) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int t_local) {
    const int st = t_local % kStagesF;
    const int64_t cell = (t_begin + t_local) * kCTF;
    double *dst = stage_base + (size_t)st * M.stage_doubles;
    mbar_arrive_expect_tx(&full[st], stage_bytes);
    bulk_g2s(dst, Yg + cell * NYT * kRec, (uint32_t)M.y_doubles * 8u, &full[st]);
    for (int w = 0; w < n_ptiles_blk; w++)
      bulk_g2s(dst + M.y_doubles + w * M.p_doubles, hp.P + (((int64_t)(pb * W + w) * L.n_cells_pad) + cell) * NPT * kRec,
               (uint32_t)M.p_doubles * 8u, &full[st]);
    bulk_g2s(dst + M.y_doubles + W * M.p_doubles, Sg + cell * kScal, (uint32_t)(kCTF * kScal) * 8u, &full[st]);
  };
  if (threadIdx.x == 0)
    for (int t = 0; t < kStagesF && t < n_my_tiles; t++) issue(t);

  double acc[NYT * NPT];
#pragma unroll
  for (int i = 0; i < NYT * NPT; i++) acc[i] = 0.0;

  for (int t = 0; t < n_my_tiles; t++) {
    const int st = t % kStagesF;
    mbar_wait(&full[st], (uint32_t)((t / kStagesF) & 1));
    const double *Ys = stage_base + (size_t)st * M.stage_doubles;
    const double *Ps = Ys + M.y_doubles;                          // [W][kCTF][NPT][kRec]
    const double *Ss = Ps + W * M.p_doubles;
    double *der = der_base + (size_t)(t & 1) * M.der_doubles;
    double2 *dE = reinterpret_cast<double2 *>(der + M.oE);       // [kCTF][NYT][32]
    int *dN = reinterpret_cast<int *>(der + M.oN);               // [kCTF][NYT][32]
    int2 *dCLS = reinterpret_cast<int2 *>(der + M.oCLS);         // [kCTF][NYT]
    double2 *dS = reinterpret_cast<double2 *>(der + M.oDS);      // [kCTF][NYT]
    double *dP = der + M.oDP;                                     // [W][kCTF][NPT][kDerP]
    double *dX = der + M.oDX;                                     // [W][kCTF][NYT][PP]
    double *dC = der + M.oDC;                                     // [kCTF][kDerC]

    // ---- tables of this tile, part 1: one warp per cell -- the slot factor e^{-mT A} of every lane, and its exponent range
#pragma unroll 1
    for (int c = warp; c < kCTF; c += W) {
      // largest pT B over the block's phi points -> a slot is dead when mT A > ln(DBL_MAX) + max pT B (high-word compare)
      double b = 0.0; int thi = (int)0x80000000; bool big = false;
      if (lane < n_ptiles_blk * NPT) {
        b = Ps[((lane / NPT) * kCTF + c) * NPT * kRec + (lane % NPT) * kRec];
        thi = __double2hiint(fma(pT, b, 709.79));
        big = !(fabs(b) * pT <= kQLimit);
      }
      const bool general = __any_sync(kFull, big);
      const int dead_hi = general ? 0x7fffffff : __reduce_max_sync(kFull, thi);
      bool live = false;
      if (lane < NYT) live = Ys[(c * NYT + lane) * kRec] < 0.5 * kDeadSlotA;       // not a padding / skipped / other-set record
      const bool cell_live = __any_sync(kFull, live);
      if (lane == 0) {
        dC[c * kDerC + 0] = Ss[c * kScal + 0];
        dC[c * kDerC + 1] = Ss[c * kScal + 1];
        dC[c * kDerC + 2] = 1.0 + Ss[c * kScal + 2];
        *reinterpret_cast<int2 *>(dC + c * kDerC + 3) = make_int2(general ? 1 : 0, cell_live ? 0 : 1);
      }
      if (lane < NYT) dS[c * NYT + lane] = make_double2(Ys[(c * NYT + lane) * kRec + 1], Ys[(c * NYT + lane) * kRec + 2]);
#pragma unroll 1
      for (int j = 0; j < NYT; j++) {
        const double a = mT * Ys[(c * NYT + j) * kRec];
        const bool pre_dead = __double2hiint(a) > dead_hi || !cell_live;
        if (__all_sync(kFull, pre_dead)) {
          if (lane == 0) dCLS[c * NYT + j] = make_int2(kNForcedDead, kNForcedDead);
          continue;
        }
        double pe; int ne;
        exp_neg_poly_at(a, tab_addr, pe, ne);                       // garbage for pre_dead lanes, which are forced dead
        if (pre_dead) ne = kNForcedDead;
        const int nemin = __reduce_min_sync(kFull, ne);
        const int nemax = __reduce_max_sync(kFull, ne);
        dE[(c * NYT + j) * 32 + lane] = make_double2(a, pe);
        dN[(c * NYT + j) * 32 + lane] = ne;
        if (lane == 0) dCLS[c * NYT + j] = make_int2(nemin, nemax);
      }
    }
    // ---- part 2: the phi factor e^{+pT B} and the pT-folded phi records; the mixed shear term
#pragma unroll 1
    for (int w = threadIdx.x; w < n_ptiles_blk * kCTF * NPT; w += nthreads) {
      const double *pr = Ps + w * kRec;                              // [w'][c][k] is contiguous
      const double q = pT * pr[0];
      double fq; int fm;
      exp_neg_poly_at(-q, tab_addr, fq, fm);                        // e^{+q} = fq 2^fm
      dP[w * kDerP + 0] = q;
      dP[w * kDerP + 1] = pT * pr[1];
      dP[w * kDerP + 2] = pT2 * pr[2];
      dP[w * kDerP + 3] = fq;
      *reinterpret_cast<int2 *>(dP + w * kDerP + 4) = make_int2(fm, 0);
      dP[w * kDerP + 5] = 0.0;
    }
    if (POLY) {
#pragma unroll 1
      for (int w = threadIdx.x; w < n_ptiles_blk * kCTF * NYT * NPT; w += nthreads) {
        const int wc = w / (NYT * NPT), r = w - wc * (NYT * NPT), j = r / NPT, k = r - j * NPT, c = wc % kCTF;
        const double *yr = Ys + (c * NYT + j) * kRec, *pr = Ps + (wc * NPT + k) * kRec;
        dX[(wc * NYT + j) * PP + k] = pT * fma(pr[4], yr[4], -(pr[3] * yr[3]));     // pT (R2 U2 - R1 U1)
      }
    }
    __syncthreads();                                   // tables of tile t complete; every warp is past the inner loop of tile t - 1
    if (threadIdx.x == 0 && t + kStagesF < n_my_tiles) issue(t + kStagesF);   // stage st has been consumed by the table pass

    if (tile_valid) {
#pragma unroll 1
      for (int c = 0; c < kCTF; c++) {
        const int2 flg = *reinterpret_cast<const int2 *>(dC + c * kDerC + 3);
        if (flg.y != 0) continue;                        // nothing alive in this cell (block-uniform)
        const bool cell_general = flg.x != 0;
        const double2 k01 = *reinterpret_cast<const double2 *>(dC + c * kDerC);
        const double K0m = k01.x * m2;
        const double K2 = k01.y;
        const double K31 = dC[c * kDerC + 2];

        double q[NPT], pd[NPT], G0[NPT], fq[NPT]; int fm[NPT];
#pragma unroll
        for (int k = 0; k < NPT; k++) {
          const double *pr = dP + ((warp * kCTF + c) * NPT + k) * kDerP;
          const double2 v0 = *reinterpret_cast<const double2 *>(pr);
          const double2 v1 = *reinterpret_cast<const double2 *>(pr + 2);
          q[k] = v0.x; pd[k] = v0.y; G0[k] = v1.x; fq[k] = v1.y;
          fm[k] = *reinterpret_cast<const int *>(pr + 4);
        }
        int fmmin = fm[0], fmmax = fm[0];
#pragma unroll
        for (int k = 1; k < NPT; k++) { fmmin = min(fmmin, fm[k]); fmmax = max(fmmax, fm[k]); }

#pragma unroll
        for (int j = 0; j < NYT; j++) {
          // exponent range of the group over all lanes: warp-uniform
          const int2 cls = dCLS[c * NYT + j];
          const int nhi = cls.y + fmmax;
          if (nhi <= kNDead && !cell_general) continue;  // every member of every lane dead: exact 0
          const bool all_ok = cls.x + fmmin >= kNNormal && !cell_general;      // every member of every lane alive and normal
          double *accj = acc + j * NPT;
          const double *xr = dX + ((warp * kCTF + c) * NYT + j) * PP;
          const double2 ea = dE[(c * NYT + j) * 32 + lane];                  // mT A, mantissa of e^{-mT A}
          const int ne = dN[(c * NYT + j) * 32 + lane];
          const double2 cq = dS[c * NYT + j];                                 // Cp, Qyy
          const double a = ea.x, pe = ea.y;
          const double H = POLY ? fma(mT2, cq.y, K0m) : 0.0;                  // mT^2 Qyy + K0 m^2
          double pv[NPT]; int n[NPT];
#pragma unroll
          for (int k = 0; k < NPT; k++) { pv[k] = fma(mT, cq.x, pd[k]); n[k] = ne + fm[k]; }     // p.dsigma; binary exponent of e^{-x}
          if (!all_ok) {
            // rare (warp-uniform, ~7 % of the live groups): the group touches the overflow / sub-normal boundary of e^{-x} on some
            // lane, or the cell is out of the factored form's range: every member that is not plainly dead is evaluated from x itself
#pragma unroll
            for (int k = 0; k < NPT; k++) {
              if (n[k] > kNDead || cell_general) {
                const double sk = POLY ? fma(mT, xr[k], H + G0[k]) : 0.0;
                accj[k] += late_member<MODEL>(a - q[k], sk, pv[k], K2, K31, sign, reg_lo, reg_hi, thr_hi);
              }
            }
            continue;
          }
          // ---- fast path, branch-free but for the occupation-factor form (warp-uniform)
          double av[NPT], g[NPT], fe[NPT];
#pragma unroll
          for (int k = 0; k < NPT; k++) {
            const double pm = pe * fq[k];
            av[k] = __hiloint2double(__double2hiint(pm) + (n[k] << 20), __double2loint(pm));
          }
          if (POLY) {
#pragma unroll
            for (int k = 0; k < NPT; k++) g[k] = df_poly<MODEL>(fma(mT, xr[k], H + G0[k]), a - q[k], K2);
          }
          if (nhi <= kNDilute) {                         // a < 2^-18 for every lane: 1 / (1 + Theta a) = 1 - Theta a + a^2 to 2^-54
#pragma unroll
            for (int k = 0; k < NPT; k++) {
              const double fb = fma(av[k], av[k], fma(nsign, av[k], 1.0));
              fe[k] = av[k] * fb; if (POLY) g[k] = fma(fb, g[k], K31);
            }
          } else {
#pragma unroll
            for (int k = 0; k < NPT; k++) {
              const double fb = rcp_fast(fma(sign, av[k], 1.0));
              fe[k] = av[k] * fb; if (POLY) g[k] = fma(fb, g[k], K31);
            }
          }
#pragma unroll
          for (int k = 0; k < NPT; k++) {
            if (POLY) fe[k] *= clamp_g(g[k], reg_lo, reg_hi, hp.reg_chk);               // regulate_deltaf
            if (__double2hiint(pv[k]) > thr_hi) accj[k] = fma(pv[k], fe[k], accj[k]);   // outflow test (smooth_kernels.cpp:285)
          }
        }
      }
    }
  }

  if (tile_valid) hot_epilogue<NYT, NPT, false>(hp, acc, nullptr, chunk, 0, ty, tp, lane_valid, ipart, ipT);
}

// ------------------------------------------------------------------------------------------------ dispatch
// Register-tile shapes of the factored kernel; is3d_options.tile_variant = 17 + k selects entry k.
struct FShape { int nyt, npt, ct, minb; };
static const FShape kFShapes[kNumFactoredVariants] = {{7, 3, 8, 3}, {7, 4, 8, 3}, {3, 6, 8, 3}, {3, 6, 8, 4}, {3, 8, 8, 3}};

bool factored_supported(int model, const Layout &L)
{
  return !L.dim2 && L.n_species >= kFactoredMinSpecies &&
         (model == M_LIN14 || model == M_LINCE || model == M_JONAHLIN || model == M_IDEAL);
}

int factored_match(int nyt, int npt)
{
  for (int v = 0; v < kNumFactoredVariants; v++)
    if (kFShapes[v].nyt == nyt && kFShapes[v].npt == npt) return v;
  return -1;
}

void factored_variant_shape(int fvariant, int *nyt, int *npt, int *ct, int *max_warps)
{
  if (fvariant < 0 || fvariant >= kNumFactoredVariants) fvariant = 0;
  const FShape &s = kFShapes[fvariant];
  *nyt = s.nyt; *npt = s.npt; *ct = s.ct; *max_warps = kFW;
}

// lanes are species, warps are phi tiles: blocks per (cell chunk, y tile) = species groups x pT points x phi blocks
void factored_blocking(int n_species, int n_pT, int n_ptiles, int *n_warps, int *n_groupblocks)
{
  const int w = n_ptiles < kFW ? n_ptiles : kFW;
  *n_warps = w;
  *n_groupblocks = ((n_species + 31) / 32) * n_pT * ((n_ptiles + w - 1) / w);
}

template <int MODEL, int NYT, int NPT, int MINB>
static cudaError_t launch_f(const HotParams &hp, cudaStream_t st, size_t *smem_out)
{
  const Layout &L = hp.L;
  const FMap<NYT, NPT> M(hp.n_warps);
  const size_t smem = M.bytes();
  if (smem_out) *smem_out = smem;
  if (hp.integ_mode) return cudaErrorInvalidValue;       // operation = 0 integrates over the pT lanes of a block: cf_kernels.cu
  if (hp.n_warps < 1 || hp.n_warps > kFW || L.ct % kCTF != 0) return cudaErrorInvalidValue;
  auto kern = cf_factored_kernel<MODEL, NYT, NPT, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t grid = (int64_t)hp.n_groupblocks * L.n_ytiles * hp.n_chunks;
  if (grid == 0) return cudaSuccess;
  kern<<<(unsigned)grid, hp.n_warps * 32, smem, st>>>(hp);
  return cudaGetLastError();
}

template <int MODEL>
static cudaError_t launch_fmodel(const HotParams &hp, int fvariant, cudaStream_t st, size_t *smem_out)
{
  switch (fvariant) {
    case 1: return launch_f<MODEL, 7, 4, 3>(hp, st, smem_out);
    case 2: return launch_f<MODEL, 3, 6, 3>(hp, st, smem_out);
    case 3: return launch_f<MODEL, 3, 6, 4>(hp, st, smem_out);
    case 4: return launch_f<MODEL, 3, 8, 3>(hp, st, smem_out);
    default: return launch_f<MODEL, 7, 3, 3>(hp, st, smem_out);
  }
}

cudaError_t launch_factored(int model, const HotParams &hp, int fvariant, cudaStream_t st, size_t *smem_out)
{
  switch (model) {
    case M_LIN14: return launch_fmodel<M_LIN14>(hp, fvariant, st, smem_out);
    case M_LINCE: return launch_fmodel<M_LINCE>(hp, fvariant, st, smem_out);
    case M_JONAHLIN: return launch_fmodel<M_JONAHLIN>(hp, fvariant, st, smem_out);
    case M_IDEAL: return launch_fmodel<M_IDEAL>(hp, fvariant, st, smem_out);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace is3d
