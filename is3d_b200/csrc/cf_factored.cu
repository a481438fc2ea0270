// cf_factored.cu -- second-generation hot kernel for the linear-delta-f models in 3+1D (sm_100a): the loop nest of
// EmissionFunctionArray::calculate_dN_pTdpTdphidy (reference src/cpp/emissionfunction_smooth_kernels.cpp:246-347) for
// df_mode 1 (14 moment), 2 (Chapman-Enskog), Jonah's linearised df and the ideal f_eq, for species lists of >= 16 hadrons.
//
// Work decomposition
//   lane   <-> one SPECIES; the 32 lanes of a warp are 32 consecutive species of the chosen list (consecutive entries of a PDG
//              list have similar masses) and every warp of a block works on the SAME pT point.  u.p/T = mT A[slot] - pT B[phi] is
//              then nearly the same for all lanes of a warp: dead / dilute / ultra-dilute groups are decided per WARP with
//              votes, and the one branch taken is the one all lanes need.  (cf_kernels.cu puts the 32 pT points of one species in
//              a warp: pT spans 0.001-40 GeV there, so a warp straddles every class.)
//   thread     register tile of NYT rapidity slots x NPT phi points, walks the cells of its chunk.
//   block      up to 4 warps (128 species) x one pT point x one (y-tile, phi-tile) x one contiguous cell chunk; cell tiles are
//              streamed global -> shared with cp.async.bulk (TMA) through a kStages-deep mbarrier pipeline.
//   grid       species blocks x pT points x bin tiles x cell chunks; partial[chunk][bin] + reduce_kernel as in cf_kernels.cu.
//
// Arithmetic per evaluation
//   1. Factored exponential: e^{-x} = e^{-mT A[slot]} e^{+pT B[phi]}.  The phi factor does not depend on the species any more: the
//      block computes it ONCE per (cell, phi) into shared memory while it turns the streamed records into the tables of the inner
//      loop (pT B, pT D, pT^2 Qpp, pT (R2 U2 - R1 U1), ...); a thread evaluates one exponential per (cell, slot) and multiplies
//      mantissas (1 DMUL) / adds binary exponents per evaluation.  The exponent sum also classifies the group, so nothing of this
//      needs x itself.
//   2. g = 1 + df is formed directly (the reference multiplies f_eq (1 + df), :330); regulate_deltaf clamps g to [0, 2] on its
//      high word.
//   3. Occupation factor 1 / (1 + Theta a), a = e^{-x}: exactly 1 for a < 2^-54 (ultra dilute), 1 - Theta a + a^2 for a < 2^-18,
//      MUFU seed + Newton otherwise -- chosen per warp.
//   4. The derived tables are double buffered: a tile costs ONE __syncthreads and its TMA stage is released before the inner loop.
// Dead slots (every member's exp(x) overflows in the reference, f = 0 exactly) are skipped before their exponential is evaluated:
// mT A > ln(DBL_MAX) + max_k pT B[k] is a high-word compare.  Members within ~3 units of the overflow / sub-normal boundary are
// evaluated by late_member() exactly as cf_kernels.cu decides them (x itself, exp_neg), so the zero pattern is the reference's.
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
constexpr int kNUltra = -56;       // n <= kNUltra: a < 2^-54, 1 + Theta a rounds to 1
constexpr int kNForcedDead = -200000;

constexpr int kDerS = 4;           // derived slot record: A, Cp, Qyy, -
constexpr int kDerP = 6;           // derived phi record (block's pT folded in): q = pT B, pd = pT D, G0 = pT^2 Qpp, fq, int fm, -
constexpr int kDerC = 6;           // derived cell record: K0, K2, 1 + K3, int2 {dead_hi, flags}, -, -
constexpr double kQLimit = 2000.0; // |pT B| beyond this leaves the exact range of the Cody-Waite reduction: late_member() for the cell
constexpr unsigned kFull = 0xffffffffu;

template <int NPT> struct PairPitch { static constexpr int v = (NPT + 1) & ~1; };

// clamp g = 1 + df to [0, 2] (regulate_deltaf: df in [-1, 1], smooth_kernels.cpp:328) on the high word: negative (sign bit) -> +0,
// >= 2 -> 2.  reg_lo / reg_hi are 0 / 0x40000000 with regulation on and INT_MIN / INT_MAX with it off.
__device__ __forceinline__ double clamp_g(double g, int reg_lo, int reg_hi)
{
  const int hi = __double2hiint(g);
  const int hc = min(max(hi, reg_lo), reg_hi);
  return __hiloint2double(hc, hc == hi ? __double2loint(g) : 0);
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
  if (MODEL != M_IDEAL) g = clamp_g(fma(fb, df_poly<MODEL>(s, x, K2), K31), reg_lo, reg_hi);
  return pv * ((av * fb) * g);
}

}  // namespace

template <int MODEL, int NYT, int NPT, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
cf_factored_kernel(const HotParams hp)
{
  constexpr bool POLY = (MODEL != M_IDEAL);
  constexpr int PP = PairPitch<NPT>::v;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Layout &L = hp.L;
  const int CT = L.ct;
  const int nthreads = blockDim.x;
  const int y_doubles = CT * NYT * kRec, p_doubles = CT * NPT * kRec, s_doubles = CT * kScal;
  const int stage_doubles = y_doubles + p_doubles + s_doubles;
  const int der_doubles = CT * (NYT * kDerS + NPT * kDerP + kDerC + NYT * PP);
  double *stage_base = reinterpret_cast<double *>(smem_raw);
  double *der_base = stage_base + (size_t)kStages * stage_doubles;
  uint64_t *full = reinterpret_cast<uint64_t *>(der_base + 2 * (size_t)der_doubles);

  // ---- task decode: blockIdx -> (species block, pT point, y tile, phi tile, cell chunk); pT runs fastest, so the blocks that are
  //      resident together stream the same cells
  const int n_bintiles = hp.n_groupblocks * L.n_ytiles * L.n_ptiles;
  const int chunk = blockIdx.x / n_bintiles;
  int bt = blockIdx.x - chunk * n_bintiles;
  const int ipT = bt % L.n_pT; bt /= L.n_pT;
  const int tp = bt % L.n_ptiles; bt /= L.n_ptiles;
  const int ty = bt % L.n_ytiles; bt /= L.n_ytiles;
  const int sb = bt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- this lane's species; the block's pT
  const int isp = (sb * hp.n_warps + warp) * 32 + lane;
  const bool lane_valid = isp < L.n_species;
  const int ipart = lane_valid ? isp : L.n_species - 1;
  const double mass = hp.mass[ipart], sign = hp.sign[ipart], pT = hp.pT[ipT];
  const double nsign = -sign;
  const double m2 = mass * mass, pT2 = pT * pT;
  const double mT2 = m2 + pT2;
  const double mT = sqrt(mT2);
  const int reg_lo = hp.reg_lo, reg_hi = hp.reg_hi;
  const int thr_hi = (int)(hp.outflow_thr >> 32);

  const int64_t t_begin = hp.chunk_tiles ? hp.chunk_tiles[chunk] : (L.n_tiles * (int64_t)chunk) / hp.n_chunks;
  const int64_t t_end = hp.chunk_tiles ? hp.chunk_tiles[chunk + 1] : (L.n_tiles * (int64_t)(chunk + 1)) / hp.n_chunks;
  const int n_my_tiles = (int)(t_end - t_begin);

  const double *Yg = hp.Y + ((int64_t)ty * L.n_cells_pad) * NYT * kRec;
  const double *Pg = hp.P + ((int64_t)tp * L.n_cells_pad) * NPT * kRec;
  const double *Sg = hp.S;
  const uint32_t stage_bytes = (uint32_t)stage_doubles * 8u;

  exp_table_init();
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
    bulk_g2s(dst, Yg + cell * NYT * kRec, (uint32_t)y_doubles * 8u, &full[st]);
    bulk_g2s(dst + y_doubles, Pg + cell * NPT * kRec, (uint32_t)p_doubles * 8u, &full[st]);
    bulk_g2s(dst + y_doubles + p_doubles, Sg + cell * kScal, (uint32_t)s_doubles * 8u, &full[st]);
  };
  if (threadIdx.x == 0)
    for (int t = 0; t < kStages && t < n_my_tiles; t++) issue(t);

  double acc[NYT * NPT];
#pragma unroll
  for (int i = 0; i < NYT * NPT; i++) acc[i] = 0.0;

  for (int t = 0; t < n_my_tiles; t++) {
    const int st = t % kStages;
    mbar_wait(&full[st], (uint32_t)((t / kStages) & 1));
    const double *Ys = stage_base + (size_t)st * stage_doubles;
    const double *Ps = Ys + y_doubles;
    const double *Ss = Ps + p_doubles;
    double *dS = der_base + (size_t)(t & 1) * der_doubles;        // [CT][NYT][kDerS]
    double *dP = dS + CT * NYT * kDerS;                           // [CT][NPT][kDerP]
    double *dC = dP + CT * NPT * kDerP;                           // [CT][kDerC]
    double *dX = dC + CT * kDerC;                                 // [CT][NYT][PP]: pT (R2 U2 - R1 U1)

    // ---- derive this tile's tables (block-cooperative, a few items per thread)
#pragma unroll 1
    for (int w = threadIdx.x; w < CT * NYT; w += nthreads) {
      const double *yr = Ys + w * kRec;
      dS[w * kDerS + 0] = yr[0];
      dS[w * kDerS + 1] = yr[1];
      dS[w * kDerS + 2] = yr[2];
      dS[w * kDerS + 3] = 0.0;
    }
#pragma unroll 1
    for (int w = threadIdx.x; w < CT * NPT; w += nthreads) {
      const double *pr = Ps + w * kRec;
      const double q = pT * pr[0];
      double fq; int fm;
      exp_neg_poly(-q, fq, fm);                                    // e^{+q} = fq 2^fm
      dP[w * kDerP + 0] = q;
      dP[w * kDerP + 1] = pT * pr[1];
      dP[w * kDerP + 2] = pT2 * pr[2];
      dP[w * kDerP + 3] = fq;
      *reinterpret_cast<int2 *>(dP + w * kDerP + 4) = make_int2(fm, 0);
      dP[w * kDerP + 5] = 0.0;
    }
    if (POLY) {
#pragma unroll 1
      for (int w = threadIdx.x; w < CT * NYT * NPT; w += nthreads) {
        const int c = w / (NYT * NPT), r = w - c * (NYT * NPT), j = r / NPT, k = r - j * NPT;
        const double *yr = Ys + (c * NYT + j) * kRec, *pr = Ps + (c * NPT + k) * kRec;
        dX[(c * NYT + j) * PP + k] = pT * fma(pr[4], yr[4], -(pr[3] * yr[3]));
      }
    }
#pragma unroll 1
    for (int c = threadIdx.x; c < CT; c += nthreads) {
      double bmax = Ps[c * NPT * kRec], babs = fabs(bmax);
      for (int k = 1; k < NPT; k++) { const double b = Ps[(c * NPT + k) * kRec]; bmax = fmax(bmax, b); babs = fmax(babs, fabs(b)); }
      bool live = false;                                 // any slot of this cell that is not a dead record (padding, skipped cell,
      for (int j = 0; j < NYT; j++) live = live || Ys[(c * NYT + j) * kRec] < 0.5 * kDeadSlotA;   // slot owned by the other record set)
      const bool general = !(babs * pT <= kQLimit);
      dC[c * kDerC + 0] = Ss[c * kScal + 0];
      dC[c * kDerC + 1] = Ss[c * kScal + 1];
      dC[c * kDerC + 2] = 1.0 + Ss[c * kScal + 2];
      // a slot is dead when mT A > ln(DBL_MAX) + max_k pT B[k]; compared on the high words, so the sliver in between stays alive
      const int dead_hi = !live ? (int)0x80000000 : general ? 0x7fffffff : __double2hiint(fma(pT, bmax, 709.79));
      *reinterpret_cast<int2 *>(dC + c * kDerC + 3) = make_int2(dead_hi, general ? 1 : 0);
      dC[c * kDerC + 4] = 0.0; dC[c * kDerC + 5] = 0.0;
    }
    __syncthreads();                                   // tables of tile t complete; every warp is past the inner loop of tile t - 1
    if (threadIdx.x == 0 && t + kStages < n_my_tiles) issue(t + kStages);     // stage st has been consumed by the derive pass

    for (int c = 0; c < CT; c++) {
      const int2 flg = *reinterpret_cast<const int2 *>(dC + c * kDerC + 3);
      const int dead_hi = flg.x;
      if (dead_hi == (int)0x80000000) continue;          // nothing alive in this cell (block-uniform)
      const bool cell_general = flg.y != 0;
      const double2 k01 = *reinterpret_cast<const double2 *>(dC + c * kDerC);
      const double K0m = k01.x * m2;
      const double K2 = k01.y;
      const double K31 = dC[c * kDerC + 2];

      double q[NPT], pd[NPT], G0[NPT], fq[NPT]; int fm[NPT];
#pragma unroll
      for (int k = 0; k < NPT; k++) {
        const double2 v0 = *reinterpret_cast<const double2 *>(dP + (c * NPT + k) * kDerP);
        const double2 v1 = *reinterpret_cast<const double2 *>(dP + (c * NPT + k) * kDerP + 2);
        q[k] = v0.x; pd[k] = v0.y; G0[k] = v1.x; fq[k] = v1.y;
        fm[k] = *reinterpret_cast<const int *>(dP + (c * NPT + k) * kDerP + 4);
      }

#pragma unroll
      for (int j = 0; j < NYT; j++) {
        const double2 s0 = *reinterpret_cast<const double2 *>(dS + (c * NYT + j) * kDerS);
        const double a = mT * s0.x;                     // mT A: the slot part of u.p / T
        const bool pre_dead = __double2hiint(a) > dead_hi;
        if (__all_sync(kFull, pre_dead)) continue;      // every member of every lane dead: exact 0
        double *accj = acc + j * NPT;
        const double *xr = dX + (c * NYT + j) * PP;
        double pe; int ne;
        exp_neg_poly(a, pe, ne);                        // e^{-a} = pe 2^ne (garbage for pre_dead lanes, which are forced dead)
        if (pre_dead) ne = kNForcedDead;
        int n[NPT], nmin, nmax;
#pragma unroll
        for (int k = 0; k < NPT; k++) n[k] = ne + fm[k];
        nmin = n[0]; nmax = n[0];
#pragma unroll
        for (int k = 1; k < NPT; k++) { nmin = min(nmin, n[k]); nmax = max(nmax, n[k]); }
        const bool dead_t = (nmax <= kNDead && !cell_general) || pre_dead;
        const bool ok_t = nmin >= kNNormal && !cell_general;           // every member alive with a normal e^{-x}: fast path
        const bool late_t = !dead_t && !ok_t;                           // some member dead / sub-normal / out of range
        const bool dil_w = __all_sync(kFull, !ok_t || nmax <= kNDilute);
        const bool ult_w = __all_sync(kFull, !ok_t || nmax <= kNUltra);
        const double cpm = mT * s0.y;                   // mT (cosh dsigma_tau + sinh dsigma_eta / tau)
        const double H = POLY ? fma(mT2, dS[(c * NYT + j) * kDerS + 2], K0m) : 0.0;     // mT^2 Qyy + K0 m^2
        double pv[NPT];
#pragma unroll
        for (int k = 0; k < NPT; k++) pv[k] = pd[k] + cpm;              // p.dsigma
        int pvlo = __double2hiint(pv[0]);
#pragma unroll
        for (int k = 1; k < NPT; k++) pvlo = min(pvlo, __double2hiint(pv[k]));
        const bool pos_w = __all_sync(kFull, !ok_t || pvlo > thr_hi);   // outflow test passes for every member of every fast lane

        if (ok_t) {
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
          if (ult_w) {                                  // feqbar = 1 exactly
#pragma unroll
            for (int k = 0; k < NPT; k++) { fe[k] = av[k]; if (POLY) g[k] = g[k] + K31; }
          } else if (dil_w) {
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
          if (POLY) {
#pragma unroll
            for (int k = 0; k < NPT; k++) fe[k] *= clamp_g(g[k], reg_lo, reg_hi);
          }
          if (pos_w) {
#pragma unroll
            for (int k = 0; k < NPT; k++) accj[k] = fma(pv[k], fe[k], accj[k]);
          } else {
#pragma unroll
            for (int k = 0; k < NPT; k++) accumulate_pos(accj[k], pv[k], fe[k], thr_hi);
          }
        }
        if (late_t) {
#pragma unroll
          for (int k = 0; k < NPT; k++) {
            const double sk = POLY ? fma(mT, xr[k], H + G0[k]) : 0.0;
            accj[k] += late_member<MODEL>(a - q[k], sk, pv[k], K2, K31, sign, reg_lo, reg_hi, thr_hi);
          }
        }
      }
    }
  }

  __syncthreads();                                     // the stage area becomes the epilogue's scratch
  hot_epilogue<NYT, NPT, false>(hp, acc, stage_base, chunk, sb, ty, tp, lane_valid, ipart, ipT);
}

// ------------------------------------------------------------------------------------------------ dispatch
// Register-tile shapes of the factored kernel; is3d_options.tile_variant = 17 + k selects entry k.
struct FShape { int nyt, npt, ct, minb, warps; };
static const FShape kFShapes[kNumFactoredVariants] = {{7, 3, 8, 3, 4}, {7, 4, 8, 3, 4}, {7, 3, 8, 6, 2}, {7, 6, 8, 2, 4}};

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
  *nyt = s.nyt; *npt = s.npt; *ct = s.ct; *max_warps = s.warps;
}

// lanes are species: blocks of n_warps x 32 species, one block column per pT point
void factored_blocking(int fvariant, int n_species, int n_pT, int *n_warps, int *n_groupblocks)
{
  if (fvariant < 0 || fvariant >= kNumFactoredVariants) fvariant = 0;
  const int max_warps = kFShapes[fvariant].warps;
  const int n_groups = (n_species + 31) / 32;
  int best_w = 1, best = 1 << 30;
  for (int w = max_warps; w >= 1; w--) {                 // widest block that launches the fewest warps
    const int launched = ((n_groups + w - 1) / w) * w;
    if (launched < best) { best = launched; best_w = w; }
  }
  *n_warps = best_w;
  *n_groupblocks = ((n_groups + best_w - 1) / best_w) * n_pT;
}

template <int MODEL, int NYT, int NPT, int WARPS, int MINB>
static cudaError_t launch_f(const HotParams &hp, cudaStream_t st, size_t *smem_out)
{
  const Layout &L = hp.L;
  constexpr int PP = PairPitch<NPT>::v;
  const size_t stage_doubles = (size_t)L.ct * (NYT * kRec + NPT * kRec + kScal);
  const size_t der_doubles = (size_t)L.ct * (NYT * kDerS + NPT * kDerP + kDerC + NYT * PP);
  const size_t smem = (kStages * stage_doubles + 2 * der_doubles) * 8 + kStages * sizeof(uint64_t);
  if (smem_out) *smem_out = smem;
  if (hp.integ_mode) return cudaErrorInvalidValue;       // operation = 0 integrates over the pT lanes of a block: cf_kernels.cu
  if (hp.n_warps > WARPS) return cudaErrorInvalidValue;
  auto kern = cf_factored_kernel<MODEL, NYT, NPT, WARPS, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t grid = (int64_t)hp.n_groupblocks * L.n_ytiles * L.n_ptiles * hp.n_chunks;
  if (grid == 0) return cudaSuccess;
  kern<<<(unsigned)grid, hp.n_warps * 32, smem, st>>>(hp);
  return cudaGetLastError();
}

template <int MODEL>
static cudaError_t launch_fmodel(const HotParams &hp, int fvariant, cudaStream_t st, size_t *smem_out)
{
  switch (fvariant) {
    case 1: return launch_f<MODEL, 7, 4, 4, 3>(hp, st, smem_out);
    case 2: return launch_f<MODEL, 7, 3, 2, 6>(hp, st, smem_out);
    case 3: return launch_f<MODEL, 7, 6, 4, 2>(hp, st, smem_out);
    default: return launch_f<MODEL, 7, 3, 4, 3>(hp, st, smem_out);
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
