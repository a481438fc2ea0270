// cf_factored.cu -- second-generation hot kernel for the linear-delta-f models in 3+1D (sm_100a): the loop nest of
// EmissionFunctionArray::calculate_dN_pTdpTdphidy (reference src/cpp/emissionfunction_smooth_kernels.cpp:246-347) for
// df_mode 1 (14 moment), 2 (Chapman-Enskog), Jonah's linearised df and the ideal f_eq.
//
// Same work decomposition and TMA cell stream as cf_kernels.cu (lane = (species, pT), thread = NYT x NPT register tile, block =
// bin tile x cell chunk).  What is new is the arithmetic per evaluation:
//
//  1. Factored exponential.  u.p/T = x = mT A[slot] - pT B[phi], so e^{-x} = e^{-mT A[slot]} e^{+pT B[phi]}: one exponential per
//     (cell, slot) and one per (cell, phi) instead of one per evaluation, each kept as mantissa x 2^n (exp_neg_poly).  An
//     evaluation multiplies two mantissas (1 DMUL) and adds two integers; the exponent n = n_slot + n_phi also classifies the
//     group (dead / sub-normal / dilute / ultra-dilute) with integer min/max, so x itself is never formed on the fast path.
//  2. Merged bilinear delta-f (14 moment).  df / feqbar = mT^2 Qyy + pT^2 Qpp + K0 m^2 + mT pT pair + K2 x^2, and
//     K2 x^2 = mT^2 K2 A^2 + pT^2 K2 B^2 - 2 mT pT K2 A B has the same three shapes: it is folded into Qyy, Qpp and the pair
//     table by the block once per (cell, slot / phi / pair) -- 2 FP64 instructions per evaluation instead of 4.
//  3. g = 1 + df is formed directly (the reference multiplies f_eq (1 + df), :330) and regulate_deltaf clamps g to [0, 2] on its
//     high word; one unsigned max3 per group decides whether any member needs the clamp.
//  4. Occupation factor 1 / (1 + Theta a), a = e^{-x}: exactly 1 for a < 2^-54 ("ultra dilute", 6 FP64 instructions per
//     evaluation), 1 - Theta a + a^2 for a < 2^-18 (10), MUFU seed + Newton otherwise (12).
//  5. Derived per-tile tables.  Right after a cell tile lands, the block turns the streamed records into the tables the inner
//     loop reads (slot: A, Cp, Qyy'; phi: B, D, Qpp'; pair'; cell: K0, K2, K3, max B) -- double buffered, so a tile costs ONE
//     __syncthreads and its TMA stage is released before the inner loop starts.
//
// Dead groups (every member's exp(x) overflows in the reference, f = 0 exactly) are skipped before their exponential is
// evaluated: mT A > ln(DBL_MAX) + pT max_k B[k] is a high-word compare.  Groups with a member near the overflow / sub-normal
// boundary take the exact per-member path of cf_device.cuh (exp_neg_slow), so the zero pattern is the reference's.
#include "cf_internal.h"
#include <algorithm>
#include "cf_device.cuh"
#include "cf_epilogue.cuh"

namespace is3d {

namespace {

// exponent classes of a = mantissa (< 4) x 2^n
constexpr int kNDead = -1031;      // n <= kNDead for every member: a < 2^-1029, x > 713 > ln(DBL_MAX): all terms exactly 0
constexpr int kNNormal = -1019;    // n >= kNNormal for every member: normal results, exponent insertion is safe
constexpr int kNDilute = -20;      // n <= kNDilute: a < 2^-18
constexpr int kNUltra = -56;       // n <= kNUltra: a < 2^-54, 1 + Theta a rounds to 1

constexpr int kDerC = 6;           // doubles per derived cell record: K0, K2, 1 + K3, max_k B, int2 {general-path flag, dead-cell flag}, -
constexpr double kQLimit = 2000.0; // |pT B| beyond this: the Cody-Waite reduction of cf_device.cuh leaves its exact range -> general path

template <int NPT> struct PairPitch { static constexpr int v = (NPT + 1) & ~1; };

// clamp g = 1 + df to [0, 2] (regulate_deltaf: df in [-1, 1], smooth_kernels.cpp:328) on the high word: negative (sign bit) -> +0,
// >= 2 -> 2.  reg_lo / reg_hi are 0 / 0x40000000 with regulation on and INT_MIN / INT_MAX with it off.
__device__ __forceinline__ double clamp_g(double g, int reg_lo, int reg_hi)
{
  const int hi = __double2hiint(g);
  const int hc = min(max(hi, reg_lo), reg_hi);
  return __hiloint2double(hc, hc == hi ? __double2loint(g) : 0);
}

// acc[k] += (pd[k] + cpm) fe[k] g[k] for the members whose p.dsigma passes the outflow test (smooth_kernels.cpp:285); the test is
// made once per group on the high words (thr_hi = 0: p.dsigma > 0; INT_MIN: outflow off), so the common all-pass case is N plain DFMAs
template <int N>
__device__ __forceinline__ void accumulate_group(double *acc, const double (&pd)[N], double cpm, const double (&fe)[N], const double *g, int thr_hi)
{
  double pv[N];
#pragma unroll
  for (int k = 0; k < N; k++) pv[k] = pd[k] + cpm;
  int lo = __double2hiint(pv[0]);
#pragma unroll
  for (int k = 1; k < N; k++) lo = min(lo, __double2hiint(pv[k]));
  if (lo > thr_hi) {
#pragma unroll
    for (int k = 0; k < N; k++) acc[k] = fma(pv[k], g ? fe[k] * g[k] : fe[k], acc[k]);
  } else {
#pragma unroll
    for (int k = 0; k < N; k++) accumulate_pos(acc[k], pv[k], g ? fe[k] * g[k] : fe[k], thr_hi);
  }
}

}  // namespace

template <int MODEL, int NYT, int NPT, int MINB>
__global__ void __launch_bounds__(kMaxWarps * 32, MINB)
cf_factored_kernel(const HotParams hp)
{
  constexpr bool MERGE = (MODEL == M_LIN14);                    // K2 x^2 folded into the bilinear form
  constexpr bool NEEDX = (MODEL == M_LINCE || MODEL == M_JONAHLIN);   // df contains 1 / (u.p): x is formed per evaluation
  constexpr int PP = PairPitch<NPT>::v;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Layout &L = hp.L;
  const int CT = L.ct;
  const int nthreads = blockDim.x;
  const int y_doubles = CT * NYT * kRec, p_doubles = CT * NPT * kRec, s_doubles = CT * kScal;
  const int stage_doubles = y_doubles + p_doubles + s_doubles;
  const int der_doubles = CT * (NYT * 4 + NPT * 4 + kDerC + NYT * PP);
  double *stage_base = reinterpret_cast<double *>(smem_raw);
  double *der_base = stage_base + (size_t)kStages * stage_doubles;
  uint64_t *full = reinterpret_cast<uint64_t *>(der_base + 2 * (size_t)der_doubles);

  // ---- task decode: blockIdx -> (group block, y tile, phi tile, cell chunk)
  const int n_bintiles = hp.n_groupblocks * L.n_ytiles * L.n_ptiles;
  const int chunk = blockIdx.x / n_bintiles;
  int bt = blockIdx.x - chunk * n_bintiles;
  const int tp = bt % L.n_ptiles; bt /= L.n_ptiles;
  const int ty = bt % L.n_ytiles; bt /= L.n_ytiles;
  const int gb = bt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- this lane's (species, pT)
  const int idx = (gb * hp.n_warps + warp) * 32 + lane;
  const bool lane_valid = idx < L.n_species * L.n_pT;
  const int ipart = lane_valid ? idx / L.n_pT : 0;
  const int ipT = lane_valid ? idx - ipart * L.n_pT : 0;
  const double mass = hp.mass[ipart], sign = hp.sign[ipart], pT = hp.pT[ipT];
  const double nsign = -sign;
  const double m2 = mass * mass, pT2 = pT * pT;
  const double mT2 = m2 + pT2;
  const double mT = sqrt(mT2);
  const double mTpT = mT * pT;
  // regulate_deltaf as integer bounds on the high word of g = 1 + df (set by the host, see clamp_g)
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
    double *dS = der_base + (size_t)(t & 1) * der_doubles;        // [CT][NYT][4]: A, Cp, Qyy', -
    double *dP = dS + CT * NYT * 4;                               // [CT][NPT][4]: B, D, Qpp', -
    double *dC = dP + CT * NPT * 4;                               // [CT][kDerC]
    double *dX = dC + CT * kDerC;                                 // [CT][NYT][PP]: pair'

    // ---- derive this tile's tables (block-cooperative, ~2 items per thread)
#pragma unroll 1
    for (int w = threadIdx.x; w < CT * NYT; w += nthreads) {
      const int c = w / NYT;
      const double *yr = Ys + w * kRec;
      const double A = yr[0];
      dS[w * 4 + 0] = A;
      dS[w * 4 + 1] = yr[1];
      dS[w * 4 + 2] = MERGE ? fma(Ss[c * kScal + 1] * A, A, yr[2]) : yr[2];
      dS[w * 4 + 3] = 0.0;
    }
#pragma unroll 1
    for (int w = threadIdx.x; w < CT * NPT; w += nthreads) {
      const int c = w / NPT;
      const double *pr = Ps + w * kRec;
      const double B = pr[0];
      dP[w * 4 + 0] = B;
      dP[w * 4 + 1] = pr[1];
      dP[w * 4 + 2] = MERGE ? fma(Ss[c * kScal + 1] * B, B, pr[2]) : pr[2];
      dP[w * 4 + 3] = 0.0;
    }
    if (MODEL != M_IDEAL) {
#pragma unroll 1
      for (int w = threadIdx.x; w < CT * NYT * NPT; w += nthreads) {
        const int c = w / (NYT * NPT), r = w - c * (NYT * NPT), j = r / NPT, k = r - j * NPT;
        const double *yr = Ys + (c * NYT + j) * kRec, *pr = Ps + (c * NPT + k) * kRec;
        double v = fma(pr[4], yr[4], -(pr[3] * yr[3]));                          // R2 U2 - R1 U1
        if (MERGE) v = fma(-2.0 * Ss[c * kScal + 1] * yr[0], pr[0], v);          // - 2 K2 A B
        dX[(c * NYT + j) * PP + k] = v;
      }
    }
#pragma unroll 1
    for (int c = threadIdx.x; c < CT; c += nthreads) {
      double bmax = Ps[c * NPT * kRec], babs = fabs(bmax);
      for (int k = 1; k < NPT; k++) { const double b = Ps[(c * NPT + k) * kRec]; bmax = fmax(bmax, b); babs = fmax(babs, fabs(b)); }
      dC[c * kDerC + 0] = Ss[c * kScal + 0];
      dC[c * kDerC + 1] = Ss[c * kScal + 1];
      dC[c * kDerC + 2] = 1.0 + Ss[c * kScal + 2];
      dC[c * kDerC + 3] = bmax;
      bool live = false;                                 // any slot of this cell that is not a dead record (padding, skipped cell,
      for (int j = 0; j < NYT; j++) live = live || Ys[(c * NYT + j) * kRec] < 0.5 * kDeadSlotA;   // slot owned by the other record set)
      *reinterpret_cast<int2 *>(dC + c * kDerC + 4) = make_int2((babs * hp.pT_max <= kQLimit) ? 0 : 1, live ? 0 : 1);
      dC[c * kDerC + 5] = 0.0;
    }
    __syncthreads();                                   // tables of tile t complete; every warp is past the inner loop of tile t - 1
    if (threadIdx.x == 0 && t + kStages < n_my_tiles) issue(t + kStages);     // stage st has been consumed by the derive pass

    for (int c = 0; c < CT; c++) {
      const double2 k01 = *reinterpret_cast<const double2 *>(dC + c * kDerC);
      const double2 k23 = *reinterpret_cast<const double2 *>(dC + c * kDerC + 2);
      const int2 flg = *reinterpret_cast<const int2 *>(dC + c * kDerC + 4);        // general-path flag, dead-cell flag
      if (flg.y != 0) continue;                         // nothing alive in this cell
      const bool cell_general = flg.x != 0;
      const double K0m = k01.x * m2;
      const double K2 = k01.y;                          // Chapman-Enskog / Jonah: coefficient of x
      const double K31 = k23.x;                         // 1 + K3 (Jonah), 1 otherwise
      // dead test of a slot: mT A > ln(DBL_MAX) + max_k pT B[k] (compared on the high words, so the sliver in between goes
      // through the exact path)
      const int dead_hi = cell_general ? 0x7fffffff : __double2hiint(fma(pT, k23.y, 709.79));

      double fq[NPT], pd[NPT], G[NPT], q[NEEDX ? NPT : 1]; int fm[NPT];
#pragma unroll
      for (int k = 0; k < NPT; k++) {
        const double2 v0 = *reinterpret_cast<const double2 *>(dP + (c * NPT + k) * 4);
        const double qpp = dP[(c * NPT + k) * 4 + 2];
        const double qk = pT * v0.x;                    // pT (cos ux + sin uy) / T
        if (NEEDX) q[k] = qk;
        pd[k] = pT * v0.y;                              // pT (cos dsigma_x + sin dsigma_y)
        G[k] = fma(pT2, qpp, K0m);                      // pT^2 Qpp' + K0 m^2
        exp_neg_poly(-qk, fq[k], fm[k]);                // e^{+q} = fq 2^fm
      }

#pragma unroll
      for (int j = 0; j < NYT; j++) {
        const double2 s0 = *reinterpret_cast<const double2 *>(dS + (c * NYT + j) * 4);
        const double a = mT * s0.x;                     // mT A: the slot part of u.p / T
        if (__double2hiint(a) > dead_hi) continue;      // every member dead: exact 0
        double *accj = acc + j * NPT;
        const double qyy = dS[(c * NYT + j) * 4 + 2];
        const double *xr = dX + (c * NYT + j) * PP;
        const double cpm = mT * s0.y;                   // mT (cosh dsigma_tau + sinh dsigma_eta / tau)
        double pe; int ne;
        exp_neg_poly(a, pe, ne);                        // e^{-a} = pe 2^ne
        int n[NPT], nmin, nmax;
#pragma unroll
        for (int k = 0; k < NPT; k++) n[k] = ne + fm[k];
        nmin = n[0]; nmax = n[0];
#pragma unroll
        for (int k = 1; k < NPT; k++) { nmin = min(nmin, n[k]); nmax = max(nmax, n[k]); }
        if (nmax <= kNDead && !cell_general) continue;
        const double H = mT2 * qyy;                     // mT^2 Qyy'

        if (__builtin_expect(nmin >= kNNormal && !cell_general, 1)) {
          // ---------------- fast paths: every member alive with a normal e^{-x}
          double av[NPT], g[NPT], fe[NPT];
#pragma unroll
          for (int k = 0; k < NPT; k++) {
            const double pm = pe * fq[k];
            av[k] = __hiloint2double(__double2hiint(pm) + (n[k] << 20), __double2loint(pm));
          }
          if (MODEL == M_IDEAL) {
            if (nmax <= kNUltra) {
#pragma unroll
              for (int k = 0; k < NPT; k++) fe[k] = av[k];
            } else if (nmax <= kNDilute) {
#pragma unroll
              for (int k = 0; k < NPT; k++) fe[k] = av[k] * fma(av[k], av[k], fma(nsign, av[k], 1.0));
            } else {
#pragma unroll
              for (int k = 0; k < NPT; k++) fe[k] = av[k] * rcp_fast(fma(sign, av[k], 1.0));
            }
            accumulate_group<NPT>(accj, pd, cpm, fe, nullptr, thr_hi);
            continue;
          }
          // delta-f polynomial (without feqbar)
          double dfs[NPT];
#pragma unroll
          for (int k = 0; k < NPT; k++) {
            const double s = fma(mTpT, xr[k], H + G[k]);
            if (NEEDX) {
              const double x = a - q[k];
              dfs[k] = fma(s, rcp_fast(x), K2 * x);     // [..] / (u.p) + (..) (u.p)
            } else dfs[k] = s;
          }
          if (nmax <= kNUltra) {                        // feqbar = 1 exactly
#pragma unroll
            for (int k = 0; k < NPT; k++) { g[k] = dfs[k] + K31; fe[k] = av[k]; }
          } else if (nmax <= kNDilute) {
#pragma unroll
            for (int k = 0; k < NPT; k++) {
              const double fb = fma(av[k], av[k], fma(nsign, av[k], 1.0));
              g[k] = fma(fb, dfs[k], K31); fe[k] = av[k] * fb;
            }
          } else {
#pragma unroll
            for (int k = 0; k < NPT; k++) {
              const double fb = rcp_fast(fma(sign, av[k], 1.0));
              g[k] = fma(fb, dfs[k], K31); fe[k] = av[k] * fb;
            }
          }
          unsigned hmax = (unsigned)__double2hiint(g[0]);
#pragma unroll
          for (int k = 1; k < NPT; k++) hmax = max(hmax, (unsigned)__double2hiint(g[k]));
          if (hmax > hp.reg_chk) {
#pragma unroll
            for (int k = 0; k < NPT; k++) g[k] = clamp_g(g[k], reg_lo, reg_hi);
          }
          accumulate_group<NPT>(accj, pd, cpm, fe, g, thr_hi);
        } else {
          // ---------------- general path: some member is dead, sub-normal, or the cell's factors are out of range; per member
          //                  exactly as cf_kernels.cu decides it (x itself, exp_neg_slow)
#pragma unroll
          for (int k = 0; k < NPT; k++) {
            const double x = a - (NEEDX ? q[k] : pT * dP[(c * NPT + k) * 4]);      // the same a - fl(pT B) as everywhere else
            double avk;
            if (cell_general) avk = exp_finite(x) ? exp_neg(x) : 0.0;
            else avk = exp_neg_slow(x, pe * fq[k], n[k]);
            const double fb = rcp_fast(fma(sign, avk, 1.0));
            const double fe1 = avk * fb;
            double g1 = 1.0;
            if (MODEL != M_IDEAL) {
              const double s = fma(mTpT, xr[k], H + G[k]);
              const double d = NEEDX ? fma(s, rcp_fast(x), K2 * x) : s;
              g1 = clamp_g(fma(fb, d, K31), reg_lo, reg_hi);
            }
            accumulate_pos(accj[k], pd[k] + cpm, fe1 * g1, thr_hi);
          }
        }
      }
    }
  }

  __syncthreads();                                     // the stage area becomes the epilogue's scratch
  hot_epilogue<NYT, NPT, false>(hp, acc, stage_base, chunk, gb, ty, tp, lane_valid, ipart, ipT);
}

// ------------------------------------------------------------------------------------------------ dispatch
// Register-tile shapes of the factored kernel; is3d_options.tile_variant = 17 + k selects entry k.
struct FShape { int nyt, npt, ct, minb; };
static const FShape kFShapes[kNumFactoredVariants] = {{7, 3, 8, 3}, {7, 4, 8, 3}, {7, 3, 8, 4}, {7, 6, 8, 2}};

bool factored_supported(int model, const Layout &L)
{
  return !L.dim2 && (model == M_LIN14 || model == M_LINCE || model == M_JONAHLIN || model == M_IDEAL);
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
  *nyt = s.nyt; *npt = s.npt; *ct = s.ct; *max_warps = kMaxWarps;
}

template <int MODEL, int NYT, int NPT, int MINB>
static cudaError_t launch_f(const HotParams &hp, cudaStream_t st, size_t *smem_out)
{
  const Layout &L = hp.L;
  constexpr int PP = PairPitch<NPT>::v;
  const size_t stage_doubles = (size_t)L.ct * (NYT * kRec + NPT * kRec + kScal);
  const size_t der_doubles = (size_t)L.ct * (NYT * 4 + NPT * 4 + kDerC + NYT * PP);
  size_t smem = (kStages * stage_doubles + 2 * der_doubles) * 8 + kStages * sizeof(uint64_t);
  smem = std::max(smem, hot_epilogue_scratch_bytes(hp, NYT, false, hp.n_warps * 32));
  if (smem_out) *smem_out = smem;
  auto kern = cf_factored_kernel<MODEL, NYT, NPT, MINB>;
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
    case 1: return launch_f<MODEL, 7, 4, 3>(hp, st, smem_out);
    case 2: return launch_f<MODEL, 7, 3, 4>(hp, st, smem_out);
    case 3: return launch_f<MODEL, 7, 6, 2>(hp, st, smem_out);
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
