// cf_prepare.cu -- per-cell pre-pass of the smooth Cooper-Frye path (sm_100a).
//
// Turns the 18-23 stored scalars of each freeze-out cell into the records the hot kernels stream:
//   * per-cell set-up of EmissionFunctionArray::calculate_dN_pTdpTdphidy (reference smooth_kernels.cpp:118-242):
//     u^tau, the u.dsigma <= 0 skip, reconstruction of pi^{tau mu}, pi^{eta eta}, the delta-f coefficients
//     (natural cubic spline in T, deltafReader.cpp:325-395) and the shear / bulk coefficient combinations;
//   * the factorisation of the per-evaluation quantities into (cell, rapidity-slot) and (cell, phi) vectors
//     ("verified algebraic restructuring"):
//        u.p / T   = mT * Ax[slot] - pT * Bx[phi]
//        p.dsigma  = mT * Cp[slot] + W[slot] * pT * Dp[phi]
//        coef * pi^{mu nu} p_mu p_nu = mT^2 Qyy[slot] + pT^2 Qpp[phi] + mT pT (R2[phi] U2[slot] - R1[phi] U1[slot])
//     with U1 = cosh(y - eta), U2 = tau sinh(y - eta).
// Compiled with -fmad=false so that the set-up arithmetic rounds like the reference's scalar C++.
#include "cf_internal.h"
#include <cmath>

namespace is3d {

__device__ __forceinline__ double spline_eval(const Spline &s, double xv, bool &range_error)
{
  const int n = s.n;
  if (!(xv >= s.x[0] && xv <= s.x[n - 1])) { range_error = true; return 0.0; }
  int lo = 0, hi = n - 1;
  while (hi > lo + 1) { int mid = (hi + lo) >> 1; if (s.x[mid] > xv) hi = mid; else lo = mid; }
  const double x_lo = s.x[lo], dx = s.x[lo + 1] - x_lo, dy = s.y[lo + 1] - s.y[lo];
  const double c_i = s.c[lo], c_ip1 = s.c[lo + 1];
  const double b = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
  const double d = (c_ip1 - c_i) / (3.0 * dx);
  const double t = xv - x_lo;
  return s.y[lo] + t * (b + t * (c_i + t * d));
}

// per-cell quantities kept in shared memory between the two phases of the prepare kernel
struct CellVH {
  double tau, eta, inv_tau, invT, ut, ux, uy, un, dat, dax, day, dan;
  double sc, pitt, pitx, pity, pitn, pixx, pixy, pixn, piyy, piyn, pinn;
  int valid;
};

// record position -> cell of the caller's arrays (-1: padding).  Positions differ from cell indices when the cells have
// been grouped by (tau, r) bin for the spacetime distributions.
__device__ __forceinline__ int64_t source_cell(const RawCells &cells, const Layout &L, int64_t pos)
{
  if (cells.gather) return pos < L.n_cells_pad ? cells.gather[pos] : -1;
  return pos < cells.n ? pos : -1;
}

constexpr int kPrepCells = 16;      // cells per prepare block
constexpr int kPrepThreads = 128;

// padding, skipped (u.dsigma <= 0) and out-of-table cells: A = kDeadSlotA makes every evaluation dead, the hot kernels skip the group
__device__ __forceinline__ void dummy_slot(double *r) { r[0] = kDeadSlotA; r[1] = 0.0; r[2] = 0.0; r[3] = 0.0; r[4] = 0.0; r[5] = 0.0; }
__device__ __forceinline__ void dummy_phi(double *r) { r[0] = 0.0; r[1] = 0.0; r[2] = 0.0; r[3] = 0.0; r[4] = 0.0; r[5] = 0.0; }

// DFM = 1 (14 moment) or 2 (Chapman-Enskog)
template <int DFM>
__global__ void __launch_bounds__(kPrepThreads)
prepare_vh_kernel(RawCells cells, PrepTables tab, Layout L, int include_shear, int include_bulk,
                  double *__restrict__ Y, double *__restrict__ P, double *__restrict__ S, PrepCounters *counters)
{
  __shared__ CellVH sc_[kPrepCells];
  const int64_t cell0 = (int64_t)blockIdx.x * kPrepCells;

  // ---- phase 1: one thread per cell, scalar set-up (smooth_kernels.cpp:118-242)
  if (threadIdx.x < kPrepCells) {
    const int64_t i = source_cell(cells, L, cell0 + threadIdx.x);
    CellVH c;
    c.valid = 0;
    double K0 = 0.0, K2 = 0.0;
    if (i >= 0) {
      const double tau = cells.tau[i], tau2 = tau * tau;
      const double ux = cells.ux[i], uy = cells.uy[i], un = cells.un[i];
      const double ut = sqrt(1.0 + ux * ux + uy * uy + tau2 * un * un);
      const double dat = cells.dat[i], dax = cells.dax[i], day = cells.day[i], dan = cells.dan[i];
      const double udsigma = ut * dat + ux * dax + uy * day + un * dan;
      if (udsigma <= 0.0) {
        atomicAdd(&counters->skipped, 1ULL);
      } else {
        const double T = cells.T[i], Pr = cells.P[i], E = cells.E[i];
        double pixx = 0, pixy = 0, pixn = 0, piyy = 0, piyn = 0, pinn = 0, pitn = 0, pity = 0, pitx = 0, pitt = 0;
        if (include_shear) {
          const double ux2 = ux * ux, uy2 = uy * uy, ut2 = ut * ut, utperp = sqrt(1.0 + ux * ux + uy * uy);
          pixx = cells.pixx[i]; pixy = cells.pixy[i]; pixn = cells.pixn[i]; piyy = cells.piyy[i]; piyn = cells.piyn[i];
          pinn = (pixx * (ux2 - ut2) + piyy * (uy2 - ut2) + 2.0 * (pixy * ux * uy + tau2 * un * (pixn * ux + piyn * uy))) / (tau2 * utperp * utperp);
          pitn = (pixn * ux + piyn * uy + tau2 * pinn * un) / ut;
          pity = (pixy * ux + piyy * uy + tau2 * piyn * un) / ut;
          pitx = (pixx * ux + pixy * uy + tau2 * pixn * un) / ut;
          pitt = (pitx * ux + pity * uy + tau2 * pitn * un) / ut;
        }
        const double bulkPi = include_bulk ? cells.bulkPi[i] : 0.0;
        bool bad = false;
        const double T4 = T * T * T * T;
        double shear_coeff, bulk0, bulk2;
        if (DFM == 1) {
          const double c0 = spline_eval(tab.c0, T, bad) / T4;
          const double c2 = spline_eval(tab.c2, T, bad) / T4;
          shear_coeff = 0.5 / (T * T * (E + Pr));
          bulk0 = c0 - c2;
          bulk2 = 4.0 * c2 - c0;
          // df_bulk = (bulk0 m^2 + bulk2 (u.p)^2) Pi with u.p = T x   (bulk1 * baryon = 0 without muB)
          K0 = bulkPi * bulk0;
          K2 = bulkPi * bulk2 * T * T;
          c.sc = shear_coeff;
        } else {
          const double F = spline_eval(tab.F, T, bad) * T;
          const double betabulk = spline_eval(tab.betabulk, T, bad) * T4;
          const double betapi = spline_eval(tab.betapi, T, bad) * T4;
          shear_coeff = 0.5 / (betapi * T);
          bulk0 = F / (T * T * betabulk);
          bulk2 = 1.0 / (3.0 * T * betabulk);
          // df = [shear_coeff pi.p.p - Pi bulk2 m^2] / (u.p) + Pi (bulk0 + bulk2) (u.p), u.p = T x
          K0 = -bulkPi * bulk2 / T;
          K2 = bulkPi * (bulk0 + bulk2) * T;
          c.sc = shear_coeff / T;
        }
        if (bad) {
          atomicAdd(&counters->range_error, 1ULL);
        } else {
          c.valid = 1;
          c.tau = tau; c.eta = cells.eta[i]; c.inv_tau = 1.0 / tau; c.invT = 1.0 / T;
          c.ut = ut; c.ux = ux; c.uy = uy; c.un = un; c.dat = dat; c.dax = dax; c.day = day; c.dan = dan;
          c.pitt = pitt; c.pitx = pitx; c.pity = pity; c.pitn = pitn; c.pixx = pixx; c.pixy = pixy; c.pixn = pixn;
          c.piyy = piyy; c.piyn = piyn; c.pinn = pinn;
        }
      }
    }
    if (!c.valid) { K0 = 0.0; K2 = 0.0; }
    sc_[threadIdx.x] = c;
    const int64_t ip = cell0 + threadIdx.x;
    if (ip < L.n_cells_pad) {
      double *s = S + ip * kScal;
      s[0] = K0; s[1] = K2; s[2] = 0.0; s[3] = 0.0;
    }
  }
  __syncthreads();

  // ---- phase 2a: slot records, one (cell, slot) per thread-iteration
  const int slots_pad = L.n_ytiles * L.nst;
  for (int w = threadIdx.x; w < kPrepCells * slots_pad; w += kPrepThreads) {
    const int lc = w / slots_pad, j = w - lc * slots_pad;
    const int64_t i = cell0 + lc;
    if (i >= L.n_cells_pad) continue;
    const int ty = j / L.nst, jj = j - ty * L.nst;
    double *r = Y + (((int64_t)ty * L.n_cells_pad + i) * L.nst + jj) * L.rec_y;
    const CellVH &c = sc_[lc];
    if (!c.valid || j >= L.n_slots) { dummy_slot(r); continue; }
    double yv, eta, wgt;
    if (L.dim2) { yv = 0.0; eta = tab.slot_y[j]; wgt = tab.slot_w[j]; }
    else        { yv = tab.slot_y[j]; eta = c.eta; wgt = 1.0; }
    const double ch = cosh(yv - eta), sh = sinh(yv - eta);
    const double tsh = c.tau * sh;
    r[0] = (ch * c.ut - tsh * c.un) * c.invT;
    r[1] = wgt * (ch * c.dat + (sh * c.inv_tau) * c.dan);
    r[2] = c.sc * (c.pitt * ch * ch + c.pinn * tsh * tsh - 2.0 * c.pitn * tsh * ch);
    r[3] = ch;
    r[4] = tsh;
    r[5] = wgt;
  }

  // ---- phase 2b: phi records
  const int phis_pad = L.n_ptiles * L.npt;
  for (int w = threadIdx.x; w < kPrepCells * phis_pad; w += kPrepThreads) {
    const int lc = w / phis_pad, k = w - lc * phis_pad;
    const int64_t i = cell0 + lc;
    if (i >= L.n_cells_pad) continue;
    const int tp = k / L.npt, kk = k - tp * L.npt;
    double *r = P + (((int64_t)tp * L.n_cells_pad + i) * L.npt + kk) * kRec;
    const CellVH &c = sc_[lc];
    if (!c.valid || k >= L.n_phi) { dummy_phi(r); continue; }
    const double cs = tab.cosphi[k], sn = tab.sinphi[k];
    r[0] = (cs * c.ux + sn * c.uy) * c.invT;
    r[1] = cs * c.dax + sn * c.day;
    r[2] = c.sc * (c.pixx * cs * cs + c.piyy * sn * sn + 2.0 * c.pixy * cs * sn);
    r[3] = 2.0 * c.sc * (c.pitx * cs + c.pity * sn);
    r[4] = 2.0 * c.sc * (c.pixn * cs + c.piyn * sn);
    r[5] = 0.0;
  }
}


// =====================================================================================================================
// Modified equilibrium (df_mode 3 Mike, 4 Jonah): calculate_dN_ptdptdphidy_feqmod, smooth_kernels.cpp:396-996
// =====================================================================================================================
// Gauss-Laguerre thermal sums (gaussThermal.cpp:7-69), summed in node order like the reference
__device__ __forceinline__ double gl_neq(const double *root, const double *w, int n, double mbar, double alphaB, double baryon, double sign)
{
  double s = 0.0;
  for (int k = 0; k < n; k++) { const double p = root[k], Ebar = sqrt(p * p + mbar * mbar); s += w[k] * (p * exp(p) / (exp(Ebar - baryon * alphaB) + sign)); }
  return s;
}
__device__ __forceinline__ double gl_J10(const double *root, const double *w, int n, double mbar, double alphaB, double baryon, double sign)
{
  double s = 0.0;
  for (int k = 0; k < n; k++) {
    const double p = root[k], Ebar = sqrt(p * p + mbar * mbar), q = exp(Ebar - baryon * alphaB) + sign;
    s += w[k] * (p * exp(p + Ebar - baryon * alphaB) / (q * q));
  }
  return s;
}
__device__ __forceinline__ double gl_J20(const double *root, const double *w, int n, double mbar, double alphaB, double baryon, double sign)
{
  double s = 0.0;
  for (int k = 0; k < n; k++) {
    const double p = root[k], Ebar = sqrt(p * p + mbar * mbar), q = exp(Ebar - baryon * alphaB) + sign;
    s += w[k] * (Ebar * exp(p + Ebar - baryon * alphaB) / (q * q));
  }
  return s;
}

// 3x3 inverse: LU with partial pivoting, then column-wise solves -- the sequence gsl_linalg_LU_decomp / LU_invert run
// at smooth_kernels.cpp:689-707
__device__ void lu_inverse3(const double Ain[9], double inv[9])
{
  double A[9]; int p[3] = {0, 1, 2};
  for (int i = 0; i < 9; i++) A[i] = Ain[i];
  for (int j = 0; j < 2; j++) {
    double mx = fabs(A[j * 3 + j]); int ip = j;
    for (int i = j + 1; i < 3; i++) { const double a = fabs(A[i * 3 + j]); if (a > mx) { mx = a; ip = i; } }
    if (ip != j) { for (int k = 0; k < 3; k++) { const double t = A[j * 3 + k]; A[j * 3 + k] = A[ip * 3 + k]; A[ip * 3 + k] = t; } const int t = p[j]; p[j] = p[ip]; p[ip] = t; }
    const double ajj = A[j * 3 + j];
    if (ajj != 0.0)
      for (int i = j + 1; i < 3; i++) {
        const double aij = A[i * 3 + j] / ajj; A[i * 3 + j] = aij;
        for (int k = j + 1; k < 3; k++) A[i * 3 + k] = A[i * 3 + k] - aij * A[j * 3 + k];
      }
  }
  for (int col = 0; col < 3; col++) {
    double x[3];
    for (int i = 0; i < 3; i++) x[i] = (p[i] == col) ? 1.0 : 0.0;
    for (int i = 1; i < 3; i++) { double s = x[i]; for (int k = 0; k < i; k++) s -= A[i * 3 + k] * x[k]; x[i] = s; }
    for (int i = 2; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < 3; k++) s -= A[i * 3 + k] * x[k]; x[i] = s / A[i * 3 + i]; }
    for (int i = 0; i < 3; i++) inv[i * 3 + col] = x[i];
  }
}

// v = A^-1 b with iterative refinement (the reference refines p' = A^-1 p per momentum, smooth_kernels.cpp:907-919; applying
// the same sweeps to the factor vectors is equivalent by linearity and keeps ill-conditioned cells, detA << 1, in parity)
__device__ __forceinline__ void solve_refined3(const double A[9], const double Ainv[9], const double b[3], double v[3])
{
  for (int i = 0; i < 3; i++) v[i] = Ainv[i * 3 + 0] * b[0] + Ainv[i * 3 + 1] * b[1] + Ainv[i * 3 + 2] * b[2];
  for (int it = 0; it < 5; it++) {
    double r[3];
    for (int i = 0; i < 3; i++) r[i] = __fma_rn(-A[i * 3 + 2], v[2], __fma_rn(-A[i * 3 + 1], v[1], __fma_rn(-A[i * 3 + 0], v[0], b[i])));
    if (r[0] == 0.0 && r[1] == 0.0 && r[2] == 0.0) break;
    for (int i = 0; i < 3; i++) v[i] += Ainv[i * 3 + 0] * r[0] + Ainv[i * 3 + 1] * r[1] + Ainv[i * 3 + 2] * r[2];
  }
}

struct CellFM {
  // common
  double tau, eta, inv_tau, ut, ux, uy, un, dat, dax, day, dan;
  // linear (breakdown / narrow) branch: Chapman-Enskog or Jonah-linear coefficients, x = u.p / T
  double invT, scl, pitt, pitx, pity, pitn, pixx, pixy, pixn, piyy, piyn, pinn;
  // feqmod branch
  double Xt, Xx, Xy, Xn, Yx, Yy, Zt, Zn, A[9], Ainv[9], invTmod, detA, eta_scale;
  int valid, breaks_down;
};

// DFM = 3 (Mike) or 4 (Jonah)
template <int DFM>
__global__ void __launch_bounds__(kPrepThreads)
prepare_feqmod_kernel(RawCells cells, PrepTables tab, Layout L, int include_shear, int include_bulk,
                      double *__restrict__ YF, double *__restrict__ PF, double *__restrict__ SF,
                      double *__restrict__ YL, double *__restrict__ PL, double *__restrict__ SL,
                      double *__restrict__ cellaux, PrepCounters *counters)
{
  __shared__ CellFM sc_[kPrepCells];
  const int64_t cell0 = (int64_t)blockIdx.x * kPrepCells;
  const double two_pi2_hbarC3 = 2.0 * pow(M_PI, 2) * pow(0.197327053, 3);

  if (threadIdx.x < kPrepCells) {
    const int64_t i = source_cell(cells, L, cell0 + threadIdx.x);
    CellFM c; c.valid = 0; c.breaks_down = 0; c.detA = 1.0; c.eta_scale = 1.0;
    double sF[4] = {1.0e12, 0.0, 0.0, 0.0};      // feqmod scalars: 1/T_mod^2 (invalid cell: E'/T_mod = 1e6 m, dead), per-cell renorm (Jonah), -, -
    double sL[4] = {0.0, 0.0, 0.0, 0.0};         // linear scalars: K0, K2, K3, -
    double aux[8] = {0, 0, 0, 0, 0, 0, 0, 0};    // per-cell inputs of the (cell, species) renormalisation kernel
    if (i >= 0) {
      const double tau = cells.tau[i], tau2 = tau * tau;
      const double ux = cells.ux[i], uy = cells.uy[i], un = cells.un[i];
      const double ut = sqrt(1.0 + ux * ux + uy * uy + tau2 * un * un);
      const double dat = cells.dat[i], dax = cells.dax[i], day = cells.day[i], dan = cells.dan[i];
      const double udsigma = ut * dat + ux * dax + uy * day + un * dan;
      if (udsigma <= 0.0) {
        atomicAdd(&counters->skipped, 1ULL);
      } else {
        const double ux2 = ux * ux, uy2 = uy * uy, ut2 = ut * ut;
        const double uperp = sqrt(ux * ux + uy * uy), utperp = sqrt(1.0 + ux * ux + uy * uy);
        const double T = cells.T[i], Pr = cells.P[i], E = cells.E[i];
        double pixx = 0, pixy = 0, pixn = 0, piyy = 0, piyn = 0, pinn = 0, pitn = 0, pity = 0, pitx = 0, pitt = 0;
        if (include_shear) {
          pixx = cells.pixx[i]; pixy = cells.pixy[i]; pixn = cells.pixn[i]; piyy = cells.piyy[i]; piyn = cells.piyn[i];
          pinn = (pixx * (ux2 - ut2) + piyy * (uy2 - ut2) + 2.0 * (pixy * ux * uy + tau2 * un * (pixn * ux + piyn * uy))) / (tau2 * utperp * utperp);
          pitn = (pixn * ux + piyn * uy + tau2 * pinn * un) / ut;
          pity = (pixy * ux + piyy * uy + tau2 * piyn * un) / ut;
          pitx = (pixx * ux + pixy * uy + tau2 * pixn * un) / ut;
          pitt = (pitx * ux + pity * uy + tau2 * pitn * un) / ut;
        }
        double bulkPi = include_bulk ? cells.bulkPi[i] : 0.0;
        if (DFM == 4) {                                               // :588-594
          const double mx = tab.bulkPi_over_Peq_max;
          if (L.dx) {                                                 // calculate_dN_dX_feqmod compares with <=, >= (:1709-1710)
            if (bulkPi <= -Pr) bulkPi = -(1.0 - 1.e-5) * Pr;
            else if (bulkPi / Pr >= mx) bulkPi = Pr * (mx - 1.e-5);
          } else {
            if (bulkPi < -Pr) bulkPi = -(1.0 - 1.e-5) * Pr;
            else if (bulkPi / Pr > mx) bulkPi = Pr * (mx - 1.e-5);
          }
        }
        bool bad = false;
        const double T4 = T * T * T * T;
        double F = 0, betabulk = 1, betapi, lambda = 0, z = 1, dlam = 0, dz = 0;
        if (DFM == 3) {
          F = spline_eval(tab.F, T, bad) * T;
          betabulk = spline_eval(tab.betabulk, T, bad) * T4;
          betapi = spline_eval(tab.betapi, T, bad) * T4;
        } else {
          const double l2 = spline_eval(tab.lam2, bulkPi / Pr, bad);
          lambda = (bulkPi < 0.0) ? -sqrt(l2) : (bulkPi > 0.0 ? sqrt(l2) : 0.0);
          z = spline_eval(tab.z, bulkPi / Pr, bad);
          betapi = spline_eval(tab.betapi, T, bad) * T4;
          dlam = bulkPi / (5.0 * betapi - 3.0 * Pr * (E + Pr) / E);
          dz = -3.0 * dlam * Pr / E;
        }
        // Milne basis (viscous_correction.cpp:10-29)
        const double sinhL = tau * un / utperp, coshL = ut / utperp;
        double Xt = uperp * coshL, Zt = sinhL, Xn = uperp * sinhL / tau, Zn = coshL / tau;
        double Xx = 1.0, Yx = 0.0, Xy = 0.0, Yy = 1.0;
        if (uperp > 1.e-5) { Xx = utperp * ux / uperp; Yx = -uy / uperp; Xy = utperp * uy / uperp; Yy = ux / uperp; }
        // pi^{mu nu} in the local rest frame (viscous_correction.cpp:121-142)
        const double pixx_LRF = pitt * Xt * Xt + pixx * Xx * Xx + piyy * Xy * Xy + tau2 * tau2 * pinn * Xn * Xn
          + 2.0 * (-Xt * (pitx * Xx + pity * Xy) + pixy * Xx * Xy + tau2 * Xn * (pixn * Xx + piyn * Xy - pitn * Xt));
        const double pixy_LRF = Yx * (-pitx * Xt + pixx * Xx + pixy * Xy + tau2 * pixn * Xn) + Yy * (-pity * Xt + pixy * Xx + piyy * Xy + tau2 * piyn * Xn);
        const double pixz_LRF = Zt * (pitt * Xt - pitx * Xx - pity * Xy - tau2 * pitn * Xn) - tau2 * Zn * (pitn * Xt - pixn * Xx - piyn * Xy - tau2 * pinn * Xn);
        const double piyy_LRF = pixx * Yx * Yx + 2.0 * pixy * Yx * Yy + piyy * Yy * Yy;
        const double piyz_LRF = -Zt * (pitx * Yx + pity * Yy) + tau2 * Zn * (pixn * Yx + piyn * Yy);
        const double pizz_LRF = -(pixx_LRF + piyy_LRF);
        double T_mod = T;
        if (DFM == 3) T_mod = T + bulkPi * F / betabulk;
        const double shear_mod = 0.5 / betapi;
        double bulk_mod = bulkPi / (3.0 * betabulk);
        if (DFM == 4) bulk_mod = lambda;
        const double Axx = 1.0 + pixx_LRF * shear_mod + bulk_mod, Axy = pixy_LRF * shear_mod, Axz = pixz_LRF * shear_mod;
        const double Ayy = 1.0 + piyy_LRF * shear_mod + bulk_mod, Ayz = piyz_LRF * shear_mod, Azz = 1.0 + pizz_LRF * shear_mod + bulk_mod;
        const double detA = Axx * (Ayy * Azz - Ayz * Ayz) - Axy * (Axy * Azz - Ayz * Axz) + Axz * (Axy * Ayz - Ayy * Axz);
        const double A[9] = {Axx, Axy, Axz, Axy, Ayy, Ayz, Axz, Ayz, Azz};
        for (int q9 = 0; q9 < 9; q9++) c.A[q9] = A[q9];
        lu_inverse3(A, c.Ainv);
        const double neq_fact = T * T * T / two_pi2_hbarC3, J20_fact = T * neq_fact;
        int breaks = 0;
        if (DFM == 3) {                                               // does_feqmod_breakdown, emissionfunction.cpp:109-138
          const double mbar0 = tab.mass_pion0 / T;
          const double neq0 = neq_fact * gl_neq(tab.gla_root1, tab.gla_w1, tab.gla_n, mbar0, 0., 0., -1.);
          const double J20_0 = J20_fact * gl_J20(tab.gla_root2, tab.gla_w2, tab.gla_n, mbar0, 0., 0., -1.);
          const double dn0 = bulkPi * (neq0 + J20_0 * F / T / T) / betabulk;
          if (detA <= tab.deta_min || (neq0 + dn0) < 0.0) breaks = 1;
        }
        if (!bad && !breaks && !(T_mod > 0.0)) bad = true;            // modified temperature must stay positive
        if (bad) {
          atomicAdd(&counters->range_error, 1ULL);
        } else {
          c.valid = 1; c.breaks_down = breaks;
          if (breaks) { atomicAdd(&counters->breakdown, 1ULL); atomicAdd(&counters->linear_items, 1ULL); }
          c.tau = tau; c.eta = cells.eta[i]; c.inv_tau = 1.0 / tau; c.ut = ut; c.ux = ux; c.uy = uy; c.un = un;
          c.dat = dat; c.dax = dax; c.day = day; c.dan = dan;
          c.pitt = pitt; c.pitx = pitx; c.pity = pity; c.pitn = pitn; c.pixx = pixx; c.pixy = pixy; c.pixn = pixn;
          c.piyy = piyy; c.piyn = piyn; c.pinn = pinn;
          c.Xt = Xt; c.Xx = Xx; c.Xy = Xy; c.Xn = Xn; c.Yx = Yx; c.Yy = Yy; c.Zt = Zt; c.Zn = Zn;
          c.invT = 1.0 / T; c.invTmod = 1.0 / T_mod; c.detA = detA;
          if (detA > tab.deta_min && (L.dx || detA < 1.0) && L.dim2) c.eta_scale = detA;      // :728-729; :1850-1853 has no detA < 1 test
          // linear-branch coefficients (:641-644 / :868-869), written in terms of x = u.p / T
          const double shear_coeff = 0.5 / (betapi * T);
          c.scl = shear_coeff / T;
          if (DFM == 3) {
            const double bulk0 = F / (T * T * betabulk), bulk2 = 1.0 / (3.0 * T * betabulk);
            sL[0] = -bulkPi * bulk2 / T; sL[1] = bulkPi * (bulk0 + bulk2) * T; sL[2] = 0.0;
          } else {
            sL[0] = -dlam / (T * T); sL[1] = dlam; sL[2] = dz - 3.0 * dlam;
          }
          sF[0] = c.invTmod * c.invTmod;
          // well-conditioned momentum rescaling (Frobenius norm of A^-1 <= 3, identity: 1.73): the hot kernel may expand |p'|^2
          {
            double fro = 0.0;
            for (int q9 = 0; q9 < 9; q9++) fro += c.Ainv[q9] * c.Ainv[q9];
            sF[2] = (!breaks && fro <= 9.0) ? 1.0 : 0.0;
          }
          // renormalisation: Jonah z / detA per cell; Mike per (cell, species) in renorm_kernel
          double rn = 1.0;
          if (include_bulk && DFM == 4) rn = z;
          if (L.dx && (isnan(rn / detA) || isinf(rn / detA))) rn = 0.0;     // :1887 tests renorm / detA in 2+1D too
          if (!L.dim2) rn /= detA;                                   // :780-784 (DIMENSION == 3)
          if (isnan(rn) || isinf(rn)) rn = 0.0;                      // the reference skips the species (:773-778)
          sF[1] = fabs(rn);
          aux[0] = T; aux[1] = T_mod; aux[2] = bulkPi / betabulk; aux[3] = F; aux[4] = (L.dim2 ? 1.0 : detA); aux[5] = 1.0;
          aux[6] = L.dx ? detA : 1.0;
        }
      }
    }
    if (!c.valid) { sF[0] = 1.0e12; sF[1] = 0.0; sL[0] = sL[1] = sL[2] = 0.0; aux[5] = 0.0; }
    sc_[threadIdx.x] = c;
    const int64_t ip = cell0 + threadIdx.x;
    if (ip < L.n_cells_pad) {
      for (int k = 0; k < kScal; k++) { SF[ip * kScal + k] = sF[k]; SL[ip * kScal + k] = sL[k]; }
      if (cellaux) for (int k = 0; k < 8; k++) cellaux[ip * 8 + k] = aux[k];
    }
  }
  __syncthreads();

  // ---- slot records: feqmod set (v = A^-1 V / T_mod) and linear set (only for breakdown cells / narrow slots)
  const int slots_pad = L.n_ytiles * L.nst;
  for (int w = threadIdx.x; w < kPrepCells * slots_pad; w += kPrepThreads) {
    const int lc = w / slots_pad, j = w - lc * slots_pad;
    const int64_t i = cell0 + lc;
    if (i >= L.n_cells_pad) continue;
    const int ty = j / L.nst, jj = j - ty * L.nst;
    const int64_t off = (((int64_t)ty * L.n_cells_pad + i) * L.nst + jj) * kRec;
    double *rf = YF + off, *rl = YL + off;
    const CellFM &c = sc_[lc];
    dummy_slot(rl);
    // slots the other record set owns (and padding): |p'| / T_mod >= 1e6 mT, every evaluation dead
    rf[0] = kDeadSlotA; rf[1] = 0.0; rf[2] = 0.0; rf[3] = kDeadSlotA * kDeadSlotA; rf[4] = 0.0; rf[5] = 0.0;
    if (!c.valid || j >= L.n_slots) continue;
    double yv, eta, wgt;
    if (L.dim2) { yv = 0.0; eta = tab.slot_y[j]; wgt = tab.slot_w[j]; }
    else        { yv = tab.slot_y[j]; eta = c.eta; wgt = 1.0; }
    bool linear = c.breaks_down != 0;
    if (!linear && !L.dim2 && !L.dx && c.detA < 0.01 && fabs(yv - eta) < c.detA) {      // narrow breakdown, :813-819 (commented out at :1927-1935)
      linear = true;
      atomicAdd(&counters->linear_items, 1ULL);
    }
    if (linear) {
      const double ch = cosh(yv - eta), sh = sinh(yv - eta), tsh = c.tau * sh;
      rl[0] = (ch * c.ut - tsh * c.un) * c.invT;
      rl[1] = L.dx ? wgt * (ch * c.dat + (sh * c.inv_tau) * c.dan)               // :1948
                   : wgt * (ch * c.dat) + (sh * c.inv_tau) * c.dan;              // eta weight not on the dsigma_eta term (:831)
      rl[2] = c.scl * (c.pitt * ch * ch + c.pinn * tsh * tsh - 2.0 * c.pitn * tsh * ch);
      rl[3] = ch; rl[4] = tsh; rl[5] = wgt;
    } else {
      const double arg = yv - c.eta_scale * eta;
      const double ch = cosh(arg), sh = sinh(arg), tsh = c.tau * sh;
      // p_LRF = mT V + pT W with V = (-Xt ch + Xn tau sh, 0, -Zt ch + Zn tau sh)   (:889-891)
      const double Vv[3] = {-c.Xt * ch + c.Xn * tsh, 0.0, -c.Zt * ch + c.Zn * tsh};
      double sol[3];
      solve_refined3(c.A, c.Ainv, Vv, sol);
      const double v0 = sol[0] * c.invTmod, v1 = sol[1] * c.invTmod, v2 = sol[2] * c.invTmod;
      rf[0] = v0; rf[1] = v1; rf[2] = v2; rf[3] = v0 * v0 + v1 * v1 + v2 * v2;
      rf[4] = L.dx ? wgt * (ch * c.dat + (sh * c.inv_tau) * c.dan)               // :2001
                   : wgt * (ch * c.dat) + (sh * c.inv_tau) * c.dan;              // :884
      rf[5] = wgt;
    }
  }

  // ---- phi records
  const int phis_pad = L.n_ptiles * L.npt;
  for (int w = threadIdx.x; w < kPrepCells * phis_pad; w += kPrepThreads) {
    const int lc = w / phis_pad, k = w - lc * phis_pad;
    const int64_t i = cell0 + lc;
    if (i >= L.n_cells_pad) continue;
    const int tp = k / L.npt, kk = k - tp * L.npt;
    const int64_t off = (((int64_t)tp * L.n_cells_pad + i) * L.npt + kk) * kRec;
    double *rf = PF + off, *rl = PL + off;
    const CellFM &c = sc_[lc];
    dummy_phi(rf); dummy_phi(rl);
    if (!c.valid || k >= L.n_phi) continue;
    const double cs = tab.cosphi[k], sn = tab.sinphi[k];
    // linear set (cells that break down, and cells with narrow slots, need it; it is cheap to always write)
    rl[0] = (cs * c.ux + sn * c.uy) * c.invT;
    rl[1] = cs * c.dax + sn * c.day;
    rl[2] = c.scl * (c.pixx * cs * cs + c.piyy * sn * sn + 2.0 * c.pixy * cs * sn);
    rl[3] = 2.0 * c.scl * (c.pitx * cs + c.pity * sn);
    rl[4] = 2.0 * c.scl * (c.pixn * cs + c.piyn * sn);
    if (!c.breaks_down) {
      const double Wv[3] = {c.Xx * cs + c.Xy * sn, c.Yx * cs + c.Yy * sn, 0.0};
      double sol[3];
      solve_refined3(c.A, c.Ainv, Wv, sol);
      const double w0 = sol[0] * c.invTmod, w1 = sol[1] * c.invTmod, w2 = sol[2] * c.invTmod;
      rf[0] = w0; rf[1] = w1; rf[2] = w2; rf[3] = w0 * w0 + w1 * w1 + w2 * w2;
      rf[4] = cs * c.dax + sn * c.day;
    }
  }
}

// Mike's renormalisation n_linear / n_mod per (cell, species), smooth_kernels.cpp:745-784
__global__ void renorm_mike_kernel(const double *__restrict__ cellaux, int64_t n_cells_pad, int n_species, PrepTables tab,
                                   const double *__restrict__ mass, const double *__restrict__ sign, const double *__restrict__ degeneracy,
                                   const double *__restrict__ baryon, int include_bulk, double *__restrict__ renorm)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int s = blockIdx.y;
  if (i >= n_cells_pad) return;
  const double *a = cellaux + i * 8;
  double rn = 0.0;
  if (a[5] != 0.0) {
    rn = 1.0;
    if (include_bulk) {
      const double two_pi2_hbarC3 = 2.0 * pow(M_PI, 2) * pow(0.197327053, 3);
      const double T = a[0], T_mod = a[1], dn_fact = a[2], F = a[3];
      const double m = mass[s], g = degeneracy[s], b = baryon[s], sg = sign[s];
      const double neq_fact = T * T * T / two_pi2_hbarC3, J20_fact = T * neq_fact, N10_fact = neq_fact;
      const double nmod_fact = T_mod * T_mod * T_mod / two_pi2_hbarC3;
      const double mbar = m / T, mbar_mod = m / T_mod;
      const double neq = neq_fact * g * gl_neq(tab.gla_root1, tab.gla_w1, tab.gla_n, mbar, 0.0, b, sg);
      const double N10 = b * N10_fact * g * gl_J10(tab.gla_root1, tab.gla_w1, tab.gla_n, mbar, 0.0, b, sg);
      const double J20 = J20_fact * g * gl_J20(tab.gla_root2, tab.gla_w2, tab.gla_n, mbar, 0.0, b, sg);
      const double n_linear = neq + dn_fact * (neq + N10 * 0.0 + J20 * F / T / T);
      const double n_mod = nmod_fact * g * gl_neq(tab.gla_root1, tab.gla_w1, tab.gla_n, mbar_mod, 0.0, b, sg);
      rn = n_linear / n_mod;
    }
    if (isnan(rn) || isinf(rn) || isnan(rn / a[6]) || isinf(rn / a[6])) rn = 0.0;     // a[6] = detA for the spacetime variant, else 1
    else { rn /= a[4]; rn = fabs(rn); }
  }
  renorm[(int64_t)s * n_cells_pad + i] = rn;
}

cudaError_t launch_prepare_feqmod(const is3d_flags &fl, const RawCells &cells, const PrepTables &tab, const Layout &L,
                                  double *YF, double *PF, double *SF, double *YL, double *PL, double *SL,
                                  const double *mass, const double *sign, const double *degeneracy, const double *baryon,
                                  double *renorm, PrepCounters *counters, cudaStream_t st)
{
  const int64_t nblk = (L.n_cells_pad + kPrepCells - 1) / kPrepCells;
  if (nblk == 0) return cudaSuccess;
  // df_mode 3 keeps 8 per-cell doubles for the (cell, species) renormalisation pass in the tail of the renorm buffer
  double *cellaux = (fl.df_mode == 3) ? renorm + (int64_t)L.n_species * L.n_cells_pad : nullptr;
  if (fl.df_mode == 3)
    prepare_feqmod_kernel<3><<<(unsigned)nblk, kPrepThreads, 0, st>>>(cells, tab, L, fl.include_shear_deltaf, fl.include_bulk_deltaf, YF, PF, SF, YL, PL, SL, cellaux, counters);
  else
    prepare_feqmod_kernel<4><<<(unsigned)nblk, kPrepThreads, 0, st>>>(cells, tab, L, fl.include_shear_deltaf, fl.include_bulk_deltaf, YF, PF, SF, YL, PL, SL, cellaux, counters);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || fl.df_mode != 3) return e;
  dim3 grid((unsigned)((L.n_cells_pad + 127) / 128), (unsigned)L.n_species);
  renorm_mike_kernel<<<grid, 128, 0, st>>>(cellaux, L.n_cells_pad, L.n_species, tab, mass, sign, degeneracy, baryon, fl.include_bulk_deltaf, renorm);
  return cudaGetLastError();
}

// =====================================================================================================================
// Anisotropic hydro, PL matching: calculate_dN_pTdpTdphidy_VAH_PL, smooth_kernels.cpp:2140-2393
// =====================================================================================================================
struct CellVAH {
  double tau, eta, inv_tau, invL, ut, ux, uy, un, dat, dax, day, dan, zt, zn;
  double pitt, pitx, pity, pitn, pixx, pixy, pixn, piyy, piyn, pinn, Wt, Wx, Wy, Wn;
  double c1Pi, c3, c4, xiL_L2;
  int valid;
};

__global__ void __launch_bounds__(kPrepThreads)
prepare_vah_kernel(RawCells cells, PrepTables tab, Layout L, int include_shear, int include_bulk,
                   double *__restrict__ Y, double *__restrict__ P, double *__restrict__ S, PrepCounters *counters)
{
  __shared__ CellVAH sc_[kPrepCells];
  const int64_t cell0 = (int64_t)blockIdx.x * kPrepCells;
  if (threadIdx.x < kPrepCells) {
    const int64_t i = source_cell(cells, L, cell0 + threadIdx.x);
    CellVAH c; c.valid = 0;
    double s4[4] = {0.0, 0.0, 0.0, 0.0};
    if (i >= 0) {
      const double tau = cells.tau[i], tau2 = tau * tau;
      const double ux = cells.ux[i], uy = cells.uy[i], un = cells.un[i];
      const double ut = sqrt(1.0 + ux * ux + uy * uy + tau2 * un * un);
      const double u0 = sqrt(1.0 + ux * ux + uy * uy);
      c.valid = 1;
      c.tau = tau; c.eta = cells.eta[i]; c.inv_tau = 1.0 / tau; c.ut = ut; c.ux = ux; c.uy = uy; c.un = un;
      c.dat = cells.dat[i]; c.dax = cells.dax[i]; c.day = cells.day[i]; c.dan = cells.dan[i];
      c.zt = tau * un / u0; c.zn = ut / (u0 * tau);
      c.pitt = cells.pitt[i]; c.pitx = cells.pitx[i]; c.pity = cells.pity[i]; c.pitn = cells.pitn[i]; c.pixx = cells.pixx[i];
      c.pixy = cells.pixy[i]; c.pixn = cells.pixn[i]; c.piyy = cells.piyy[i]; c.piyn = cells.piyn[i]; c.pinn = cells.pinn[i];
      const double Wx = cells.Wx[i], Wy = cells.Wy[i];
      const double Wt = (ux * Wx + uy * Wy) * ut / (u0 * u0);
      c.Wt = Wt; c.Wx = Wx; c.Wy = Wy; c.Wn = Wt * un / ut;
      const double Lambda = cells.Lambda[i], aL = cells.aL[i];
      const double bulkPi = include_bulk ? cells.bulkPi[i] : 0.0;
      const double xiL = 1.0 / (aL * aL) - 1.0;
      c.invL = 1.0 / Lambda;
      c.xiL_L2 = xiL * c.invL * c.invL;
      c.c3 = include_shear ? cells.c3[i] : 0.0;
      c.c4 = include_shear ? cells.c4[i] : 0.0;
      c.c1Pi = cells.c1[i] * bulkPi;
      s4[0] = cells.c0[i] * bulkPi;                         // x m^2
      s4[1] = cells.c2[i] * bulkPi * Lambda * Lambda;       // x (u.p / Lambda)^2
    }
    sc_[threadIdx.x] = c;
    const int64_t ip = cell0 + threadIdx.x;
    if (ip < L.n_cells_pad) { double *s = S + ip * kScal; s[0] = s4[0]; s[1] = s4[1]; s[2] = 0.0; s[3] = 0.0; }
  }
  __syncthreads();

  const int slots_pad = L.n_ytiles * L.nst;
  for (int w = threadIdx.x; w < kPrepCells * slots_pad; w += kPrepThreads) {
    const int lc = w / slots_pad, j = w - lc * slots_pad;
    const int64_t i = cell0 + lc;
    if (i >= L.n_cells_pad) continue;
    const int ty = j / L.nst, jj = j - ty * L.nst;
    double *r = Y + (((int64_t)ty * L.n_cells_pad + i) * L.nst + jj) * kRecVah;
    const CellVAH &c = sc_[lc];
    if (!c.valid || j >= L.n_slots) { r[0] = kDeadSlotA; for (int q = 1; q < kRecVah; q++) r[q] = 0.0; continue; }
    double yv, eta, wgt;
    if (L.dim2) { yv = 0.0; eta = tab.slot_y[j]; wgt = tab.slot_w[j] * tab.eta_delta; }        // :2175-2183
    else        { yv = tab.slot_y[j]; eta = c.eta; wgt = 1.0; }
    const double ch = cosh(yv - eta), sh = sinh(yv - eta), tsh = c.tau * sh;
    const double Z = ch * c.zt - tsh * c.zn;                // z.p / mT
    const double WY = c.Wt * ch - c.Wn * tsh;               // W.p / mT (rapidity part)
    r[0] = (ch * c.ut - tsh * c.un) * c.invL;               // u.p / (mT Lambda)
    r[1] = wgt * (ch * c.dat + (sh * c.inv_tau) * c.dan);
    r[2] = c.c4 * (c.pitt * ch * ch + c.pinn * tsh * tsh - 2.0 * c.pitn * tsh * ch) + c.c3 * Z * WY + c.c1Pi * Z * Z;
    r[3] = ch; r[4] = tsh;
    r[5] = c.c3 * Z;
    r[6] = c.xiL_L2 * Z * Z;
    r[7] = wgt;
  }
  const int phis_pad = L.n_ptiles * L.npt;
  for (int w = threadIdx.x; w < kPrepCells * phis_pad; w += kPrepThreads) {
    const int lc = w / phis_pad, k = w - lc * phis_pad;
    const int64_t i = cell0 + lc;
    if (i >= L.n_cells_pad) continue;
    const int tp = k / L.npt, kk = k - tp * L.npt;
    double *r = P + (((int64_t)tp * L.n_cells_pad + i) * L.npt + kk) * kRec;
    const CellVAH &c = sc_[lc];
    if (!c.valid || k >= L.n_phi) { dummy_phi(r); continue; }
    const double cs = tab.cosphi[k], sn = tab.sinphi[k];
    r[0] = (cs * c.ux + sn * c.uy) * c.invL;
    r[1] = cs * c.dax + sn * c.day;
    r[2] = c.c4 * (c.pixx * cs * cs + c.piyy * sn * sn + 2.0 * c.pixy * cs * sn);
    r[3] = 2.0 * c.c4 * (c.pitx * cs + c.pity * sn);
    r[4] = 2.0 * c.c4 * (c.pixn * cs + c.piyn * sn);
    r[5] = c.Wx * cs + c.Wy * sn;
  }
}

cudaError_t launch_prepare_vah(const is3d_flags &fl, const RawCells &cells, const PrepTables &tab, const Layout &L,
                               double *Y, double *P, double *S, PrepCounters *counters, cudaStream_t st)
{
  const int64_t nblk = (L.n_cells_pad + kPrepCells - 1) / kPrepCells;
  if (nblk == 0) return cudaSuccess;
  prepare_vah_kernel<<<(unsigned)nblk, kPrepThreads, 0, st>>>(cells, tab, L, fl.include_shear_deltaf, fl.include_bulk_deltaf, Y, P, S, counters);
  return cudaGetLastError();
}

cudaError_t launch_prepare_vh(const is3d_flags &fl, const RawCells &cells, const PrepTables &tab, const Layout &L,
                              double *Y, double *P, double *S, PrepCounters *counters, cudaStream_t st)
{
  const int64_t nblk = (L.n_cells_pad + kPrepCells - 1) / kPrepCells;
  if (nblk == 0) return cudaSuccess;
  if (fl.df_mode == 1)
    prepare_vh_kernel<1><<<(unsigned)nblk, kPrepThreads, 0, st>>>(cells, tab, L, fl.include_shear_deltaf, fl.include_bulk_deltaf, Y, P, S, counters);
  else
    prepare_vh_kernel<2><<<(unsigned)nblk, kPrepThreads, 0, st>>>(cells, tab, L, fl.include_shear_deltaf, fl.include_bulk_deltaf, Y, P, S, counters);
  return cudaGetLastError();
}

// =====================================================================================================================
// Sampler mean yield: EmissionFunctionArray::calculate_total_yield, emissionfunction_sampling_kernels.cpp
// :653-831.  Without baryon diffusion the species sum factorises, sum_s [u.dsigma (n_eq,s + Pi dn_bulk,s)] (df_mode 1-3) or
// u.dsigma z(Pi/P) n_eq,s (df_mode 4), so the device only reduces three per-cell quantities over the surface:
//   S0 = sum u.dsigma,  S1 = sum u.dsigma Pi,  S2 = sum u.dsigma z(Pi/P),  cells with u.dsigma <= 0 skipped (:690).
// Deterministic: block b owns cells [b * kYieldSpan, (b + 1) * kYieldSpan), threads stride through them, fixed-order tree in
// shared memory, one triple per block; the host adds the triples in block order.
constexpr int kYieldThreads = 256, kYieldSpan = 8192;
__global__ void __launch_bounds__(kYieldThreads)
yield_kernel(RawCells cells, PrepTables tab, int df_mode, int include_bulk, double *__restrict__ partial, PrepCounters *counters)
{
  __shared__ double red[3][kYieldThreads];
  const int64_t lo = (int64_t)blockIdx.x * kYieldSpan, hi = min(lo + (int64_t)kYieldSpan, cells.n);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  unsigned long long skipped = 0, bad_cells = 0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += kYieldThreads) {
    const double tau = cells.tau[i], tau2 = tau * tau;
    const double ux = cells.ux[i], uy = cells.uy[i], un = cells.un[i];
    const double ut = sqrt(1.0 + ux * ux + uy * uy + tau2 * un * un);
    const double udsigma = ut * cells.dat[i] + ux * cells.dax[i] + uy * cells.day[i] + un * cells.dan[i];
    if (udsigma <= 0.0) { skipped++; continue; }
    double bulkPi = include_bulk ? cells.bulkPi[i] : 0.0;
    double z = 0.0;
    if (df_mode == 4) {
      const double Pr = cells.P[i], mx = tab.bulkPi_over_Peq_max;
      if (bulkPi <= -Pr) bulkPi = -(1.0 - 1.e-5) * Pr;                    // :753-754
      else if (bulkPi / Pr >= mx) bulkPi = Pr * (mx - 1.e-5);
      bool bad = false;
      z = spline_eval(tab.z, bulkPi / Pr, bad);
      if (bad) { bad_cells++; continue; }
    }
    s0 += udsigma; s1 += udsigma * bulkPi; s2 += udsigma * z;
  }
  red[0][threadIdx.x] = s0; red[1][threadIdx.x] = s1; red[2][threadIdx.x] = s2;
  __syncthreads();
  for (int w = kYieldThreads / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) for (int q = 0; q < 3; q++) red[q][threadIdx.x] += red[q][threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) for (int q = 0; q < 3; q++) partial[(int64_t)blockIdx.x * 3 + q] = red[q][0];
  if (skipped) atomicAdd(&counters->skipped, skipped);
  if (bad_cells) atomicAdd(&counters->range_error, bad_cells);
}

cudaError_t launch_yield(const RawCells &cells, const PrepTables &tab, int df_mode, int include_bulk, double *partial, int *n_blocks,
                         PrepCounters *counters, cudaStream_t st)
{
  *n_blocks = (int)((cells.n + kYieldSpan - 1) / kYieldSpan);
  if (*n_blocks == 0) return cudaSuccess;
  yield_kernel<<<(unsigned)*n_blocks, kYieldThreads, 0, st>>>(cells, tab, df_mode, include_bulk, partial, counters);
  return cudaGetLastError();
}

}  // namespace is3d
