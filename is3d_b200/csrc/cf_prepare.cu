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

constexpr int kPrepCells = 16;      // cells per prepare block
constexpr int kPrepThreads = 128;

__device__ __forceinline__ void dummy_slot(double *r) { r[0] = 8.0; r[1] = 0.0; r[2] = 0.0; r[3] = 0.0; r[4] = 0.0; r[5] = 0.0; }
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
    const int64_t i = cell0 + threadIdx.x;
    CellVH c;
    c.valid = 0;
    double K0 = 0.0, K2 = 0.0;
    if (i < cells.n) {
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
    double *r = Y + (((int64_t)ty * L.n_cells_pad + i) * L.nst + jj) * kRec;
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

}  // namespace is3d
