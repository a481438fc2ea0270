// cf_strict.cu -- "strict" diagnostic variant of the linear-delta-f spectra kernel ("hard parts"): evaluates every
// (cell, species, pT, phi, y [, eta]) term in the reference's own operation order -- no hoisting, no factorisation, libm-style
// cosh / sinh / exp per evaluation, divisions where the reference divides -- and is compiled with -fmad=false.  It is slow
// (one thread per momentum bin, a serial loop over the cells, ~50x the hot kernel) and exists for one purpose: in bins where a
// single cell dominates and its 1 + df nearly cancels, the reference's own value carries rounding noise of ~ulp / (1 + df); the
// restructured kernels and this one then both differ from the oracle by comparable, independent amounts, which is the evidence
// behind the conditioning allowance of tests/common.py (tests/test_gpu_parity.py::test_ill_conditioned_bins_strict_variant).
// Mirrors emissionfunction_smooth_kernels.cpp:118-339 (df_mode 1, 2; 3+1D and 2+1D); selected with is3d_options.tile_variant = 99.
#include "cf_internal.h"
#include <cmath>

namespace is3d {

namespace {

struct StrictCell {            // per-cell quantities of :118-242, in the reference's units
  double tau, tau2, eta, dat, dax, day, dan, ut, ux, uy, un, T;
  double pitt, pitx, pity, pitn, pixx, pixy, pixn, piyy, piyn, pinn, bulkPi;
  double shear_coeff, bulk0_coeff, bulk2_coeff;
  int valid;
};

__device__ double spline_strict(const Spline &s, double xv, bool &bad)
{
  const int n = s.n;
  if (!(xv >= s.x[0] && xv <= s.x[n - 1])) { bad = true; return 0.0; }
  int lo = 0, hi = n - 1;
  while (hi > lo + 1) { int mid = (hi + lo) >> 1; if (s.x[mid] > xv) hi = mid; else lo = mid; }
  const double x_lo = s.x[lo], dx = s.x[lo + 1] - x_lo, dy = s.y[lo + 1] - s.y[lo];
  const double c_i = s.c[lo], c_ip1 = s.c[lo + 1];
  const double b = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
  const double d = (c_ip1 - c_i) / (3.0 * dx);
  const double t = xv - x_lo;
  return s.y[lo] + t * (b + t * (c_i + t * d));
}

__global__ void strict_cells_kernel(RawCells cells, PrepTables tab, int df_mode, int include_shear, int include_bulk, StrictCell *out, PrepCounters *counters)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cells.n) return;
  StrictCell c; c.valid = 0;
  const double tau = cells.tau[i], tau2 = tau * tau;
  const double ux = cells.ux[i], uy = cells.uy[i], un = cells.un[i];
  const double ut = sqrt(1.0 + ux * ux + uy * uy + tau2 * un * un);
  const double dat = cells.dat[i], dax = cells.dax[i], day = cells.day[i], dan = cells.dan[i];
  const double udsigma = ut * dat + ux * dax + uy * day + un * dan;
  if (udsigma <= 0.0) { atomicAdd(&counters->skipped, 1ULL); out[i] = c; return; }
  const double ux2 = ux * ux, uy2 = uy * uy, ut2 = ut * ut, utperp = sqrt(1.0 + ux * ux + uy * uy);
  const double T = cells.T[i], P = cells.P[i], E = cells.E[i];
  double pitt = 0, pitx = 0, pity = 0, pitn = 0, pixx = 0, pixy = 0, pixn = 0, piyy = 0, piyn = 0, pinn = 0;
  if (include_shear) {
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
  if (df_mode == 1) {                                              // deltafReader.cpp:339-352 + smooth_kernels.cpp:220-226
    const double c0 = spline_strict(tab.c0, T, bad) / T4, c2 = spline_strict(tab.c2, T, bad) / T4;
    c.shear_coeff = 0.5 / (T * T * (E + P)); c.bulk0_coeff = c0 - c2; c.bulk2_coeff = 4.0 * c2 - c0;
  } else {                                                         // :353-360 + :227-234
    const double F = spline_strict(tab.F, T, bad) * T, betabulk = spline_strict(tab.betabulk, T, bad) * T4, betapi = spline_strict(tab.betapi, T, bad) * T4;
    c.shear_coeff = 0.5 / (betapi * T); c.bulk0_coeff = F / (T * T * betabulk); c.bulk2_coeff = 1.0 / (3.0 * T * betabulk);
  }
  if (bad) { atomicAdd(&counters->range_error, 1ULL); out[i] = c; return; }
  c.valid = 1; c.tau = tau; c.tau2 = tau2; c.eta = cells.eta[i]; c.dat = dat; c.dax = dax; c.day = day; c.dan = dan;
  c.ut = ut; c.ux = ux; c.uy = uy; c.un = un; c.T = T;
  c.pitt = pitt; c.pitx = pitx; c.pity = pity; c.pitn = pitn; c.pixx = pixx; c.pixy = pixy; c.pixn = pixn; c.piyy = piyy; c.piyn = piyn; c.pinn = pinn;
  c.bulkPi = bulkPi;
  out[i] = c;
}

// one thread per bin (ipart fastest, like the output array); cells in index order; :246-339
__global__ void strict_kernel(const StrictCell *__restrict__ cells, int64_t n_cells, Layout L, const double *__restrict__ mass, const double *__restrict__ sign,
                              const double *__restrict__ degeneracy, const double *__restrict__ pT_tab, const double *__restrict__ cosphi,
                              const double *__restrict__ sinphi, const double *__restrict__ y_tab, const double *__restrict__ eta_tab,
                              const double *__restrict__ eta_w, int n_eta, int df_mode, int regulate, int outflow, double prefactor, double *__restrict__ out)
{
  const int64_t n_bins = (int64_t)L.n_species * L.n_pT * L.n_phi * (L.dim2 ? 1 : L.n_y_out);
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_bins) return;
  const int ipart = (int)(b % L.n_species);
  const int ipT = (int)((b / L.n_species) % L.n_pT);
  const int iphip = (int)((b / ((int64_t)L.n_species * L.n_pT)) % L.n_phi);
  const int iy = (int)(b / ((int64_t)L.n_species * L.n_pT * L.n_phi));
  const double m = mass[ipart], mass2 = m * m, sgn = sign[ipart], g = degeneracy[ipart];
  const double pT = pT_tab[ipT], mT = sqrt(mass2 + pT * pT);
  const double px = pT * cosphi[iphip], py = pT * sinphi[iphip];
  const double y = L.dim2 ? 0.0 : y_tab[iy];
  const int eta_pts = L.dim2 ? n_eta : 1;
  double total = 0.0;
  for (int64_t ic = 0; ic < n_cells; ic++) {
    const StrictCell c = cells[ic];
    if (!c.valid) continue;
    const double mT_over_tau = mT / c.tau;
    double pdotdsigma_f_eta_sum = 0.0;
    for (int ieta = 0; ieta < eta_pts; ieta++) {
      const double eta = L.dim2 ? eta_tab[ieta] : c.eta, eta_weight = L.dim2 ? eta_w[ieta] : 1.0;
      const double pt = mT * cosh(y - eta);
      const double pn = mT_over_tau * sinh(y - eta);
      const double tau2_pn = c.tau2 * pn;
      const double pdotdsigma = eta_weight * (pt * c.dat + px * c.dax + py * c.day + pn * c.dan);
      if (outflow && pdotdsigma <= 0.0) continue;
      const double pdotu = pt * c.ut - px * c.ux - py * c.uy - tau2_pn * c.un;
      const double feq = 1.0 / (exp(pdotu / c.T) + sgn);
      const double feqbar = 1.0 - sgn * feq;
      const double pimunu_pmu_pnu = c.pitt * pt * pt + c.pixx * px * px + c.piyy * py * py + c.pinn * tau2_pn * tau2_pn
        + 2.0 * (-(c.pitx * px + c.pity * py) * pt + c.pixy * px * py + tau2_pn * (c.pixn * px + c.piyn * py - c.pitn * pt));
      double df;
      if (df_mode == 1) {
        const double df_shear = c.shear_coeff * pimunu_pmu_pnu;
        const double df_bulk = (c.bulk0_coeff * mass2 + (0.0 + c.bulk2_coeff * pdotu) * pdotu) * c.bulkPi;
        df = feqbar * (df_shear + df_bulk + 0.0);
      } else {
        const double df_shear = c.shear_coeff * pimunu_pmu_pnu / pdotu;
        const double df_bulk = (c.bulk0_coeff * pdotu + 0.0 + c.bulk2_coeff * (pdotu - mass2 / pdotu)) * c.bulkPi;
        df = feqbar * (df_shear + df_bulk + 0.0);
      }
      if (regulate) df = fmax(-1.0, fmin(df, 1.0));
      const double f = feq * (1.0 + df);
      pdotdsigma_f_eta_sum += (pdotdsigma * f);
    }
    total += (prefactor * g * pdotdsigma_f_eta_sum);
  }
  out[b] += total;
}

}  // namespace

size_t strict_cell_bytes() { return sizeof(StrictCell); }

cudaError_t launch_strict(const is3d_flags &fl, const RawCells &cells, const PrepTables &tab, const Layout &L, const double *mass, const double *sign,
                          const double *degeneracy, const double *pT, int n_eta, double prefactor, void *cell_scratch, double *dN_dev,
                          PrepCounters *counters, cudaStream_t st)
{
  if (cells.n == 0) return cudaSuccess;
  StrictCell *sc = reinterpret_cast<StrictCell *>(cell_scratch);
  strict_cells_kernel<<<(unsigned)((cells.n + 127) / 128), 128, 0, st>>>(cells, tab, fl.df_mode, fl.include_shear_deltaf, fl.include_bulk_deltaf, sc, counters);
  const int64_t n_bins = (int64_t)L.n_species * L.n_pT * L.n_phi * (L.dim2 ? 1 : L.n_y_out);
  strict_kernel<<<(unsigned)((n_bins + 127) / 128), 128, 0, st>>>(sc, cells.n, L, mass, sign, degeneracy, pT, tab.cosphi, tab.sinphi, tab.slot_y,
                                                                  tab.slot_y, tab.slot_w, n_eta, fl.df_mode, fl.regulate_deltaf, fl.outflow, prefactor, dN_dev);
  return cudaGetLastError();
}

}  // namespace is3d
