/* oracle/cf_oracle.c -- plain-C CPU restatement of iS3D's smooth Cooper-Frye spectra path.
 *
 * TEST INFRASTRUCTURE ONLY (see cf_oracle.h).  Each function cites the reference lines it follows.  The arithmetic
 * keeps the reference's operation order per evaluation (compiled with -ffp-contract=off), so that it agrees with the
 * compiled reference (oracle/_ref) to a few ulp; the only deliberate differences are
 *   - cells with u.dsigma <= 0 contribute exactly 0 (the reference leaves stale scratch data there),
 *   - per bin, cell contributions are summed in cell order (the reference's SIMD reduction order is unspecified),
 *   - the loop nest is species-outermost so that OpenMP can split species without changing any sum.
 */
#include "cf_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static const double hbarC = 0.197327053; /* iS3D.h:9 */

/* ---------------------------------------------------------------- natural cubic spline (GSL cspline restated) */
void cfo_spline_init(const double *x, const double *y, int n, double *c)
{
  int i;
  c[0] = 0.0; c[n - 1] = 0.0;
  int N = n - 2;
  if (N < 1) return;
  double *g = (double *)calloc(N, sizeof(double)), *diag = (double *)calloc(N, sizeof(double));
  double *off = (double *)calloc(N, sizeof(double)), *gamma = (double *)calloc(N, sizeof(double));
  double *alpha = (double *)calloc(N, sizeof(double)), *z = (double *)calloc(N, sizeof(double));
  double *cc = (double *)calloc(N, sizeof(double));
  for (i = 0; i < N; i++) {
    double h_i = x[i + 1] - x[i], h_ip1 = x[i + 2] - x[i + 1];
    double yd_i = y[i + 1] - y[i], yd_ip1 = y[i + 2] - y[i + 1];
    double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    off[i] = h_ip1; diag[i] = 2.0 * (h_ip1 + h_i); g[i] = 3.0 * (yd_ip1 * g_ip1 - yd_i * g_i);
  }
  if (N == 1) { c[1] = g[0] / diag[0]; }
  else {
    alpha[0] = diag[0]; gamma[0] = off[0] / alpha[0];
    for (i = 1; i < N - 1; i++) { alpha[i] = diag[i] - off[i - 1] * gamma[i - 1]; gamma[i] = off[i] / alpha[i]; }
    alpha[N - 1] = diag[N - 1] - off[N - 2] * gamma[N - 2];
    z[0] = g[0];
    for (i = 1; i < N; i++) z[i] = g[i] - gamma[i - 1] * z[i - 1];
    for (i = 0; i < N; i++) cc[i] = z[i] / alpha[i];
    c[N] = cc[N - 1];
    for (i = N - 2; i >= 0; i--) c[i + 1] = cc[i] - gamma[i] * c[i + 2];
  }
  free(g); free(diag); free(off); free(gamma); free(alpha); free(z); free(cc);
}

double cfo_spline_eval(const double *x, const double *y, const double *c, int n, double xv, int *err)
{
  if (!(xv >= x[0] && xv <= x[n - 1])) { if (err) *err = 1; return 0.0; }
  int lo = 0, hi = n - 1;
  while (hi > lo + 1) { int mid = (hi + lo) / 2; if (x[mid] > xv) hi = mid; else lo = mid; }
  double dx = x[lo + 1] - x[lo], dy = y[lo + 1] - y[lo];
  double b = (dy / dx) - dx * (c[lo + 1] + 2.0 * c[lo]) / 3.0;
  double d = (c[lo + 1] - c[lo]) / (3.0 * dx);
  double t = xv - x[lo];
  return y[lo] + t * (b + t * (c[lo] + t * d));
}

/* spline second-derivative arrays are rebuilt per call site from the tables; cache them per table set */
typedef struct {
  const cfo_df_tables *tab;
  double *c0, *c2, *F, *betabulk, *betapi, *lam2, *z;
} spline_cache;

static void cache_build(spline_cache *sc, const cfo_df_tables *tab)
{
  int n = tab->n_T;
  sc->tab = tab;
  sc->c0 = (double *)calloc(n, sizeof(double)); sc->c2 = (double *)calloc(n, sizeof(double));
  sc->F = (double *)calloc(n, sizeof(double)); sc->betabulk = (double *)calloc(n, sizeof(double));
  sc->betapi = (double *)calloc(n, sizeof(double));
  cfo_spline_init(tab->T, tab->c0, n, sc->c0); cfo_spline_init(tab->T, tab->c2, n, sc->c2);
  cfo_spline_init(tab->T, tab->F, n, sc->F); cfo_spline_init(tab->T, tab->betabulk, n, sc->betabulk);
  cfo_spline_init(tab->T, tab->betapi, n, sc->betapi);
  sc->lam2 = sc->z = NULL;
  if (tab->n_jonah > 2 && tab->jonah_x) {
    sc->lam2 = (double *)calloc(tab->n_jonah, sizeof(double)); sc->z = (double *)calloc(tab->n_jonah, sizeof(double));
    cfo_spline_init(tab->jonah_x, tab->jonah_lambda2, tab->n_jonah, sc->lam2);
    cfo_spline_init(tab->jonah_x, tab->jonah_z, tab->n_jonah, sc->z);
  }
}
static void cache_free(spline_cache *sc)
{ free(sc->c0); free(sc->c2); free(sc->F); free(sc->betabulk); free(sc->betapi); free(sc->lam2); free(sc->z); }

/* Deltaf_Data::cubic_spline, deltafReader.cpp:325-395 (include_baryon = 0 path of evaluate_df_coefficients :486-504) */
static int df_eval(const spline_cache *sc, int df_mode, double T, double E, double P, double bulkPi, cfo_dfcoef *df)
{
  const cfo_df_tables *t = sc->tab;
  int n = t->n_T, err = 0;
  memset(df, 0, sizeof(*df));
  double T4 = T * T * T * T;
  switch (df_mode) {
    case 1:
      df->c0 = cfo_spline_eval(t->T, t->c0, sc->c0, n, T, &err) / T4;
      df->c1 = 0.0;
      df->c2 = cfo_spline_eval(t->T, t->c2, sc->c2, n, T, &err) / T4;
      df->c3 = 0.0; df->c4 = 0.0;
      df->shear14_coeff = 2.0 * T * T * (E + P);
      break;
    case 2: case 3:
      df->F = cfo_spline_eval(t->T, t->F, sc->F, n, T, &err) * T;
      df->G = 0.0;
      df->betabulk = cfo_spline_eval(t->T, t->betabulk, sc->betabulk, n, T, &err) * T4;
      df->betaV = 1.0;
      df->betapi = cfo_spline_eval(t->T, t->betapi, sc->betapi, n, T, &err) * T4;
      break;
    case 4: {
      if (!sc->lam2) return 2;
      double lambda_squared = cfo_spline_eval(t->jonah_x, t->jonah_lambda2, sc->lam2, t->n_jonah, bulkPi / P, &err);
      /* the reference leaves lambda uninitialised for bulkPi == 0; 0 is the value the spline gives there */
      if (bulkPi < 0.0) df->lambda = -sqrt(lambda_squared);
      else if (bulkPi > 0.0) df->lambda = sqrt(lambda_squared);
      else df->lambda = 0.0;
      df->z = cfo_spline_eval(t->jonah_x, t->jonah_z, sc->z, t->n_jonah, bulkPi / P, &err);
      df->betapi = cfo_spline_eval(t->T, t->betapi, sc->betapi, n, T, &err) * T4;
      df->delta_lambda = bulkPi / (5.0 * df->betapi - 3.0 * P * (E + P) / E);
      df->delta_z = -3.0 * df->delta_lambda * P / E;
      break;
    }
    default: return 3;
  }
  return err;
}

int cfo_df_coefficients(const cfo_df_tables *tab, int df_mode, double T, double E, double P, double bulkPi, cfo_dfcoef *out)
{
  spline_cache sc; cache_build(&sc, tab);
  int rc = df_eval(&sc, df_mode, T, E, P, bulkPi, out);
  cache_free(&sc);
  return rc;
}

/* ---------------------------------------------------------------- thermal integrals, gaussThermal.cpp:7-115 */
static double neq_int(double pbar, double mbar, double alphaB, double baryon, double sign)
{ double Ebar = sqrt(pbar * pbar + mbar * mbar); return pbar * exp(pbar) / (exp(Ebar - baryon * alphaB) + sign); }
static double J10_int(double pbar, double mbar, double alphaB, double baryon, double sign)
{ double Ebar = sqrt(pbar * pbar + mbar * mbar); double q = exp(Ebar - baryon * alphaB) + sign;
  return pbar * exp(pbar + Ebar - baryon * alphaB) / (q * q); }
static double J20_int(double pbar, double mbar, double alphaB, double baryon, double sign)
{ double Ebar = sqrt(pbar * pbar + mbar * mbar); double q = exp(Ebar - baryon * alphaB) + sign;
  return Ebar * exp(pbar + Ebar - baryon * alphaB) / (q * q); }
typedef double (*thermal_fn)(double, double, double, double, double);
static double gauss_thermal(thermal_fn f, const double *root, const double *weight, int n, double mbar, double alphaB, double baryon, double sign)
{ double s = 0.0; for (int k = 0; k < n; k++) s += weight[k] * f(root[k], mbar, alphaB, baryon, sign); return s; }
static double E_mod_int(double pbar, double mbar, double lambda, double sign)
{ double scale2 = (1.0 + lambda) * (1.0 + lambda); double Ebar = sqrt(pbar * pbar + mbar * mbar);
  return sqrt(pbar * pbar * scale2 + mbar * mbar) * exp(pbar) / (exp(Ebar) + sign); }
static double P_mod_int(double pbar, double mbar, double lambda, double sign)
{ double scale2 = (1.0 + lambda) * (1.0 + lambda); double Ebar = sqrt(pbar * pbar + mbar * mbar);
  return pbar * pbar * scale2 / sqrt(pbar * pbar * scale2 + mbar * mbar) * exp(pbar) / (exp(Ebar) + sign); }
typedef double (*mod_fn)(double, double, double, double);
static double gauss_mod(mod_fn f, const double *root, const double *weight, int n, double mbar, double lambda, double sign)
{ double s = 0.0; for (int k = 0; k < n; k++) s += weight[k] * f(root[k], mbar, lambda, sign); return s; }

/* Deltaf_Data::compute_jonah_coefficients, deltafReader.cpp:222-297 */
double cfo_jonah_tables(int n_particles, const double *mass, const double *degeneracy, const double *sign, double T,
                        const cfo_laguerre *gla, double *x, double *lambda2, double *z)
{
  const int jonah_points = 301;
  const double lambda_min = -1.0, lambda_max = 2.0;
  const double delta_lambda = (lambda_max - lambda_min) / ((double)jonah_points - 1.0);
  double xmax = -1.0;
  for (int i = 0; i < jonah_points; i++) {
    double lambda = lambda_min + (double)i * delta_lambda;
    double E = 0.0, P = 0.0, E_mod = 0.0, P_mod = 0.0;
    for (int n = 0; n < n_particles; n++) {
      double mbar = mass[n] / T;
      if (mass[n] == 0.0) continue;
      E += degeneracy[n] * gauss_mod(E_mod_int, gla->root2, gla->weight2, gla->n_points, mbar, 0.0, sign[n]);
      P += (1.0 / 3.0) * degeneracy[n] * gauss_mod(P_mod_int, gla->root2, gla->weight2, gla->n_points, mbar, 0.0, sign[n]);
      E_mod += degeneracy[n] * gauss_mod(E_mod_int, gla->root2, gla->weight2, gla->n_points, mbar, lambda, sign[n]);
      P_mod += (1.0 / 3.0) * degeneracy[n] * gauss_mod(P_mod_int, gla->root2, gla->weight2, gla->n_points, mbar, lambda, sign[n]);
    }
    double zz = E / E_mod;
    double r = (P_mod / P) * zz - 1.0;
    lambda2[i] = lambda * lambda; z[i] = zz; x[i] = r;
    xmax = fmax(xmax, r);
  }
  return xmax;
}

/* FO_data_reader::read_surf_VH averages, readindata.cpp:423-466 */
void cfo_surface_averages(const cfo_cells *c, double *out5)
{
  double Tavg = 0, Eavg = 0, Pavg = 0, muBavg = 0, nBavg = 0, vol = 0;
  for (int64_t i = 0; i < c->n_cells; i++) {
    double tau = c->tau[i], ux = c->ux[i], uy = c->uy[i], un = c->un[i];
    double ut = sqrt(1.0 + ux * ux + uy * uy + tau * tau * un * un);
    double dat = c->dat[i], dax = c->dax[i], day = c->day[i], dan = c->dan[i];
    double udsigma = ut * dat + ux * dax + uy * day + un * dan;
    double dsds = dat * dat - dax * dax - day * day - dan * dan / (tau * tau);
    double mag = fabs(udsigma) + sqrt(fabs(udsigma * udsigma - dsds));
    double muB = c->muB ? c->muB[i] : 0.0, nB = c->nB ? c->nB[i] : 0.0;
    vol += mag;
    Eavg += (c->E[i] * mag); Tavg += (c->T[i] * mag); Pavg += (c->P[i] * mag);
    muBavg += (muB * mag); nBavg += (nB * mag);
  }
  out5[0] = Tavg / vol; out5[1] = Eavg / vol; out5[2] = Pavg / vol; out5[3] = muBavg / vol; out5[4] = nBavg / vol;
}

/* ---------------------------------------------------------------- per-cell set-up shared by the kernels */
typedef struct {
  int skip;
  double tau, tau2, eta, dat, dax, day, dan, ut, ux, uy, un, T, P, E;
  double pitt, pitx, pity, pitn, pixx, pixy, pixn, piyy, piyn, pinn, bulkPi;
  double alphaB, Vt, Vx, Vy, Vn, baryon_enthalpy_ratio;
  cfo_dfcoef df;
  double shear_coeff, bulk0_coeff, bulk1_coeff, bulk2_coeff;
} cell_setup;

/* emissionfunction_smooth_kernels.cpp:118-197 (identical in :504-584) */
static void cell_common(const cfo_flags *fl, const cfo_cells *c, int64_t i, cell_setup *s)
{
  memset(s, 0, sizeof(*s));
  double tau = c->tau[i], tau2 = tau * tau;
  s->tau = tau; s->tau2 = tau2; s->eta = c->eta[i];
  s->dat = c->dat[i]; s->dax = c->dax[i]; s->day = c->day[i]; s->dan = c->dan[i];
  double ux = c->ux[i], uy = c->uy[i], un = c->un[i];
  double ut = sqrt(1.0 + ux * ux + uy * uy + tau2 * un * un);
  s->ux = ux; s->uy = uy; s->un = un; s->ut = ut;
  double udsigma = ut * s->dat + ux * s->dax + uy * s->day + un * s->dan;
  if (udsigma <= 0.0) { s->skip = 1; return; }
  double ux2 = ux * ux, uy2 = uy * uy, ut2 = ut * ut;
  double utperp = sqrt(1.0 + ux * ux + uy * uy);
  s->T = c->T[i]; s->P = c->P[i]; s->E = c->E[i];
  if (fl->include_shear) {
    double pixx = c->pixx[i], pixy = c->pixy[i], pixn = c->pixn[i], piyy = c->piyy[i], piyn = c->piyn[i];
    double pinn = (pixx * (ux2 - ut2) + piyy * (uy2 - ut2) + 2.0 * (pixy * ux * uy + tau2 * un * (pixn * ux + piyn * uy))) / (tau2 * utperp * utperp);
    double pitn = (pixn * ux + piyn * uy + tau2 * pinn * un) / ut;
    double pity = (pixy * ux + piyy * uy + tau2 * piyn * un) / ut;
    double pitx = (pixx * ux + pixy * uy + tau2 * pixn * un) / ut;
    double pitt = (pitx * ux + pity * uy + tau2 * pitn * un) / ut;
    s->pixx = pixx; s->pixy = pixy; s->pixn = pixn; s->piyy = piyy; s->piyn = piyn;
    s->pinn = pinn; s->pitn = pitn; s->pity = pity; s->pitx = pitx; s->pitt = pitt;
  }
  if (fl->include_bulk) s->bulkPi = c->bulkPi[i];
  if (fl->include_baryon && fl->include_diff) {
    double muB = c->muB[i], nB = c->nB[i];
    s->Vx = c->Vx[i]; s->Vy = c->Vy[i]; s->Vn = c->Vn[i];
    s->Vt = (s->Vx * ux + s->Vy * uy + tau2 * s->Vn * un) / ut;
    s->alphaB = muB / s->T;
    s->baryon_enthalpy_ratio = nB / (s->E + s->P);
  }
}

static int grid_dims(const cfo_flags *fl, const cfo_grid *g, int *y_pts, int *eta_pts)
{
  *y_pts = g->n_y; *eta_pts = 1;
  if (fl->dimension == 2) { *y_pts = 1; *eta_pts = g->n_eta; }
  return 0;
}

/* ---------------------------------------------------------------- a1: linear delta-f kernel, :28-393 */
int64_t cfo_smooth_vh(const cfo_flags *fl, const cfo_cells *c, const cfo_species *sp, const cfo_grid *g,
                      const cfo_df_tables *tab, double *dN, double *dN_abs)
{
  if (fl->df_mode != 1 && fl->df_mode != 2) return -1;
  if (fl->include_baryon) return -2;                    /* bilinear (T, muB) lookup: reference indexes out of bounds (R8) */
  const double prefactor = pow(2.0 * M_PI * hbarC, -3);
  const int npart = sp->n, npT = g->n_pT, nphi = g->n_phi;
  int y_pts, eta_pts; grid_dims(fl, g, &y_pts, &eta_pts);
  const int64_t n = c->n_cells;
  spline_cache sc; cache_build(&sc, tab);
  cell_setup *cs = (cell_setup *)malloc(sizeof(cell_setup) * (size_t)(n > 0 ? n : 1));
  int64_t skipped = 0; int bad = 0;
  for (int64_t i = 0; i < n; i++) {
    cell_setup *s = &cs[i];
    cell_common(fl, c, i, s);
    if (s->skip) { skipped++; continue; }
    if (df_eval(&sc, fl->df_mode, s->T, s->E, s->P, s->bulkPi, &s->df)) bad = 1;
    if (fl->df_mode == 1) {                              /* :222-229 */
      s->shear_coeff = 0.5 / (s->T * s->T * (s->E + s->P));
      s->bulk0_coeff = s->df.c0 - s->df.c2;
      s->bulk1_coeff = s->df.c1;
      s->bulk2_coeff = 4.0 * s->df.c2 - s->df.c0;
    } else {                                             /* :230-237 */
      s->shear_coeff = 0.5 / (s->df.betapi * s->T);
      s->bulk0_coeff = s->df.F / (s->T * s->T * s->df.betabulk);
      s->bulk1_coeff = s->df.G / s->df.betabulk;
      s->bulk2_coeff = 1.0 / (3.0 * s->T * s->df.betabulk);
    }
  }
  cache_free(&sc);
  if (bad) { free(cs); return -3; }
  double *cosphi = (double *)malloc(sizeof(double) * nphi), *sinphi = (double *)malloc(sizeof(double) * nphi);
  for (int k = 0; k < nphi; k++) { cosphi[k] = cos(g->phi[k]); sinphi[k] = sin(g->phi[k]); }

#pragma omp parallel for schedule(dynamic, 1)
  for (int ipart = 0; ipart < npart; ipart++) {
    double mass = sp->mass[ipart], mass2 = mass * mass, sign = sp->sign[ipart];
    double degeneracy = sp->degeneracy[ipart], baryon = sp->baryon[ipart];
    for (int64_t icell = 0; icell < n; icell++) {
      const cell_setup *s = &cs[icell];
      if (s->skip) continue;
      double chem = baryon * s->alphaB;
      for (int ipT = 0; ipT < npT; ipT++) {
        double pT = g->pT[ipT];
        double mT = sqrt(mass2 + pT * pT);
        double mT_over_tau = mT / s->tau;
        for (int iphip = 0; iphip < nphi; iphip++) {
          double px = pT * cosphi[iphip], py = pT * sinphi[iphip];
          for (int iy = 0; iy < y_pts; iy++) {
            double y = (fl->dimension == 2) ? 0.0 : g->y[iy];
            double sum = 0.0, sum_abs = 0.0;
            for (int ieta = 0; ieta < eta_pts; ieta++) {
              double eta = (fl->dimension == 2) ? g->eta[ieta] : s->eta;
              double eta_weight = (fl->dimension == 2) ? g->eta_weight[ieta] : 1.0;
              double pt = mT * cosh(y - eta);
              double pn = mT_over_tau * sinh(y - eta);
              double tau2_pn = s->tau2 * pn;
              double pdotdsigma = eta_weight * (pt * s->dat + px * s->dax + py * s->day + pn * s->dan);
              if (fl->outflow && pdotdsigma <= 0.0) continue;
              double pdotu = pt * s->ut - px * s->ux - py * s->uy - tau2_pn * s->un;
              double feq = 1.0 / (exp(pdotu / s->T - chem) + sign);
              double mag = 0.0;           /* |df_shear| + |df_bulk| + |df_diff| built from absolute values of every term */
              double feqbar = 1.0 - sign * feq;
              double pimunu_pmu_pnu = s->pitt * pt * pt + s->pixx * px * px + s->piyy * py * py + s->pinn * tau2_pn * tau2_pn
                + 2.0 * (-(s->pitx * px + s->pity * py) * pt + s->pixy * px * py + tau2_pn * (s->pixn * px + s->piyn * py - s->pitn * pt));
              double Vmu_pmu = s->Vt * pt - s->Vx * px - s->Vy * py - s->Vn * tau2_pn;
              double df;
              if (fl->df_mode == 1) {
                double df_shear = s->shear_coeff * pimunu_pmu_pnu;
                double df_bulk = (s->bulk0_coeff * mass2 + (s->bulk1_coeff * baryon + s->bulk2_coeff * pdotu) * pdotu) * s->bulkPi;
                double df_diff = (s->df.c3 * baryon + s->df.c4 * pdotu) * Vmu_pmu;
                df = feqbar * (df_shear + df_bulk + df_diff);
                mag = fabs(s->shear_coeff) * (fabs(s->pitt) * pt * pt + fabs(s->pixx) * px * px + fabs(s->piyy) * py * py + fabs(s->pinn) * tau2_pn * tau2_pn
                      + 2.0 * ((fabs(s->pitx * px) + fabs(s->pity * py)) * pt + fabs(s->pixy * px * py) + fabs(tau2_pn) * (fabs(s->pixn * px) + fabs(s->piyn * py) + fabs(s->pitn) * pt)))
                      + (fabs(s->bulk0_coeff) * mass2 + fabs(s->bulk2_coeff) * pdotu * pdotu) * fabs(s->bulkPi) + fabs(df_diff);
              } else {
                double df_shear = s->shear_coeff * pimunu_pmu_pnu / pdotu;
                double df_bulk = (s->bulk0_coeff * pdotu + s->bulk1_coeff * baryon + s->bulk2_coeff * (pdotu - mass2 / pdotu)) * s->bulkPi;
                double df_diff = (s->baryon_enthalpy_ratio - baryon / pdotu) * Vmu_pmu / s->df.betaV;
                df = feqbar * (df_shear + df_bulk + df_diff);
                mag = fabs(s->shear_coeff) / pdotu * (fabs(s->pitt) * pt * pt + fabs(s->pixx) * px * px + fabs(s->piyy) * py * py + fabs(s->pinn) * tau2_pn * tau2_pn
                      + 2.0 * ((fabs(s->pitx * px) + fabs(s->pity * py)) * pt + fabs(s->pixy * px * py) + fabs(tau2_pn) * (fabs(s->pixn * px) + fabs(s->piyn * py) + fabs(s->pitn) * pt)))
                      + (fabs(s->bulk0_coeff) * pdotu + fabs(s->bulk2_coeff) * (pdotu + mass2 / pdotu)) * fabs(s->bulkPi) + fabs(df_diff);
              }
              int clamped = 0;
              if (fl->regulate_deltaf) { clamped = (df <= -1.0 || df >= 1.0); df = fmax(-1.0, fmin(df, 1.0)); }
              double f = feq * (1.0 + df);
              sum += (pdotdsigma * f);
              /* magnitude of the terms that make up f: its rounding noise is a few ulp of this, however small f itself is */
              sum_abs += fabs(pdotdsigma) * feq * (1.0 + (clamped ? 1.0 : fabs(feqbar) * mag));
            }
            int64_t iS3D = (int64_t)ipart + (int64_t)npart * ((int64_t)ipT + (int64_t)npT * ((int64_t)iphip + (int64_t)nphi * (int64_t)iy));
            dN[iS3D] += (prefactor * degeneracy * sum);
            if (dN_abs) dN_abs[iS3D] += (prefactor * degeneracy * sum_abs);
          }
        }
      }
    }
  }
  free(cs); free(cosphi); free(sinphi);
  return skipped;
}

/* ---------------------------------------------------------------- N2: spacetime distributions, linear delta-f, :1000-1446 */
/* the (tau, r) bin of a cell and the histogram updates of :1381-1400 */
static void spacetime_bin(const cfo_spacetime_spec *spec, double tau, double x, double y, double value,
                          double *h_tau, double *h_r, double *h_taur)
{
  const double tau_width = (spec->tau_max - spec->tau_min) / (double)spec->tau_bins;
  const double r_width = (spec->r_max - spec->r_min) / (double)spec->r_bins;
  double r = sqrt(x * x + y * y);
  int itau = (int)floor((tau - spec->tau_min) / tau_width);
  int ir = (int)floor((r - spec->r_min) / r_width);
  if (itau >= 0 && itau < spec->tau_bins) {
    h_tau[itau] += value;
    if (ir >= 0 && ir < spec->r_bins) h_taur[(int64_t)itau * spec->r_bins + ir] += value;
  }
  if (ir >= 0 && ir < spec->r_bins) h_r[ir] += value;
}

int64_t cfo_spacetime_vh(const cfo_flags *fl, const cfo_cells *c, const cfo_species *sp, const cfo_grid *g,
                         const cfo_df_tables *tab, const cfo_spacetime_spec *spec,
                         double *dN_tau, double *dN_r, double *dN_taur, double *dN_dydeta, double *dN_dy)
{
  if (fl->df_mode != 1 && fl->df_mode != 2) return -1;
  if (fl->include_baryon) return -2;
  const double prefactor = pow(2.0 * M_PI * hbarC, -3);
  const int npart = sp->n, npT = g->n_pT, nphi = g->n_phi;
  int y_pts, eta_pts; grid_dims(fl, g, &y_pts, &eta_pts);
  const int64_t n = c->n_cells;
  spline_cache sc; cache_build(&sc, tab);
  cell_setup *cs = (cell_setup *)malloc(sizeof(cell_setup) * (size_t)(n > 0 ? n : 1));
  int64_t skipped = 0; int bad = 0;
  for (int64_t i = 0; i < n; i++) {                      /* per-cell set-up :1146-1283, identical to :118-242 */
    cell_setup *s = &cs[i];
    cell_common(fl, c, i, s);
    if (s->skip) { skipped++; continue; }
    if (df_eval(&sc, fl->df_mode, s->T, s->E, s->P, s->bulkPi, &s->df)) bad = 1;
    if (fl->df_mode == 1) {
      s->shear_coeff = 0.5 / (s->T * s->T * (s->E + s->P));
      s->bulk0_coeff = s->df.c0 - s->df.c2;
      s->bulk1_coeff = s->df.c1;
      s->bulk2_coeff = 4.0 * s->df.c2 - s->df.c0;
    } else {
      s->shear_coeff = 0.5 / (s->df.betapi * s->T);
      s->bulk0_coeff = s->df.F / (s->T * s->T * s->df.betabulk);
      s->bulk1_coeff = s->df.G / s->df.betabulk;
      s->bulk2_coeff = 1.0 / (3.0 * s->T * s->df.betabulk);
    }
  }
  cache_free(&sc);
  if (bad) { free(cs); return -3; }
  double *cosphi = (double *)malloc(sizeof(double) * nphi), *sinphi = (double *)malloc(sizeof(double) * nphi);
  for (int k = 0; k < nphi; k++) { cosphi[k] = cos(g->phi[k]); sinphi[k] = sin(g->phi[k]); }

#pragma omp parallel for schedule(dynamic, 1)
  for (int ipart = 0; ipart < npart; ipart++) {
    const double mass = sp->mass[ipart], mass2 = mass * mass, sign = sp->sign[ipart];
    const double degeneracy = sp->degeneracy[ipart], baryon = sp->baryon[ipart];
    double *h_tau = dN_tau + (int64_t)ipart * spec->tau_bins, *h_r = dN_r + (int64_t)ipart * spec->r_bins;
    double *h_taur = dN_taur + (int64_t)ipart * spec->tau_bins * spec->r_bins;
    double *h_eta = dN_dydeta + (int64_t)ipart * eta_pts;
    for (int64_t icell = 0; icell < n; icell++) {
      const cell_setup *s = &cs[icell];
      if (s->skip) continue;
      const double chem = baryon * s->alphaB;
      double dN_dy_cell = 0.0;
      for (int ipT = 0; ipT < npT; ipT++) {
        const double pT = g->pT[ipT], mT = sqrt(mass2 + pT * pT), mT_over_tau = mT / s->tau;
        const double pT_weight = spec->pT_weight[ipT];
        for (int iphip = 0; iphip < nphi; iphip++) {
          const double px = pT * cosphi[iphip], py = pT * sinphi[iphip], phi_weight = g->phi_weight[iphip];
          for (int iy = 0; iy < y_pts; iy++) {           /* 3+1D: every y point is summed, without y weights (:1299) */
            const double y = (fl->dimension == 2) ? 0.0 : g->y[iy];
            double eta_sum = 0.0;
            for (int ieta = 0; ieta < eta_pts; ieta++) {
              const double eta = (fl->dimension == 2) ? g->eta[ieta] : s->eta;
              const double eta_weight = (fl->dimension == 2) ? g->eta_weight[ieta] : 1.0;
              const double pt = mT * cosh(y - eta), pn = mT_over_tau * sinh(y - eta), tau2_pn = s->tau2 * pn;
              const double pdotdsigma = eta_weight * (pt * s->dat + px * s->dax + py * s->day + pn * s->dan);
              if (fl->outflow && pdotdsigma <= 0.0) continue;
              const double pdotu = pt * s->ut - px * s->ux - py * s->uy - tau2_pn * s->un;
              const double feq = 1.0 / (exp(pdotu / s->T - chem) + sign);
              const double feqbar = 1.0 - sign * feq;
              const double pimunu_pmu_pnu = s->pitt * pt * pt + s->pixx * px * px + s->piyy * py * py + s->pinn * tau2_pn * tau2_pn
                + 2.0 * (-(s->pitx * px + s->pity * py) * pt + s->pixy * px * py + tau2_pn * (s->pixn * px + s->piyn * py - s->pitn * pt));
              const double Vmu_pmu = s->Vt * pt - s->Vx * px - s->Vy * py - s->Vn * tau2_pn;
              double df;
              if (fl->df_mode == 1) {
                const double df_shear = s->shear_coeff * pimunu_pmu_pnu;
                const double df_bulk = (s->bulk0_coeff * mass2 + (s->bulk1_coeff * baryon + s->bulk2_coeff * pdotu) * pdotu) * s->bulkPi;
                const double df_diff = (s->df.c3 * baryon + s->df.c4 * pdotu) * Vmu_pmu;
                df = feqbar * (df_shear + df_bulk + df_diff);
              } else {
                const double df_shear = s->shear_coeff * pimunu_pmu_pnu / pdotu;
                const double df_bulk = (s->bulk0_coeff * pdotu + s->bulk1_coeff * baryon + s->bulk2_coeff * (pdotu - mass2 / pdotu)) * s->bulkPi;
                const double df_diff = (s->baryon_enthalpy_ratio - baryon / pdotu) * Vmu_pmu / s->df.betaV;
                df = feqbar * (df_shear + df_bulk + df_diff);
              }
              if (fl->regulate_deltaf) df = fmax(-1.0, fmin(df, 1.0));
              const double f = feq * (1.0 + df);
              eta_sum += (pdotdsigma * f);
              h_eta[ieta] += (pT_weight * phi_weight * prefactor * degeneracy * pdotdsigma * f / eta_weight);
            }
            dN_dy_cell += (pT_weight * phi_weight * prefactor * degeneracy * eta_sum);
          }
        }
      }
      dN_dy[ipart] += dN_dy_cell;
      spacetime_bin(spec, s->tau, spec->x[icell], spec->y[icell], dN_dy_cell, h_tau, h_r, h_taur);
    }
  }
  free(cs); free(cosphi); free(sinphi);
  return skipped;
}

/* ---------------------------------------------------------------- a2: modified-equilibrium kernel, :396-996 */
typedef struct {
  double Xt, Xx, Xy, Xn, Yx, Yy, Zt, Zn;
} milne_basis;

/* Milne_Basis ctor, viscous_correction.cpp:10-29 */
static void milne(milne_basis *b, double ut, double ux, double uy, double un, double uperp, double utperp, double tau)
{
  double sinhL = tau * un / utperp, coshL = ut / utperp;
  b->Xt = uperp * coshL; b->Zt = sinhL; b->Xn = uperp * sinhL / tau; b->Zn = coshL / tau;
  b->Xx = 1.0; b->Yx = 0.0; b->Xy = 0.0; b->Yy = 1.0;
  if (uperp > 1.e-5) { b->Xx = utperp * ux / uperp; b->Yx = -uy / uperp; b->Xy = utperp * uy / uperp; b->Yy = ux / uperp; }
}

/* 3x3 inverse by LU with partial pivoting (gsl_linalg_LU_decomp / LU_invert as used at :689-707) */
static void lu_inverse3(const double A_in[9], double inv[9])
{
  double A[9]; int p[3] = {0, 1, 2};
  memcpy(A, A_in, sizeof(A));
  for (int j = 0; j < 2; j++) {
    double max = fabs(A[j * 3 + j]); int ip = j;
    for (int i = j + 1; i < 3; i++) { double a = fabs(A[i * 3 + j]); if (a > max) { max = a; ip = i; } }
    if (ip != j) { for (int k = 0; k < 3; k++) { double t = A[j * 3 + k]; A[j * 3 + k] = A[ip * 3 + k]; A[ip * 3 + k] = t; } int t = p[j]; p[j] = p[ip]; p[ip] = t; }
    double ajj = A[j * 3 + j];
    if (ajj != 0.0)
      for (int i = j + 1; i < 3; i++) {
        double aij = A[i * 3 + j] / ajj; A[i * 3 + j] = aij;
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

static void matvec3(const double A[9], const double x[3], double y[3])
{ for (int i = 0; i < 3; i++) { y[i] = 0.0; for (int j = 0; j < 3; j++) y[i] += A[i * 3 + j] * x[j]; } }

typedef struct {
  cell_setup s;
  milne_basis b;
  double T_mod, alphaB_mod, detA, eta_scale;
  double A[9], Ainv[9];
  int breaks_down;
  double neq_fact, dn_fact, J20_fact, N10_fact, nmod_fact;
} feqmod_setup;

/* One restatement serves both feqmod routines.  spec == NULL: calculate_dN_ptdptdphidy_feqmod (:396-996), result in dN.
 * spec != NULL: calculate_dN_dX_feqmod (:1449-2135), result in the histograms; that routine differs from the first in
 *   - the Jonah bulk-pressure clamp uses <= / >= (:1709-1710 vs :591-592),
 *   - eta_scale = detA whenever detA > detA_min in 2+1D, without the detA < 1 condition (:1850-1853 vs :729),
 *   - the renormalisation NaN/Inf test is made on renorm / detA (:1887 vs :773),
 *   - the "narrow" per-rapidity breakdown is commented out (:1927-1935 vs :813-819),
 *   - p.dsigma = eta_weight * (pt dat + px dax + py day + pn dan) in both branches (:1948, :2001 vs :832, :883),
 *   - its `breakdown` counter runs inside the species loop (npart times the cell count); the cell count is returned here. */
static int64_t feqmod_core(const cfo_flags *fl, const cfo_cells *c, const cfo_species *sp, const cfo_grid *g,
                           const cfo_df_tables *tab, const cfo_laguerre *gla, double *dN, int64_t *breakdown_out,
                           const cfo_spacetime_spec *spec, double *dN_tau, double *dN_r, double *dN_taur, double *dN_dydeta, double *dN_dy)
{
  const int dX = (spec != NULL);
  if (fl->df_mode != 3 && fl->df_mode != 4) return -1;
  if (fl->include_baryon) return -2;
  const double prefactor = pow(2.0 * M_PI * hbarC, -3);
  const double two_pi2_hbarC3 = 2.0 * pow(M_PI, 2) * pow(hbarC, 3);
  const int npart = sp->n, npT = g->n_pT, nphi = g->n_phi, DF_MODE = fl->df_mode;
  int y_pts, eta_pts; grid_dims(fl, g, &y_pts, &eta_pts);
  const int64_t n = c->n_cells;
  const double detA_min = fl->deta_min;
  spline_cache sc; cache_build(&sc, tab);
  feqmod_setup *cs = (feqmod_setup *)malloc(sizeof(feqmod_setup) * (size_t)(n > 0 ? n : 1));
  int64_t skipped = 0, breakdown = 0; int bad = 0;
  for (int64_t i = 0; i < n; i++) {
    feqmod_setup *f = &cs[i]; cell_setup *s = &f->s;
    cell_common(fl, c, i, s);
    if (s->skip) { skipped++; continue; }
    double uperp = sqrt(s->ux * s->ux + s->uy * s->uy), utperp = sqrt(1.0 + s->ux * s->ux + s->uy * s->uy);
    if (DF_MODE == 4) {                                              /* :588-594 */
      double mx = tab->bulkPi_over_Peq_max;
      if (dX) {
        if (s->bulkPi <= -s->P) s->bulkPi = -(1.0 - 1.e-5) * s->P;
        else if (s->bulkPi / s->P >= mx) s->bulkPi = s->P * (mx - 1.e-5);
      } else {
        if (s->bulkPi < -s->P) s->bulkPi = -(1.0 - 1.e-5) * s->P;
        else if (s->bulkPi / s->P > mx) s->bulkPi = s->P * (mx - 1.e-5);
      }
    }
    if (df_eval(&sc, DF_MODE, s->T, s->E, s->P, s->bulkPi, &s->df)) bad = 1;
    const cfo_dfcoef *df = &s->df;
    milne(&f->b, s->ut, s->ux, s->uy, s->un, uperp, utperp, s->tau);
    const milne_basis *b = &f->b;
    double tau2 = s->tau2;
    /* Shear_Stress::boost_pimunu_to_lrf, viscous_correction.cpp:121-142 */
    double Xt = b->Xt, Xx = b->Xx, Xy = b->Xy, Xn = b->Xn, Yx = b->Yx, Yy = b->Yy, Zt = b->Zt, Zn = b->Zn;
    double pitt = s->pitt, pitx = s->pitx, pity = s->pity, pitn = s->pitn, pixx = s->pixx, pixy = s->pixy, pixn = s->pixn, piyy = s->piyy, piyn = s->piyn, pinn = s->pinn;
    double pixx_LRF = pitt * Xt * Xt + pixx * Xx * Xx + piyy * Xy * Xy + tau2 * tau2 * pinn * Xn * Xn
      + 2.0 * (-Xt * (pitx * Xx + pity * Xy) + pixy * Xx * Xy + tau2 * Xn * (pixn * Xx + piyn * Xy - pitn * Xt));
    double pixy_LRF = Yx * (-pitx * Xt + pixx * Xx + pixy * Xy + tau2 * pixn * Xn) + Yy * (-pity * Xt + pixy * Xx + piyy * Xy + tau2 * piyn * Xn);
    double pixz_LRF = Zt * (pitt * Xt - pitx * Xx - pity * Xy - tau2 * pitn * Xn) - tau2 * Zn * (pitn * Xt - pixn * Xx - piyn * Xy - tau2 * pinn * Xn);
    double piyy_LRF = pixx * Yx * Yx + 2.0 * pixy * Yx * Yy + piyy * Yy * Yy;
    double piyz_LRF = -Zt * (pitx * Yx + pity * Yy) + tau2 * Zn * (pixn * Yx + piyn * Yy);
    double pizz_LRF = -(pixx_LRF + piyy_LRF);
    f->T_mod = s->T; f->alphaB_mod = s->alphaB;
    if (DF_MODE == 3) { f->T_mod = s->T + s->bulkPi * df->F / df->betabulk; f->alphaB_mod = s->alphaB + s->bulkPi * df->G / df->betabulk; }
    s->shear_coeff = 0.5 / (df->betapi * s->T);                        /* :641-644 */
    s->bulk0_coeff = df->F / (s->T * s->T * df->betabulk);
    s->bulk1_coeff = df->G / df->betabulk;
    s->bulk2_coeff = 1.0 / (3.0 * s->T * df->betabulk);
    double shear_mod = 0.5 / df->betapi;
    double bulk_mod = s->bulkPi / (3.0 * df->betabulk);
    if (DF_MODE == 4) bulk_mod = df->lambda;
    double Axx = 1.0 + pixx_LRF * shear_mod + bulk_mod, Axy = pixy_LRF * shear_mod, Axz = pixz_LRF * shear_mod;
    double Ayy = 1.0 + piyy_LRF * shear_mod + bulk_mod, Ayz = piyz_LRF * shear_mod, Azz = 1.0 + pizz_LRF * shear_mod + bulk_mod;
    f->detA = Axx * (Ayy * Azz - Ayz * Ayz) - Axy * (Axy * Azz - Ayz * Axz) + Axz * (Axy * Ayz - Ayy * Axz);
    double A[9] = {Axx, Axy, Axz, Axy, Ayy, Ayz, Axz, Ayz, Azz};
    memcpy(f->A, A, sizeof(A));
    lu_inverse3(A, f->Ainv);
    f->neq_fact = s->T * s->T * s->T / two_pi2_hbarC3;
    f->dn_fact = s->bulkPi / df->betabulk;
    f->J20_fact = s->T * f->neq_fact;
    f->N10_fact = f->neq_fact;
    f->nmod_fact = f->T_mod * f->T_mod * f->T_mod / two_pi2_hbarC3;
    /* does_feqmod_breakdown, emissionfunction.cpp:109-150 */
    f->breaks_down = 0;
    if (DF_MODE == 3) {
      double mbar_pion0 = fl->mass_pion0 / s->T;
      double neq_pion0 = f->neq_fact * gauss_thermal(neq_int, gla->root1, gla->weight1, gla->n_points, mbar_pion0, 0., 0., -1.);
      double J20_pion0 = f->J20_fact * gauss_thermal(J20_int, gla->root2, gla->weight2, gla->n_points, mbar_pion0, 0., 0., -1.);
      double dn_pion0 = s->bulkPi * (neq_pion0 + J20_pion0 * df->F / s->T / s->T) / df->betabulk;
      double nlinear_pion0 = neq_pion0 + dn_pion0;
      if (f->detA <= detA_min || nlinear_pion0 < 0.0) f->breaks_down = 1;
    }
    if (f->breaks_down) breakdown++;
    f->eta_scale = 1.0;
    if (f->detA > detA_min && (dX || f->detA < 1.0) && fl->dimension == 2) f->eta_scale = f->detA;
  }
  cache_free(&sc);
  if (bad) { free(cs); return -3; }
  double *cosphi = (double *)malloc(sizeof(double) * nphi), *sinphi = (double *)malloc(sizeof(double) * nphi);
  for (int k = 0; k < nphi; k++) { cosphi[k] = cos(g->phi[k]); sinphi[k] = sin(g->phi[k]); }

#pragma omp parallel for schedule(dynamic, 1)
  for (int ipart = 0; ipart < npart; ipart++) {
    double mass = sp->mass[ipart], mass2 = mass * mass, sign = sp->sign[ipart];
    double degeneracy = sp->degeneracy[ipart], baryon = sp->baryon[ipart];
    double *h_tau = dX ? dN_tau + (int64_t)ipart * spec->tau_bins : NULL, *h_r = dX ? dN_r + (int64_t)ipart * spec->r_bins : NULL;
    double *h_taur = dX ? dN_taur + (int64_t)ipart * spec->tau_bins * spec->r_bins : NULL;
    double *h_eta = dX ? dN_dydeta + (int64_t)ipart * eta_pts : NULL;
    for (int64_t icell = 0; icell < n; icell++) {
      const feqmod_setup *fs = &cs[icell]; const cell_setup *s = &fs->s; const milne_basis *b = &fs->b;
      if (s->skip) continue;
      double dN_dy_cell = 0.0;
      const cfo_dfcoef *df = &s->df;
      double chem = baryon * s->alphaB, chem_mod = baryon * fs->alphaB_mod;
      double renorm = 1.0;
      if (fl->include_bulk) {                                            /* :747-771 */
        if (DF_MODE == 3) {
          double mbar = mass / s->T, mbar_mod = mass / fs->T_mod;
          double neq = fs->neq_fact * degeneracy * gauss_thermal(neq_int, gla->root1, gla->weight1, gla->n_points, mbar, s->alphaB, baryon, sign);
          double N10 = baryon * fs->N10_fact * degeneracy * gauss_thermal(J10_int, gla->root1, gla->weight1, gla->n_points, mbar, s->alphaB, baryon, sign);
          double J20 = fs->J20_fact * degeneracy * gauss_thermal(J20_int, gla->root2, gla->weight2, gla->n_points, mbar, s->alphaB, baryon, sign);
          double n_linear = neq + fs->dn_fact * (neq + N10 * df->G + J20 * df->F / s->T / s->T);
          double n_mod = fs->nmod_fact * degeneracy * gauss_thermal(neq_int, gla->root1, gla->weight1, gla->n_points, mbar_mod, fs->alphaB_mod, baryon, sign);
          renorm = n_linear / n_mod;
        } else renorm = df->z;
      }
      if (dX ? (isnan(renorm / fs->detA) || isinf(renorm / fs->detA)) : (isnan(renorm) || isinf(renorm))) continue;   /* :773-778, :1887-1891 */
      if (fl->dimension == 3) renorm /= fs->detA;
      for (int ipT = 0; ipT < npT; ipT++) {
        double pT = g->pT[ipT];
        double mT = sqrt(mass2 + pT * pT);
        double mT_over_tau = mT / s->tau;
        const double w_pT_phi_base = dX ? spec->pT_weight[ipT] : 0.0;
        for (int iphip = 0; iphip < nphi; iphip++) {
          double px = pT * cosphi[iphip], py = pT * sinphi[iphip];
          const double pT_weight = w_pT_phi_base, phi_weight = g->phi_weight ? g->phi_weight[iphip] : 0.0;
          for (int iy = 0; iy < y_pts; iy++) {
            double y = (fl->dimension == 2) ? 0.0 : g->y[iy];
            double sum = 0.0;
            for (int ieta = 0; ieta < eta_pts; ieta++) {
              double eta = (fl->dimension == 2) ? g->eta[ieta] : s->eta;
              double eta_weight = (fl->dimension == 2) ? g->eta_weight[ieta] : 1.0;
              int narrow = 0;
              if (!dX && fl->dimension == 3 && !fs->breaks_down) { if (fs->detA < 0.01 && fabs(y - eta) < fs->detA) narrow = 1; }
              double pdotdsigma, f = 0.0;
              if (fs->breaks_down || narrow) {                           /* :825-877 */
                double pt = mT * cosh(y - eta);
                double pn = mT_over_tau * sinh(y - eta);
                double tau2_pn = s->tau2 * pn;
                pdotdsigma = dX ? eta_weight * (pt * s->dat + px * s->dax + py * s->day + pn * s->dan)
                                : eta_weight * (pt * s->dat + px * s->dax + py * s->day) + pn * s->dan;
                if (fl->outflow && pdotdsigma <= 0.0) continue;
                double pdotu = pt * s->ut - px * s->ux - py * s->uy - tau2_pn * s->un;
                double pimunu_pmu_pnu = s->pitt * pt * pt + s->pixx * px * px + s->piyy * py * py + s->pinn * tau2_pn * tau2_pn
                  + 2.0 * (-(s->pitx * px + s->pity * py) * pt + s->pixy * px * py + tau2_pn * (s->pixn * px + s->piyn * py - s->pitn * pt));
                if (DF_MODE == 3) {
                  double feq = 1.0 / (exp(pdotu / s->T - chem) + sign);
                  double feqbar = 1.0 - sign * feq;
                  double Vmu_pmu = s->Vt * pt - s->Vx * px - s->Vy * py - s->Vn * tau2_pn;
                  double df_shear = s->shear_coeff * pimunu_pmu_pnu / pdotu;
                  double df_bulk = (s->bulk0_coeff * pdotu + s->bulk1_coeff * baryon + s->bulk2_coeff * (pdotu - mass2 / pdotu)) * s->bulkPi;
                  double df_diff = (s->baryon_enthalpy_ratio - baryon / pdotu) * Vmu_pmu / df->betaV;
                  double dfv = feqbar * (df_shear + df_bulk + df_diff);
                  if (fl->regulate_deltaf) dfv = fmax(-1.0, fmin(dfv, 1.0));
                  f = feq * (1.0 + dfv);
                } else {
                  double feq = 1.0 / (exp(pdotu / s->T) + sign);
                  double feqbar = 1.0 - sign * feq;
                  double df_shear = feqbar * s->shear_coeff * pimunu_pmu_pnu / pdotu;
                  double df_bulk = df->delta_z - 3.0 * df->delta_lambda + feqbar * df->delta_lambda * (pdotu - mass2 / pdotu) / s->T;
                  double dfv = df_shear + df_bulk;
                  if (fl->regulate_deltaf) dfv = fmax(-1.0, fmin(dfv, 1.0));
                  f = feq * (1.0 + dfv);
                }
              } else {                                                   /* :878-928 */
                double pt = mT * cosh(y - fs->eta_scale * eta);
                double pn = mT_over_tau * sinh(y - fs->eta_scale * eta);
                double tau2_pn = s->tau2 * pn;
                pdotdsigma = dX ? eta_weight * (pt * s->dat + px * s->dax + py * s->day + pn * s->dan)
                                : eta_weight * (pt * s->dat + px * s->dax + py * s->day) + pn * s->dan;
                if (fl->outflow && pdotdsigma <= 0.0) continue;
                double pLRF[3] = {-b->Xt * pt + b->Xx * px + b->Xy * py + b->Xn * tau2_pn, b->Yx * px + b->Yy * py, -b->Zt * pt + b->Zn * tau2_pn};
                double pmod[3], pmod_prev[3], pprev[3], dp3[3], dpmod[3];
                matvec3(fs->Ainv, pLRF, pmod);
                for (int it = 0; it < 5; it++) {
                  memcpy(pmod_prev, pmod, sizeof(pmod));
                  matvec3(fs->A, pmod_prev, pprev);
                  for (int k = 0; k < 3; k++) dp3[k] = pLRF[k] - pprev[k];
                  double dp = sqrt(dp3[0] * dp3[0] + dp3[1] * dp3[1] + dp3[2] * dp3[2]);
                  if (dp <= 1.e-16) break;
                  matvec3(fs->Ainv, dp3, dpmod);
                  for (int k = 0; k < 3; k++) pmod[k] = pmod_prev[k] + dpmod[k];
                }
                double E_mod = sqrt(mass2 + pmod[0] * pmod[0] + pmod[1] * pmod[1] + pmod[2] * pmod[2]);
                f = fabs(renorm) / (exp(E_mod / fs->T_mod - chem_mod) + sign);
              }
              sum += (pdotdsigma * f);
              if (dX) h_eta[ieta] += (pT_weight * phi_weight * prefactor * degeneracy * pdotdsigma * f / eta_weight);
            }
            if (dX) { dN_dy_cell += (pT_weight * phi_weight * prefactor * degeneracy * sum); continue; }
            int64_t iS3D = (int64_t)ipart + (int64_t)npart * ((int64_t)ipT + (int64_t)npT * ((int64_t)iphip + (int64_t)nphi * (int64_t)iy));
            dN[iS3D] += (prefactor * degeneracy * sum);
          }
        }
      }
      if (dX) {
        dN_dy[ipart] += dN_dy_cell;
        spacetime_bin(spec, s->tau, spec->x[icell], spec->y[icell], dN_dy_cell, h_tau, h_r, h_taur);
      }
    }
  }
  free(cs); free(cosphi); free(sinphi);
  if (breakdown_out) *breakdown_out = breakdown;
  return skipped;
}

int64_t cfo_smooth_feqmod(const cfo_flags *fl, const cfo_cells *c, const cfo_species *sp, const cfo_grid *g,
                          const cfo_df_tables *tab, const cfo_laguerre *gla, double *dN, int64_t *breakdown_out)
{
  return feqmod_core(fl, c, sp, g, tab, gla, dN, breakdown_out, NULL, NULL, NULL, NULL, NULL, NULL);
}

int64_t cfo_spacetime_feqmod(const cfo_flags *fl, const cfo_cells *c, const cfo_species *sp, const cfo_grid *g,
                             const cfo_df_tables *tab, const cfo_laguerre *gla, const cfo_spacetime_spec *spec,
                             double *dN_tau, double *dN_r, double *dN_taur, double *dN_dydeta, double *dN_dy, int64_t *breakdown_out)
{
  if (!spec) return -1;
  return feqmod_core(fl, c, sp, g, tab, gla, NULL, breakdown_out, spec, dN_tau, dN_r, dN_taur, dN_dydeta, dN_dy);
}

/* ---------------------------------------------------------------- N4: sampler mean yield */
static double J11_int(double pbar, double mbar, double alphaB, double baryon, double sign)
{ double Ebar = sqrt(pbar * pbar + mbar * mbar); double q = exp(Ebar - baryon * alphaB) + sign;
  return pbar * pbar * pbar / (Ebar * Ebar) * exp(pbar + Ebar - baryon * alphaB) / (q * q); }
static double J30_int(double pbar, double mbar, double alphaB, double baryon, double sign)
{ double Ebar = sqrt(pbar * pbar + mbar * mbar); double q = exp(Ebar - baryon * alphaB) + sign;
  return Ebar * Ebar / pbar * exp(pbar + Ebar - baryon * alphaB) / (q * q); }
static double J31_int(double pbar, double mbar, double alphaB, double baryon, double sign)
{ double Ebar = sqrt(pbar * pbar + mbar * mbar); double q = exp(Ebar - baryon * alphaB) + sign;
  return pbar * exp(pbar + Ebar - baryon * alphaB) / (q * q); }

/* Deltaf_Data::compute_particle_densities, deltafReader.cpp:536-650: densities at the surface-average T, E, P (avg5 = T, E, P,
 * muB, nB; muB = nB = 0 here).  root3 / weight3 are the alpha = 3 Gauss-Laguerre nodes (df_mode 1 only). */
int cfo_particle_densities(int n, const double *mass, const double *degeneracy, const double *baryon, const double *sign,
                           const double *avg5, int df_mode, const cfo_df_tables *tab, const cfo_laguerre *gla,
                           const double *root3, const double *weight3, double *neq_out, double *bulk_out, double *diff_out)
{
  const double two_pi2_hbarC3 = 2.0 * pow(M_PI, 2) * pow(hbarC, 3);
  const double T = avg5[0], E = avg5[1], P = avg5[2], muB = avg5[3], nB = avg5[4];
  cfo_dfcoef df;
  if (cfo_df_coefficients(tab, df_mode, T, E, P, 0.0, &df)) return -3;
  const double alphaB = muB / T, baryon_enthalpy_ratio = nB / (E + P);
  const int pts = gla->n_points;
  for (int i = 0; i < n; i++) {
    const double m = mass[i], g = degeneracy[i], b = baryon[i], sg = sign[i], mbar = m / T;
    const double neq_fact = g * pow(T, 3) / two_pi2_hbarC3;
    const double neq = neq_fact * gauss_thermal(neq_int, gla->root1, gla->weight1, pts, mbar, alphaB, b, sg);
    double dn_bulk = 0.0, dn_diff = 0.0;
    if (df_mode == 1) {
      const double J10_fact = g * pow(T, 3) / two_pi2_hbarC3, J20_fact = g * pow(T, 4) / two_pi2_hbarC3;
      const double J30_fact = g * pow(T, 5) / two_pi2_hbarC3, J31_fact = g * pow(T, 5) / two_pi2_hbarC3 / 3.0;
      const double J10 = J10_fact * gauss_thermal(J10_int, gla->root1, gla->weight1, pts, mbar, alphaB, b, sg);
      const double J20 = J20_fact * gauss_thermal(J20_int, gla->root2, gla->weight2, pts, mbar, alphaB, b, sg);
      const double J30 = J30_fact * gauss_thermal(J30_int, root3, weight3, pts, mbar, alphaB, b, sg);
      const double J31 = J31_fact * gauss_thermal(J31_int, root3, weight3, pts, mbar, alphaB, b, sg);
      dn_bulk = ((df.c0 - df.c2) * m * m * J10 + df.c1 * b * J20 + (4.0 * df.c2 - df.c0) * J30);
      dn_diff = b * df.c3 * neq * T + df.c4 * J31;
    } else if (df_mode == 2 || df_mode == 3) {
      const double J10_fact = g * pow(T, 3) / two_pi2_hbarC3, J11_fact = g * pow(T, 3) / two_pi2_hbarC3 / 3.0;
      const double J20_fact = g * pow(T, 4) / two_pi2_hbarC3;
      const double J10 = J10_fact * gauss_thermal(J10_int, gla->root1, gla->weight1, pts, mbar, alphaB, b, sg);
      const double J11 = J11_fact * gauss_thermal(J11_int, gla->root1, gla->weight1, pts, mbar, alphaB, b, sg);
      const double J20 = J20_fact * gauss_thermal(J20_int, gla->root2, gla->weight2, pts, mbar, alphaB, b, sg);
      dn_bulk = (neq + (b * J10 * df.G) + (J20 * df.F / pow(T, 2))) / df.betabulk;
      dn_diff = (neq * T * baryon_enthalpy_ratio - b * J11) / df.betaV;
    }
    neq_out[i] = neq; bulk_out[i] = dn_bulk; diff_out[i] = dn_diff;
  }
  return 0;
}

/* EmissionFunctionArray::calculate_total_yield, emissionfunction_sampling_kernels.cpp:653-831, without baryon diffusion
 * (V.dsigma = 0, so the reference's uninitialised ds_space never contributes): per cell and species
 *   df_mode 1-3: u.dsigma (n_eq + Pi dn_bulk);   df_mode 4: u.dsigma z(Pi/P) n_eq  (does_feqmod_breakdown is false for Jonah).
 * Returns the skipped-cell count, *Ntot gets the yield (times 2 y_cut in 2+1D). */
int64_t cfo_total_yield(const cfo_flags *fl, const cfo_cells *c, int n_species, const double *neq, const double *bulk,
                        const cfo_df_tables *tab, double y_cut, double *Ntot_out)
{
  spline_cache sc; cache_build(&sc, tab);
  double Ntot = 0.0; int64_t skipped = 0; int bad = 0;
  for (int64_t i = 0; i < c->n_cells; i++) {
    const double tau = c->tau[i], tau2 = tau * tau;
    const double ux = c->ux[i], uy = c->uy[i], un = c->un[i];
    const double ut = sqrt(1.0 + ux * ux + uy * uy + tau2 * un * un);
    const double udsigma = ut * c->dat[i] + ux * c->dax[i] + uy * c->day[i] + un * c->dan[i];
    if (udsigma <= 0.0) { skipped++; continue; }
    double bulkPi = fl->include_bulk ? c->bulkPi[i] : 0.0;
    const double P = c->P[i];
    cfo_dfcoef df; memset(&df, 0, sizeof(df));
    if (fl->df_mode == 4) {
      const double mx = tab->bulkPi_over_Peq_max;
      if (bulkPi <= -P) bulkPi = -(1.0 - 1.e-5) * P;
      else if (bulkPi / P >= mx) bulkPi = P * (mx - 1.e-5);
    }
    if (df_eval(&sc, fl->df_mode, c->T[i], c->E[i], P, bulkPi, &df)) bad = 1;
    const double ds_time = udsigma;                   /* Surface_Element_Vector::boost_dsigma_to_lrf, viscous_correction.cpp:76 */
    for (int s = 0; s < n_species; s++) {
      if (fl->df_mode == 4) Ntot += ds_time * df.z * neq[s];
      else Ntot += ds_time * (neq[s] + bulkPi * bulk[s]);
    }
  }
  cache_free(&sc);
  if (bad) return -3;
  if (fl->dimension == 2) Ntot *= (2.0 * y_cut);
  *Ntot_out = Ntot;
  return skipped;
}

/* ---------------------------------------------------------------- a3: anisotropic PL-matching kernel, :2140-2393 */
int64_t cfo_smooth_vah(const cfo_flags *fl, const cfo_cells *c, const cfo_species *sp, const cfo_grid *g, double *dN)
{
  const double prefactor = 1.0 / (8.0 * (M_PI * M_PI * M_PI)) / hbarC / hbarC / hbarC;
  const int npart = sp->n, npT = g->n_pT, nphi = g->n_phi;
  int y_pts, eta_pts; grid_dims(fl, g, &y_pts, &eta_pts);
  const int64_t n = c->n_cells;
  double delta_eta = (g->n_eta > 1) ? g->eta[1] - g->eta[0] : 0.0;        /* :2175 */
  double *cosphi = (double *)malloc(sizeof(double) * nphi), *sinphi = (double *)malloc(sizeof(double) * nphi);
  for (int k = 0; k < nphi; k++) { cosphi[k] = cos(g->phi[k]); sinphi[k] = sin(g->phi[k]); }

#pragma omp parallel for schedule(dynamic, 1)
  for (int ipart = 0; ipart < npart; ipart++) {
    double mass = sp->mass[ipart], mass2 = mass * mass, sign = sp->sign[ipart], degeneracy = sp->degeneracy[ipart];
    for (int64_t i = 0; i < n; i++) {
      double tau = c->tau[i], tau2 = tau * tau;
      double dat = c->dat[i], dax = c->dax[i], day = c->day[i], dan = c->dan[i];
      double ux = c->ux[i], uy = c->uy[i], un = c->un[i];
      double ut = sqrt(1.0 + ux * ux + uy * uy + tau2 * un * un);
      double u0 = sqrt(1.0 + ux * ux + uy * uy);
      double zt = tau * un / u0, zn = ut / (u0 * tau);
      double pitt = c->pitt[i], pitx = c->pitx[i], pity = c->pity[i], pitn = c->pitn[i], pixx = c->pixx[i], pixy = c->pixy[i];
      double pixn = c->pixn[i], piyy = c->piyy[i], piyn = c->piyn[i], pinn = c->pinn[i];
      double bulkPi = c->bulkPi[i];
      double Wx = c->Wx[i], Wy = c->Wy[i];
      double Wt = (ux * Wx + uy * Wy) * ut / (u0 * u0);
      double Wn = Wt * un / ut;
      double Lambda = c->Lambda[i], aL = c->aL[i];
      double c0 = c->c0[i], c1 = c->c1[i], c2 = c->c2[i], c3 = c->c3[i], c4 = c->c4[i];
      for (int ipT = 0; ipT < npT; ipT++) {
        double pT = g->pT[ipT];
        double mT = sqrt(mass2 + pT * pT);
        double mT_over_tau = mT / tau;
        for (int iphip = 0; iphip < nphi; iphip++) {
          double px = pT * cosphi[iphip], py = pT * sinphi[iphip];
          for (int iy = 0; iy < y_pts; iy++) {
            double y = (fl->dimension == 2) ? 0.0 : g->y[iy];
            double sum = 0.0;
            for (int ieta = 0; ieta < eta_pts; ieta++) {
              double eta = (fl->dimension == 2) ? g->eta[ieta] : c->eta[i];
              double eta_weight = (fl->dimension == 2) ? g->eta_weight[ieta] * delta_eta : 1.0;
              double pt = mT * cosh(y - eta);
              double pn = mT_over_tau * sinh(y - eta);
              double tau2_pn = tau2 * pn;
              double pdotdsigma = pt * dat + px * dax + py * day + pn * dan;
              double pdotu = pt * ut - px * ux - py * uy - tau2_pn * un;
              double pdotz = pt * zt - tau2_pn * zn;
              double xiL = 1.0 / (aL * aL) - 1.0;
              double Ea = sqrt(pdotu * pdotu + xiL * pdotz * pdotz);
              double fa = 1.0 / (exp(Ea / Lambda) + sign);
              double fabar = 1.0 - sign * fa;
              double df_shear = 0.0;
              if (fl->include_shear) {
                double Wmu_pmu_pz = pdotz * (Wt * pt - Wx * px - Wy * py - Wn * tau2_pn);
                double pimunu_pmu_pnu = pitt * pt * pt + pixx * px * px + piyy * py * py + pinn * tau2_pn * tau2_pn
                  + 2.0 * (-(pitx * px + pity * py) * pt + pixy * px * py + tau2_pn * (pixn * px + piyn * py - pitn * pt));
                df_shear = c3 * Wmu_pmu_pz + c4 * pimunu_pmu_pnu;
              }
              double df_bulk = 0.0;
              if (fl->include_bulk) df_bulk = (c0 * mass2 + c1 * pdotz * pdotz + c2 * pdotu * pdotu) * bulkPi;
              double df = df_shear + df_bulk;
              if (fl->regulate_deltaf) {
                double reg_df = fmax(-1.0, fmin(fabar * df, 1.0));
                sum += (eta_weight * pdotdsigma * fa * (1.0 + reg_df));
              } else sum += (eta_weight * pdotdsigma * fa * (1.0 + fabar * df));
            }
            int64_t iS3D = (int64_t)ipart + (int64_t)npart * ((int64_t)ipT + (int64_t)npT * ((int64_t)iphip + (int64_t)nphi * (int64_t)iy));
            dN[iS3D] += (prefactor * degeneracy * sum);
          }
        }
      }
    }
  }
  free(cosphi); free(sinphi);
  return 0;
}

/* ---------------------------------------------------------------- VAH helpers, arsenal.cpp:999-1066 */
double cfo_aL_fit(double x)
{
  static const double num[15] = {2.307660683188896e-22, 1.7179667824677117e-16, 7.2725449826862375e-12, 4.2846163672079405e-8,
    0.00004757224421671691, 0.011776118846199547, 0.7235583305942909, 11.582755440134724, 44.45243622597357, 12.673594148032494,
    -33.75866652773691, 8.04299287188939, 1.462901772148128, -0.6320131889637761, 0.048528166213735346};
  static const double den[15] = {5.595674409987461e-19, 8.059757191879689e-14, 1.2033043382301483e-9, 2.9819348588423508e-6,
    0.0015212379997299082, 0.18185453852532632, 5.466199358534425, 40.1581708710626, 44.38310108782752, -55.213789667214364,
    1.5449108423263358, 11.636087951096759, -4.005934533735304, 0.4703844693488544, -0.014599143701745957};
  double xp[15]; xp[0] = 1.0; xp[1] = x;
  for (int k = 2; k < 15; k++) xp[k] = xp[k - 1] * x;
  double a = num[0], b = den[0];
  for (int k = 1; k < 15; k++) { a += num[k] * xp[k]; b += den[k] * xp[k]; }
  return a / b;
}

double cfo_R200(double aL)
{
  double x = (1.0 / (aL * aL)) - 1.0, t200 = 0.0, delta = 0.01;
  if (x > delta) t200 = 1.0 + (1.0 + x) * atan(sqrt(x)) / sqrt(x);
  else if (x < -delta && x > -1.0) t200 = 1.0 + (1.0 + x) * atanh(sqrt(-x)) / sqrt(-x);
  else if (x >= -delta && x <= delta)
    t200 = 2.0 + x * (0.6666666666666667 + x * (-0.1333333333333333 + x * (0.05714285714285716 + x * (-0.031746031746031744 + x * (0.020202020202020193 +
      x * (-0.013986013986013984 + (0.010256410256410262 - 0.00784313725490196 * x) * x))))));
  else return NAN;
  return aL * t200;
}
