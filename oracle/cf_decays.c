/* oracle/cf_decays.c -- CPU restatement of iS3D's resonance-decay feed-down.  TEST INFRASTRUCTURE ONLY.
 *
 * Follows src/cpp/emissionfunction_resonance_decays.cpp of the reference operation by operation (file:line cited at each function),
 * quirks included.  Parity status: PINNED against the reference's own routine, compiled unmodified behind oracle/ref_decays_prefix.h
 * (the routine's first statements are a printf + exit(-1) left by its author, :126-129; see that header): tests/test_decays.py and
 * tests/golden/decays_*.npz.  The reference author flags the handling of MTmax in the interpolation as unfinished.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "cf_oracle.h"

#define GP 12
static const double GL_ROOT[GP] = {-0.98156063424672, -0.90411725637048, -0.76990267419431, -0.58731795428662, -0.3678314989982, -0.12523340851147,
   0.12523340851147, 0.36783149899818, 0.58731795428662, 0.76990267419431, 0.90411725637048, 0.98156063424672};
static const double GL_WEIGHT[GP] = {0.04717533638651, 0.1069393259953, 0.16007832854335, 0.20316742672307, 0.23349253653836, 0.2491470458134,
   0.2491470458134, 0.23349253653836, 0.20316742672307, 0.1600783285433, 0.10693932599532, 0.04717533638651};

typedef struct { double constant, slope; } mt_fit;

typedef struct {
  const cfo_particles *pdg;
  int n_chosen; const int32_t *chosen;            /* chosen_particles_sampling_table: pdg index of every chosen species */
  int n_pT, n_phi, n_y_tab, y_pts, dim;
  const double *pT, *phi, *y;
  double *dN;                                     /* [y][phi][pT][species] */
  double *logdN;                                  /* [y][phi][pT] of the current parent */
  mt_fit *fit;                                    /* [y_pts][n_phi] */
  double MTValues[512];
  int err;
} ctx_t;

/* particle_index, :59-79 */
static int particle_index(const cfo_particles *p, int mc_id, int *err)
{
  if (mc_id == 0) { *err = -11; return 0; }
  for (int i = 0; i < p->n_particles; i++) if (p->mcid[i] == mc_id) return i;
  *err = -12; return 0;
}
/* EmissionFunctionArray::particle_chosen_index, :82-97 */
static int particle_chosen_index(const ctx_t *c, int pdg_index, int *err)
{
  for (int i = 0; i < c->n_chosen; i++) if (c->chosen[i] == pdg_index) return i;
  *err = -13; return 0;
}

/* calculate_Q_factor, :99-121 */
static double q_factor(double mass_parent, double mass_1, double mass_2, double mass_3)
{
  static const double x_root[24] = {-0.99518721999702,-0.97472855597131,-0.93827455200273,-0.8864155270044,-0.8200019859739,-0.74012419157855,-0.64809365193698,-0.54542147138884,-0.43379350762605,-0.31504267969616,-0.19111886747362,-0.064056892862606,0.06405689286261,0.19111886747362,0.31504267969616,0.43379350762605,0.54542147138884,0.64809365193698,0.74012419157855,0.8200019859739,0.8864155270044,0.93827455200273,0.97472855597131,0.99518721999702};
  static const double x_weight[24] = {0.01234122979999,0.02853138862893,0.0442774388174,0.059298584915437,0.0733464814111,0.08619016153195,0.0976186521041,0.107444270116,0.11550566805373,0.1216704729278,0.12583745634683,0.1279381953468,0.1279381953468,0.1258374563468,0.1216704729278,0.1155056680537,0.107444270116,0.09761865210411,0.08619016153195,0.07334648141108,0.05929858491544,0.04427743881742,0.02853138862893,0.01234122979999};
  const double a = (mass_parent + mass_1) * (mass_parent + mass_1), b = (mass_parent - mass_1) * (mass_parent - mass_1);
  const double c = (mass_2 + mass_3) * (mass_2 + mass_3), d = (mass_2 - mass_3) * (mass_2 - mass_3);
  double Q = 0.0;
  for (int i = 0; i < 24; i++) {
    const double s = c + (b - c) * (1.0 + x_root[i]) / 2.0;
    Q += x_weight[i] * (b - c) * sqrt(fabs((a - s) * (b - s) * (s - c) * (s - d))) / (2.0 * s);
  }
  return Q;
}

/* LUP_decomposition + LUP_solve for n = 2, arsenal.cpp:1072-1207 */
static void lup2_solve(double A[2][2], double b[2])
{
  const int n = 2;
  int pvector[2] = {0, 1}, imax = 0;
  double implicit_scale[2], big, sum, temp;
  for (int i = 0; i < n; i++) {
    big = 0.0;
    for (int j = 0; j < n; j++) { temp = fabs(A[i][j]); if (temp > big) big = temp; }
    if (big == 0.0) break;
    implicit_scale[i] = 1.0 / big;
  }
  for (int j = 0; j < n; j++) {
    for (int i = 0; i < j; i++) { sum = A[i][j]; for (int k = 0; k < i; k++) sum -= A[i][k] * A[k][j]; A[i][j] = sum; }
    big = 0.0;
    for (int i = j; i < n; i++) {
      sum = A[i][j];
      for (int k = 0; k < j; k++) sum -= A[i][k] * A[k][j];
      A[i][j] = sum;
      temp = implicit_scale[i] * fabs(sum);
      if (temp >= big) { big = temp; imax = i; }
    }
    if (j != imax) {
      for (int k = 0; k < n; k++) { temp = A[imax][k]; A[imax][k] = A[j][k]; A[j][k] = temp; }
      implicit_scale[imax] = implicit_scale[j];
    }
    pvector[j] = imax;
    if (A[j][j] == 0.0) A[j][j] = 1.0e-16;
    if (j != n - 1) { temp = 1.0 / A[j][j]; for (int i = j + 1; i < n; i++) A[i][j] *= temp; }
  }
  for (int i = 0; i < n; i++) {
    const int ip = pvector[i];
    sum = b[ip]; b[ip] = b[i];
    for (int j = 0; j < i; j++) sum -= A[i][j] * b[j];
    b[i] = sum;
  }
  for (int i = n - 1; i >= 0; i--) {
    sum = b[i];
    for (int j = i + 1; j < n; j++) sum -= A[i][j] * b[j];
    b[i] = sum / A[i][i];
  }
}

/* EmissionFunctionArray::estimate_MT_function_of_dNdypTdpTdphi, :2032-2158: least-squares line through (mT, log dN) of the parent
 * for mT > sqrt(2.73) M, up to the first non-finite log */
static mt_fit estimate_mt_fit(ctx_t *c, int iy, int iphip, double mass_parent)
{
  mt_fit out = {0.0, 0.0};
  double mT_points[512], logdN_points[512];
  int n = 0;
  for (int ipT = 0; ipT < c->n_pT; ipT++) {
    const long iS = ipT + (long)c->n_pT * (iphip + (long)c->n_phi * iy);
    const double logdN = c->logdN[iS];
    if (isfinite(logdN)) {
      const double pT = c->pT[ipT];
      const double mT = sqrt(mass_parent * mass_parent + pT * pT);
      if (mT > sqrt(2.73) * mass_parent) { mT_points[n] = mT; logdN_points[n] = logdN; n++; }
    } else break;
  }
  if (n < 2) { c->err = -14; return out; }
  /* f = A^T y, M = A^T A with A = [1 mT], sums in point order */
  double f[2] = {0.0, 0.0}, M[2][2];
  for (int i = 0; i < 2; i++) { double sum = 0.0; for (int k = 0; k < n; k++) sum += ((i == 0 ? 1.0 : mT_points[k]) * logdN_points[k]); f[i] = sum; }
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++) {
      double sum = 0.0;
      for (int k = 0; k < n; k++) sum += ((i == 0 ? 1.0 : mT_points[k]) * (j == 0 ? 1.0 : mT_points[k]));
      M[i][j] = sum;
    }
  lup2_solve(M, f);
  out.constant = f[0]; out.slope = f[1];
  return out;
}

/* left / right interpolation points in the azimuthal table, the block repeated at :1439-1489, 1586-1636, 1706-1757, 1882-1932 */
static void phi_points(const ctx_t *c, double *Phip, int *iL, int *iR, double *PL, double *PR)
{
  const double Phip_min = c->phi[0], Phip_max = c->phi[c->n_phi - 1];
  if (*Phip >= Phip_min && *Phip <= Phip_max) {
    int r = 1;
    while (*Phip > c->phi[r]) r++;
    *iR = r; *iL = r - 1; *PL = c->phi[*iL]; *PR = c->phi[*iR];
  } else {
    *iL = c->n_phi - 1; *iR = 0;
    *PL = c->phi[*iL] - 2.0 * M_PI; *PR = c->phi[*iR];
    *Phip -= floor(*Phip / M_PI) * (2.0 * M_PI);
  }
}

/* dN_dYMTdMTdPhi_boost_invariant, :1413-1676 */
static double parent_2d(const ctx_t *c, double MT, double Phip1, double Phip2, double MTmax)
{
  double logdN1 = 0.0, logdN2 = 0.0;
  int i1L, i1R, i2L, i2R; double P1L, P1R, P2L, P2R;
  const int npT = c->n_pT;
  if (MT <= MTmax) {
    phi_points(c, &Phip1, &i1L, &i1R, &P1L, &P1R);
    phi_points(c, &Phip2, &i2L, &i2R, &P2L, &P2R);
    int iMTR = 1;
    while (MT > c->MTValues[iMTR]) iMTR++;
    const int iMTL = iMTR - 1;
    const double MTL = c->MTValues[iMTL], MTR = c->MTValues[iMTR];
    const double dPhip1 = P1R - P1L, dPhip2 = P2R - P2L, dMT = MTR - MTL;
    const double a_LL = c->logdN[iMTL + npT * i1L], a_RL = c->logdN[iMTL + npT * i1R], a_LR = c->logdN[iMTR + npT * i1L], a_RR = c->logdN[iMTR + npT * i1R];
    const double b_LL = c->logdN[iMTL + npT * i2L], b_RL = c->logdN[iMTL + npT * i2R], b_LR = c->logdN[iMTR + npT * i2L], b_RR = c->logdN[iMTR + npT * i2R];
    logdN1 = ((a_LL * (P1R - Phip1) + a_RL * (Phip1 - P1L)) * (MTR - MT) + (a_LR * (P1R - Phip1) + a_RR * (Phip1 - P1L)) * (MT - MTL)) / (dPhip1 * dMT);
    logdN2 = ((b_LL * (P2R - Phip2) + b_RL * (Phip2 - P2L)) * (MTR - MT) + (b_LR * (P2R - Phip2) + b_RR * (Phip2 - P2L)) * (MT - MTL)) / (dPhip2 * dMT);
  } else {
    phi_points(c, &Phip1, &i1L, &i1R, &P1L, &P1R);
    phi_points(c, &Phip2, &i2L, &i2R, &P2L, &P2R);
    const double dPhip1 = P1R - P1L, dPhip2 = P2R - P2L;
    const mt_fit f1L = c->fit[i1L], f1R = c->fit[i1R], f2L = c->fit[i2L], f2R = c->fit[i2R];
    const double a_L = f1L.constant + f1L.slope * MT, a_R = f1R.constant + f1R.slope * MT;
    const double b_L = f2L.constant + f2L.slope * MT, b_R = f2R.constant + f2R.slope * MT;
    logdN1 = (a_L * (P1R - Phip1) + a_R * (Phip1 - P1L)) / dPhip1;
    logdN2 = (b_L * (P2R - Phip2) + b_R * (Phip2 - P2L)) / dPhip2;
  }
  return (exp(logdN1) + exp(logdN2));
}

/* dN_dYMTdMTdPhi_non_boost_invariant, :1680-2028 */
static double parent_3d(const ctx_t *c, int iYL, int iYR, double YL, double YR, double MT, double Phip1, double Phip2, double Y, double MTmax)
{
  double logdN1 = 0.0, logdN2 = 0.0;
  int i1L, i1R, i2L, i2R; double P1L, P1R, P2L, P2R;
  const long npT = c->n_pT, nphi = c->n_phi;
  phi_points(c, &Phip1, &i1L, &i1R, &P1L, &P1R);
  phi_points(c, &Phip2, &i2L, &i2R, &P2L, &P2R);
  const double dY = YR - YL, dPhip1 = P1R - P1L, dPhip2 = P2R - P2L;
  if (MT <= MTmax) {
    int iMTR = 1;
    while (MT > c->MTValues[iMTR]) iMTR++;
    const int iMTL = iMTR - 1;
    const double MTL = c->MTValues[iMTL], MTR = c->MTValues[iMTR], dMT = MTR - MTL;
#define LG(im, ip, iy) c->logdN[(im) + npT * ((ip) + nphi * (iy))]
    const double a_LLL = LG(iMTL, i1L, iYL), a_RLL = LG(iMTL, i1L, iYR), a_LRL = LG(iMTL, i1R, iYL), a_RRL = LG(iMTL, i1R, iYR);
    const double a_LLR = LG(iMTR, i1L, iYL), a_RLR = LG(iMTR, i1L, iYR), a_LRR = LG(iMTR, i1R, iYL), a_RRR = LG(iMTR, i1R, iYR);
    const double b_LLL = LG(iMTL, i2L, iYL), b_RLL = LG(iMTL, i2L, iYR), b_LRL = LG(iMTL, i2R, iYL), b_RRL = LG(iMTL, i2R, iYR);
    const double b_LLR = LG(iMTR, i2L, iYL), b_RLR = LG(iMTR, i2L, iYR), b_LRR = LG(iMTR, i2R, iYL), b_RRR = LG(iMTR, i2R, iYR);
#undef LG
    logdN1 = (MTR - MT) * ((a_LLL * (YR - Y) + a_RLL * (Y - YL)) * (P1R - Phip1) + (a_LRL * (YR - Y) + a_RRL * (Y - YL)) * (Phip1 - P1L))
           + (MT - MTL) * ((a_LLR * (YR - Y) + a_RLR * (Y - YL)) * (P1R - Phip1) + (a_LRR * (YR - Y) + a_RRR * (Y - YL)) * (Phip1 - P1L));
    logdN1 /= (dY * dPhip1 * dMT);
    logdN2 = (MTR - MT) * ((b_LLL * (YR - Y) + b_RLL * (Y - YL)) * (P2R - Phip2) + (b_LRL * (YR - Y) + b_RRL * (Y - YL)) * (Phip2 - P2L))
           + (MT - MTL) * ((b_LLR * (YR - Y) + b_RLR * (Y - YL)) * (P2R - Phip2) + (b_LRR * (YR - Y) + b_RRR * (Y - YL)) * (Phip2 - P2L));
    logdN2 /= (dY * dPhip2 * dMT);
  } else {
    const mt_fit f1LL = c->fit[iYL * nphi + i1L], f1RL = c->fit[iYR * nphi + i1L], f1LR = c->fit[iYL * nphi + i1R], f1RR = c->fit[iYR * nphi + i1R];
    const mt_fit f2LL = c->fit[iYL * nphi + i2L], f2RL = c->fit[iYR * nphi + i2L], f2LR = c->fit[iYL * nphi + i2R], f2RR = c->fit[iYR * nphi + i2R];
    const double a_LL = f1LL.constant + f1LL.slope * MT, a_LR = f1LR.constant + f1LR.slope * MT, a_RL = f1RL.constant + f1RL.slope * MT, a_RR = f1RR.constant + f1RR.slope * MT;
    const double b_LL = f2LL.constant + f2LL.slope * MT, b_LR = f2LR.constant + f2LR.slope * MT, b_RL = f2RL.constant + f2RL.slope * MT, b_RR = f2RR.constant + f2RR.slope * MT;
    logdN1 = (a_LL * (YR - Y) + a_RL * (Y - YL)) * (P1R - Phip1) + (a_LR * (YR - Y) + a_RR * (Y - YL)) * (Phip1 - P1L);
    logdN1 /= (dY * dPhip1);
    logdN2 = (b_LL * (YR - Y) + b_RL * (Y - YL)) * (P2R - Phip2) + (b_LR * (YR - Y) + b_RR * (Y - YL)) * (Phip2 - P2L);
    logdN2 /= (dY * dPhip2);
  }
  return (exp(logdN1) + exp(logdN2));
}

/* the (zeta) integral at fixed (v, s): shared tail of the four integration loops (:588-628, 753-784, 1183-1213, 1352-1380) */
static double zeta_integral_at(const ctx_t *c, double MTbar, double DeltaMT, double mT_coshvDeltaY_over_pT, double Estar_M_over_pT,
                               double parent_mass2, double phip, double MTmax, int dim, int cutoff_Y, int iYL, int iYR, double YL, double YR, double Y)
{
  const double two_Pi = 2.0 * M_PI;
  double zeta_integral = 0.0;
  for (int izeta = 0; izeta < GP; izeta++) {
    if (cutoff_Y) break;
    const double coszeta = cos((M_PI / 2.0) * (1.0 + GL_ROOT[izeta]));
    const double MT = MTbar + (DeltaMT * coszeta);
    const double PT = sqrt(MT * MT - parent_mass2);
    const double cosPhip_tilde = (MT * mT_coshvDeltaY_over_pT - Estar_M_over_pT) / PT;
    const double Phip_tilde = acos(cosPhip_tilde);
    double Phip_1 = fmod(Phip_tilde + phip, two_Pi), Phip_2 = fmod(-Phip_tilde + phip, two_Pi);
    if (Phip_1 < 0.0) Phip_1 += two_Pi;
    if (Phip_2 < 0.0) Phip_2 += two_Pi;
    const double integrand = MT * (dim == 2 ? parent_2d(c, MT, Phip_1, Phip_2, MTmax) : parent_3d(c, iYL, iYR, YL, YR, MT, Phip_1, Phip_2, Y, MTmax));
    zeta_integral += (GL_WEIGHT[izeta] * integrand);
  }
  return zeta_integral;
}

/* parent rapidity Y -> interpolation points (:719-738, 1308-1327) */
static int y_points(const ctx_t *c, double Y, double Ymax, int *iYL, int *iYR, double *YL, double *YR)
{
  if (fabs(Y) <= Ymax) {
    int r = 1;
    while (Y > c->y[r]) r++;
    *iYR = r; *iYL = r - 1; *YL = c->y[*iYL]; *YR = c->y[*iYR];
    return 0;
  }
  return 1;
}

/* group the daughters that are chosen species by type (:307-371, 828-899) */
static int group_daughters(const ctx_t *c, const int *prod, int nprod, int *groups, int *members)
{
  int found[3] = {0, 0, 0}, n_sel = 0, sel[3], ng = 0;
  for (int ic = 0; ic < c->n_chosen; ic++) {
    int all = 1;
    for (int k = 0; k < nprod; k++) { if (prod[k] == c->chosen[ic] && !found[k]) found[k] = 1; all = all && found[k]; }
    if (all) break;
  }
  for (int k = 0; k < nprod; k++) if (found[k]) sel[n_sel++] = prod[k];
  for (int s = 0; s < n_sel; s++) {
    int put = 0;
    for (int gi = 0; gi < ng; gi++) if (sel[s] == groups[gi]) { members[gi] += 1; put = 1; break; }
    if (!put) { groups[ng] = sel[s]; members[ng] = 1; ng++; }
  }
  return ng;
}

static void setup_parent_tables(ctx_t *c, double mass_parent)
{
  for (int ipT = 0; ipT < c->n_pT; ipT++) c->MTValues[ipT] = sqrt(fabs(c->pT[ipT] * c->pT[ipT] + mass_parent * mass_parent));
  for (int iphip = 0; iphip < c->n_phi; iphip++)
    for (int iy = 0; iy < c->y_pts; iy++) c->fit[iy * c->n_phi + iphip] = estimate_mt_fit(c, iy, iphip, mass_parent);
}

/* EmissionFunctionArray::two_body_decay, :296-812 */
static void two_body_decay(ctx_t *c, double branch_ratio, int parent_chosen_index, int particle_1, int particle_2, double mass_parent)
{
  const cfo_particles *pd = c->pdg;
  const int prod[2] = {particle_1, particle_2};
  int groups[2], members[2];
  const int ng = group_daughters(c, prod, 2, groups, members);
  (void)parent_chosen_index;
  if (ng == 0) return;
  int chosen_index[2]; double mass_squared[2], momentum_star[2], energy_star[2];
  for (int gi = 0; gi < ng; gi++) {
    chosen_index[gi] = particle_chosen_index(c, groups[gi], &c->err);
    const double mass = pd->mass[groups[gi]];
    mass_squared[gi] = mass * mass;
    /* the reference takes the mass of particle_2 as the recoil mass whichever daughter is looked at (:411-413) */
    const double mass_secondary = pd->mass[particle_2];
    const double W2 = mass_secondary * mass_secondary;
    const double Estar = (mass_parent * mass_parent + mass * mass - W2) / (2.0 * mass_parent);
    energy_star[gi] = Estar;
    momentum_star[gi] = sqrt(Estar * Estar - mass * mass);
  }
  setup_parent_tables(c, mass_parent);
  if (c->err) return;
  const double Ymax = fabs(c->dim == 2 ? 0.0 : c->y[c->y_pts - 1]);
  const double MTmax = c->MTValues[c->n_pT - 1];
  for (int gi = 0; gi < ng; gi++) {
    const double multiplicity = (double)members[gi];
    const double parent_mass2 = mass_parent * mass_parent;
    const double mass2 = mass_squared[gi], Estar = energy_star[gi], Estar2 = Estar * Estar, pstar = momentum_star[gi];
    const double Estar_M = Estar * mass_parent;
    const double prefactor = multiplicity * mass_parent * branch_ratio / (8.0 * pstar);
    for (int ipT = 0; ipT < c->n_pT; ipT++) {
      const double pT = c->pT[ipT], pT2 = pT * pT, mT2 = pT2 + mass2, mT = sqrt(mT2);
      const double M_pT = mass_parent * pT, Estar_M_mT = Estar_M * mT, Estar_M_over_pT = Estar_M / pT, Estar2_plus_pT2 = Estar2 + pT2;
      const double DeltaY = log((pstar + sqrt(Estar2_plus_pT2)) / mT);
      double MTbar_table[GP], DeltaMT_table[GP], mTc_table[GP], vw_table[GP];
      for (int k = 0; k < GP; k++) {
        const double coshvDeltaY = cosh(GL_ROOT[k] * DeltaY);
        const double mT2_coshvDeltaY2 = mT2 * coshvDeltaY * coshvDeltaY;
        const double den = mT2_coshvDeltaY2 - pT2;
        MTbar_table[k] = Estar_M_mT * coshvDeltaY / den;
        DeltaMT_table[k] = M_pT * sqrt(fabs(Estar2_plus_pT2 - mT2_coshvDeltaY2)) / den;
        mTc_table[k] = mT * coshvDeltaY / pT;
        vw_table[k] = DeltaY * GL_WEIGHT[k] / sqrt(fabs(den));
      }
      for (int iphip = 0; iphip < c->n_phi; iphip++) {
        const double phip = c->phi[iphip];
        for (int iy = 0; iy < c->y_pts; iy++) {
          const double y = (c->dim == 2) ? 0.0 : c->y[iy];
          double decay2D_integral = 0.0;
          for (int iv = 0; iv < GP; iv++) {
            int iYL = 0, iYR = 0, cutoff_Y = 0; double YL = 0.0, YR = 0.0, Y = 0.0;
            if (c->dim == 3) { Y = y + GL_ROOT[iv] * DeltaY; cutoff_Y = y_points(c, Y, Ymax, &iYL, &iYR, &YL, &YR); }
            const double zi = zeta_integral_at(c, MTbar_table[iv], DeltaMT_table[iv], mTc_table[iv], Estar_M_over_pT, parent_mass2, phip, MTmax,
                                               c->dim, cutoff_Y, iYL, iYR, YL, YR, Y);
            decay2D_integral += (vw_table[iv] * zi);
          }
          const long iS3D = chosen_index[gi] + (long)c->n_chosen * (ipT + (long)c->n_pT * (iphip + (long)c->n_phi * iy));
          c->dN[iS3D] += prefactor * decay2D_integral;
        }
      }
    }
  }
}

/* EmissionFunctionArray::three_body_decay, :816-1409 */
static void three_body_decay(ctx_t *c, double branch_ratio, int particle_1, int particle_2, int particle_3, double mass_parent)
{
  const cfo_particles *pd = c->pdg;
  const int prod[3] = {particle_1, particle_2, particle_3};
  int groups[3], members[3];
  const int ng = group_daughters(c, prod, 3, groups, members);
  if (ng == 0) return;
  int chosen_index[3]; double m1g[3], m2g[3], m3g[3], Qg[3];
  for (int gi = 0; gi < ng; gi++) {
    chosen_index[gi] = particle_chosen_index(c, groups[gi], &c->err);
    m1g[gi] = pd->mass[groups[gi]];
    int rest[2], nr = 0, removed = 0;
    for (int k = 0; k < 3; k++) { if (!removed && prod[k] == groups[gi]) { removed = 1; continue; } rest[nr++] = prod[k]; }
    m2g[gi] = pd->mass[rest[0]]; m3g[gi] = pd->mass[rest[1]];
    Qg[gi] = q_factor(mass_parent, m1g[gi], m2g[gi], m3g[gi]);
  }
  setup_parent_tables(c, mass_parent);
  if (c->err) return;
  const double Ymax = fabs(c->dim == 2 ? 0.0 : c->y[c->y_pts - 1]);
  const double MTmax = c->MTValues[c->n_pT - 1];
  for (int gi = 0; gi < ng; gi++) {
    const double multiplicity = (double)members[gi];
    const double parent_mass2 = mass_parent * mass_parent;
    const double mass_1 = m1g[gi], mass_1_squared = mass_1 * mass_1, mass_2 = m2g[gi], mass_3 = m3g[gi];
    const double s_plus = (mass_parent - mass_1) * (mass_parent - mass_1), s_minus = (mass_2 + mass_3) * (mass_2 + mass_3);
    const double d = (mass_2 - mass_3) * (mass_2 - mass_3), Q_norm = Qg[gi];
    double s_root[GP], Estar_table[GP], pstar_table[GP], sw_table[GP];
    for (int k = 0; k < GP; k++) {
      const double s = s_minus + (s_plus - s_minus) * (1.0 + GL_ROOT[k]) / 2.0;
      const double Estar = (parent_mass2 + mass_1_squared - s) / (2.0 * mass_parent);
      s_root[k] = s; Estar_table[k] = Estar;
      sw_table[k] = GL_WEIGHT[k] * sqrt(fabs((s - s_minus) * (s - d))) / s;
      pstar_table[k] = sqrt(Estar * Estar - mass_1_squared);
    }
    const double prefactor = multiplicity * parent_mass2 * (s_plus - s_minus) * branch_ratio / (8.0 * Q_norm);
    for (int ipT = 0; ipT < c->n_pT; ipT++) {
      const double pT = c->pT[ipT], pT2 = pT * pT, mT2 = pT2 + mass_1_squared, mT = sqrt(mT2);
      const double M_pT = mass_parent * pT, M_mT = mass_parent * mT, M_over_pT = mass_parent / pT, mT_over_pT = mT / pT;
      for (int iphip = 0; iphip < c->n_phi; iphip++) {
        const double phip = c->phi[iphip];
        for (int iy = 0; iy < c->y_pts; iy++) {
          const double y = (c->dim == 2) ? 0.0 : c->y[iy];
          double decay3D_integral = 0.0;
          for (int is = 0; is < GP; is++) {
            const double Estar = Estar_table[is], Estar2 = Estar * Estar, pstar = pstar_table[is];
            const double Estar_M_mT = Estar * M_mT, Estar2_plus_pT2 = Estar2 + pT2, Estar_M_over_pT = Estar * M_over_pT;
            const double DeltaY = log((pstar + sqrt(Estar2_plus_pT2)) / mT);
            double v_integral = 0.0;
            for (int iv = 0; iv < GP; iv++) {
              const double v = GL_ROOT[iv];
              int iYL = 0, iYR = 0, cutoff_Y = 0; double YL = 0.0, YR = 0.0, Y = 0.0;
              if (c->dim == 3) { Y = y + v * DeltaY; cutoff_Y = y_points(c, Y, Ymax, &iYL, &iYR, &YL, &YR); }
              const double coshvDeltaY = cosh(v * DeltaY);
              /* 2+1D groups mT^2 cosh cosh left to right (:1167), 3+1D squares cosh first (:1329-1330) */
              const double mT2_coshvDeltaY2 = (c->dim == 2) ? mT2 * coshvDeltaY * coshvDeltaY : mT2 * (coshvDeltaY * coshvDeltaY);
              const double den = mT2_coshvDeltaY2 - pT2;
              const double mTc = mT_over_pT * coshvDeltaY;
              const double MTbar = Estar_M_mT * coshvDeltaY / den;
              const double DeltaMT = M_pT * sqrt(fabs(Estar2_plus_pT2 - mT2_coshvDeltaY2)) / den;
              const double vw = DeltaY * GL_WEIGHT[iv] / sqrt(fabs(den));
              const double zi = zeta_integral_at(c, MTbar, DeltaMT, mTc, Estar_M_over_pT, parent_mass2, phip, MTmax, c->dim, cutoff_Y, iYL, iYR, YL, YR, Y);
              v_integral += vw * zi;
            }
            /* 2+1D uses the tabulated s weight, 3+1D recomputes it with the same expression */
            (void)s_root;
            decay3D_integral += sw_table[is] * v_integral;
          }
          const long iS3D = chosen_index[gi] + (long)c->n_chosen * (ipT + (long)c->n_pT * (iphip + (long)c->n_phi * iy));
          c->dN[iS3D] += prefactor * decay3D_integral;
        }
      }
    }
  }
}

/* EmissionFunctionArray::do_resonance_decays + resonance_decay_channel, :124-292.  dN is amended in place. */
int cfo_resonance_decays(const cfo_particles *pdg, int32_t n_chosen, const int32_t *chosen_pdg_index, const cfo_grid *g, int32_t dimension,
                         double *dN)
{
  if (n_chosen - 1 <= 0) return -10;
  if (g->n_pT > 512) return -15;
  ctx_t c; memset(&c, 0, sizeof(c));
  c.pdg = pdg; c.n_chosen = n_chosen; c.chosen = chosen_pdg_index;
  c.n_pT = g->n_pT; c.n_phi = g->n_phi; c.n_y_tab = g->n_y; c.dim = dimension; c.y_pts = (dimension == 2) ? 1 : g->n_y;
  c.pT = g->pT; c.phi = g->phi; c.y = g->y; c.dN = dN;
  c.logdN = (double *)calloc((size_t)c.n_pT * c.n_phi * c.n_y_tab, sizeof(double));
  c.fit = (mt_fit *)calloc((size_t)c.y_pts * c.n_phi, sizeof(mt_fit));
  for (int ichosen = n_chosen - 1; ichosen > 0 && !c.err; ichosen--) {
    const int ipart = chosen_pdg_index[ichosen];
    if (pdg->stable[ipart]) continue;
    for (int ipT = 0; ipT < c.n_pT; ipT++)
      for (int iphip = 0; iphip < c.n_phi; iphip++)
        for (int iy = 0; iy < c.y_pts; iy++) {
          const long iS3D = ichosen + (long)n_chosen * (ipT + (long)c.n_pT * (iphip + (long)c.n_phi * iy));
          c.logdN[ipT + (long)c.n_pT * (iphip + (long)c.n_phi * iy)] = log(dN[iS3D]);
        }
    for (int ich = 0; ich < pdg->decays[ipart] && !c.err; ich++) {
      const int row = pdg->dec_first[ipart] + ich;
      const int decay_products = abs(pdg->dec_npart[row]);
      int idx[5];
      for (int k = 0; k < decay_products && k < 5; k++) idx[k] = particle_index(pdg, pdg->dec_part[row * 5 + k], &c.err);
      if (c.err) break;
      if (decay_products == 1) continue;
      const double branch_ratio = pdg->dec_br[row];
      if (decay_products == 2) {
        double mass_parent = pdg->mass[ipart], mass_1 = pdg->mass[idx[0]], mass_2 = pdg->mass[idx[1]];
        while ((mass_1 + mass_2) > mass_parent) {                           /* :241-256 */
          mass_parent += 0.25 * pdg->width[ipart];
          mass_1 -= 0.5 * pdg->width[idx[0]];
          mass_2 -= 0.5 * pdg->width[idx[1]];
          if (mass_1 < 0.0 || mass_2 < 0.0) { c.err = -16; break; }
        }
        if (!c.err) two_body_decay(&c, branch_ratio, ichosen, idx[0], idx[1], mass_parent);
      } else if (decay_products == 3) {
        three_body_decay(&c, branch_ratio, idx[0], idx[1], idx[2], pdg->mass[ipart]);
      } else if (decay_products == 4) {
        /* skipped by the reference (:279-282) */
      } else c.err = -17;
    }
  }
  free(c.logdN); free(c.fit);
  return c.err;
}
