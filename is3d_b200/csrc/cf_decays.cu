// cf_decays.cu -- resonance-decay feed-down of the smooth spectra on the GPU (sm_100a).
//
// Replaces EmissionFunctionArray::do_resonance_decays and the routines below it (reference
// src/cpp/emissionfunction_resonance_decays.cpp:124-2158): parents are taken from the last chosen species down to the second; for
// every unstable parent the logarithm of its (already amended) spectrum is tabulated, and each 2- / 3-body channel adds, for every
// daughter that is a chosen species, the phase-space integral of the parent spectrum -- 12-point Gauss-Legendre in (v, zeta)
// [and s], bi- / tri-linear interpolation of log dN in (M_T, Phi[, Y]), exponential extrapolation in M_T beyond the table -- to
// the daughter's bins.  NOTE: in the reference snapshot the routine is disabled by an exit(-1) at entry (":126-129, I need to change
// the linear interpolation's MTmax ..."); what is implemented here is the body behind it, checked against that body (oracle/).
//
// B200 mapping: the spectra array stays in HBM (39 MB at 305 species); the work per parent is (channels x daughter groups) "terms" x
// momentum bins, one thread per (term, bin) for 2-body terms (144 integrand points each) and 12 threads -- one per s node -- for
// 3-body terms (1728 points), whose partial sums are folded in node order.  Terms write prefactor x integral to a scratch array
// and a second kernel adds them to the daughters in (channel, group) order, so the result does not depend on the launch geometry
// and follows the reference's accumulation order.  The parent's log table (129 KB) is read through L1/L2.  Parents are inherently
// sequential (a parent's spectrum must have received its own feed-down first).
// Compiled with -fmad=false: products and sums round one by one like the reference's scalar C++.
#include "cf_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace is3d {

namespace {

constexpr int GP = 12;                  // Gauss-Legendre points of every decay integral (:493-498, 1001-1006)
__constant__ double kGLRoot[GP] = {-0.98156063424672, -0.90411725637048, -0.76990267419431, -0.58731795428662, -0.3678314989982, -0.12523340851147,
   0.12523340851147, 0.36783149899818, 0.58731795428662, 0.76990267419431, 0.90411725637048, 0.98156063424672};
__constant__ double kGLWeight[GP] = {0.04717533638651, 0.1069393259953, 0.16007832854335, 0.20316742672307, 0.23349253653836, 0.2491470458134,
   0.2491470458134, 0.23349253653836, 0.20316742672307, 0.1600783285433, 0.10693932599532, 0.04717533638651};

struct DecayTerm {
  int body;                 // 2 or 3
  int daughter;             // chosen index of the daughter species the term adds to
  double mass_parent;       // (for 2-body channels possibly shifted to satisfy energy conservation, :241-256)
  double prefactor;
  double mass2, Estar, pstar;                  // 2-body: daughter mass^2, energy and momentum in the parent rest frame
  double m1sq, s_plus, s_minus, d;             // 3-body: daughter mass^2 and the invariant-mass range of the other pair
};

struct DecayGrid {
  int n_species, n_pT, n_phi, n_y_tab, y_pts, dim;
  const double *pT, *phi, *y;
};

struct Fit { double constant, slope; };

// ---- per-term tables: parent transverse masses and the exponential M_T fits
__device__ void lup2_solve(double A[2][2], double b[2])                        // arsenal.cpp:1072-1207 for n = 2
{
  const int n = 2;
  int pvector[2] = {0, 1}, imax = 0;
  double implicit_scale[2] = {0.0, 0.0}, big, sum, temp;
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

__global__ void log_parent_kernel(DecayGrid G, const double *__restrict__ dN, int parent, double *__restrict__ logdN)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = G.n_pT * G.n_phi * G.y_pts;
  if (i >= n) return;
  logdN[i] = log(dN[(int64_t)parent + (int64_t)G.n_species * i]);               // :163-171; i = ipT + n_pT (iphi + n_phi iy)
}

// one block per term: MT[ipT] and, per (iy, iphi), the least-squares line through (mT, log dN) for mT > sqrt(2.73) M (:2032-2158)
__global__ void decay_setup_kernel(DecayGrid G, const DecayTerm *__restrict__ terms, const double *__restrict__ logdN,
                                   double *__restrict__ MT_all, Fit *__restrict__ fit_all, int *__restrict__ error)
{
  const DecayTerm T = terms[blockIdx.x];
  double *MT = MT_all + (size_t)blockIdx.x * G.n_pT;
  Fit *fit = fit_all + (size_t)blockIdx.x * G.y_pts * G.n_phi;
  const double M = T.mass_parent;
  for (int ipT = threadIdx.x; ipT < G.n_pT; ipT += blockDim.x) MT[ipT] = sqrt(fabs(G.pT[ipT] * G.pT[ipT] + M * M));
  for (int w = threadIdx.x; w < G.y_pts * G.n_phi; w += blockDim.x) {
    const int iy = w / G.n_phi, iphip = w - iy * G.n_phi;
    double f[2] = {0.0, 0.0}, A[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    int n = 0;
    // the reference collects the points first and then forms A^T y and A^T A with sums in point order; accumulating them here in
    // the same order gives the same sums
    for (int ipT = 0; ipT < G.n_pT; ipT++) {
      const double l = logdN[ipT + G.n_pT * (iphip + G.n_phi * iy)];
      if (!isfinite(l)) break;
      const double pT = G.pT[ipT];
      const double mT = sqrt(M * M + pT * pT);
      if (mT > sqrt(2.73) * M) {
        f[0] += (1.0 * l); f[1] += (mT * l);
        A[0][0] += (1.0 * 1.0); A[0][1] += (1.0 * mT); A[1][0] += (mT * 1.0); A[1][1] += (mT * mT);
        n++;
      }
    }
    if (n < 2) { atomicExch(error, 1); fit[w].constant = 0.0; fit[w].slope = 0.0; continue; }    // the reference exits here (:2078-2082)
    lup2_solve(A, f);
    fit[w].constant = f[0]; fit[w].slope = f[1];
  }
}

// ---- interpolation of the parent's log spectrum
struct ParentView {
  const double *logdN, *MT; const Fit *fit; const double *phi, *y;
  int n_pT, n_phi; double MTmax;
};

// left / right points in the azimuthal table; angles outside [phi_0, phi_last] wrap around 2 pi (:1439-1489 and its three copies)
__device__ __forceinline__ void phi_points(const ParentView &P, double &Phip, int &iL, int &iR, double &PL, double &PR)
{
  if (Phip >= P.phi[0] && Phip <= P.phi[P.n_phi - 1]) {
    int r = 1;
    while (Phip > P.phi[r]) r++;
    iR = r; iL = r - 1; PL = P.phi[iL]; PR = P.phi[iR];
  } else {
    iL = P.n_phi - 1; iR = 0;
    PL = P.phi[iL] - 2.0 * M_PI; PR = P.phi[iR];
    Phip -= floor(Phip / M_PI) * (2.0 * M_PI);
  }
}

__device__ double parent_2d(const ParentView &P, double MT, double Phip1, double Phip2)                     // :1413-1676
{
  double logdN1, logdN2;
  int i1L, i1R, i2L, i2R; double P1L, P1R, P2L, P2R;
  phi_points(P, Phip1, i1L, i1R, P1L, P1R);
  phi_points(P, Phip2, i2L, i2R, P2L, P2R);
  const double dPhip1 = P1R - P1L, dPhip2 = P2R - P2L;
  if (MT <= P.MTmax) {
    int iMTR = 1;
    while (MT > P.MT[iMTR]) iMTR++;
    const int iMTL = iMTR - 1, npT = P.n_pT;
    const double MTL = P.MT[iMTL], MTR = P.MT[iMTR], dMT = MTR - MTL;
    const double a_LL = P.logdN[iMTL + npT * i1L], a_RL = P.logdN[iMTL + npT * i1R], a_LR = P.logdN[iMTR + npT * i1L], a_RR = P.logdN[iMTR + npT * i1R];
    const double b_LL = P.logdN[iMTL + npT * i2L], b_RL = P.logdN[iMTL + npT * i2R], b_LR = P.logdN[iMTR + npT * i2L], b_RR = P.logdN[iMTR + npT * i2R];
    logdN1 = ((a_LL * (P1R - Phip1) + a_RL * (Phip1 - P1L)) * (MTR - MT) + (a_LR * (P1R - Phip1) + a_RR * (Phip1 - P1L)) * (MT - MTL)) / (dPhip1 * dMT);
    logdN2 = ((b_LL * (P2R - Phip2) + b_RL * (Phip2 - P2L)) * (MTR - MT) + (b_LR * (P2R - Phip2) + b_RR * (Phip2 - P2L)) * (MT - MTL)) / (dPhip2 * dMT);
  } else {
    const Fit f1L = P.fit[i1L], f1R = P.fit[i1R], f2L = P.fit[i2L], f2R = P.fit[i2R];
    const double a_L = f1L.constant + f1L.slope * MT, a_R = f1R.constant + f1R.slope * MT;
    const double b_L = f2L.constant + f2L.slope * MT, b_R = f2R.constant + f2R.slope * MT;
    logdN1 = (a_L * (P1R - Phip1) + a_R * (Phip1 - P1L)) / dPhip1;
    logdN2 = (b_L * (P2R - Phip2) + b_R * (Phip2 - P2L)) / dPhip2;
  }
  return (exp(logdN1) + exp(logdN2));
}

__device__ double parent_3d(const ParentView &P, int iYL, int iYR, double YL, double YR, double MT, double Phip1, double Phip2, double Y)   // :1680-2028
{
  double logdN1, logdN2;
  int i1L, i1R, i2L, i2R; double P1L, P1R, P2L, P2R;
  const int npT = P.n_pT, nphi = P.n_phi;
  phi_points(P, Phip1, i1L, i1R, P1L, P1R);
  phi_points(P, Phip2, i2L, i2R, P2L, P2R);
  const double dY = YR - YL, dPhip1 = P1R - P1L, dPhip2 = P2R - P2L;
  if (MT <= P.MTmax) {
    int iMTR = 1;
    while (MT > P.MT[iMTR]) iMTR++;
    const int iMTL = iMTR - 1;
    const double MTL = P.MT[iMTL], MTR = P.MT[iMTR], dMT = MTR - MTL;
#define LG(im, ip, iy) P.logdN[(im) + npT * ((ip) + nphi * (iy))]
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
    const Fit f1LL = P.fit[iYL * nphi + i1L], f1RL = P.fit[iYR * nphi + i1L], f1LR = P.fit[iYL * nphi + i1R], f1RR = P.fit[iYR * nphi + i1R];
    const Fit f2LL = P.fit[iYL * nphi + i2L], f2RL = P.fit[iYR * nphi + i2L], f2LR = P.fit[iYL * nphi + i2R], f2RR = P.fit[iYR * nphi + i2R];
    const double a_LL = f1LL.constant + f1LL.slope * MT, a_LR = f1LR.constant + f1LR.slope * MT, a_RL = f1RL.constant + f1RL.slope * MT, a_RR = f1RR.constant + f1RR.slope * MT;
    const double b_LL = f2LL.constant + f2LL.slope * MT, b_LR = f2LR.constant + f2LR.slope * MT, b_RL = f2RL.constant + f2RL.slope * MT, b_RR = f2RR.constant + f2RR.slope * MT;
    logdN1 = (a_LL * (YR - Y) + a_RL * (Y - YL)) * (P1R - Phip1) + (a_LR * (YR - Y) + a_RR * (Y - YL)) * (Phip1 - P1L);
    logdN1 /= (dY * dPhip1);
    logdN2 = (b_LL * (YR - Y) + b_RL * (Y - YL)) * (P2R - Phip2) + (b_LR * (YR - Y) + b_RR * (Y - YL)) * (Phip2 - P2L);
    logdN2 /= (dY * dPhip2);
  }
  return (exp(logdN1) + exp(logdN2));
}

// zeta integral at fixed (v [, s]): the innermost loop of the four integration routines (:588-628, 753-784, 1183-1213, 1352-1380)
__device__ double zeta_integral_at(const ParentView &P, double MTbar, double DeltaMT, double mTc_over_pT, double Estar_M_over_pT, double parent_mass2,
                                   double phip, int dim, bool cutoff_Y, int iYL, int iYR, double YL, double YR, double Y)
{
  const double two_Pi = 2.0 * M_PI;
  double zeta_integral = 0.0;
  if (cutoff_Y) return zeta_integral;
  for (int izeta = 0; izeta < GP; izeta++) {
    const double coszeta = cos((M_PI / 2.0) * (1.0 + kGLRoot[izeta]));
    const double MT = MTbar + (DeltaMT * coszeta);
    const double PT = sqrt(MT * MT - parent_mass2);
    const double cosPhip_tilde = (MT * mTc_over_pT - Estar_M_over_pT) / PT;
    const double Phip_tilde = acos(cosPhip_tilde);
    double Phip_1 = fmod(Phip_tilde + phip, two_Pi), Phip_2 = fmod(-Phip_tilde + phip, two_Pi);
    if (Phip_1 < 0.0) Phip_1 += two_Pi;
    if (Phip_2 < 0.0) Phip_2 += two_Pi;
    const double integrand = MT * (dim == 2 ? parent_2d(P, MT, Phip_1, Phip_2) : parent_3d(P, iYL, iYR, YL, YR, MT, Phip_1, Phip_2, Y));
    zeta_integral += (kGLWeight[izeta] * integrand);
  }
  return zeta_integral;
}

__device__ __forceinline__ bool y_points(const DecayGrid &G, double Y, double Ymax, int &iYL, int &iYR, double &YL, double &YR)
{
  if (fabs(Y) <= Ymax) {
    int r = 1;
    while (Y > G.y[r]) r++;
    iYR = r; iYL = r - 1; YL = G.y[iYL]; YR = G.y[iYR];
    return false;
  }
  return true;                                           // parent rapidity outside the table: cut off (:735-738)
}

__device__ __forceinline__ ParentView parent_view(const DecayGrid &G, const double *logdN, const double *MT_all, const Fit *fit_all, int term)
{
  ParentView P;
  P.logdN = logdN; P.MT = MT_all + (size_t)term * G.n_pT; P.fit = fit_all + (size_t)term * G.y_pts * G.n_phi;
  P.phi = G.phi; P.y = G.y; P.n_pT = G.n_pT; P.n_phi = G.n_phi; P.MTmax = P.MT[G.n_pT - 1];
  return P;
}

// EmissionFunctionArray::two_body_decay, integration part (:521-806): one thread per (term, bin)
__global__ void __launch_bounds__(128) decay2_kernel(DecayGrid G, const DecayTerm *__restrict__ terms, const int *__restrict__ term_ids, int n_terms2,
                                                    const double *__restrict__ logdN, const double *__restrict__ MT_all, const Fit *__restrict__ fit_all,
                                                    double *__restrict__ scratch)
{
  const int n_bins = G.n_pT * G.n_phi * G.y_pts;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)n_terms2 * n_bins) return;
  const int it = (int)(gid / n_bins), bin = (int)(gid - (int64_t)it * n_bins);
  const int term = term_ids[it];
  const DecayTerm T = terms[term];
  const int ipT = bin % G.n_pT, iphip = (bin / G.n_pT) % G.n_phi, iy = bin / (G.n_pT * G.n_phi);
  const ParentView P = parent_view(G, logdN, MT_all, fit_all, term);
  const double mass_parent = T.mass_parent, parent_mass2 = mass_parent * mass_parent;
  const double Estar = T.Estar, Estar2 = Estar * Estar, pstar = T.pstar, Estar_M = Estar * mass_parent;
  const double pT = G.pT[ipT], pT2 = pT * pT, mT2 = pT2 + T.mass2, mT = sqrt(mT2);
  const double M_pT = mass_parent * pT, Estar_M_mT = Estar_M * mT, Estar_M_over_pT = Estar_M / pT, Estar2_plus_pT2 = Estar2 + pT2;
  const double DeltaY = log((pstar + sqrt(Estar2_plus_pT2)) / mT);
  const double phip = G.phi[iphip], y = (G.dim == 2) ? 0.0 : G.y[iy];
  const double Ymax = (G.dim == 2) ? 0.0 : fabs(G.y[G.y_pts - 1]);
  double decay2D_integral = 0.0;
  for (int iv = 0; iv < GP; iv++) {
    const double v = kGLRoot[iv];
    int iYL = 0, iYR = 0; double YL = 0.0, YR = 0.0, Y = 0.0; bool cutoff_Y = false;
    if (G.dim == 3) { Y = y + v * DeltaY; cutoff_Y = y_points(G, Y, Ymax, iYL, iYR, YL, YR); }
    const double coshvDeltaY = cosh(v * DeltaY);
    const double mT2_coshvDeltaY2 = mT2 * coshvDeltaY * coshvDeltaY;
    const double den = mT2_coshvDeltaY2 - pT2;
    const double MTbar = Estar_M_mT * coshvDeltaY / den;
    const double DeltaMT = M_pT * sqrt(fabs(Estar2_plus_pT2 - mT2_coshvDeltaY2)) / den;
    const double mTc = mT * coshvDeltaY / pT;
    const double vw = DeltaY * kGLWeight[iv] / sqrt(fabs(den));
    const double zi = zeta_integral_at(P, MTbar, DeltaMT, mTc, Estar_M_over_pT, parent_mass2, phip, G.dim, cutoff_Y, iYL, iYR, YL, YR, Y);
    decay2D_integral += (vw * zi);
  }
  scratch[(size_t)term * n_bins + bin] = T.prefactor * decay2D_integral;
}

// EmissionFunctionArray::three_body_decay, integration part (:1048-1398): 12 threads per (term, bin), one per s node
constexpr int kBins3 = 8;               // bins per block of decay3_kernel
__global__ void __launch_bounds__(GP * kBins3) decay3_kernel(DecayGrid G, const DecayTerm *__restrict__ terms, const int *__restrict__ term_ids, int n_terms3,
                                                            const double *__restrict__ logdN, const double *__restrict__ MT_all, const Fit *__restrict__ fit_all,
                                                            double *__restrict__ scratch)
{
  __shared__ double part[kBins3][GP];
  const int n_bins = G.n_pT * G.n_phi * G.y_pts;
  const int is = threadIdx.x, lb = threadIdx.y;
  const int64_t gid = (int64_t)blockIdx.x * kBins3 + lb;
  const bool valid = gid < (int64_t)n_terms3 * n_bins;
  int term = 0, bin = 0;
  double contrib = 0.0, prefactor = 0.0;
  if (valid) {
    const int it = (int)(gid / n_bins);
    bin = (int)(gid - (int64_t)it * n_bins);
    term = term_ids[it];
    const DecayTerm T = terms[term];
    prefactor = T.prefactor;
    const int ipT = bin % G.n_pT, iphip = (bin / G.n_pT) % G.n_phi, iy = bin / (G.n_pT * G.n_phi);
    const ParentView P = parent_view(G, logdN, MT_all, fit_all, term);
    const double mass_parent = T.mass_parent, parent_mass2 = mass_parent * mass_parent, mass_1_squared = T.m1sq;
    const double s = T.s_minus + (T.s_plus - T.s_minus) * (1.0 + kGLRoot[is]) / 2.0;
    const double s_integrand_weight = kGLWeight[is] * sqrt(fabs((s - T.s_minus) * (s - T.d))) / s;
    const double Estar = (parent_mass2 + mass_1_squared - s) / (2.0 * mass_parent), Estar2 = Estar * Estar;
    const double pstar = sqrt(Estar * Estar - mass_1_squared);
    const double pT = G.pT[ipT], pT2 = pT * pT, mT2 = pT2 + mass_1_squared, mT = sqrt(mT2);
    const double M_pT = mass_parent * pT, M_mT = mass_parent * mT, M_over_pT = mass_parent / pT, mT_over_pT = mT / pT;
    const double Estar_M_mT = Estar * M_mT, Estar2_plus_pT2 = Estar2 + pT2, Estar_M_over_pT = Estar * M_over_pT;
    const double DeltaY = log((pstar + sqrt(Estar2_plus_pT2)) / mT);
    const double phip = G.phi[iphip], y = (G.dim == 2) ? 0.0 : G.y[iy];
    const double Ymax = (G.dim == 2) ? 0.0 : fabs(G.y[G.y_pts - 1]);
    double v_integral = 0.0;
    for (int iv = 0; iv < GP; iv++) {
      const double v = kGLRoot[iv];
      int iYL = 0, iYR = 0; double YL = 0.0, YR = 0.0, Y = 0.0; bool cutoff_Y = false;
      if (G.dim == 3) { Y = y + v * DeltaY; cutoff_Y = y_points(G, Y, Ymax, iYL, iYR, YL, YR); }
      const double coshvDeltaY = cosh(v * DeltaY);
      // the boost-invariant branch multiplies mT^2 cosh cosh left to right (:1167), the 3+1D branch squares cosh first (:1329-1330)
      const double mT2_coshvDeltaY2 = (G.dim == 2) ? mT2 * coshvDeltaY * coshvDeltaY : mT2 * (coshvDeltaY * coshvDeltaY);
      const double den = mT2_coshvDeltaY2 - pT2;
      const double mTc = mT_over_pT * coshvDeltaY;
      const double MTbar = Estar_M_mT * coshvDeltaY / den;
      const double DeltaMT = M_pT * sqrt(fabs(Estar2_plus_pT2 - mT2_coshvDeltaY2)) / den;
      const double vw = DeltaY * kGLWeight[iv] / sqrt(fabs(den));
      const double zi = zeta_integral_at(P, MTbar, DeltaMT, mTc, Estar_M_over_pT, parent_mass2, phip, G.dim, cutoff_Y, iYL, iYR, YL, YR, Y);
      v_integral += vw * zi;
    }
    contrib = s_integrand_weight * v_integral;
  }
  part[lb][is] = contrib;
  __syncthreads();
  if (valid && is == 0) {
    double decay3D_integral = 0.0;
    for (int k = 0; k < GP; k++) decay3D_integral += part[lb][k];          // s nodes in order, like the reference's loop
    scratch[(size_t)term * n_bins + bin] = prefactor * decay3D_integral;
  }
}

// dN[daughter] += scratch[term], terms in (channel, group) order
__global__ void decay_add_kernel(DecayGrid G, const DecayTerm *__restrict__ terms, int n_terms, const double *__restrict__ scratch, double *__restrict__ dN)
{
  const int n_bins = G.n_pT * G.n_phi * G.y_pts;
  const int bin = blockIdx.x * blockDim.x + threadIdx.x;
  if (bin >= n_bins) return;
  for (int t = 0; t < n_terms; t++) dN[(int64_t)terms[t].daughter + (int64_t)G.n_species * bin] += scratch[(size_t)t * n_bins + bin];
}

// calculate_Q_factor, :99-121
double q_factor(double mass_parent, double mass_1, double mass_2, double mass_3)
{
  static const double x_root[24] = {-0.99518721999702,-0.97472855597131,-0.93827455200273,-0.8864155270044,-0.8200019859739,-0.74012419157855,-0.64809365193698,-0.54542147138884,-0.43379350762605,-0.31504267969616,-0.19111886747362,-0.064056892862606,0.06405689286261,0.19111886747362,0.31504267969616,0.43379350762605,0.54542147138884,0.64809365193698,0.74012419157855,0.8200019859739,0.8864155270044,0.93827455200273,0.97472855597131,0.99518721999702};
  static const double x_weight[24] = {0.01234122979999,0.02853138862893,0.0442774388174,0.059298584915437,0.0733464814111,0.08619016153195,0.0976186521041,0.107444270116,0.11550566805373,0.1216704729278,0.12583745634683,0.1279381953468,0.1279381953468,0.1258374563468,0.1216704729278,0.1155056680537,0.107444270116,0.09761865210411,0.08619016153195,0.07334648141108,0.05929858491544,0.04427743881742,0.02853138862893,0.01234122979999};
  const double a = (mass_parent + mass_1) * (mass_parent + mass_1), b = (mass_parent - mass_1) * (mass_parent - mass_1);
  const double c = (mass_2 + mass_3) * (mass_2 + mass_3), d = (mass_2 - mass_3) * (mass_2 - mass_3);
  double Q = 0.0;
  for (int i = 0; i < 24; i++) {
    const double s = c + (b - c) * (1.0 + x_root[i]) / 2.0;
    Q += x_weight[i] * (b - c) * std::sqrt(std::fabs((a - s) * (b - s) * (s - c) * (s - d))) / (2.0 * s);
  }
  return Q;
}

struct Host {                       // bookkeeping helpers, all on the host and in integers / the reference's scalar order
  const is3d_particle_list *pdg; int n_chosen; const int32_t *chosen;
  int particle_index(int mc_id, std::string *err) const                           // :59-79
  {
    if (mc_id == 0) { *err = "a decay product has mc_id 0 (null particle in the particle list)"; return -1; }
    for (int i = 0; i < pdg->n_particles; i++) if (pdg->mcid[i] == mc_id) return i;
    *err = "decay product " + std::to_string(mc_id) + " is not in the particle list"; return -1;
  }
  int chosen_index(int pdg_index, std::string *err) const                          // :82-97
  {
    for (int i = 0; i < n_chosen; i++) if (chosen[i] == pdg_index) return i;
    *err = "daughter is not a chosen particle"; return -1;
  }
  // daughters that are chosen species, grouped by type in order of first appearance (:307-371, 828-899)
  int group(const int *prod, int nprod, int *groups, int *members) const
  {
    bool found[3] = {false, false, false};
    for (int ic = 0; ic < n_chosen; ic++) {
      bool all = true;
      for (int k = 0; k < nprod; k++) { if (prod[k] == chosen[ic] && !found[k]) found[k] = true; all = all && found[k]; }
      if (all) break;
    }
    int ng = 0;
    for (int k = 0; k < nprod; k++) {
      if (!found[k]) continue;
      bool put = false;
      for (int gi = 0; gi < ng; gi++) if (prod[k] == groups[gi]) { members[gi] += 1; put = true; break; }
      if (!put) { groups[ng] = prod[k]; members[ng] = 1; ng++; }
    }
    return ng;
  }
};

}  // namespace

// Host driver.  dN_dev: device pointer to the spectra [y][phi][pT][species], amended in place.
int resonance_decays_device(const is3d_particle_list *pdg, int n_chosen, const int32_t *chosen, const is3d_grid *gr, int dimension,
                            double *dN_dev, cudaStream_t st, int *launches, std::string *err)
{
  if (n_chosen - 1 <= 0) { *err = "need at least two chosen particles for the resonance decay routine"; return IS3D_ERR_ARGUMENT; }
  DecayGrid G; memset(&G, 0, sizeof(G));
  G.n_species = n_chosen; G.n_pT = gr->n_pT; G.n_phi = gr->n_phi; G.n_y_tab = gr->n_y; G.dim = dimension;
  G.y_pts = (dimension == 2) ? 1 : gr->n_y;
  if (G.n_pT < 2 || G.n_phi < 2 || (dimension == 3 && G.y_pts < 2)) { *err = "momentum tables too short for the interpolation"; return IS3D_ERR_ARGUMENT; }
  const int n_bins = G.n_pT * G.n_phi * G.y_pts;
  Host H{pdg, n_chosen, chosen};

  // device buffers (freed on every path out of this function)
  struct Buffers {
    double *tables = nullptr, *logdN = nullptr, *MT = nullptr, *scratch = nullptr; Fit *fit = nullptr; DecayTerm *terms = nullptr; int *ids = nullptr, *error = nullptr;
    ~Buffers() { cudaFree(tables); cudaFree(logdN); cudaFree(MT); cudaFree(scratch); cudaFree(fit); cudaFree(terms); cudaFree(ids); cudaFree(error); }
  } B;
  const int max_terms = 50 * 3;                       // <= 50 channels per particle (readindata.h Maxdecaychannel) x 3 daughter groups
  auto ck = [&](cudaError_t e, const char *what) { if (e != cudaSuccess) { *err = std::string(what) + ": " + cudaGetErrorString(e); return false; } return true; };
  if (!ck(cudaMalloc(&B.tables, sizeof(double) * (size_t)(G.n_pT + G.n_phi + G.n_y_tab + 8)), "cudaMalloc") ||
      !ck(cudaMalloc(&B.logdN, sizeof(double) * (size_t)n_bins), "cudaMalloc") ||
      !ck(cudaMalloc(&B.MT, sizeof(double) * (size_t)max_terms * G.n_pT), "cudaMalloc") ||
      !ck(cudaMalloc(&B.fit, sizeof(Fit) * (size_t)max_terms * G.y_pts * G.n_phi), "cudaMalloc") ||
      !ck(cudaMalloc(&B.scratch, sizeof(double) * (size_t)max_terms * n_bins), "cudaMalloc") ||
      !ck(cudaMalloc(&B.terms, sizeof(DecayTerm) * (size_t)max_terms), "cudaMalloc") ||
      !ck(cudaMalloc(&B.ids, sizeof(int) * (size_t)max_terms * 2), "cudaMalloc") ||
      !ck(cudaMalloc(&B.error, sizeof(int)), "cudaMalloc")) return IS3D_ERR_CUDA;
  if (!ck(cudaMemcpyAsync(B.tables, gr->pT, sizeof(double) * G.n_pT, cudaMemcpyHostToDevice, st), "H2D") ||
      !ck(cudaMemcpyAsync(B.tables + G.n_pT, gr->phi, sizeof(double) * G.n_phi, cudaMemcpyHostToDevice, st), "H2D") ||
      !ck(cudaMemcpyAsync(B.tables + G.n_pT + G.n_phi, gr->y, sizeof(double) * G.n_y_tab, cudaMemcpyHostToDevice, st), "H2D") ||
      !ck(cudaMemsetAsync(B.error, 0, sizeof(int), st), "memset")) return IS3D_ERR_CUDA;
  G.pT = B.tables; G.phi = B.tables + G.n_pT; G.y = B.tables + G.n_pT + G.n_phi;

  for (int ichosen = n_chosen - 1; ichosen > 0; ichosen--) {
    const int ipart = chosen[ichosen];
    if (pdg->stable[ipart]) continue;
    // ---- the parent's channels -> terms (host; resonance_decay_channel + the set-up halves of two_/three_body_decay)
    std::vector<DecayTerm> terms;
    for (int ich = 0; ich < pdg->decays[ipart]; ich++) {
      const int row = pdg->dec_first[ipart] + ich;
      const int decay_products = std::abs(pdg->dec_npart[row]);
      if (decay_products > 5) { *err = "a decay channel lists more than five products"; return IS3D_ERR_ARGUMENT; }
      int idx[5];
      for (int k = 0; k < decay_products; k++) { idx[k] = H.particle_index(pdg->dec_part[row * 5 + k], err); if (idx[k] < 0) return IS3D_ERR_ARGUMENT; }
      if (decay_products == 1 || decay_products == 4) continue;              // trivial; 4-body channels are skipped by the reference (:279-282)
      if (decay_products != 2 && decay_products != 3) { *err = "number of decay products = 0 or > 4"; return IS3D_ERR_ARGUMENT; }
      const double branch_ratio = pdg->dec_br[row];
      int groups[3], members[3];
      if (decay_products == 2) {
        double mass_parent = pdg->mass[ipart], mass_1 = pdg->mass[idx[0]], mass_2 = pdg->mass[idx[1]];
        while ((mass_1 + mass_2) > mass_parent) {                             // :241-256
          mass_parent += 0.25 * pdg->width[ipart];
          mass_1 -= 0.5 * pdg->width[idx[0]];
          mass_2 -= 0.5 * pdg->width[idx[1]];
          if (mass_1 < 0.0 || mass_2 < 0.0) { *err = "one daughter mass went negative while enforcing energy conservation"; return IS3D_ERR_ARGUMENT; }
        }
        const int ng = H.group(idx, 2, groups, members);
        for (int gi = 0; gi < ng; gi++) {
          DecayTerm T; memset(&T, 0, sizeof(T));
          T.body = 2; T.daughter = H.chosen_index(groups[gi], err);
          if (T.daughter < 0) return IS3D_ERR_ARGUMENT;
          const double mass = pdg->mass[groups[gi]];
          const double mass_secondary = pdg->mass[idx[1]];                      // always the second product (:411-413)
          const double W2 = mass_secondary * mass_secondary;
          const double Estar = (mass_parent * mass_parent + mass * mass - W2) / (2.0 * mass_parent);
          T.mass_parent = mass_parent; T.mass2 = mass * mass; T.Estar = Estar; T.pstar = std::sqrt(Estar * Estar - mass * mass);
          T.prefactor = (double)members[gi] * mass_parent * branch_ratio / (8.0 * T.pstar);
          terms.push_back(T);
        }
      } else {
        const double mass_parent = pdg->mass[ipart];
        const int ng = H.group(idx, 3, groups, members);
        for (int gi = 0; gi < ng; gi++) {
          DecayTerm T; memset(&T, 0, sizeof(T));
          T.body = 3; T.daughter = H.chosen_index(groups[gi], err);
          if (T.daughter < 0) return IS3D_ERR_ARGUMENT;
          int rest[2], nr = 0; bool removed = false;
          for (int k = 0; k < 3; k++) { if (!removed && idx[k] == groups[gi]) { removed = true; continue; } rest[nr++] = idx[k]; }
          const double mass_1 = pdg->mass[groups[gi]], mass_2 = pdg->mass[rest[0]], mass_3 = pdg->mass[rest[1]];
          const double Q_norm = q_factor(mass_parent, mass_1, mass_2, mass_3);
          T.mass_parent = mass_parent; T.m1sq = mass_1 * mass_1;
          T.s_plus = (mass_parent - mass_1) * (mass_parent - mass_1); T.s_minus = (mass_2 + mass_3) * (mass_2 + mass_3);
          T.d = (mass_2 - mass_3) * (mass_2 - mass_3);
          T.prefactor = (double)members[gi] * (mass_parent * mass_parent) * (T.s_plus - T.s_minus) * branch_ratio / (8.0 * Q_norm);
          terms.push_back(T);
        }
      }
    }
    if (terms.empty()) continue;
    if ((int)terms.size() > max_terms) { *err = "too many decay terms for one parent"; return IS3D_ERR_ARGUMENT; }
    std::vector<int> ids2, ids3;
    for (int t = 0; t < (int)terms.size(); t++) (terms[t].body == 2 ? ids2 : ids3).push_back(t);
    std::vector<int> ids(ids2); ids.insert(ids.end(), ids3.begin(), ids3.end());
    // pageable host memory: these copies complete before cudaMemcpyAsync returns, so the vectors may go out of scope
    if (!ck(cudaMemcpyAsync(B.terms, terms.data(), sizeof(DecayTerm) * terms.size(), cudaMemcpyHostToDevice, st), "H2D") ||
        !ck(cudaMemcpyAsync(B.ids, ids.data(), sizeof(int) * ids.size(), cudaMemcpyHostToDevice, st), "H2D")) return IS3D_ERR_CUDA;
    log_parent_kernel<<<(n_bins + 255) / 256, 256, 0, st>>>(G, dN_dev, ichosen, B.logdN);
    decay_setup_kernel<<<(unsigned)terms.size(), 128, 0, st>>>(G, B.terms, B.logdN, B.MT, B.fit, B.error);
    *launches += 2;
    if (!ids2.empty()) {
      const int64_t n = (int64_t)ids2.size() * n_bins;
      decay2_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(G, B.terms, B.ids, (int)ids2.size(), B.logdN, B.MT, B.fit, B.scratch);
      (*launches)++;
    }
    if (!ids3.empty()) {
      const int64_t n = (int64_t)ids3.size() * n_bins;
      decay3_kernel<<<(unsigned)((n + kBins3 - 1) / kBins3), dim3(GP, kBins3), 0, st>>>(G, B.terms, B.ids + ids2.size(), (int)ids3.size(), B.logdN, B.MT, B.fit, B.scratch);
      (*launches)++;
    }
    decay_add_kernel<<<(n_bins + 255) / 256, 256, 0, st>>>(G, B.terms, (int)terms.size(), B.scratch, dN_dev);
    (*launches)++;
    if (!ck(cudaGetLastError(), "decay kernels")) return IS3D_ERR_CUDA;
    // the term list of the next parent reuses B.terms: the stream must be done with it (also surfaces kernel faults per parent)
    if (!ck(cudaStreamSynchronize(st), "decay kernels")) return IS3D_ERR_CUDA;
  }
  int flag = 0;
  if (!ck(cudaMemcpyAsync(&flag, B.error, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H") || !ck(cudaStreamSynchronize(st), "sync")) return IS3D_ERR_CUDA;
  if (flag) { *err = "not enough positive points of a parent spectrum beyond mT = 1.65 M to fit its large-mT tail (the reference exits here)"; return IS3D_ERR_ARGUMENT; }
  return IS3D_OK;
}

}  // namespace is3d
