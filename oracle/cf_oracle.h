/* oracle/cf_oracle.h -- CPU restatement of iS3D's smooth Cooper-Frye path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (is3d_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED for df_mode 1-4 (mode 1 surfaces, 3+1D and 2+1D) against the unmodified reference sources
 * compiled here (oracle/_ref/is3d_ref, see oracle/Makefile) -- tests/test_oracle_vs_reference.py and the committed
 * vectors under tests/golden/.  The anisotropic kernel (cf_oracle_smooth_vah) is pinned against a direct call of the
 * reference's (otherwise uncalled) calculate_dN_pTdpTdphidy_VAH_PL; its (Lambda, alpha_L) coefficient lookup is pinned against
 * the reference's only reader of those tables, src/cuda/deltafReader.cu:192-277, compiled into oracle/_ref/vah_ref
 * (tests/golden/vah_coefficients.npz).  The resonance-decay routine (cf_decays.c) is pinned against the reference's routine run
 * behind oracle/ref_decays_prefix.h (its author disabled it with an exit(-1) at entry).
 * GSL is absent from the image: the natural cubic spline and the 3x3 LU inverse restate GSL's published
 * algorithms (see oracle/gsl_shim); "parity unpinned" against a real libgsl at the 1e-16 level.
 */
#ifndef CF_ORACLE_H
#define CF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int64_t n_cells;
  const double *tau, *eta, *dat, *dax, *day, *dan, *ux, *uy, *un, *T, *P, *E;
  const double *pixx, *pixy, *pixn, *piyy, *piyn, *bulkPi;
  const double *muB, *nB, *Vx, *Vy, *Vn;
  /* anisotropic (mode 2) only */
  const double *pitt, *pitx, *pity, *pitn, *pinn, *Wx, *Wy, *Lambda, *aL, *c0, *c1, *c2, *c3, *c4;
} cfo_cells;

typedef struct {
  int32_t n;
  const double *mass, *sign, *degeneracy, *baryon;
} cfo_species;

typedef struct {
  int32_t n_pT, n_phi, n_y, n_eta;       /* table lengths (y table length even in 2+1D) */
  const double *pT, *phi, *phi_weight, *y, *eta, *eta_weight;
} cfo_grid;

typedef struct {
  int32_t df_mode, dimension, include_baryon, include_bulk, include_shear, include_diff, regulate_deltaf, outflow;
  double deta_min, mass_pion0;
} cfo_flags;

/* muB = 0 rows of the coefficient tables + the Jonah lambda/z tables */
typedef struct {
  int32_t n_T;
  const double *T, *c0, *c1, *c2, *c3, *c4, *F, *G, *betabulk, *betaV, *betapi;
  int32_t n_jonah;
  const double *jonah_x, *jonah_lambda2, *jonah_z;
  double bulkPi_over_Peq_max;
} cfo_df_tables;

typedef struct {
  int32_t n_points;                      /* Gauss-Laguerre points per alpha */
  const double *root1, *weight1, *root2, *weight2;
} cfo_laguerre;

/* per-cell coefficient struct, reference readindata.h:105-131 */
typedef struct {
  double c0, c1, c2, c3, c4, shear14_coeff, F, G, betabulk, betaV, betapi, lambda, z, delta_lambda, delta_z;
} cfo_dfcoef;

/* natural cubic spline (GSL cspline restated): c has n entries */
void cfo_spline_init(const double *x, const double *y, int n, double *c);
double cfo_spline_eval(const double *x, const double *y, const double *c, int n, double xv, int *err);

/* Deltaf_Data::cubic_spline, deltafReader.cpp:325-395; returns nonzero if T or Pi/P is outside the table */
int cfo_df_coefficients(const cfo_df_tables *tab, int df_mode, double T, double E, double P, double bulkPi, cfo_dfcoef *out);

/* Deltaf_Data::compute_jonah_coefficients, deltafReader.cpp:222-297: fills x[301], lambda2[301], z[301]; returns max x */
double cfo_jonah_tables(int n_particles, const double *mass, const double *degeneracy, const double *sign, double T_avg,
                        const cfo_laguerre *gla, double *x, double *lambda2, double *z);

/* surface averages as in FO_data_reader::read_surf_VH, readindata.cpp:423-466: out[5] = T,E,P,muB,nB */
void cfo_surface_averages(const cfo_cells *c, double *out5);

/* EmissionFunctionArray::calculate_dN_pTdpTdphidy, emissionfunction_smooth_kernels.cpp:28-393 (df_mode 1,2).
 * dN is [n_y_tab][n_phi][n_pT][n_species] (species fastest) and is ADDED into.  Returns #cells skipped (u.dsigma<=0)
 * or a negative error code.
 * dN_abs (optional, may be NULL) receives, per bin, the same sum with every term of delta-f replaced by its absolute value:
 * the bin's floating-point noise floor is a few ulp of dN_abs, which matters where 1 + df nearly cancels. */
int64_t cfo_smooth_vh(const cfo_flags *fl, const cfo_cells *c, const cfo_species *sp, const cfo_grid *g,
                      const cfo_df_tables *tab, double *dN, double *dN_abs);

/* EmissionFunctionArray::calculate_dN_ptdptdphidy_feqmod, :396-996 (df_mode 3,4).  *breakdown gets the cell count. */
int64_t cfo_smooth_feqmod(const cfo_flags *fl, const cfo_cells *c, const cfo_species *sp, const cfo_grid *g,
                          const cfo_df_tables *tab, const cfo_laguerre *gla, double *dN, int64_t *breakdown);

/* EmissionFunctionArray::calculate_dN_pTdpTdphidy_VAH_PL, :2140-2393 */
int64_t cfo_smooth_vah(const cfo_flags *fl, const cfo_cells *c, const cfo_species *sp, const cfo_grid *g, double *dN);

/* operation = 0: spacetime distributions.  x, y are the transverse cell positions; the result arrays are
 * the RAW sums the reference accumulates before its writers divide by the bin widths (zero-initialised by the caller):
 *   dN_tau [n_species][tau_bins], dN_r [n_species][r_bins], dN_taur [n_species][tau_bins][r_bins],
 *   dN_dydeta [n_species][eta_pts] (eta_pts = 1 in 3+1D, n_eta in 2+1D), dN_dy [n_species]. */
typedef struct {
  double tau_min, tau_max, r_min, r_max;
  int32_t tau_bins, r_bins;
  const double *x, *y;                   /* [n_cells] */
  const double *pT_weight;               /* [n_pT] */
} cfo_spacetime_spec;

/* EmissionFunctionArray::calculate_dN_dX, emissionfunction_smooth_kernels.cpp:1000-1446 (df_mode 1,2).
 * Returns #cells skipped (u.dsigma <= 0) or a negative error code. */
int64_t cfo_spacetime_vh(const cfo_flags *fl, const cfo_cells *c, const cfo_species *sp, const cfo_grid *g,
                         const cfo_df_tables *tab, const cfo_spacetime_spec *spec,
                         double *dN_tau, double *dN_r, double *dN_taur, double *dN_dydeta, double *dN_dy);

/* EmissionFunctionArray::calculate_dN_dX_feqmod, :1449-2135 (df_mode 3,4); *breakdown gets the number of cells that break down
 * (the reference's printed counter is n_species times that, it is incremented inside the species loop). */
int64_t cfo_spacetime_feqmod(const cfo_flags *fl, const cfo_cells *c, const cfo_species *sp, const cfo_grid *g,
                             const cfo_df_tables *tab, const cfo_laguerre *gla, const cfo_spacetime_spec *spec,
                             double *dN_tau, double *dN_r, double *dN_taur, double *dN_dydeta, double *dN_dy, int64_t *breakdown);

/* Sampler mean yield: Deltaf_Data::compute_particle_densities (deltafReader.cpp:536-650) and
 * EmissionFunctionArray::calculate_total_yield (emissionfunction_sampling_kernels.cpp:653-831), include_baryon = 0. */
int cfo_particle_densities(int n, const double *mass, const double *degeneracy, const double *baryon, const double *sign,
                           const double *avg5, int df_mode, const cfo_df_tables *tab, const cfo_laguerre *gla,
                           const double *root3, const double *weight3, double *neq_out, double *bulk_out, double *diff_out);
int64_t cfo_total_yield(const cfo_flags *fl, const cfo_cells *c, int n_species, const double *neq, const double *bulk,
                        const cfo_df_tables *tab, double y_cut, double *Ntot_out);

/* Resonance-decay feed-down: EmissionFunctionArray::do_resonance_decays and everything below it,
 * emissionfunction_resonance_decays.cpp:124-2158 (oracle/cf_decays.c).  The particle list is the reference's particle_info array
 * (readindata.cpp:1440-1568: anti-baryons synthesised, decay channels flattened: channel rows dec_first[i] .. + decays[i]).
 * dN [y][phi][pT][chosen species] is amended in place; returns 0 or a negative error code (the reference exits there). */
typedef struct {
  int32_t n_particles;
  const int32_t *mcid; const double *mass, *width; const int32_t *stable, *decays, *dec_first;
  const int32_t *dec_npart; const double *dec_br; const int32_t *dec_part;       /* dec_part[row * 5 + k] */
} cfo_particles;
int cfo_resonance_decays(const cfo_particles *pdg, int32_t n_chosen, const int32_t *chosen_pdg_index, const cfo_grid *g, int32_t dimension,
                         double *dN);

/* VAH helpers: aL_fit / R200 (arsenal.cpp:999-1066) and the (Lambda, aL) bilinear lookup of
 * src/cuda/deltafReader.cu:192-277 */
double cfo_aL_fit(double pl_over_peq);
double cfo_R200(double aL);

#ifdef __cplusplus
}
#endif
#endif
