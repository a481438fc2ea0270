/* include/is3d_b200.h -- C ABI of the B200-native smooth Cooper-Frye spectra path.
 *
 * The reference (derekeverett/iS3D) has no FFI/plugin interface; the hot path sits behind the C++ member functions
 *   EmissionFunctionArray::calculate_dN_pTdpTdphidy          src/cpp/emissionfunction.h:179, emissionfunction_smooth_kernels.cpp:28
 *   EmissionFunctionArray::calculate_dN_ptdptdphidy_feqmod   src/cpp/emissionfunction.h:182, emissionfunction_smooth_kernels.cpp:396
 *   EmissionFunctionArray::calculate_dN_pTdpTdphidy_VAH_PL   src/cpp/emissionfunction.h:(VAH_PL), emissionfunction_smooth_kernels.cpp:2140
 * called from EmissionFunctionArray::calculate_spectra (src/cpp/emissionfunction.cpp:1519, 1584, 1650).  The entry
 * points below take what those calls take -- the per-species arrays, the structure-of-arrays freeze-out surface, the
 * momentum tables and the delta-f coefficient tables -- as plain pointers and sizes, and add the result into a
 * caller-owned spectra array with the reference's layout.  INTEGRATION.md shows the few lines that replace the three
 * call sites.  No C++/torch types cross this boundary; all functions return 0 or an IS3D_ERR_* code, never exit().
 * Further down: operation = 0 (calculate_dN_dX{,_feqmod}, smooth_kernels.cpp:1000-2135), the sampler's mean-yield pass
 * (calculate_total_yield, emissionfunction_sampling_kernels.cpp:653-831) and the file-level host layer.
 *
 * Threading: every entry point works on the CUDA device that is current in the calling thread and keeps one workspace per
 * device, so different host threads may drive different devices concurrently (calls on the same device are serialised).
 * Multi-GPU: either one process per GPU, each calling with its own contiguous shard of cells and all-reducing the spectra
 * itself (is3d_b200/distributed.py under torchrun), or ONE process calling the *_multi entry points below.
 */
#ifndef IS3D_B200_H
#define IS3D_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  IS3D_OK = 0,
  IS3D_ERR_ARGUMENT = 1,      /* NULL / inconsistent argument */
  IS3D_ERR_UNSUPPORTED = 2,   /* flag combination outside the path (e.g. include_baryon = 1) */
  IS3D_ERR_TABLE_RANGE = 3,   /* a cell's T (or Pi/P for df_mode 4) lies outside the coefficient table; the reference aborts here */
  IS3D_ERR_CUDA = 4,          /* CUDA runtime failure; is3d_b200_last_error() has the text */
  IS3D_ERR_NO_DEVICE = 5,     /* no CUDA device: there is deliberately no CPU fallback */
  IS3D_ERR_IO = 6,            /* host layer: missing / malformed input file */
  IS3D_ERR_NCCL = 7           /* multi-GPU: libnccl.so.2 could not be loaded, or an NCCL call failed */
};

/* Freeze-out surface, structure of arrays, GeV / fm units exactly as EmissionFunctionArray::calculate_spectra packs
 * them (emissionfunction.cpp:1327-1499).  Pointers for switched-off terms may be NULL.  Read-only, caller-owned. */
typedef struct {
  int64_t n_cells;
  const double *tau, *eta, *dat, *dax, *day, *dan, *ux, *uy, *un, *T, *P, *E;
  const double *pixx, *pixy, *pixn, *piyy, *piyn, *bulkPi;
  const double *muB, *nB, *Vx, *Vy, *Vn;
  /* anisotropic hydro (mode 2) only: all ten pi_perp components, W_perp, Lambda, alpha_L and per-cell c0..c4 */
  const double *pitt, *pitx, *pity, *pitn, *pinn, *Wx, *Wy, *Lambda, *aL, *c0, *c1, *c2, *c3, *c4;
  /* transverse cell positions: read by operation = 0 (spacetime distributions) only, NULL otherwise */
  const double *x, *y;
} is3d_surface;

/* Chosen species in output order (emissionfunction.cpp:1293-1307). */
typedef struct {
  int32_t n;
  const double *mass, *sign, *degeneracy, *baryon;
} is3d_species;

/* Momentum / rapidity tables (iS3D.cpp:161-167).  eta/eta_weight are used when dimension = 2, y when dimension = 3. */
typedef struct {
  int32_t n_pT, n_phi, n_y, n_eta;
  const double *pT, *phi, *y, *eta, *eta_weight;
} is3d_grid;

/* Switches captured by the EmissionFunctionArray constructor (emissionfunction.cpp:170-191). */
typedef struct {
  int32_t mode;              /* surface type: 0,1,4,5,6,7 = viscous hydro; 2 = anisotropic hydro, PL matching */
  int32_t df_mode;           /* 1 14-moment, 2 Chapman-Enskog, 3 feqmod (Mike), 4 feqmod (Jonah) */
  int32_t dimension;         /* 2 boost invariant (y = 0, eta quadrature) or 3 */
  int32_t include_baryon, include_bulk_deltaf, include_shear_deltaf, include_baryondiff_deltaf;
  int32_t regulate_deltaf, outflow;
  double deta_min, mass_pion0;
} is3d_flags;

/* muB = 0 rows of deltaf_coefficients/vh/<eos>/{c0..c4,F,G,betabulk,betaV,betapi}.dat as stored in the files
 * (T-scaled; deltafReader.cpp:65-219), plus the Jonah lambda^2 / z tables (deltafReader.cpp:222-297). */
typedef struct {
  int32_t n_T;
  const double *T, *c0, *c1, *c2, *c3, *c4, *F, *G, *betabulk, *betaV, *betapi;
  int32_t n_jonah;
  const double *jonah_x, *jonah_lambda2, *jonah_z;
  double bulkPi_over_Peq_max;
} is3d_df_tables;

/* generalized Gauss-Laguerre nodes alpha = 1, 2 (tables/gla_roots_weights_32_points.txt), feqmod only */
typedef struct {
  int32_t n_points;
  const double *root1, *weight1, *root2, *weight2;
} is3d_laguerre;

typedef struct {
  int32_t memory;            /* 0: surface arrays and dN live in host memory; 1: they are device pointers on the current device */
  void *stream;              /* cudaStream_t to launch on (NULL = default stream) */
  int32_t n_chunks;          /* cell-range split used for load balance; 0 = choose */
  int32_t tile_variant;      /* kernel variant: 0 = tuned default for the model (df_mode 1 / 2 and ideal f_eq on 3+1D tiles with <= 64 pT
                                points: shifted-factor kernel 23; everything else: cf_kernel); 1..16 = register-tile table of cf_kernel;
                                17..21 = shapes of the factored kernel (df_mode 1/2, 3+1D, >= 16 species); 22..25 = shapes of the
                                shifted-factor kernel (df_mode 1/2, 3+1D, <= 64 pT points); 99 = strict diagnostic kernel (df_mode 1/2 in
                                the reference's own operation order, one thread per bin: ~50x slower, tests only).  A variant that does
                                not apply to the call returns IS3D_ERR_ARGUMENT. */
  int32_t reserved[4];
} is3d_options;

typedef struct {
  int64_t cells_skipped_udsigma;   /* cells with u.dsigma <= 0 (smooth_kernels.cpp:137) */
  int64_t cells_feqmod_breakdown;  /* the reference's `breakdown` counter (smooth_kernels.cpp:721, 992) */
  int64_t evaluations;             /* cells x species x pT x phi x y x eta */
  double h2d_ms, prepare_ms, kernel_ms, reduce_ms, d2h_ms, total_ms;   /* CUDA-event times on the launch stream */
  int32_t gpu_launches;            /* kernels launched by this call */
  int32_t n_chunks, tile_variant;
  int32_t n_gpus;                  /* devices that worked on this call (times above: slowest device) */
  double allreduce_ms;             /* multi-GPU: the NCCL all-reduce of the spectra array */
  int32_t n_chunks_wanted;         /* chunks the load-balance rule asked for; > n_chunks when the 4 GiB cap on the partial-sum buffer cut it down */
  int32_t reserved;
} is3d_stats;

/* Bind to the current CUDA device and create its workspace (one per CUDA ordinal).  Safe to call more than once. */
int is3d_b200_init(void);
/* ---- one process, several GPUs (the reference has ONE entry, IS3D::run_particlization, src/cpp/iS3D.cpp:73-191;
 * a drop-in for it has to use the whole box by itself).  is3d_b200_init_devices(n) selects the first n visible devices (n <= 0: all
 * of them, or the number in the environment variable IS3D_B200_GPUS) and creates one NCCL communicator over them (libnccl.so.2 is
 * loaded at run time).  is3d_b200_smooth_spectra_multi() then takes HOST arrays, splits the cells into contiguous shards of
 * ceil(n_cells / n) (species, grids and coefficient tables are replicated), runs one host thread and one stream per device, combines
 * the per-device spectra with a single ncclAllReduce(sum, double) and adds device 0's copy into dN_out.  With one device it is
 * is3d_b200_smooth_spectra().  is3d_b200_run_workdir / is3d_b200_run_surface and the IS3D class use it, so the file-level
 * drop-in runs on every visible GPU.  Results differ from the 1-GPU result by summation order only (non-negative terms: ~1e-15). */
int is3d_b200_init_devices(int n_gpus);
int is3d_b200_device_count(void);     /* devices the *_multi entry points use (1 before is3d_b200_init_devices) */
int is3d_b200_shutdown(void);
const char *is3d_b200_strerror(int code);
const char *is3d_b200_last_error(void);
int is3d_b200_version(void);

/* dN_out has n_species * n_pT * n_phi * n_y doubles, index ipart + n_species*(ipT + n_pT*(iphi + n_phi*iy))
 * (emissionfunction_smooth_kernels.cpp:363); the result is ADDED into it (reference: `+=` into a zeroed array).
 * Dispatch follows calculate_spectra: mode 2 -> anisotropic kernel; df_mode 1,2 -> linear delta-f; 3,4 -> feqmod. */
int is3d_b200_smooth_spectra(const is3d_flags *flags, const is3d_surface *surface, const is3d_species *species,
                             const is3d_grid *grid, const is3d_df_tables *df, const is3d_laguerre *laguerre,
                             const is3d_options *options, double *dN_out, is3d_stats *stats);

int is3d_b200_smooth_spectra_multi(const is3d_flags *flags, const is3d_surface *surface, const is3d_species *species,
                                   const is3d_grid *grid, const is3d_df_tables *df, const is3d_laguerre *laguerre,
                                   const is3d_options *options, double *dN_out, is3d_stats *stats);

/* ---- operation = 0: spacetime distributions of the momentum-integrated yield -------------------
 * Replaces EmissionFunctionArray::calculate_dN_dX (emissionfunction_smooth_kernels.cpp:1000-1446, df_mode 1, 2) and
 * calculate_dN_dX_feqmod (:1449-2135, df_mode 3, 4), called from calculate_spectra (emissionfunction.cpp:1514, 1579).
 * Every cell's yield  sum_{pT, phi, y} w_pT w_phi g/(2 pi hbar c)^3 sum_eta p.dsigma f  (3+1D: all y points, unweighted,
 * like the reference) is binned by the cell's tau and r = sqrt(x^2 + y^2) (is3d_surface.x, .y):
 *    itau = floor((tau - tau_min) / ((tau_max - tau_min) / tau_bins)),  ir likewise (:1376-1400).
 * Results are the RAW sums the reference accumulates; its writers divide by the bin volumes (:1404-1435), see
 * is3d_b200_write_spacetime().  Arrays are caller-allocated host memory and are OVERWRITTEN:
 *    dN_tau [n_species][tau_bins]   dN_r [n_species][r_bins]   dN_taur [n_species][tau_bins][r_bins]
 *    dN_dydeta [n_species][eta_pts] (eta_pts = 1 in 3+1D, n_eta in 2+1D)   dN_dy [n_species]                      */
typedef struct {
  double tau_min, tau_max, r_min, r_max;   /* iS3D_parameters.dat: tau_min, tau_max, r_min, r_max */
  int32_t tau_bins, r_bins;                /* tau_bins, r_bins */
  const double *pT_weight, *phi_weight;    /* host: second column of the pT and phi tables */
} is3d_spacetime_bins;
typedef struct { double *dN_tau, *dN_r, *dN_taur, *dN_dydeta, *dN_dy; } is3d_spacetime_result;
int is3d_b200_spacetime_distributions(const is3d_flags *flags, const is3d_surface *surface, const is3d_species *species,
                                      const is3d_grid *grid, const is3d_df_tables *df, const is3d_laguerre *laguerre,
                                      const is3d_spacetime_bins *bins, const is3d_options *options,
                                      is3d_spacetime_result *result, is3d_stats *stats);
/* Same over the devices of is3d_b200_init_devices(): HOST arrays, contiguous cell shards, one thread per device; the histograms are
 * linear in the cells, so the per-device raw sums (a few MB) are added on the host in device order. */
int is3d_b200_spacetime_distributions_multi(const is3d_flags *flags, const is3d_surface *surface, const is3d_species *species,
                                            const is3d_grid *grid, const is3d_df_tables *df, const is3d_laguerre *laguerre,
                                            const is3d_spacetime_bins *bins, const is3d_options *options,
                                            is3d_spacetime_result *result, is3d_stats *stats);

/* ---- resonance-decay feed-down of the smooth spectra ------------------------------------------------
 * Replaces EmissionFunctionArray::do_resonance_decays and the routines below it (src/cpp/emissionfunction_resonance_decays.cpp:
 * 124-2158, called from calculate_spectra, emissionfunction.cpp:1689-1698).  The particle list is the reference's particle_info
 * array (readindata.cpp:1440-1568: anti-baryons right behind their baryon, `stable` = first channel has one product) with the
 * decay channels flattened: channel rows dec_first[i] .. dec_first[i] + decays[i] - 1 of dec_npart / dec_br / dec_part[row * 5 + k].
 * chosen_pdg_index is the reference's chosen_particles_sampling_table (particle-list index of every chosen species).
 * dN [y][phi][pT][species] (host memory, or device memory with options->memory = 1) is amended in place: parents from the last
 * chosen species down to the second, 2- and 3-body channels, daughters that are chosen species.
 * NOTE: the reference snapshot disables its own routine with an exit(-1) at entry (:126-129, the author's note on the MTmax
 * handling of the interpolation); this entry point implements the body behind that guard and is checked against it. */
typedef struct {
  int32_t n_particles;
  const int32_t *mcid;
  const double *mass, *width;
  const int32_t *stable, *decays, *dec_first;
  const int32_t *dec_npart;
  const double *dec_br;
  const int32_t *dec_part;
} is3d_particle_list;
int is3d_b200_resonance_decays(const is3d_particle_list *particles, int32_t n_chosen, const int32_t *chosen_pdg_index,
                               const is3d_grid *grid, int32_t dimension, const is3d_options *options, double *dN, is3d_stats *stats);

/* FP64 FMA peak of the current device measured with a dependency-free DFMA chain (the roofline denominator;
 * MEASURED_PEAKS.json carries no FP64 figure).  Returns TFLOP/s in *tflops, SM clock not touched. */
int is3d_b200_measure_fp64_peak(double *tflops, double *ms);
/* Same probe launched back to back for `seconds` of device time: the sustained (power-capped) FP64 roof. */
int is3d_b200_measure_fp64_sustained(double seconds, double *tflops);

/* ---- host layer: the drop-in behind iS3D_parameters.dat / input/surface.dat / PDG / deltaf_coefficients / tables ----
 * Equivalent of IS3D::run_particlization(1) with operation = 1 (src/cpp/iS3D.cpp:73-191): reads the CWD-relative input
 * files under `workdir`, runs the spectra on the GPU, writes results/dN_pTdpTdphidy*.dat, results/vn_continuous/ and
 * results/dN_dy_*.dat with the reference's formats; with do_resonance_decays = 1 it then runs the feed-down and writes
 * results/dN_pTdpTdphidy_resonance_decays.dat and results/dN_dpTdphidy_resonance_decays.dat (emissionfunction.cpp:452-488, 555-590),
 * and dN_raw receives the amended spectra.  If dN_raw != NULL it receives the spectra (n_raw doubles max).
 * operation = 0 runs the spacetime distributions instead and writes results/spacetime_distribution/dN_taudtaudy_<mcid>.dat,
 * dN_twopirdrdy_<mcid>.dat, dN_twopitaurdtaudrdy_<mcid>.dat, dN_dydeta_<mcid>_<eta_pts>pt.dat (smooth_kernels.cpp:1112-1126,
 * 1404-1435); dN_raw then receives the raw sums concatenated as [dN_tau | dN_r | dN_taur | dN_dydeta | dN_dy]. */
int is3d_b200_run_workdir(const char *workdir, double *dN_raw, int64_t n_raw, int32_t *mcid_out, int32_t n_mcid_max,
                          is3d_stats *stats);
/* Same with the freeze-out cells passed in memory (reference: IS3D::read_fo_surf_from_memory + run_particlization(0),
 * src/cpp/iS3D.cpp:26-71): parameters, particle list and tables still come from `workdir`. */
int is3d_b200_run_surface(const char *workdir, const is3d_surface *surface, double *dN_raw, int64_t n_raw,
                          int32_t *mcid_out, int32_t n_mcid_max, is3d_stats *stats);

/* Host-side table builders for callers that drive is3d_b200_smooth_spectra() directly (no working directory):
 *  - surface averages T, E, P, muB, nB weighted as in FO_data_reader::read_surf_VH (readindata.cpp:423-466), passed through
 *    the reference's 15-significant-digit side file round trip (average_thermodynamic_quantities.dat);
 *  - the Jonah lambda^2(Pi/P), z(Pi/P) tables of Deltaf_Data::compute_jonah_coefficients (deltafReader.cpp:222-297):
 *    301 points each, built from the full particle list at temperature T_avg with the alpha = 2 Gauss-Laguerre nodes. */
int is3d_b200_surface_averages(const is3d_surface *surface, double *out5);
int is3d_b200_jonah_tables(int32_t n_particles, const double *mass, const double *degeneracy, const double *sign, double T_avg,
                           int32_t n_points, const double *root2, const double *weight2,
                           double *x301, double *lambda2_301, double *z301, double *bulkPi_over_Peq_max);
/* Anisotropic hydro, in-memory callers: alpha_L and Lambda [GeV] from the surface file's T, P, PL columns [fm units]
 * (readindata.cpp:905-918), and the per-cell c0..c4 from the (Lambda [fm^-1], alpha_L) tables, each table given as
 * c[iL * naL + iaL] with the values of deltaf_coefficients/vah/c{k}_vah1.dat (src/cuda/deltafReader.cu:192-277). */
int is3d_b200_vah_anisotropy(int64_t n, const double *T_fm, const double *P_fm, const double *PL_fm, double *aL_out, double *Lambda_GeV_out);
int is3d_b200_vah_coefficients(int32_t nL, int32_t naL, const double *L_fm, const double *aL_grid, const double *c0, const double *c1,
                               const double *c2, const double *c3, const double *c4, int64_t n, const double *Lambda_GeV,
                               const double *aL, double *o0, double *o1, double *o2, double *o3, double *o4);
/* ---- sampler mean yield -----------------------------------------------------------------------
 * is3d_b200_particle_densities: n_eq, dn_bulk, dn_diff per species at the surface averages avg5 = (T, E, P, muB, nB) -- what
 * Deltaf_Data::compute_particle_densities (deltafReader.cpp:536-650) stores in particle_info.  Host computation; rootK/weightK are
 * the alpha = K rows of tables/gla_roots_weights_32_points.txt (alpha = 3 only for df_mode 1).
 * is3d_b200_mean_yield: EmissionFunctionArray::calculate_total_yield (emissionfunction_sampling_kernels.cpp:653-831) for the
 * chosen species' densities: sum over cells with u.dsigma > 0 of u.dsigma (n_eq + Pi dn_bulk) (df_mode 1-3) or
 * u.dsigma z(Pi/P) n_eq (df_mode 4), times 2 y_cut in 2+1D.  include_baryon = 0 only.  The surface reduction runs on the GPU. */
int is3d_b200_particle_densities(int32_t n, const double *mass, const double *degeneracy, const double *baryon, const double *sign,
                                 const double *avg5, int32_t df_mode, const is3d_df_tables *df, int32_t n_points,
                                 const double *root1, const double *weight1, const double *root2, const double *weight2,
                                 const double *root3, const double *weight3, double *neq, double *dn_bulk, double *dn_diff);
int is3d_b200_mean_yield(const is3d_flags *flags, const is3d_surface *surface, int32_t n_species, const double *neq,
                         const double *dn_bulk, const is3d_df_tables *df, double y_cut, const is3d_options *options,
                         double *Ntot, is3d_stats *stats);

/* Writers only: produce the results/ files of `workdir` from a spectra array in the reference layout. */
int is3d_b200_write_results(const char *workdir, const double *dN, int64_t n);
/* Host-layer inspection without GPU work: dumps what the readers derived (named double records) to out_path. */
int is3d_b200_host_dump(const char *workdir, const char *out_path);
const char *is3d_b200_host_error(void);

#ifdef __cplusplus
}
#endif
#endif
