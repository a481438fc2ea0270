// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A driver translation unit for the UNMODIFIED reference sources under /root/reference/src/cpp.  It is
// compiled together with them (oracle/Makefile -> oracle/_ref/is3d_ref[_omp]) and exists only so that the
// parity tests and the CPU baseline can (a) call the reference's smooth Cooper-Frye kernels directly and
// time exactly that call, and (b) get the spectra array out as raw doubles (the reference's text writers
// keep 9 significant digits, emissionfunction.cpp:413).
//
// It repeats the set-up sequence of IS3D::run_particlization (src/cpp/iS3D.cpp:81-169) using the reference's
// own public classes, packs the structure-of-arrays inputs as EmissionFunctionArray::calculate_spectra does
// (src/cpp/emissionfunction.cpp:1282-1499), and then calls one of
//   EmissionFunctionArray::calculate_dN_pTdpTdphidy          (emissionfunction_smooth_kernels.cpp:28)
//   EmissionFunctionArray::calculate_dN_ptdptdphidy_feqmod   (:396)
//   EmissionFunctionArray::calculate_dN_pTdpTdphidy_VAH_PL   (:2140, uncalled in the reference itself)
// No reference file is modified or copied; `#define private public` is the only trick used.
//
// usage (run inside a directory laid out like an iS3D checkout):
//   is3d_ref kernel  [out.bin]   hot-path call only, raw dump + timing to stdout as a JSON line
//   is3d_ref full    [out.bin]   calculate_spectra() incl. the reference's text writers, then raw dump
//   is3d_ref yield   [out.bin]   sampler mean yield: calculate_total_yield() with the reference's own species densities
//   is3d_ref vah     [out.bin]   mode-2 surface + input/vah_coefficients.bin (c0..c4 per cell) -> VAH_PL kernel
//   is3d_ref_decays decays [out.bin]   (only in the binary built with oracle/ref_decays_prefix.h) input/spectra_in.bin -> the spectra
//                                array, then EmissionFunctionArray::do_resonance_decays (emissionfunction_resonance_decays.cpp:124)
#include <iostream>
#include <sstream>
#include <fstream>
#include <string>
#include <vector>
#include <random>
#include <complex>
#include <array>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>

#define private public
#include "emissionfunction.h"
#undef private
#include "iS3D.h"
#include "readindata.h"
#include "Table.h"
#include "ParameterReader.h"
#include "deltafReader.h"

using namespace std;

static double *zeros(long n) { return (double *)calloc(n > 0 ? n : 1, sizeof(double)); }

#ifdef IS3D_REF_DECAYS
// see oracle/ref_decays_prefix.h: the first exit() of the resonance-decay translation unit is the author's "unfinished" guard
extern "C" void is3d_ref_exit_hook(int code)
{
  static int swallowed = 0;
  if (!swallowed) { swallowed = 1; return; }
  std::exit(code);
}
#endif

int main(int argc, char **argv)
{
  string what = (argc > 1) ? argv[1] : "kernel";
  string out = (argc > 2) ? argv[2] : "results/dN_raw.bin";

  ParameterReader *paraRdr = new ParameterReader;
  paraRdr->readFromFile("iS3D_parameters.dat");

  FO_data_reader freeze_out_data(paraRdr, "input");
  long FO_length = freeze_out_data.get_number_cells();
  FO_surf *surf = new FO_surf[FO_length];
  memset((void *)surf, 0, sizeof(FO_surf) * FO_length);
  freeze_out_data.read_surf_switch(FO_length, surf);

  particle_info *particle_data = new particle_info[Maxparticle];
  PDG_Data pdg(paraRdr);
  int Nparticle = pdg.read_resonances(particle_data);

  int mode = paraRdr->getVal("mode");
  int df_mode = paraRdr->getVal("df_mode");

  Deltaf_Data *df_data = new Deltaf_Data(paraRdr);
  df_data->load_df_coefficient_data();
  df_data->construct_cubic_splines();
  if (mode != 2 && mode != 3)
  {
    // these read average_thermodynamic_quantities.dat, which the VAH readers never write
    df_data->compute_jonah_coefficients(particle_data, Nparticle);
    df_data->compute_particle_densities(particle_data, Nparticle);
  }

  Table chosen_particles("PDG/chosen_particles.dat");
  Table pT_tab("tables/pT_gauss_legendre_table.dat");
  Table phi_tab("tables/phi_gauss_legendre_table.dat");
  Table y_tab("tables/y_trapezoid_table_21pt.dat");
  Table eta_tab("tables/eta/eta_trapezoid_table_241pt.dat");

  EmissionFunctionArray efa(paraRdr, &chosen_particles, &pT_tab, &phi_tab, &y_tab, &eta_tab,
                            particle_data, Nparticle, surf, FO_length, df_data);

  const int npart = efa.number_of_chosen_particles;
  const long nbins = (long)npart * efa.pT_tab_length * efa.phi_tab_length * efa.y_tab_length;
  double seconds = 0.0;

#ifdef IS3D_REF_DECAYS
  if (what == "decays")
  {
    FILE *f = fopen("input/spectra_in.bin", "rb");
    if (!f) { fprintf(stderr, "ref_driver: input/spectra_in.bin missing\n"); return 2; }
    if ((long)fread(efa.dN_pTdpTdphidy, sizeof(double), nbins, f) != nbins) { fprintf(stderr, "ref_driver: short spectra_in.bin\n"); return 2; }
    fclose(f);
    auto t0 = chrono::steady_clock::now();
    efa.do_resonance_decays(particle_data);
    seconds = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
    efa.write_dN_pTdpTdphidy_with_resonance_decays_toFile();      // the two files calculate_spectra writes next (emissionfunction.cpp:1695-1696)
    efa.write_dN_dpTdphidy_with_resonance_decays_toFile();
  }
  else
#endif
  if (what == "full")
  {
    std::vector<std::vector<Sampled_Particle> > dummy;
    auto t0 = chrono::steady_clock::now();
    efa.calculate_spectra(dummy);
    seconds = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
  }
  else
  {
    // per-species arrays, emissionfunction.cpp:1293-1307
    double *Mass = zeros(npart), *Sign = zeros(npart), *Degeneracy = zeros(npart), *Baryon = zeros(npart);
    for (int ipart = 0; ipart < npart; ipart++)
    {
      particle_info *p = &particle_data[efa.chosen_particles_sampling_table[ipart]];
      Mass[ipart] = p->mass;  Sign[ipart] = p->sign;  Degeneracy[ipart] = p->gspin;  Baryon[ipart] = p->baryon;
    }
    const long N = FO_length;
    double *T = zeros(N), *P = zeros(N), *E = zeros(N), *tau = zeros(N), *eta = zeros(N);
    double *ux = zeros(N), *uy = zeros(N), *un = zeros(N);
    double *dat = zeros(N), *dax = zeros(N), *day = zeros(N), *dan = zeros(N);
    double *pitt = zeros(N), *pitx = zeros(N), *pity = zeros(N), *pitn = zeros(N), *pinn = zeros(N);
    double *pixx = zeros(N), *pixy = zeros(N), *pixn = zeros(N), *piyy = zeros(N), *piyn = zeros(N);
    double *bulkPi = zeros(N), *muB = zeros(N), *nB = zeros(N), *Vx = zeros(N), *Vy = zeros(N), *Vn = zeros(N);
    double *Wx = zeros(N), *Wy = zeros(N), *Lambda = zeros(N), *aL = zeros(N);
    double *c0 = zeros(N), *c1 = zeros(N), *c2 = zeros(N), *c3 = zeros(N), *c4 = zeros(N);
    for (long i = 0; i < N; i++)
    {
      const FO_surf &s = surf[i];
      T[i] = s.T; P[i] = s.P; E[i] = s.E; tau[i] = s.tau; eta[i] = s.eta;
      ux[i] = s.ux; uy[i] = s.uy; un[i] = s.un;
      dat[i] = s.dat; dax[i] = s.dax; day[i] = s.day; dan[i] = s.dan;
      pitt[i] = s.pitt; pitx[i] = s.pitx; pity[i] = s.pity; pitn[i] = s.pitn; pinn[i] = s.pinn;
      pixx[i] = s.pixx; pixy[i] = s.pixy; pixn[i] = s.pixn; piyy[i] = s.piyy; piyn[i] = s.piyn;
      bulkPi[i] = s.bulkPi; muB[i] = s.muB; nB[i] = s.nB; Vx[i] = s.Vx; Vy[i] = s.Vy; Vn[i] = s.Vn;
      Wx[i] = s.Wx; Wy[i] = s.Wy; Lambda[i] = s.Lambda; aL[i] = s.aL;
    }
    Gauss_Laguerre *gla = new Gauss_Laguerre;
    gla->load_roots_and_weights("tables/gla_roots_weights_32_points.txt");

    if (what == "yield")
    {
      // sampler mean yield: the densities come from compute_particle_densities (called above), emissionfunction.cpp:1296-1306
      double *Eq = zeros(npart), *Bk = zeros(npart), *Df = zeros(npart);
      for (int ipart = 0; ipart < npart; ipart++)
      {
        particle_info *p = &particle_data[efa.chosen_particles_sampling_table[ipart]];
        Eq[ipart] = p->equilibrium_density;  Bk[ipart] = p->bulk_density;  Df[ipart] = p->diff_density;
      }
      auto t0 = chrono::steady_clock::now();
      double Ntot = efa.calculate_total_yield(Eq, Bk, Df, T, P, E, tau, ux, uy, un, dat, dax, day, dan, pixx, pixy, pixn, piyy, piyn,
                                              bulkPi, muB, nB, Vx, Vy, Vn, df_data, gla);
      seconds = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
      printf("\nREF_YIELD %.17g\n", Ntot);
      printf("REF_DENSITIES");
      for (int ipart = 0; ipart < npart; ipart++) printf(" %.17g %.17g", Eq[ipart], Bk[ipart]);
      printf("\n");
    }
    else if (what == "vah")
    {
      FILE *f = fopen("input/vah_coefficients.bin", "rb");
      if (!f) { fprintf(stderr, "ref_driver: input/vah_coefficients.bin missing\n"); return 2; }
      double *cs[5] = {c0, c1, c2, c3, c4};
      for (int k = 0; k < 5; k++)
        if ((long)fread(cs[k], sizeof(double), N, f) != N) { fprintf(stderr, "ref_driver: short vah_coefficients.bin\n"); return 2; }
      fclose(f);
      auto t0 = chrono::steady_clock::now();
      efa.calculate_dN_pTdpTdphidy_VAH_PL(Mass, Sign, Degeneracy, tau, eta, ux, uy, un, dat, dax, day, dan, T,
                                          pitt, pitx, pity, pitn, pixx, pixy, pixn, piyy, piyn, pinn, bulkPi,
                                          Wx, Wy, Lambda, aL, c0, c1, c2, c3, c4);
      seconds = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
    }
    else if (df_mode == 1 || df_mode == 2)
    {
      auto t0 = chrono::steady_clock::now();
      efa.calculate_dN_pTdpTdphidy(Mass, Sign, Degeneracy, Baryon, T, P, E, tau, eta, ux, uy, un, dat, dax, day, dan,
                                   pixx, pixy, pixn, piyy, piyn, bulkPi, muB, nB, Vx, Vy, Vn, df_data);
      seconds = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
    }
    else
    {
      auto t0 = chrono::steady_clock::now();
      efa.calculate_dN_ptdptdphidy_feqmod(Mass, Sign, Degeneracy, Baryon, T, P, E, tau, eta, ux, uy, un, dat, dax, day, dan,
                                          pixx, pixy, pixn, piyy, piyn, bulkPi, muB, nB, Vx, Vy, Vn, gla, df_data);
      seconds = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
    }
  }

  FILE *fo = fopen(out.c_str(), "wb");
  if (!fo) { fprintf(stderr, "ref_driver: cannot open %s\n", out.c_str()); return 2; }
  fwrite(efa.dN_pTdpTdphidy, sizeof(double), nbins, fo);
  fclose(fo);

  // species order actually used (integer bookkeeping must match bit-for-bit)
  printf("\nREF_JSON {\"n_cells\": %ld, \"n_species\": %d, \"n_pT\": %d, \"n_phi\": %d, \"n_y_tab\": %d, \"n_eta_tab\": %d, "
         "\"seconds\": %.9g, \"mcid\": [", FO_length, npart, efa.pT_tab_length, efa.phi_tab_length, efa.y_tab_length,
         efa.eta_tab_length, seconds);
  for (int ipart = 0; ipart < npart; ipart++)
    printf("%s%ld", ipart ? ", " : "", particle_data[efa.chosen_particles_sampling_table[ipart]].mc_id);
  printf("]}\n");
  return 0;
}
