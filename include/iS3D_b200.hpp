// iS3D_b200.hpp -- header-only C++ mirror of the reference's library entry `class IS3D` (src/cpp/iS3D.h:19-97, iS3D.cpp:26-191)
// on top of the C ABI in is3d_b200.h, for JETSCAPE-style callers that hand the freeze-out surface over in memory:
//
//     IS3D particlization;
//     particlization.read_fo_surf_from_memory(tau, x, y, eta, dsigma_tau, ..., Pi);     // same 21 vectors, same order
//     particlization.run_particlization(0);                                             // 1: read input/surface.dat instead
//
// Same contract as the reference: iS3D_parameters.dat, PDG/, tables/, deltaf_coefficients/ are read relative to the current
// directory and the results/ files are written there.  Differences: operation = 1 (spectra) and operation = 0 (spacetime
// distributions) only -- no sampler, hence no final_particles_ -- and errors throw std::runtime_error instead of exit(-1).
// The reference does not copy pinn on this path (iS3D.cpp:99-134) and neither does the kernel need it (it is reconstructed).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "is3d_b200.h"

class IS3D {
 public:
  IS3D() {}
  ~IS3D() {}

  // the freeze-out surface (member names of the reference class)
  std::vector<double> tau, x, y, eta;                               // contravariant position
  std::vector<double> dsigma_tau, dsigma_x, dsigma_y, dsigma_eta;   // covariant surface normal vector
  std::vector<double> E, T, P;                                      // energy density, temperature, pressure [GeV, fm]
  std::vector<double> ux, uy, un;                                   // contravariant flow velocity
  std::vector<double> pixx, pixy, pixn, piyy, piyn, pinn;           // shear stress
  std::vector<double> Pi;                                           // bulk pressure

  is3d_stats last_stats{};                                          // timings / counters of the last run (not in the reference)

  void read_fo_surf_from_memory(std::vector<double> tau_in, std::vector<double> x_in, std::vector<double> y_in, std::vector<double> eta_in,
                                std::vector<double> dsigma_tau_in, std::vector<double> dsigma_x_in, std::vector<double> dsigma_y_in,
                                std::vector<double> dsigma_eta_in, std::vector<double> E_in, std::vector<double> T_in,
                                std::vector<double> P_in, std::vector<double> ux_in, std::vector<double> uy_in, std::vector<double> un_in,
                                std::vector<double> pixx_in, std::vector<double> pixy_in, std::vector<double> pixn_in,
                                std::vector<double> piyy_in, std::vector<double> piyn_in, std::vector<double> pinn_in,
                                std::vector<double> Pi_in)
  {
    tau = std::move(tau_in); x = std::move(x_in); y = std::move(y_in); eta = std::move(eta_in);
    dsigma_tau = std::move(dsigma_tau_in); dsigma_x = std::move(dsigma_x_in); dsigma_y = std::move(dsigma_y_in); dsigma_eta = std::move(dsigma_eta_in);
    E = std::move(E_in); T = std::move(T_in); P = std::move(P_in);
    ux = std::move(ux_in); uy = std::move(uy_in); un = std::move(un_in);
    pixx = std::move(pixx_in); pixy = std::move(pixy_in); pixn = std::move(pixn_in); piyy = std::move(piyy_in); piyn = std::move(piyn_in);
    pinn = std::move(pinn_in); Pi = std::move(Pi_in);
  }

  // fo_from_file = 1: input/surface.dat; 0: the vectors handed to read_fo_surf_from_memory (iS3D.cpp:90-134)
  void run_particlization(int fo_from_file)
  {
    int rc;
    if (fo_from_file) {
      rc = is3d_b200_run_workdir(".", nullptr, 0, nullptr, 0, &last_stats);
    } else {
      const size_t n = tau.size();
      const std::vector<double> *all[] = {&x, &y, &eta, &dsigma_tau, &dsigma_x, &dsigma_y, &dsigma_eta, &E, &T, &P, &ux, &uy, &un,
                                          &pixx, &pixy, &pixn, &piyy, &piyn, &Pi};
      for (const std::vector<double> *v : all)
        if (v->size() != n) throw std::runtime_error("IS3D: freeze-out vectors differ in length");
      is3d_surface s{};
      s.n_cells = (int64_t)n;
      s.tau = tau.data(); s.eta = eta.data(); s.x = x.data(); s.y = y.data();
      s.dat = dsigma_tau.data(); s.dax = dsigma_x.data(); s.day = dsigma_y.data(); s.dan = dsigma_eta.data();
      s.ux = ux.data(); s.uy = uy.data(); s.un = un.data(); s.T = T.data(); s.P = P.data(); s.E = E.data();
      s.pixx = pixx.data(); s.pixy = pixy.data(); s.pixn = pixn.data(); s.piyy = piyy.data(); s.piyn = piyn.data(); s.bulkPi = Pi.data();
      rc = is3d_b200_run_surface(".", &s, nullptr, 0, nullptr, 0, &last_stats);
    }
    if (rc != IS3D_OK) {
      const char *h = is3d_b200_host_error(), *k = is3d_b200_last_error();
      throw std::runtime_error(std::string("IS3D::run_particlization: ") + is3d_b200_strerror(rc) + ": " + ((h && *h) ? h : k));
    }
  }
};
