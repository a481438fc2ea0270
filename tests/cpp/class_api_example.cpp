// Drives the header-only IS3D class (include/iS3D_b200.hpp) the way a JETSCAPE module drives the reference's iS3D_lib:
// reads a surface in the mode-1 text format into 21 vectors (GeV / fm units applied like FO_data_reader::read_surf_VH),
// hands them over in memory and runs the particlization in the current directory.
#include <cstdio>
#include <fstream>
#include <sstream>
#include "iS3D_b200.hpp"

int main(int argc, char **argv)
{
  const char *path = argc > 1 ? argv[1] : "input/surface.dat";
  std::ifstream in(path);
  if (!in) { std::fprintf(stderr, "cannot open %s\n", path); return 2; }
  const double hbarC = 0.197327053;
  std::vector<double> c[20];
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    double v[20];
    int k = 0;
    while (k < 20 && (ls >> v[k])) k++;
    if (k < 20) continue;
    for (int j = 0; j < 20; j++) c[j].push_back(j >= 11 ? v[j] * hbarC : v[j]);       // E, T, P, pi, Pi: fm^-n -> GeV
  }
  IS3D is3d;
  //                            tau   x     y     eta   dat   dax   day   dan   E      T      P      ux    uy    un     pixx   pixy   pixn   piyy   piyn   pinn (unused)                Pi
  is3d.read_fo_surf_from_memory(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[11], c[12], c[13], c[8], c[9], c[10], c[14], c[15], c[16], c[17], c[18], std::vector<double>(c[0].size(), 0.0), c[19]);
  try { is3d.run_particlization(0); }
  catch (const std::exception &e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
  std::printf("cells %lld skipped %lld kernel_ms %.3f launches %d\n", (long long)is3d.tau.size(), (long long)is3d.last_stats.cells_skipped_udsigma,
              is3d.last_stats.kernel_ms, is3d.last_stats.gpu_launches);
  return 0;
}
