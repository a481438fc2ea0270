// main.cpp -- is3d_b200_run: the RuniS3D.cpp equivalent (reference src/cpp/RuniS3D.cpp:3-12) for operation = 1 (spectra) and operation = 0 (spacetime distributions).
// Run it from a directory laid out like an iS3D checkout (iS3D_parameters.dat, input/, PDG/, tables/,
// deltaf_coefficients/, results/); it writes the same results/ files as the reference.
#include "../../include/is3d_b200.h"
#include <cstdio>
#include <cstring>

int main(int argc, char **argv)
{
  const char *dir = (argc > 1) ? argv[1] : ".";
  is3d_stats st;
  std::memset(&st, 0, sizeof(st));
  std::printf("is3d_b200: smooth Cooper-Frye spectra on the GPU (drop-in for iS3D operation = 1 and 0)\n");
  const int rc = is3d_b200_run_workdir(dir, nullptr, 0, nullptr, 0, &st);
  if (rc != IS3D_OK) {
    std::fprintf(stderr, "is3d_b200: failed: %s\n", is3d_b200_strerror(rc));
    return rc;
  }
  std::printf("cells skipped (u.dsigma <= 0): %lld, feqmod breakdown cells: %lld\n", (long long)st.cells_skipped_udsigma,
              (long long)st.cells_feqmod_breakdown);
  std::printf("%lld evaluations, kernel %.3f ms (%.3e evals/s), total %.3f ms incl. copies, %d kernel launches\n",
              (long long)st.evaluations, st.kernel_ms, st.kernel_ms > 0 ? st.evaluations / (st.kernel_ms * 1e-3) : 0.0, st.total_ms,
              st.gpu_launches);
  std::printf("Done calculating particle spectra. Output stored in results folder.\n");
  return 0;
}
