// oracle/vah_ref_driver.cpp -- TEST INFRASTRUCTURE: calls the reference's only (Lambda, alpha_L) coefficient reader,
// DeltafReader::load_coefficients (src/cuda/deltafReader.cu:192-277, host C++ inside a .cu file), compiled unmodified with g++.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "main.cuh"
#include "ParameterReader.cuh"
#include "readindata.cuh"
#include "deltafReader.cuh"
int main(int argc, char **argv)
{
  if (argc < 3) { fprintf(stderr, "usage: vah_ref <in.bin: n doubles aL | n doubles Lambda [GeV]> <out.bin>\n"); return 2; }
  FILE *f = fopen(argv[1], "rb"); if (!f) return 2;
  fseek(f, 0, SEEK_END); long n = ftell(f) / 16; fseek(f, 0, SEEK_SET);
  std::vector<double> aL(n), Lam(n);
  if ((long)fread(aL.data(), 8, n, f) != n || (long)fread(Lam.data(), 8, n, f) != n) return 2;
  fclose(f);
  ParameterReader prm;
  prm.setVal("mode", 2); prm.setVal("df_mode", 4); prm.setVal("include_baryon", 0);
  FO_surf *surf = (FO_surf *)calloc(n, sizeof(FO_surf));
  for (long i = 0; i < n; i++) { surf[i].aL = aL[i]; surf[i].Lambda = Lam[i]; surf[i].c0 = surf[i].c1 = surf[i].c2 = surf[i].c3 = surf[i].c4 = -12345.0; }
  DeltafReader rd(&prm, "deltaf_coefficients");
  rd.load_coefficients(surf, n);
  FILE *o = fopen(argv[2], "wb"); if (!o) return 2;
  for (long i = 0; i < n; i++) { double c[5] = {surf[i].c0, surf[i].c1, surf[i].c2, surf[i].c3, surf[i].c4}; fwrite(c, 8, 5, o); }
  fclose(o);
  return 0;
}
