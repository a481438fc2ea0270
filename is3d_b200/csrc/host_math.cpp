#include "host_math.h"
#include <vector>

namespace is3d {

void host_spline_init(const double *x, const double *y, int n, double *c)
{
  c[0] = 0.0;
  if (n > 1) c[n - 1] = 0.0;
  const int m = n - 2;                       // interior unknowns c[1..n-2]
  if (m < 1) return;
  std::vector<double> rhs(m), diag(m), off(m), gam(m), alp(m), z(m);
  for (int i = 0; i < m; i++) {
    const double h0 = x[i + 1] - x[i], h1 = x[i + 2] - x[i + 1];
    const double d0 = y[i + 1] - y[i], d1 = y[i + 2] - y[i + 1];
    const double r0 = (h0 != 0.0) ? 1.0 / h0 : 0.0, r1 = (h1 != 0.0) ? 1.0 / h1 : 0.0;
    off[i] = h1;
    diag[i] = 2.0 * (h1 + h0);
    rhs[i] = 3.0 * (d1 * r1 - d0 * r0);
  }
  if (m == 1) { c[1] = rhs[0] / diag[0]; return; }
  alp[0] = diag[0];
  gam[0] = off[0] / alp[0];
  for (int i = 1; i < m - 1; i++) {
    alp[i] = diag[i] - off[i - 1] * gam[i - 1];
    gam[i] = off[i] / alp[i];
  }
  alp[m - 1] = diag[m - 1] - off[m - 2] * gam[m - 2];
  z[0] = rhs[0];
  for (int i = 1; i < m; i++) z[i] = rhs[i] - gam[i - 1] * z[i - 1];
  for (int i = 0; i < m; i++) z[i] = z[i] / alp[i];
  c[m] = z[m - 1];
  for (int i = m - 2; i >= 0; i--) c[i + 1] = z[i] - gam[i] * c[i + 2];
}

bool host_spline_eval(const double *x, const double *y, const double *c, int n, double xv, double *out)
{
  if (!(xv >= x[0] && xv <= x[n - 1])) return false;
  int lo = 0, hi = n - 1;
  while (hi > lo + 1) { const int mid = (hi + lo) / 2; if (x[mid] > xv) hi = mid; else lo = mid; }
  const double dx = x[lo + 1] - x[lo], dy = y[lo + 1] - y[lo];
  const double b = (dy / dx) - dx * (c[lo + 1] + 2.0 * c[lo]) / 3.0;
  const double d = (c[lo + 1] - c[lo]) / (3.0 * dx);
  const double t = xv - x[lo];
  *out = y[lo] + t * (b + t * (c[lo] + t * d));
  return true;
}

}  // namespace is3d
