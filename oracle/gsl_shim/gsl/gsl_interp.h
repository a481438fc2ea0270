/* Minimal stand-in for <gsl/gsl_interp.h> / <gsl/gsl_spline.h>: natural cubic spline only.
   TEST INFRASTRUCTURE: lets the unmodified reference sources compile in a container without GSL.
   Restates the published GSL algorithm (interpolation/cspline.c, linalg/tridiag.c):
     c[0] = c[n-1] = 0; interior c from the symmetric tridiagonal system
       diag_i = 2 (h_i + h_{i+1}), offdiag_i = h_{i+1}, rhs_i = 3 (dy_{i+1}/h_{i+1} - dy_i/h_i)
     solved by the L D L^T recurrence; evaluation on [x_i, x_{i+1}]:
       b = dy/h - h (c_{i+1} + 2 c_i)/3,  d = (c_{i+1} - c_i)/(3 h),  y = y_i + t (b + t (c_i + t d)).
   Out-of-range x aborts, like GSL's default error handler. */
#ifndef IS3D_ORACLE_GSL_INTERP_H
#define IS3D_ORACLE_GSL_INTERP_H
#include <stdlib.h>
#include <stdio.h>

typedef struct { int kind; } gsl_interp_type;
static const gsl_interp_type is3d_shim_cspline_type = {1};
static const gsl_interp_type * const gsl_interp_cspline = &is3d_shim_cspline_type;

typedef struct { size_t cache; } gsl_interp_accel;

typedef struct {
  size_t size;
  double *x, *y, *c;
} gsl_spline;

static inline gsl_interp_accel *gsl_interp_accel_alloc(void)
{ return (gsl_interp_accel *)calloc(1, sizeof(gsl_interp_accel)); }
static inline void gsl_interp_accel_free(gsl_interp_accel *a) { free(a); }

static inline gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size)
{
  (void)T;
  gsl_spline *s = (gsl_spline *)calloc(1, sizeof(gsl_spline));
  s->size = size;
  s->x = (double *)calloc(size, sizeof(double));
  s->y = (double *)calloc(size, sizeof(double));
  s->c = (double *)calloc(size, sizeof(double));
  return s;
}
static inline void gsl_spline_free(gsl_spline *s)
{ if (s) { free(s->x); free(s->y); free(s->c); free(s); } }

static inline int gsl_spline_init(gsl_spline *s, const double xa[], const double ya[], size_t size)
{
  size_t i;
  for (i = 0; i < size; i++) { s->x[i] = xa[i]; s->y[i] = ya[i]; }
  const size_t max_index = size - 1;
  const size_t sys = max_index - 1;          /* number of interior unknowns */
  s->c[0] = 0.0; s->c[max_index] = 0.0;
  if (size < 3) return 0;
  double *g = (double *)calloc(sys, sizeof(double));
  double *diag = (double *)calloc(sys, sizeof(double));
  double *off = (double *)calloc(sys, sizeof(double));
  for (i = 0; i < sys; i++) {
    const double h_i = xa[i + 1] - xa[i], h_ip1 = xa[i + 2] - xa[i + 1];
    const double yd_i = ya[i + 1] - ya[i], yd_ip1 = ya[i + 2] - ya[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0;
    const double g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    off[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g[i] = 3.0 * (yd_ip1 * g_ip1 - yd_i * g_i);
  }
  if (sys == 1) {
    s->c[1] = g[0] / diag[0];
  } else {
    const size_t N = sys;
    double *gamma = (double *)calloc(N, sizeof(double));
    double *alpha = (double *)calloc(N, sizeof(double));
    double *cc = (double *)calloc(N, sizeof(double));
    double *z = (double *)calloc(N, sizeof(double));
    double *xs = s->c + 1;
    alpha[0] = diag[0];
    gamma[0] = off[0] / alpha[0];
    for (i = 1; i < N - 1; i++) {
      alpha[i] = diag[i] - off[i - 1] * gamma[i - 1];
      gamma[i] = off[i] / alpha[i];
    }
    alpha[N - 1] = diag[N - 1] - off[N - 2] * gamma[N - 2];
    z[0] = g[0];
    for (i = 1; i < N; i++) z[i] = g[i] - gamma[i - 1] * z[i - 1];
    for (i = 0; i < N; i++) cc[i] = z[i] / alpha[i];
    xs[N - 1] = cc[N - 1];
    if (N >= 2) { size_t j; for (i = N - 2, j = 0; j <= N - 2; j++, i--) xs[i] = cc[i] - gamma[i] * xs[i + 1]; }
    free(gamma); free(alpha); free(cc); free(z);
  }
  free(g); free(diag); free(off);
  return 0;
}

static inline double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *a)
{
  (void)a;
  const size_t n = s->size;
  if (!(x >= s->x[0] && x <= s->x[n - 1])) {
    fprintf(stderr, "gsl shim: interpolation error: x = %.17g outside [%.17g, %.17g]\n", x, s->x[0], s->x[n - 1]);
    abort();
  }
  size_t lo = 0, hi = n - 1;                 /* bisection: x[lo] <= x < x[lo+1] */
  while (hi > lo + 1) { size_t mid = (hi + lo) / 2; if (s->x[mid] > x) hi = mid; else lo = mid; }
  const double x_lo = s->x[lo], x_hi = s->x[lo + 1], y_lo = s->y[lo], y_hi = s->y[lo + 1];
  const double dx = x_hi - x_lo, dy = y_hi - y_lo;
  const double c_i = s->c[lo], c_ip1 = s->c[lo + 1];
  const double b_i = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
  const double d_i = (c_ip1 - c_i) / (3.0 * dx);
  const double delx = x - x_lo;
  return y_lo + delx * (b_i + delx * (c_i + delx * d_i));
}
#endif
