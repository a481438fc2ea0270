// host_math.h -- small host-side numerics of the host layer (natural cubic spline, Gauss-Laguerre thermal sums).
#pragma once
#include <cstdint>

namespace is3d {

// Natural cubic spline second-derivative coefficients c[n] for knots x[n], values y[n] -- the algorithm behind
// gsl_interp_cspline, which the reference uses for every delta-f coefficient (deltafReader.cpp:300-322):
// c[0] = c[n-1] = 0, interior from the symmetric tridiagonal system solved by an L D L^T recurrence.
void host_spline_init(const double *x, const double *y, int n, double *c);
// Evaluate; returns false when xv is outside [x[0], x[n-1]] (GSL's default handler aborts there).
bool host_spline_eval(const double *x, const double *y, const double *c, int n, double xv, double *out);

}  // namespace is3d
