/* Minimal stand-in for <gsl/gsl_linalg.h>: the handful of dense-matrix calls the reference makes for its
   3x3 inverse (src/cpp/emissionfunction_smooth_kernels.cpp:689-707).  TEST INFRASTRUCTURE only.
   LU_decomp = Gaussian elimination with partial (row) pivoting; LU_invert = solve A X = I column by column. */
#ifndef IS3D_ORACLE_GSL_LINALG_H
#define IS3D_ORACLE_GSL_LINALG_H
#include <stdlib.h>
#include <math.h>

typedef struct { size_t size1, size2, tda; double *data; int owner; } gsl_matrix;
typedef struct { gsl_matrix matrix; } gsl_matrix_view;
typedef struct { size_t size; size_t *data; } gsl_permutation;

static inline gsl_matrix_view gsl_matrix_view_array(double *base, size_t n1, size_t n2)
{ gsl_matrix_view v; v.matrix.size1 = n1; v.matrix.size2 = n2; v.matrix.tda = n2; v.matrix.data = base; v.matrix.owner = 0; return v; }
static inline gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2)
{ gsl_matrix *m = (gsl_matrix *)calloc(1, sizeof(gsl_matrix)); m->size1 = n1; m->size2 = n2; m->tda = n2;
  m->data = (double *)calloc(n1 * n2, sizeof(double)); m->owner = 1; return m; }
static inline void gsl_matrix_free(gsl_matrix *m) { if (m) { if (m->owner) free(m->data); free(m); } }
static inline double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j) { return m->data[i * m->tda + j]; }
static inline void gsl_matrix_set(gsl_matrix *m, size_t i, size_t j, double x) { m->data[i * m->tda + j] = x; }

static inline gsl_permutation *gsl_permutation_calloc(size_t n)
{ gsl_permutation *p = (gsl_permutation *)calloc(1, sizeof(gsl_permutation)); p->size = n;
  p->data = (size_t *)calloc(n, sizeof(size_t)); for (size_t i = 0; i < n; i++) p->data[i] = i; return p; }
static inline void gsl_permutation_free(gsl_permutation *p) { if (p) { free(p->data); free(p); } }

static inline int gsl_linalg_LU_decomp(gsl_matrix *A, gsl_permutation *p, int *signum)
{
  const size_t N = A->size1;
  *signum = 1;
  for (size_t i = 0; i < N; i++) p->data[i] = i;
  for (size_t j = 0; j + 1 < N; j++) {
    double max = fabs(gsl_matrix_get(A, j, j)); size_t i_pivot = j;
    for (size_t i = j + 1; i < N; i++) { double aij = fabs(gsl_matrix_get(A, i, j)); if (aij > max) { max = aij; i_pivot = i; } }
    if (i_pivot != j) {
      for (size_t k = 0; k < N; k++) { double t = gsl_matrix_get(A, j, k); gsl_matrix_set(A, j, k, gsl_matrix_get(A, i_pivot, k)); gsl_matrix_set(A, i_pivot, k, t); }
      size_t t = p->data[j]; p->data[j] = p->data[i_pivot]; p->data[i_pivot] = t;
      *signum = -(*signum);
    }
    const double ajj = gsl_matrix_get(A, j, j);
    if (ajj != 0.0) {
      for (size_t i = j + 1; i < N; i++) {
        const double aij = gsl_matrix_get(A, i, j) / ajj;
        gsl_matrix_set(A, i, j, aij);
        for (size_t k = j + 1; k < N; k++) gsl_matrix_set(A, i, k, gsl_matrix_get(A, i, k) - aij * gsl_matrix_get(A, j, k));
      }
    }
  }
  return 0;
}

static inline int gsl_linalg_LU_invert(const gsl_matrix *LU, const gsl_permutation *p, gsl_matrix *inverse)
{
  const size_t N = LU->size1;
  for (size_t col = 0; col < N; col++) {
    double x[16];
    for (size_t i = 0; i < N; i++) x[i] = (p->data[i] == col) ? 1.0 : 0.0;   /* permuted unit vector */
    for (size_t i = 1; i < N; i++) { double s = x[i]; for (size_t k = 0; k < i; k++) s -= gsl_matrix_get(LU, i, k) * x[k]; x[i] = s; }
    for (size_t ii = N; ii-- > 0;) { double s = x[ii]; for (size_t k = ii + 1; k < N; k++) s -= gsl_matrix_get(LU, ii, k) * x[k]; x[ii] = s / gsl_matrix_get(LU, ii, ii); }
    for (size_t i = 0; i < N; i++) gsl_matrix_set(inverse, i, col, x[i]);
  }
  return 0;
}
#endif
