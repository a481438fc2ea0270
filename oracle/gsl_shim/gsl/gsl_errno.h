/* Minimal stand-in for <gsl/gsl_errno.h>: test infrastructure only (see oracle/README.md). */
#ifndef IS3D_ORACLE_GSL_ERRNO_H
#define IS3D_ORACLE_GSL_ERRNO_H
#define GSL_SUCCESS 0
#define GSL_EDOM 1
#endif
