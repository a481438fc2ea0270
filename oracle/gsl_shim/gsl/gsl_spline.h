#ifndef IS3D_ORACLE_GSL_SPLINE_H
#define IS3D_ORACLE_GSL_SPLINE_H
#include "gsl_interp.h"
#endif
