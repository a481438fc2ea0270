/* Minimal stand-in for <gsl/gsl_sf_bessel.h>.  The reference only names gsl_sf_bessel_Kn inside a
   commented-out block (src/cpp/emissionfunction.cpp:54-75), so nothing has to be provided. */
#ifndef IS3D_ORACLE_GSL_SF_BESSEL_H
#define IS3D_ORACLE_GSL_SF_BESSEL_H
#endif
