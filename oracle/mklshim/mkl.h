/* Stand-in for Intel MKL's mkl.h (CBLAS level-1 subset) -- ORACLE BUILD ONLY; see mkl_types.h.
 * libtorch_cpu.so exports only cblas_daxpy of the five routines pcg.cpp uses, so all five are restated
 * as plain loops in oracle/mkl_adapter.cpp. */
#ifndef RCHOL_B200_MKLSHIM_MKL_H
#define RCHOL_B200_MKLSHIM_MKL_H
#include "mkl_types.h"
#include "mkl_spblas.h"
#ifdef __cplusplus
extern "C" {
#endif
void rchol_b200_cblas_dcopy(MKL_INT n, const double *x, MKL_INT incx, double *y, MKL_INT incy);
double rchol_b200_cblas_dnrm2(MKL_INT n, const double *x, MKL_INT incx);
double rchol_b200_cblas_ddot(MKL_INT n, const double *x, MKL_INT incx, const double *y, MKL_INT incy);
void rchol_b200_cblas_dscal(MKL_INT n, double a, double *x, MKL_INT incx);
void rchol_b200_cblas_daxpy(MKL_INT n, double a, const double *x, MKL_INT incx, double *y, MKL_INT incy);
#define cblas_dcopy rchol_b200_cblas_dcopy
#define cblas_dnrm2 rchol_b200_cblas_dnrm2
#define cblas_ddot rchol_b200_cblas_ddot
#define cblas_dscal rchol_b200_cblas_dscal
#define cblas_daxpy rchol_b200_cblas_daxpy
#ifdef __cplusplus
}
#endif
#endif
