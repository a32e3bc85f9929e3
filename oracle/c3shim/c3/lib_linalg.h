/* C3 stand-in: the three CBLAS entry points the reference path calls
 * (cblas_ddot bellman.c:95 / valuefunc.c:528-578, cblas_dgemv
 * valuefunc.c:424-575, cblas_daxpy nodeutil.c old path).  Implemented as
 * plain sequential loops in c3shim.c -- the summation order of a real BLAS
 * is implementation-defined, see DESIGN.md "tolerances". */
#ifndef C3SHIM_LINALG_H
#define C3SHIM_LINALG_H
#include <stddef.h>
#include "array.h"
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112 };
double cblas_ddot(int n, const double *x, int incx, const double *y, int incy);
void cblas_daxpy(int n, double a, const double *x, int incx, double *y, int incy);
void cblas_dgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n,
                 double alpha, const double *a, int lda, const double *x, int incx,
                 double beta, double *y, int incy);
#endif
