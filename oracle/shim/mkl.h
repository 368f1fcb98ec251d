/* oracle/shim/mkl.h -- TEST INFRASTRUCTURE, not product code.
 * Stand-in for Intel MKL's <mkl.h>: the reference needs only LAPACK spotrf
 * (mcpar.cc:480) and relies on this header for abort() (mcutil.hh:8). */
#ifndef ORACLE_SHIM_MKL_H_
#define ORACLE_SHIM_MKL_H_
#include <stdlib.h>
#include "mkl_vsl.h"
#ifdef __cplusplus
extern "C" {
#endif
/* Cholesky factorisation, column-major.  uplo='U' on a column-major matrix is the
 * row-major LOWER factor in place; the other triangle is left untouched. */
void spotrf(const char *uplo, const int *n, float *a, const int *lda, int *info);
#ifdef __cplusplus
}
#endif
#endif
