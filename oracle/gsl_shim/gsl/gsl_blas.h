/* oracle/gsl_shim — TEST INFRASTRUCTURE ONLY; see gsl_matrix.h for provenance. */
#ifndef NDNET_ORACLE_GSL_BLAS_H
#define NDNET_ORACLE_GSL_BLAS_H
#include <gsl/gsl_matrix.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE_t;
typedef enum { CblasUpper = 121, CblasLower = 122 } CBLAS_UPLO_t;
typedef enum { CblasNonUnit = 131, CblasUnit = 132 } CBLAS_DIAG_t;
int gsl_blas_dgemm(CBLAS_TRANSPOSE_t TransA, CBLAS_TRANSPOSE_t TransB, double alpha,
                   const gsl_matrix *A, const gsl_matrix *B, double beta, gsl_matrix *C);
int gsl_blas_ddot(const gsl_vector *X, const gsl_vector *Y, double *result);
#ifdef __cplusplus
}
#endif
#endif
