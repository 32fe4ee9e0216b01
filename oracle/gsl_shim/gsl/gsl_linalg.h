/* oracle/gsl_shim — TEST INFRASTRUCTURE ONLY; see gsl_matrix.h for provenance. */
#ifndef NDNET_ORACLE_GSL_LINALG_H
#define NDNET_ORACLE_GSL_LINALG_H
#include <gsl/gsl_matrix.h>
#include <gsl/gsl_vector.h>
#include <gsl/gsl_permutation.h>
#include <gsl/gsl_blas.h>
#ifdef __cplusplus
extern "C" {
#endif
int gsl_linalg_LU_decomp(gsl_matrix *A, gsl_permutation *p, int *signum);
double gsl_linalg_LU_det(gsl_matrix *LU, int signum);
int gsl_linalg_LU_sgndet(gsl_matrix *LU, int signum);
int gsl_linalg_LU_invert(const gsl_matrix *LU, const gsl_permutation *p, gsl_matrix *inverse);
#ifdef __cplusplus
}
#endif
#endif
