/* oracle/gsl_shim — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Minimal stand-in for the GNU Scientific Library headers that
 * core_legacy/include/ndnet_core/kullback_leibler.h:33-34 includes.  GSL is not
 * installed in this image and cannot be fetched (no network); the reference pins
 * it only through `apt install libgsl-dev` on Ubuntu 22.04 (Dockerfile:12), i.e.
 * GSL 2.7.1.  The routines below restate the published GSL 2.7.1 / gslcblas
 * algorithms for exactly the entry points kullback_leibler.c:48-124 calls.
 * PARITY WITH REAL GSL IS UNPINNED: no GSL source, binary or golden vector is
 * available offline to check this restatement against.
 */
#ifndef NDNET_ORACLE_GSL_MATRIX_H
#define NDNET_ORACLE_GSL_MATRIX_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { size_t size; size_t stride; double *data; void *block; int owner; } gsl_vector;
typedef struct { gsl_vector vector; } gsl_vector_view;
typedef struct { size_t size1; size_t size2; size_t tda; double *data; void *block; int owner; } gsl_matrix;
typedef struct { gsl_matrix matrix; } gsl_matrix_view;

gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2);
void gsl_matrix_free(gsl_matrix *m);
gsl_matrix_view gsl_matrix_view_array(double *base, size_t n1, size_t n2);
gsl_vector_view gsl_vector_view_array(double *base, size_t n);
int gsl_matrix_memcpy(gsl_matrix *dest, const gsl_matrix *src);
int gsl_matrix_sub(gsl_matrix *a, const gsl_matrix *b);
int gsl_matrix_transpose_memcpy(gsl_matrix *dest, const gsl_matrix *src);
double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j);

#ifdef __cplusplus
}
#endif
#endif
