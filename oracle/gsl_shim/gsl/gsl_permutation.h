/* oracle/gsl_shim — TEST INFRASTRUCTURE ONLY; see gsl_matrix.h for provenance. */
#ifndef NDNET_ORACLE_GSL_PERMUTATION_H
#define NDNET_ORACLE_GSL_PERMUTATION_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { size_t size; size_t *data; } gsl_permutation;
gsl_permutation *gsl_permutation_alloc(size_t n);
void gsl_permutation_free(gsl_permutation *p);
void gsl_permutation_init(gsl_permutation *p);
#ifdef __cplusplus
}
#endif
#endif
