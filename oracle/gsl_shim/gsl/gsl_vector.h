#ifndef NDNET_ORACLE_GSL_VECTOR_H
#define NDNET_ORACLE_GSL_VECTOR_H
#include <gsl/gsl_matrix.h>
#endif
