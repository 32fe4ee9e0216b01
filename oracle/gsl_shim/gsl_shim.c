/* oracle/gsl_shim/gsl_shim.c — TEST INFRASTRUCTURE ONLY.
 *
 * Restatement of the GSL 2.7.1 + gslcblas routines that
 * /root/reference/core_legacy/src/kullback_leibler.c:48-124 calls, so the
 * reference C core can be compiled where it lies (oracle/Makefile) without the
 * real library, which is absent from this image.  Written from the published
 * algorithms (LAPACK-style level-2 LU / triangular inverse / U*L product that
 * GSL uses below its recursion cross-over of 24; reference-CBLAS loop orders
 * of gslcblas).  PARITY WITH REAL GSL IS UNPINNED — see gsl/gsl_matrix.h.
 *
 * Every routine is written for general sizes / strides (the way the library
 * is), not specialised to 3x3: the specialised closed forms live in
 * oracle/ndt_oracle.c and in the CUDA kernels, and tests/ compares all three.
 */
#include <gsl/gsl_linalg.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

/* ---------------------------------------------------------------- containers */

gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2) {
    gsl_matrix *m = (gsl_matrix *)malloc(sizeof(gsl_matrix));
    m->size1 = n1; m->size2 = n2; m->tda = n2;
    m->data = (double *)malloc(sizeof(double) * (n1 * n2 ? n1 * n2 : 1));
    m->block = NULL; m->owner = 1;
    return m;
}

void gsl_matrix_free(gsl_matrix *m) {
    if (!m) return;
    if (m->owner) free(m->data);
    free(m);
}

gsl_matrix_view gsl_matrix_view_array(double *base, size_t n1, size_t n2) {
    gsl_matrix_view v;
    v.matrix.size1 = n1; v.matrix.size2 = n2; v.matrix.tda = n2;
    v.matrix.data = base; v.matrix.block = NULL; v.matrix.owner = 0;
    return v;
}

gsl_vector_view gsl_vector_view_array(double *base, size_t n) {
    gsl_vector_view v;
    v.vector.size = n; v.vector.stride = 1; v.vector.data = base;
    v.vector.block = NULL; v.vector.owner = 0;
    return v;
}

int gsl_matrix_memcpy(gsl_matrix *dest, const gsl_matrix *src) {
    for (size_t i = 0; i < src->size1; i++)
        for (size_t j = 0; j < src->size2; j++)
            dest->data[i * dest->tda + j] = src->data[i * src->tda + j];
    return 0;
}

int gsl_matrix_sub(gsl_matrix *a, const gsl_matrix *b) {
    for (size_t i = 0; i < a->size1; i++)
        for (size_t j = 0; j < a->size2; j++)
            a->data[i * a->tda + j] -= b->data[i * b->tda + j];
    return 0;
}

int gsl_matrix_transpose_memcpy(gsl_matrix *dest, const gsl_matrix *src) {
    for (size_t i = 0; i < dest->size1; i++)
        for (size_t j = 0; j < dest->size2; j++)
            dest->data[i * dest->tda + j] = src->data[j * src->tda + i];
    return 0;
}

double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j) { return m->data[i * m->tda + j]; }

gsl_permutation *gsl_permutation_alloc(size_t n) {
    gsl_permutation *p = (gsl_permutation *)malloc(sizeof(gsl_permutation));
    p->size = n;
    p->data = (size_t *)malloc(sizeof(size_t) * (n ? n : 1));
    return p;
}

void gsl_permutation_free(gsl_permutation *p) {
    if (!p) return;
    free(p->data);
    free(p);
}

void gsl_permutation_init(gsl_permutation *p) {
    for (size_t i = 0; i < p->size; i++) p->data[i] = i;
}

/* ------------------------------------------------ gslcblas level 1/2/3 loops */

/* idamax: first index of the strictly largest |x|; 0 when nothing exceeds 0. */
static size_t cb_idamax(size_t n, const double *x, size_t inc) {
    double max = 0.0;
    size_t result = 0;
    for (size_t i = 0; i < n; i++) {
        if (fabs(x[i * inc]) > max) { max = fabs(x[i * inc]); result = i; }
    }
    return result;
}

static void cb_dswap(size_t n, double *x, size_t incx, double *y, size_t incy) {
    for (size_t i = 0; i < n; i++) { double t = x[i * incx]; x[i * incx] = y[i * incy]; y[i * incy] = t; }
}

static void cb_dscal(size_t n, double alpha, double *x, size_t inc) {
    for (size_t i = 0; i < n; i++) x[i * inc] *= alpha;
}

static double cb_ddot(size_t n, const double *x, size_t incx, const double *y, size_t incy) {
    double r = 0.0;
    for (size_t i = 0; i < n; i++) r += x[i * incx] * y[i * incy];
    return r;
}

/* dger, row-major: A += alpha * x * y^T */
static void cb_dger(size_t M, size_t N, double alpha, const double *x, size_t incx,
                    const double *y, size_t incy, double *A, size_t lda) {
    for (size_t i = 0; i < M; i++) {
        const double tmp = alpha * x[i * incx];
        for (size_t j = 0; j < N; j++) A[lda * i + j] += y[j * incy] * tmp;
    }
}

/* dtrmv, row-major, NoTrans: x := T*x */
static void cb_dtrmv_notrans(int upper, int nonunit, size_t N, const double *A, size_t lda,
                             double *x, size_t incx) {
    if (upper) {
        for (size_t i = 0; i < N; i++) {
            double temp = 0.0;
            for (size_t j = i + 1; j < N; j++) temp += x[j * incx] * A[lda * i + j];
            if (nonunit) x[i * incx] = temp + x[i * incx] * A[lda * i + i];
            else x[i * incx] += temp;
        }
    } else {
        for (size_t i = N; i > 0 && i--;) {
            double temp = 0.0;
            for (size_t j = 0; j < i; j++) temp += x[j * incx] * A[lda * i + j];
            if (nonunit) x[i * incx] = temp + x[i * incx] * A[lda * i + i];
            else x[i * incx] += temp;
        }
    }
}

/* dgemv, row-major: y := alpha*op(A)*x + beta*y, A is M x N */
static void cb_dgemv(int trans, size_t M, size_t N, double alpha, const double *A, size_t lda,
                     const double *x, size_t incx, double beta, double *y, size_t incy) {
    if (M == 0 || N == 0) return;
    if (alpha == 0.0 && beta == 1.0) return;
    const size_t lenX = trans ? M : N, lenY = trans ? N : M;
    if (beta == 0.0) { for (size_t i = 0; i < lenY; i++) y[i * incy] = 0.0; }
    else if (beta != 1.0) { for (size_t i = 0; i < lenY; i++) y[i * incy] *= beta; }
    if (alpha == 0.0) return;
    if (!trans) {
        for (size_t i = 0; i < lenY; i++) {
            double temp = 0.0;
            for (size_t j = 0; j < lenX; j++) temp += x[j * incx] * A[lda * i + j];
            y[i * incy] += alpha * temp;
        }
    } else {
        for (size_t j = 0; j < lenX; j++) {
            const double temp = alpha * x[j * incx];
            if (temp != 0.0) {
                for (size_t i = 0; i < lenY; i++) y[i * incy] += temp * A[lda * j + i];
            }
        }
    }
}

int gsl_blas_dgemm(CBLAS_TRANSPOSE_t TransA, CBLAS_TRANSPOSE_t TransB, double alpha,
                   const gsl_matrix *A, const gsl_matrix *B, double beta, gsl_matrix *C) {
    /* only the NoTrans/NoTrans case is reachable from kullback_leibler.c:98,107 */
    if (TransA != CblasNoTrans || TransB != CblasNoTrans) {
        fprintf(stderr, "gsl_shim: dgemm transpose variants not restated\n");
        abort();
    }
    const size_t M = C->size1, N = C->size2, K = A->size2;
    if (A->size1 != M || B->size2 != N || B->size1 != K) {
        fprintf(stderr, "gsl_shim: invalid length\n");
        return 19; /* GSL_EBADLEN */
    }
    double *c = C->data; const size_t ldc = C->tda;
    const double *a = A->data; const size_t lda = A->tda;
    const double *b = B->data; const size_t ldb = B->tda;
    if (alpha == 0.0 && beta == 1.0) return 0;
    if (beta == 0.0) {
        for (size_t i = 0; i < M; i++) for (size_t j = 0; j < N; j++) c[ldc * i + j] = 0.0;
    } else if (beta != 1.0) {
        for (size_t i = 0; i < M; i++) for (size_t j = 0; j < N; j++) c[ldc * i + j] *= beta;
    }
    if (alpha == 0.0) return 0;
    for (size_t k = 0; k < K; k++) {
        for (size_t i = 0; i < M; i++) {
            const double temp = alpha * a[lda * i + k];
            if (temp != 0.0) {
                for (size_t j = 0; j < N; j++) c[ldc * i + j] += temp * b[ldb * k + j];
            }
        }
    }
    return 0;
}

int gsl_blas_ddot(const gsl_vector *X, const gsl_vector *Y, double *result) {
    *result = cb_ddot(X->size, X->data, X->stride, Y->data, Y->stride);
    return 0;
}

/* ------------------------------------------------------------- gsl_linalg LU */

/* Level-2 right-looking LU with partial pivoting (the path GSL 2.7 takes for
 * N <= 24): idamax pivot, whole-row swap, reciprocal scaling of the sub-column
 * when |a_jj| >= DBL_MIN (element-wise division otherwise), rank-1 update. */
int gsl_linalg_LU_decomp(gsl_matrix *A, gsl_permutation *p, int *signum) {
    const size_t M = A->size1, N = A->size2, lda = A->tda;
    const size_t minMN = M < N ? M : N;
    double *a = A->data;
    size_t ipiv[64];
    if (minMN > 64 || p->size != minMN) { fprintf(stderr, "gsl_shim: bad LU size\n"); abort(); }

    for (size_t j = 0; j < minMN; ++j) {
        size_t j_pivot = j + cb_idamax(M - j, a + lda * j + j, lda);
        ipiv[j] = j_pivot;
        if (j_pivot != j) cb_dswap(N, a + lda * j, 1, a + lda * j_pivot, 1);
        if (j < M - 1) {
            double Ajj = a[lda * j + j];
            if (fabs(Ajj) >= DBL_MIN) {
                cb_dscal(M - j - 1, 1.0 / Ajj, a + lda * (j + 1) + j, lda);
            } else {
                for (size_t i = 1; i < M - j; ++i) a[lda * (j + i) + j] /= Ajj;
            }
        }
        if (j < minMN - 1) {
            cb_dger(M - j - 1, N - j - 1, -1.0, a + lda * (j + 1) + j, lda, a + lda * j + (j + 1), 1,
                    a + lda * (j + 1) + (j + 1), lda);
        }
    }

    gsl_permutation_init(p);
    *signum = 1;
    for (size_t i = 0; i < minMN; ++i) {
        size_t pivi = ipiv[i];
        if (p->data[pivi] != p->data[i]) {
            size_t tmp = p->data[pivi];
            p->data[pivi] = p->data[i];
            p->data[i] = tmp;
            *signum = -(*signum);
        }
    }
    return 0;
}

double gsl_linalg_LU_det(gsl_matrix *LU, int signum) {
    const size_t n = LU->size1;
    double det = (double)signum;
    for (size_t i = 0; i < n; i++) det *= LU->data[i * LU->tda + i];
    return det;
}

int gsl_linalg_LU_sgndet(gsl_matrix *LU, int signum) {
    const size_t n = LU->size1;
    int s = signum;
    for (size_t i = 0; i < n; i++) {
        double u = LU->data[i * LU->tda + i];
        if (u < 0) s *= -1;
        else if (u == 0) { s = 0; break; }
    }
    return s;
}

/* in-place inverse of a triangular matrix, level-2 (LAPACK dtrti2 shape) */
static void tri_invert_L2(int upper, int nonunit, size_t N, double *T, size_t ld) {
    if (upper) {
        for (size_t i = 0; i < N; ++i) {
            double aii;
            if (nonunit) { T[ld * i + i] = 1.0 / T[ld * i + i]; aii = -T[ld * i + i]; }
            else aii = -1.0;
            if (i > 0) {
                cb_dtrmv_notrans(1, nonunit, i, T, ld, T + i, ld);
                cb_dscal(i, aii, T + i, ld);
            }
        }
    } else {
        for (size_t i = 0; i < N; ++i) {
            const size_t j = N - i - 1;
            double ajj;
            if (nonunit) { T[ld * j + j] = 1.0 / T[ld * j + j]; ajj = -T[ld * j + j]; }
            else ajj = -1.0;
            if (j < N - 1) {
                cb_dtrmv_notrans(0, nonunit, N - j - 1, T + ld * (j + 1) + (j + 1), ld,
                                 T + ld * (j + 1) + j, ld);
                cb_dscal(N - j - 1, ajj, T + ld * (j + 1) + j, ld);
            }
        }
    }
}

/* in-place product U*L (U upper non-unit, L unit lower, packed in one array), level-2 */
static void tri_UL_L2(size_t N, double *A, size_t ld) {
    if (N == 1) return;
    for (size_t i = 0; i < N; ++i) {
        double *Aii = A + ld * i + i;
        const double Uii = *Aii;
        if (i < N - 1) {
            double *lb = A + ld * (i + 1) + i;   /* column i below the diagonal, stride ld */
            double *ur = A + ld * i + (i + 1);   /* row i right of the diagonal, stride 1 */
            *Aii += cb_ddot(N - i - 1, lb, ld, ur, 1);
            if (i > 0) {
                double *U_TR = A + (i + 1);            /* i x (N-i-1) */
                double *L_BL = A + ld * (i + 1);       /* (N-i-1) x i */
                double *ut = A + i;                    /* column i above the diagonal, stride ld */
                double *ll = A + ld * i;               /* row i left of the diagonal, stride 1 */
                cb_dgemv(1, N - i - 1, i, 1.0, L_BL, ld, ur, 1, Uii, ll, 1);
                cb_dgemv(0, i, N - i - 1, 1.0, U_TR, ld, lb, ld, 1.0, ut, ld);
            }
        } else {
            cb_dscal(N - 1, Uii, A + ld * (N - 1), 1);
        }
    }
}

int gsl_linalg_LU_invert(const gsl_matrix *LU, const gsl_permutation *p, gsl_matrix *inverse) {
    const size_t N = LU->size1;
    gsl_matrix_memcpy(inverse, LU);
    double *a = inverse->data; const size_t ld = inverse->tda;
    for (size_t i = 0; i < N; i++) {
        if (a[ld * i + i] == 0.0) {
            /* GSL_ERROR("matrix is singular", GSL_EDOM) with the default (aborting) handler.
             * Unreachable from kullback_leibler.c, which returns at :66-78 first. */
            fprintf(stderr, "gsl_shim: matrix is singular\n");
            abort();
        }
    }
    tri_invert_L2(1, 1, N, a, ld);   /* U^{-1} */
    tri_invert_L2(0, 0, N, a, ld);   /* L^{-1} (unit) */
    tri_UL_L2(N, a, ld);             /* U^{-1} L^{-1} */
    /* apply the inverse permutation to the entries of every row */
    double tmp[64];
    for (size_t i = 0; i < N; i++) {
        for (size_t k = 0; k < N; k++) tmp[p->data[k]] = a[ld * i + k];
        for (size_t k = 0; k < N; k++) a[ld * i + k] = tmp[k];
    }
    return 0;
}
