/* oracle/ndt_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the reference's NDT hot path
 * (carlostojal/NDT-Net core_legacy), used ONLY as the parity checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product path
 * (ndt-net_b200/csrc/*.cu behind include/ndnet_b200.h) never links, loads or
 * calls it, and has no CPU fallback.
 *
 * It restates what the COMPILED reference does, quirks
 * included, in the canonical deterministic schedule (the 8 workers of
 * normal_distributions.c run one after the other, so each voxel sees its points
 * in ascending index).  Reference lines followed are cited per function as
 * core_legacy/src/<file>:<lines>.
 *
 * Pinning: the reference publishes no numeric golden vectors for this path
 * (SURVEY.md §4); this restatement is pinned against (a) the reference's own
 * known answers that exist (bounding box, neighbour indices, 16-point cube counts)
 * and (b) outputs of the reference's C sources themselves, compiled unmodified by
 * oracle/Makefile into oracle/_ref/ with GSL replaced by oracle/gsl_shim
 * (tests/test_oracle_vs_ref.py, fixtures in tests/golden/).  The GSL arithmetic
 * itself (third-party, GSL 2.7.1, not vendored) is restated from its published
 * algorithm: PARITY WITH REAL GSL IS UNPINNED.
 *
 * Build: gcc -O2 -ffp-contract=off (x86-64 baseline: no FMA, like the reference).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* core_legacy/include/ndnet_core/ndt.h:38-43, normal_distributions.h:39 */
#define DOWNSAMPLE_UPPER_THRESHOLD 0.2
#define MIN_VOXEL_GUESS 0.01
#define MAX_VOXEL_GUESS 30.0
#define MAX_GUESS_ITERATIONS 15
#define NUM_PCL_WORKERS 8
#define NDIR 6

typedef struct {
    /* inputs of the last run */
    unsigned long n_points;
    /* grid of the accepted (or last) guess */
    unsigned int len[3];
    double offset[3];
    double voxel_size;
    double limits[6];           /* max x,y,z then min x,y,z */
    int evaluations;            /* number of estimate passes */
    int ret;
    /* per-point voxel index on the final grid; -1 = never voxelised (A4/A5) */
    long *point_voxel;
    /* per-cell statistics, G = len[0]*len[1]*len[2] */
    unsigned long G;
    unsigned long *num_samples;
    double *mean;               /* G*3 */
    double *cov;                /* G*9 (LU-mangled after the KL stage) */
    double *cov0;               /* G*9 copy taken before the KL stage */
    unsigned long *num_samples0;/* copy before pruning */
    unsigned short *cls;
    unsigned int *hist;         /* G*(num_classes+1) or NULL */
    unsigned short num_classes;
    int has_classes;
    unsigned long num_nds;      /* occupied cells of the accepted pass */
    /* divergence list */
    unsigned long num_kl0;      /* entries right after calculate_kl_divergences */
    unsigned long num_kl;       /* after prune */
    double *kl_div;             /* sorted list (pre-prune order is kept in *_0 copies) */
    long *kl_p, *kl_q;
    double *kl_div0; long *kl_p0, *kl_q0;
    unsigned long num_valid;    /* after prune */
    unsigned long num_valid0;   /* before prune */
    int prune_ret;
    unsigned long prune_walk;   /* idx_to_remove at the end of the walk */
    /* outputs */
    unsigned long num_out;      /* rows to_point_cloud produced (may exceed num_desired, A15) */
    double *out_pts, *out_cov; unsigned short *out_cls; long *out_voxel;
} ndt_oracle_t;

static void oracle_free_arrays(ndt_oracle_t *o) {
    free(o->point_voxel); free(o->num_samples); free(o->mean); free(o->cov); free(o->cov0);
    free(o->num_samples0); free(o->cls); free(o->hist);
    free(o->kl_div); free(o->kl_p); free(o->kl_q); free(o->kl_div0); free(o->kl_p0); free(o->kl_q0);
    free(o->out_pts); free(o->out_cov); free(o->out_cls); free(o->out_voxel);
    o->point_voxel = NULL; o->num_samples = NULL; o->mean = NULL; o->cov = NULL; o->cov0 = NULL;
    o->num_samples0 = NULL; o->cls = NULL; o->hist = NULL;
    o->kl_div = NULL; o->kl_p = NULL; o->kl_q = NULL; o->kl_div0 = NULL; o->kl_p0 = NULL; o->kl_q0 = NULL;
    o->out_pts = NULL; o->out_cov = NULL; o->out_cls = NULL; o->out_voxel = NULL;
}

ndt_oracle_t *ndt_oracle_create(void) { return (ndt_oracle_t *)calloc(1, sizeof(ndt_oracle_t)); }
void ndt_oracle_destroy(ndt_oracle_t *o) { if (o) { oracle_free_arrays(o); free(o); } }

/* ---- core_legacy/src/pointclouds.c:28-66 ------------------------------------------------- */
static double maxf_(double a, double b) { return a > b ? a : b; }
static double minf_(double a, double b) { return a < b ? a : b; }

void ndt_oracle_limits(const double *pc, unsigned long n, double lim[6]) {
    /* max starts at DBL_MIN (smallest positive double), not -DBL_MAX (pointclouds.c:44-46) */
    lim[0] = lim[1] = lim[2] = DBL_MIN;
    lim[3] = lim[4] = lim[5] = DBL_MAX;
    for (unsigned long i = 0; i < n; i++) {
        for (int a = 0; a < 3; a++) {
            lim[a] = maxf_(pc[i * 3 + a], lim[a]);
            lim[3 + a] = minf_(pc[i * 3 + a], lim[3 + a]);
        }
    }
}

/* ---- core_legacy/src/voxel.c:61-81 --------------------------------------------------------- */
void ndt_oracle_grid(const double lim[6], double vs, int len[3], double off[3]) {
    for (int a = 0; a < 3; a++) {
        double dim = lim[a] - lim[3 + a];
        len[a] = (int)ceil(dim / vs);
        off[a] = lim[3 + a];
    }
}

/* ---- core_legacy/src/voxel.c:116-175 (6-neighbourhood, enum order voxel.h:35-43) ------------ */
/* returns -4 when the neighbour is outside the grid (unsigned wrap of x-1 at x==0 included) */
int ndt_oracle_neighbor(unsigned long index, unsigned int lx, unsigned int ly, unsigned int lz, int dir,
                        unsigned long *out) {
    static const int dx[6] = {1, -1, 0, 0, 0, 0}, dy[6] = {0, 0, 1, -1, 0, 0}, dz[6] = {0, 0, 0, 0, 1, -1};
    if (index >= (unsigned long)lx * ly * lz) return -1;
    unsigned int z = index / (lx * ly), y = (index % (lx * ly)) / lx, x = index % lx;  /* voxel.c:198-200 */
    x += dx[dir]; y += dy[dir]; z += dz[dir];
    if (x >= lx || y >= ly || z >= lz) return -4;
    *out = (unsigned long)z * lx * ly + (unsigned long)y * lx + x;                         /* voxel.c:186 */
    return 0;
}

/* ---- core_legacy/src/normal_distributions.c:28-137,139-285 ---------------------------------- */
/* One estimate pass on a dense grid.  Returns the number of occupied cells.  If point_voxel
 * is non-NULL it receives the cell of every point (or -1). */
static unsigned long estimate_pass(const double *pc, unsigned long n, const unsigned short *classes,
                                   unsigned short num_classes, double vs, const int len[3], const double off[3],
                                   unsigned long *ns, double *mean, double *cov, double *m2, unsigned short *cls,
                                   unsigned int *hist, long *point_voxel) {
    const unsigned long G = (unsigned long)((long)len[0] * len[1] * len[2]);
    const unsigned int nb = (unsigned int)num_classes + 1;
    memset(ns, 0, G * sizeof(*ns));
    memset(mean, 0, G * 3 * sizeof(double));
    memset(cov, 0, G * 9 * sizeof(double));
    memset(m2, 0, G * 3 * sizeof(double));
    if (cls) memset(cls, 0, G * sizeof(*cls));
    if (hist) memset(hist, 0, G * nb * sizeof(*hist));
    if (point_voxel) for (unsigned long i = 0; i < n; i++) point_voxel[i] = -1;

    const unsigned long chunk = n / NUM_PCL_WORKERS;           /* :34-35, tail n%8 never visited */
    for (int w = 0; w < NUM_PCL_WORKERS; w++) {
        for (unsigned long i = w * chunk; i < (w + 1) * chunk; i++) {
            unsigned int v[3];
            for (int a = 0; a < 3; a++) v[a] = (unsigned int)floor((pc[i * 3 + a] - off[a]) / vs);  /* voxel.c:89-91 */
            if (v[0] >= (unsigned int)len[0] || v[1] >= (unsigned int)len[1] || v[2] >= (unsigned int)len[2])
                break;                                          /* worker returns: rest of its chunk dropped (:47-52) */
            const unsigned long idx = (unsigned long)v[2] * len[0] * len[1] + (unsigned long)v[1] * len[0] + v[0];
            if (point_voxel) point_voxel[i] = (long)idx;
            double *mu = mean + idx * 3, *S = cov + idx * 9, *M2 = m2 + idx * 3;
            ns[idx]++;
            const double cnt = (double)ns[idx];
            for (int j = 0; j < 3; j++) {                       /* :78-104 */
                const double x = pc[i * 3 + j];
                const double old = mu[j];
                mu[j] += (x - mu[j]) / cnt;
                M2[j] += (x - old) * (x - mu[j]);
                S[j * 3 + j] = M2[j] / cnt;
                if (isnan(S[j * 3 + j])) S[j * 3 + j] = 0.0;
                for (int k = j + 1; k < 3; k++) {
                    /* mu[j] is already updated, mu[k] (k>j) is not yet: the order-dependent "covariance" */
                    S[j * 3 + k] += (x - mu[j]) * (pc[i * 3 + k] - mu[k]) / cnt;
                    if (isnan(S[j * 3 + k])) S[j * 3 + k] = 0.0;
                    S[k * 3 + j] = S[j * 3 + k];
                }
            }
            if (classes) {                                      /* :107-121 */
                unsigned int *h = hist + idx * nb;
                h[classes[i]]++;
                unsigned int best = 0;
                for (unsigned int c = 0; c <= num_classes; c++)
                    if (h[c] > best) { best = h[c]; cls[idx] = (unsigned short)c; }
            }
        }
    }
    unsigned long occupied = 0;
    for (unsigned long g = 0; g < G; g++) occupied += ns[g] > 0;   /* :265-269 */
    return occupied;
}

/* ---- GSL 2.7.1 restated for 3x3 (see oracle/gsl_shim/gsl_shim.c for the general loops) ------ */

/* gsl_linalg_LU_decomp on a row-major 3x3, in place.  perm[i] = source row of row i. */
void ndt_oracle_lu3(double a[9], int perm[3], int *signum) {
    int ipiv[3];
    for (int j = 0; j < 3; j++) {
        double mx = 0.0; int piv = 0;                               /* idamax: first strictly largest */
        for (int i = 0; i < 3 - j; i++) {
            double v = fabs(a[(j + i) * 3 + j]);
            if (v > mx) { mx = v; piv = i; }
        }
        piv += j;
        ipiv[j] = piv;
        if (piv != j) for (int c = 0; c < 3; c++) { double t = a[j * 3 + c]; a[j * 3 + c] = a[piv * 3 + c]; a[piv * 3 + c] = t; }
        if (j < 2) {
            const double ajj = a[j * 3 + j];
            if (fabs(ajj) >= DBL_MIN) {
                const double r = 1.0 / ajj;
                for (int i = j + 1; i < 3; i++) a[i * 3 + j] *= r;
            } else {
                for (int i = j + 1; i < 3; i++) a[i * 3 + j] /= ajj;
            }
            for (int i = j + 1; i < 3; i++) {                       /* dger, alpha = -1 */
                const double tmp = -1.0 * a[i * 3 + j];
                for (int c = j + 1; c < 3; c++) a[i * 3 + c] += a[j * 3 + c] * tmp;
            }
        }
    }
    perm[0] = 0; perm[1] = 1; perm[2] = 2;
    *signum = 1;
    for (int i = 0; i < 3; i++) {
        const int pv = ipiv[i];
        if (perm[pv] != perm[i]) { int t = perm[pv]; perm[pv] = perm[i]; perm[i] = t; *signum = -*signum; }
    }
}

double ndt_oracle_lu3_det(const double lu[9], int signum) {
    double det = (double)signum;
    det *= lu[0]; det *= lu[4]; det *= lu[8];
    return det;
}

int ndt_oracle_lu3_sgndet(const double lu[9], int signum) {
    int s = signum;
    for (int i = 0; i < 3; i++) {
        const double u = lu[i * 4];
        if (u < 0) s *= -1;
        else if (u == 0) { s = 0; break; }
    }
    return s;
}

/* gsl_linalg_LU_invert: U^{-1} (dtrti2), unit L^{-1}, in-place U^{-1}*L^{-1}, inverse column permutation */
void ndt_oracle_lu3_invert(const double lu[9], const int perm[3], double inv[9]) {
    double t[9];
    memcpy(t, lu, sizeof(t));
    /* upper, non-unit */
    t[0] = 1.0 / t[0];
    t[4] = 1.0 / t[4];
    t[1] = (0.0 + t[1] * t[0]) * (-t[4]);
    t[8] = 1.0 / t[8];
    {
        const double x0 = (0.0 + t[5] * t[1]) + t[2] * t[0];
        const double x1 = 0.0 + t[5] * t[4];
        t[2] = x0 * (-t[8]);
        t[5] = x1 * (-t[8]);
    }
    /* lower, unit */
    t[7] = (t[7] + 0.0) * -1.0;
    {
        const double x1 = t[6] + (0.0 + t[3] * t[7]);
        const double x0 = t[3] + 0.0;
        t[3] = x0 * -1.0;
        t[6] = x1 * -1.0;
    }
    /* U^{-1} * L^{-1} in place */
    {
        /* i = 0 */
        t[0] += (0.0 + t[3] * t[1]) + t[6] * t[2];
        /* i = 1 */
        const double u11 = t[4];
        t[4] += 0.0 + t[7] * t[5];
        if (u11 == 0.0) t[3] = 0.0;                   /* dgemv: beta==0 overwrites, beta==1 is skipped */
        else if (u11 != 1.0) t[3] *= u11;             /* beta * ll */
        { const double tmp = 1.0 * t[5]; if (tmp != 0.0) t[3] += tmp * t[6]; }
        { const double tmp = 0.0 + t[7] * t[2]; t[1] += 1.0 * tmp; }
        /* i = 2 */
        t[6] *= t[8];
        t[7] *= t[8];
    }
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) inv[i * 3 + perm[k]] = t[i * 3 + k];
}

/* ---- core_legacy/src/kullback_leibler.c:28-127 ---------------------------------------------- */
/* Mutates both covariances in place exactly like the reference does through gsl_matrix_view_array. */
int ndt_oracle_kl_pair(unsigned long np_, double *pcov, unsigned long nq, double *qcov, double *div) {
    *div = 0;
    if (np_ <= 1 || nq <= 1) return -1;                             /* :42-45 */
    int pperm[3], qperm[3], psign, qsign;
    ndt_oracle_lu3(pcov, pperm, &psign);                            /* :57 in place */
    ndt_oracle_lu3(qcov, qperm, &qsign);                            /* :58 in place */
    const double p_det = ndt_oracle_lu3_det(pcov, psign);
    const double q_det = ndt_oracle_lu3_det(qcov, qsign);
    if (p_det == 0 || q_det == 0) return -2;                        /* :66 */
    if (ndt_oracle_lu3_sgndet(pcov, psign) == 0 || ndt_oracle_lu3_sgndet(qcov, qsign) == 0) return -2;  /* :71-78 */
    double qinv[9];
    ndt_oracle_lu3_invert(qcov, qperm, qinv);                       /* :92 */
    /* trace of q_inverse * (packed LU of p): cblas_dgemm, beta=0, k-outer, zero A entries skipped (:96-102) */
    double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 3; i++) {
            const double temp = 1.0 * qinv[i * 3 + k];
            if (temp != 0.0) for (int j = 0; j < 3; j++) c[i * 3 + j] += temp * pcov[k * 3 + j];
        }
    double trace = 0;
    for (int i = 0; i < 3; i++) trace += c[i * 4];
    /* Mahalanobis term: dgemm with C aliased to A and beta=0 zeroes A first => exactly 0 (:105-112) */
    const double first_part_result = 0.0;
    *div = 0.5 * (first_part_result + trace - log(q_det / p_det) - 3);   /* :115 */
    return 0;
}

/* ---- driver: core_legacy/src/ndt.c:119-222 --------------------------------------------------- */
int ndt_oracle_run(ndt_oracle_t *o, const double *pc, unsigned long n, const unsigned short *classes,
                   unsigned short num_classes, unsigned long num_desired) {
    oracle_free_arrays(o);
    o->n_points = n; o->num_classes = num_classes; o->has_classes = classes != NULL;
    o->num_out = 0; o->num_kl = o->num_kl0 = 0; o->num_valid = o->num_valid0 = 0; o->prune_ret = 0; o->prune_walk = 0;
    o->G = 0; o->num_nds = 0;
    ndt_oracle_limits(pc, n, o->limits);

    double guess = (double)(MAX_VOXEL_GUESS - MIN_VOXEL_GUESS) / 2.0;   /* ndt.c:136 */
    double min_guess = MIN_VOXEL_GUESS, max_guess = MAX_VOXEL_GUESS;
    unsigned int iter = 0;
    int len[3]; double off[3];
    unsigned long num_nds = 0;
    double *m2 = NULL;
    o->evaluations = 0;
    o->point_voxel = (long *)malloc(sizeof(long) * (n ? n : 1));
    int accepted = 0;
    do {
        ndt_oracle_grid(o->limits, guess, len, off);
        const long Gl = (long)len[0] * len[1] * len[2];
        if (len[0] < 0 || len[1] < 0 || len[2] < 0 || Gl > (1L << 28)) { o->ret = -1; return -1; }  /* stands in for malloc failure ndt.c:151-155 */
        const unsigned long G = (unsigned long)Gl, Ga = G ? G : 1;
        free(o->num_samples); free(o->mean); free(o->cov); free(o->cls); free(o->hist); free(m2);
        o->num_samples = (unsigned long *)malloc(Ga * sizeof(unsigned long));
        o->mean = (double *)malloc(Ga * 3 * sizeof(double));
        o->cov = (double *)malloc(Ga * 9 * sizeof(double));
        m2 = (double *)malloc(Ga * 3 * sizeof(double));
        o->cls = (unsigned short *)malloc(Ga * sizeof(unsigned short));
        o->hist = classes ? (unsigned int *)malloc(Ga * ((size_t)num_classes + 1) * sizeof(unsigned int)) : NULL;
        o->G = G;
        num_nds = estimate_pass(pc, n, classes, num_classes, guess, len, off, o->num_samples, o->mean, o->cov, m2,
                                o->cls, o->hist, o->point_voxel);
        o->evaluations++;
        if (num_nds > num_desired * (1 + DOWNSAMPLE_UPPER_THRESHOLD)) min_guess = guess;   /* ndt.c:169 */
        else if (num_nds < num_desired) max_guess = guess;                                 /* ndt.c:171 */
        else { accepted = 1; break; }
        guess = min_guess + (max_guess - min_guess) / 2.0;                                 /* ndt.c:183 */
        iter++;
    } while (iter < MAX_GUESS_ITERATIONS);
    free(m2);
    for (int a = 0; a < 3; a++) { o->len[a] = (unsigned int)len[a]; o->offset[a] = off[a]; }
    o->voxel_size = guess;
    o->num_nds = num_nds;
    if (!accepted) { o->ret = -3; return -3; }                                             /* ndt.c:191-194 */

    const unsigned long G = o->G;
    const unsigned int lx = o->len[0], ly = o->len[1], lz = o->len[2];
    o->cov0 = (double *)malloc((G ? G : 1) * 9 * sizeof(double));
    memcpy(o->cov0, o->cov, G * 9 * sizeof(double));

    /* calculate_kl_divergences: kullback_leibler.c:129-202.  Literal descending insertion. */
    const unsigned long cap = (G ? G : 1) * NDIR;
    o->kl_div = (double *)malloc(cap * sizeof(double));
    o->kl_p = (long *)malloc(cap * sizeof(long));
    o->kl_q = (long *)malloc(cap * sizeof(long));
    unsigned long K = 0, valid = 0;
    for (unsigned long idx = 0; idx < G; idx++) {           /* z,y,x ascending == linear index ascending */
        if (o->num_samples[idx] == 0) continue;
        valid++;
        for (int d = 0; d < NDIR; d++) {
            unsigned long nb;
            if (ndt_oracle_neighbor(idx, lx, ly, lz, d, &nb) == -4) continue;
            if (o->num_samples[nb] == 0) continue;
            double div = 0;
            if (ndt_oracle_kl_pair(o->num_samples[idx], o->cov + idx * 9, o->num_samples[nb], o->cov + nb * 9, &div) == -2)
                continue;
            unsigned long j = 0;
            while (j < K) { if (o->kl_div[j] < div) break; j++; }       /* :181-186 */
            memmove(o->kl_div + j + 1, o->kl_div + j, (K - j) * sizeof(double));
            memmove(o->kl_p + j + 1, o->kl_p + j, (K - j) * sizeof(long));
            memmove(o->kl_q + j + 1, o->kl_q + j, (K - j) * sizeof(long));
            o->kl_div[j] = div; o->kl_p[j] = (long)idx; o->kl_q[j] = (long)nb;
            K++;
        }
    }
    o->num_kl0 = K; o->num_valid0 = valid;
    o->kl_div0 = (double *)malloc((K ? K : 1) * sizeof(double));
    o->kl_p0 = (long *)malloc((K ? K : 1) * sizeof(long));
    o->kl_q0 = (long *)malloc((K ? K : 1) * sizeof(long));
    memcpy(o->kl_div0, o->kl_div, K * sizeof(double));
    memcpy(o->kl_p0, o->kl_p, K * sizeof(long));
    memcpy(o->kl_q0, o->kl_q, K * sizeof(long));
    o->num_samples0 = (unsigned long *)malloc((G ? G : 1) * sizeof(unsigned long));
    memcpy(o->num_samples0, o->num_samples, G * sizeof(unsigned long));

    /* prune_nds: ndt.c:28-73 */
    o->prune_ret = 0;
    if (num_desired > valid) {
        o->prune_ret = -1;
    } else {
        const unsigned int to_remove = (unsigned int)(valid - num_desired);
        unsigned long walk = 0;
        for (unsigned long i = 0; i < to_remove; walk++) {
            if (walk >= K) { o->prune_ret = -2; break; }
            if (o->num_samples[o->kl_p[walk]] == 0) continue;
            o->num_samples[o->kl_p[walk]] = 0;
            valid--; K--; i++;
        }
        o->prune_walk = walk;
        if (o->prune_ret == 0) {
            /* ndt.c:70-72 shifts K entries left by `walk`; entries past the original end are
             * undefined in the reference (A15) — the oracle keeps only the well-defined ones. */
            const unsigned long avail = o->num_kl0 - walk;
            const unsigned long keep = K < avail ? K : avail;
            memmove(o->kl_div, o->kl_div + walk, keep * sizeof(double));
            memmove(o->kl_p, o->kl_p + walk, keep * sizeof(long));
            memmove(o->kl_q, o->kl_q + walk, keep * sizeof(long));
        }
    }
    o->num_kl = K; o->num_valid = valid;

    /* to_point_cloud: ndt.c:75-117 (ascending linear index) */
    unsigned long rows = 0;
    for (unsigned long g = 0; g < G; g++) rows += o->num_samples[g] > 0;
    o->out_pts = (double *)malloc((rows ? rows : 1) * 3 * sizeof(double));
    o->out_cov = (double *)malloc((rows ? rows : 1) * 9 * sizeof(double));
    o->out_cls = (unsigned short *)calloc(rows ? rows : 1, sizeof(unsigned short));
    o->out_voxel = (long *)malloc((rows ? rows : 1) * sizeof(long));
    unsigned long r = 0;
    for (unsigned long g = 0; g < G; g++) {
        if (o->num_samples[g] == 0) continue;
        memcpy(o->out_pts + r * 3, o->mean + g * 3, 3 * sizeof(double));
        memcpy(o->out_cov + r * 9, o->cov + g * 9, 9 * sizeof(double));
        if (classes) o->out_cls[r] = o->cls[g];
        o->out_voxel[r] = (long)g;
        r++;
    }
    o->num_out = r;
    o->ret = 0;
    return 0;
}

/* ---- accessors for the ctypes wrapper (oracle/ndt_oracle.py) --------------------------------- */
#define GETTER(type, name, expr) type ndt_oracle_get_##name(const ndt_oracle_t *o) { return expr; }
GETTER(int, ret, o->ret)
GETTER(int, evaluations, o->evaluations)
GETTER(double, voxel_size, o->voxel_size)
GETTER(unsigned long, G, o->G)
GETTER(unsigned long, num_nds, o->num_nds)
GETTER(unsigned long, num_kl0, o->num_kl0)
GETTER(unsigned long, num_kl, o->num_kl)
GETTER(unsigned long, num_valid0, o->num_valid0)
GETTER(unsigned long, num_valid, o->num_valid)
GETTER(unsigned long, num_out, o->num_out)
GETTER(int, prune_ret, o->prune_ret)
GETTER(unsigned long, prune_walk, o->prune_walk)
GETTER(const unsigned int *, len, o->len)
GETTER(const double *, offset, o->offset)
GETTER(const double *, limits, o->limits)
GETTER(const long *, point_voxel, o->point_voxel)
GETTER(const unsigned long *, num_samples, o->num_samples)
GETTER(const unsigned long *, num_samples0, o->num_samples0)
GETTER(const double *, mean, o->mean)
GETTER(const double *, cov, o->cov)
GETTER(const double *, cov0, o->cov0)
GETTER(const unsigned short *, cls, o->cls)
GETTER(const double *, kl_div0, o->kl_div0)
GETTER(const long *, kl_p0, o->kl_p0)
GETTER(const long *, kl_q0, o->kl_q0)
GETTER(const double *, kl_div, o->kl_div)
GETTER(const long *, kl_p, o->kl_p)
GETTER(const long *, kl_q, o->kl_q)
GETTER(const double *, out_pts, o->out_pts)
GETTER(const double *, out_cov, o->out_cov)
GETTER(const unsigned short *, out_cls, o->out_cls)
GETTER(const long *, out_voxel, o->out_voxel)
