// Host build of the product's voxel-coordinate arithmetic (ndt-net_b200/csrc/ndt_cell.cuh) for tests/test_ndt_cell.py.
#include "../../ndt-net_b200/csrc/ndt_cell.cuh"

// the reference: core_legacy/src/voxel.c:89-91  (unsigned)floor((p - off) / vs), fp64
static inline unsigned reference_cell(double p, double off, double vs) { return (unsigned)floor((p - off) / vs); }

// returns the number of mismatches of the exact fp64 shortcut; *first = index of the first one
extern "C" long ndt_host_check_exact64(const double *p, const double *off, const double *vs, long n, long *first) {
    long bad = 0;
    for (long i = 0; i < n; i++) {
        const double rv = 1.0 / vs[i];
        if (ndt::cell_exact64(p[i], off[i], vs[i], rv) != reference_cell(p[i], off[i], vs[i])) { if (!bad) *first = i; bad++; }
    }
    return bad;
}

// fp32 prefilter: mismatches among the DECIDED points; *decided = how many the prefilter decided itself
extern "C" long ndt_host_check_prefilter32(const float *p, const float *off, const double *vs, long n, long *first, long *decided) {
    long bad = 0, dec = 0;
    for (long i = 0; i < n; i++) {
        const float rv32 = (float)(1.0 / vs[i]);
        unsigned cell;
        if (!ndt::cell_prefilter32(p[i], off[i], rv32, cell)) continue;
        dec++;
        if (cell != reference_cell((double)p[i], (double)off[i], vs[i])) { if (!bad) *first = i; bad++; }
    }
    *decided = dec;
    return bad;
}
