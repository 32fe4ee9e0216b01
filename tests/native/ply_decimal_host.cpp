// Host build of the product's decimal parser (ndt-net_b200/csrc/ply_decimal.cuh) for tests/test_ply_decimal.py:
// the header is integer arithmetic only, so g++ gives the bits the device gives.
#include "../../ndt-net_b200/csrc/ply_decimal.cuh"

extern "C" int ply_host_parse_float(const char *s, long n, double *out) {
    return ply::parse_float(reinterpret_cast<const unsigned char *>(s), 0, n, out);
}
extern "C" int ply_host_parse_int(const char *s, long n, long long *out) {
    return ply::parse_int(reinterpret_cast<const unsigned char *>(s), 0, n, out);
}
extern "C" int ply_host_parse_many(const char *buf, const long *off, long count, double *out, int *status) {
    for (long i = 0; i < count; i++)
        status[i] = ply::parse_float(reinterpret_cast<const unsigned char *>(buf), off[i], off[i + 1], &out[i]);
    return 0;
}
extern "C" int ply_host_parse_line(const char *s, long n, double *xyz, long long *tag) {
    return ply::parse_line(reinterpret_cast<const unsigned char *>(s), 0, n, xyz, tag);
}
