"""Generates tests/golden/ndt_ref_golden.npz by running the REFERENCE's own C sources (compiled by
oracle/Makefile into oracle/_ref/libndnet_ref_det.so: unmodified core_legacy/src/*.c + oracle/gsl_shim,
8 workers run in id order) on the seeded cases of tests/cases.py.  Run in the build container, where
/root/reference exists:   python tests/golden/make_golden.py
The fixture pins oracle/ndt_oracle.c and, through the C ABI, the CUDA path."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ndt-net_b200"))
from oracle import ref_ctypes as rc  # noqa: E402
from tests import cases  # noqa: E402


def main():
    lib = rc.load(rc.ref_lib_path("det"))
    out = {}
    names = []
    for name, pts, labels, ncls, d in cases.small_cases() + cases.medium_cases()[:3]:
        r = rc.downsample(lib, np.asarray(pts, np.float64), d, labels, ncls, introspect=False)
        names.append(name)
        out[name + "/sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(pts).tobytes()).digest(), np.uint8)
        out[name + "/hdr"] = np.array([r.ret, r.lens[0], r.lens[1], r.lens[2], r.num_out, r.num_valid, r.num_kl], np.int64)
        out[name + "/vs"] = np.array([r.voxel_size, *r.offsets], np.float64)
        n = r.num_out if r.ret == 0 else 0
        out[name + "/pts"] = r.points[:n].copy()
        out[name + "/cov"] = r.covs[:n].copy()
        out[name + "/cls"] = r.classes[:n].copy()
        print(f"{name:36s} ret {r.ret:3d} lens {r.lens} vs {r.voxel_size:.6g} out {r.num_out} valid {r.num_valid} kl {r.num_kl}")
    out["names"] = np.array(names)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ndt_ref_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
