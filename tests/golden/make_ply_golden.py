"""Generates tests/golden/carla_like.ply and tests/golden/ply_ref_golden.npz by running the REFERENCE's own reader
(/root/reference/ndnet/datasets/CARLA_Seg.py:96-183) in this container.  `open3d` (absent here, imported at
CARLA_Seg.py:4 but unused by get_data_pcl) and `ndnet.preprocessing.ndt_legacy` (loads /usr/local/lib/libndnet.so at
import) are stubbed.  Run here (needs /root/reference):   python tests/golden/make_ply_golden.py"""
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
N_CLASSES, N_SAMPLES, SEED = 28, 256, 1234


def load_reference_dataset():
    saved = {k: v for k, v in sys.modules.items() if k == "ndnet" or k.startswith("ndnet.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, "/root/reference")
    stub = types.ModuleType("ndnet.preprocessing.ndt_legacy")
    stub.NDT_Sampler = object
    had_o3d = "open3d" in sys.modules
    if not had_o3d:
        sys.modules["open3d"] = types.ModuleType("open3d")
    try:
        import ndnet  # noqa: F401
        sys.modules["ndnet.preprocessing.ndt_legacy"] = stub
        ref = importlib.import_module("ndnet.datasets.CARLA_Seg")
        assert ref.__file__.startswith("/root/reference"), ref.__file__
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k == "ndnet" or k.startswith("ndnet.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        if not had_o3d:
            del sys.modules["open3d"]
    return ref


def carla_like_text(n_points: int, seed: int) -> str:
    """A CARLA semantic-LiDAR style ASCII PLY: 10 header lines, then `x y z cos_angle object_idx object_tag`
    (the layout tools/viz.py:13-42 reads with class_pos=5).  Mixed literal styles on purpose."""
    rng = np.random.default_rng(seed)
    head = ["ply", "format ascii 1.0", f"element vertex {n_points}", "property float32 x", "property float32 y",
            "property float32 z", "property float32 CosAngle", "property uint32 ObjIdx", "property uint32 ObjTag", "end_header"]
    rows = []
    for i in range(n_points):
        x, y, z = (float(v) for v in rng.normal(0, 30, 3))
        tag = int(rng.integers(0, N_CLASSES + 1))
        style = i % 5
        if style == 0:
            xyz = f"{x:.4f} {y:.4f} {z:.4f}"
        elif style == 1:
            xyz = f"{x!r} {y!r} {z!r}"                          # 17 significant digits
        elif style == 2:
            xyz = f"{x:.6e} {y:.3E} {z:+.5f}"
        elif style == 3:
            xyz = f"{np.float32(x).item()!r} {np.float32(y).item()!r} {np.float32(z).item()!r}"
        else:
            xyz = f"  {x:.2f}\t{y:.0f}   {int(z)}"
        rows.append(f"{xyz} {rng.uniform(-1, 1):.4f} {int(rng.integers(0, 500))} {tag}")
    return "\n".join(head + rows) + "\n"


def main():
    ref = load_reference_dataset()
    path = os.path.join(HERE, "carla_like.ply")
    with open(path, "w") as f:
        f.write(carla_like_text(700, 0))
    ds = ref.CARLA_Seg.__new__(ref.CARLA_Seg)                   # __init__ only lists a directory (CARLA_Seg.py:18-32)
    ds.n_classes, ds.n_samples, ds.path = N_CLASSES, N_SAMPLES, HERE
    np.random.seed(SEED)
    points, gt = ds.get_data_pcl(path)
    np.random.seed(SEED)
    indexes = np.random.choice(700, N_SAMPLES, replace=False)    # the draw get_data_pcl made (:141)
    np.savez_compressed(os.path.join(HERE, "ply_ref_golden.npz"), points=points.numpy(), gt=gt.numpy(), indexes=indexes,
                        n_classes=N_CLASSES, n_samples=N_SAMPLES, seed=SEED)
    print("wrote", path, points.shape, gt.shape)


if __name__ == "__main__":
    main()
