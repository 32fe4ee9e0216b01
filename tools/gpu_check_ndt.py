"""Developer probe (not a test): detailed GPU-vs-oracle comparison of the NDT path."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ndt-net_b200"))
import torch
from ndnet_b200.engine import NdtEngine
from ndnet_b200.synth import lidar_cloud, modelnet_cloud
from oracle import ndt_oracle
from tests.helpers import same_bits, first_diff

eng = NdtEngine(0)
cases = [(16000, 1000, lidar_cloud, True), (120000, 1000, lidar_cloud, False), (2048, 512, modelnet_cloud, False),
         (120003, 4096, lidar_cloud, True), (120000, 256, lidar_cloud, False), (4099, 100, modelnet_cloud, False)]
for (n, d, gen, lab) in cases:
    B = 3
    pts = np.stack([gen(n, s) if not (lab and gen is lidar_cloud) else gen(n, s) for s in range(B)])
    labels = None
    if lab:
        labels = np.stack([np.random.default_rng(100 + s).integers(0, 29, n).astype(np.uint16) for s in range(B)])
    t = time.time()
    out = eng.downsample(torch.from_numpy(pts).cuda(), d, None if labels is None else torch.from_numpy(labels.astype(np.int16)).cuda(),
                         28 if lab else 0, nan_to_num=False, want_f64=True, want_voxel=True)
    torch.cuda.synchronize(); dt = time.time() - t
    pv = eng.last_point_voxels(B, n).cpu().numpy()
    for b in range(B):
        ref = ndt_oracle.run(pts[b], d, None if labels is None else labels[b], 28 if lab else 0)
        info = out.info[b]
        ok_hdr = (info["status"] == ref.ret and tuple(int(x) for x in info["len"]) == ref.lens and info["voxel_size"] == ref.voxel_size
                  and info["evaluations"] == ref.evaluations)
        ok_pv = np.array_equal(pv[b], ref.point_voxel)
        if ref.ret != 0:
            print(n, d, b, "ret", ref.ret, info["status"], ok_hdr); continue
        f = out.feat64[b].cpu().numpy(); vox = out.voxel[b].cpu().numpy()
        ok_sel = np.array_equal(vox[:ref.num_out], ref.out_voxel[:d])
        ok_mean = same_bits(f[:ref.num_out, :3], ref.out_pts[:d])
        ok_cov = same_bits(f[:ref.num_out, 3:], ref.out_cov[:d])
        ok_lab = labels is None or np.array_equal(out.labels[b].cpu().numpy().astype(np.uint16)[:ref.num_out], ref.out_cls[:d])
        div, p, q = eng.last_kl_list(b, int(info["num_kl"]) + 8)
        same_len = len(div) == ref.num_kl0
        ok_list = same_len and same_bits(div, ref.kl_div0) and np.array_equal(p, ref.kl_p0) and np.array_equal(q, ref.kl_q0)
        nd = 0; maxrel = 0.0
        if same_len and not ok_list:
            # compare as sets keyed by (p,q)
            ka = {(int(a), int(c)): v for a, c, v in zip(p, q, div)}
            kb = {(int(a), int(c)): v for a, c, v in zip(ref.kl_p0, ref.kl_q0, ref.kl_div0)}
            for key, v in kb.items():
                w = ka.get(key)
                if w is None: nd += 1; continue
                if not (v == w or (np.isnan(v) and np.isnan(w))):
                    nd += 1
                    if np.isfinite(v) and np.isfinite(w) and v != 0: maxrel = max(maxrel, abs(v - w) / abs(v))
        print(n, d, b, "hdr", ok_hdr, "pv", ok_pv, "sel", ok_sel, "mean", ok_mean, "cov", ok_cov, "lab", ok_lab, "list", ok_list,
              "K", len(div), ref.num_kl0, "ndiff", nd, "maxrel %.2e" % maxrel, "valid", info["num_valid"], ref.num_valid,
              "evals", info["evaluations"], "batch %.1f ms" % (dt * 1e3))
        if not ok_mean: print("   mean diff", first_diff(f[:ref.num_out, :3], ref.out_pts[:d]))
        if not ok_cov: print("   cov diff", first_diff(f[:ref.num_out, 3:], ref.out_cov[:d]))
