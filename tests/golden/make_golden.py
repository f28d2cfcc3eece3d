"""Writes tests/golden/small16.npz: a small seeded 16-beam sequence, the CPU oracle's outputs on it at every
stage boundary, and the outputs of the REFERENCE's own vendored kd-tree (oracle/_ref, compiled from
/root/reference/src/global_fusion/include/Scancontext/nanoflann.hpp) on the same map/query sets.

The reference ships no tests or vectors and cannot be built offline, so apart from the kd-tree rows these
vectors pin the restated algorithm, not the reference binary ("parity unpinned").
Run in the build container:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402
from vil_fusion_b200 import synth  # noqa: E402

N_FRAMES = 6


def main():
    orc.build()
    seq = synth.Sequence("vlp16", N_FRAMES, seed=11)
    cfg = orc.config(n_scan=16, n_rings=16)
    out = {}
    scans = [seq[i] for i in range(N_FRAMES)]
    for i, (x, r) in enumerate(scans):
        out[f"scan{i}"] = x
    x0 = scans[0][0]
    e, es, s, ss = orc.extract(cfg, x0)
    out.update(edge=e, edge_src=es, surf=s, surf_src=ss)
    for leaf in (0.4, 0.8):
        v, _ = orc.voxel_grid(s, leaf)
        out[f"vox_surf_{leaf}"] = v
    ve, _ = orc.voxel_grid(e, 0.4)
    out["vox_edge_0.4"] = ve
    c = np.array([2.0, -1.0, 0.0])
    out["crop_center"] = c
    out["crop_surf"] = orc.crop_box(s, c - 15.0, c + 15.0)
    # kd-tree: oracle restatement and the reference's vendored nanoflann, on the voxel-filtered surf cloud
    mp = out["vox_surf_0.8"]
    rng = np.random.default_rng(5)
    q = mp[rng.integers(0, mp.shape[0], 400)].copy()
    q[:, :3] += rng.normal(0, 0.25, (400, 3)).astype(np.float32)
    out["knn_map"], out["knn_q"] = mp, q
    oi, od = orc.knn(mp, q)
    out["knn_idx"], out["knn_d2"] = oi, od
    ri, rd = orc.ref_knn(mp, q)
    out["ref_knn_idx"], out["ref_knn_d2"] = ri, rd
    # full sequence: per-frame poses, final maps, one frame's factors and solve trace
    od_ = orc.Odometry(cfg)
    poses = []
    for i, (x, r) in enumerate(scans):
        if i == 3:  # snapshot before frame 3 for the association / solve vectors
            st = od_.state()
            me, ms = od_.cloud(orc.MAP_EDGE), od_.cloud(orc.MAP_SURF)
        p, ne, ns = od_.process_scan(x)
        poses.append(p)
        if i == 3:
            out["f3_state"], out["f3_map_edge"], out["f3_map_surf"] = st, me, ms
            out["f3_ds_edge"], out["f3_ds_surf"] = od_.cloud(orc.DS_EDGE), od_.cloud(orc.DS_SURF)
            out["f3_solves"] = od_.solves()
    out["poses"] = np.asarray(poses)
    out["final_map_edge"], out["final_map_surf"] = od_.cloud(orc.MAP_EDGE), od_.cloud(orc.MAP_SURF)
    # association at the predicted pose of frame 3 (identity-ish): use the pose the oracle ended frame 2 with
    pose = poses[2]
    f = orc.factors(cfg, pose, out["f3_ds_edge"], out["f3_ds_surf"], out["f3_map_edge"], out["f3_map_surf"])
    for k, v in f.items():
        out[f"fac_{k}"] = v
    out["fac_pose"] = pose
    pab, pnd = orc.pack_factors(out["f3_ds_edge"], out["f3_ds_surf"], f)
    H, g, cost = orc.normal_eq(cfg.huber, pose, pab, pnd)
    out.update(ne_H=H, ne_g=g, ne_cost=np.array([cost]))
    p2, tr, term = orc.solve(cfg.huber, 4, pose, pab, pnd)
    out.update(solve_pose=p2, solve_trace=tr, solve_term=np.array([term]))
    path = os.path.join(ROOT, "tests", "golden", "small16.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", {k: v.shape for k, v in out.items() if k in ("edge", "surf", "poses", "knn_idx")})


if __name__ == "__main__":
    main()
