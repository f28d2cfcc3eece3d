"""Developer smoke: runs every stage of the CUDA path against the CPU oracle and prints diagnostics
(does not stop at the first mismatch).  Usage: python tools/dev_check.py [n_frames]"""
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, ".")
from oracle import orc  # noqa: E402
from vil_fusion_b200 import cabi, synth  # noqa: E402


def section(name):
    print(f"\n=== {name} ===", flush=True)


def main():
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    seq = synth.Sequence("hdl64", nf, seed=0)
    ocfg = orc.config()
    g = cabi.Odometry(cabi.default_config())
    xyzi, ring = seq[0]

    section("stage 1 extract")
    try:
        oe, oes, os_, oss = orc.extract(ocfg, xyzi)
        ne, ns = g.feature_extract(xyzi)
        ge, ges = g.features(0)
        gs, gss = g.features(1)
        print("oracle", oe.shape[0], os_.shape[0], "gpu", ne, ns)
        print("edge src equal:", np.array_equal(oes, ges), " surf src equal:", np.array_equal(oss, gss))
        print("edge pts equal:", np.array_equal(oe, ge), " surf pts equal:", np.array_equal(os_, gs))
        if not np.array_equal(oes, ges):
            k = min(len(oes), len(ges))
            bad = np.nonzero(oes[:k] != ges[:k])[0]
            print(" first edge mismatch at", bad[:5], oes[bad[:5]], ges[bad[:5]])
        if not np.array_equal(oss, gss):
            k = min(len(oss), len(gss))
            bad = np.nonzero(oss[:k] != gss[:k])[0]
            print(" first surf mismatch at", bad[:5], oss[bad[:5]], gss[bad[:5]], "n mismatches", len(bad))
    except Exception:
        traceback.print_exc()

    section("voxel grid")
    try:
        for leaf, pts in ((0.4, oe), (0.8, os_), (0.2, os_)):
            o, ok = orc.voxel_grid(pts, leaf)
            v, guard = g.voxel_downsample(pts, leaf)
            print(f"leaf {leaf}: oracle {o.shape[0]} gpu {v.shape[0]} guard {guard} equal {np.array_equal(o, v)}")
            if o.shape == v.shape and not np.array_equal(o, v):
                print("  max abs diff", np.abs(o - v).max(), "rows differing", int((o != v).any(axis=1).sum()))
        c = np.array([5.0, -3.0, 0.5])
        o = orc.crop_box(os_, c - 20, c + 20)
        v = g.crop_box(os_, c - 20, c + 20)
        print("crop box:", o.shape[0], v.shape[0], np.array_equal(o, v))
        o2, _ = orc.voxel_grid(o, 0.8)
        v2 = g.crop_voxel_downsample(os_, c, 20.0, 0.8)
        print("crop+voxel:", o2.shape[0], v2.shape[0], np.array_equal(o2, v2))
    except Exception:
        traceback.print_exc()

    section("knn5")
    try:
        mp, _ = orc.voxel_grid(os_, 0.8)
        rng = np.random.default_rng(0)
        q = mp[rng.integers(0, mp.shape[0], 3000)].copy()
        q[:, :3] += rng.normal(0, 0.2, (3000, 3)).astype(np.float32)
        oi, od = orc.knn(mp, q)
        gi, gd = g.knn5(mp, q)
        inside = od < 1.0
        print("ranks inside gate:", int(inside.sum()), "d2 equal:", np.array_equal(od[inside], gd[inside]),
              "idx equal:", int((oi[inside] == gi[inside]).sum()), "/", int(inside.sum()))
        full = inside[:, 4]
        print("queries with all 5 inside:", int(full.sum()), "exact rows:", int((oi[full] == gi[full]).all(axis=1).sum()))
        # raw (un-downsampled) map: many points per cell
        oi, od = orc.knn(os_, q)
        gi, gd = g.knn5(os_, q)
        inside = od < 1.0
        print("raw map: ranks inside gate:", int(inside.sum()), "d2 equal:", np.array_equal(od[inside], gd[inside]),
              "idx equal:", int((oi[inside] == gi[inside]).sum()))
    except Exception:
        traceback.print_exc()

    section("full sequence (free running)")
    try:
        o = orc.Odometry(ocfg)
        g2 = cabi.Odometry(cabi.default_config())
        t_gpu = 0.0
        for i in range(nf):
            xyzi, ring = seq[i]
            po, _, _ = o.process_scan(xyzi)
            t0 = time.time()
            pg = g2.process_scan(xyzi)
            t_gpu += time.time() - t0
            c = g2.counts()
            dq = np.abs(po[:4] - pg[:4]).max()
            dt = np.abs(po[4:] - pg[4:]).max()
            print(i, "dq %.2e dt %.2e" % (dq, dt), "maps o", o.cloud(0).shape[0], o.cloud(1).shape[0], "g", c["n_map_edge"], c["n_map_surf"],
                  "ds", c["n_ds_edge"], c["n_ds_surf"], "status", c["status"])
            if i in (1, nf - 1):
                print("   oracle solves", o.solves()[:, :6].tolist())
                print("   gpu solves   ", g2.solves()[:, :6].tolist())
        me, ge_ = o.cloud(0), g2.cloud(0)
        ms, gs_ = o.cloud(1), g2.cloud(1)
        print("final maps equal:", me.shape == ge_.shape and np.array_equal(me, ge_), ms.shape == gs_.shape and np.array_equal(ms, gs_))
        if me.shape == ge_.shape:
            print("  edge map max diff", np.abs(me - ge_).max())
        if ms.shape == gs_.shape:
            print("  surf map max diff", np.abs(ms - gs_).max())
        print("gpu wall per frame (blocking API, incl. first-call overheads): %.3f ms" % (1e3 * t_gpu / nf), "launches", g2.launch_count())
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()
