"""kNN + map-update roofline sweep on large synthetic maps (BASELINE.json configs[2] / configs[4]: ~1e5 .. 4e6 map points,
1e4 .. 2.6e5 queries), device resident, through the C ABI (vilf_bench_stage).  Run on the GPU box; prints one JSON line per case.

Map: a ground plane sampled on a jittered lattice of `spacing` metres inside +-100 m (a voxel-filtered surf map looks like
that) plus vertical walls; queries: map points displaced by a few centimetres (a scan registered with a slightly wrong pose).
Algorithmic bytes: hash build 32 B per map point; query 136 B (16 B query + 5 x 16 B neighbours + 40 B result);
map update 16 B read per input point + 16 B written per voxel.
"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vil_fusion_b200 import cabi

peak = 6544.7
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def make_map(m, rng):
    n_ground = int(m * 0.7)
    side = int(np.sqrt(n_ground))
    sp = 200.0 / side
    gx, gy = np.meshgrid((np.arange(side) + 0.5) * sp - 100.0, (np.arange(side) + 0.5) * sp - 100.0)
    g = np.stack([gx.ravel(), gy.ravel(), np.full(side * side, -1.73)], 1)
    g[:, :2] += rng.uniform(-0.3, 0.3, (g.shape[0], 2)) * sp
    n_wall = m - g.shape[0]
    wy = rng.choice([-15.0, 15.0, -40.0, 40.0], n_wall)
    w = np.stack([rng.uniform(-100, 100, n_wall), wy, rng.uniform(-1.7, 12.0, n_wall)], 1)
    pts = np.concatenate([g, w]).astype(np.float32)
    rng.shuffle(pts)
    return np.concatenate([pts, rng.random((pts.shape[0], 1), dtype=np.float32)], 1)


def main():
    cases = [(100_000, 20_000), (1_000_000, 20_000), (1_000_000, 260_000), (4_000_000, 260_000)]
    if len(sys.argv) > 1:
        cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
    rng = np.random.default_rng(7)
    g = cabi.Odometry(cabi.default_config(max_scan_points=300000, max_map_points=max(c[0] for c in cases) + 1024))
    for m, nq in cases:
        mp = make_map(m, rng)
        q = mp[rng.integers(0, mp.shape[0], nq)].copy()
        q[:, :3] += rng.normal(0, 0.03, (nq, 3)).astype(np.float32)
        spacing = 200.0 / int(np.sqrt(int(m * 0.7)))
        leaf = 0.8 if spacing >= 0.6 else 0.4 if spacing >= 0.3 else 0.2 if spacing >= 0.15 else 0.1  # the leaf such a map would be filtered at
        ms = g.bench_stage(0, mp, q, leaf=leaf, iters=10)
        idx, d2 = g.knn5(mp[: min(m, 200000)], q[:1000])  # sanity: the stage returns real neighbours
        b_build, b_query = 32.0 * mp.shape[0], 136.0 * nq
        out = dict(case="knn", map_points=int(mp.shape[0]), queries=nq, build_ms=ms[0], query_ms=ms[1],
                   queries_per_s=nq / (ms[1] * 1e-3), build_gbs=b_build / (ms[0] * 1e-3) / 1e9, query_gbs=b_query / (ms[1] * 1e-3) / 1e9,
                   build_frac=b_build / (ms[0] * 1e-3) / 1e9 / peak, query_frac=b_query / (ms[1] * 1e-3) / 1e9 / peak, peak_gbs=peak,
                   found5=float((idx[:, 4] >= 0).mean()), grid_leaf=leaf, grid_cell_m=ms[3], grid_shells=int(ms[2]))
        print(json.dumps(out), flush=True)
        for leaf in (0.4, 0.2):
            ms = g.bench_stage(1, mp, leaf=leaf, iters=10)
            b = 16.0 * mp.shape[0] + 16.0 * ms[2]
            print(json.dumps(dict(case="map_update", map_points=int(mp.shape[0]), leaf=leaf, voxels_out=int(ms[2]), ms=ms[0],
                                  points_per_s=mp.shape[0] / (ms[0] * 1e-3), gbs=b / (ms[0] * 1e-3) / 1e9, frac=b / (ms[0] * 1e-3) / 1e9 / peak, peak_gbs=peak)), flush=True)
    g.close()


if __name__ == "__main__":
    main()
