"""ncu target: the large-map stages of the cell-ordered path once each (1e6-point map, 260 k queries, 20 k new points)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import roofline_sweep as rs
from vil_fusion_b200 import cabi
rng = np.random.default_rng(7)
g = cabi.Odometry(cabi.default_config(max_scan_points=300000, max_map_points=(1 << 20) + 1024, flags=int(sys.argv[1]) if len(sys.argv) > 1 else 0))
mp = rs.make_map(1_000_000, rng)
nq = 260_000
q = mp[rng.integers(0, mp.shape[0], nq)].copy(); q[:, :3] += rng.normal(0, 0.03, (nq, 3)).astype(np.float32)
newp = mp[rng.integers(0, mp.shape[0], 20000)].copy(); newp[:, :3] += rng.normal(0, 0.15, (20000, 3)).astype(np.float32)
print("knn", g.bench_stage(0, mp, q, leaf=0.2, iters=2))
print("update", g.bench_stage(2, mp, newp, leaf=0.2, iters=2))
g.close()
