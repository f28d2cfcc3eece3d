import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import orc
from vil_fusion_b200 import cabi, synth
seq = synth.Sequence("hdl64", 2, seed=13)
g = cabi.Odometry(cabi.default_config(max_scan_points=116000, max_map_points=1 << 18))
_, _, surf, _ = orc.extract(orc.config(), seq[0][0])
rng = np.random.default_rng(0)
for n in (1000, 20000, len(surf)):
    mp = surf[:n]
    q = mp[rng.integers(0, n, 3000)].copy(); q[:, :3] += rng.normal(0, 0.2, (3000, 3)).astype(np.float32)
    print("map", n, flush=True)
    ia, da = g.knn5(mp, q)
    io, do = orc.knn(mp, q, 5, canonical=True)
    ins = do < 1.0
    print("  d equal", np.array_equal(da[ins], do[ins]), "i equal", np.array_equal(ia[ins], io[ins]), flush=True)
print("sequence", flush=True)
for i in range(2):
    print(g.process_scan(seq[i][0]), flush=True)
