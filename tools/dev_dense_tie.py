"""Dev: the first frame of the dense sequence where GPU and oracle poses split by more than rounding — which factor differs?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import orc
from vil_fusion_b200 import cabi, synth

D = synth.DENSE
seq = synth.Sequence(D["sensor"], 300, seed=7, density=D["density"], speed=D["speed"])
cfg = orc.config(edge_leaf=D["edge_leaf"], surf_leaf=D["surf_leaf"])
o = orc.Odometry(cfg)
g = cabi.Odometry(cabi.default_config(edge_leaf=D["edge_leaf"], surf_leaf=D["surf_leaf"], max_scan_points=116000, max_map_points=D["max_map_points"], max_ring_points=1864))
np.set_printoptions(precision=9, linewidth=250)
for i in range(60):
    x = np.ascontiguousarray(seq[i][0])
    me, ms, st = o.cloud(0).copy(), o.cloud(1).copy(), o.state().copy()
    po, _, _ = o.process_scan(x)
    pg = g.process_scan(x)
    if np.abs(po - pg).max() > 1e-12:
        print("frame", i, "pose diff", np.abs(po - pg).max())
        de, ds = o.cloud(2), o.cloud(3)
        # factors of both sides against the PRE-update maps at the oracle's final pose (any pose near the solution shows the same sets)
        g2 = cabi.Odometry(cabi.default_config(edge_leaf=D["edge_leaf"], surf_leaf=D["surf_leaf"], max_scan_points=116000, max_map_points=D["max_map_points"], max_ring_points=1864))
        g2.set_state(st, me, ms)
        ppred = g2.predict()
        f0 = orc.factors(cfg, ppred, de, ds, me, ms)
        pab, pnd = orc.pack_factors(de, ds, f0)
        pose1 = orc.solve(0.1, 4, ppred, pab, pnd)[0]
        print("pose after outer 1", pose1)
        for pose in (ppred, pose1):
            fg = g2.factors(pose, de, ds)
            fo = orc.factors(cfg, pose, de, ds, me, ms)
            for k in ("edge", "surf"):
                vi = fg[k + "_valid"] != fo[k + "_valid"]
                ni = (fg[k + "_nn"] != fo[k + "_nn"]).any(axis=1)
                gate = fo[k + "_d2"][:, 4] < 1.0
                print(k, "valid differs", int(vi.sum()), "nn differs (inside gate)", int((ni & gate).sum()), "of", len(vi))
                for r in np.nonzero(ni & gate)[0][:5]:
                    print("  q", r, "gpu nn", fg[k + "_nn"][r], fg[k + "_d2"][r], "\n       orc nn", fo[k + "_nn"][r], fo[k + "_d2"][r])
        break
