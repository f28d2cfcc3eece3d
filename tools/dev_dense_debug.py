"""Dev: where do the GPU maps of the dense (configs[2]) sequence differ from the oracle's?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import orc
from vil_fusion_b200 import cabi, synth

D = synth.DENSE
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 50
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cap = int(sys.argv[3]) if len(sys.argv) > 3 else D["max_map_points"]
world_frames = int(sys.argv[4]) if len(sys.argv) > 4 else frames
seq = synth.Sequence(D["sensor"], world_frames, seed=7, density=D["density"], speed=D["speed"])
o = orc.Odometry(orc.config(edge_leaf=D["edge_leaf"], surf_leaf=D["surf_leaf"]))
g = cabi.Odometry(cabi.default_config(edge_leaf=D["edge_leaf"], surf_leaf=D["surf_leaf"], max_scan_points=116000, max_map_points=cap, max_ring_points=1864, flags=flags))
for i in range(frames):
    x = np.ascontiguousarray(seq[i][0])
    po, _, _ = o.process_scan(x)
    pg = g.process_scan(x)
    bad = False
    for which in (0, 1, 2, 3):
        mo, mg = o.cloud(which), g.cloud(which)
        if mo.shape != mg.shape:
            print(i, which, "shape", mo.shape, mg.shape); bad = True; continue
        d = np.abs(mo - mg).max(axis=1) if len(mo) else np.zeros(0)
        nz = np.nonzero(d > 0)[0]
        if len(nz):
            bad = True
            print(i, which, "n", len(mo), "rows differing", len(nz), "first", nz[:8], "max", d.max())
            for r in nz[:4]:
                print("   o", mo[r], "g", mg[r])
            so = set(map(bytes, mo)); sg = set(map(bytes, mg))
            print("   as sets: only oracle", len(so - sg), "only gpu", len(sg - so))
    if np.abs(po - pg).max() > 1e-12:
        np.set_printoptions(precision=17, linewidth=250)
        print("oracle solves\n", o.solves()); print("gpu solves\n", g.solves())
    print(i, "pose diff", np.abs(po - pg).max(), "maps", o.cloud(0).shape[0], o.cloud(1).shape[0], flush=True)
    if bad and i > 0:
        break
