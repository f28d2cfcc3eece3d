"""The cell-ordered map path (k_cellmap.cu: merge update + cell table + warp-per-query 5-NN; the default for maps beyond one
cluster's 2^19 points, forced here with VILF_FLAG_CELL_MAP) against the radix-sorted path (VILF_FLAG_LEGACY_MAP: voxel filter
of the whole map + hashed grid rebuilt every frame), bit for bit, and against the oracle.  The dense configs[2] sequence
(tests/test_gpu_sequence.py) runs the cell-ordered path by default; test_cell_map_long_sequences_vs_oracle below holds it to the
oracle on the standard sequences as well."""
import numpy as np
import pytest

from conftest import check_knn

pytestmark = pytest.mark.gpu

ALL_CLOUDS = (0, 1, 2, 3, 4, 5)


def pair(cabi, **kw):
    return cabi.Odometry(cabi.default_config(flags=cabi.FLAG_CELL_MAP, **kw)), cabi.Odometry(cabi.default_config(flags=cabi.FLAG_LEGACY_MAP, **kw))


@pytest.mark.parametrize("sensor,kw,frames", [
    ("hdl64", dict(max_scan_points=116000, max_map_points=1 << 18), 40),
    ("vlp32", dict(n_scan=32, n_rings=32, max_scan_points=58000, max_map_points=1 << 18), 40),
    ("beams128", dict(n_scan=0, n_rings=128, edge_leaf=0.2, surf_leaf=0.4, max_scan_points=263000, max_map_points=1 << 18, max_ring_points=2048 + 64), 10),
    ("hdl64", dict(edge_leaf=0.1, surf_leaf=0.1, max_scan_points=116000, max_map_points=1 << 20), 30),   # four shells of 0.4 m cells
    ("hdl64", dict(edge_leaf=0.3, surf_leaf=1.1, max_scan_points=116000, max_map_points=1 << 18), 12),   # leaves that are no power-of-two fraction of the gate
])
def test_sequences_are_bit_identical_to_the_legacy_path(cabi, synth, sensor, kw, frames):
    seq = synth.Sequence(sensor, frames, seed=13)
    a, b = pair(cabi, **kw)
    for i in range(frames):
        x, r = seq[i]
        ring = r if sensor == "beams128" else None
        pa, pb = a.process_scan(x, ring), b.process_scan(x, ring)
        assert np.array_equal(pa, pb), i
        if i % 5 == 0 or i == frames - 1:
            for which in ALL_CLOUDS:
                assert np.array_equal(a.cloud(which), b.cloud(which)), (i, which)
            assert np.array_equal(a.solves(), b.solves()), i
    ca, cb = a.counts(), b.counts()
    assert ca == cb and ca["status"] == 0
    a.close(); b.close()


def test_knn5_and_factors_equal_legacy(cabi, orc, synth):
    seq = synth.Sequence("hdl64", 4, seed=5)
    kw = dict(max_scan_points=116000, max_map_points=1 << 18)
    a, b = pair(cabi, **kw)
    rng = np.random.default_rng(2)
    for i in range(4):
        x, _ = seq[i]
        a.process_scan(x); b.process_scan(x)
        pose = a.pose()[0]
        de, ds = a.cloud(cabi.DS_EDGE), a.cloud(cabi.DS_SURF)
        fa, fb = a.factors(pose, de, ds), b.factors(pose, de, ds)
        for k in fa:
            assert np.array_equal(fa[k], fb[k]), (i, k)   # incl. neighbour indices in the reference's map order
    # explicit maps: a filtered map, a raw (many points per voxel) one, a lattice full of exact distance ties, tiny and empty maps
    _, _, surf, _ = orc.extract(orc.config(), seq[0][0])
    filt, _ = orc.voxel_grid(surf, 0.8)
    lattice = np.zeros((20000, 4), np.float32); lattice[:, :3] = np.round(rng.uniform(-8, 8, (20000, 3)) * 4) / 4
    for mp in (filt, surf, lattice, filt[:3], filt[:0]):
        q = (mp if len(mp) else filt)[rng.integers(0, max(len(mp), 1) if len(mp) else len(filt), 3000)].copy()
        q[:, :3] += rng.normal(0, 0.2, (3000, 3)).astype(np.float32)
        if mp is lattice:
            q[:, :3] = np.round(q[:, :3] * 2) / 2
        ia, da = a.knn5(mp, q)
        ib, db = b.knn5(mp, q)
        assert np.array_equal(da, db) and np.array_equal(ia, ib)
        if len(mp) >= 5:
            io, do = orc.knn(mp, q, 5, canonical=True)
            ins = do < np.float32(1.0)
            assert np.array_equal(da[ins], do[ins]) and np.array_equal(ia[ins], io[ins])
    a.close(); b.close()


def test_map_update_edge_cases(cabi, orc):
    """createSubMap through the merge: nothing new, everything new, everything cropped away, many points per voxel on both
    sides, points exactly on voxel faces; maps compared with the oracle's crop box + voxel filter of the concatenation."""
    rng = np.random.default_rng(4)
    g = cabi.Odometry(cabi.default_config(max_scan_points=60000, max_map_points=1 << 17, flags=cabi.FLAG_CELL_MAP))

    def cloud(n, lo, hi, snap=None):
        c = np.zeros((n, 4), np.float32)
        c[:, :3] = rng.uniform(lo, hi, (n, 3))
        if snap:
            c[:, :3] = np.round(c[:, :3] / snap) * snap   # exactly on multiples of `snap` (0.4 and 0.8 m faces included)
        c[:, 3] = rng.random(n)
        return c

    def expect(me, ms, ne, ns, ctr):
        out = []
        for m, n, leaf in ((me, ne, 0.4), (ms, ns, 0.8)):
            w = n.copy()   # pointAssociaToMap (EM:355-363) with the identity rotation: fp64 add of the translation, fp32 store
            w[:, :3] = (n[:, :3].astype(np.float64) + np.asarray(ctr)).astype(np.float32)
            cat = np.concatenate([m, w])
            box = orc.crop_box(cat, [c - 100.0 for c in ctr], [c + 100.0 for c in ctr])
            out.append(orc.voxel_grid(box, leaf)[0])
        return out

    me, ms = cloud(3000, -30, 30), cloud(20000, -30, 30)
    g.map_init(me, ms)
    assert np.array_equal(g.cloud(0), me) and np.array_equal(g.cloud(1), ms)   # localMapInited keeps the raw clouds, in order
    ctr = [0.0, 0.0, 0.0]
    steps = [
        (cloud(0, 0, 1), cloud(0, 0, 1)),                          # nothing new: the first update only filters
        (cloud(2000, -30, 30), cloud(9000, -30, 30)),              # new points into old and new voxels
        (cloud(1500, -20, 20, snap=0.2), cloud(4000, -20, 20, snap=0.2)),  # on voxel faces
        (cloud(4000, 200, 260), cloud(9000, 200, 260)),            # every new point outside the crop box
        (cloud(50000, -2, 2), cloud(50000, -2, 2)),                # hundreds of new points per voxel
    ]
    for k, (ne, ns) in enumerate(steps):
        g.set_pose([0, 0, 0, 1] + ctr)
        g.create_submap(ne, ns)
        me, ms = expect(me, ms, ne, ns, ctr)
        assert np.array_equal(g.cloud(0), me), k
        assert np.array_equal(g.cloud(1), ms), k
    # move the crop box so that most of the map leaves it, then all of it
    for ctr in ([95.0, 0.0, 0.0], [400.0, 0.0, 0.0]):
        g.set_pose([0, 0, 0, 1] + ctr)
        ne, ns = cloud(100, -5, 5), cloud(300, -5, 5)   # sensor frame
        g.create_submap(ne, ns)
        me, ms = expect(me, ms, ne, ns, ctr)
        assert np.array_equal(g.cloud(0), me) and np.array_equal(g.cloud(1), ms)
    assert len(me) <= 100 and len(ms) <= 300
    # 5-NN against the maps as they are now
    c = g.counts()
    assert c["n_map_edge"] == len(me) and c["n_map_surf"] == len(ms) and c["status"] == 0
    g.close()


def test_pcl_guard_is_reported(cabi):
    """A leaf so small that PCL's int32 voxel-index guard would skip the filter is refused loudly by the cell-ordered path
    (the legacy path reproduces PCL's pass-through)."""
    rng = np.random.default_rng(1)
    g = cabi.Odometry(cabi.default_config(edge_leaf=0.01, surf_leaf=0.01, max_scan_points=4096, max_map_points=1 << 14, flags=cabi.FLAG_CELL_MAP))
    c = np.zeros((3000, 4), np.float32); c[:, :3] = rng.uniform(-90, 90, (3000, 3))
    with pytest.raises(cabi.VilfError) as e:
        g.map_init(c[:500], c)
        g.set_pose([0, 0, 0, 1, 0, 0, 0])
        g.create_submap(c[:100], c[:1000])
    assert e.value.code == 4
    g.close()


@pytest.mark.parametrize("sensor,n_scan,cap,frames", [("hdl64", 64, 116000, 300), ("vlp32", 32, 58000, 300)])
def test_cell_map_long_sequences_vs_oracle(cabi, orc, synth, sensor, n_scan, cap, frames):
    """Free-running against the CPU oracle with the cell-ordered maps forced on a standard-size sequence: every frame inside the
    north-star tolerance, final maps within 1e-5 m, identical solver summaries."""
    from conftest import pose_err
    seq = synth.Sequence(sensor, frames, seed=23)
    o = orc.Odometry(orc.config(n_scan=n_scan, n_rings=n_scan))
    g = cabi.Odometry(cabi.default_config(n_scan=n_scan, n_rings=n_scan, max_scan_points=cap, max_map_points=1 << 18, max_ring_points=1864, flags=cabi.FLAG_CELL_MAP))
    for i in range(frames):
        x = np.ascontiguousarray(seq[i][0])
        po, _, _ = o.process_scan(x)
        pg = g.process_scan(x)
        e = pose_err(po, pg)
        assert e[0] <= 1e-4 and e[1] <= 1e-3, (i, e)
    for which in (0, 1):
        mo, mg = o.cloud(which), g.cloud(which)
        assert mo.shape == mg.shape and np.abs(mo - mg).max() <= 1e-5
    assert np.array_equal(o.solves()[:, :4], g.solves()[:, :4])
    assert g.counts()["status"] == 0
    g.close()
