"""Ring-field / range-image feature extractor: class featureExtract of
src/visual_inertial_lidar/feature_tracker/include/featureExtract.hpp, the alternative stage 1 for drivers that supply ring ids.

CPU part: the oracle restatement (oracle/orc_rangeimage.hpp) against an independent numpy restatement of the projection and of the
curvature / occlusion marks, and against the invariants of the selection.  GPU part: the CUDA path (VILF_FLAG_RANGE_IMAGE) through
the C ABI against the oracle, bit for bit (selected input indices and points), then a free-running sequence with this stage 1.
"""
import numpy as np
import pytest

from conftest import pose_err


def numpy_projection(P, x, ring):
    """projectPointCloud + inverProjectCloud (FX:293-370), vectorised: (src, col, range) of the image points in row-major order."""
    x = x.astype(np.float32)
    fin = np.isfinite(x[:, :3]).all(axis=1)
    rng = np.sqrt(x[:, 0] * x[:, 0] + x[:, 1] * x[:, 1] + x[:, 2] * x[:, 2]).astype(np.float32)
    rxy = np.sqrt(x[:, 0] * x[:, 0] + x[:, 1] * x[:, 1]).astype(np.float32)
    row = ring.astype(np.int64)
    # float atan2: computed in double and rounded (numpy's own float32 arctan2 is a few ulp off, glibc's atan2f is not)
    ang = (np.arctan2(x[:, 0].astype(np.float64), x[:, 1].astype(np.float64)).astype(np.float32).astype(np.float64) * 180.0 / np.pi).astype(np.float32)
    res = np.float32(360.0 / np.float32(P.horizon_scan))
    # std::round = half away from zero
    v = (ang.astype(np.float64) - 90.0) / np.float64(res)
    col = (-(np.sign(v) * np.floor(np.abs(v) + 0.5)) + P.horizon_scan // 2).astype(np.int64)
    col = np.where(col >= P.horizon_scan, col - P.horizon_scan, col)
    ok = fin & ~(rxy.astype(np.float64) < P.lidar_min) & ~(rxy.astype(np.float64) > P.lidar_max) & (row < P.n_scan) & (row % P.downsample_rate == 0) & (col >= 0) & (col < P.horizon_scan)
    idx = np.nonzero(ok)[0]
    cell = row[idx] * P.horizon_scan + col[idx]
    order = np.lexsort((idx, cell))          # by cell, then by input index: the first of each cell is its owner
    cell_s, idx_s = cell[order], idx[order]
    first = np.ones(len(cell_s), bool)
    first[1:] = cell_s[1:] != cell_s[:-1]
    own = idx_s[first]
    return own, (cell_s[first] % P.horizon_scan), rng[own]


def off_grid(x, deg=0.137):
    """The synthetic sensors fire exactly on the boundaries of a Horizon_SCAN = n_az image (azimuth = (j + 0.5) steps), where the
    column of a return hangs on the last bit of atan2f; a real sensor has no such alignment.  A small yaw takes the scan off the grid."""
    a = np.deg2rad(deg)
    c, s = np.float32(np.cos(a)), np.float32(np.sin(a))
    y = x.copy()
    y[:, 0] = c * x[:, 0] - s * x[:, 1]
    y[:, 1] = s * x[:, 0] + c * x[:, 1]
    return np.ascontiguousarray(y)


def numpy_curvature_and_marks(col, rng):
    n = len(rng)
    curv = np.zeros(n, np.float32)
    picked = np.zeros(n, np.int32)
    r = rng.astype(np.float32)
    for i in range(5, n - 5):
        d = np.float32(0)
        for k in (-5, -4, -3, -2, -1, 5, 4, 3, 2, 1):
            d = np.float32(d + r[i + k])
        d = np.float32(d - np.float32(r[i] * np.float32(10)))
        curv[i] = np.float32(d * d)
    for i in range(5, n - 6):
        d1, d2 = r[i], r[i + 1]
        if abs(int(col[i + 1]) - int(col[i])) < 10:
            if float(np.float32(d1 - d2)) > 0.3:
                picked[i - 5:i + 1] = 1
            elif float(np.float32(d2 - d1)) > 0.3:
                picked[i + 1:i + 7] = 1
        a, b = abs(np.float32(r[i - 1] - r[i])), abs(np.float32(r[i + 1] - r[i]))
        if float(a) > 0.02 * float(r[i]) and float(b) > 0.02 * float(r[i]):
            picked[i] = 1
    return curv, picked


@pytest.fixture(scope="module")
def scan16(synth):
    seq = synth.Sequence("vlp16", 2, seed=3)
    x, _ = seq[1]
    # vlp16: 16 rings x 600 azimuth steps, ring-major; the ring id of a return = its beam (recovered from the elevation)
    el = np.degrees(np.arctan2(x[:, 2], np.hypot(x[:, 0], x[:, 1])))
    ring = np.clip(np.round((el + 15.0) / 2.0), 0, 15).astype(np.uint16)
    return off_grid(x), ring


def test_oracle_projection_matches_numpy(orc, scan16, synth):
    x, ring = scan16
    for P in (orc.ri_params(n_scan=16, horizon_scan=600), orc.ri_params(n_scan=16, horizon_scan=1800, downsample_rate=2, lidar_max=40.0)):
        e, es, s, ss, d = orc.ri_extract(P, x, ring, debug=True)
        own, col, rng = numpy_projection(P, x, ring)
        assert np.array_equal(d["src"], own) and np.array_equal(d["col"], col) and np.array_equal(d["range"], rng)
        curv, picked = numpy_curvature_and_marks(d["col"], d["range"])
        assert np.array_equal(d["curvature"], curv) and np.array_equal(d["picked"], picked)
    seq = synth.Sequence("beams128", 1, seed=2)
    x, ring = seq[0]
    x = off_grid(x)
    P = orc.ri_params(n_scan=128, horizon_scan=2048)
    d = orc.ri_extract(P, x, ring, debug=True)[4]
    own, col, rng = numpy_projection(P, x, ring)
    assert np.array_equal(d["src"], own) and np.array_equal(d["col"], col) and np.array_equal(d["range"], rng)


def test_oracle_selection_invariants(orc, synth):
    seq = synth.Sequence("beams128", 1, seed=5)
    x, ring = seq[0]
    x = off_grid(x)
    P = orc.ri_params(n_scan=128, horizon_scan=2048)
    e, es, s, ss, d = orc.ri_extract(P, x, ring, debug=True)
    pos_of = {int(v): i for i, v in enumerate(d["src"])}
    epos = np.array([pos_of[int(v)] for v in es])
    spos = np.array([pos_of[int(v)] for v in ss])
    assert len(e) > 500 and len(s) > 100000
    assert np.array_equal(e, x[es]) and np.array_equal(s, x[ss])
    assert (d["curvature"][epos] > P.edge_threshold).all()          # FX:142
    assert len(set(epos.tolist()) & set(spos.tolist())) == 0          # an edge is never a surf point
    assert (np.diff(spos) > 0).all()                                   # surf points come out in image order (FX:207-211, :222)
    se = d["ring_start_end"]
    for r in range(P.n_scan):                                          # <= 20 edges per sector (FX:147)
        st, en = int(se[r, 0]), int(se[r, 1])
        for j in range(6):
            sp = (st * (6 - j) + en * j) // 6
            ep = (st * (5 - j) + en * (j + 1)) // 6 - 1
            if sp >= ep:
                continue
            k = int(np.count_nonzero((epos >= sp) & (epos <= ep)))
            assert k <= 20
            assert int(np.count_nonzero((spos >= sp) & (spos <= ep))) == ep - sp + 1 - k  # everything else of [sp, ep] is surf
    # the first return of an image cell wins: shuffling the input changes which duplicates survive, not the occupied cells
    perm = np.random.default_rng(1).permutation(len(x))
    d2 = orc.ri_extract(P, x[perm], ring[perm], debug=True)[4]
    assert np.array_equal(d2["col"], d["col"]) and np.array_equal(d2["ring_start_end"], se)


def test_oracle_matches_committed_goldens(orc, golden2):
    for i in (0, 3):
        P = orc.ri_params(n_scan=16, horizon_scan=600)
        e, es, s, ss, d = orc.ri_extract(P, golden2[f"ri{i}_scan"], golden2[f"ri{i}_ring"], debug=True)
        assert np.array_equal(es, golden2[f"ri{i}_edge_src"]) and np.array_equal(ss, golden2[f"ri{i}_surf_src"])
        assert np.array_equal(d["col"], golden2[f"ri{i}_col"]) and np.array_equal(d["curvature"], golden2[f"ri{i}_curvature"])
        assert np.array_equal(d["picked"], golden2[f"ri{i}_picked"])


def test_oracle_degenerate_inputs(orc):
    P = orc.ri_params(n_scan=16, horizon_scan=600)
    for x in (np.zeros((0, 4), np.float32), np.full((40, 4), np.nan, np.float32), np.array([[5, 0, 0, 1]] * 9, np.float32)):
        e, es, s, ss = orc.ri_extract(P, x, np.zeros(len(x), np.uint16))
        assert len(e) == 0 and len(s) == 0


# ---------------------------------------------------------------------------------------------------------
# CUDA path through the C ABI
# ---------------------------------------------------------------------------------------------------------
def gpu_cfg(cabi, n_rings, horizon, cap, **kw):
    return cabi.default_config(n_scan=0, n_rings=n_rings, horizon_scan=horizon, flags=cabi.FLAG_RANGE_IMAGE, lidar_max=200.0, max_scan_points=cap,
                               max_map_points=1 << 18, **kw)


@pytest.mark.gpu
def test_gpu_range_image_extract_bit_exact(cabi, orc, synth, scan16):
    cases = []
    x, ring = scan16
    cases.append((x, ring, 16, 600, dict()))
    cases.append((x, ring, 16, 1800, dict(downsample_rate=2)))
    seq = synth.Sequence("beams128", 2, seed=5)
    xb, rb = seq[1]
    xb = off_grid(xb)
    cases.append((xb, rb, 128, 2048, dict()))
    perm = np.random.default_rng(4).permutation(len(xb))      # firing order != ring-major: the first return of a cell still wins
    cases.append((np.ascontiguousarray(xb[perm]), np.ascontiguousarray(rb[perm]), 128, 2048, dict()))
    bad = xb.copy()
    bad[::977, 2] = np.nan; bad[5::1201, 0] = np.inf
    cases.append((bad, rb, 128, 2048, dict()))
    cases.append((xb, rb, 128, 1024, dict(ri_edge_threshold=0.5, ri_surf_threshold=0.05)))  # coarser image: many cell collisions
    for x, ring, R, H, kw in cases:
        g = cabi.Odometry(gpu_cfg(cabi, R, H, 270000, **kw))
        P = orc.ri_params(n_scan=R, horizon_scan=H, downsample_rate=kw.get("downsample_rate", 1), edge_threshold=kw.get("ri_edge_threshold", 1.0),
                          surf_threshold=kw.get("ri_surf_threshold", 0.1))
        oe, oes, os_, oss = orc.ri_extract(P, x, ring)
        ne, ns = g.feature_extract(x, ring)
        ge, ges = g.features(0)
        gs, gss = g.features(1)
        assert (ne, ns) == (len(oe), len(os_)), (R, H, kw, ne, ns, len(oe), len(os_))
        assert np.array_equal(ges, oes) and np.array_equal(gss, oss)
        assert np.array_equal(ge.view(np.uint32), oe.view(np.uint32)) and np.array_equal(gs.view(np.uint32), os_.view(np.uint32))
        g.close()
    # degenerate inputs: nothing extracted, no error
    g = cabi.Odometry(gpu_cfg(cabi, 16, 600, 20000))
    assert g.feature_extract(np.array([[5, 0, 0, 1]] * 9, np.float32), np.zeros(9, np.uint16)) == (0, 0)
    g.close()
    with pytest.raises(cabi.VilfError):
        cabi.Odometry(cabi.default_config(n_scan=64, flags=cabi.FLAG_RANGE_IMAGE))  # needs explicit ring ids


@pytest.mark.gpu
def test_gpu_range_image_goldens(cabi, golden2):
    """The CUDA path against the committed golden vectors (no oracle in the loop)."""
    g = cabi.Odometry(gpu_cfg(cabi, 16, 600, 20000))
    for i in (0, 3):
        x, ring = golden2[f"ri{i}_scan"], golden2[f"ri{i}_ring"]
        ne, ns = g.feature_extract(x, ring)
        ge, ges = g.features(0)
        gs, gss = g.features(1)
        assert np.array_equal(ges, golden2[f"ri{i}_edge_src"]) and np.array_equal(gss, golden2[f"ri{i}_surf_src"])
        assert np.array_equal(ge, x[ges]) and np.array_equal(gs, x[gss]) and (ne, ns) == (len(ges), len(gss))
    g.close()


@pytest.mark.gpu
def test_gpu_sequence_with_range_image_stage1(cabi, orc, synth):
    """configs[4] shape (128 beams x 2048, leaf 0.2 / 0.4) with the range-image extractor as stage 1, free-running against the oracle
    (its features fed to the same EstimationMapping restatement)."""
    frames = 30
    seq = synth.Sequence("beams128", frames, seed=9)
    P = orc.ri_params(n_scan=128, horizon_scan=2048)
    o = orc.Odometry(orc.config(n_scan=0, n_rings=128, edge_leaf=0.2, surf_leaf=0.4))
    g = cabi.Odometry(gpu_cfg(cabi, 128, 2048, 270000, edge_leaf=0.2, surf_leaf=0.4))
    for i in range(frames):
        x, ring = seq[i]
        x = off_grid(x)
        e, _, s, _ = orc.ri_extract(P, x, ring)
        if i == 0:
            o.init_map(e, s); po = np.array([0, 0, 0, 1, 0, 0, 0.0])
        else:
            po = o.update(e, s)
        pg = g.process_scan(x, ring)
        er = pose_err(po, pg)
        assert er[0] <= 1e-4 and er[1] <= 1e-3, (i, er)
    for which in (0, 1):
        mo, mg = o.cloud(which), g.cloud(which)
        assert mo.shape == mg.shape and np.abs(mo - mg).max() <= 1e-5
    assert np.array_equal(o.solves()[:, :4], g.solves()[:, :4])
    c = g.counts()
    assert c["status"] == 0 and c["n_edge"] > 500
    g.close()
