"""ScanContext place recognition: SCManager of src/global_fusion/include/Scancontext/Scancontext.h.

CPU part: the oracle restatement (oracle/orc_scancontext.hpp) against an independent numpy restatement of the descriptor, against
the properties the algorithm has by construction (a yaw rotation of the cloud by one sector is a column shift), and its ring-key
search against the REFERENCE's own kd-tree compiled from /root/reference (oracle/_ref/libref_sckeys.so) index for index.
GPU part: every vilf_sc_* entry point against the oracle, bit for bit.
"""
import os

import numpy as np
import pytest


def cloud_of(seq, i, stride=7):
    """A key-frame cloud as global_fusion sees it: sensor-frame points, thinned (the node publishes voxel-filtered features)."""
    return np.ascontiguousarray(seq[i][0][::stride])


def numpy_descriptor(pts, lidar_height=2.0, R=20, S=60, max_radius=80.0):
    """Scancontext.h:42-83 written independently of the oracle (vectorised; float32 where the reference holds floats)."""
    x = pts[:, 0].astype(np.float32); y = pts[:, 1].astype(np.float32)
    z = (pts[:, 2].astype(np.float64) + lidar_height).astype(np.float32)
    rng = np.sqrt(x * x + y * y).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        at = np.arctan(np.where((x >= 0) == (y >= 0), y / x, np.where(x < 0, y / (-x), (-y) / x)).astype(np.float32)).astype(np.float32)
    k = 180.0 / np.pi
    th = np.where((x >= 0) & (y >= 0), k * at.astype(np.float64),
                  np.where((x < 0) & (y >= 0), 180 - k * at.astype(np.float64),
                           np.where((x < 0) & (y < 0), 180 + k * at.astype(np.float64), 360 - k * at.astype(np.float64)))).astype(np.float32)
    keep = ~(rng.astype(np.float64) > max_radius)
    ring = np.clip(np.ceil(rng.astype(np.float64) / max_radius * R).astype(np.int64), 1, R)
    sec = np.clip(np.ceil(th.astype(np.float64) / 360.0 * S).astype(np.int64), 1, S)
    d = np.full((R, S), -1000.0)
    np.maximum.at(d, (ring[keep] - 1, sec[keep] - 1), z[keep].astype(np.float64))
    d[d == -1000.0] = 0
    return d


def rotz(pts, deg):
    a = np.deg2rad(deg)
    c, s = np.cos(a), np.sin(a)
    out = pts.copy()
    out[:, 0] = (c * pts[:, 0] - s * pts[:, 1]).astype(np.float32)
    out[:, 1] = (s * pts[:, 0] + c * pts[:, 1]).astype(np.float32)
    return out


@pytest.fixture(scope="module")
def keyframes(synth):
    seq = synth.Sequence("hdl64", 48, seed=11)
    return [cloud_of(seq, i) for i in range(48)]


# ---------------------------------------------------------------------------------------------------------
# oracle (CPU)
# ---------------------------------------------------------------------------------------------------------
def test_oracle_descriptor_matches_numpy_restatement(orc, keyframes):
    for pts in keyframes[:4]:
        d, rk, sk = orc.sc_make(pts)
        assert np.array_equal(d, numpy_descriptor(pts))
        assert np.allclose(rk, d.mean(axis=1), rtol=1e-13, atol=0) and np.allclose(sk, d.mean(axis=0), rtol=1e-13, atol=0)
    # out-of-range points are ignored, empty bins are 0, a bin holds the maximum height + LIDAR_HEIGHT
    pts = np.array([[10, 0.5, 1.0, 0], [10, 0.5, 3.0, 0], [100, 0, 9.0, 0], [-20, -0.1, -1.0, 0]], np.float32)
    d, _, _ = orc.sc_make(pts)
    assert d[2, 0] == 5.0 and d[5, 30] == 1.0 and np.count_nonzero(d) == 2


def numpy_sc_distance(a, b, search_ratio=0.1):
    """distanceBtnScanContext (SC:163-193) written independently with numpy reductions (as Eigen would vectorise them)."""
    S = a.shape[1]
    v1, v2 = a.mean(axis=0), b.mean(axis=0)
    norms = [np.linalg.norm(v1 - np.roll(v2, sh)) for sh in range(S)]
    arg = int(np.argmin(norms))
    radius = int(np.floor(0.5 * search_ratio * S + 0.5))
    space = sorted({arg} | {(arg + i + S) % S for i in range(1, radius + 1)} | {(arg - i + S) % S for i in range(1, radius + 1)})
    best, best_sh = 10000000.0, 0
    for sh in space:
        bs = np.roll(b, sh, axis=1)
        n1, n2 = np.linalg.norm(a, axis=0), np.linalg.norm(bs, axis=0)
        ok = (n1 != 0) & (n2 != 0)
        if not ok.any():
            continue
        sim = ((a * bs).sum(axis=0)[ok] / (n1[ok] * n2[ok])).sum() / ok.sum()
        if 1.0 - sim < best:
            best, best_sh = 1.0 - sim, sh
    return best, best_sh


def test_oracle_distance_matches_numpy_restatement(orc, keyframes):
    descs = [orc.sc_make(p)[0] for p in keyframes[:10]] + [orc.sc_make(rotz(keyframes[2], 33.0))[0]]
    for i in range(len(descs)):
        for j in range(len(descs)):
            do, so = orc.sc_distance(descs[i], descs[j])
            dn, sn = numpy_sc_distance(descs[i], descs[j])
            assert so == sn and abs(do - dn) < 1e-12, (i, j, do, dn, so, sn)


def test_oracle_empty_and_degenerate_clouds(orc):
    d, rk, sk = orc.sc_make(np.zeros((0, 4), np.float32))
    assert not d.any() and not rk.any() and not sk.any()
    dist, sh = orc.sc_distance(d, d)  # no populated sector pair: 0 / 0 -> NaN never beats the initial minimum (SC:149, :158)
    assert dist == 10000000 and sh == 0
    # the origin itself: atan(0 / 0) = NaN -> no quadrant matches; ring 1 / sector 1 by the clamps (SC:66-67)
    d, _, _ = orc.sc_make(np.array([[0, 0, 1, 0]], np.float32))
    assert d[0, 0] == 3.0


def test_oracle_yaw_rotation_is_a_column_shift(orc, keyframes):
    pts = keyframes[5]
    d0, _, _ = orc.sc_make(pts)
    for k in (1, 7, 31):
        # half a sector of margin is not needed for the DISTANCE: a 6-degree yaw moves nearly every point one sector on
        d1, _, _ = orc.sc_make(rotz(pts, 6.0 * k))
        dist, sh = orc.sc_distance(d1, d0)
        assert sh == k and dist < 0.05, (k, sh, dist)
        dist, sh = orc.sc_distance(d0, d1)
        assert sh == 60 - k and dist < 0.05
    # exact column roll: distance exactly 0 at the inverse shift
    dist, sh = orc.sc_distance(d0, np.roll(d0, 9, axis=1))
    assert sh == 51 and abs(dist) < 1e-15


def test_oracle_key_search_matches_the_reference_kdtree(orc, keyframes):
    """Ring-key candidates: the oracle's exhaustive search against the reference's own InvKeyTree (nanoflann, compiled from
    /root/reference by oracle/Makefile) — same indices, same float distances."""
    if not os.path.exists(os.path.join(os.path.dirname(orc.__file__), "_ref", "libref_sckeys.so")):
        pytest.skip("oracle/_ref/libref_sckeys.so not built (no /root/reference on this machine)")
    keys = np.array([orc.sc_make(p)[1] for p in keyframes], np.float64).astype(np.float32)
    rng = np.random.default_rng(3)
    more = (keys[rng.integers(0, len(keys), 600)] + rng.normal(0, 0.05, (600, 20))).astype(np.float32)
    allk = np.concatenate([keys, more])
    for qi in range(0, 40):
        q = allk[rng.integers(0, len(allk))] + np.float32(0.001)
        idx, d2 = orc.ref_sc_key_knn(allk, q, 3)
        mine = sorted(((orc.sc_key_dist(q, allk[i]), i) for i in range(len(allk))))[:3]
        assert [m[1] for m in mine] == list(idx), qi
        assert np.array_equal(np.array([m[0] for m in mine], np.float32), d2)


def test_oracle_matches_committed_goldens(orc, golden, golden2):
    """Descriptors, keys and pairwise distances of the six golden scans; ring-key candidates equal to what the REFERENCE's kd-tree
    returned in the build container (tests/golden/make_golden_r2.py)."""
    descs = []
    for i in range(6):
        d, rk, sk = orc.sc_make(golden[f"scan{i}"])
        assert np.array_equal(d, golden2[f"sc{i}_desc"]) and np.array_equal(rk, golden2[f"sc{i}_ringkey"]) and np.array_equal(sk, golden2[f"sc{i}_sectorkey"])
        descs.append(d)
    for i in range(6):
        for j in range(6):
            dist, sh = orc.sc_distance(descs[i], descs[j])
            assert dist == golden2["sc_dist"][i, j] and sh == golden2["sc_shift"][i, j]
    bank, q = golden2["sc_key_bank"], golden2["sc_key_q"]
    for k in range(len(q)):
        mine = sorted(((orc.sc_key_dist(q[k], bank[i]), i) for i in range(len(bank))))[:3]
        assert [m[1] for m in mine] == list(golden2["ref_sc_key_idx"][k])
        assert np.array_equal(np.array([m[0] for m in mine], np.float32), golden2["ref_sc_key_d2"][k])


def test_oracle_loop_detection_sequence(orc, keyframes):
    """detectLoopClosureID: early return below 31 key frames, then a revisit of key frame 4 (rotated by 3 sectors) is found.
    (With the reference's TREE_MAKING_PERIOD_ = 30 the search set would still be the single key of the first rebuild, SC:227-238.)"""
    m = orc.SCManager(orc.sc_params(tree_making_period=5))
    for i, pts in enumerate(keyframes[:40]):
        m.add(pts)
        r = m.detect()
        if i < 30:
            assert r == (-1, 0.0, 10000000.0, 0)
    m.add(rotz(keyframes[4], 18.0))
    lid, yaw, md, nn = m.detect()
    assert lid == 4 and nn == 4 and md < 0.1
    assert abs(yaw - np.float32(np.deg2rad(18.0))) < 1e-6


# ---------------------------------------------------------------------------------------------------------
# CUDA path through the C ABI
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_descriptor_keys_and_distance_bit_exact(cabi, orc, keyframes):
    g = cabi.SCManager()
    frames = keyframes[:12] + [rotz(keyframes[3], 42.0), np.zeros((0, 4), np.float32), np.array([[0, 0, 1, 0], [3, -4, 2, 0]], np.float32)]
    descs = []
    for pts in frames:
        g.add(pts)
        d, rk, sk = g.get(-1)
        do, rko, sko = orc.sc_make(pts)
        assert np.array_equal(d, do) and np.array_equal(rk, rko) and np.array_equal(sk, sko)
        descs.append(do)
    assert len(g) == len(frames)
    for i, j in [(0, 1), (3, 12), (12, 3), (5, 5), (2, 13), (13, 13), (14, 7)]:
        assert g.distance_between(i, j) == orc.sc_distance(descs[i], descs[j]), (i, j)
        assert g.distance(descs[i], descs[j]) == orc.sc_distance(descs[i], descs[j])
    assert g.launch_count() > 0
    g.close()


@pytest.mark.gpu
def test_gpu_scancontext_goldens(cabi, golden, golden2):
    """The CUDA path against the committed golden vectors (no oracle in the loop)."""
    g = cabi.SCManager()
    for i in range(6):
        g.add(golden[f"scan{i}"])
        d, rk, sk = g.get(-1)
        assert np.array_equal(d, golden2[f"sc{i}_desc"]) and np.array_equal(rk, golden2[f"sc{i}_ringkey"]) and np.array_equal(sk, golden2[f"sc{i}_sectorkey"])
    for i in range(6):
        for j in range(6):
            dist, sh = g.distance_between(i, j)
            assert dist == golden2["sc_dist"][i, j] and sh == golden2["sc_shift"][i, j]
    g.close()


@pytest.mark.gpu
def test_gpu_loop_detection_matches_oracle(cabi, orc, synth):
    """A longer key-frame stream with revisits: every detectLoopClosureID result (id, yaw, distance, nearest index) equals the
    oracle's, through the tree-rebuild period and the exclude-recent window."""
    seq = synth.Sequence("vlp32", 90, seed=5)
    clouds = [cloud_of(seq, i, 5) for i in range(90)]
    stream = clouds[:50] + [rotz(clouds[7], 24.0), clouds[12]] + clouds[50:] + [rotz(clouds[20], 300.0), clouds[51]]
    p = dict(tree_making_period=10, num_exclude_recent=20)
    g = cabi.SCManager(cabi.sc_params(**p))
    o = orc.SCManager(orc.sc_params(**p))
    loops = 0
    for i, pts in enumerate(stream):
        g.add(pts); o.add(pts)
        rg, ro = g.detect(), o.detect()
        assert rg == ro, (i, rg, ro)
        loops += rg[0] >= 0
    assert loops >= 3
    g.close()


@pytest.mark.gpu
def test_gpu_resident_cloud_equals_host_cloud(cabi, orc, synth):
    """vilf_sc_make_and_save_resident takes getMapCloud(MapCloud) = /GlobalMap where it lies on the device."""
    seq = synth.Sequence("hdl64", 4, seed=2)
    od = cabi.Odometry(cabi.default_config(max_scan_points=116000, max_map_points=1 << 18))
    g = cabi.SCManager()
    for i in range(4):
        od.process_scan(seq[i][0])
        g.add_resident(od)
        pts = od.cloud(cabi.NO_REGISTERED)
        d, rk, sk = g.get(-1)
        do, rko, sko = orc.sc_make(pts)
        assert pts.shape[0] > 1000 and np.array_equal(d, do) and np.array_equal(rk, rko) and np.array_equal(sk, sko)
    g.close(); od.close()


@pytest.mark.gpu
def test_gpu_sc_argument_errors(cabi):
    with pytest.raises(cabi.VilfError):
        cabi.SCManager(cabi.sc_params(num_candidates=9))
    g = cabi.SCManager(cabi.sc_params(max_keyframes=2, max_points=100))
    with pytest.raises(cabi.VilfError):
        g.detect()  # nothing stored
    g.add(np.zeros((3, 4), np.float32))
    with pytest.raises(cabi.VilfError):
        g.add(np.zeros((101, 4), np.float32))
    g.add(np.zeros((3, 4), np.float32))
    with pytest.raises(cabi.VilfError):
        g.add(np.zeros((3, 4), np.float32))  # max_keyframes
    with pytest.raises(cabi.VilfError):
        g.get(5)
    g.close()
