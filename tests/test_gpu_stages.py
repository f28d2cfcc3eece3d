"""Stage-level parity of the CUDA path (through the C ABI) against the CPU oracle and the committed goldens.
Bit-exact for indices / fp32 geometry; fp64 fits bit-exact (same operation order, no FMA); solve within 1e-9."""
import numpy as np
import pytest

from conftest import check_knn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g16(cabi):
    o = cabi.Odometry(cabi.default_config(n_scan=16, n_rings=16, max_scan_points=20000, max_map_points=1 << 17))
    yield o
    o.close()


@pytest.fixture(scope="module")
def g64(cabi):
    o = cabi.Odometry(cabi.default_config(max_scan_points=120000, max_map_points=1 << 18))
    yield o
    o.close()


def test_native_library_is_the_path(cabi, g16):
    import os
    assert os.path.exists(cabi.LIB_PATH)
    with open("/proc/self/maps") as f:
        assert "libvilf_cuda.so" in f.read()
    assert g16.launch_count() >= 0


# ---------------- stage 1 ----------------
def test_extract_golden(g16, golden):
    ne, ns = g16.feature_extract(golden["scan0"])
    e, es = g16.features(0)
    s, ss = g16.features(1)
    assert (ne, ns) == (len(golden["edge_src"]), len(golden["surf_src"]))
    assert np.array_equal(es, golden["edge_src"]) and np.array_equal(ss, golden["surf_src"])
    assert np.array_equal(e, golden["edge"]) and np.array_equal(s, golden["surf"])


def test_extract_hdl64_vs_oracle(g64, orc, hdl64_frames):
    cfg = orc.config()
    for xyzi, _ in hdl64_frames[:3]:
        oe, oes, os_, oss = orc.extract(cfg, xyzi)
        g64.feature_extract(xyzi)
        e, es = g64.features(0)
        s, ss = g64.features(1)
        assert np.array_equal(es, oes) and np.array_equal(ss, oss)
        assert np.array_equal(e, oe) and np.array_equal(s, os_)


def test_extract_firing_order_input(cabi, orc, synth):
    """Arrival order != ring-major (real drivers emit column by column): the ring binning must be a stable partition."""
    seq = synth.Sequence("vlp32", 1, seed=2, order=1)
    xyzi, _ = seq[0]
    g = cabi.Odometry(cabi.default_config(n_scan=32, n_rings=32, max_scan_points=65536, max_map_points=1 << 16))
    oe, oes, os_, oss = orc.extract(orc.config(n_scan=32, n_rings=32), xyzi)
    g.feature_extract(xyzi)
    assert np.array_equal(g.features(0)[1], oes) and np.array_equal(g.features(1)[1], oss)
    g.close()


def test_extract_explicit_rings_128(cabi, orc, synth):
    seq = synth.Sequence("beams128", 1, seed=4)
    xyzi, ring = seq[0]
    g = cabi.Odometry(cabi.default_config(n_scan=0, n_rings=128, max_scan_points=270000, max_map_points=1 << 16))
    oe, oes, os_, oss = orc.extract(orc.config(n_scan=0, n_rings=128), xyzi, ring)
    ne, ns = g.feature_extract(xyzi, ring)
    assert ne == len(oes) and ns == len(oss)
    assert np.array_equal(g.features(0)[1], oes) and np.array_equal(g.features(1)[1], oss)
    with pytest.raises(cabi.VilfError):
        g.feature_extract(xyzi)  # n_scan == 0 needs ring ids
    g.close()


def test_extract_edge_cases(g16, orc, golden, cabi):
    cfg = orc.config(n_scan=16, n_rings=16)
    empty = np.zeros((0, 4), np.float32)
    assert g16.feature_extract(empty) == (0, 0)
    x = golden["scan0"]
    # ragged: rings below the 131-point floor (FE:179) are skipped, rings exactly at 131 are used
    for keep in (100, 130, 131, 140, 300):
        parts, cnt = [], {}
        for p in x:
            d = float(np.sqrt(np.float32(p[0] * p[0] + p[1] * p[1])))
            if d < 3.0:
                continue
            rid = int((np.arctan(float(p[2]) / d) * 180 / np.pi + 15) / 2 + 0.5)
            if 0 <= rid <= 15 and cnt.get(rid, 0) < keep:
                cnt[rid] = cnt.get(rid, 0) + 1
                parts.append(p)
        sub = np.asarray(parts, np.float32)
        oe, oes, os_, oss = orc.extract(cfg, sub)
        g16.feature_extract(sub)
        assert np.array_equal(g16.features(0)[1], oes) and np.array_equal(g16.features(1)[1], oss), keep
    # every point out of range
    far = x.copy(); far[:, :2] *= 1000
    assert g16.feature_extract(far) == (0, 0)
    # over capacity -> loud error, not truncation
    with pytest.raises(cabi.VilfError) as e:
        g16.feature_extract(np.zeros((20001, 4), np.float32))
    assert e.value.code == 3


def test_wrong_scan_number_goes_to_ring0(cabi, orc, golden):
    """FE:103-106: an unsupported N_SCAN bins everything into ring 0; a ring longer than the kernel's limit must fail loudly."""
    x = golden["scan0"][:9000]
    g = cabi.Odometry(cabi.default_config(n_scan=7, n_rings=7, max_scan_points=20000, max_map_points=1 << 16))
    oe, oes, os_, oss = orc.extract(orc.config(n_scan=7, n_rings=7), x)
    g.feature_extract(x)
    assert np.array_equal(g.features(0)[1], oes) and np.array_equal(g.features(1)[1], oss)
    with pytest.raises(cabi.VilfError) as e:
        g.feature_extract(np.concatenate([golden["scan0"], golden["scan1"]])[:19000])
    assert e.value.code == 4
    g.close()


def test_extract_drops_non_finite_returns(cabi, orc, synth, g64, hdl64_frames):
    """Organized clouds of real drivers carry NaN / Inf rows.  FE:56-57 leaves them in the cloud; the x86 reference drops them
    at the scanID range test (int(NaN) == INT_MIN).  The device must drop them too: same edge / surf indices as the oracle,
    which must equal the extraction of the scan with those rows deleted (indices remapped)."""
    x = hdl64_frames[1][0].copy()
    rng = np.random.default_rng(9)
    bad = np.sort(rng.choice(x.shape[0], 700, replace=False))
    x[bad[0::4], 0] = np.nan
    x[bad[1::4], 2] = np.nan
    x[bad[2::4], 1] = np.inf
    x[bad[3::4], 2] = -np.inf
    cfg = orc.config()
    oe, oes, os_, oss = orc.extract(cfg, x)
    g64.feature_extract(x)
    e, es = g64.features(0)
    s, ss = g64.features(1)
    assert np.array_equal(es, oes) and np.array_equal(ss, oss)
    assert np.array_equal(e, oe) and np.array_equal(s, os_)
    assert np.isfinite(e).all() and np.isfinite(s).all() and len(es) > 1000
    keep = np.setdiff1d(np.arange(x.shape[0]), bad)
    ce, ces, cs, css = orc.extract(cfg, x[keep])
    assert np.array_equal(keep[ces], oes) and np.array_equal(keep[css], oss)
    # explicit ring ids: non-finite returns are dropped in that mode as well
    seq = synth.Sequence("beams128", 1, seed=4)
    y, ring = seq[0]
    y = y.copy()
    y[5::97, 1] = np.nan
    g = cabi.Odometry(cabi.default_config(n_scan=0, n_rings=128, max_scan_points=270000, max_map_points=1 << 16))
    _, oes, _, oss = orc.extract(orc.config(n_scan=0, n_rings=128), y, ring)
    g.feature_extract(y, ring)
    assert np.array_equal(g.features(0)[1], oes) and np.array_equal(g.features(1)[1], oss)
    assert np.isfinite(g.features(1)[0]).all()
    g.close()


# ---------------- voxel grid / crop box ----------------
def test_voxel_goldens(g16, golden):
    for leaf in (0.4, 0.8):
        v, guard = g16.voxel_downsample(golden["surf"], leaf)
        assert not guard and np.array_equal(v, golden[f"vox_surf_{leaf}"])
    assert np.array_equal(g16.voxel_downsample(golden["edge"], 0.4)[0], golden["vox_edge_0.4"])
    c = golden["crop_center"]
    assert np.array_equal(g16.crop_box(golden["surf"], c - 15.0, c + 15.0), golden["crop_surf"])


def test_voxel_vs_oracle_hdl64(g64, orc, hdl64_frames):
    cfg = orc.config()
    _, _, surf, _ = orc.extract(cfg, hdl64_frames[0][0])
    for leaf in (0.2, 0.4, 0.8, 3.0):
        o, _ = orc.voxel_grid(surf, leaf)
        v, guard = g64.voxel_downsample(surf, leaf)
        assert not guard and np.array_equal(o, v), leaf
    c = np.array([4.0, -2.0, 0.3])
    oc = orc.crop_box(surf, c - 25, c + 25)
    assert np.array_equal(g64.crop_box(surf, c - 25, c + 25), oc)
    assert np.array_equal(g64.crop_voxel_downsample(surf, c, 25.0, 0.8), orc.voxel_grid(oc, 0.8)[0])


def test_voxel_edge_cases_and_properties(g64, orc):
    empty = np.zeros((0, 4), np.float32)
    assert g64.voxel_downsample(empty, 0.4)[0].shape == (0, 4)
    one = np.array([[1.5, -2.5, 0.25, 0.7]], np.float32)
    assert np.array_equal(g64.voxel_downsample(one, 0.4)[0], one)
    same = np.repeat(one, 1000, axis=0)
    assert np.array_equal(g64.voxel_downsample(same, 0.4)[0], orc.voxel_grid(same, 0.4)[0])
    big = np.array([[0, 0, 0, 0], [3e3, 3e3, 3e3, 0]], np.float32)
    out, guard = g64.voxel_downsample(big, 0.01)  # PCL's int32 guard: output = input
    assert guard and np.array_equal(out, big)
    # crop that removes everything
    assert g64.crop_box(same, [10, 10, 10], [11, 11, 11]).shape == (0, 4)
    assert g64.crop_voxel_downsample(same, [100.0, 100.0, 100.0], 1.0, 0.4).shape == (0, 4)
    # closed box: points exactly on the bound stay (CropBox uses < / >)
    edge = np.array([[1.0, 1.0, 1.0, 0], [1.0000001, 1.0, 1.0, 0]], np.float32)
    assert np.array_equal(g64.crop_box(edge, [0, 0, 0], [1, 1, 1]), orc.crop_box(edge, [0, 0, 0], [1, 1, 1]))
    # size-independent properties on a large random cloud (260k points, 0.2 m leaf: config 5 shape)
    rng = np.random.default_rng(7)
    pts = rng.uniform(-60, 60, (260000, 4)).astype(np.float32)
    pts[:, 2] = rng.uniform(-3, 12, 260000)
    v, _ = g64.voxel_downsample(pts, 0.2)
    inv = np.float32(1) / np.float32(0.2)
    cells = np.floor(v[:, :3] * inv).astype(np.int64)
    assert len(np.unique(cells, axis=0)) == len(v)
    assert len(v) == len(np.unique(np.floor(pts[:, :3] * inv).astype(np.int64), axis=0))
    assert np.array_equal(g64.voxel_downsample(v, 0.2)[0], v)  # idempotent
    key = (cells[:, 2] - cells[:, 2].min()) * 10**8 + (cells[:, 1] - cells[:, 1].min()) * 10**4 + (cells[:, 0] - cells[:, 0].min())
    assert np.all(np.diff(key) > 0)  # ascending voxel index = PCL's output order
    assert abs(float(v[:, 3].astype(np.float64).mean()) - float(pts[:, 3].astype(np.float64).mean())) < 0.01


# ---------------- 5-NN ----------------
def test_knn_golden_and_reference_kdtree(g16, golden):
    gi, gd = g16.knn5(golden["knn_map"], golden["knn_q"])
    check_knn(gi, gd, golden["ref_knn_idx"], golden["ref_knn_d2"])
    full = golden["ref_knn_d2"][:, 4] < 1.0
    assert full.sum() > 50 and np.array_equal(gi[full], golden["ref_knn_idx"][full])


def test_knn_vs_oracle(g64, orc, hdl64_frames):
    cfg = orc.config()
    _, _, surf, _ = orc.extract(cfg, hdl64_frames[0][0])
    mp, _ = orc.voxel_grid(surf, 0.8)
    rng = np.random.default_rng(0)
    q = mp[rng.integers(0, len(mp), 4000)].copy()
    q[:, :3] += rng.normal(0, 0.2, (4000, 3)).astype(np.float32)
    for m in (mp, surf):  # one-per-voxel map and the raw first-frame map (many points per cell)
        oi, od = orc.knn(m, q)
        gi, gd = g64.knn5(m, q)
        check_knn(gi, gd, oi, od)


def test_knn_edge_cases(g64, orc):
    q = np.array([[0.1, 0.2, 0.3, 0], [50, 50, 50, 0]], np.float32)
    gi, gd = g64.knn5(np.zeros((0, 4), np.float32), q)
    assert (gi == -1).all()
    mp = np.array([[0, 0, 0, 0], [0.5, 0, 0, 0], [0, 0.5, 0, 0]], np.float32)
    gi, gd = g64.knn5(mp, q)
    oi, od = orc.knn(mp, q)
    assert np.array_equal(gi[0, :3], oi[0, :3]) and (gi[0, 3:] == -1).all() and (gi[1] == -1).all()
    assert np.all(gd[0, 3:] > 1e30) and np.all(gd[1] > 1e30)  # nothing inside the gate: FLT_MAX
    # exact ties: duplicate map points -> lowest index first, distances equal
    dup = np.repeat(np.array([[1, 1, 1, 0]], np.float32), 8, axis=0)
    gi, gd = g64.knn5(dup, np.array([[1.1, 1, 1, 0]], np.float32))
    assert list(gi[0]) == [0, 1, 2, 3, 4] and len(set(gd[0])) == 1
    # negative coordinates / cell boundaries
    rng = np.random.default_rng(3)
    mp = np.zeros((20000, 4), np.float32); mp[:, :3] = np.round(rng.uniform(-8, 8, (20000, 3)) * 4) / 4
    qq = np.zeros((2000, 4), np.float32); qq[:, :3] = np.round(rng.uniform(-8, 8, (2000, 3)) * 2) / 2
    oi, od = orc.knn(mp, qq)
    gi, gd = g64.knn5(mp, qq)
    inside = od < 1.0
    assert np.array_equal(gd[inside], od[inside])


def test_knn_bruteforce_large(cabi):
    """Full-size property (1e6-point map, config 3/5 scale): inside the gate the result equals brute force on a sample."""
    g64 = cabi.Odometry(cabi.default_config(max_scan_points=4096, max_map_points=1 << 20))
    rng = np.random.default_rng(11)
    m = 1_000_000
    mp = np.zeros((m, 4), np.float32)
    mp[:, 0] = rng.uniform(-100, 100, m); mp[:, 1] = rng.uniform(-100, 100, m); mp[:, 2] = rng.uniform(-2, 6, m)
    q = mp[rng.integers(0, m, 200000)].copy()
    q[:, :3] += rng.normal(0, 0.15, (200000, 3)).astype(np.float32)
    gi, gd = g64.knn5(mp, q)
    for r in rng.integers(0, len(q), 40):
        d = q[r, :3][None] - mp[:, :3]
        bf = ((d[:, 0] * d[:, 0]) + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        o = np.argsort(bf, kind="stable")[:5]
        ins = bf[o] < 1.0
        assert np.array_equal(gd[r][ins], bf[o][ins])
        assert np.array_equal(gi[r][ins], o[ins]) or len(set(bf[o][ins])) < ins.sum()
    assert np.all(np.diff(gd, axis=1) >= 0)  # ascending
    g64.close()


# ---------------- factors / normal equations / solve ----------------
def test_factors_golden(g16, golden):
    g16.set_state(golden["f3_state"], golden["f3_map_edge"], golden["f3_map_surf"])
    f = g16.factors(golden["fac_pose"], golden["f3_ds_edge"], golden["f3_ds_surf"])
    assert np.array_equal(f["edge_valid"], golden["fac_edge_valid"]) and np.array_equal(f["surf_valid"], golden["fac_surf_valid"])
    ev, sv = f["edge_valid"].astype(bool), f["surf_valid"].astype(bool)
    assert np.array_equal(f["edge_nn"][ev], golden["fac_edge_nn"][ev]) and np.array_equal(f["surf_nn"][sv], golden["fac_surf_nn"][sv])
    assert np.array_equal(f["edge_d2"][ev], golden["fac_edge_d2"][ev]) and np.array_equal(f["surf_d2"][sv], golden["fac_surf_d2"][sv])
    # line end points / plane parameters: bit-exact (same operation order, no FMA); eigenvector sign is immaterial (T4)
    assert np.array_equal(f["surf_nd"], golden["fac_surf_nd"])
    a, b = f["edge_ab"][:, :3], f["edge_ab"][:, 3:]
    ga, gb = golden["fac_edge_ab"][:, :3], golden["fac_edge_ab"][:, 3:]
    same = np.all(a == ga, axis=1) & np.all(b == gb, axis=1)
    flip = np.all(a == gb, axis=1) & np.all(b == ga, axis=1)
    assert np.all(same | flip)


def test_factors_vs_oracle_hdl64(cabi, orc, hdl64_frames):
    cfg = orc.config()
    o = orc.Odometry(cfg)
    g = cabi.Odometry(cabi.default_config(max_scan_points=120000, max_map_points=1 << 18))
    for xyzi, _ in hdl64_frames[:3]:
        o.process_scan(xyzi)
        g.process_scan(xyzi)
    pose = o.state()[:7]
    de, ds = o.cloud(orc.DS_EDGE), o.cloud(orc.DS_SURF)
    fo = orc.factors(cfg, pose, de, ds, o.cloud(orc.MAP_EDGE), o.cloud(orc.MAP_SURF))
    fg = g.factors(pose, de, ds)
    assert np.array_equal(fg["edge_valid"], fo["edge_valid"]) and np.array_equal(fg["surf_valid"], fo["surf_valid"])
    assert np.array_equal(fg["surf_nd"], fo["surf_nd"])
    assert np.abs(np.abs(fg["edge_ab"][:, :3] - fg["edge_ab"][:, 3:]) - np.abs(fo["edge_ab"][:, :3] - fo["edge_ab"][:, 3:])).max() == 0
    g.close()


def test_normal_equations_and_solve_golden(g16, orc, golden):
    cfg = orc.config(n_scan=16, n_rings=16)
    f = {k[4:]: golden[k] for k in golden.files if k.startswith("fac_") and k != "fac_pose"}
    pab, pnd = orc.pack_factors(golden["f3_ds_edge"], golden["f3_ds_surf"], f)
    H, gvec, cost = g16.normal_equations(golden["fac_pose"], pab, pnd)
    scale = np.abs(golden["ne_H"]).max()
    assert np.abs(H - golden["ne_H"]).max() <= 1e-12 * scale
    assert np.abs(gvec - golden["ne_g"]).max() <= 1e-12 * max(1.0, np.abs(golden["ne_g"]).max())
    assert abs(cost - golden["ne_cost"][0]) <= 1e-13 * cost
    pose, tr, term = g16.solve(golden["fac_pose"], pab, pnd, 4)
    gt = golden["solve_trace"]
    assert term == golden["solve_term"][0] and tr.shape == gt.shape
    assert np.array_equal(tr[:, :3], gt[:, :3])                      # same accept / reject schedule
    assert np.abs(tr[:, 3] - gt[:, 3]).max() <= 1e-10 * gt[0, 3]     # costs
    assert np.abs(tr[:, 7] / gt[:, 7] - 1).max() < 1e-6              # trust-region radii
    assert np.abs(pose - golden["solve_pose"]).max() < 1e-9
    # no factors at all: pose untouched (ceres returns immediately)
    p0, tr0, term0 = g16.solve(golden["fac_pose"], np.zeros((0, 9)), np.zeros((0, 7)), 4)
    assert np.array_equal(p0, golden["fac_pose"]) and term0 == 4
    # max_iters = 0: evaluation only
    p1, tr1, _ = g16.solve(golden["fac_pose"], pab, pnd, 0)
    assert np.array_equal(p1, golden["fac_pose"]) and len(tr1) == 1


def test_solve_random_problems_vs_oracle(g16, orc):
    rng = np.random.default_rng(9)
    for trial in range(6):
        x = np.zeros(7); x[:3] = rng.normal(0, 0.02, 3); x[3] = np.sqrt(1 - (x[:3] ** 2).sum()); x[4:] = rng.normal(0, 0.3, 3)
        ke, ks = int(rng.integers(0, 200)), int(rng.integers(1, 800))
        p = rng.uniform(-30, 30, (ke, 3)); d = rng.normal(size=(ke, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        c = p + rng.normal(0, 0.05, (ke, 3))
        pab = np.concatenate([p, c + 0.1 * d, c - 0.1 * d], axis=1)
        ps = rng.uniform(-30, 30, (ks, 3)); n = rng.normal(size=(ks, 3)); n /= np.linalg.norm(n, axis=1, keepdims=True)
        pnd = np.concatenate([ps, n, (-(n * ps).sum(1) + rng.normal(0, 0.05, ks))[:, None]], axis=1)
        po, tro, termo = orc.solve(0.1, 4, x, pab, pnd)
        pg, trg, termg = g16.solve(x, pab, pnd, 4)
        assert termo == termg and tro.shape == trg.shape, trial
        assert np.array_equal(tro[:, :3], trg[:, :3])
        assert np.abs(po - pg).max() < 1e-9, trial


@pytest.mark.parametrize("ke,ks", [(3000, 18000), (0, 5780), (0, 5840), (510, 0), (4000, 1)])
def test_solve_factor_pool_boundary_and_overflow(g16, orc, ke, ks):
    """k_solve keeps the accepted factors of each CTA in a shared-memory pool (5100 doubles: 10 per line, 7 per plane
    factor) and evaluates a CTA whose share does not fit from global memory instead: sizes far beyond the pool, sizes at
    which only some of the 8 CTAs fit, and one-kind-only problems must all follow the oracle's schedule."""
    rng = np.random.default_rng(ke * 7 + ks)
    x = np.zeros(7); x[:3] = rng.normal(0, 0.01, 3); x[3] = np.sqrt(1 - (x[:3] ** 2).sum()); x[4:] = rng.normal(0, 0.2, 3)
    p = rng.uniform(-30, 30, (ke, 3)); d = rng.normal(size=(ke, 3)); d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-9)
    c = p + rng.normal(0, 0.05, (ke, 3))
    pab = np.concatenate([p, c + 0.1 * d, c - 0.1 * d], axis=1)
    ps = rng.uniform(-30, 30, (ks, 3)); n = rng.normal(size=(ks, 3)); n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-9)
    pnd = np.concatenate([ps, n, (-(n * ps).sum(1) + rng.normal(0, 0.05, ks))[:, None]], axis=1)
    H, gv, cost = g16.normal_equations(x, pab, pnd)
    Ho, go, co = orc.normal_eq(0.1, x, pab, pnd)
    assert np.abs(H - Ho).max() <= 1e-11 * np.abs(Ho).max() and np.abs(gv - go).max() <= 1e-11 * max(1.0, np.abs(go).max())
    assert abs(cost - co) <= 1e-12 * co
    po, tro, termo = orc.solve(0.1, 4, x, pab, pnd)
    pg, trg, termg = g16.solve(x, pab, pnd, 4)
    assert termo == termg and tro.shape == trg.shape
    assert np.array_equal(tro[:, :3], trg[:, :3])
    assert np.abs(po - pg).max() < 1e-9


@pytest.mark.parametrize("leaf,spacing", [(0.2, 0.2), (0.1, 0.12), (0.4, 0.4)])
def test_knn5_dense_maps_fine_cells_early_exit(cabi, orc, leaf, spacing):
    """Maps filtered at a fine leaf get search cells finer than the gate radius and a shell-by-shell walk with early exit
    (k_knn.cu: group_knn5): the neighbours must still be exactly the oracle's (FLANN-order) five, index for index."""
    rng = np.random.default_rng(int(leaf * 1000))
    side = int(40.0 / spacing)
    gx, gy = np.meshgrid(np.arange(side) * spacing - 20.0, np.arange(side) * spacing - 20.0)
    ground = np.stack([gx.ravel(), gy.ravel(), np.full(side * side, -1.7)], 1)
    ground += rng.uniform(-0.3, 0.3, ground.shape) * spacing
    wall = np.stack([rng.uniform(-20, 20, 20000), np.full(20000, 7.0) + rng.normal(0, 0.02, 20000), rng.uniform(-1.7, 6.0, 20000)], 1)
    sparse = rng.uniform(-20, 20, (300, 3))  # isolated points: queries near them need every shell or have fewer than 5 inside the gate
    mp = np.concatenate([ground, wall, sparse]).astype(np.float32)
    mp = np.concatenate([mp, rng.random((mp.shape[0], 1), dtype=np.float32)], 1)
    q = np.concatenate([mp[rng.integers(0, mp.shape[0], 4000), :3] + rng.normal(0, 0.05, (4000, 3)), rng.uniform(-21, 21, (1000, 3))]).astype(np.float32)
    q = np.concatenate([q, np.zeros((q.shape[0], 1), np.float32)], 1)
    g = cabi.Odometry(cabi.default_config(edge_leaf=leaf, surf_leaf=2 * leaf, max_scan_points=20000, max_map_points=1 << 19))
    gi, gd = g.knn5(mp, q)
    oi, od = orc.knn(mp, q)
    check_knn(gi, gd, oi, od)
    full = od[:, 4] < 1.0
    assert full.sum() > 3000 and (~full).sum() > 100  # both regimes are exercised
    g.close()


def test_voxel_fuzz_run_lengths_and_sizes(cabi, orc):
    """Randomised clouds that stress the cluster path's centroid emitter (runs shorter / longer than a 32-position step, runs
    crossing warp and CTA boundaries, one giant voxel, everything cropped but a corner) and its radix passes (1..4 passes,
    sizes around the chunk / tile boundaries), on BOTH paths, bit-exact against the oracle's PCL restatement."""
    rng = np.random.default_rng(99)
    a = cabi.Odometry(cabi.default_config(max_scan_points=70000, max_map_points=1 << 16))
    b = cabi.Odometry(cabi.default_config(max_scan_points=70000, max_map_points=1 << 16, flags=cabi.FLAG_NO_CLUSTER))
    sizes = [1, 2, 31, 32, 33, 511, 512, 513, 4095, 4096, 4097, 8191, 12345, 32768, 60001]
    for case, n in enumerate(sizes):
        kind = case % 5
        if kind == 0:    # uniform noise, mostly singletons
            pts = rng.uniform(-30, 30, (n, 4))
        elif kind == 1:  # a few giant voxels (hundreds to thousands of points each) + noise
            centres = rng.uniform(-20, 20, (max(1, n // 1500), 3))
            pts = np.concatenate([centres[rng.integers(0, len(centres), n)] + rng.uniform(0, 0.3, (n, 3)), rng.random((n, 1))], 1)
        elif kind == 2:  # everything in ONE voxel
            pts = np.concatenate([np.array([3.0, -7.0, 1.0]) + rng.uniform(0.05, 0.3, (n, 3)), rng.random((n, 1))], 1)
        elif kind == 3:  # run lengths 1..70 in random order of arrival (runs straddle the 32-position steps)
            reps = rng.integers(1, 70, n)
            base = rng.uniform(-40, 40, (n, 3))
            idx = np.repeat(np.arange(n), reps)[:n]
            pts = np.concatenate([base[idx] + rng.uniform(0, 0.05, (n, 3)), rng.random((n, 1))], 1)
            pts = pts[rng.permutation(n)]
        else:            # thin wide slab: many key bits (4 passes at a fine leaf)
            pts = np.concatenate([rng.uniform(-900, 900, (n, 2)), rng.uniform(-2, 2, (n, 1)), rng.random((n, 1))], 1)
        pts = pts.astype(np.float32)
        for leaf in (0.4, 0.05 if kind == 4 else 0.8):
            o, ok = orc.voxel_grid(pts, leaf)
            va, ga = a.voxel_downsample(pts, leaf)
            vb, gb = b.voxel_downsample(pts, leaf)
            assert ga == gb == (not ok), (n, kind, leaf)
            assert np.array_equal(va, o) and np.array_equal(vb, o), (n, kind, leaf)
        c = pts[rng.integers(0, n), :3].astype(np.float64)
        oc = orc.crop_box(pts, c - 5.0, c + 5.0)
        want = orc.voxel_grid(oc, 0.4)[0] if len(oc) else np.zeros((0, 4), np.float32)
        assert np.array_equal(a.crop_voxel_downsample(pts, c, 5.0, 0.4), want), (n, kind)
        assert np.array_equal(b.crop_voxel_downsample(pts, c, 5.0, 0.4), want), (n, kind)
    a.close(); b.close()
