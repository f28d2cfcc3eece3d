"""TEST INFRASTRUCTURE — ctypes binding of the CPU oracle (oracle/_build/liborc.so).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
PARITY UNPINNED: the reference has no tests or golden vectors; see orc_pipeline.hpp.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liborc.so")
_REF = os.path.join(_HERE, "_ref", "libref_nanoflann.so")
_REF_SC = os.path.join(_HERE, "_ref", "libref_sckeys.so")


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_LIB) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB) for f in ("orc_capi.cpp", "orc_pipeline.hpp", "orc_math.hpp", "orc_depth.hpp", "orc_scancontext.hpp", "orc_rangeimage.hpp")
    ):
        subprocess.run(["make", "-C", _HERE, "_build/liborc.so"], check=True, capture_output=True)
    if os.path.isdir("/root/reference") and (force or not os.path.exists(_REF) or not os.path.exists(_REF_SC)):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


class Config(C.Structure):
    _fields_ = [
        ("n_scan", C.c_int), ("n_rings", C.c_int),
        ("lidar_min", C.c_double), ("lidar_max", C.c_double), ("edge_threshold", C.c_double),
        ("edge_leaf", C.c_double), ("surf_leaf", C.c_double), ("crop_half", C.c_double),
        ("knn_gate", C.c_double), ("huber", C.c_double),
        ("outer_iters", C.c_int), ("lm_max_iters", C.c_int), ("voxel_order", C.c_int), ("knn_ties", C.c_int),
        ("eig_alg", C.c_int), ("plane_alg", C.c_int), ("centroid_div", C.c_int), ("lm_solver", C.c_int),  # sensitivity alternates (orc_pipeline.hpp: Config)
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_odom_create.restype = C.c_void_p
        _lib.orc_odom_cloud_size.restype = C.c_int
        _lib.orc_odom_get_solves.restype = C.c_int
        _lib.orc_solve.restype = C.c_int
        _lib.orc_voxel_grid.restype = C.c_int
    return _lib


def ref_lib():
    """The reference's own vendored nanoflann kd-tree (None if oracle/_ref was never built)."""
    if not os.path.exists(_REF):
        return None
    return C.CDLL(_REF)


def config(**kw) -> Config:
    c = Config()
    lib().orc_default_config(C.byref(c))
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def extract(cfg: Config, xyzi, ring=None):
    """Stage 1 -> (edge [ne,4], edge_src [ne], surf [ns,4], surf_src [ns])."""
    xyzi = _f32(xyzi)
    n = xyzi.shape[0]
    edge = np.empty((max(n, 1), 4), np.float32); surf = np.empty((max(n, 1), 4), np.float32)
    es = np.empty(max(n, 1), np.int32); ss = np.empty(max(n, 1), np.int32)
    ne = C.c_int(); ns = C.c_int()
    rp = None
    if ring is not None:
        ring = np.ascontiguousarray(ring, dtype=np.uint16)
        rp = _p(ring, C.c_uint16)
    lib().orc_extract(C.byref(cfg), _p(xyzi, C.c_float), n, rp, _p(edge, C.c_float), _p(es, C.c_int), C.byref(ne),
                      _p(surf, C.c_float), _p(ss, C.c_int), C.byref(ns))
    return edge[:ne.value].copy(), es[:ne.value].copy(), surf[:ns.value].copy(), ss[:ns.value].copy()


def voxel_grid(pts, leaf: float, order_mode: int = 0):
    pts = _f32(pts)
    n = pts.shape[0]
    out = np.empty((max(n, 1), 4), np.float32)
    no = C.c_int()
    ok = lib().orc_voxel_grid(_p(pts, C.c_float), n, C.c_float(leaf), order_mode, _p(out, C.c_float), C.byref(no))
    return out[:no.value].copy(), bool(ok)


def crop_box(pts, mn, mx):
    pts = _f32(pts)
    n = pts.shape[0]
    out = np.empty((max(n, 1), 4), np.float32)
    no = C.c_int()
    mn = _f64(mn); mx = _f64(mx)
    lib().orc_crop_box(_p(pts, C.c_float), n, _p(mn, C.c_double), _p(mx, C.c_double), _p(out, C.c_float), C.byref(no))
    return out[:no.value].copy()


def _knn(fn, mp, q, k):
    mp = _f32(mp); q = _f32(q)
    nq = q.shape[0]
    idx = np.empty((nq, k), np.int32); d2 = np.empty((nq, k), np.float32)
    fn(_p(mp, C.c_float), mp.shape[0], _p(q, C.c_float), nq, k, _p(idx, C.c_int), _p(d2, C.c_float))
    return idx, d2


def knn(mp, q, k: int = 5, canonical: bool = False):
    """Oracle kd-tree (restated FLANN single index): exact k-NN, ascending squared fp32 distances.  canonical=True resolves
    equal distances (tie class T2) by ascending map index — the CUDA path's rule — instead of FLANN's visiting order."""
    return _knn(lib().orc_knn_canonical if canonical else lib().orc_knn, mp, q, k)


def ref_knn(mp, q, k: int = 5):
    """The reference's vendored nanoflann 1.3.2 on the same inputs (oracle/_ref)."""
    r = ref_lib()
    if r is None:
        raise RuntimeError("oracle/_ref/libref_nanoflann.so not built")
    return _knn(r.ref_nanoflann_knn, mp, q, k)


def camera_cloud(scan, T):
    """feature_tracker_node.cpp:348-361: camera field-of-view filter + pcl::transformPointCloud(LIDAR_CAMERA_EX)."""
    scan = _f32(scan); T = _f64(np.asarray(T).reshape(16))
    out = np.empty((max(scan.shape[0], 1), 4), np.float32)
    no = C.c_int()
    lib().orc_camera_cloud(_p(scan, C.c_float), scan.shape[0], _p(T, C.c_double), _p(out, C.c_float), C.byref(no))
    return out[:no.value].copy()


def feature_depth(cloud, feats, num_bins: int = 360):
    """getFeatureDepth steps 4.1-4.4 (feature_tracker_node.cpp:54-140): (depth [m], 3-NN indices [m,3])."""
    cloud = _f32(cloud); feats = np.ascontiguousarray(feats, dtype=np.float32)
    m = feats.shape[0]
    d = np.empty(max(m, 1), np.float32); nn = np.empty((max(m, 1), 3), np.int32)
    lib().orc_feature_depth(_p(cloud, C.c_float), cloud.shape[0], _p(feats, C.c_float), m, num_bins, _p(d, C.c_float), _p(nn, C.c_int))
    return d[:m].copy(), nn[:m].copy()


def node_outputs(rt12, last):
    """feature_tracker_node.cpp:388-401, :445-446: (relative pose [7], path pose [7], new last [7])."""
    rt12 = _f64(rt12); last = _f64(last).copy()
    rel = np.empty(7); path = np.empty(7)
    lib().orc_node_outputs(_p(rt12, C.c_double), _p(last, C.c_double), _p(rel, C.c_double), _p(path, C.c_double))
    return rel, path, last


def factors(cfg: Config, pose, edge, surf, map_e, map_s):
    """Data association at `pose` -> dict of per-point outputs for edge and surf features."""
    pose = _f64(pose); edge = _f32(edge); surf = _f32(surf); map_e = _f32(map_e); map_s = _f32(map_s)
    ne, ns = edge.shape[0], surf.shape[0]
    ev = np.zeros(max(ne, 1), np.uint8); eab = np.zeros((max(ne, 1), 6)); enn = np.zeros((max(ne, 1), 5), np.int32); ed2 = np.zeros((max(ne, 1), 5), np.float32)
    sv = np.zeros(max(ns, 1), np.uint8); snd = np.zeros((max(ns, 1), 4)); snn = np.zeros((max(ns, 1), 5), np.int32); sd2 = np.zeros((max(ns, 1), 5), np.float32)
    lib().orc_factors(C.byref(cfg), _p(pose, C.c_double), _p(edge, C.c_float), ne, _p(surf, C.c_float), ns,
                      _p(map_e, C.c_float), map_e.shape[0], _p(map_s, C.c_float), map_s.shape[0],
                      _p(ev, C.c_uint8), _p(eab, C.c_double), _p(enn, C.c_int), _p(ed2, C.c_float),
                      _p(sv, C.c_uint8), _p(snd, C.c_double), _p(snn, C.c_int), _p(sd2, C.c_float))
    return dict(edge_valid=ev[:ne], edge_ab=eab[:ne], edge_nn=enn[:ne], edge_d2=ed2[:ne],
                surf_valid=sv[:ns], surf_nd=snd[:ns], surf_nn=snn[:ns], surf_d2=sd2[:ns])


def pack_factors(edge, surf, f):
    """(edge_pab [ke,9], surf_pnd [ks,7]) of the accepted factors, in residual-block order."""
    ev = f["edge_valid"].astype(bool); sv = f["surf_valid"].astype(bool)
    pab = np.concatenate([np.asarray(edge, np.float32)[ev, :3].astype(np.float64), f["edge_ab"][ev]], axis=1)
    pnd = np.concatenate([np.asarray(surf, np.float32)[sv, :3].astype(np.float64), f["surf_nd"][sv]], axis=1)
    return np.ascontiguousarray(pab), np.ascontiguousarray(pnd)


def normal_eq(huber: float, pose, pab, pnd):
    pose = _f64(pose); pab = _f64(pab); pnd = _f64(pnd)
    H = np.zeros(21); g = np.zeros(6); cost = C.c_double()
    lib().orc_normal_eq(C.c_double(huber), _p(pose, C.c_double), _p(pab, C.c_double), pab.shape[0], _p(pnd, C.c_double), pnd.shape[0],
                        _p(H, C.c_double), _p(g, C.c_double), C.byref(cost))
    return H, g, cost.value


def solve(huber: float, max_iters: int, pose, pab, pnd):
    """ceres::Solve restated -> (pose_out [7], trace [rows,16], termination)."""
    pose = _f64(pose).copy(); pab = _f64(pab); pnd = _f64(pnd)
    tr = np.zeros((max_iters + 2, 16)); nr = C.c_int()
    term = lib().orc_solve(C.c_double(huber), max_iters, _p(pose, C.c_double), _p(pab, C.c_double), pab.shape[0], _p(pnd, C.c_double), pnd.shape[0],
                           _p(tr, C.c_double), tr.shape[0], C.byref(nr))
    return pose, tr[:nr.value].copy(), term


def se3_plus(x, delta):
    x = _f64(x); delta = _f64(delta); out = np.zeros(7)
    lib().orc_se3_plus(_p(x, C.c_double), _p(delta, C.c_double), _p(out, C.c_double))
    return out


def edge_eval(pose, pab):
    pose = _f64(pose); pab = _f64(pab); r = np.zeros(3); J = np.zeros((3, 6))
    lib().orc_edge_eval(_p(pose, C.c_double), _p(pab, C.c_double), _p(r, C.c_double), _p(J, C.c_double))
    return r, J


def surf_eval(pose, pnd):
    pose = _f64(pose); pnd = _f64(pnd); r = np.zeros(1); J = np.zeros((1, 6))
    lib().orc_surf_eval(_p(pose, C.c_double), _p(pnd, C.c_double), _p(r, C.c_double), _p(J, C.c_double))
    return r, J


def eig3(Cm, alt: bool = False):
    """EM:150.  alt: Eigen 3.3.7's tridiagonal QR iteration instead of cyclic Jacobi (sensitivity alternate)."""
    Cm = _f64(Cm); w = np.zeros(3); V = np.zeros((3, 3))
    (lib().orc_eig3_alt if alt else lib().orc_eig3)(_p(Cm, C.c_double), _p(w, C.c_double), _p(V, C.c_double))
    return w, V


def lstsq5x3(A, b, alt: bool = False):
    """EM:198.  alt: Householder QR without column pivoting instead of the column-pivoted one (sensitivity alternate)."""
    A = _f64(A); b = _f64(b); n = np.zeros(3)
    (lib().orc_lstsq5x3_alt if alt else lib().orc_lstsq5x3)(_p(A, C.c_double), _p(b, C.c_double), _p(n, C.c_double))
    return n


MAP_EDGE, MAP_SURF, DS_EDGE, DS_SURF, REGISTERED, NO_REGISTERED = range(6)


class Odometry:
    """EstimationMapping (+ featureExtraction via process_scan) on the CPU."""

    def __init__(self, cfg: Config):
        self.cfg = cfg
        self._h = C.c_void_p(lib().orc_odom_create(C.byref(cfg)))
        self.frames = 0

    def close(self):
        if self._h:
            lib().orc_odom_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_map(self, edge, surf):
        edge = _f32(edge); surf = _f32(surf)
        lib().orc_odom_init_map(self._h, _p(edge, C.c_float), edge.shape[0], _p(surf, C.c_float), surf.shape[0])

    def update(self, edge, surf):
        edge = _f32(edge); surf = _f32(surf); pose = np.zeros(7)
        lib().orc_odom_update(self._h, _p(edge, C.c_float), edge.shape[0], _p(surf, C.c_float), surf.shape[0], _p(pose, C.c_double))
        return pose

    def process_scan(self, xyzi, ring=None):
        xyzi = _f32(xyzi); pose = np.zeros(7); ne = C.c_int(); ns = C.c_int()
        rp = None
        if ring is not None:
            ring = np.ascontiguousarray(ring, dtype=np.uint16)
            rp = _p(ring, C.c_uint16)
        lib().orc_odom_process_scan(self._h, _p(xyzi, C.c_float), xyzi.shape[0], rp, int(self.frames == 0), _p(pose, C.c_double), C.byref(ne), C.byref(ns))
        self.frames += 1
        return pose, ne.value, ns.value

    def cloud(self, which: int):
        n = lib().orc_odom_cloud_size(self._h, which)
        out = np.empty((max(n, 1), 4), np.float32)
        lib().orc_odom_get_cloud(self._h, which, _p(out, C.c_float))
        return out[:n].copy()

    def set_cloud(self, which: int, pts):
        pts = _f32(pts)
        lib().orc_odom_set_cloud(self._h, which, _p(pts, C.c_float), pts.shape[0])

    def state(self):
        s = np.zeros(31)
        lib().orc_odom_get_state(self._h, _p(s, C.c_double))
        return s

    def set_state(self, s):
        s = _f64(s)
        lib().orc_odom_set_state(self._h, _p(s, C.c_double))

    def solves(self):
        out = np.zeros((8, 8))
        n = lib().orc_odom_get_solves(self._h, _p(out, C.c_double), 8)
        return out[:n].copy()

    def timing(self):
        t = np.zeros(7)
        lib().orc_odom_get_timing(self._h, _p(t, C.c_double))
        return dict(extract=t[0], scan_ds=t[1], kd_build=t[2], assoc=t[3], solve=t[4], map_update=t[5], frames=int(t[6]))


# ---------------------------------------------------------------------------------------------------------
# ScanContext (src/global_fusion/include/Scancontext/Scancontext.h)
# ---------------------------------------------------------------------------------------------------------
class SCParams(C.Structure):
    _fields_ = [("lidar_height", C.c_double), ("num_ring", C.c_int), ("num_sector", C.c_int), ("max_radius", C.c_double),
                ("num_exclude_recent", C.c_int), ("num_candidates", C.c_int), ("search_ratio", C.c_double), ("dist_thres", C.c_double),
                ("tree_making_period", C.c_int), ("pad_", C.c_int)]


def sc_params(**kw) -> SCParams:
    p = SCParams()
    lib().orc_sc_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def sc_make(pts, params: SCParams | None = None):
    """makeScancontext + ring key + sector key -> (desc [ring, sector], ringkey [ring], sectorkey [sector])."""
    p = params or sc_params()
    pts = _f32(pts)
    d = np.zeros((p.num_ring, p.num_sector)); rk = np.zeros(p.num_ring); sk = np.zeros(p.num_sector)
    lib().orc_sc_make(C.byref(p), _p(pts, C.c_float), pts.shape[0], _p(d, C.c_double), _p(rk, C.c_double), _p(sk, C.c_double))
    return d, rk, sk


def sc_distance(sc1, sc2, params: SCParams | None = None):
    """distanceBtnScanContext -> (distance, column shift)."""
    p = params or sc_params()
    a = _f64(sc1); b = _f64(sc2)
    dist = C.c_double(); sh = C.c_int()
    lib().orc_sc_distance(C.byref(p), _p(a, C.c_double), _p(b, C.c_double), C.byref(dist), C.byref(sh))
    return dist.value, sh.value


def sc_key_dist(a, b):
    a = _f32(a); b = _f32(b)
    lib().orc_sc_key_dist.restype = C.c_float
    return float(lib().orc_sc_key_dist(_p(a, C.c_float), _p(b, C.c_float), a.shape[0]))


class SCManager:
    def __init__(self, params: SCParams | None = None):
        self.p = params or sc_params()
        lib().orc_sc_create.restype = C.c_void_p
        self._h = C.c_void_p(lib().orc_sc_create(C.byref(self.p)))

    def add(self, pts):
        pts = _f32(pts)
        lib().orc_sc_add(self._h, _p(pts, C.c_float), pts.shape[0])

    def detect(self):
        """detectLoopClosureID -> (loop id or -1, yaw difference [rad], nearest distance, nearest index)."""
        lid = C.c_int(); yaw = C.c_float(); md = C.c_double(); nn = C.c_int()
        lib().orc_sc_detect(self._h, C.byref(lid), C.byref(yaw), C.byref(md), C.byref(nn))
        return lid.value, yaw.value, md.value, nn.value

    def close(self):
        if self._h:
            lib().orc_sc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ref_sc_key_knn(keys, query, k: int = 3):
    """The reference's own ring-key kd-tree (oracle/_ref/libref_sckeys.so): (indices [k], squared distances [k])."""
    if not os.path.exists(_REF_SC):
        raise RuntimeError("oracle/_ref/libref_sckeys.so not built")
    r = C.CDLL(_REF_SC)
    keys = _f32(keys); query = _f32(query)
    idx = np.zeros(k, np.int64); d2 = np.zeros(k, np.float32)
    r.ref_sc_key_knn(_p(keys, C.c_float), keys.shape[0], keys.shape[1], _p(query, C.c_float), k, _p(idx, C.c_longlong), _p(d2, C.c_float))
    return idx, d2


# ---------------------------------------------------------------------------------------------------------
# Ring-field / range-image feature extractor (src/visual_inertial_lidar/feature_tracker/include/featureExtract.hpp)
# ---------------------------------------------------------------------------------------------------------
class RIParams(C.Structure):
    _fields_ = [("n_scan", C.c_int), ("horizon_scan", C.c_int), ("downsample_rate", C.c_int), ("pad_", C.c_int),
                ("lidar_min", C.c_double), ("lidar_max", C.c_double), ("edge_threshold", C.c_double), ("surf_threshold", C.c_double)]


def ri_params(**kw) -> RIParams:
    p = RIParams()
    lib().orc_ri_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def ri_extract(params: RIParams, xyzi, ring, debug: bool = False):
    """featureExtract::extractFeature -> (edge [ne,4], edge_src [ne], surf [ns,4], surf_src [ns]) (+ dict of the intermediate arrays)."""
    xyzi = _f32(xyzi)
    ring = np.ascontiguousarray(ring, dtype=np.uint16)
    n = xyzi.shape[0]
    m = max(n, 1)
    edge = np.empty((m, 4), np.float32); surf = np.empty((m, 4), np.float32); es = np.empty(m, np.int32); ss = np.empty(m, np.int32)
    ne = C.c_int(); ns = C.c_int(); nsem = C.c_int()
    dsrc = np.zeros(m, np.int32); dcol = np.zeros(m, np.int32); drng = np.zeros(m, np.float32); dcurv = np.zeros(m, np.float32); dpick = np.zeros(m, np.int32)
    rse = np.zeros((params.n_scan, 2), np.int32)
    lib().orc_ri_extract(C.byref(params), _p(xyzi, C.c_float), _p(ring, C.c_uint16), n, _p(edge, C.c_float), _p(es, C.c_int), C.byref(ne),
                         _p(surf, C.c_float), _p(ss, C.c_int), C.byref(ns), _p(dsrc, C.c_int), _p(dcol, C.c_int), _p(drng, C.c_float), _p(dcurv, C.c_float),
                         _p(dpick, C.c_int), C.byref(nsem), _p(rse, C.c_int))
    out = (edge[:ne.value].copy(), es[:ne.value].copy(), surf[:ns.value].copy(), ss[:ns.value].copy())
    if debug:
        k = nsem.value
        return out + (dict(src=dsrc[:k].copy(), col=dcol[:k].copy(), range=drng[:k].copy(), curvature=dcurv[:k].copy(), picked=dpick[:k].copy(), ring_start_end=rse),)
    return out
