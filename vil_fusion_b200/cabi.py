"""ctypes binding of libvilf_cuda.so (include/vilf.h) — the test / bench harness side of the C ABI.

The shared library is the product; this module only marshals numpy arrays into it.  There is no CPU
fallback: importing works without a GPU (so that the symbol table can be checked), every call needs one.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VILF_LIB_PATH", os.path.join(_HERE, "libvilf_cuda.so"))  # the override is for A/B runs of kernel variants

STATUS = {0: "OK", 1: "INVALID", 2: "CUDA", 3: "CAPACITY", 4: "UNSUPPORTED", 5: "STATE"}


class VilfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vilf status {code} ({STATUS.get(code, '?')}): {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("n_scan", C.c_int32), ("n_rings", C.c_int32),
        ("lidar_min", C.c_double), ("lidar_max", C.c_double), ("edge_threshold", C.c_double),
        ("edge_leaf", C.c_double), ("surf_leaf", C.c_double), ("crop_half", C.c_double),
        ("knn_gate", C.c_double), ("huber", C.c_double),
        ("outer_iters", C.c_int32), ("lm_max_iters", C.c_int32),
        ("max_scan_points", C.c_int32), ("max_map_points", C.c_int32), ("max_ring_points", C.c_int32), ("flags", C.c_int32),
        ("horizon_scan", C.c_int32), ("downsample_rate", C.c_int32), ("ri_edge_threshold", C.c_double), ("ri_surf_threshold", C.c_double),
    ]


# every symbol include/vilf.h declares (tests/test_abi.py checks the .so exports all of them)
SYMBOLS = [
    "vilf_default_config", "vilf_create", "vilf_create_batch", "vilf_destroy", "vilf_last_error", "vilf_host_alloc", "vilf_host_free", "vilf_memcpy_h2d_async", "vilf_pack_pointcloud2", "vilf_unpack_pointcloud2", "vilf_node_outputs",
    "vilf_process_scan", "vilf_submit_scan", "vilf_wait", "vilf_submit_scan_batch", "vilf_wait_batch", "vilf_submit_scan_batch_dev",
    "vilf_feature_extract", "vilf_get_features", "vilf_map_init", "vilf_map_init_points", "vilf_update", "vilf_update_points",
    "vilf_get_pose", "vilf_set_pose", "vilf_predict", "vilf_create_submap", "vilf_get_cloud", "vilf_voxel_downsample", "vilf_crop_voxel_downsample", "vilf_crop_box", "vilf_knn5",
    "vilf_factors", "vilf_normal_equations", "vilf_solve", "vilf_get_solves", "vilf_state_export", "vilf_state_import",
    "vilf_profile_enable", "vilf_profile_read", "vilf_profile_read_kernels", "vilf_profile_kernel_name", "vilf_launch_count", "vilf_get_stream", "vilf_get_counts", "vilf_debug_voxel_phases", "vilf_bench_stage", "vilf_feature_depth",
    "vilf_sc_default_params", "vilf_sc_create", "vilf_sc_destroy", "vilf_sc_last_error", "vilf_sc_make_and_save", "vilf_sc_make_and_save_resident",
    "vilf_sc_detect_loop_closure", "vilf_sc_get", "vilf_sc_size", "vilf_sc_distance", "vilf_sc_distance_between", "vilf_sc_launch_count",
]

_lib = None


def lib():
    """Load the shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc)")
        _lib = C.CDLL(LIB_PATH)
        _lib.vilf_last_error.restype = C.c_char_p
        _lib.vilf_last_error.argtypes = [C.c_void_p]
    return _lib


def default_config(**kw) -> Config:
    c = Config()
    lib().vilf_default_config(C.byref(c))
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
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


FLAG_NO_CLUSTER = 1
FLAG_NO_GRAPH = 2
FLAG_LEGACY_MAP = 4
FLAG_CELL_MAP = 8
FLAG_RANGE_IMAGE = 16
MAP_EDGE, MAP_SURF, DS_EDGE, DS_SURF, REGISTERED, NO_REGISTERED = range(6)


class Odometry:
    """One lidar sequence: featureExtraction + EstimationMapping of the reference, on the GPU."""

    def __init__(self, cfg: Config | None = None, device: int = 0, _handle=None, _owner=True):
        self.cfg = cfg if cfg is not None else default_config()
        self._owner = _owner
        if _handle is None:
            h = C.c_void_p()
            rc = lib().vilf_create(C.byref(self.cfg), device, C.byref(h))
            if rc:
                raise VilfError(rc, "vilf_create failed (is a CUDA device visible?)")
            self._h = h
        else:
            self._h = _handle

    def _ck(self, rc):
        if rc:
            raise VilfError(rc, lib().vilf_last_error(self._h).decode())

    def close(self):
        if self._h and self._owner:
            lib().vilf_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- per-frame path ----
    def process_scan(self, xyzi, ring=None):
        xyzi = _f32(xyzi)
        pose = np.zeros(7)
        r = None if ring is None else np.ascontiguousarray(ring, dtype=np.uint16)
        self._ck(lib().vilf_process_scan(self._h, _p(xyzi, C.c_float), xyzi.shape[0], _p(r, C.c_uint16), _p(pose, C.c_double)))
        return pose

    def submit_scan(self, xyzi, ring=None) -> int:
        """xyzi / ring must stay alive (and unmodified) until wait() of the returned ticket."""
        t = C.c_int64()
        self._ck(lib().vilf_submit_scan(self._h, _p(xyzi, C.c_float), xyzi.shape[0], _p(ring, C.c_uint16), C.byref(t)))
        return t.value

    def wait(self, ticket: int):
        pose = np.zeros(7)
        self._ck(lib().vilf_wait(self._h, C.c_int64(ticket), _p(pose, C.c_double)))
        return pose

    # ---- reference method surface ----
    def feature_extract(self, xyzi, ring=None):
        xyzi = _f32(xyzi)
        r = None if ring is None else np.ascontiguousarray(ring, dtype=np.uint16)
        ne, ns = C.c_int(), C.c_int()
        self._ck(lib().vilf_feature_extract(self._h, _p(xyzi, C.c_float), xyzi.shape[0], _p(r, C.c_uint16), C.byref(ne), C.byref(ns)))
        return ne.value, ns.value

    def features(self, which: int):
        n = C.c_int()
        self._ck(lib().vilf_get_features(self._h, which, None, None, 0, C.byref(n)))
        pts = np.empty((max(n.value, 1), 4), np.float32)
        src = np.empty(max(n.value, 1), np.int32)
        self._ck(lib().vilf_get_features(self._h, which, _p(pts, C.c_float), _p(src, C.c_int32), pts.shape[0], C.byref(n)))
        return pts[:n.value].copy(), src[:n.value].copy()

    def map_init(self, edge=None, surf=None):
        if edge is None:
            self._ck(lib().vilf_map_init(self._h))
        else:
            edge, surf = _f32(edge), _f32(surf)
            self._ck(lib().vilf_map_init_points(self._h, _p(edge, C.c_float), edge.shape[0], _p(surf, C.c_float), surf.shape[0]))

    def update(self, edge=None, surf=None):
        pose = np.zeros(7)
        if edge is None:
            self._ck(lib().vilf_update(self._h, _p(pose, C.c_double)))
        else:
            edge, surf = _f32(edge), _f32(surf)
            self._ck(lib().vilf_update_points(self._h, _p(edge, C.c_float), edge.shape[0], _p(surf, C.c_float), surf.shape[0], _p(pose, C.c_double)))
        return pose

    def pose(self):
        pose = np.zeros(7)
        rt = np.zeros(12)
        self._ck(lib().vilf_get_pose(self._h, _p(pose, C.c_double), _p(rt, C.c_double)))
        return pose, rt

    def set_pose(self, pose, update_odom: bool = True):
        pose = _f64(pose)
        self._ck(lib().vilf_set_pose(self._h, _p(pose, C.c_double), int(update_odom)))

    def predict(self):
        pose = np.zeros(7)
        self._ck(lib().vilf_predict(self._h, _p(pose, C.c_double)))
        return pose

    def create_submap(self, edge_ds, surf_ds):
        e, s = _f32(edge_ds), _f32(surf_ds)
        self._ck(lib().vilf_create_submap(self._h, _p(e, C.c_float), e.shape[0], _p(s, C.c_float), s.shape[0]))

    def cloud(self, which: int):
        n = C.c_int()
        self._ck(lib().vilf_get_cloud(self._h, which, None, 0, C.byref(n)))
        out = np.empty((max(n.value, 1), 4), np.float32)
        self._ck(lib().vilf_get_cloud(self._h, which, _p(out, C.c_float), out.shape[0], C.byref(n)))
        return out[:n.value].copy()

    # ---- stage-level entry points ----
    def voxel_downsample(self, pts, leaf: float):
        pts = _f32(pts)
        out = np.empty((max(pts.shape[0], 1), 4), np.float32)
        n, g = C.c_int(), C.c_int()
        self._ck(lib().vilf_voxel_downsample(self._h, _p(pts, C.c_float), pts.shape[0], C.c_float(leaf), _p(out, C.c_float), out.shape[0], C.byref(n), C.byref(g)))
        return out[:n.value].copy(), bool(g.value)

    def crop_voxel_downsample(self, pts, center, half: float, leaf: float):
        pts = _f32(pts)
        center = _f64(center)
        out = np.empty((max(pts.shape[0], 1), 4), np.float32)
        n = C.c_int()
        self._ck(lib().vilf_crop_voxel_downsample(self._h, _p(pts, C.c_float), pts.shape[0], _p(center, C.c_double), C.c_double(half), C.c_float(leaf),
                                                  _p(out, C.c_float), out.shape[0], C.byref(n)))
        return out[:n.value].copy()

    def crop_box(self, pts, mn, mx):
        pts = _f32(pts)
        mn, mx = _f64(mn), _f64(mx)
        out = np.empty((max(pts.shape[0], 1), 4), np.float32)
        n = C.c_int()
        self._ck(lib().vilf_crop_box(self._h, _p(pts, C.c_float), pts.shape[0], _p(mn, C.c_double), _p(mx, C.c_double), _p(out, C.c_float), out.shape[0], C.byref(n)))
        return out[:n.value].copy()

    def knn5(self, mp, q):
        mp, q = _f32(mp), _f32(q)
        nq = q.shape[0]
        idx = np.empty((max(nq, 1), 5), np.int32)
        d2 = np.empty((max(nq, 1), 5), np.float32)
        self._ck(lib().vilf_knn5(self._h, _p(mp, C.c_float), mp.shape[0], _p(q, C.c_float), nq, _p(idx, C.c_int32), _p(d2, C.c_float)))
        return idx[:nq], d2[:nq]

    def factors(self, pose, edge, surf):
        pose, edge, surf = _f64(pose), _f32(edge), _f32(surf)
        ne, ns = edge.shape[0], surf.shape[0]
        ev = np.zeros(max(ne, 1), np.uint8); eab = np.zeros((max(ne, 1), 6)); enn = np.zeros((max(ne, 1), 5), np.int32); ed2 = np.zeros((max(ne, 1), 5), np.float32)
        sv = np.zeros(max(ns, 1), np.uint8); snd = np.zeros((max(ns, 1), 4)); snn = np.zeros((max(ns, 1), 5), np.int32); sd2 = np.zeros((max(ns, 1), 5), np.float32)
        self._ck(lib().vilf_factors(self._h, _p(pose, C.c_double), _p(edge, C.c_float), ne, _p(surf, C.c_float), ns,
                                    _p(ev, C.c_uint8), _p(eab, C.c_double), _p(enn, C.c_int32), _p(ed2, C.c_float),
                                    _p(sv, C.c_uint8), _p(snd, C.c_double), _p(snn, C.c_int32), _p(sd2, C.c_float)))
        return dict(edge_valid=ev[:ne], edge_ab=eab[:ne], edge_nn=enn[:ne], edge_d2=ed2[:ne],
                    surf_valid=sv[:ns], surf_nd=snd[:ns], surf_nn=snn[:ns], surf_d2=sd2[:ns])

    def normal_equations(self, pose, pab, pnd):
        pose, pab, pnd = _f64(pose), _f64(pab), _f64(pnd)
        H = np.zeros(21); g = np.zeros(6); cost = C.c_double()
        self._ck(lib().vilf_normal_equations(self._h, _p(pose, C.c_double), _p(pab, C.c_double), pab.shape[0], _p(pnd, C.c_double), pnd.shape[0],
                                             _p(H, C.c_double), _p(g, C.c_double), C.byref(cost)))
        return H, g, cost.value

    def solve(self, pose, pab, pnd, max_iters: int = 4):
        pose = _f64(pose).copy(); pab, pnd = _f64(pab), _f64(pnd)
        tr = np.zeros((10, 16)); nr = C.c_int(); term = C.c_int()
        self._ck(lib().vilf_solve(self._h, _p(pose, C.c_double), _p(pab, C.c_double), pab.shape[0], _p(pnd, C.c_double), pnd.shape[0], max_iters,
                                  _p(tr, C.c_double), tr.shape[0], C.byref(nr), C.byref(term)))
        return pose, tr[:nr.value].copy(), term.value

    def solves(self):
        out = np.zeros((8, 8)); n = C.c_int()
        self._ck(lib().vilf_get_solves(self._h, _p(out, C.c_double), 8, C.byref(n)))
        return out[:n.value].copy()

    def state(self):
        s = np.zeros(31)
        self._ck(lib().vilf_state_export(self._h, _p(s, C.c_double)))
        return s

    def set_state(self, s, map_edge, map_surf):
        s, me, ms = _f64(s), _f32(map_edge), _f32(map_surf)
        self._ck(lib().vilf_state_import(self._h, _p(s, C.c_double), _p(me, C.c_float), me.shape[0], _p(ms, C.c_float), ms.shape[0]))

    # ---- measurement ----
    def profile(self, on: bool):
        self._ck(lib().vilf_profile_enable(self._h, int(on)))

    def profile_read(self, reset: bool = True):
        ms = np.zeros(7); fr = C.c_int64()
        self._ck(lib().vilf_profile_read(self._h, _p(ms, C.c_double), C.byref(fr), int(reset)))
        names = ["extract", "scan_ds", "grid_build", "knn_fit", "solve", "map_update", "frame"]
        return dict(zip(names, ms.tolist())), fr.value

    def profile_kernels(self, reset: bool = True):
        """{(phase, kernel name): (ms, launches)} accumulated since the last reset."""
        ms = np.zeros(320); cnt = np.zeros(320, np.int64)
        self._ck(lib().vilf_profile_read_kernels(self._h, _p(ms, C.c_double), _p(cnt, C.c_int64), 320, int(reset)))
        lib().vilf_profile_kernel_name.restype = C.c_char_p
        phases = ["extract", "scan_ds", "assoc_solve", "map_update", "grid_build"]
        out = {}
        for tag in np.nonzero(cnt)[0]:
            out[(phases[tag // 64], lib().vilf_profile_kernel_name(int(tag % 64)).decode())] = (float(ms[tag]), int(cnt[tag]))
        return out

    def launch_count(self) -> int:
        n = C.c_int64()
        self._ck(lib().vilf_launch_count(self._h, C.byref(n)))
        return n.value

    def stream(self) -> int:
        p = C.c_void_p()
        self._ck(lib().vilf_get_stream(self._h, C.byref(p)))
        return p.value or 0

    def bench_stage(self, stage: int, mp, q=None, leaf: float = 0.4, iters: int = 10):
        mp = _f32(mp)
        q = None if q is None else _f32(q)
        ms = np.zeros(4)
        self._ck(lib().vilf_bench_stage(self._h, stage, _p(mp, C.c_float), mp.shape[0], _p(q, C.c_float), 0 if q is None else q.shape[0],
                                        C.c_float(leaf), iters, _p(ms, C.c_double)))
        return ms

    def feature_depth(self, feats, cloud_cam=None, T_lidar_cam=None, num_bins: int = 360):
        """(depth [m], nn [m,3], points searched).  cloud_cam=None: use the resident scan + T_lidar_cam (4x4)."""
        feats = np.ascontiguousarray(feats, dtype=np.float32)
        m = feats.shape[0]
        d = np.empty(max(m, 1), np.float32); nn = np.empty((max(m, 1), 3), np.int32); cnt = C.c_int()
        cl = None if cloud_cam is None else _f32(cloud_cam)
        T = None if T_lidar_cam is None else _f64(np.asarray(T_lidar_cam).reshape(16))
        self._ck(lib().vilf_feature_depth(self._h, _p(cl, C.c_float), 0 if cl is None else cl.shape[0], _p(T, C.c_double), _p(feats, C.c_float), m, num_bins,
                                          _p(d, C.c_float), _p(nn, C.c_int32), C.byref(cnt)))
        return d[:m].copy(), nn[:m].copy(), cnt.value

    def voxel_phases(self, job: int):
        t = np.zeros(8, np.int64)
        self._ck(lib().vilf_debug_voxel_phases(self._h, job, _p(t, C.c_int64)))
        return t

    def counts(self):
        c = np.zeros(8, np.int32)
        self._ck(lib().vilf_get_counts(self._h, _p(c, C.c_int32)))
        return dict(n_edge=int(c[0]), n_surf=int(c[1]), n_ds_edge=int(c[2]), n_ds_surf=int(c[3]), n_map_edge=int(c[4]), n_map_surf=int(c[5]),
                    status=int(c[6]), frames=int(c[7]))


class Batch:
    """`count` independent sequences stepped in lock-step on one GPU (one launch sequence for all)."""

    def __init__(self, cfg: Config | None = None, count: int = 1, device: int = 0):
        self.cfg = cfg if cfg is not None else default_config()
        self.count = count
        self._hs = (C.c_void_p * count)()
        rc = lib().vilf_create_batch(C.byref(self.cfg), device, count, self._hs)
        if rc:
            raise VilfError(rc, "vilf_create_batch failed (is a CUDA device visible?)")
        self.seqs = [Odometry(self.cfg, device, _handle=C.c_void_p(self._hs[i]), _owner=False) for i in range(count)]

    def _ck(self, rc):
        if rc:
            raise VilfError(rc, lib().vilf_last_error(C.c_void_p(self._hs[0])).decode())

    def close(self):
        if self._hs is not None:
            lib().vilf_destroy(C.c_void_p(self._hs[0]))
            self._hs = None
            for s in self.seqs:
                s._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _submit(self, fn, ptrs, ns, ring_ptrs):
        n = len(ptrs)
        arr = (C.c_void_p * n)(*ptrs)
        cnt = (C.c_int * n)(*ns)
        rarr = None if ring_ptrs is None else (C.c_void_p * n)(*ring_ptrs)
        t = C.c_int64()
        self._ck(fn(self._hs, n, arr, cnt, rarr, C.byref(t)))
        return t.value

    def submit(self, scans, rings=None) -> int:
        """scans: list of float32 [n_i,4] host arrays (kept alive by the caller until wait)."""
        ptrs = [s.ctypes.data for s in scans]
        rp = None if rings is None else [r.ctypes.data for r in rings]
        return self._submit(lib().vilf_submit_scan_batch, ptrs, [s.shape[0] for s in scans], rp)

    def submit_dev(self, dev_ptrs, ns, ring_dev_ptrs=None) -> int:
        """dev_ptrs: device addresses (ints) of float32 [n_i,4] arrays already resident in HBM."""
        return self._submit(lib().vilf_submit_scan_batch_dev, list(dev_ptrs), list(ns), None if ring_dev_ptrs is None else list(ring_dev_ptrs))

    def wait(self, ticket: int):
        poses = np.zeros((self.count, 7))
        self._ck(lib().vilf_wait_batch(self._hs, self.count, C.c_int64(ticket), _p(poses, C.c_double)))
        return poses


def host_alloc(nbytes: int) -> np.ndarray:
    """Pinned host buffer as a uint8 numpy array (freed with host_free)."""
    p = C.c_void_p()
    rc = lib().vilf_host_alloc(C.byref(p), C.c_uint64(nbytes))
    if rc:
        raise VilfError(rc, "vilf_host_alloc failed")
    buf = (C.c_uint8 * nbytes).from_address(p.value)
    a = np.frombuffer(buf, dtype=np.uint8)
    a.flags.writeable = True
    return a


def host_free(a: np.ndarray) -> None:
    lib().vilf_host_free(C.c_void_p(a.ctypes.data))


def memcpy_h2d_async(dst_dev_ptr: int, src_host_ptr: int, nbytes: int, stream: int) -> None:
    """cudaMemcpyAsync(HostToDevice) on `stream` through the CUDA runtime the library links (bench.py: link-speed probe)."""
    rc = lib().vilf_memcpy_h2d_async(C.c_void_p(dst_dev_ptr), C.c_void_p(src_host_ptr), C.c_uint64(nbytes), C.c_void_p(stream))
    if rc:
        raise VilfError(rc, "cudaMemcpyAsync failed")


def pack_pointcloud2(data: np.ndarray, n_points: int, point_step: int, off_x: int, off_y: int, off_z: int, off_intensity: int = -1, out: np.ndarray | None = None):
    """sensor_msgs/PointCloud2 payload (uint8 array) -> packed float32 [n, 4] (host side, no GPU needed)."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    if out is None:
        out = np.empty((max(n_points, 1), 4), np.float32)
    rc = lib().vilf_pack_pointcloud2(_p(data, C.c_uint8), n_points, point_step, off_x, off_y, off_z, off_intensity, _p(out, C.c_float))
    if rc:
        raise VilfError(rc, "bad PointCloud2 layout")
    return out[:n_points]


def unpack_pointcloud2(xyzi: np.ndarray, point_step: int = 32, off_x: int = 0, off_y: int = 4, off_z: int = 8, off_intensity: int = 16):
    """Packed float32 [n, 4] -> PointCloud2 data bytes; the defaults are pcl::toROSMsg's layout of pcl::PointXYZI."""
    xyzi = np.ascontiguousarray(xyzi, dtype=np.float32).reshape(-1, 4)
    n = xyzi.shape[0]
    out = np.empty(max(n, 1) * point_step, np.uint8)
    rc = lib().vilf_unpack_pointcloud2(_p(xyzi, C.c_float), n, point_step, off_x, off_y, off_z, off_intensity, _p(out, C.c_uint8))
    if rc:
        raise VilfError(rc, "bad PointCloud2 layout")
    return out[:n * point_step]


def node_outputs(rt12, last):
    """feature_tracker_node.cpp:388-401, :445-446 from get_pose()'s rt12: (relative pose [7], path pose [7], new last [7])."""
    rt12 = np.ascontiguousarray(rt12, dtype=np.float64).reshape(12)
    last = np.array(last, dtype=np.float64).reshape(7)
    rel = np.empty(7); path = np.empty(7)
    rc = lib().vilf_node_outputs(_p(rt12, C.c_double), _p(last, C.c_double), _p(rel, C.c_double), _p(path, C.c_double))
    if rc:
        raise VilfError(rc, "vilf_node_outputs")
    return rel, path, last


# ---------------------------------------------------------------------------------------------------------
# ScanContext place recognition (include/vilf.h: vilf_sc_*; SCManager of src/global_fusion/include/Scancontext/Scancontext.h)
# ---------------------------------------------------------------------------------------------------------
class SCParams(C.Structure):
    _fields_ = [("lidar_height", C.c_double), ("num_ring", C.c_int32), ("num_sector", C.c_int32), ("max_radius", C.c_double),
                ("num_exclude_recent", C.c_int32), ("num_candidates", C.c_int32), ("search_ratio", C.c_double), ("dist_thres", C.c_double),
                ("tree_making_period", C.c_int32), ("max_keyframes", C.c_int32), ("max_points", C.c_int32), ("pad_", C.c_int32)]


def sc_params(**kw) -> SCParams:
    p = SCParams()
    lib().vilf_sc_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class SCManager:
    """SCManager of the reference on the GPU: add() = makeAndSaveScancontextAndKeys, detect() = detectLoopClosureID."""

    def __init__(self, params: SCParams | None = None, device: int = 0):
        self.p = params if params is not None else sc_params()
        h = C.c_void_p()
        rc = lib().vilf_sc_create(C.byref(self.p), device, C.byref(h))
        if rc:
            raise VilfError(rc, "vilf_sc_create failed (is a CUDA device visible?)")
        self._h = h
        lib().vilf_sc_last_error.restype = C.c_char_p
        lib().vilf_sc_last_error.argtypes = [C.c_void_p]

    def _ck(self, rc):
        if rc:
            raise VilfError(rc, lib().vilf_sc_last_error(self._h).decode())

    def close(self):
        if self._h:
            lib().vilf_sc_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add(self, xyzi):
        xyzi = _f32(xyzi)
        self._ck(lib().vilf_sc_make_and_save(self._h, _p(xyzi, C.c_float), xyzi.shape[0]))

    def add_resident(self, odom: "Odometry"):
        self._ck(lib().vilf_sc_make_and_save_resident(self._h, odom._h))

    def detect(self):
        """-> (loop id or -1, yaw difference [rad], nearest distance, nearest index)"""
        lid = C.c_int(); yaw = C.c_float(); md = C.c_double(); nn = C.c_int()
        self._ck(lib().vilf_sc_detect_loop_closure(self._h, C.byref(lid), C.byref(yaw), C.byref(md), C.byref(nn)))
        return lid.value, yaw.value, md.value, nn.value

    def get(self, index: int = -1):
        d = np.zeros((self.p.num_ring, self.p.num_sector)); rk = np.zeros(self.p.num_ring); sk = np.zeros(self.p.num_sector)
        self._ck(lib().vilf_sc_get(self._h, index, _p(d, C.c_double), _p(rk, C.c_double), _p(sk, C.c_double)))
        return d, rk, sk

    def __len__(self):
        n = C.c_int()
        self._ck(lib().vilf_sc_size(self._h, C.byref(n)))
        return n.value

    def distance(self, sc1, sc2):
        a = _f64(sc1); b = _f64(sc2)
        dist = C.c_double(); sh = C.c_int()
        self._ck(lib().vilf_sc_distance(self._h, _p(a, C.c_double), _p(b, C.c_double), C.byref(dist), C.byref(sh)))
        return dist.value, sh.value

    def distance_between(self, i: int, j: int):
        dist = C.c_double(); sh = C.c_int()
        self._ck(lib().vilf_sc_distance_between(self._h, i, j, C.byref(dist), C.byref(sh)))
        return dist.value, sh.value

    def launch_count(self) -> int:
        n = C.c_int64()
        self._ck(lib().vilf_sc_launch_count(self._h, C.byref(n)))
        return n.value
