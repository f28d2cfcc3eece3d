"""The C-ABI shared library: loads, exports every symbol include/vilf.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vilf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vilf_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(cabi):
    assert header_symbols() == sorted(cabi.SYMBOLS)


def test_library_exports_every_declared_symbol(cabi):
    lib = cabi.lib()
    for s in header_symbols():
        assert hasattr(lib, s), f"libvilf_cuda.so does not export {s}"


def test_default_config_is_the_reference_yaml(cabi):
    c = cabi.default_config()
    # config/kitti/velodyne_param_64.yaml:9-23 + hard-coded constants
    assert (c.n_scan, c.lidar_min, c.lidar_max, c.edge_threshold) == (64, 3.0, 90.0, 0.1)
    assert (c.edge_leaf, c.surf_leaf, c.crop_half, c.knn_gate, c.huber) == (0.4, 0.8, 100.0, 1.0, 0.1)
    assert (c.outer_iters, c.lm_max_iters) == (2, 4)


def test_config_struct_layout_matches_header(cabi, tmp_path):
    """The ctypes mirrors of vilf_config / vilf_sc_params have the size and field offsets the C compiler gives include/vilf.h."""
    import subprocess
    src = tmp_path / "layout.c"
    fields_cfg = [f[0] for f in cabi.Config._fields_]
    fields_sc = [f[0] for f in cabi.SCParams._fields_]
    body = "".join(f'printf("%zu ", offsetof(vilf_config, {f}));' for f in fields_cfg) + 'printf("%zu\\n", sizeof(vilf_config));'
    body += "".join(f'printf("%zu ", offsetof(vilf_sc_params, {f}));' for f in fields_sc) + 'printf("%zu\\n", sizeof(vilf_sc_params));'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vilf.h"\nint main(void) {' + body + "return 0; }\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert [int(v) for v in out[0].split()] == [getattr(cabi.Config, f).offset for f in fields_cfg] + [C.sizeof(cabi.Config)]
    assert [int(v) for v in out[1].split()] == [getattr(cabi.SCParams, f).offset for f in fields_sc] + [C.sizeof(cabi.SCParams)]
    assert C.sizeof(cabi.Config) == 120  # 2 x int32, 8 x double, 6 x int32, 2 x int32, 2 x double


def test_invalid_arguments_are_rejected_without_touching_cuda(cabi):
    lib = cabi.lib()
    h = C.c_void_p()
    cfg = cabi.default_config()
    assert lib.vilf_create(None, 0, C.byref(h)) == 1
    cfg.outer_iters = 0
    assert lib.vilf_create(C.byref(cfg), 0, C.byref(h)) == 1
    assert lib.vilf_process_scan(None, None, 0, None, None) == 1
    assert lib.vilf_destroy(None) == 1


def test_no_cpu_fallback(cabi):
    """Without a CUDA device the product path must refuse to run (no oracle / CPU route behind the ABI)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cabi.VilfError) as e:
        cabi.Odometry(cabi.default_config())
    assert e.value.code == 2


def test_product_package_does_not_import_the_oracle():
    """Nothing under vil_fusion_b200/ (the product) may import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "vil_fusion_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "liborc" not in txt and "orc_capi" not in txt and "oracle/" not in txt, f


def test_pack_pointcloud2(cabi):
    """PointCloud2 -> packed float4 (feature_tracker_node.cpp:339-340, pcl::fromROSMsg): a velodyne-style layout
    x y z intensity ring(u16) time(f32), point_step 22, and a layout without intensity."""
    import numpy as np
    rng = np.random.default_rng(1)
    n = 1000
    dt = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"], "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                   "offsets": [0, 4, 8, 12, 16, 18], "itemsize": 22})
    msg = np.zeros(n, dt)
    for k in ("x", "y", "z", "intensity", "time"):
        msg[k] = rng.normal(size=n).astype(np.float32)
    msg["ring"] = rng.integers(0, 64, n)
    raw = msg.view(np.uint8).reshape(-1)
    out = cabi.pack_pointcloud2(raw, n, 22, 0, 4, 8, 12)
    assert np.array_equal(out, np.stack([msg["x"], msg["y"], msg["z"], msg["intensity"]], 1))
    out = cabi.pack_pointcloud2(raw, n, 22, 0, 4, 8, -1)
    assert np.array_equal(out[:, :3], np.stack([msg["x"], msg["y"], msg["z"]], 1)) and (out[:, 3] == 0).all()
    with pytest.raises(cabi.VilfError):
        cabi.pack_pointcloud2(raw, n, 22, 0, 4, 20, 12)  # z would read past the point


def test_unpack_pointcloud2_round_trip(cabi):
    """packed float4 -> PointCloud2 bytes in pcl::toROSMsg's PointXYZI layout (feature_tracker_node.cpp:439-440) and back."""
    import numpy as np
    rng = np.random.default_rng(2)
    pts = rng.normal(size=(777, 4)).astype(np.float32)
    raw = cabi.unpack_pointcloud2(pts)
    assert raw.size == 777 * 32
    rec = raw.view(np.dtype({"names": ["x", "y", "z", "intensity"], "formats": ["<f4"] * 4, "offsets": [0, 4, 8, 16], "itemsize": 32}))
    assert np.array_equal(np.stack([rec["x"], rec["y"], rec["z"], rec["intensity"]], 1), pts)
    pad = raw.reshape(-1, 32)
    assert not pad[:, 12:16].any() and not pad[:, 20:].any()
    assert np.array_equal(cabi.pack_pointcloud2(raw, 777, 32, 0, 4, 8, 16), pts)
    assert cabi.unpack_pointcloud2(np.zeros((0, 4), np.float32)).size == 0
    with pytest.raises(cabi.VilfError):
        cabi.unpack_pointcloud2(pts, 16, 0, 4, 8, 16)  # intensity would land past the point


def test_node_outputs_match_the_oracle_and_scipy(cabi):
    """/Odometry relative pose and /path pose (feature_tracker_node.cpp:388-401): the library's host arithmetic against the
    oracle bit for bit, and against scipy's rotation algebra to rounding, over a random walk that includes rotations
    beyond 120 degrees (the non-trace branches of the matrix -> quaternion conversion)."""
    import numpy as np
    from scipy.spatial.transform import Rotation as Rot
    from oracle import orc
    rng = np.random.default_rng(3)
    last_a = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)
    last_b = last_a.copy()
    Rprev, tprev = np.eye(3), np.zeros(3)
    R, t = np.eye(3), np.zeros(3)
    branches = set()
    for k in range(400):
        R = R @ Rot.from_rotvec(rng.normal(size=3) * (0.02 if k % 7 else 1.5)).as_matrix()
        t = t + rng.normal(size=3)
        rt12 = np.concatenate([R.reshape(-1), t])
        rel_a, path_a, last_a = cabi.node_outputs(rt12, last_a)
        rel_b, path_b, last_b = orc.node_outputs(rt12, last_b)
        assert np.array_equal(rel_a, rel_b) and np.array_equal(path_a, path_b) and np.array_equal(last_a, last_b)
        branches.add(bool(np.trace(R) > 0))
        q = Rot.from_matrix(R).as_quat()
        assert min(np.abs(path_a[:4] - q).max(), np.abs(path_a[:4] + q).max()) < 1e-12 and np.array_equal(path_a[4:], t)
        qr = Rot.from_matrix(Rprev.T @ R).as_quat()
        assert min(np.abs(rel_a[:4] - qr).max(), np.abs(rel_a[:4] + qr).max()) < 1e-12
        assert np.abs(rel_a[4:] - Rprev.T @ (t - tprev)).max() < 1e-11
        Rprev, tprev = R.copy(), t.copy()
    assert branches == {True, False}
