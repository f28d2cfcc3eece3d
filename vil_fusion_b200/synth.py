"""Seeded synthetic lidar sequences for the odometry hot path (test / bench input only).

The reference ships no data, so every input is generated here:
an analytic ray cast of a ground plane plus axis-aligned boxes (buildings and
poles) seen from a smooth vehicle trajectory, with spinning-lidar elevation
tables that invert the reference's own ring formulas
(featureExtraction.hpp:75-102), so that `getLaserCloud` bins every return into
the ring it was fired from.

Output layout per scan: float32 [n, 4] = x, y, z, intensity in the SENSOR
frame, ring-major and azimuth-ascending inside a ring (the reference assumes
arrival order == azimuth order inside a ring, featureExtraction.hpp:108), plus
the uint16 ring id of every return (needed for beam counts the reference's
angle formulas do not cover, featureExtraction.hpp:103-106).

numpy only: runs identically in the CPU container and on the GPU box.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

GROUND_Z = -1.73  # KITTI mount height


@dataclass(frozen=True)
class Sensor:
    name: str
    elev_deg: np.ndarray  # [rings], ring id == index
    n_az: int
    n_scan_param: int  # value for the reference's /N_SCAN parameter (0 = explicit ring ids)

    @property
    def rings(self) -> int:
        return int(self.elev_deg.shape[0])


def hdl64() -> Sensor:
    """HDL-64E-like table inverting featureExtraction.hpp:93-96.

    ring r<32: 2 - r/3 deg, ring r>=32: -8.83 - (r-32)/2 deg.  The first and
    last beams sit exactly on the reference's reject limits (angle > 2,
    angle < -24.33, featureExtraction.hpp:98); they are pulled 0.05 deg
    inwards so that an ulp of atan() cannot drop a whole ring.
    """
    r = np.arange(64, dtype=np.float64)
    e = np.where(r < 32, 2.0 - r / 3.0, -8.83 - (r - 32.0) / 2.0)
    e[0] -= 0.05
    e[63] += 0.05
    return Sensor("hdl64", e, 1800, 64)


def vlp32() -> Sensor:
    """32-beam table hitting the reference's uniform bins mid-bin.

    featureExtraction.hpp:85: scanID = int((angle + 92/3) * 3/4).
    """
    r = np.arange(32, dtype=np.float64)
    e = -92.0 / 3.0 + (r + 0.5) * 4.0 / 3.0
    return Sensor("vlp32", e, 1800, 32)


def beams128() -> Sensor:
    """128-beam sensor; outside the reference's angle formulas -> explicit ring ids."""
    e = np.linspace(12.0, -25.0, 128)
    return Sensor("beams128", e, 2048, 0)


SENSORS = {"hdl64": hdl64, "vlp32": vlp32, "beams128": beams128}


@dataclass
class World:
    boxes: np.ndarray  # [B, 6] xmin ymin zmin xmax ymax zmax
    ground_z: float = GROUND_Z


def make_world(seed: int = 0, length: float = 2200.0, density: float = 1.0) -> World:
    """Corridor world along +x: buildings on both sides, poles nearer the lane.

    `density` scales the number of objects (config 3 / 5 use > 1 to grow the
    local map).
    """
    rng = np.random.default_rng(1000 + seed)
    boxes = []
    for side in (-1.0, 1.0):
        x = -150.0
        while x < length:
            w = rng.uniform(10.0, 30.0)
            d = rng.uniform(10.0, 30.0)
            h = rng.uniform(6.0, 25.0)
            y0 = rng.uniform(12.0, 22.0)
            lo, hi = (y0, y0 + d) if side > 0 else (-y0 - d, -y0)
            boxes.append([x, lo, GROUND_Z, x + w, hi, GROUND_Z + h])
            x += w + rng.uniform(2.0, 14.0) / density
        # poles
        x = -150.0
        while x < length:
            y = side * rng.uniform(8.5, 10.5)
            s = 0.15
            h = rng.uniform(3.0, 8.0)
            boxes.append([x - s, y - s, GROUND_Z, x + s, y + s, GROUND_Z + h])
            x += rng.uniform(7.0, 25.0) / density
        # low walls / parked-vehicle-like boxes
        x = -150.0
        while x < length:
            y = side * rng.uniform(6.5, 8.0)
            lx = rng.uniform(2.0, 5.0)
            boxes.append([x, y - 0.9, GROUND_Z, x + lx, y + 0.9, GROUND_Z + rng.uniform(1.2, 2.2)])
            x += rng.uniform(15.0, 60.0) / density
    return World(np.asarray(boxes, dtype=np.float64))


def _rot(roll: float, pitch: float, yaw: float) -> np.ndarray:
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cy, sy = np.cos(yaw), np.sin(yaw)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


def trajectory(n_frames: int, seed: int = 0, speed: float = 1.0):
    """Ground-truth sensor poses (R [n,3,3], t [n,3]); frame 0 is the identity.

    ~`speed` m/frame forward, gentle yaw weave (rate well under 0.02 rad/frame),
    small sinusoidal roll/pitch.  Kept for sanity checks only: parity is GPU
    vs oracle, not vs ground truth.
    """
    rng = np.random.default_rng(2000 + seed)
    ph = rng.uniform(0, 2 * np.pi, 3)
    f = np.arange(n_frames, dtype=np.float64)
    yaw = 0.08 * (np.sin(2 * np.pi * f / 120.0 + ph[0]) - np.sin(ph[0]))
    roll = 0.01 * (np.sin(2 * np.pi * f / 37.0 + ph[1]) - np.sin(ph[1]))
    pitch = 0.008 * (np.sin(2 * np.pi * f / 53.0 + ph[2]) - np.sin(ph[2]))
    t = np.zeros((n_frames, 3))
    for i in range(1, n_frames):
        t[i, 0] = t[i - 1, 0] + speed * np.cos(yaw[i - 1])
        t[i, 1] = t[i - 1, 1] + speed * np.sin(yaw[i - 1])
    R = np.stack([_rot(roll[i], pitch[i], yaw[i]) for i in range(n_frames)])
    return R, t


def _ray_dirs(sensor: Sensor) -> np.ndarray:
    e = np.deg2rad(sensor.elev_deg)[:, None]
    a = (-np.pi + (np.arange(sensor.n_az) + 0.5) * (2 * np.pi / sensor.n_az))[None, :]
    d = np.stack([np.cos(e) * np.cos(a), np.cos(e) * np.sin(a), np.sin(e) * np.ones_like(a)], axis=-1)
    return d  # [rings, n_az, 3]


def _cast_numpy(d, t, near, ground_z):
    """Nearest positive hit distance per ray (inf = no hit): ground plane + boxes (slab test)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        s = (ground_z - t[2]) / d[:, 2]
    s = np.where((d[:, 2] < -1e-9) & (s > 0), s, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
    for j in range(0, near.shape[0], 16):
        bb = near[j:j + 16]
        with np.errstate(invalid="ignore"):
            lo = (bb[None, :, 0:3] - t[None, None, :]) * inv[:, None, :]
            hi = (bb[None, :, 3:6] - t[None, None, :]) * inv[:, None, :]
        tmin = np.fmin(lo, hi).max(axis=2)
        tmax = np.fmax(lo, hi).min(axis=2)
        hit = (tmax >= np.maximum(tmin, 0.0))
        sb = np.where(hit, np.where(tmin > 0, tmin, np.inf), np.inf).min(axis=1)
        s = np.minimum(s, sb)
    return s


try:  # same arithmetic, ~100x faster; numpy path kept as the fallback
    import numba as _nb

    @_nb.njit(parallel=True, cache=False)
    def _cast_numba(d, t, near, ground_z):
        n = d.shape[0]
        nb = near.shape[0]
        out = np.empty(n, dtype=np.float64)
        for i in _nb.prange(n):
            best = np.inf
            dz = d[i, 2]
            if dz < -1e-9:
                sg = (ground_z - t[2]) / dz
                if sg > 0:
                    best = sg
            i0 = 1.0 / d[i, 0]
            i1 = 1.0 / d[i, 1]
            i2 = 1.0 / d[i, 2]
            for j in range(nb):
                lo = (near[j, 0] - t[0]) * i0
                hi = (near[j, 3] - t[0]) * i0
                tmin = min(lo, hi)
                tmax = max(lo, hi)
                lo = (near[j, 1] - t[1]) * i1
                hi = (near[j, 4] - t[1]) * i1
                tmin = max(tmin, min(lo, hi))
                tmax = min(tmax, max(lo, hi))
                lo = (near[j, 2] - t[2]) * i2
                hi = (near[j, 5] - t[2]) * i2
                tmin = max(tmin, min(lo, hi))
                tmax = min(tmax, max(lo, hi))
                if tmax >= tmin and tmin > 0.0 and tmin < best:
                    best = tmin
            out[i] = best
        return out

    _cast = _cast_numba
except Exception:  # pragma: no cover
    _cast = _cast_numpy


def raycast(world: World, sensor: Sensor, R: np.ndarray, t: np.ndarray, rng: np.random.Generator,
            noise: float = 0.01, max_range: float = 120.0):
    """One scan: returns (xyzi float32 [n,4], ring uint16 [n])."""
    dirs_s = _ray_dirs(sensor)  # sensor frame
    rings, n_az, _ = dirs_s.shape
    d = dirs_s.reshape(-1, 3) @ R.T  # world-frame directions
    # cull boxes that cannot be hit inside max_range
    b = world.boxes
    cx = np.clip(t[0], b[:, 0], b[:, 3]) - t[0]
    cy = np.clip(t[1], b[:, 1], b[:, 4]) - t[1]
    near = b[cx * cx + cy * cy < max_range * max_range]
    s = _cast(np.ascontiguousarray(d), np.ascontiguousarray(t, dtype=np.float64),
              np.ascontiguousarray(near), float(world.ground_z))
    ok = np.isfinite(s) & (s < max_range)
    s = s + rng.normal(0.0, noise, s.shape)
    with np.errstate(invalid="ignore"):
        p = dirs_s.reshape(-1, 3) * s[:, None]
    inten = rng.uniform(0.0, 1.0, s.shape)
    ring = np.repeat(np.arange(rings, dtype=np.uint16), n_az)
    ok &= s > 0.5
    xyzi = np.concatenate([p, inten[:, None]], axis=1)[ok].astype(np.float32)
    return np.ascontiguousarray(xyzi), np.ascontiguousarray(ring[ok])


class Sequence:
    """Lazy seeded sequence of scans: seq[i] -> (xyzi, ring); deterministic per (seed, i)."""

    def __init__(self, sensor: str = "hdl64", n_frames: int = 100, seed: int = 0,
                 density: float = 1.0, noise: float = 0.01, speed: float = 1.0):
        self.sensor = SENSORS[sensor]()
        self.n_frames = n_frames
        self.seed = seed
        self.noise = noise
        self.world = make_world(seed, length=max(400.0, n_frames * speed + 300.0), density=density)
        self.R, self.t = trajectory(n_frames, seed, speed)

    def __len__(self) -> int:
        return self.n_frames

    def __getitem__(self, i: int):
        rng = np.random.default_rng([self.seed, 77, i])
        return raycast(self.world, self.sensor, self.R[i], self.t[i], rng, noise=self.noise)

    def gt_pose(self, i: int):
        return self.R[i], self.t[i]
