"""Synthetic HDL-64-shaped LiDAR scenes shared by tests, smoke() and bench.py.

Not part of the reference: the reference ships no data (README.md:92-95 points at a KITTI rosbag).
Frames are float32 XYZI, stride 16 B, ring-major (ring 0 = highest elevation), deterministic.

A frame = 64 rings x 1875 azimuth steps = 120,000 rays from a sensor at the origin:
  * ground plane z = -1.73 m (+ gaussian range noise),
  * K box obstacles (cars 4.5x1.8x1.5 m, pedestrians 0.6x0.6x1.7 m) ray-cast with occlusion,
  * rays that hit nothing end on a 6 m high wall ring at r = 45 m.
Objects move (static / constant-velocity with reflection at the ROI edge / constant-turn-rate) with
dt = 0.1 s per frame, so the tracker sees persistent, associable detections.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SENSOR_Z_GROUND = -1.73
WALL_R = 45.0
WALL_TOP = SENSOR_Z_GROUND + 6.0
DT_US = 100_000.0


@dataclass
class SceneConfig:
    rings: int = 64
    azimuths: int = 1875
    elev_top_deg: float = 2.0
    elev_bot_deg: float = -24.8
    n_objects: int = 72
    ped_fraction: float = 0.25
    lattice_pitch: float = 5.5
    roi_half: float = 23.0
    range_noise: float = 0.03
    seed: int = 1

    @property
    def n_points(self) -> int:
        return self.rings * self.azimuths


def dense_config(**kw) -> SceneConfig:
    """1M-point 'dense' frame of BASELINE.json configs[4]: 128 rings x 7813 azimuths (truncated to 1e6)."""
    return SceneConfig(rings=128, azimuths=7813, **kw)


@dataclass
class Scene:
    cfg: SceneConfig
    pos: np.ndarray = field(default=None)      # (K,2)
    yaw: np.ndarray = field(default=None)      # (K,)
    size: np.ndarray = field(default=None)     # (K,3) length,width,height
    vel: np.ndarray = field(default=None)      # (K,) speed m/s
    yawrate: np.ndarray = field(default=None)  # (K,) rad/s
    frame: int = 0


def make_scene(cfg: SceneConfig) -> Scene:
    rng = np.random.default_rng(cfg.seed)
    p = cfg.lattice_pitch
    n_side = int(2 * cfg.roi_half // p) + 1
    coords = (np.arange(n_side) - (n_side - 1) / 2.0) * p
    gx, gy = np.meshgrid(coords, coords, indexing="ij")
    sites = np.stack([gx.ravel(), gy.ravel()], 1)
    sites = sites[np.hypot(sites[:, 0], sites[:, 1]) > 6.0]
    rng.shuffle(sites)
    k = min(cfg.n_objects, len(sites))
    pos = sites[:k] + rng.uniform(-0.6, 0.6, size=(k, 2))
    is_ped = rng.random(k) < cfg.ped_fraction
    size = np.where(is_ped[:, None], np.array([0.6, 0.6, 1.7]), np.array([4.5, 1.8, 1.5]))
    yaw = rng.uniform(-math.pi, math.pi, size=k)
    kind = rng.integers(0, 3, size=k)  # 0 static, 1 CV, 2 CTRV
    vel = np.where(kind == 0, 0.0, rng.uniform(0.3, 2.5, size=k))
    vel = np.where(is_ped & (kind != 0), rng.uniform(0.3, 1.5, size=k), vel)
    yawrate = np.where(kind == 2, rng.uniform(-0.6, 0.6, size=k), 0.0)
    return Scene(cfg, pos.astype(np.float64), yaw, size.astype(np.float64), vel, yawrate, 0)


def advance(scene: Scene, dt: float = 0.1) -> None:
    """Move every object one frame (CV / CTRV), reflecting headings at the ROI edge."""
    s = scene
    s.yaw = s.yaw + s.yawrate * dt
    s.pos = s.pos + (s.vel * dt)[:, None] * np.stack([np.cos(s.yaw), np.sin(s.yaw)], 1)
    lim = s.cfg.roi_half
    out_x = np.abs(s.pos[:, 0]) > lim
    out_y = np.abs(s.pos[:, 1]) > lim
    s.yaw = np.where(out_x, math.pi - s.yaw, s.yaw)
    s.yaw = np.where(out_y, -s.yaw, s.yaw)
    s.pos = np.clip(s.pos, -lim, lim)
    # keep objects out of the sensor's blind disc
    r = np.hypot(s.pos[:, 0], s.pos[:, 1])
    near = r < 5.0
    s.yaw = np.where(near, np.arctan2(s.pos[:, 1], s.pos[:, 0]), s.yaw)
    s.yaw = (s.yaw + math.pi) % (2 * math.pi) - math.pi
    s.frame += 1


def render(scene: Scene) -> np.ndarray:
    """Ray-cast the scene -> (N,4) float32 XYZI, ring-major."""
    cfg = scene.cfg
    rng = np.random.default_rng((cfg.seed << 20) + scene.frame + 1)
    R, A = cfg.rings, cfg.azimuths
    elev = np.deg2rad(np.linspace(cfg.elev_top_deg, cfg.elev_bot_deg, R))
    az = (np.arange(A) + 0.5) * (2 * math.pi / A) - math.pi
    ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
    dx = ce * np.cos(az)[None, :]
    dy = ce * np.sin(az)[None, :]
    dz = np.broadcast_to(se, (R, A))
    # ground / wall
    with np.errstate(divide="ignore"):
        t_ground = np.where(dz < -1e-9, SENSOR_Z_GROUND / dz, np.inf)
    t_wall = WALL_R / np.maximum(np.hypot(dx, dy), 1e-9)
    t = np.minimum(t_ground, t_wall)
    hit_kind = np.where(t_ground <= t_wall, 0, 1).astype(np.int8)
    inten = np.where(hit_kind == 0, 0.2, 0.5)
    # boxes: slab test on the azimuth columns each box can cover
    for k in range(len(scene.pos)):
        cx, cy = scene.pos[k]
        L, W, H = scene.size[k]
        c, s = math.cos(scene.yaw[k]), math.sin(scene.yaw[k])
        corners = np.array([[+L / 2, +W / 2], [+L / 2, -W / 2], [-L / 2, -W / 2], [-L / 2, +W / 2]])
        wc = np.stack([cx + c * corners[:, 0] - s * corners[:, 1], cy + s * corners[:, 0] + c * corners[:, 1]], 1)
        ca = np.arctan2(wc[:, 1], wc[:, 0])
        ref = math.atan2(cy, cx)
        d = (ca - ref + math.pi) % (2 * math.pi) - math.pi
        a_lo, a_hi = ref + d.min(), ref + d.max()
        i_lo = int(math.floor((a_lo + math.pi) / (2 * math.pi) * A)) - 1
        i_hi = int(math.ceil((a_hi + math.pi) / (2 * math.pi) * A)) + 1
        cols = np.arange(i_lo, i_hi + 1) % A
        # ray in box frame
        ox, oy = -cx, -cy
        obx, oby = c * ox + s * oy, -s * ox + c * oy
        ddx, ddy, ddz = dx[:, cols], dy[:, cols], dz[:, cols]
        bdx, bdy = c * ddx + s * ddy, -s * ddx + c * ddy
        zlo, zhi = SENSOR_Z_GROUND, SENSOR_Z_GROUND + H
        tmin = np.full(bdx.shape, -np.inf)
        tmax = np.full(bdx.shape, np.inf)
        for o, dd, lo, hi in ((obx, bdx, -L / 2, L / 2), (oby, bdy, -W / 2, W / 2), (0.0, ddz, zlo, zhi)):
            with np.errstate(divide="ignore", invalid="ignore"):
                inv = 1.0 / np.where(np.abs(dd) < 1e-12, 1e-12, dd)
            t1, t2 = (lo - o) * inv, (hi - o) * inv
            tmin = np.maximum(tmin, np.minimum(t1, t2))
            tmax = np.minimum(tmax, np.maximum(t1, t2))
        hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 0.0)
        tc = t[:, cols]
        better = hit & (tmin < tc)
        tc = np.where(better, tmin, tc)
        t[:, cols] = tc
        ic = inten[:, cols]
        inten[:, cols] = np.where(better, 0.8, ic)
    t = t + rng.normal(0.0, cfg.range_noise, size=t.shape)
    pts = np.stack([dx * t, dy * t, dz * t, inten], -1).reshape(-1, 4).astype(np.float32)
    return pts


def frames(cfg: SceneConfig, n_frames: int, only=None):
    """Yield (timestamp_us, points) for n_frames consecutive frames of one scene.  `only(f)`: render just the frames it accepts
    (the scene still advances through all of them) -- a rank that holds every world-th frame of a frame-sharded sequence."""
    sc = make_scene(cfg)
    for f in range(n_frames):
        if only is None or only(f):
            yield (f + 1) * DT_US, render(sc)
        advance(sc)


def uniform_cloud(n: int, seed: int, half: float = 60.0) -> np.ndarray:
    """Uniform random XYZI cloud (stress / ragged-input tests)."""
    rng = np.random.default_rng(seed)
    p = np.empty((n, 4), np.float32)
    p[:, 0:2] = rng.uniform(-half, half, size=(n, 2))
    p[:, 2] = rng.uniform(-3.0, 1.5, size=n)
    p[:, 3] = rng.random(n)
    return p
