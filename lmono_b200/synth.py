"""Deterministic synthetic LiDAR data for the lmono hot path (numpy only).

SURVEY.md section 8(d) defines the workloads: a procedural street scene (ground plane,
axis-aligned boxes = buildings, vertical poles) along a loop road, an HDL-64/HDL-32 shaped
ray-caster whose beam elevations sit at the centres of the reference's ring bins
(inverse of Aloam/src/scanRegistration.cpp:178-192), and direct samplers for map points
and per-sweep feature clouds.  Output layout is the KITTI ``.bin`` layout the reference's
kittiHelper reads (Aloam/src/kittiHelper.cpp:25-35): float32 x, y, z, intensity.

Nothing here touches the oracle or the GPU; it only makes inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SENSOR_HEIGHT = 1.73


# --------------------------------------------------------------------------- world
@dataclass
class World:
    boxes: np.ndarray          # [nb, 6] xmin, ymin, zmin, xmax, ymax, zmax
    poles: np.ndarray          # [np, 4] x, y, radius, height
    loop_radius: float
    seed: int
    ground_z: float = 0.0
    extra: dict = field(default_factory=dict)


def make_world(seed: int = 20261017, n_boxes: int = 200, n_poles: int = 300,
               loop_length: float = 2000.0) -> World:
    """Ground plane + boxes + poles along a circular loop road of the given length."""
    rng = np.random.default_rng(seed)
    R = loop_length / (2.0 * math.pi)
    boxes = []
    for _ in range(n_boxes):
        ang = rng.uniform(0, 2 * math.pi)
        side = rng.choice([-1.0, 1.0])
        off = rng.uniform(9.0, 30.0) * side
        cx, cy = (R + off) * math.cos(ang), (R + off) * math.sin(ang)
        sx, sy = rng.uniform(6.0, 30.0, size=2)
        h = rng.uniform(4.0, 18.0)
        boxes.append([cx - sx / 2, cy - sy / 2, 0.0, cx + sx / 2, cy + sy / 2, h])
    poles = []
    for _ in range(n_poles):
        ang = rng.uniform(0, 2 * math.pi)
        side = rng.choice([-1.0, 1.0])
        off = rng.uniform(4.5, 8.0) * side
        poles.append([(R + off) * math.cos(ang), (R + off) * math.sin(ang), 0.15, rng.uniform(4.0, 9.0)])
    return World(np.asarray(boxes, np.float64), np.asarray(poles, np.float64), R, seed)


def loop_pose(world: World, s: float):
    """Pose (q xyzw, t) of the sensor after driving arc length s along the loop (z up, x forward)."""
    R = world.loop_radius
    ang = s / R
    t = np.array([R * math.cos(ang), R * math.sin(ang), SENSOR_HEIGHT])
    yaw = ang + math.pi / 2.0
    q = np.array([0.0, 0.0, math.sin(yaw / 2), math.cos(yaw / 2)])
    return q, t


def quat_to_rot(q):
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def small_rot_quat(rx, ry, rz):
    """Quaternion for a small rotation vector (rad)."""
    v = np.array([rx, ry, rz], np.float64)
    n = np.linalg.norm(v)
    if n < 1e-15:
        return np.array([0, 0, 0, 1.0])
    s = math.sin(n / 2) / n
    return np.array([v[0] * s, v[1] * s, v[2] * s, math.cos(n / 2)])


def perturb_pose(q, t, rng, dt=0.2, drot_deg=1.0):
    """U(+-dt m, +-drot deg) perturbation."""
    dq = small_rot_quat(*np.deg2rad(rng.uniform(-drot_deg, drot_deg, 3)))
    return quat_mul(q, dq), t + rng.uniform(-dt, dt, 3)


# --------------------------------------------------------------------------- direct samplers
def _nearby(world: World, center, radius):
    c = np.asarray(center[:2])
    bx = world.boxes
    bc = 0.5 * (bx[:, 0:2] + bx[:, 3:5])
    bh = 0.5 * np.linalg.norm(bx[:, 3:5] - bx[:, 0:2], axis=1)
    bsel = np.linalg.norm(bc - c, axis=1) < radius + bh
    psel = np.linalg.norm(world.poles[:, :2] - c, axis=1) < radius
    return world.boxes[bsel], world.poles[psel]


def sample_surface_points(world: World, center, radius, n, rng, noise=0.0):
    """n points on the ground and on box walls within `radius` of center (world frame)."""
    boxes, _ = _nearby(world, center, radius)
    n_ground = n // 2 if len(boxes) else n
    out = []
    r = radius * np.sqrt(rng.uniform(0, 1, n_ground))
    a = rng.uniform(0, 2 * math.pi, n_ground)
    g = np.stack([center[0] + r * np.cos(a), center[1] + r * np.sin(a), np.full(n_ground, world.ground_z)], 1)
    out.append(g)
    n_wall = n - n_ground
    if n_wall > 0 and len(boxes):
        bi = rng.integers(0, len(boxes), n_wall)
        b = boxes[bi]
        face = rng.integers(0, 4, n_wall)
        u = rng.uniform(0, 1, n_wall)
        v = rng.uniform(0, 1, n_wall)
        x = np.where(face == 0, b[:, 0], np.where(face == 1, b[:, 3], b[:, 0] + u * (b[:, 3] - b[:, 0])))
        y = np.where(face == 2, b[:, 1], np.where(face == 3, b[:, 4], b[:, 1] + u * (b[:, 4] - b[:, 1])))
        z = b[:, 2] + v * (b[:, 5] - b[:, 2])
        w = np.stack([x, y, z], 1)
        keep = np.linalg.norm(w[:, :2] - np.asarray(center[:2]), axis=1) < radius
        out.append(w[keep])
    p = np.concatenate(out, 0)
    if noise > 0:
        p = p + rng.normal(0, noise, p.shape)
    return p


def sample_edge_points(world: World, center, radius, n, rng, noise=0.0):
    """n points on pole axes and vertical box edges within `radius` of center."""
    boxes, poles = _nearby(world, center, radius)
    segs = []
    for p in poles:
        segs.append([p[0], p[1], 0.0, p[3]])
    for b in boxes:
        for (x, y) in ((b[0], b[1]), (b[0], b[4]), (b[3], b[1]), (b[3], b[4])):
            if math.hypot(x - center[0], y - center[1]) < radius:
                segs.append([x, y, b[2], b[5]])
    if not segs:
        return np.zeros((0, 3))
    segs = np.asarray(segs)
    si = rng.integers(0, len(segs), n)
    s = segs[si]
    z = s[:, 2] + rng.uniform(0, 1, n) * (s[:, 3] - s[:, 2])
    p = np.stack([s[:, 0], s[:, 1], z], 1)
    if noise > 0:
        p = p + rng.normal(0, noise, p.shape)
    return p


def to_xyzi(p, intensity=0.0):
    out = np.zeros((len(p), 4), np.float32)
    out[:, :3] = p
    out[:, 3] = intensity
    return out


def world_to_sensor(p_world, q, t):
    R = quat_to_rot(q)
    return (p_world - t) @ R     # R^T (p - t)


def sample_sweep_features(world: World, q, t, rng, n_corner=5000, n_surf=40000, max_range=60.0,
                          noise=0.02):
    """Feature clouds of one sweep in the SENSOR frame (what laserOdometry hands to
    laserMapping as /laser_cloud_corner_last and /laser_cloud_surf_last)."""
    c = sample_edge_points(world, t, max_range, n_corner, rng, noise)
    s = sample_surface_points(world, t, max_range, n_surf, rng, noise)
    c = c[c[:, 2] < t[2] + 12.0]
    return (to_xyzi(world_to_sensor(c, q, t)), to_xyzi(world_to_sensor(s, q, t)))


def sample_map(world: World, center, half_xy=125.0, n_surf=3_000_000, n_corner=1_000_000, seed=7):
    """Dense world-frame samples over the 250 x 250 m window; the caller voxel-filters them
    (map import) to obtain the ~750k surf / ~250k corner centroid map of SURVEY 8d C-3."""
    rng = np.random.default_rng(seed)
    radius = half_xy * math.sqrt(2.0)
    s = sample_surface_points(world, center, radius, n_surf, rng, 0.01)
    c = sample_edge_points(world, center, radius, n_corner, rng, 0.01)
    ks = (np.abs(s[:, 0] - center[0]) < half_xy) & (np.abs(s[:, 1] - center[1]) < half_xy)
    kc = (np.abs(c[:, 0] - center[0]) < half_xy) & (np.abs(c[:, 1] - center[1]) < half_xy)
    return to_xyzi(c[kc]), to_xyzi(s[ks])


def dense_map(center, half_xy=125.0, n_surf=750_000, n_corner=250_000, seed=7,
              plane_res=0.8, line_res=0.4):
    """SURVEY 8d C-3 map: planes sampled on a 0.8 m lattice and line segments on a 0.4 m
    lattice inside a 250 x 250 x 150 m box, jittered, until ~n_surf / ~n_corner points."""
    rng = np.random.default_rng(seed)
    cx, cy, cz = center
    surf = []
    # ground lattice
    g = np.arange(-half_xy, half_xy, plane_res) + 0.37
    gx, gy = np.meshgrid(g, g, indexing="ij")
    ground = np.stack([cx + gx.ravel(), cy + gy.ravel(), np.zeros(gx.size)], 1)
    surf.append(ground)
    n_have = len(ground)
    # vertical walls until enough
    while n_have < n_surf:
        L = rng.uniform(20, 120)
        Hh = rng.uniform(4, 40)
        x0, y0 = rng.uniform(-half_xy, half_xy - 1, 2)
        horiz = rng.random() < 0.5
        u = np.arange(0, L, plane_res)
        v = np.arange(0.4, Hh, plane_res)
        uu, vv = np.meshgrid(u, v, indexing="ij")
        if horiz:
            w = np.stack([cx + x0 + uu.ravel(), np.full(uu.size, cy + y0), vv.ravel()], 1)
        else:
            w = np.stack([np.full(uu.size, cx + x0), cy + y0 + uu.ravel(), vv.ravel()], 1)
        k = (np.abs(w[:, 0] - cx) < half_xy) & (np.abs(w[:, 1] - cy) < half_xy)
        surf.append(w[k])
        n_have += int(k.sum())
    surf = np.concatenate(surf, 0)[:n_surf]
    surf = surf + rng.normal(0, 0.02, surf.shape)
    corner = []
    n_have = 0
    while n_have < n_corner:
        x0, y0 = rng.uniform(-half_xy, half_xy, 2)
        Hh = rng.uniform(3, 30)
        z = np.arange(0.2, Hh, line_res)
        seg = np.stack([np.full(z.size, cx + x0), np.full(z.size, cy + y0), z], 1)
        corner.append(seg)
        n_have += len(seg)
    corner = np.concatenate(corner, 0)[:n_corner]
    corner = corner + rng.normal(0, 0.01, corner.shape)
    return to_xyzi(corner), to_xyzi(surf)


# --------------------------------------------------------------------------- ray caster
def beam_elevations_deg(n_scans: int):
    """Beam elevations at the centres of the reference's ring bins
    (Aloam/src/scanRegistration.cpp:169-192)."""
    if n_scans == 64:
        ids = np.arange(64)
        return np.where(ids < 32, 2.0 - ids / 3.0, -8.83 - (ids - 32) / 2.0)
    if n_scans == 32:
        ids = np.arange(32)
        return (ids + 0.5) * 4.0 / 3.0 - 92.0 / 3.0
    if n_scans == 16:
        return np.arange(16) * 2.0 - 15.0
    raise ValueError("n_scans must be 16, 32 or 64")


def raycast_sweep(world: World, q, t, n_scans=64, n_az=1875, rng=None, noise=0.02,
                  max_range=80.0, min_range=0.5):
    """Ray-cast one sweep; returns float32 [n,4] x,y,z,intensity in the SENSOR frame,
    ring-major order (ring 0 first), azimuth decreasing within a ring as a Velodyne
    spinning clockwise produces (so that -atan2(y,x) increases)."""
    rng = rng or np.random.default_rng(0)
    elev = np.deg2rad(beam_elevations_deg(n_scans))
    az0 = rng.uniform(-0.01, 0.01)
    az = az0 - (np.arange(n_az) + 0.25) * (2 * math.pi / n_az)
    ee, aa = np.meshgrid(elev, az, indexing="ij")
    d_s = np.stack([np.cos(ee) * np.cos(aa), np.cos(ee) * np.sin(aa), np.sin(ee)], -1).reshape(-1, 3)
    R = quat_to_rot(q)
    d_w = d_s @ R.T
    o = np.asarray(t, np.float64)
    best = np.full(len(d_w), np.inf)
    # ground
    dz = d_w[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = (world.ground_z - o[2]) / dz
    tg = np.where((dz < -1e-9) & (tg > 0), tg, np.inf)
    best = np.minimum(best, tg)
    boxes, poles = _nearby(world, o, max_range)
    # boxes: slab method
    for b in boxes:
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = (b[0:3] - o) / d_w
            t2 = (b[3:6] - o) / d_w
        tmin = np.nanmax(np.minimum(t1, t2), axis=1)
        tmax = np.nanmin(np.maximum(t1, t2), axis=1)
        hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 0)
        best = np.where(hit & (tmin < best), tmin, best)
    # poles: vertical cylinders
    dxy2 = d_w[:, 0] ** 2 + d_w[:, 1] ** 2
    for p in poles:
        ox, oy = o[0] - p[0], o[1] - p[1]
        bq = ox * d_w[:, 0] + oy * d_w[:, 1]
        cq = ox * ox + oy * oy - p[2] * p[2]
        disc = bq * bq - dxy2 * cq
        with np.errstate(divide="ignore", invalid="ignore"):
            th = (-bq - np.sqrt(np.where(disc > 0, disc, np.nan))) / dxy2
        zh = o[2] + th * d_w[:, 2]
        hit = (disc > 0) & (th > 0) & (zh >= 0) & (zh <= p[3])
        best = np.where(hit & (th < best), th, best)
    ok = np.isfinite(best) & (best < max_range) & (best > min_range)
    rr = best + rng.normal(0, noise, best.shape)
    pts = d_s * np.where(ok, rr, 0.0)[:, None]
    out = np.zeros((int(ok.sum()), 4), np.float32)
    out[:, :3] = pts[ok]
    out[:, 3] = 0.5
    return out


def raycast_sweep_torch(world: World, q, t, n_scans=64, n_az=1875, rng=None, noise=0.02,
                        max_range=80.0, min_range=0.5, device="cuda"):
    """raycast_sweep with the ray / primitive tests evaluated by torch on `device` (float64, chunked over the
    primitives): the same geometry and output layout, fast enough for the pole-dense bench city (thousands of
    cylinders in range).  Workload generation only -- not part of the product path."""
    import torch
    rng = rng or np.random.default_rng(0)
    elev = np.deg2rad(beam_elevations_deg(n_scans))
    az0 = rng.uniform(-0.01, 0.01)
    az = az0 - (np.arange(n_az) + 0.25) * (2 * math.pi / n_az)
    ee, aa = np.meshgrid(elev, az, indexing="ij")
    d_s = np.stack([np.cos(ee) * np.cos(aa), np.cos(ee) * np.sin(aa), np.sin(ee)], -1).reshape(-1, 3)
    R = quat_to_rot(q)
    o = np.asarray(t, np.float64)
    boxes, poles = _nearby(world, o, max_range)
    dev = torch.device(device)
    dw = torch.from_numpy(d_s @ R.T).to(dev)                      # [n, 3]
    ot = torch.from_numpy(o).to(dev)
    inf = torch.tensor(float("inf"), dtype=torch.float64, device=dev)
    best = torch.full((dw.shape[0],), float("inf"), dtype=torch.float64, device=dev)
    dz = dw[:, 2]
    tg = (world.ground_z - ot[2]) / dz
    best = torch.minimum(best, torch.where((dz < -1e-9) & (tg > 0), tg, inf))
    inv = 1.0 / dw                                                # +-inf for axis-parallel rays, as numpy's division
    for b0 in range(0, len(boxes), 32):
        bx = torch.from_numpy(boxes[b0:b0 + 32]).to(dev)          # [m, 6]
        t1 = (bx[None, :, 0:3] - ot) * inv[:, None, :]
        t2 = (bx[None, :, 3:6] - ot) * inv[:, None, :]
        lo = torch.nan_to_num(torch.minimum(t1, t2), nan=-float("inf"))
        hi = torch.nan_to_num(torch.maximum(t1, t2), nan=float("inf"))
        tmin = lo.amax(dim=2)
        tmax = hi.amin(dim=2)
        hit = (tmax >= torch.clamp(tmin, min=0.0)) & (tmin > 0)
        best = torch.minimum(best, torch.where(hit, tmin, inf).amin(dim=1))
    dxy2 = dw[:, 0] ** 2 + dw[:, 1] ** 2
    for p0 in range(0, len(poles), 64):
        pl = torch.from_numpy(poles[p0:p0 + 64]).to(dev)          # [m, 4] x, y, radius, height
        ox, oy = ot[0] - pl[None, :, 0], ot[1] - pl[None, :, 1]
        bq = ox * dw[:, None, 0] + oy * dw[:, None, 1]
        cq = ox * ox + oy * oy - pl[None, :, 2] ** 2
        disc = bq * bq - dxy2[:, None] * cq
        th = (-bq - torch.sqrt(torch.clamp(disc, min=0.0))) / dxy2[:, None]
        zh = ot[2] + th * dw[:, None, 2]
        hit = (disc > 0) & (th > 0) & (zh >= 0) & (zh <= pl[None, :, 3])
        best = torch.minimum(best, torch.where(hit, th, inf).amin(dim=1))
    best = best.cpu().numpy()
    ok = np.isfinite(best) & (best < max_range) & (best > min_range)
    rr = best + rng.normal(0, noise, best.shape)
    pts = d_s * np.where(ok, rr, 0.0)[:, None]
    out = np.zeros((int(ok.sum()), 4), np.float32)
    out[:, :3] = pts[ok]
    out[:, 3] = 0.5
    return out


# --------------------------------------------------------------------------- bench city (C-3)
def make_city(center=(0.0, 0.0), half=125.0, seed=7, pitch=15.0, footprint=(9.0, 12.0),
              height=(12.0, 40.0), pole_pitch=4.2, street_radius=30.0) -> World:
    """Dense block city filling the 250 x 250 m mapping window of SURVEY 8d C-3: a lattice of
    buildings (planes -> ~750 k surf voxels at 0.8 m) and a lattice of poles plus the building
    edges (lines -> ~250 k corner voxels at 0.4 m).  A ring road of radius `street_radius`
    around the centre is kept free of buildings for the sensor to drive on."""
    rng = np.random.default_rng(seed)
    cx, cy = center
    boxes, poles = [], []
    g = np.arange(-half + pitch / 2, half, pitch)
    for x in g:
        for y in g:
            r = math.hypot(x, y)
            if abs(r - street_radius) < 9.0:
                continue
            sx, sy = rng.uniform(*footprint, size=2)
            h = rng.uniform(*height)
            boxes.append([cx + x - sx / 2, cy + y - sy / 2, 0.0, cx + x + sx / 2, cy + y + sy / 2, h])
    gp = np.arange(-half + 1.0, half, pole_pitch)
    for x in gp:
        for y in gp:
            px, py = x + rng.uniform(-0.8, 0.8), y + rng.uniform(-0.8, 0.8)
            inside = False
            for b in boxes:
                if b[0] - 0.3 < cx + px < b[3] + 0.3 and b[1] - 0.3 < cy + py < b[4] + 0.3:
                    inside = True
                    break
            if not inside:
                poles.append([cx + px, cy + py, 0.1, rng.uniform(8.0, 36.0)])
    w = World(np.asarray(boxes, np.float64), np.asarray(poles, np.float64), street_radius, seed)
    w.extra["center"] = (cx, cy)
    return w


def city_pose(world: World, s: float):
    """Pose on the ring road of a make_city() world after arc length s."""
    cx, cy = world.extra.get("center", (0.0, 0.0))
    R = world.loop_radius
    ang = s / R
    t = np.array([cx + R * math.cos(ang), cy + R * math.sin(ang), SENSOR_HEIGHT])
    yaw = ang + math.pi / 2.0
    return np.array([0.0, 0.0, math.sin(yaw / 2), math.cos(yaw / 2)]), t


def sample_box_roofs(world: World, center, radius, n, rng):
    boxes, _ = _nearby(world, center, radius)
    if not len(boxes) or n <= 0:
        return np.zeros((0, 3))
    b = boxes[rng.integers(0, len(boxes), n)]
    return np.stack([b[:, 0] + rng.uniform(0, 1, n) * (b[:, 3] - b[:, 0]),
                     b[:, 1] + rng.uniform(0, 1, n) * (b[:, 4] - b[:, 1]), b[:, 5]], 1)
