"""Deterministic synthetic scan pairs + PLY IO for the FCCF-PCR path.

The reference (FCCF.cpp:1655-1661) reads PLY files of float32 x/y/z; its datasets (ETH, RESSO) are
not available offline, so benchmarks and parity tests use these seeded synthetic planar scenes:
an office-like room (6 walls + rotated boxes + curved clutter) and an outdoor block scene.
"""
from __future__ import annotations

import math

import numpy as np

GT_YAW_PITCH_ROLL_DEG = (25.0, 3.0, 2.0)
GT_TRANSLATION = (1.5, -0.8, 0.2)


def _rot_z(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def _rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def _rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])


def ground_truth(yaw_pitch_roll_deg=GT_YAW_PITCH_ROLL_DEG, t=GT_TRANSLATION):
    """4x4 float64 rigid transform mapping the SRC cloud (argv[1]) into the TAR frame (argv[2])."""
    y, p, r = (math.radians(v) for v in yaw_pitch_roll_deg)
    T = np.eye(4)
    T[:3, :3] = _rot_z(y) @ _rot_y(p) @ _rot_x(r)
    T[:3, 3] = t
    return T


class _Scene:
    def __init__(self):
        self.rects = []      # (origin, u, v) parallelogram, or (origin, u, v, True) triangle
        self.spheres = []    # (centre, radius)
        self.cyls = []       # (centre, radius, height)
        self.rough = {}      # rect index -> (amplitude, wavelength) of a long-wave undulation along u ("rough" planes)

    def add_box(self, centre, size, yaw, bottom=False):
        R = _rot_z(yaw)
        c = np.asarray(centre, float)
        hx, hy, hz = (0.5 * s for s in size)
        ax, ay, az = R[:, 0] * hx, R[:, 1] * hy, R[:, 2] * hz
        faces = [(c + ax - ay - az, 2 * ay, 2 * az), (c - ax - ay - az, 2 * ay, 2 * az),
                 (c + ay - ax - az, 2 * ax, 2 * az), (c - ay - ax - az, 2 * ax, 2 * az),
                 (c + az - ax - ay, 2 * ax, 2 * ay)]
        if bottom:
            faces.append((c - az - ax - ay, 2 * ax, 2 * ay))
        self.rects.extend(faces)

    def areas(self):
        a = [np.linalg.norm(np.cross(r[1], r[2])) * (0.5 if len(r) > 3 else 1.0) for r in self.rects]
        a += [4 * math.pi * r * r for (_, r) in self.spheres]
        a += [2 * math.pi * r * h for (_, r, h) in self.cyls]
        return np.asarray(a)

    def sample(self, rng, n, clutter_frac, sigma, undulation=None):
        nr, ns = len(self.rects), len(self.spheres)
        w = self.areas().copy()
        planar, curved = w[:nr].sum(), w[nr:].sum()
        if curved > 0:
            w[:nr] *= (1.0 - clutter_frac) / planar
            w[nr:] *= clutter_frac / curved
        else:
            w /= planar
        counts = rng.multinomial(n, w / w.sum())
        out = []
        for k, rc in enumerate(self.rects):
            o, u, v = rc[0], rc[1], rc[2]
            m = counts[k]
            a, b = rng.random(m), rng.random(m)
            if len(rc) > 3:
                flip = (a + b) > 1.0
                a = np.where(flip, 1.0 - a, a)
                b = np.where(flip, 1.0 - b, b)
            p = o[None, :] + a[:, None] * u[None, :] + b[:, None] * v[None, :]
            if k in self.rough:
                # neighbouring 1 m voxels see normals a few degrees apart: a plane whose roughness (mean angle
                # between the face normal and its voxels' normals, FCCF.cpp:660-667) exceeds rough_threshold_gl
                amp, lam = self.rough[k]
                nrm = np.cross(u, v); nrm = nrm / np.linalg.norm(nrm)
                p = p + (amp * np.sin(2.0 * math.pi * a * np.linalg.norm(u) / lam))[:, None] * nrm[None, :]
            if undulation is not None and k == 0:
                p[:, 2] += undulation(p[:, 0], p[:, 1])
            out.append(p)
        for k, (c, r) in enumerate(self.spheres):
            m = counts[nr + k]
            d = rng.normal(size=(m, 3))
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            out.append(np.asarray(c)[None, :] + r * d)
        for k, (c, r, h) in enumerate(self.cyls):
            m = counts[nr + ns + k]
            th = rng.random(m) * 2 * math.pi
            z = (rng.random(m) - 0.5) * h
            out.append(np.asarray(c)[None, :] + np.stack([r * np.cos(th), r * np.sin(th), z], axis=1))
        p = np.concatenate(out, axis=0)
        p += rng.normal(scale=sigma, size=p.shape)
        return p[rng.permutation(len(p))]


def _indoor_scene(rng):
    sc = _Scene()
    H = 3.0
    ez = np.array([0, 0, H])
    # irregular quadrilateral footprint: no two walls are parallel (>10 deg off), so the room has no
    # 180-degree plane symmetry for the plane-level scoring to fall into
    A, B, Cc, D = (np.array([x, y, -H / 2]) for (x, y) in [(-5.0, -4.0), (5.0, -4.0), (4.0, 2.8), (-4.6, 4.6)])
    for a, b in ((A, B), (B, Cc), (Cc, D), (D, A)):
        sc.rects.append((a, b - a, ez))
    for z in (0.0, H):
        off = np.array([0, 0, z])
        sc.rects.append((A + off, B - A, D - A, True))
        sc.rects.append((Cc + off, B - Cc, D - Cc, True))
    sc.rects.append((np.array([-3.6, 0.6, -H / 2]), np.array([3.3, 1.2, 0.0]), 0.9 * ez))  # partition, ~20 deg
    nb = int(rng.integers(4, 7))
    for _ in range(nb):
        size = (rng.uniform(0.8, 2.2), rng.uniform(0.6, 1.4), rng.uniform(0.7, 2.0))
        c = (rng.uniform(-3.5, 3.5), rng.uniform(-2.8, 2.8), -H / 2 + size[2] / 2)
        sc.add_box(c, size, math.radians(rng.uniform(0, 40)))
    for _ in range(6):
        r = rng.uniform(0.2, 0.6)
        sc.spheres.append(((rng.uniform(-4, 4), rng.uniform(-3, 3), rng.uniform(-1.0, 0.8)), r))
    for _ in range(4):
        r = rng.uniform(0.2, 0.5)
        h = rng.uniform(1.0, 2.2)
        sc.cyls.append(((rng.uniform(-4, 4), rng.uniform(-3, 3), -H / 2 + h / 2), r, h))
    return sc


def _outdoor_scene(rng):
    sc = _Scene()
    G = 200.0
    sc.rects.append((np.array([-G / 2, -G / 2, 0.0]), np.array([G, 0, 0.0]), np.array([0, G, 0.0])))
    for _ in range(int(rng.integers(20, 41))):
        size = (rng.uniform(10, 30), rng.uniform(10, 30), rng.uniform(6, 25))
        c = (rng.uniform(-85, 85), rng.uniform(-85, 85), size[2] / 2)
        sc.add_box(c, size, math.radians(rng.uniform(0, 90)))
    for _ in range(60):
        r = rng.uniform(1.5, 4.0)
        sc.spheres.append(((rng.uniform(-95, 95), rng.uniform(-95, 95), rng.uniform(3, 8)), r))
    return sc


def make_pair(kind="indoor", n_points=50_000, seed=1, overlap=0.7, sigma=0.005, clutter_frac=0.2,
              yaw_pitch_roll_deg=GT_YAW_PITCH_ROLL_DEG, t=GT_TRANSLATION):
    """Returns (src, tar, T_gt): float32 [n,3] clouds and the float64 4x4 with tar ≈ T_gt · src.

    src is argv[1] of the reference CLI, tar argv[2]; the printed matrix maps src into tar's frame
    (SURVEY.md Q1).  Both clouds are independent samplings of one scene (seeds differ); each is
    cropped by a half space so that roughly `overlap` of it is seen by the other.
    """
    rng_scene = np.random.Generator(np.random.PCG64(seed))
    sc = _indoor_scene(rng_scene) if kind in ("indoor", "indoor_rough") else _outdoor_scene(rng_scene)
    if kind == "indoor_rough":      # every second surface undulates: roughness types 1 (rough-rough) and 2 (mixed) appear
        for k in range(1, len(sc.rects), 2):
            sc.rough[k] = (0.045, 4.0)
    und = None
    if kind not in ("indoor", "indoor_rough"):
        und = lambda x, y: 0.4 * np.sin(x / 17.0) * np.cos(y / 23.0)  # noqa: E731
    ext = 5.0 if kind in ("indoor", "indoor_rough") else 100.0
    clouds = []
    for which in (0, 1):
        rng = np.random.Generator(np.random.PCG64([seed, 1000 + which]))
        frac = 0.5 + 0.5 * overlap
        m = int(n_points / frac * 1.08) + 64
        cut = ext * (2.0 * frac - 1.0)
        parts, have = [], 0
        while have < n_points:
            p = sc.sample(rng, m, clutter_frac, sigma, und)
            # cloud 0 drops the far +x end, cloud 1 the far -y end (crops that do not map onto each
            # other under the room's near-symmetries)
            p = p[p[:, 0] < cut] if which == 0 else p[p[:, 1] > -0.8 * cut]
            parts.append(p)
            have += len(p)
        clouds.append(np.concatenate(parts, axis=0)[:n_points])
    T = ground_truth(yaw_pitch_roll_deg, t)
    tar = clouds[0]
    Tinv = np.linalg.inv(T)
    src = clouds[1] @ Tinv[:3, :3].T + Tinv[:3, 3]
    return np.ascontiguousarray(src, np.float32), np.ascontiguousarray(tar, np.float32), T


# --------------------------------------------------------------------------------------------
# PLY (binary little endian / ascii), float32 x y z — the reference's only input format
# --------------------------------------------------------------------------------------------
def write_ply(path, xyz, binary=True):
    xyz = np.ascontiguousarray(xyz, np.float32)
    hdr = ("ply\nformat %s 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nend_header\n"
           % ("binary_little_endian" if binary else "ascii", len(xyz)))
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        if binary:
            f.write(xyz.astype("<f4").tobytes())
        else:
            for p in xyz:
                f.write(("%.9g %.9g %.9g\n" % (p[0], p[1], p[2])).encode("ascii"))


def rotation_error_deg(Ta, Tb):
    """Angle between two rotation blocks; the chord form stays accurate for tiny angles
    (acos of a float32 trace bottoms out around 0.02 degrees)."""
    Ra, Rb = np.asarray(Ta, float)[:3, :3], np.asarray(Tb, float)[:3, :3]
    chord = np.linalg.norm(Ra - Rb) / math.sqrt(2.0)      # = 2 sin(angle / 2) for rotations
    return math.degrees(2.0 * math.asin(min(1.0, chord / 2.0)))


def read_ply(path):
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    hdr = data[:end].decode("ascii").split("\n")
    n = [int(line.split()[2]) for line in hdr if line.startswith("element vertex")][0]
    if any("binary_little_endian" in line for line in hdr):
        return np.frombuffer(data[end:end + 12 * n], "<f4").reshape(n, 3).copy()
    return np.array([[float(v) for v in line.split()[:3]] for line in data[end:].decode().split("\n")[:n]], np.float32)


def translation_error(Ta, Tb):
    return float(np.linalg.norm(np.asarray(Ta, float)[:3, 3] - np.asarray(Tb, float)[:3, 3]))
