"""Seeded synthetic RGB-D frames for tests and bench.

A planar colour "poster" of random filled rectangles is viewed by a pinhole
camera (ICL intrinsics, Examples/RGB-D/ICL.yaml:8-11 in the reference) moving
on a smooth 6-DoF trajectory.  Everything is numpy float64 + PCG64, so the same
seed gives the same bytes on every box.  It only makes inputs; it checks nothing.
"""
from __future__ import annotations

import numpy as np

ICL = dict(fx=481.2, fy=-480.0, cx=319.5, cy=239.5, bf=40.0, depth_factor=5000.0)


def make_poster(seed: int, size: int = 2048, nrect: int = 3000) -> np.ndarray:
    """size×size×3 u8 poster: random rectangles, 3-tap binomial blur, N(0,2) noise."""
    rng = np.random.default_rng(seed)
    img = np.full((size, size, 3), 128, np.int32)
    xs = rng.integers(0, size, nrect)
    ys = rng.integers(0, size, nrect)
    ws = rng.integers(8, 121, nrect)
    hs = rng.integers(8, 121, nrect)
    cols = rng.integers(0, 256, (nrect, 3))
    for x, y, w, h, c in zip(xs, ys, ws, hs, cols):
        img[y:y + h, x:x + w] = c
    # binomial [1 2 1]/4 separable blur, integer arithmetic, edge replicate
    p = np.pad(img, ((1, 1), (0, 0), (0, 0)), mode="edge")
    img = (p[:-2] + 2 * p[1:-1] + p[2:] + 2) >> 2
    p = np.pad(img, ((0, 0), (1, 1), (0, 0)), mode="edge")
    img = (p[:, :-2] + 2 * p[:, 1:-1] + p[:, 2:] + 2) >> 2
    noise = np.rint(rng.normal(0.0, 2.0, img.shape)).astype(np.int32)
    return np.clip(img + noise, 0, 255).astype(np.uint8)


def make_lowtex(seed: int, w: int = 640, h: int = 480) -> np.ndarray:
    """Low-texture grey frame (cfg 3): flat-shaded convex quads on a uniform wall."""
    rng = np.random.default_rng(seed)
    img = np.full((h, w), int(rng.integers(90, 140)), np.float64)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(int(rng.integers(6, 13))):
        cx, cy = rng.uniform(0.1 * w, 0.9 * w), rng.uniform(0.1 * h, 0.9 * h)
        rx, ry = rng.uniform(40, 160), rng.uniform(40, 120)
        th = rng.uniform(0, np.pi)
        u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
        v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        k = rng.uniform(-0.3, 0.3)  # keystone -> non-parallel edges
        inside = (np.abs(u) <= rx * (1 + k * v / ry)) & (np.abs(v) <= ry)
        base = img[int(np.clip(cy, 0, h - 1)), int(np.clip(cx, 0, w - 1))]
        grey = base + rng.choice([-1, 1]) * rng.uniform(40, 100)
        img[inside] = np.clip(grey, 0, 255)
    img = img + rng.normal(0.0, 1.5, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def trajectory(n: int, seed: int) -> np.ndarray:
    """n world->camera poses Tcw (n,4,4) float64; ≤2 cm and ≤0.5° between frames."""
    rng = np.random.default_rng(seed + 7919)
    ph = rng.uniform(0, 2 * np.pi, 6)
    out = np.zeros((n, 4, 4))
    for i in range(n):
        t = i * 0.05
        ang = np.deg2rad([4.0 * np.sin(0.9 * t + ph[0]), 4.0 * np.sin(0.7 * t + ph[1]),
                          6.0 * np.sin(0.5 * t + ph[2])])
        twc = np.array([0.25 * np.sin(0.6 * t + ph[3]), 0.2 * np.sin(0.8 * t + ph[4]),
                        -2.2 + 0.3 * np.sin(0.4 * t + ph[5])])
        cx_, sx_ = np.cos(ang[0]), np.sin(ang[0])
        cy_, sy_ = np.cos(ang[1]), np.sin(ang[1])
        cz_, sz_ = np.cos(ang[2]), np.sin(ang[2])
        Rx = np.array([[1, 0, 0], [0, cx_, -sx_], [0, sx_, cx_]])
        Ry = np.array([[cy_, 0, sy_], [0, 1, 0], [-sy_, 0, cy_]])
        Rz = np.array([[cz_, -sz_, 0], [sz_, cz_, 0], [0, 0, 1]])
        Rwc = Rz @ Ry @ Rx
        T = np.eye(4)
        T[:3, :3] = Rwc.T
        T[:3, 3] = -Rwc.T @ twc
        out[i] = T
    return out


def render(poster: np.ndarray, Tcw: np.ndarray, w: int = 640, h: int = 480, K: dict = ICL,
           metres_per_px: float = 0.0022, noise_seed: int | None = None):
    """Render one RGB-D view of the poster plane z=0.  Returns (rgb u8 h×w×3, depth u16 h×w)."""
    S = poster.shape[0]
    sx, sy = w / 640.0, h / 480.0
    fx, fy, cx, cy = K["fx"] * sx, K["fy"] * sy, (K["cx"] + 0.5) * sx - 0.5, (K["cy"] + 0.5) * sy - 0.5
    v, u = np.mgrid[0:h, 0:w].astype(np.float64)
    d = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], -1)
    Rwc = Tcw[:3, :3].T
    twc = -Rwc @ Tcw[:3, 3]
    dw = d @ Rwc.T
    t = -twc[2] / dw[..., 2]
    px = (twc[0] + t * dw[..., 0]) / metres_per_px + S / 2
    py = (twc[1] + t * dw[..., 1]) / metres_per_px + S / 2
    px = np.clip(px, 0, S - 1.001)
    py = np.clip(py, 0, S - 1.001)
    x0 = np.floor(px).astype(np.int64)
    y0 = np.floor(py).astype(np.int64)
    ax = (px - x0)[..., None]
    ay = (py - y0)[..., None]
    P = poster.astype(np.float64)
    val = ((1 - ay) * ((1 - ax) * P[y0, x0] + ax * P[y0, x0 + 1])
           + ay * ((1 - ax) * P[y0 + 1, x0] + ax * P[y0 + 1, x0 + 1]))
    if noise_seed is not None:
        val = val + np.random.default_rng(noise_seed).normal(0.0, 1.0, val.shape)
    rgb = np.clip(np.rint(val), 0, 255).astype(np.uint8)
    depth = np.clip(np.rint(t * K["depth_factor"]), 0, 65535).astype(np.uint16)
    return rgb, depth


def rgb_to_gray(rgb: np.ndarray) -> np.ndarray:
    """cv2-4.x RGB→Y in Q15; the arithmetic Tracking.cc:219-232 triggers."""
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    return ((r * 9798 + g * 19235 + b * 3735 + 16384) >> 15).astype(np.uint8)


def sequence(seed: int, n: int, w: int = 640, h: int = 480, poster_size: int = 2048):
    """n consecutive (gray, depth, Tcw) of the textured scene."""
    poster = make_poster(seed, poster_size)
    T = trajectory(n, seed)
    grays, depths = [], []
    for i in range(n):
        rgb, dep = render(poster, T[i], w, h, noise_seed=seed * 100003 + i)
        grays.append(rgb_to_gray(rgb))
        depths.append(dep)
    return np.stack(grays), np.stack(depths), T
