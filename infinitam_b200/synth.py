"""Deterministic synthetic depth sequences.

The reference ships no sample frames, so both arms (the B200 engines and the
reference CPU engines used as oracle / CPU baseline) are fed the same
closed-form "analytic room":

  * the inside of an axis-aligned box 4.0 x 2.6 x 4.0 m centred at the origin,
  * four spheres and four boxes at different depths across the field of view (SPHERES / BOXES below; the first
    sphere and the first box are the two objects SURVEY.md 8d names),

seen by a pinhole camera that moves on a circle of radius 0.25 m in the xz
plane around (0, 0, -0.6) while yawing 0.25 deg per frame and pitching
2 deg * sin(2 pi k / 100).  Depth is the z-depth of the first hit in metres,
stored like a sensor would: (short)(z * 1000 + 0.5) millimetres.

Why more than the two objects of SURVEY.md 8d: against the planar walls, with one sphere and one box only, lateral
translation and yaw are nearly indistinguishable for point-to-plane ICP, and the REFERENCE's own tracker (which stops a
pyramid level as soon as the step is < 6e-3) loses the trajectory after ~10 frames (0.35 rad off after 30 frames,
measured with oracle/_ref).  With the additional objects the reference tracks all 100 frames to ~1e-3 rad / 4 mm, so the
benchmark measures a working SLAM loop instead of a diverged one.

Only numpy is used; nothing here is on the product path.
"""
from __future__ import annotations

import math
import os

import numpy as np

ROOM_HALF = np.array([2.0, 1.3, 2.0])
SPHERES = [((0.6, 0.3, 1.2), 0.5), ((-0.3, -0.6, 1.6), 0.35), ((1.4, 0.8, 0.6), 0.4), ((-1.5, -0.2, 0.9), 0.3)]
BOXES = [((-0.8, 0.9, 1.3), (0.3, 0.4, 0.25)), ((0.2, 1.0, 0.4), (0.25, 0.3, 0.25)), ((1.2, -0.7, 1.7), (0.3, 0.25, 0.3)),
         ((-1.6, 0.7, -0.2), (0.3, 0.6, 0.3))]


def intrinsics_for(width: int, height: int):
    """fx = fy = 580 at 640x480 (ITMLib/Objects/ITMIntrinsics.h:49), scaled with the width."""
    s = width / 640.0
    return (580.0 * s, 580.0 * s, width / 2.0, height / 2.0)


def _rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def _rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def camera_to_world(k: int, n_period: int = 100):
    """4x4 world-from-camera transform of frame k (double precision)."""
    phase = 2.0 * math.pi * k / n_period
    pos = np.array([0.25 * math.cos(phase), 0.0, -0.6 + 0.25 * math.sin(phase)])
    yaw = math.radians(0.25 * k)
    pitch = math.radians(2.0 * math.sin(phase))
    T = np.eye(4)
    T[:3, :3] = _rot_y(yaw) @ _rot_x(pitch)
    T[:3, 3] = pos
    return T


def ground_truth_pose(k: int, n_period: int = 100):
    """Camera-from-world M_k in the engine's world frame (= camera frame 0), 4x4 row-major double."""
    return np.linalg.inv(camera_to_world(k, n_period)) @ camera_to_world(0, n_period)


def render_depth(k: int, width: int = 640, height: int = 480, intr=None, noise: bool = False,
                 n_period: int = 100) -> np.ndarray:
    """Depth frame k as int16 millimetres, shape (height, width)."""
    fx, fy, cx, cy = intr if intr is not None else intrinsics_for(width, height)
    T = camera_to_world(k, n_period)
    R, o = T[:3, :3], T[:3, 3]
    u = (np.arange(width, dtype=np.float64) - cx) / fx
    v = (np.arange(height, dtype=np.float64) - cy) / fy
    d_cam = np.stack(np.broadcast_arrays(u[None, :], v[:, None], np.ones((1, 1))), axis=-1)  # z = 1
    d = d_cam @ R.T  # world direction, parameter t == camera z-depth
    with np.errstate(divide="ignore", invalid="ignore"):
        # room: we are inside, take the exit distance
        t1 = (ROOM_HALF - o) / d
        t2 = (-ROOM_HALF - o) / d
        t_room = np.min(np.maximum(t1, t2), axis=-1)
        z = t_room
        a = np.sum(d * d, axis=-1)
        for centre, radius in SPHERES:
            oc = o - np.asarray(centre)
            b = 2.0 * np.sum(d * oc, axis=-1)
            c = float(oc @ oc) - radius ** 2
            disc = b * b - 4 * a * c
            t_s = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
            z = np.minimum(z, np.where(t_s > 0, t_s, np.inf))
        for centre, half in BOXES:  # slab test, entry distance
            lo = (np.asarray(centre) - np.asarray(half) - o) / d
            hi = (np.asarray(centre) + np.asarray(half) - o) / d
            t_in = np.max(np.minimum(lo, hi), axis=-1)
            t_out = np.min(np.maximum(lo, hi), axis=-1)
            z = np.minimum(z, np.where((t_in < t_out) & (t_in > 0), t_in, np.inf))
    if noise:
        rng = np.random.default_rng(20261017 + k)
        z = z + rng.normal(0.0, 0.001, size=z.shape)
        drop = rng.random(z.shape) < 0.01
        z = np.where(drop, 0.0, z)
    mm = np.floor(z * 1000.0 + 0.5)
    mm = np.clip(mm, 0, 32767)
    return mm.astype(np.int16)


def sequence(n_frames: int, width: int = 640, height: int = 480, noise: bool = False,
             start: int = 0) -> np.ndarray:
    """(n_frames, height, width) int16, C-contiguous."""
    out = np.empty((n_frames, height, width), dtype=np.int16)
    intr = intrinsics_for(width, height)

    def one(i):
        out[i] = render_depth(start + i, width, height, intr, noise)

    workers = min(n_frames, os.cpu_count() or 1, 32)
    if workers > 1:  # numpy releases the GIL inside its array loops: frames render in parallel, results are unchanged
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=workers) as ex:
            list(ex.map(one, range(n_frames)))
    else:
        for i in range(n_frames):
            one(i)
    return out
