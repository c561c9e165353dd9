"""Synthetic multi-view frames in the reference's JSON frame schema.

frame = {camera_name: [json.dumps([skeleton, ...]), timestamp, 'no_image', bodies_3D]}
skeleton = {"<joint_id>": [joint_id, x_px, y_px, valid, prob]}
(schema: reference panoptic_conversor/get_joints_from_panoptic_model_multi.py:236,281,287)

Datasets/checkpoints are not available offline, so inputs are random 3D skeletons projected
through the configuration's own camera matrices with the full Brown distortion model.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional

import numpy as np

from .config import CameraConfig, N_JOINTS


def project_points(cfg: CameraConfig, c: int, X: np.ndarray):
    """World points [n,3] -> distorted pixel coordinates [n,2] and depth [n] for camera c."""
    T = cfg.T_root2cam[c]
    Xc = X @ T[:3, :3].T + T[:3, 3]
    z = Xc[:, 2]
    zs = np.where(np.abs(z) < 1e-9, 1e-9, z)
    x = Xc[:, 0] / zs
    y = Xc[:, 1] / zs
    k1, k2, p1, p2, k3 = cfg.dist64(c)
    r2 = x * x + y * y
    rad = 1 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
    xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    u = cfg.fx[c] * xd + cfg.cx[c]
    v = cfg.fy[c] * yd + cfg.cy[c]
    return np.stack([u, v], axis=1), z


def random_people(cfg: CameraConfig, rng: np.random.Generator, n_persons: int) -> np.ndarray:
    """[P,18,3] joints in metres: centre U[-1,1]^2 on the ground plane, joints = centre + N(0,0.2^2)
    horizontally, height U(0.1,1.7) along the configuration's up axis."""
    up = cfg.up_axis
    ground = [a for a in range(3) if a != up]
    people = np.zeros((n_persons, N_JOINTS, 3))
    for p in range(n_persons):
        centre = rng.uniform(-1.0, 1.0, size=2)
        horiz = centre[None, :] + rng.normal(0.0, 0.2, size=(N_JOINTS, 2))
        h = rng.uniform(0.1, 1.7, size=N_JOINTS)
        people[p, :, ground[0]] = horiz[:, 0]
        people[p, :, ground[1]] = horiz[:, 1]
        people[p, :, up] = cfg.up_sign * h
    return people


def make_frame(cfg: CameraConfig, seed: int, n_persons: int, *, drop_joint_p: float = 0.0,
               drop_view_p: float = 0.0, rand_conf: bool = False, keep_empty: bool = False,
               with_gt: bool = False, camera_order: Optional[List[int]] = None, all_cameras: bool = False,
               pixel_noise: float = 0.0) -> Dict[str, list]:
    """One synthetic frame. Joints are kept iff in front of the camera and inside the image.

    drop_joint_p / drop_view_p / rand_conf add detector-like raggedness (missing joint keys,
    persons unseen in a view, valid=0 joints and non-unit confidences; keep_empty keeps skeletons
    with no joints, which the reference skips as heads) for the parity tests. all_cameras: the frame carries every
    camera of the rig (parameters.camera_names), used or not. pixel_noise: sigma [px] of Gaussian detector noise on the
    kept joints (its own random stream, so the noiseless frames of a seed do not change).
    """
    rng = np.random.default_rng(seed)
    rng_noise = np.random.default_rng([seed, 0x5eed])
    people = random_people(cfg, rng, n_persons)
    cams = list(cfg.used_sm) if camera_order is None else list(camera_order)
    if all_cameras and camera_order is None:
        cams = list(range(cfg.n_cameras))
    frame: Dict[str, list] = {}
    for c in cams:
        skeletons = []
        kept = []
        for p in range(n_persons):
            if drop_view_p > 0 and rng.random() < drop_view_p:
                continue
            uv, z = project_points(cfg, c, people[p])
            sk = {}
            for j in range(N_JOINTS):
                if not (z[j] > 0 and 0 <= uv[j, 0] < cfg.image_width and 0 <= uv[j, 1] < cfg.image_height):
                    continue
                if drop_joint_p > 0 and rng.random() < drop_joint_p:
                    continue
                valid, prob = 1, 1
                if rand_conf:
                    prob = float(np.round(rng.uniform(0.05, 1.0), 6))
                    valid = 1 if prob > 0.2 else 0
                u, v = float(uv[j, 0]), float(uv[j, 1])
                if pixel_noise > 0:
                    du, dv = rng_noise.normal(0.0, pixel_noise, size=2)
                    u, v = u + float(du), v + float(dv)
                sk[str(j)] = [j, u, v, valid, prob]
            if sk or keep_empty:
                skeletons.append(sk)
                kept.append(p)
        gt = []
        if with_gt:      # bodies_3D in cm, one entry per skeleton of this camera (test/sm_metrics.py:118-121 indexes them together)
            gt = [dict({str(j): (people[p, j] * 100.0).tolist() for j in range(N_JOINTS)}, **{'-1': [0, 0, 0]})
                  for p in kept]
        frame[cfg.camera_names[c]] = [json.dumps(skeletons), 0.0, 'no_image', gt]
    return frame


def make_frames(cfg: CameraConfig, n_frames: int, n_persons: int, base_seed: int = 0, **kw) -> List[Dict[str, list]]:
    """Independent seeds per frame (seed = base + frame index), SURVEY.md 8d."""
    return [make_frame(cfg, base_seed + i, n_persons, **kw) for i in range(n_frames)]
