"""Camera/calibration configuration for the per-frame inference path.

Mirrors what the reference derives at module-import time from `parameters.parameters`
and the pickled TransformManager (reference: skeleton_matching/graph_generator.py:32-52,
utils/pose_estimator_dataset_from_json.py:28-47, parameters.py:45-118).

A `CameraConfig` is plain numpy data so it can be (a) built from the reference's own
`parameters` module + pickle when running as a drop-in, or (b) loaded from a small `.npz`
fixture on a box where the reference tree does not exist.
"""
from __future__ import annotations

import dataclasses
import pickle
from typing import List, Sequence

import numpy as np

N_JOINTS = 18            # COCO-18, reference parameters.py:8
JOINT_FEATS_SM = 10      # graph features per joint, graph_generator.py:128-140
JOINT_FEATS_MLP = 14     # parameters.numbers_per_joint


@dataclasses.dataclass
class CameraConfig:
    """All per-configuration constants of the hot path.

    `camera_names` is `parameters.camera_names` (all cameras); `used_sm` / `used_pe` are
    indices into it for `used_cameras_skeleton_matching` / `used_cameras`.
    """
    name: str
    image_width: float
    image_height: float
    camera_names: List[str]
    fx: np.ndarray
    fy: np.ndarray
    cx: np.ndarray
    cy: np.ndarray
    kd0: np.ndarray  # k1
    kd1: np.ndarray  # k2
    kd2: np.ndarray  # k3
    p1: np.ndarray
    p2: np.ndarray
    T_root2cam: np.ndarray      # [C,4,4] float64, tm.get_transform('root', cam)
    used_sm: List[int]
    used_pe: List[int]
    min_number_of_views: int = 2
    up_axis: int = 1            # axes_3D['Z'][0]
    up_sign: float = -1.0       # axes_3D['Z'][1]
    median_axis: int = 2        # axes_3D['Y'][0]  (triangulate() median filter axis)

    # ---- derived tables, same arithmetic as the reference ------------------------------
    @property
    def n_cameras(self) -> int:
        return len(self.camera_names)

    @property
    def V_sm(self) -> int:
        return len(self.used_sm)

    @property
    def V_pe(self) -> int:
        return len(self.used_pe)

    @property
    def n_features_sm(self) -> int:
        return 2 + N_JOINTS * JOINT_FEATS_SM * self.V_sm

    @property
    def mlp_in(self) -> int:
        return N_JOINTS * JOINT_FEATS_MLP * self.V_pe

    @property
    def used_sm_names(self) -> List[str]:
        return [self.camera_names[i] for i in self.used_sm]

    @property
    def used_pe_names(self) -> List[str]:
        return [self.camera_names[i] for i in self.used_pe]

    def sm_table_camera(self, c: int) -> int:
        """Camera whose fp32 tables (K^-1, camera->root, centre) the head features of camera c are built from.
        graph_generator.py:38-52 appends the tables walking camera_names and keeping the cameras that are in
        used_cameras_skeleton_matching; HumanGraphFromView indexes them with used_cameras_skeleton_matching.index(camera)
        (:232-233): slot s holds the s-th used camera in rig order. c itself whenever the used list is in rig order."""
        if c not in self.used_sm:
            return c
        return sorted(self.used_sm)[self.used_sm.index(c)]

    def K32(self, c: int) -> np.ndarray:
        """3x3 float32 intrinsics, as torch.tensor([[fx,0,cx],...]) (pose_estimator_utils.py:17-30)."""
        return np.array([[self.fx[c], 0.0, self.cx[c]],
                         [0.0, self.fy[c], self.cy[c]],
                         [0.0, 0.0, 1.0]], dtype=np.float32)

    def Kinv32(self, c: int) -> np.ndarray:
        """torch.inverse(K_fp32) (graph_generator.py:50). torch's CPU inverse is LAPACK
        sgetrf/sgetri; we call torch itself once per camera on the host (setup time, not the
        hot path) so the bits are the library's own."""
        import torch
        return torch.inverse(torch.from_numpy(self.K32(c))).numpy().copy()

    def T_root2cam32(self, c: int) -> np.ndarray:
        return self.T_root2cam[c].astype(np.float32)

    def T_cam2root32(self, c: int) -> np.ndarray:
        """float32 cast of inv(T_root2cam) (graph_generator.py:45-46 via pytransform3d)."""
        return np.linalg.inv(self.T_root2cam[c]).astype(np.float32)

    def centre32(self, c: int) -> np.ndarray:
        """T_cam2root32 @ [0,0,0,1] (graph_generator.py:52) == last column."""
        return self.T_cam2root32(c)[:, 3].copy()

    def dist64(self, c: int) -> np.ndarray:
        """[k1,k2,p1,p2,k3] float64 (pose_estimator_dataset_from_json.py:45)."""
        return np.array([self.kd0[c], self.kd1[c], self.p1[c], self.p2[c], self.kd2[c]], dtype=np.float64)

    def K64_from32(self, c: int) -> np.ndarray:
        """camera_matrix(...).numpy() is float32; cv2 promotes to float64 (dataset.py:43)."""
        return self.K32(c).astype(np.float64)

    def P64(self, c: int) -> np.ndarray:
        """projection matrix = T_root2cam[0:3,:] float64 (dataset.py:47)."""
        return self.T_root2cam[c][0:3, :].astype(np.float64)

    # ---- (de)serialisation --------------------------------------------------------------
    def to_npz(self, path: str) -> None:
        np.savez(path, name=self.name, image_width=self.image_width, image_height=self.image_height,
                 camera_names=np.array(self.camera_names), fx=self.fx, fy=self.fy, cx=self.cx, cy=self.cy,
                 kd0=self.kd0, kd1=self.kd1, kd2=self.kd2, p1=self.p1, p2=self.p2,
                 T_root2cam=self.T_root2cam, used_sm=np.array(self.used_sm), used_pe=np.array(self.used_pe),
                 min_number_of_views=self.min_number_of_views, up_axis=self.up_axis, up_sign=self.up_sign,
                 median_axis=self.median_axis)

    @staticmethod
    def from_npz(path: str) -> "CameraConfig":
        z = np.load(path, allow_pickle=False)
        return CameraConfig(
            name=str(z['name']), image_width=float(z['image_width']), image_height=float(z['image_height']),
            camera_names=[str(s) for s in z['camera_names']],
            fx=z['fx'], fy=z['fy'], cx=z['cx'], cy=z['cy'], kd0=z['kd0'], kd1=z['kd1'], kd2=z['kd2'],
            p1=z['p1'], p2=z['p2'], T_root2cam=z['T_root2cam'],
            used_sm=[int(i) for i in z['used_sm']], used_pe=[int(i) for i in z['used_pe']],
            min_number_of_views=int(z['min_number_of_views']), up_axis=int(z['up_axis']),
            up_sign=float(z['up_sign']), median_axis=int(z['median_axis']))

    @staticmethod
    def from_parameters(parameters, tm=None, name: str = "parameters") -> "CameraConfig":
        """Build from the reference's `parameters.parameters` namedtuple (+ pickled TransformManager)."""
        if tm is None:
            tm = pickle.load(open(parameters.transformations_path, 'rb'))
        names = list(parameters.camera_names)
        idx = list(parameters.cameras)
        f = lambda field: np.array([getattr(parameters, field)[i] for i in idx], dtype=np.float64)
        T = np.stack([np.asarray(tm.get_transform("root", n), dtype=np.float64) for n in names])
        return CameraConfig(
            name=name, image_width=float(parameters.image_width), image_height=float(parameters.image_height),
            camera_names=names, fx=f('fx'), fy=f('fy'), cx=f('cx'), cy=f('cy'),
            kd0=f('kd0'), kd1=f('kd1'), kd2=f('kd2'), p1=f('p1'), p2=f('p2'), T_root2cam=T,
            used_sm=[names.index(c) for c in parameters.used_cameras_skeleton_matching],
            used_pe=[names.index(c) for c in parameters.used_cameras],
            min_number_of_views=int(parameters.min_number_of_views),
            up_axis=int(parameters.axes_3D['Z'][0]), up_sign=float(parameters.axes_3D['Z'][1]),
            median_axis=int(parameters.axes_3D['Y'][0]))


def ring_config(n_views: int = 10, radius: float = 4.0, height: float = 1.5, f: float = 1400.0,
                width: int = 1920, height_px: int = 1080) -> CameraConfig:
    """Synthetic stress configuration (SURVEY.md 8d cfg 5): cameras on a ring looking at the origin,
    world up = -Y like Panoptic, no lens distortion."""
    names, Ts = [], []
    for i in range(n_views):
        ang = 2.0 * np.pi * i / n_views
        centre = np.array([radius * np.cos(ang), -height, radius * np.sin(ang)])
        z = -centre / np.linalg.norm(centre)          # optical axis towards the origin
        down = np.array([0.0, 1.0, 0.0])              # image y points down; world down = +Y
        x = np.cross(down, z); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R = np.stack([x, y, z])                       # rows: camera axes in world coords
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = -R @ centre
        names.append('ring%02d' % i); Ts.append(T)
    V = n_views
    z0 = np.zeros(V)
    return CameraConfig(
        name='ring%d' % V, image_width=float(width), image_height=float(height_px), camera_names=names,
        fx=np.full(V, f), fy=np.full(V, f), cx=np.full(V, width / 2.0), cy=np.full(V, height_px / 2.0),
        kd0=z0.copy(), kd1=z0.copy(), kd2=z0.copy(), p1=z0.copy(), p2=z0.copy(),
        T_root2cam=np.stack(Ts), used_sm=list(range(V)), used_pe=list(range(V)),
        min_number_of_views=2, up_axis=1, up_sign=-1.0, median_axis=2)
