"""Drop-in for the reference's utils/pose_estimator_utils.py (:17-75).

`triangulate` (the triangulation baseline, :52-75) runs on the GPU: fp64 undistortion (cv2.undistortPoints),
pairwise register-resident 4x4 DLT solves (cv2.triangulatePoints), upper median + 5 cm filter, mean. The small
tensor helpers keep the reference's signatures.
"""
import numpy as np
import torch

import _b200pose_runtime as rt

if torch.cuda.is_available() is True:
    device = torch.device('cuda')
else:
    device = torch.device('cpu')


def camera_matrix(cam_idx, use_cuda=True):
    """3x3 fp32 intrinsics of camera `cam_idx` (pose_estimator_utils.py:17-30)."""
    p = rt.parameters()
    dev = torch.device('cuda') if (torch.cuda.is_available() and use_cuda) else torch.device('cpu')
    K = torch.zeros((3, 3), dtype=torch.float32)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[2, 2] = p.fx[cam_idx], p.fy[cam_idx], p.cx[cam_idx], p.cy[cam_idx], 1.0
    return K.to(dev)


def from_homogeneous(v):
    return (v / v[-1])[:-1]


def from_homogeneous2(v):
    return v / v[-1]


def get_distortion_coefficients(cam_idx):
    p = rt.parameters()
    return torch.tensor([p.kd0[cam_idx], p.kd1[cam_idx], p.kd2[cam_idx]], device=device)


def apply_distortion(kd, v):
    """Radial part of the Brown model on normalised coordinates (training-loss helper, :44-50)."""
    r2 = (v[:-1] * v[:-1]).sum(dim=0)
    radial = 1 + kd[0] * r2 + kd[1] * r2 * r2 + kd[2] * r2 * r2 * r2
    out = v.clone()
    out[0] = v[0] * radial
    out[1] = v[1] * radial
    return out


def triangulate(points_2D, camera_matrices, distortion_coefficients, projection_matrices, median_chek_axis):
    """points_2D: {joint: {camera: [x, y]}}; the three tables are indexed by the same camera keys
    (test/metrics_from_triangulation.py:237-249). Returns {joint: 3x1 float64 array} like the reference."""
    ctx = rt.context()
    joints = list(range(rt.pkg.N_JOINTS))                      # parameters.joint_list (COCO-18)
    cams = []
    for j in joints:
        for c in points_2D.get(str(j), {}):
            if c not in cams:
                cams.append(c)
    if len(cams) < 2:
        return dict()
    C = len(cams)
    k64 = np.zeros((C, 4)); dist = np.zeros((C, 5)); p64 = np.zeros((C, 3, 4))
    for i, c in enumerate(cams):
        K = np.asarray(camera_matrices[c], dtype=np.float64)
        k64[i] = [K[0, 0], K[1, 1], K[0, 2], K[1, 2]]
        d = np.asarray(distortion_coefficients[c], dtype=np.float64).reshape(-1)
        dist[i, :min(5, d.size)] = d[:5]
        p64[i] = np.asarray(projection_matrices[c], dtype=np.float64)
    xy = np.zeros((C, rt.pkg.N_JOINTS, 2)); mask = np.zeros(C, dtype=np.uint32)
    for j in joints:
        for c, pt in points_2D.get(str(j), {}).items():
            i = cams.index(c)
            xy[i, j] = np.asarray(pt, dtype=np.float64).reshape(-1)[:2]
            mask[i] |= np.uint32(1 << j)
    xyz, m = ctx.triangulate_tables(k64, dist, p64, xy, mask, int(median_chek_axis))
    return {str(j): xyz[j].reshape(3, 1) for j in joints if m[j]}
