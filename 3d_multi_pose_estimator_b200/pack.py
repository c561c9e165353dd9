"""Host-side frame packer: reference JSON frames -> the packed skeleton batch of include/b200pose.h.

Replaces the per-frame `json.loads` + per-joint Python loops of the reference
(skeleton_matching/graph_generator.py:573-605, utils/pose_estimator_dataset_from_json.py:237-289) with
one pass that keeps pixel coordinates as the float64 values the JSON carried (the reference normalises
them in float64 before rounding to fp32, graph_generator.py:496-497).
"""
from __future__ import annotations

import dataclasses
import json
from typing import Dict, List, Optional, Sequence

import numpy as np

from .config import CameraConfig, N_JOINTS


@dataclasses.dataclass
class PackedBatch:
    n_frames: int
    sk_xy: np.ndarray        # [S,18,2] f64
    sk_vp: np.ndarray        # [S,18,2] f32
    sk_mask: np.ndarray      # [S] u32
    sk_cam: np.ndarray       # [S] i32   camera index (camera_names order)
    head_off: np.ndarray     # [B+1] i32
    node_off: np.ndarray     # [B+1] i32
    max_heads: int
    max_enodes: int
    skeletons: Optional[List[List[dict]]] = None      # per frame, per head: the skeleton dict (jsons_for_head)
    skeleton_index: Optional[List[List[int]]] = None  # per frame, per head: index within its camera's list

    @property
    def n_heads(self) -> int:
        return int(self.head_off[-1])

    @property
    def n_nodes(self) -> int:
        return int(self.node_off[-1])

    @property
    def n_enodes(self) -> int:
        return self.n_nodes - self.n_heads

    @property
    def n_edges(self) -> int:
        return self.n_heads + 5 * self.n_enodes

    def input_bytes(self) -> int:
        return sum(a.nbytes for a in (self.sk_xy, self.sk_vp, self.sk_mask, self.sk_cam, self.head_off, self.node_off))

    def tile(self, reps: int) -> "PackedBatch":
        """The same frames repeated `reps` times (large synthetic batches)."""
        H = np.diff(self.head_off); N = np.diff(self.node_off)
        return PackedBatch(
            n_frames=self.n_frames * reps, sk_xy=np.tile(self.sk_xy, (reps, 1, 1)), sk_vp=np.tile(self.sk_vp, (reps, 1, 1)),
            sk_mask=np.tile(self.sk_mask, reps), sk_cam=np.tile(self.sk_cam, reps),
            head_off=np.concatenate([[0], np.cumsum(np.tile(H, reps))]).astype(np.int32),
            node_off=np.concatenate([[0], np.cumsum(np.tile(N, reps))]).astype(np.int32),
            max_heads=self.max_heads, max_enodes=self.max_enodes)

    def slice(self, lo: int, hi: int) -> "PackedBatch":
        """Frames [lo, hi) as their own batch (multi-GPU sharding)."""
        h0, h1 = int(self.head_off[lo]), int(self.head_off[hi])
        return PackedBatch(
            n_frames=hi - lo, sk_xy=self.sk_xy[h0:h1], sk_vp=self.sk_vp[h0:h1], sk_mask=self.sk_mask[h0:h1],
            sk_cam=self.sk_cam[h0:h1], head_off=(self.head_off[lo:hi + 1] - self.head_off[lo]).astype(np.int32),
            node_off=(self.node_off[lo:hi + 1] - self.node_off[lo]).astype(np.int32),
            max_heads=self.max_heads, max_enodes=self.max_enodes,
            skeletons=None if self.skeletons is None else self.skeletons[lo:hi],
            skeleton_index=None if self.skeleton_index is None else self.skeleton_index[lo:hi])


def edge_nodes_of_groups(sizes: Sequence[int]) -> int:
    """M = sum_{i<j} n_i n_j (graph_generator.py:854-864)."""
    s = int(sum(sizes))
    return (s * s - int(sum(n * n for n in sizes))) // 2


def pack_skeleton(sk: dict):
    """One skeleton dict {"<joint>": [joint, x, y, valid, prob]} -> (xy [18,2] f64, vp [18,2] f32, joint mask).
    The "ID" key is ignored (graph_generator.py:483, pose_estimator_dataset_from_json.py:73,255)."""
    a = np.zeros((N_JOINTS, 2), dtype=np.float64)
    b = np.zeros((N_JOINTS, 2), dtype=np.float32)
    m = 0
    for j, v in sk.items():
        if j == "ID":
            continue
        ji = int(j)
        a[ji, 0] = v[1]; a[ji, 1] = v[2]
        b[ji, 0] = v[3]; b[ji, 1] = v[4]
        m |= 1 << ji
    return a, b, m


def pack_frames(frames: Sequence[Dict[str, list]], cfg: CameraConfig, keep_json: bool = True) -> PackedBatch:
    """frames: reference frame dicts {camera: [json_string | list_of_skeletons, ...]}.

    Head order = frame-dict camera order restricted to used_cameras_skeleton_matching, then skeleton
    order; skeletons with no joint keys are skipped (graph_generator.py:583-601). The "ID" key is ignored
    (graph_generator.py:483)."""
    sm_names = {cfg.camera_names[i]: i for i in cfg.used_sm}
    xy, vp, mask, cam = [], [], [], []
    head_off, node_off = [0], [0]
    all_sk, all_idx = [], []
    max_heads = max_enodes = 0
    for frame in frames:
        sizes = []
        f_sk, f_idx = [], []
        for camera, payload in frame.items():
            c = sm_names.get(camera)
            if c is None:
                continue
            skeletons = payload[0]
            if isinstance(skeletons, str):
                skeletons = json.loads(skeletons)
            n = 0
            for idx, sk in enumerate(skeletons):
                a, b, m = pack_skeleton(sk)
                if m == 0:
                    continue
                xy.append(a); vp.append(b); mask.append(m); cam.append(c)
                if keep_json:
                    f_sk.append(sk); f_idx.append(idx)
                n += 1
            if n:
                sizes.append(n)
        H = sum(sizes)
        M = edge_nodes_of_groups(sizes)
        head_off.append(head_off[-1] + H)
        node_off.append(node_off[-1] + H + M)
        max_heads = max(max_heads, H)
        max_enodes = max(max_enodes, M)
        all_sk.append(f_sk); all_idx.append(f_idx)
    S = len(xy)
    return PackedBatch(
        n_frames=len(frames),
        sk_xy=np.stack(xy) if S else np.zeros((0, N_JOINTS, 2), np.float64),
        sk_vp=np.stack(vp) if S else np.zeros((0, N_JOINTS, 2), np.float32),
        sk_mask=np.array(mask, dtype=np.uint32), sk_cam=np.array(cam, dtype=np.int32),
        head_off=np.array(head_off, dtype=np.int32), node_off=np.array(node_off, dtype=np.int32),
        max_heads=max_heads, max_enodes=max_enodes,
        skeletons=all_sk if keep_json else None, skeleton_index=all_idx if keep_json else None)
