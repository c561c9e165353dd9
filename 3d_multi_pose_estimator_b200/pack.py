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

    def validate(self) -> None:
        """The declared frame plan (max_heads / max_enodes size shared-memory plans and the clustering scratch) must cover
        every frame - a stale value in a hand-built or loaded batch would make kernels skip the frames that exceed it."""
        if self.n_frames == 0:
            return
        H = np.diff(np.asarray(self.head_off, dtype=np.int64))
        M = np.diff(np.asarray(self.node_off, dtype=np.int64)) - H
        if len(H) != self.n_frames or (H < 0).any() or (M < 0).any():
            raise ValueError('PackedBatch: head_off / node_off are not non-decreasing offset arrays of %d frames' % self.n_frames)
        if int(H.max()) > self.max_heads or int(M.max()) > self.max_enodes:
            raise ValueError('PackedBatch: a frame has %d heads / %d edge-nodes but the batch declares max_heads=%d, max_enodes=%d'
                             % (int(H.max()), int(M.max()), self.max_heads, self.max_enodes))
        if self.head_off[-1] != len(self.sk_cam):
            raise ValueError('PackedBatch: head_off ends at %d but there are %d skeletons' % (int(self.head_off[-1]), len(self.sk_cam)))

    def input_bytes(self) -> int:
        return sum(a.nbytes for a in (self.sk_xy, self.sk_vp, self.sk_mask, self.sk_cam, self.head_off, self.node_off))

    def save(self, path: str) -> None:
        """Binary ingest format: the packed arrays as an .npz (440 B per skeleton instead of ~1 KB of JSON text); loading
        it costs a memcpy, so a recorded sequence can be fed at GPU speed."""
        np.savez(path, n_frames=self.n_frames, sk_xy=self.sk_xy, sk_vp=self.sk_vp, sk_mask=self.sk_mask, sk_cam=self.sk_cam,
                 head_off=self.head_off, node_off=self.node_off, max_heads=self.max_heads, max_enodes=self.max_enodes)

    @staticmethod
    def load(path: str) -> "PackedBatch":
        z = np.load(path)
        return PackedBatch(n_frames=int(z['n_frames']), sk_xy=z['sk_xy'], sk_vp=z['sk_vp'], sk_mask=z['sk_mask'], sk_cam=z['sk_cam'],
                           head_off=z['head_off'], node_off=z['node_off'], max_heads=int(z['max_heads']), max_enodes=int(z['max_enodes']))

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


def pack_json(text, cfg: CameraConfig, n_threads: int = 0, pinned: bool = False) -> PackedBatch:
    """Native packer: the JSON text of a test file (a list of reference frames) or of one frame -> PackedBatch, parsed
    in parallel by libb200pose.so (b200pose_pack_json). Equivalent to pack_frames(json.loads(text), cfg, keep_json=False)
    but ~100x faster; with pinned=True the arrays are views of page-locked torch tensors, ready for asynchronous
    host->device copies."""
    import ctypes as C
    from . import _lib
    L = _lib.lib()
    data = text.encode('utf-8') if isinstance(text, str) else bytes(text)
    names = cfg.used_sm_names
    arr_names = (C.c_char_p * len(names))(*[n.encode('utf-8') for n in names])
    arr_idx = (C.c_int32 * len(names))(*[int(i) for i in cfg.used_sm])
    handle = C.c_void_p()
    _lib.check(L.b200pose_pack_json(data, len(data), len(names), arr_names, arr_idx, n_threads, C.byref(handle)), 'pack_json')
    try:
        sz = [C.c_int32() for _ in range(5)]
        _lib.check(L.b200pose_packed_sizes(handle, *[C.byref(x) for x in sz]), 'packed_sizes')
        B, S, N, max_heads, max_enodes = [int(x.value) for x in sz]
        if pinned:
            import torch
            mk = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
            sk_xy, sk_vp = mk((S, N_JOINTS, 2), torch.float64), mk((S, N_JOINTS, 2), torch.float32)
            sk_mask, sk_cam = mk((S,), torch.int32).view(np.uint32), mk((S,), torch.int32)
            head_off, node_off, sk_idx = mk((B + 1,), torch.int32), mk((B + 1,), torch.int32), mk((S,), torch.int32)
        else:
            sk_xy, sk_vp = np.empty((S, N_JOINTS, 2), np.float64), np.empty((S, N_JOINTS, 2), np.float32)
            sk_mask, sk_cam = np.empty(S, np.uint32), np.empty(S, np.int32)
            head_off, node_off, sk_idx = np.empty(B + 1, np.int32), np.empty(B + 1, np.int32), np.empty(S, np.int32)
        p = lambda a: C.c_void_p(a.ctypes.data)
        _lib.check(L.b200pose_packed_copy(handle, p(sk_xy), p(sk_vp), p(sk_mask), p(sk_cam), p(head_off), p(node_off), p(sk_idx),
                                          n_threads), 'packed_copy')
    finally:
        L.b200pose_packed_free(handle)
    skeleton_index = [sk_idx[head_off[b]:head_off[b + 1]].tolist() for b in range(B)] if B <= 4096 else None
    return PackedBatch(n_frames=B, sk_xy=sk_xy, sk_vp=sk_vp, sk_mask=sk_mask, sk_cam=sk_cam, head_off=head_off, node_off=node_off,
                       max_heads=max_heads, max_enodes=max_enodes, skeletons=None, skeleton_index=skeleton_index)


def pack_frames_fast(frames: Sequence[Dict[str, list]], cfg: CameraConfig, keep_json: bool = True) -> PackedBatch:
    """pack_frames for frames whose camera payloads carry the skeleton list as a JSON *string* (the reference's wire format,
    SURVEY.md App. A): the strings are handed to the native packer instead of being walked joint by joint in Python - ~8x
    faster for a live frame (1.2 ms -> 0.15 ms), which is most of a live frame's host time. Frames with inline lists go to
    pack_frames. The result is identical (same arrays bit for bit, same `skeletons` / `skeleton_index`)."""
    for f in frames:
        for payload in f.values():
            if not isinstance(payload[0], str):
                return pack_frames(frames, cfg, keep_json=keep_json)
    dq = json.dumps
    text = '[' + ','.join('{' + ','.join(dq(c) + ':[' + dq(p[0]) + ']' for c, p in f.items()) + '}' for f in frames) + ']'
    pb = pack_json(text, cfg, n_threads=1 if len(frames) < 16 else 0)
    if keep_json:
        sm = set(cfg.used_sm_names)
        all_sk, all_idx = [], []
        for b, f in enumerate(frames):
            f_sk, f_idx = [], []
            for c, p in f.items():
                if c in sm:
                    for i, sk in enumerate(json.loads(p[0])):
                        if any(k != 'ID' for k in sk):          # skeletons without joint keys are not heads (graph_generator.py:590)
                            f_sk.append(sk); f_idx.append(i)
            all_sk.append(f_sk); all_idx.append(f_idx)
        pb.skeletons, pb.skeleton_index = all_sk, all_idx
    else:
        pb.skeletons = None                 # skeleton_index stays: it came with the packed arrays, no Python object was built for it
    return pb
