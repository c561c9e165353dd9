"""Frame sharding across the GPUs of one box.

Frames are independent (the reference keeps no cross-frame state, test/metrics_from_model.py:120-300), so a batch
is cut into contiguous blocks of frames, one block per rank, with weights and camera tables replicated. There is no
collective on the hot path; the only exchange is the gather of fixed-size result records, issued asynchronously behind
each step (ResultGather: one packing kernel + one NCCL all-gather on NCCL's own stream) and waited for once, at the end
of the job.
The record layout is plain int32 words so it travels through NCCL (device tensors) and gloo (CPU tensors, tests)
alike: [n_persons_total | n_persons[F] | person_sk[Pcap, C] | joints[Pcap, J] as raw fp32 bits].
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch


def shard_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `rank`: ceil(n_frames / world) frames per rank, the tail ranks may be short/empty."""
    per = -(-n_frames // world) if world > 0 else n_frames
    lo = min(n_frames, rank * per)
    return lo, min(n_frames, lo + per)


def record_words(frames_cap: int, persons_cap: int, n_cameras: int, n_out: int) -> int:
    return 1 + frames_cap + persons_cap * n_cameras + persons_cap * n_out


def pack_record(n_persons: torch.Tensor, person_sk: torch.Tensor, joints: torch.Tensor, frames_cap: int, persons_cap: int,
                n_cameras: int, n_out: int, head_base: int = 0) -> torch.Tensor:
    """One rank's results as a fixed-size int32 record (padded with -1 / 0). `head_base` turns the rank-local
    skeleton indices of person_sk into indices of the whole batch."""
    dev = n_persons.device
    F, P = int(n_persons.numel()), int(person_sk.shape[0])
    if F > frames_cap or P > persons_cap:
        raise ValueError('result does not fit the record (%d/%d frames, %d/%d persons)' % (F, frames_cap, P, persons_cap))
    # padding words are never read back (unpack_records slices by the counts), so the record is not cleared
    rec = torch.empty(record_words(frames_cap, persons_cap, n_cameras, n_out), dtype=torch.int32, device=dev)
    rec[0] = P
    rec[1:1 + F] = n_persons
    o = 1 + frames_cap
    if P:
        sk = rec[o:o + P * n_cameras].view(P, n_cameras)
        sk.copy_(person_sk)
        if head_base:
            sk.add_((sk >= 0).to(torch.int32) * int(head_base))
    o += persons_cap * n_cameras
    if P and n_out:
        rec[o:o + P * n_out].view(torch.float32).view(P, n_out).copy_(joints)
    return rec


def pack_record_device(res: dict, n_frames: int, frames_cap: int, persons_cap: int, n_cameras: int, n_out: int,
                       out: torch.Tensor, head_base: int = 0, stream=None) -> torch.Tensor:
    """pack_record for results resident on the GPU, as ONE kernel launch (b200pose_pack_record): no eager tensor ops, no
    host-side person count - it reads person_off[n_frames] on the device - so it can be enqueued right behind the step that
    produced `res` (PosePipeline.infer). `out`: int32 [record_words] device buffer."""
    from . import _lib
    joints = res.get('joints')
    ld = int(joints.stride(0)) if joints is not None and joints.numel() else n_out
    _lib.check(_lib.lib().b200pose_pack_record(n_frames, n_cameras, n_out, _lib.ptr(res['n_persons']), _lib.ptr(res['person_off']),
                                               _lib.ptr(res['person_sk']),
                                               _lib.ptr(joints) if joints is not None and joints.numel() else None, ld,
                                               frames_cap, persons_cap, head_base, _lib.ptr(out), stream), 'pack_record')
    return out


class ResultGather:
    """The job's one exchange, taken off the critical path: every step's record is packed by one kernel behind the step and
    all-gathered asynchronously (NCCL's own stream waits for the packing kernel; the compute stream never waits for the
    collective), `depth` steps in flight. finish() waits for what is outstanding - the "final gather" of SURVEY.md 8e."""

    def __init__(self, world: int, frames_cap: int, persons_cap: int, n_cameras: int, n_out: int, device, depth: int = 4, group=None):
        self.world, self.group, self.depth = world, group, depth
        self.dims = (frames_cap, persons_cap, n_cameras, n_out)
        words = record_words(*self.dims)
        self.rec = [torch.empty(words, dtype=torch.int32, device=device) for _ in range(depth)]
        self.out = [torch.empty(world * words, dtype=torch.int32, device=device) for _ in range(depth)]
        self.work = [None] * depth
        self.n = 0

    def submit(self, res: dict, n_frames: int, head_base: int = 0, stream=None):
        import torch.distributed as dist
        k = self.n % self.depth
        if self.work[k] is not None:
            self.work[k].wait()                      # stream-level: the slot's previous gather must have read its record
            self.work[k] = None
        frames_cap, persons_cap, n_cameras, n_out = self.dims
        pack_record_device(res, n_frames, frames_cap, persons_cap, n_cameras, n_out, self.rec[k], head_base, stream)
        if self.world > 1:
            self.work[k] = dist.all_gather_into_tensor(self.out[k], self.rec[k], group=self.group, async_op=True)
        self.n += 1
        return k

    def finish(self):
        """Waits for the outstanding gathers; returns the gathered records [world, words] of the last `depth` steps, oldest first."""
        for w in self.work:
            if w is not None:
                w.wait()
        self.work = [None] * self.depth
        ks = [(i % self.depth) for i in range(max(0, self.n - self.depth), self.n)]
        if self.world == 1:
            return [self.rec[k].reshape(1, -1) for k in ks]
        return [self.out[k].reshape(self.world, -1) for k in ks]


def all_gather_records(rec: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """The single collective of a step: [world, words] records of every rank (identity for world == 1)."""
    if world == 1:
        return rec.reshape(1, -1)
    import torch.distributed as dist
    out = torch.empty(world * rec.numel(), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec, group=group)
    return out.reshape(world, -1)


def unpack_records(gathered: torch.Tensor, frames_per_rank: List[int], frames_cap: int, persons_cap: int, n_cameras: int,
                   n_out: int) -> Dict[str, np.ndarray]:
    """Concatenates the ranks' results in frame order: n_persons[F_total], person_off[F_total+1], person_sk[P_total, C],
    joints[P_total, n_out]."""
    g = gathered.cpu()
    n_persons, sks, joints = [], [], []
    for r, F in enumerate(frames_per_rank):
        rec = g[r]
        P = int(rec[0])
        n_persons.append(rec[1:1 + F].numpy())
        o = 1 + frames_cap
        sks.append(rec[o:o + persons_cap * n_cameras].reshape(persons_cap, n_cameras)[:P].numpy())
        o += persons_cap * n_cameras
        joints.append(rec[o:o + P * n_out].contiguous().view(torch.float32).reshape(P, n_out).numpy())
    n_persons = np.concatenate(n_persons) if n_persons else np.zeros(0, np.int32)
    return dict(n_persons=n_persons, person_off=np.concatenate([[0], np.cumsum(n_persons)]).astype(np.int32),
                person_sk=np.concatenate(sks) if sks else np.zeros((0, n_cameras), np.int32),
                joints=np.concatenate(joints) if joints else np.zeros((0, n_out), np.float32))
