"""CPU, world_size 2 over gloo: the multi-GPU host logic (contiguous frame blocks per rank, one all-gather of
fixed-size records, frame-order concatenation). The per-rank results come from the oracle here; on the GPU box
bench.py --gpus N drives the same code with the CUDA path."""
import importlib
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers

sharding = importlib.import_module('3d_multi_pose_estimator_b200.sharding')


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 8, 1024, 1000003):
        for world in (1, 2, 3, 4, 8):
            blocks = [sharding.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) == -(-n // world) and all(s >= 0 for s in sizes)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_results(frames, cfg):
    from oracle import pose_oracle as O
    tabs = O.CameraTables(cfg)
    gat, mlp = helpers.golden_weights('panoptic')
    gw, mw = helpers.np_state(gat), helpers.np_state(mlp)
    pack = importlib.import_module('3d_multi_pose_estimator_b200.pack')
    pb = pack.pack_frames(frames, cfg)
    n_persons, sk_rows, joints = [], [], []
    for b, f in enumerate(frames):
        out = O.infer_frame(f, tabs, gw, mw)
        if out is None:
            n_persons.append(0)
            continue
        n_persons.append(len(out['proposals']))
        for person in out['proposals']:
            row = np.full(cfg.n_cameras, -1, np.int32)
            for s, h in enumerate(person):
                if h >= 0:
                    row[cfg.used_sm[s]] = pb.head_off[b] + h
            sk_rows.append(row)
        joints.append(out['joints'])
    return (np.array(n_persons, np.int32), np.stack(sk_rows) if sk_rows else np.zeros((0, cfg.n_cameras), np.int32),
            np.concatenate(joints) if joints else np.zeros((0, 54), np.float32), pb)


def _worker(rank, world, port, tags, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        cfg, npz, meta = helpers.load_golden('panoptic')
        frames = [{c: v for c, v in meta['frames'][t].items() if json.loads(v[0])} for t in tags]
        lo, hi = sharding.shard_range(len(frames), rank, world)
        pack = importlib.import_module('3d_multi_pose_estimator_b200.pack')
        head_base = int(pack.pack_frames(frames, cfg, keep_json=False).head_off[lo])
        n_persons, person_sk, joints, _ = _oracle_results(frames[lo:hi], cfg)
        frames_cap, persons_cap = -(-len(frames) // world), 64
        rec = sharding.pack_record(torch.from_numpy(n_persons), torch.from_numpy(person_sk), torch.from_numpy(joints), frames_cap,
                                   persons_cap, cfg.n_cameras, 54, head_base=head_base)
        gathered = sharding.all_gather_records(rec, world)
        per_rank = [sharding.shard_range(len(frames), r, world) for r in range(world)]
        merged = sharding.unpack_records(gathered, [b - a for a, b in per_rank], frames_cap, persons_cap, cfg.n_cameras, 54)
        if rank == 0:
            ret.put({k: v.tolist() for k, v in merged.items()})
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    cfg, npz, meta = helpers.load_golden('panoptic')
    tags = ['p3', 'p4a', 'onecam', 'rag1', 'm1']                 # ragged: one frame without a graph, uneven split 3 + 2
    ctx = mp.get_context('spawn')
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, tags, ret)) for r in range(2)]
    for p in procs:
        p.start()
    merged = ret.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    frames = [{c: v for c, v in meta['frames'][t].items() if json.loads(v[0])} for t in tags]
    n_persons, person_sk, joints, pb = _oracle_results(frames, cfg)
    assert merged['n_persons'] == n_persons.tolist()
    assert merged['person_off'] == np.concatenate([[0], np.cumsum(n_persons)]).tolist()
    assert merged['person_sk'] == person_sk.tolist()
    assert np.array_equal(np.array(merged['joints'], np.float32), joints)       # records carry raw fp32 bits


def test_record_roundtrip_single_rank():
    n_persons = torch.tensor([2, 0, 1], dtype=torch.int32)
    person_sk = torch.tensor([[0, -1, 3], [1, 2, -1], [5, 6, 7]], dtype=torch.int32)
    joints = torch.randn(3, 54)
    rec = sharding.pack_record(n_persons, person_sk, joints, 4, 8, 3, 54, head_base=10)
    out = sharding.unpack_records(sharding.all_gather_records(rec, 1), [3], 4, 8, 3, 54)
    assert out['n_persons'].tolist() == [2, 0, 1] and out['person_off'].tolist() == [0, 2, 2, 3]
    assert out['person_sk'].tolist() == [[10, -1, 13], [11, 12, -1], [15, 16, 17]]
    assert np.array_equal(out['joints'], joints.numpy())
    with pytest.raises(ValueError):
        sharding.pack_record(n_persons, person_sk, joints, 2, 8, 3, 54)
