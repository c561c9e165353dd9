#!/usr/bin/env python
"""Benchmark of the per-frame inference hot path (BASELINE.json: frames/sec, 5-view, N persons).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

One step = one pass of the whole path (graph build -> GAT -> clustering -> encoder -> MLP) over one batch
of synthetic frames: Panoptic 5-view, 1024 frames, 4 persons/frame (BASELINE.json configs[1]).
  value : frames/s with the packed inputs already resident in HBM (CUDA events, max over ranks)
  e2e   : frames/s through the public host API (pinned host buffers -> results on the host)
Frames are independent, so N GPUs each process their own 1024 frames (weak scaling); when N > 1 the
per-frame results are gathered with one NCCL all_gather per step.
`--impl reference` times the CPU port of the reference path (oracle/) on a bounded sample instead.
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, 'tests', 'golden')

METRIC = 'frames/sec (5-view, 4 persons/frame, 1024-frame batch)'
UNIT = 'frames/s'


def load_workload(config, n_frames, n_persons, seed0):
    pkg = importlib.import_module('3d_multi_pose_estimator_b200')
    synth = importlib.import_module('3d_multi_pose_estimator_b200.synth')
    cfg = pkg.CameraConfig.from_npz(os.path.join(GOLDEN, 'cameras_%s.npz' % config))
    frames = synth.make_frames(cfg, n_frames, n_persons, base_seed=seed0)
    frames = [{c: f[c] for c in f if json.loads(f[c][0])} for f in frames]
    return cfg, frames


def load_weights(config, cfg):
    """Random-init weights of the reference architectures (seeded) + the stored last-layer calibration."""
    import torch
    W = importlib.import_module('3d_multi_pose_estimator_b200.weights')
    meta = json.load(open(os.path.join(GOLDEN, 'golden_%s.json' % config)))
    npz = np.load(os.path.join(GOLDEN, 'golden_%s.npz' % config))
    gat = W.make_gat_state(cfg.n_features_sm, meta['gat_seed'])
    gat['layers.4.fc2.weight'] = torch.from_numpy(npz['gat_last_fc2_weight'].copy())
    gat['layers.4.fc2.bias'] = torch.from_numpy(npz['gat_last_fc2_bias'].copy())
    mlp = W.make_mlp_state(cfg.n_cameras * 18 * 14, 54, meta['mlp_seed'])
    return gat, mlp


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap'}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


def cpu_reference_run(cfg, frames, gat_state, mlp_state, budget_s, max_frames=None):
    """The CPU port of the reference path (oracle/pose_oracle.py), one frame at a time like the reference's
    drivers, torch-free numpy with all BLAS threads. Returns (frames/s, frames processed, seconds)."""
    from oracle import pose_oracle as O
    tabs = O.CameraTables(cfg)
    gw = {k: v.numpy() for k, v in gat_state.items()}
    mw = {k: v.numpy() for k, v in mlp_state.items()}
    n, t0 = 0, time.perf_counter()
    for f in frames:
        O.infer_frame(f, tabs, gw, mw)
        n += 1
        if time.perf_counter() - t0 > budget_s or (max_frames and n >= max_frames):
            break
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def algorithmic_work(pb, pipe_gat_dims, mlp_dims, n_persons):
    """Algorithmic bytes / flops per step: edge-softmax compulsory bytes and executed GEMM flops."""
    N, E, S = pb.n_nodes, pb.n_edges, pb.n_heads
    agg_bytes = 0
    gemm_flops = 0
    for l, (din, h, d) in enumerate(pipe_gat_dims):
        rows = (S + 1) if l == 0 else N
        hd = h * d
        gemm_flops += 2 * rows * (din * din + din * (hd + 2 * h))
        agg_bytes += (rows * (hd + 2 * h) + N * hd) * 4 + E * 4 + (N + 1) * 4
    mlp_flops = 0
    for k, n in mlp_dims:
        mlp_flops += 2 * n_persons * k * n
    return agg_bytes, gemm_flops, mlp_flops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--frames', type=int, default=1024)
    ap.add_argument('--persons', type=int, default=4)
    ap.add_argument('--config', default='panoptic')
    ap.add_argument('--cpu-budget', type=float, default=15.0, help='seconds of CPU work for the cpu_baseline sample')
    ap.add_argument('--gemm-impl', type=int, default=0)
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    config_desc = {'workload': 'Panoptic 5-view, %d-frame batch, %d persons/frame, skeleton matching + MLP lift'
                   % (args.frames, args.persons), 'camera_config': args.config, 'frames_per_gpu': args.frames,
                   'persons_per_frame': args.persons, 'weights': 'random-init (seeded) + last-layer bias calibration',
                   'l2': 'L2 flushed (256 MiB write) between timed iterations'}

    # ------------------------------------------------------------------ reference arm (CPU port)
    if args.impl == 'reference':
        if rank != 0:
            return
        import torch
        cfg, frames = load_workload(args.config, min(args.frames, 64), args.persons, 0)
        gat, mlp = load_weights(args.config, cfg)
        per_step = []
        sample_frames = 0
        for i in range(args.warmup + args.steps):
            fps, n, dt = cpu_reference_run(cfg, frames, gat, mlp, budget_s=max(2.0, 60.0 / max(1, args.steps + args.warmup)), max_frames=16)
            if i >= args.warmup:
                per_step.append(dt / n)
                sample_frames = n
        ms_frame = 1e3 * float(np.mean(per_step))
        value = 1e3 / ms_frame
        cores = os.cpu_count()
        line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_frame * args.frames, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config_desc,
                'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                 'sample': '%d frames per step, one frame at a time (numpy/BLAS, %d host threads available); '
                                           'ms_per_step extrapolated to the %d-frame batch' % (sample_frames, cores, args.frames)},
                'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    pm = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
    pack = importlib.import_module('3d_multi_pose_estimator_b200.pack')
    W = importlib.import_module('3d_multi_pose_estimator_b200.weights')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cfg, frames = load_workload(args.config, args.frames, args.persons, seed0=rank * args.frames)
    gat, mlp = load_weights(args.config, cfg)
    pb = pack.pack_frames(frames, cfg, keep_json=False)
    hb = pm.HostBatch(pb)
    pipe = pm.PosePipeline(cfg, gat, mlp, device=dev, gemm_impl=args.gemm_impl)
    db = hb.to_device(dev)
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    P_cap = pb.n_heads // 2 + 1

    def gather_results(res):
        if world == 1:
            return
        rec = torch.zeros(args.frames + P_cap * 54, dtype=torch.float32, device=dev)
        rec[:args.frames] = res['n_persons'].float()
        j = res['joints'].reshape(-1)
        rec[args.frames:args.frames + j.numel()] = j
        out = torch.empty(world * rec.numel(), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(out, rec)

    def step_device():
        res = pipe.infer(db)
        gather_results(res)
        return res

    def step_host():
        out = pipe.infer_host(hb)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (also settles the caching allocator)
    for _ in range(warmup):
        res = step_device()
        step_host()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- device-resident timing: one event pair per step, L2 flushed in between
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    pipe.launches = 0
    barrier()
    for a, b in ev:
        flush.fill_(1)
        a.record()
        res = step_device()
        b.record()
    barrier()
    launches = pipe.launches // max(1, args.steps)
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    # ---- end-to-end through the host API
    barrier()
    e2e_t = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = step_host()
        torch.cuda.synchronize()
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    e2e_ms = 1e3 * float(np.mean(e2e_t))
    d2h = sum(v.numel() * v.element_size() for v in out.values() if hasattr(v, 'numel'))
    # ---- per-kernel-class timing for the roofline (separate pass, CUDA events around each class)
    kern = profile_classes(pipe, db, pm, torch) if rank == 0 else None
    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_max = float(t[0]), float(t[1])
    total_frames = args.frames * world
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        tc_peak = peaks.get('bf16_tflops_sustained', 1400.0)
        peak_src = 'measured' if peaks else 'fallback'
        P = res['n_persons_total']
        gat_dims = W.gat_layer_dims(cfg.n_features_sm)
        mlp_dims = [(l['k'], l['n']) for l in pipe.mlp]
        agg_bytes, gat_flops, mlp_flops = algorithmic_work(pb, gat_dims, mlp_dims, P)
        kernels = []
        for name, k_ms in kern.items():
            entry = {'kernel': name, 'ms_per_step': k_ms}
            if name == 'gat_projection_gemm':
                entry.update(bound='tensor', achieved=3 * gat_flops / k_ms / 1e9, peak=tc_peak, unit='TFLOP/s')
            elif name == 'mlp_gemm':
                entry.update(bound='tensor', achieved=3 * mlp_flops / k_ms / 1e9, peak=tc_peak, unit='TFLOP/s')
            elif name == 'edge_softmax_aggregate':
                entry.update(bound='hbm', achieved=agg_bytes / k_ms / 1e6, peak=hbm_peak, unit='GB/s')
            if 'achieved' in entry:
                entry['frac'] = entry['achieved'] / entry['peak']
            kernels.append(entry)
        dominant = max([k for k in kernels if 'frac' in k], key=lambda k: k['ms_per_step'])
        roofline = {'kernel': dominant['kernel'], 'bound': dominant['bound'], 'achieved': dominant['achieved'],
                    'peak': dominant['peak'], 'unit': dominant['unit'], 'frac': dominant['frac'], 'traffic': None,
                    'peak_source': peak_src + (' (bf16 sustained; achieved counts the 3 executed split-bf16 MMAs)'
                                               if dominant['bound'] == 'tensor' else '')}
        cpu_fps, cpu_n, cpu_dt = cpu_reference_run(cfg, frames, gat, mlp, budget_s=args.cpu_budget)
        line = {'metric': METRIC, 'value': total_frames / ms_max * 1e3, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': warmup, 'ms_per_step': ms_max, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'bf16x3 split (fp32-accurate), fp32 accumulate; fp64 geometry', 'data': 'synthetic',
                'config': config_desc,
                'e2e': {'value': total_frames / e2e_max * 1e3, 'unit': UNIT, 'h2d_bytes_per_step': hb.nbytes(),
                        'd2h_bytes_per_step': int(d2h), 'ms_per_step': e2e_max},
                'gpu_launches': launches, 'clocks': sampler.summary(), 'roofline': roofline, 'kernels': kernels,
                'cpu_baseline': {'value': cpu_fps, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
                                 'sample': 'first %d frames of the same batch, one at a time, %.1f s (numpy/BLAS threads = host cores)' % (cpu_n, cpu_dt)},
                'persons_found_per_frame': P / args.frames,
                'p50_frame_latency_ms': None}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def profile_classes(pipe, db, pm, torch):
    """ms per step spent in each kernel class, measured live with CUDA events on the launching stream."""
    import collections
    acc = collections.OrderedDict()
    orig_linear, orig_agg = pipe.linear, pipe.aggregate
    orig = {n: getattr(pipe, n) for n in ('build_graph', 'head_feature_planes', 'cluster', 'gather_persons', 'encode_persons')}
    state = {'phase': 'gat'}

    def timed(name, fn):
        def wrapper(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            acc.setdefault(name() if callable(name) else name, []).append((e0, e1))
            return r
        return wrapper
    pipe.linear = timed(lambda: 'gat_projection_gemm' if state['phase'] == 'gat' else 'mlp_gemm', orig_linear)
    pipe.aggregate = timed('edge_softmax_aggregate', orig_agg)
    names = {'build_graph': 'graph_build', 'head_feature_planes': 'node_features', 'cluster': 'cluster',
             'gather_persons': 'gather_persons', 'encode_persons': 'encode_dlt'}
    for n, f in orig.items():
        setattr(pipe, n, timed(names[n], f))
    orig_mlp = pipe.mlp_forward

    def mlp_wrapper(*a, **k):
        state['phase'] = 'mlp'
        r = orig_mlp(*a, **k)
        state['phase'] = 'gat'
        return r
    pipe.mlp_forward = mlp_wrapper
    reps = 3
    for _ in range(reps):
        pipe.infer(db)
    torch.cuda.synchronize()
    out = {k: sum(a.elapsed_time(b) for a, b in v) / reps for k, v in acc.items()}
    pipe.linear, pipe.aggregate, pipe.mlp_forward = orig_linear, orig_agg, orig_mlp
    for n, f in orig.items():
        setattr(pipe, n, f)
    return out


if __name__ == '__main__':
    main()
