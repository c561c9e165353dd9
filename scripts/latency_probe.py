"""Single-frame latency breakdown: wall time of every stage of one-frame infer() calls, and a cProfile of the host side."""
import importlib, os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import load_workload, load_weights
pm = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
pack = importlib.import_module('3d_multi_pose_estimator_b200.pack')
cfg, frames = load_workload('panoptic', 32, 4, 0)
gat, mlp = load_weights('panoptic', cfg)
pipe = pm.PosePipeline(cfg, gat, mlp, device='cuda:0')
pb = pack.pack_frames(frames, cfg, keep_json=False)
singles = [pm.HostBatch(pb.slice(i, i + 1)) for i in range(32)]
for i in range(40):
    pipe.infer_host(singles[i % 32])
sync = torch.cuda.synchronize
acc = {}
def timed(name, fn):
    def w(*a, **k):
        sync(); t0 = time.perf_counter(); r = fn(*a, **k); sync(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
    return w
for n in ('build_graph', 'head_feature_planes', 'linear', 'aggregate', 'cluster', 'encode_persons', 'mlp_forward'):
    setattr(pipe, n, timed(n, getattr(pipe, n)))
N = 100
t0 = time.perf_counter()
for i in range(N):
    pipe.infer_host(singles[i % 32])
tot = time.perf_counter() - t0
print('with per-stage syncs: %.3f ms per frame' % (1e3 * tot / N))
for k, v in acc.items():
    print('  %-22s %.3f ms' % (k, 1e3 * v / N))

# ---- un-instrumented: CPU enqueue time vs GPU busy time of one frame
pipe2 = pm.PosePipeline(cfg, gat, mlp, device='cuda:0')
dbs = [h.to_device('cuda:0') for h in singles]
for i in range(40):
    pipe2.infer(dbs[i % 32])
sync()
import numpy as np
cpu_a, gpu_a, wall = [], [], []
for i in range(100):
    db = dbs[i % 32]
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    st = pipe2.stage_a(db)
    e1.record()
    t1 = time.perf_counter()
    sync()
    cpu_a.append(t1 - t0); gpu_a.append(e0.elapsed_time(e1) * 1e-3)
    t0 = time.perf_counter(); pipe2.stage_b(db, st); sync(); wall.append(time.perf_counter() - t0)
print('stage A: CPU enqueue %.3f ms, GPU span %.3f ms; stage B wall %.3f ms' % (1e3 * np.median(cpu_a), 1e3 * np.median(gpu_a), 1e3 * np.median(wall)))
lat = []
for i in range(200):
    sync(); t0 = time.perf_counter(); pipe2.infer_host(singles[i % 32]); lat.append(time.perf_counter() - t0)
print('infer_host p50 %.3f ms' % (1e3 * np.median(lat)))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(100):
    pipe2.infer_host(singles[i % 32])
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
