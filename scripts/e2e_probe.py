"""Host-overhead probe: end-to-end time vs number of chunks, and the CPU cost of enqueueing one step."""
import importlib, json, os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import load_workload, load_weights
pm = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
pack = importlib.import_module('3d_multi_pose_estimator_b200.pack')
cfg, frames = load_workload('panoptic', 1024, 4, 0)
gat, mlp = load_weights('panoptic', cfg)
pb = pack.pack_frames(frames, cfg, keep_json=False)
pipe = pm.PosePipeline(cfg, gat, mlp, device='cuda:0')
hb = pm.HostBatch(pb)
db = hb.to_device('cuda:0')
for _ in range(3):
    pipe.infer(db)
torch.cuda.synchronize()
# CPU time to enqueue stage A (no sync inside)
t0 = time.perf_counter()
for _ in range(5):
    pipe.stage_a(db)
t_cpu = (time.perf_counter() - t0) / 5
torch.cuda.synchronize()
print('stage_a enqueue CPU time %.3f ms' % (1e3 * t_cpu))
for n in (1, 2, 4, 8):
    for _ in range(3):
        pipe.infer_host(hb, n_chunks=n)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe.infer_host(hb, n_chunks=n)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print('n_chunks %d: e2e %.3f ms (min %.3f)' % (n, 1e3 * sum(ts) / len(ts), 1e3 * min(ts)))
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    pipe.infer_host(hb, n_chunks=4)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
