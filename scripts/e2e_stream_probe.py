"""Where a streamed end-to-end step (PosePipeline.infer_host_stream) spends its time: wall time per batch over a long
run, against the device-resident step, plus the host time of the enqueue / finalize halves."""
import importlib, os, sys, time
import numpy as np
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
    pipe.infer(db, sync=False)
for _ in pipe.infer_host_stream([hb] * 3):
    pass
torch.cuda.synchronize()
K = 100
t0 = time.perf_counter()
for _ in range(K):
    pipe.infer(db, sync=False)
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_dev = time.perf_counter() - t0
print('device-resident, %d steps back to back: %.3f ms per step (host enqueue time %.3f ms per step)' % (K, 1e3 * t_dev / K, 1e3 * t_host / K))
t0 = time.perf_counter()
gaps = []
last = t0
for out in pipe.infer_host_stream([hb] * K):
    now = time.perf_counter(); gaps.append(now - last); last = now
torch.cuda.synchronize()
t_e2e = time.perf_counter() - t0
print('streamed end to end, %d batches: %.3f ms per batch; yield-to-yield median %.3f ms, p90 %.3f' % (K, 1e3 * t_e2e / K, 1e3 * np.median(gaps[3:]), 1e3 * np.percentile(gaps[3:], 90)))
# H2D alone
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    d = hb.to_device('cuda:0')
torch.cuda.synchronize()
print('H2D of one batch alone: %.3f ms' % (1e3 * (time.perf_counter() - t0) / 20))
