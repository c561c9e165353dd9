"""Host-side cost of a live-frame submission (3d_multi_pose_estimator_b200/live.py): per-call wall time of submit(), of each
graph replay, and the GPU-side latency to each event."""
import importlib, os, sys, time
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import load_workload, load_weights
pm = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
pack = importlib.import_module('3d_multi_pose_estimator_b200.pack')
live = importlib.import_module('3d_multi_pose_estimator_b200.live')
cfg, frames = load_workload('panoptic', 64, 4, 0)
gat, mlp = load_weights('panoptic', cfg)
pipe = pm.PosePipeline(cfg, gat, mlp, device='cuda:0')
lv = live.LiveFrames(pipe)
lv.model_key = 'x'
pbs = [pack.pack_frames_fast([f], cfg, keep_json=False) for f in frames]
for pb in pbs[:16]:
    h = lv.submit(pb); h.proposals(); h.stage3(); torch.cuda.synchronize()
t_sub, t_a, t_b, t_c = [], [], [], []
for pb in pbs:
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    h = lv.submit(pb)
    t1 = time.perf_counter()
    h.ent.ev_a.synchronize(); t2 = time.perf_counter()
    h.ent.ev_b.synchronize(); t3 = time.perf_counter()
    h.ent.ev_c.synchronize(); t4 = time.perf_counter()
    t_sub.append(t1 - t0); t_a.append(t2 - t0); t_b.append(t3 - t0); t_c.append(t4 - t0)
med = lambda v: 1e3 * float(np.median(v))
print('submit %.3f ms | event A at %.3f | B at %.3f | C at %.3f (host wall time from the start of submit, median of %d frames)' % (
    med(t_sub), med(t_a), med(t_b), med(t_c), len(pbs)))
ent = h.ent
for name in ('g1', 'g2', 'g3'):
    g = getattr(ent, name)
    ts = []
    for _ in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(lv.stream):
            g.replay()
        ts.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
    print(name, 'replay call %.3f ms' % med(ts))
