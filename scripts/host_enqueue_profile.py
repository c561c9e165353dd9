import importlib, os, sys, time, cProfile, pstats, io
import torch
sys.path.insert(0, '/root/repo')
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
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    pipe.infer(db, sync=False)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(28); print(s.getvalue()[:6000])
