"""One profiled step of the bench workload (for ncu): 2 warm steps, then cudaProfilerStart .. Stop around one."""
import importlib, json, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import load_workload, load_weights
pm = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
pack = importlib.import_module('3d_multi_pose_estimator_b200.pack')
frames_n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
config = sys.argv[2] if len(sys.argv) > 2 else 'panoptic'
persons = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cfg, frames = load_workload(config, frames_n, persons, 0)
gat, mlp = load_weights(config, cfg)
pb = pack.pack_frames(frames, cfg, keep_json=False)
pipe = pm.PosePipeline(cfg, gat, mlp, device='cuda:0')
db = pm.HostBatch(pb).to_device('cuda:0')
for _ in range(2):
    pipe.infer(db)
torch.cuda.synchronize()
torch.cuda.profiler.start()
res = pipe.infer(db)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('persons', res['n_persons_total'])
