"""Times the aggregation class (and the whole step) of the bench workload: A/B runs of kernel variants selected by agg_impl
(0 dispatch, 6 one-CTA-per-frame specialised, 5 generic frame-resident, 1 gather, 2 large-frame).  python scripts/agg_quick.py [agg_impl [config persons frames]]"""
import importlib, os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import load_workload, load_weights, profile_classes
pm = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
pack = importlib.import_module('3d_multi_pose_estimator_b200.pack')
config = sys.argv[2] if len(sys.argv) > 2 else 'panoptic'
persons = int(sys.argv[3]) if len(sys.argv) > 3 else 4
n_frames = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
cfg, frames = load_workload(config, n_frames, persons, 0)
gat, mlp = load_weights(config, cfg)
pb = pack.pack_frames(frames, cfg, keep_json=False)
pipe = pm.PosePipeline(cfg, gat, mlp, device='cuda:0')
pipe.agg_impl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
pipe.fuse_small_fc2 = os.environ.get('B200POSE_NO_FUSE2') != '1'
db = pm.HostBatch(pb).to_device('cuda:0')
for _ in range(3):
    pipe.infer(db)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for a, b in ev:
    flush.fill_(1); a.record(); pipe.infer(db, sync=False); b.record()
torch.cuda.synchronize()
k = profile_classes(pipe, db, pm, torch)
print(config, persons, 'heads/frame', pb.max_heads, 'fuse2', pipe.fuse_small_fc2, 'impl %d: step %.4f ms | agg %.4f | gemm %.4f | mlp %.4f' % (pipe.agg_impl, np.mean([a.elapsed_time(b) for a, b in ev]),
      k['edge_softmax_aggregate'], k['gat_projection_gemm'], k['mlp_gemm']))
