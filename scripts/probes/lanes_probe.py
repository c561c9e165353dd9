"""Do two compute lanes help? K device-resident 1024-frame steps on one stream vs alternating between two streams (two pipelines
sharing the weight planes, each with its own workspaces): total time per step."""
import importlib, os, sys, time, copy
import torch
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from bench import load_workload, load_weights
pm = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
pack = importlib.import_module('3d_multi_pose_estimator_b200.pack')
cfg, frames = load_workload('panoptic', 1024, 4, 0)
gat, mlp = load_weights('panoptic', cfg)
pb = pack.pack_frames(frames, cfg, keep_json=False)
dev = torch.device('cuda:0')
pipe = pm.PosePipeline(cfg, gat, mlp, device=dev)
twin = pm.PosePipeline(cfg, gat, mlp, device=dev)
db = pm.HostBatch(pb).to_device(dev)
K = 20
def run(lanes):
    streams = [torch.cuda.Stream(dev) for _ in lanes]
    for _ in range(3):
        for p, s in zip(lanes, streams):
            with torch.cuda.stream(s):
                p.infer(db, sync=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        p, s = lanes[i % len(lanes)], streams[i % len(lanes)]
        with torch.cuda.stream(s):
            p.infer(db, sync=False)
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / K
for rep in range(2):
    print('one lane  %.3f ms/step' % run([pipe]))
    print('two lanes %.3f ms/step' % run([pipe, twin]))
third = pm.PosePipeline(cfg, gat, mlp, device=dev)
print('three lanes %.3f ms/step' % run([pipe, twin, third]))
