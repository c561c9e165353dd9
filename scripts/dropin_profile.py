"""cProfile of the reference driver's per-frame loop on the drop-in modules (where do the milliseconds per frame go)."""
import cProfile, io, os, pstats, sys, contextlib
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import bench, dropin_env
from test_dropin_gpu import run_frame
cfg, frames = bench.load_workload('panoptic', 40, 4, 0)
gat, mlp_state = bench.load_weights('panoptic', cfg)
mods = dropin_env.activate(cfg)
dev = torch.device('cuda')
torch.set_grad_enabled(False)   # test/metrics_from_model.py:54
model = mods['gat2'].GAT2(None, 5, cfg.n_features_sm, 1, [40, 40, 40, 30], [10, 10, 8, 5], torch.nn.LeakyReLU(), torch.nn.Sigmoid(), 0., 0., 0.15, False, bias=True)
model.load_state_dict(gat); model = model.to(dev)
mlp = mods['mlp'].PoseEstimatorMLP(input_dimensions=cfg.n_cameras * 18 * 14, output_dimensions=54)
mlp.load_state_dict(mlp_state); mlp = mlp.to(dev)
with contextlib.redirect_stdout(io.StringIO()):
    for f in frames[:8]:
        run_frame(mods, cfg, model, mlp, f)
    pr = cProfile.Profile()
    pr.enable()
    for f in frames:
        run_frame(mods, cfg, model, mlp, f)
    torch.cuda.synchronize()
    pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
print(s.getvalue()[:9000])
