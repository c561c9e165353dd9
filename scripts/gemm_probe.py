"""Attribution probe for the persistent GEMM: times the GAT-shaped and MLP-shaped projections with parts of the kernel
switched off (b200pose_set_debug), single CTAs vs CTA pairs."""
import importlib, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
pm = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
L = importlib.import_module('3d_multi_pose_estimator_b200._lib').lib()
ptr, check = pm.ptr, pm.check
dev = 'cuda:0'
stream = None
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

def run(m, n, k, planes_out, impl, dbg, reps=5):
    A = pm.Planes(m, k, dev); W = pm.Planes(n, k, dev)
    A.hi.normal_(); A.lo.normal_(std=1e-3); W.hi.normal_(std=0.05); W.lo.normal_(std=1e-4)
    bias = torch.zeros(n, device=dev)
    out_p = pm.Planes(m, n, dev) if planes_out else None
    ldo = (n + 3) // 4 * 4
    out_f = None if planes_out else torch.empty((m, ldo), device=dev)
    s = torch.cuda.current_stream().cuda_stream
    L.b200pose_set_debug(dbg)
    ts = []
    for i in range(reps + 2):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(L.b200pose_linear(ptr(A.hi), ptr(A.lo), A.ld, ptr(W.hi), ptr(W.lo), W.ld, ptr(bias), m, n, k, 0.15, 1.0,
                                ptr(out_f), ldo if out_f is not None else 0, ptr(out_p.hi) if out_p else None,
                                ptr(out_p.lo) if out_p else None, out_p.ld if out_p else 0, impl, s), 'linear')
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    L.b200pose_set_debug(0)
    return sum(ts) / len(ts)

shapes = [('gat fc1 400x400 planes', 184320, 400, 400, True), ('gat fc2 420x400 f32', 184320, 420, 400, False),
          ('gat fc1 150x150 planes', 184320, 150, 150, True), ('mlp 3072x3072 planes', 3638, 3072, 3072, True),
          ('mlp 1024x1024 planes', 3638, 1024, 1024, True)]
for name, m, n, k, pl in shapes:
    flops = 3 * 2.0 * m * n * k
    for impl in (4, 5):
        row = []
        for dbg in (0, 1, 4, 5, 2, 7):
            t = run(m, n, k, pl, impl, dbg)
            row.append('dbg%d %.1fus' % (dbg, 1e3 * t))
        t0 = run(m, n, k, pl, impl, 0)
        print('%-26s impl %d  %s   | full: %.0f TFLOP/s executed' % (name, impl, '  '.join(row), flops / t0 / 1e9))
