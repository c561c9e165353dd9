"""Runs one of the reference's OWN driver scripts (test/metrics_from_model.py, show_results_from_model.py, ...) exactly as
its README says - cwd = <reference>/test, `python <script> <args>` - either on the reference's modules (mode 'reference':
CPU, with the dgl / pytransform3d shims of oracle/shims) or on the B200 drop-in (mode 'dropin': the shadow directory first
on sys.path, which is all a user changes). The script file is executed unmodified through runpy.

While it runs, observers record what goes through three functions the script imports - the person proposals
(skeleton_matching_utils.get_person_proposal_from_network_output), the 3D outputs (mlp.PoseEstimatorMLP.forward) and the
triangulations (pose_estimator_utils.triangulate) - plus stdout and, for the viewers, the headless pyqtgraph scene. The
observers wrap the module attributes BEFORE the script's `from X import f`, so they see every call and change no result.

    python tests/drivers/run_driver.py --mode dropin --refroot baseline/_ref --script metrics_from_model.py \
        --record out.json -- --testfiles f.json --tmdir d --modelsdir m --datastep 1
"""
import argparse
import io
import json
import os
import runpy
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
REPO = os.path.dirname(TESTS)
SHADOW = os.path.join(REPO, '3d_multi_pose_estimator_b200', 'shadow')


class Tee(io.TextIOBase):
    def __init__(self, real):
        self.real, self.buf = real, io.StringIO()

    def write(self, s):
        self.buf.write(s)
        return self.real.write(s)

    def flush(self):
        self.real.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mode', choices=['reference', 'dropin'], required=True)
    ap.add_argument('--refroot', required=True)
    ap.add_argument('--script', required=True)
    ap.add_argument('--record', required=True)
    ap.add_argument('--seed', type=int, default=0, help='seeds random / numpy / torch before the script starts (the training-side loaders draw from the global generators)')
    ap.add_argument('rest', nargs=argparse.REMAINDER)
    a = ap.parse_args()
    rest = a.rest[1:] if a.rest and a.rest[0] == '--' else a.rest
    refroot = os.path.abspath(a.refroot)
    record_path = os.path.abspath(a.record)
    stubs = os.path.join(HERE, 'stubs')
    if a.mode == 'dropin':
        first = [SHADOW, stubs]                    # what a user does: PYTHONPATH=<repo>/3d_multi_pose_estimator_b200/shadow
    else:
        first = [os.path.join(REPO, 'oracle', 'shims'), stubs]
    for p in reversed(first):
        sys.path.insert(0, p)
    os.chdir(os.path.join(refroot, 'test'))
    for p in ('../skeleton_matching', '../utils', '../'):      # the scripts append exactly these (metrics_from_model.py:12,17,23)
        sys.path.append(p)
    import numpy as np
    import torch

    rec = dict(mode=a.mode, script=a.script, argv=rest, proposals=[], scores=[], mlp_out=[], triangulate=[], drawn=[])
    import skeleton_matching_utils as smu
    import mlp as mlp_mod
    import pose_estimator_utils as pu
    rec['modules'] = {m.__name__: os.path.relpath(m.__file__, REPO) for m in (smu, mlp_mod, pu)}
    gpp = smu.get_person_proposal_from_network_output

    def gpp_obs(outputs, subgraph, indices, nodes_camera, *args, **kw):
        res = gpp(outputs, subgraph, indices, nodes_camera, *args, **kw)
        sc = outputs.detach().cpu().reshape(-1).tolist() if hasattr(outputs, 'detach') else [float(x) for x in outputs]
        rec['scores'].append(sc)
        rec['proposals'].append([{c: (None if h is None else int(h)) for c, h in person.items()} for person in res])
        return res
    smu.get_person_proposal_from_network_output = gpp_obs
    fwd = mlp_mod.PoseEstimatorMLP.forward

    def fwd_obs(self, x):
        y = fwd(self, x)
        rec['mlp_out'].append(y.detach().cpu().float().numpy().tolist())
        return y
    mlp_mod.PoseEstimatorMLP.forward = fwd_obs
    tri = pu.triangulate

    def tri_obs(*args, **kw):
        r = tri(*args, **kw)
        rec['triangulate'].append({str(j): np.asarray(v, dtype=np.float64).reshape(-1).tolist() for j, v in r.items()})
        return r
    pu.triangulate = tri_obs

    import random
    random.seed(a.seed); np.random.seed(a.seed); torch.manual_seed(a.seed)
    tee = Tee(sys.stdout)
    sys.stdout = tee
    sys.argv = [a.script] + rest
    code = 0
    t0 = time.time()
    try:
        runpy.run_path(os.path.join(refroot, 'test', a.script), run_name='__main__')
    except SystemExit as e:
        code = e.code if isinstance(e.code, int) else (0 if e.code is None else 1)
    finally:
        sys.stdout = tee.real
    rec['seconds'] = time.time() - t0
    rec['exit'] = code
    rec['stdout'] = tee.buf.getvalue().splitlines()
    if 'pyqtgraph' in sys.modules:
        rec['drawn'] = list(sys.modules['pyqtgraph']._RECORD)
    rec['cuda'] = bool(torch.cuda.is_available())
    json.dump(rec, open(record_path, 'w'))
    sys.exit(code)


if __name__ == '__main__':
    main()
