"""Runs the reference's OWN training script skeleton_matching/train_skeleton_matching.py, unmodified, through runpy - on the
reference's modules (mode 'reference': CPU, dgl / pytransform3d shims of oracle/shims) or on the B200 drop-in (mode 'dropin').

The script imports `gat2` and `graph_generator` from its own directory, which Python puts FIRST on sys.path when a file is run
as `python train_skeleton_matching.py` - so PYTHONPATH alone cannot shadow them. The drop-in way to run it is as a module from
the same working directory:

    cd <reference>/skeleton_matching
    PYTHONPATH=<repo>/3d_multi_pose_estimator_b200/shadow:. python -m train_skeleton_matching --trainset ... --devset ... --testset ...

which is what this runner reproduces (shadow directory, then the script's directory). It records stdout (the per-epoch losses
the script prints), the exit status and a digest of the checkpoint the script saves.

    python tests/drivers/run_train_driver.py --mode dropin --refroot <copy of baseline/_ref> --record out.json -- --trainset a.json ...
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
    ap.add_argument('--record', required=True)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('rest', nargs=argparse.REMAINDER)
    a = ap.parse_args()
    rest = a.rest[1:] if a.rest and a.rest[0] == '--' else a.rest
    refroot = os.path.abspath(a.refroot)
    record_path = os.path.abspath(a.record)
    sm_dir = os.path.join(refroot, 'skeleton_matching')
    # (the reference's requirements.txt packages that this image lacks - pytransform3d, needed to unpickle the tm_*.pickle - are
    # stood in for by tests/drivers/stubs, as for the other drivers)
    first = [SHADOW, os.path.join(HERE, 'stubs')] if a.mode == 'dropin' else [os.path.join(REPO, 'oracle', 'shims')]
    os.chdir(sm_dir)
    sys.path[:0] = first + [sm_dir]
    import numpy as np
    import torch
    import random
    import gat2
    import graph_generator
    rec = dict(mode=a.mode, argv=rest, modules={m.__name__: os.path.relpath(m.__file__, REPO) for m in (gat2, graph_generator)})
    random.seed(a.seed); np.random.seed(a.seed); torch.manual_seed(a.seed)
    tee = Tee(sys.stdout)
    sys.stdout = tee
    sys.argv = ['train_skeleton_matching.py'] + rest
    code, t0 = 0, time.time()
    try:
        runpy.run_path(os.path.join(sm_dir, 'train_skeleton_matching.py'), run_name='__main__')
    except SystemExit as e:
        code = e.code if isinstance(e.code, int) else (0 if e.code is None else 1)
    finally:
        sys.stdout = tee.real
    rec['seconds'] = time.time() - t0
    rec['exit'] = code
    rec['stdout'] = tee.buf.getvalue().splitlines()
    rec['cuda'] = bool(torch.cuda.is_available())
    ck = os.path.join(sm_dir, 'skeleton_matching.tch')
    if os.path.exists(ck):
        st = torch.load(ck, map_location='cpu')
        rec['checkpoint'] = {k: [float(v.double().sum()), float(v.double().abs().sum()), list(v.shape)] for k, v in st.items()}
    rec['prms'] = os.path.exists(os.path.join(sm_dir, 'skeleton_matching.prms'))
    json.dump(rec, open(record_path, 'w'))
    sys.exit(code)


if __name__ == '__main__':
    main()
