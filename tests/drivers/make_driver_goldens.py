"""Runs every driver script of the UNMODIFIED reference (baseline/_ref, CPU, dgl/pytransform3d shims) on the synthetic
fixtures of make_fixtures.py and stores the records under tests/golden/drivers/ - what the drop-in runs of the same
scripts are compared with on the GPU box. Also writes the digest manifest of the reference's files.

    python oracle/make_ref.py && python tests/drivers/make_driver_goldens.py      (build container only)
"""
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import harness                      # noqa: E402
import make_fixtures                # noqa: E402


def main():
    assert harness.reference_staged(), 'run oracle/make_ref.py first'
    os.makedirs(harness.GOLDEN_DIR, exist_ok=True)
    shutil.copyfile(os.path.join(harness.REF, 'MANIFEST.sha256.json'), harness.MANIFEST)
    tmp = tempfile.mkdtemp(prefix='b200pose_drv_')
    fx = make_fixtures.make(os.path.join(tmp, 'fixtures'))
    root = harness.working_copy(tmp, fx)
    for script in harness.SCRIPTS:
        rec = harness.run('reference', root, script, fx, os.path.join(tmp, script + '.json'))
        rec['stdout'] = [l for l in rec['stdout'] if '%|' not in l]
        rec.pop('modules', None)
        json.dump(rec, open(os.path.join(harness.GOLDEN_DIR, script.replace('.py', '.json')), 'w'))
        print('%-36s exit %d  %4.1f s  %d matched frames, %d MLP calls, %d triangulations, %d drawn' %
              (script, rec['exit'], rec['seconds'], len(rec['proposals']), len(rec['mlp_out']), len(rec['triangulate']), len(rec['drawn'])))
        for l in rec['stdout'][-6:]:
            print('      ' + l)
    shutil.rmtree(tmp)


if __name__ == '__main__':
    main()
