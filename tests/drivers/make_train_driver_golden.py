"""Runs the UNMODIFIED skeleton_matching/train_skeleton_matching.py of the staged reference copy (baseline/_ref, CPU, import
shims) on the synthetic single-person files of make_fixtures.make_training_files and stores what it printed - the training
loss of each of its 100 epochs, the validation lines, the final test-set MSE - plus a digest of the checkpoint it saved, as
tests/golden/drivers/train_skeleton_matching.json. Build container only (about half a minute).

    python tests/drivers/make_train_driver_golden.py
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import harness  # noqa: E402
import make_fixtures  # noqa: E402


def run(mode, tmp, timeout=1800):
    files = make_fixtures.make_training_files(os.path.join(tmp, 'files'))
    root = os.path.join(tmp, 'reference')
    shutil.copytree(harness.REF, root)
    rec_path = os.path.join(tmp, 'record_%s.json' % mode)
    cmd = [sys.executable, os.path.join(HERE, 'run_train_driver.py'), '--mode', mode, '--refroot', root, '--record', rec_path, '--',
           '--trainset'] + files['trainset'] + ['--devset'] + files['devset'] + ['--testset'] + files['testset']
    env = dict(os.environ)
    env.pop('PYTHONPATH', None)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout, env=env)
    if p.returncode != 0 or not os.path.exists(rec_path):
        raise AssertionError('train_skeleton_matching.py (%s) exited %d:\n%s' % (mode, p.returncode, p.stdout.decode(errors='replace')[-4000:]))
    rec = json.load(open(rec_path))
    strip = os.path.join(tmp, 'files') + '/'
    rec['argv'] = [a.replace(strip, '') for a in rec['argv']]
    rec['stdout'] = [l.replace(strip, '') for l in rec['stdout']]
    rec['modules'] = {k: os.path.basename(os.path.dirname(v)) + '/' + os.path.basename(v) for k, v in rec['modules'].items()}
    return rec


def main():
    assert harness.reference_staged() and not harness.verify_reference_unmodified()
    with tempfile.TemporaryDirectory() as tmp:
        rec = run('reference', tmp)
    out = os.path.join(harness.GOLDEN_DIR, 'train_skeleton_matching.json')
    json.dump(rec, open(out, 'w'), indent=0)
    print('written', out, '-', len([l for l in rec['stdout'] if l.startswith('loss:')]), 'epochs,', rec['stdout'][-1])


if __name__ == '__main__':
    main()
