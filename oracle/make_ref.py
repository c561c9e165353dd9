"""TEST / BASELINE INFRASTRUCTURE - not product code.

Stages an UNMODIFIED copy of the reference tree under baseline/_ref/ (git-ignored, so it never enters the history, but
not gpurun-ignored, so it travels to the GPU box like the built .so). The reference is plain Python with no packaging
metadata (`pip install /root/reference` has nothing to build: no setup.py / pyproject.toml), so "installing" it is a
byte-for-byte copy of its files; a manifest with the sha256 of every file is written next to it and re-checked by the
tests that execute the reference's own drivers (tests/test_reference_drivers_*.py), so "unmodified" is verifiable.

What uses baseline/_ref: the driver tests (the unmodified test/*.py scripts run with the drop-in directory first on
PYTHONPATH), and bench.py's cpu_baseline / --impl reference legs (the unmodified reference path under oracle/shims, timed
on the host cores: cpu_baseline.kind = "reference").

    python oracle/make_ref.py            # needs /root/reference (only exists in the build container)
"""
import hashlib
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('B200POSE_REFERENCE_SRC', '/root/reference')
DST = os.path.join(REPO, 'baseline', '_ref')
KEEP_EXT = ('.py', '.json', '.pickle', '.txt', '.md', '.yaml')


def sha256(path):
    h = hashlib.sha256()
    with open(path, 'rb') as f:
        h.update(f.read())
    return h.hexdigest()


def stage(verbose=True):
    if not os.path.isdir(os.path.join(SRC, 'skeleton_matching')):
        if verbose:
            print('make_ref: no reference tree at %s - keeping %s as it is' % (SRC, DST))
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d not in ('.git', '__pycache__')]
        for fn in files:
            if not (fn.endswith(KEEP_EXT) or fn == 'LICENSE'):
                continue
            s = os.path.join(root, fn)
            rel = os.path.relpath(s, SRC)
            d = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            os.chmod(d, 0o644)
            manifest[rel] = sha256(d)
    json.dump(manifest, open(os.path.join(DST, 'MANIFEST.sha256.json'), 'w'), indent=1, sort_keys=True)
    if verbose:
        print('make_ref: %d files of the reference staged under %s' % (len(manifest), DST))
    return True


def verify():
    """True when baseline/_ref exists and every file still has the digest recorded when it was staged."""
    mf = os.path.join(DST, 'MANIFEST.sha256.json')
    if not os.path.exists(mf):
        return False
    manifest = json.load(open(mf))
    return all(os.path.exists(os.path.join(DST, rel)) and sha256(os.path.join(DST, rel)) == dig for rel, dig in manifest.items())


if __name__ == '__main__':
    stage()
    sys.exit(0 if verify() else 1)
