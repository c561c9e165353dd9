"""CPU: pins the driver goldens. With the staged reference copy present (build container), the unmodified
test/metrics_from_model.py is run again on the reference's own modules and must reproduce the committed record; the
staged files must carry the committed digests; the synthetic fixtures must be the ones the goldens were made from."""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'drivers'))
import harness                      # noqa: E402
import make_fixtures                # noqa: E402
import helpers                      # noqa: E402


def test_fixture_frames_are_deterministic():
    """The GPU box regenerates the fixtures; they must be the files the goldens were recorded on."""
    tmp = tempfile.mkdtemp(prefix='b200pose_fx_')
    fx = make_fixtures.make(tmp)
    digest = hashlib.sha256(open(fx['test_file'], 'rb').read()).hexdigest()
    want = json.load(open(os.path.join(harness.GOLDEN_DIR, 'fixtures.sha256.json')))
    assert digest == want['test_file']
    assert hashlib.sha256(open(fx['viewer_file'], 'rb').read()).hexdigest() == want['viewer_file']
    for i, p in enumerate(fx['single_files']):
        assert hashlib.sha256(open(p, 'rb').read()).hexdigest() == want['single_%d' % i]


def test_every_driver_has_a_golden_record():
    for script in harness.SCRIPTS:
        rec = json.load(open(os.path.join(harness.GOLDEN_DIR, script.replace('.py', '.json'))))
        assert rec['exit'] == 0 and rec['mode'] == 'reference' and len(rec['proposals']) > 0


@pytest.mark.skipif(not harness.reference_staged(), reason='baseline/_ref not staged (python oracle/make_ref.py)')
def test_staged_reference_is_unmodified():
    assert harness.verify_reference_unmodified() == []
    if os.path.isdir('/root/reference'):                     # and the digests are those of the tree it was copied from
        want = json.load(open(harness.MANIFEST))
        for rel, dig in want.items():
            assert harness.sha256(os.path.join('/root/reference', rel)) == dig, rel


@pytest.mark.skipif(not harness.reference_staged(), reason='baseline/_ref not staged (python oracle/make_ref.py)')
def test_reference_run_reproduces_the_golden_record():
    cfg, npz, meta = helpers.load_golden('panoptic')
    tmp = tempfile.mkdtemp(prefix='b200pose_drv_')
    fx = make_fixtures.make(os.path.join(tmp, 'fixtures'))
    root = harness.working_copy(tmp, fx)
    script = 'metrics_from_model.py'
    got = harness.run('reference', root, script, fx, os.path.join(tmp, 'rec.json'))
    ref = json.load(open(os.path.join(harness.GOLDEN_DIR, script.replace('.py', '.json'))))
    assert got['proposals'] == ref['proposals']
    assert np.array_equal(np.concatenate([np.asarray(m).ravel() for m in got['mlp_out']]),
                          np.concatenate([np.asarray(m).ravel() for m in ref['mlp_out']]))
    harness.compare(got, ref, fx, cfg)
