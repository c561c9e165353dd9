"""GPU: the reference's OWN driver scripts, unmodified (baseline/_ref/test/*.py, digests checked against the committed
manifest), executed the way a user switches to this repository - the shadow directory first on sys.path, cwd =
<reference>/test, the script's documented command line - on synthetic JSON / tm pickle / checkpoint files, and compared
with the records of the same scripts running on the reference's own modules (CPU, tests/drivers/make_driver_goldens.py).

Covers SURVEY.md 7-8 and 8f-2: metrics_from_model.py and show_results_from_model.py (the two drivers the north star
names; the viewer runs headless on a stub pyqtgraph that records the scene), metrics_from_triangulation.py,
show_results_from_triangulation.py, sm_metrics.py, reprojection_error.py and the training-side sm_metrics_without_gt.py.
What is compared: exit status, which modules the script actually imported, per-frame person proposals (exact; a difference
needs the near-tie attribution of oracle/check.py), scores (1e-4 relative), 3D joints (0.5 mm), triangulations (1e-7 m),
the printed result block and everything handed to the scene."""
import json
import os
import sys
import tempfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'drivers'))
import harness                      # noqa: E402
import make_fixtures                # noqa: E402
import helpers                      # noqa: E402

pytestmark = pytest.mark.gpu

_state = {}


def setup():
    if 'fx' not in _state:
        if not harness.reference_staged():
            pytest.skip('NO REFERENCE COPY under baseline/_ref (python oracle/make_ref.py stages it in the build container): '
                        'the unmodified driver scripts cannot run')
        bad = harness.verify_reference_unmodified()
        assert not bad, 'baseline/_ref differs from the committed digests: %s' % bad
        tmp = tempfile.mkdtemp(prefix='b200pose_drv_')
        fx = make_fixtures.make(os.path.join(tmp, 'fixtures'))
        _state.update(tmp=tmp, fx=fx, root=harness.working_copy(tmp, fx))
    return _state


@pytest.mark.parametrize('script', list(harness.SCRIPTS))
def test_unmodified_driver_on_dropin(script):
    st = setup()
    cfg, npz, meta = helpers.load_golden('panoptic')
    ref = json.load(open(os.path.join(harness.GOLDEN_DIR, script.replace('.py', '.json'))))
    got = harness.run('dropin', st['root'], script, st['fx'], os.path.join(st['tmp'], script + '.dropin.json'))
    assert got['cuda'], 'the drop-in run saw no CUDA device'
    for name, path in got['modules'].items():
        assert path.startswith(os.path.join('3d_multi_pose_estimator_b200', 'shadow')), (name, path)
    rep = harness.compare(got, ref, st['fx'], cfg)
    timing = [l for l in got['stdout'] if l.startswith('Mean time')]
    print('\n%s on the drop-in: exit 0 in %.1f s (reference on CPU here: %.1f s); %d matched frames, worst score deviation %.2e relative, '
          'worst joint deviation %.4f mm, worst triangulation deviation %.2e m, explained near-ties %s'
          % (script, got['seconds'], ref['seconds'], rep['calls'], rep['worst_score_rel'], rep['worst_joint_mm'], rep['worst_tri_m'], rep['explained']))
    for l in timing:
        print('   ', l)
