"""Shared by tests/test_reference_drivers_{cpu,gpu}.py and tests/drivers/make_driver_goldens.py: the list of the
reference's driver scripts with the command line each gets, a sub-process runner around tests/drivers/run_driver.py, and
the comparison of a drop-in run with the record of the unmodified reference on the same files."""
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
REPO = os.path.dirname(TESTS)
REF = os.path.join(REPO, 'baseline', '_ref')
GOLDEN_DIR = os.path.join(TESTS, 'golden', 'drivers')
MANIFEST = os.path.join(GOLDEN_DIR, 'reference_manifest.sha256.json')     # committed digests of the reference's files

# script -> builder of its command line from the fixture paths (README.md of the reference: usage of each script)
SCRIPTS = {
    'metrics_from_model.py': lambda f: ['--testfiles', f['test_file'], '--tmdir', f['tm_dir'], '--modelsdir', f['models'], '--datastep', '1'],
    'show_results_from_model.py': lambda f: ['--testfile', f['viewer_file'], '--showgt', '--tmfile', f['tm_file'], '--modelsdir', f['models'], '--datastep', '1'],
    'metrics_from_triangulation.py': lambda f: ['--testfiles', f['test_file'], '--tmdir', f['tm_dir'], '--modelsdir', f['models'], '--datastep', '1'],
    'show_results_from_triangulation.py': lambda f: ['--testfile', f['viewer_file'], '--showgt', '--tmfile', f['tm_file'], '--modelsdir', f['models'], '--datastep', '1'],
    'sm_metrics.py': lambda f: ['--testfiles', f['test_file'], '--tmdir', f['tm_dir'], '--modelsdir', f['models'], '--datastep', '1'],
    'reprojection_error.py': lambda f: ['--testfiles', f['test_file'], '--showgt', '--tmdir', f['tm_dir'], '--modelsdir', f['models'], '--datastep', '1'],
    'sm_metrics_without_gt.py': lambda f: ['--testfiles'] + f['single_files'] + ['--modelsdir', f['models'], '--datastep', '1'],
}


def reference_staged():
    return os.path.isdir(os.path.join(REF, 'test'))


def sha256(path):
    h = hashlib.sha256()
    h.update(open(path, 'rb').read())
    return h.hexdigest()


def verify_reference_unmodified():
    """Every staged file has the digest committed in tests/golden/drivers/reference_manifest.sha256.json."""
    want = json.load(open(MANIFEST))
    bad = [rel for rel, dig in want.items() if not os.path.exists(os.path.join(REF, rel)) or sha256(os.path.join(REF, rel)) != dig]
    return bad


def working_copy(tmpdir, fixtures):
    """A scratch copy of baseline/_ref to run in (the scripts write a cache/ directory next to themselves, and
    test/sm_metrics.py:81-84 loads '../models_panoptic/*' whatever --modelsdir says)."""
    root = os.path.join(tmpdir, 'reference')
    shutil.copytree(REF, root)
    shutil.copytree(fixtures['models'], os.path.join(root, 'models_panoptic'))
    return root


def relocate(rec, fixtures_dir):
    """Paths in argv / stdout depend on the scratch directory: keep file names only."""
    strip = lambda s: s.replace(fixtures_dir.rstrip('/') + '/', '').replace(fixtures_dir, '')
    rec['argv'] = [strip(a) for a in rec['argv']]
    rec['stdout'] = [strip(l) for l in rec['stdout']]
    return rec


def run(mode, refroot, script, fixtures, record_path, timeout=900):
    cmd = [sys.executable, os.path.join(HERE, 'run_driver.py'), '--mode', mode, '--refroot', refroot, '--script', script,
           '--record', record_path, '--'] + SCRIPTS[script](fixtures)
    env = dict(os.environ)
    env.pop('PYTHONPATH', None)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout, env=env)
    if p.returncode != 0 or not os.path.exists(record_path):
        raise AssertionError('%s (%s) exited %d:\n%s' % (script, mode, p.returncode, p.stdout.decode(errors='replace')[-4000:]))
    return relocate(json.load(open(record_path)), fixtures['tm_dir'])


_NUM = re.compile(r'[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?|nan|inf')
_TIMING = ('Mean time', 'it/s', 's/it', 'cache/MergedMultipleHumansDataset')     # timing means, tqdm, the DGL cache-file notice


def metric_lines(stdout):
    """The lines a script prints as its result (timing means and tqdm bars excluded): [(text with numbers blanked, [numbers])]."""
    out = []
    for l in stdout:
        if any(t in l for t in _TIMING) or '%|' in l:
            continue
        nums = [float(x) for x in _NUM.findall(l)]
        out.append((_NUM.sub('#', l).strip(), nums))
    return out


def compare(got, ref, fixtures, cfg, score_rtol=1e-4, joint_tol_m=0.5e-3, tri_tol_m=1e-7, metric_rtol=2e-3):
    """got: record of a drop-in run; ref: record of the unmodified reference on the same files. Returns a report dict;
    raises AssertionError on any difference the tolerances of BASELINE.json do not cover."""
    from oracle import check as OC
    from oracle import pose_oracle as O
    assert got['exit'] == 0 and ref['exit'] == 0
    rep = dict(script=ref['script'], calls=len(ref['proposals']), explained=[], worst_score_rel=0.0, worst_joint_mm=0.0, worst_tri_m=0.0)
    assert len(got['proposals']) == len(ref['proposals']), 'number of matched frames differs'
    skip_stage3 = set()
    tabs = O.CameraTables(cfg)
    graph_frames = None
    for i, (gp, rp) in enumerate(zip(got['proposals'], ref['proposals'])):
        gs, rs = np.asarray(got['scores'][i], np.float64), np.asarray(ref['scores'][i], np.float64)
        assert gs.shape == rs.shape, (i, gs.shape, rs.shape)
        is_label_vector = set(np.unique(rs)) <= {0.0, 1.0}          # sm_metrics_without_gt feeds the labels through the same call
        if gp != rp:
            # allowed only with the attribution of oracle.check (needs the frame's graph: the i-th frame that has one)
            if graph_frames is None:
                graph_frames = [f for f in ({c: fr[c] for c in fr if json.loads(fr[c][0])} for fr in fixtures['frames'])
                                if O.build_graph(f, tabs) is not None]
            assert ref['script'] in ('metrics_from_model.py', 'metrics_from_triangulation.py', 'sm_metrics.py', 'reprojection_error.py'), \
                'proposals differ in call %d' % i
            og = O.build_graph(graph_frames[i], tabs)
            names = cfg.used_sm_names
            arr = np.array([[-1 if p[c] is None else p[c] for c in names] for p in gp], dtype=np.int32).reshape(-1, len(names))
            rep['explained'].append((i, OC.explain_assignment_mismatch(rs, gs, og, cfg, arr)))
            skip_stage3.add(i)
        if not is_label_vector:
            rel = np.abs(gs - rs) / np.maximum(np.abs(rs), 1e-30)
            # head rows of the output are not edge scores (nothing reads them), but they obey the same tolerance in practice
            rep['worst_score_rel'] = max(rep['worst_score_rel'], float(rel.max()))
            assert rel.max() <= score_rtol, 'scores of call %d differ by %g relative' % (i, rel.max())
    if not rep['explained']:
        assert len(got['mlp_out']) == len(ref['mlp_out'])
        for g, r in zip(got['mlp_out'], ref['mlp_out']):
            g, r = np.asarray(g) * 10.0, np.asarray(r) * 10.0
            assert g.shape == r.shape
            if g.size:
                rep['worst_joint_mm'] = max(rep['worst_joint_mm'], float(np.abs(g - r).max()) * 1e3)
                assert np.abs(g - r).max() <= joint_tol_m, '3D joints differ by %.4f mm' % (np.abs(g - r).max() * 1e3)
        assert len(got['triangulate']) == len(ref['triangulate'])
        for g, r in zip(got['triangulate'], ref['triangulate']):
            assert sorted(g) == sorted(r)
            for j in r:
                d = float(np.abs(np.asarray(g[j]) - np.asarray(r[j])).max())
                rep['worst_tri_m'] = max(rep['worst_tri_m'], d)
                assert d <= tri_tol_m * max(1.0, float(np.abs(np.asarray(r[j])).max())), 'triangulated joint differs by %g m' % d
        gm, rm = metric_lines(got['stdout']), metric_lines(ref['stdout'])
        gm = [m for m in gm if not m[0].startswith('MLP input size')]      # printed by the module; same line in both, position may differ
        rm = [m for m in rm if not m[0].startswith('MLP input size')]
        assert [m[0] for m in gm] == [m[0] for m in rm], 'printed result block differs:\n%s\n--- reference:\n%s' % (gm, rm)
        for (t, a), (_, b) in zip(gm, rm):
            for x, y in zip(a, b):
                assert abs(x - y) <= metric_rtol * max(abs(y), 1e-6) + 1e-9, 'printed metric differs: %s: %r vs %r' % (t, x, y)
        assert len(got['drawn']) == len(ref['drawn'])
        for (gk, gv), (rk, rv) in zip(got['drawn'], ref['drawn']):
            assert gk == rk
            gv, rv = np.asarray(gv, np.float64), np.asarray(rv, np.float64)
            assert gv.shape == rv.shape
            if gv.size:
                assert np.abs(gv - rv).max() <= joint_tol_m, 'drawn %s differs by %g m' % (gk, np.abs(gv - rv).max())
        rep['drawn'] = len(ref['drawn'])
    rep['metric_block'] = [l for l in got['stdout'] if not ('%|' in l)]
    return rep
