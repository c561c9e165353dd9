"""TEST INFRASTRUCTURE - not product code. Only tests/, tests/golden/make_golden.py,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import anything under oracle/.

Sets up an interpreter so the UNMODIFIED reference modules under /root/reference import and run
on CPU in this container: the two shims (oracle/shims: dgl, pytransform3d) go first on sys.path,
then the reference's own directories in the order its scripts append them
(test/metrics_from_model.py:12,17,23), cwd = <reference>/test because the reference resolves
'../tm_panoptic.pickle' relative to cwd (parameters.py:70).

The reference tree only exists in the build container; on the GPU box the staged copy under baseline/_ref (same bytes,
digests in tests/golden/drivers/reference_manifest.sha256.json) is what this activates.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
# the reference tree itself (build container) or the unmodified copy staged by oracle/make_ref.py (travels to the GPU box)
_STAGED = os.path.join(os.path.dirname(_HERE), 'baseline', '_ref')
REFERENCE_ROOT = os.environ.get('B200POSE_REFERENCE_ROOT') or ('/root/reference' if os.path.isdir('/root/reference/skeleton_matching') else _STAGED)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'skeleton_matching'))


def activate(configuration='PANOPTIC', parameters_override=None):
    """Make `import gat2, graph_generator, mlp, ...` resolve to the reference's files.

    configuration: 'PANOPTIC' (as shipped) or 'ARPLAB'. The reference selects it by editing the
    CONFIGURATION constant (parameters.py:47); here the file's text is executed in memory with that
    one constant substituted, so the ARPLAB numbers are the reference's own (parameters.py:80-118).
    parameters_override: optional callable(namedtuple) -> namedtuple for derived configurations
    (e.g. ARP cut to 3 cameras) applied on top.
    """
    if not reference_available():
        raise RuntimeError('reference tree not found at %s' % REFERENCE_ROOT)
    for p in (os.path.join(REFERENCE_ROOT, ''), os.path.join(REFERENCE_ROOT, 'utils'),
              os.path.join(REFERENCE_ROOT, 'skeleton_matching'), os.path.join(_HERE, 'shims')):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    os.chdir(os.path.join(REFERENCE_ROOT, 'test'))
    import types
    src = open(os.path.join(REFERENCE_ROOT, 'parameters.py')).read()
    marker = "CONFIGURATION = 'PANOPTIC'"
    assert src.count(marker) == 1, 'reference parameters.py layout changed'
    src = src.replace(marker, "CONFIGURATION = %r" % configuration)
    mod = types.ModuleType('parameters')
    mod.__file__ = os.path.join(REFERENCE_ROOT, 'parameters.py')
    exec(compile(src, mod.__file__, 'exec'), mod.__dict__)
    if parameters_override is not None:
        mod.parameters = parameters_override(mod.parameters)
    sys.modules['parameters'] = mod
    return mod.parameters


CONFIGS = ('panoptic', 'arp3', 'ring10', 'arp6', 'arp_robot2', 'pansub')


def activate_config(config):
    """The named configurations of the goldens and of bench.py, as `parameters` of the unmodified reference:
    panoptic = PANOPTIC as shipped; arp6 = ARPLAB as shipped; arp3 = ARPLAB cut to its first three cameras (BASELINE.json
    configs[2]); arp_robot2 = ARPLAB with the robot-cameras-only lines of parameters.py:110-112; pansub = PANOPTIC with a
    permuted 4-of-5 skeleton-matching set and a 3-camera pose-estimator set; ring10 = a synthetic 10-camera ring
    (BASELINE.json configs[4]) whose TransformManager pickle is written to a temporary file."""
    tm_arp = os.path.join(REFERENCE_ROOT, 'tm_arp.pickle')
    tm_pan = os.path.join(REFERENCE_ROOT, 'tm_panoptic.pickle')
    if config == 'panoptic':
        return activate('PANOPTIC', lambda p: p._replace(transformations_path=tm_pan))
    if config == 'arp6':
        return activate('ARPLAB', lambda p: p._replace(transformations_path=tm_arp))
    if config == 'arp3':
        return activate('ARPLAB', lambda p: p._replace(
            cameras=[0, 1, 2], camera_names=p.camera_names[:3], used_cameras=p.camera_names[:3],
            used_cameras_skeleton_matching=p.camera_names[:3], transformations_path=tm_arp))
    if config == 'arp_robot2':
        return activate('ARPLAB', lambda p: p._replace(
            used_cameras=['orinbot_l', 'orinbot_r'], used_cameras_skeleton_matching=['orinbot_l', 'orinbot_r'],
            transformations_path=tm_arp))
    if config == 'pansub':
        return activate('PANOPTIC', lambda p: p._replace(
            used_cameras_skeleton_matching=['trackerd', 'trackerb', 'trackere', 'trackerc'],
            used_cameras=['trackerb', 'trackere', 'trackerd'], transformations_path=tm_pan))
    if config == 'ring10':
        import importlib
        import pickle
        import tempfile
        repo = os.path.dirname(_HERE)
        if repo not in sys.path:
            sys.path.insert(0, repo)
        sys.path.insert(0, os.path.join(_HERE, 'shims'))
        from pytransform3d.transform_manager import TransformManager
        cfg = importlib.import_module('3d_multi_pose_estimator_b200').ring_config(10)
        tm = TransformManager.__new__(TransformManager)
        tm.transforms = {('root', n): cfg.T_root2cam[i] for i, n in enumerate(cfg.camera_names)}
        f = tempfile.NamedTemporaryFile(prefix='tm_ring10_', suffix='.pickle', delete=False)
        pickle.dump(tm, f)
        f.close()
        V = 10
        return activate('PANOPTIC', lambda p: p._replace(
            cameras=list(range(V)), camera_names=cfg.camera_names, used_cameras=cfg.camera_names,
            used_cameras_skeleton_matching=cfg.camera_names,
            fx=[float(x) for x in cfg.fx], fy=[float(x) for x in cfg.fy], cx=[float(x) for x in cfg.cx],
            cy=[float(x) for x in cfg.cy],
            kd0=[0.] * V, kd1=[0.] * V, kd2=[0.] * V, p1=[0.] * V, p2=[0.] * V, transformations_path=f.name))
    raise ValueError(config)
