"""CPU: the native JSON frame packer (b200pose_pack_json) against the Python packer / the reference's own parsing."""
import importlib
import json
import time

import numpy as np
import pytest

import helpers

pack = importlib.import_module('3d_multi_pose_estimator_b200.pack')
libmod = importlib.import_module('3d_multi_pose_estimator_b200._lib')


def same(a, b):
    assert a.n_frames == b.n_frames and a.max_heads == b.max_heads and a.max_enodes == b.max_enodes
    for f in ('sk_xy', 'sk_vp', 'sk_mask', 'sk_cam', 'head_off', 'node_off'):
        x, y = getattr(a, f), getattr(b, f)
        assert x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y), f


@pytest.mark.parametrize('config', helpers.CONFIGS)
def test_golden_frames_pack_identically(config):
    cfg, npz, meta = helpers.load_golden(config)
    frames = [meta['frames'][t] for t in meta['cases']]                 # full reference frames: 4 elements per camera
    want = pack.pack_frames(frames, cfg)
    got = pack.pack_json(json.dumps(frames), cfg)
    same(got, want)
    assert got.skeleton_index == want.skeleton_index
    one = pack.pack_json(json.dumps(frames[0]), cfg, n_threads=1)         # a single frame object
    same(one, pack.pack_frames(frames[:1], cfg))


def test_number_parsing_is_exact_and_schema_variants():
    cfg, npz, meta = helpers.load_golden('panoptic')
    rng = np.random.default_rng(3)
    vals = [0.1, 1e-7, 123456789.123456789, 5e-324, 1.7976931348623157e308, 959.9999999999999, 2.5e-5, 1e22, 1e23, 3.0, 7,
            0.30000000000000004, 1234567890123456789012.0, -0.0, -17.25] + [float(x) for x in rng.uniform(0, 1920, 40)] + \
           [float(np.float32(x)) for x in rng.uniform(0, 1080, 20)]
    sk = {str(j % 18): [j % 18, vals[(2 * j) % len(vals)], vals[(2 * j + 1) % len(vals)], 1, 0.5] for j in range(18)}
    sk2 = {"ID": 7, "3": [3, vals[6], vals[9], 0, 0.25], "17": [17, 12, 13, 1, 1]}
    frames = [
        {"trackerb": [json.dumps([sk, {}, sk2]), 0.0, "no_image", [{"0": [1, 2, 3], "-1": [0, 0, 0]}]],
         "not_a_camera": [json.dumps([sk]), 1.0],
         "trackera": [[sk2, sk], 2.5, 'x " [ ] { } \\ y']},                                       # inline list + tricky trailing string
        {},                                                                                       # empty frame
        {"trackerd": ["[]", 0.0], "trackere": [json.dumps([{}]), 0.0]},                           # nothing usable
        {"trackerc": [json.dumps([sk2], indent=2), 0.0, "no_image", []]},                         # whitespace inside the string
    ]
    text = json.dumps(frames, indent=1)
    got = pack.pack_json(text, cfg)
    same(got, pack.pack_frames(json.loads(text), cfg))
    assert got.head_off.tolist() == [0, 4, 4, 4, 5] and got.node_off.tolist() == [0, 8, 8, 8, 9]
    assert got.skeleton_index == [[0, 2, 0, 1], [], [], [0]]
    pinned = pack.pack_json(text.encode(), cfg, pinned=False, n_threads=3)
    same(pinned, got)


def test_malformed_input_is_reported():
    cfg, npz, meta = helpers.load_golden('panoptic')
    for bad in ['', '[{"trackera": [', '[{"trackera": ["[{\\"1\\": [1, 2]}]"]}]', '[{"trackera": ["[{\\"99\\": [1,2,3,4,5]}]"]}]', '17']:
        with pytest.raises(libmod.B200PoseError):
            pack.pack_json(bad, cfg)


def test_native_packer_is_fast():
    cfg, npz, meta = helpers.load_golden('panoptic')
    frames = [helpers.synth.make_frame(cfg, 900 + i, 4, with_gt=True) for i in range(64)] * 8
    text = json.dumps(frames)
    t0 = time.perf_counter()
    got = pack.pack_json(text, cfg)
    t_native = time.perf_counter() - t0
    t0 = time.perf_counter()
    want = pack.pack_frames(json.loads(text), cfg, keep_json=False)
    t_python = time.perf_counter() - t0
    same(got, want)
    print('native %.1f ms, python %.1f ms for %d frames (%.1f MB)' % (1e3 * t_native, 1e3 * t_python, len(frames), len(text) / 1e6))
    assert t_native < t_python


def test_number_fast_paths_against_python_float():
    """Every conversion path of the native parser (short decimals, 16-19 digit decimals through the x87 division,
    strtod fallback) against Python's float() on a few hundred thousand adversarial and random literals."""
    cfg, npz, meta = helpers.load_golden('panoptic')
    rng = np.random.default_rng(11)
    lits = []
    for x in rng.uniform(0, 2000, 60000):
        lits.append(repr(float(x)))
    for x in rng.uniform(0, 1, 30000):
        lits.append(repr(float(x) * 10.0 ** int(rng.integers(-12, 12))))
    for m in rng.integers(1, 2 ** 63 - 1, 40000):                       # 19-digit mantissas, many near rounding boundaries
        lits.append('%d.%de%d' % (m // 10 ** 9, m % 10 ** 9, int(rng.integers(-20, 10))))
    for k in range(20000):                                              # exact half-way cases between adjacent doubles
        a = float(rng.uniform(1, 2000))
        mid = (np.float64(a).view(np.int64) | 1)
        lo = np.int64(mid).view(np.float64)
        from decimal import Decimal
        hw = (Decimal(float(lo)) + Decimal(float(np.nextafter(lo, np.inf)))) / 2
        lits.append(format(hw, 'f')[:22])
    want = np.array([float(t) for t in lits])
    per = 18
    frames = []
    for i in range(0, len(lits) - 2 * per, 2 * per):
        sk = '{' + ', '.join('"%d": [%d, %s, %s, 1, 1]' % (j, j, lits[i + 2 * j], lits[i + 2 * j + 1]) for j in range(per)) + '}'
        frames.append('{"trackera": [[%s], 0.0]}' % sk)
    got = pack.pack_json('[' + ', '.join(frames) + ']', cfg)
    n = got.sk_xy.shape[0]
    assert n == len(frames)
    assert np.array_equal(got.sk_xy.reshape(-1), want[:n * 2 * per])


def test_binary_ingest_roundtrip(tmp_path):
    cfg, npz, meta = helpers.load_golden('arp3')
    pb = pack.pack_frames([meta['frames'][t] for t in meta['cases']], cfg, keep_json=False)
    pb.save(str(tmp_path / 'batch.npz'))
    same(pack.PackedBatch.load(str(tmp_path / 'batch.npz')), pb)


@pytest.mark.parametrize('config', helpers.CONFIGS)
def test_pack_frames_fast_is_identical(config):
    """pack_frames_fast (camera payload strings handed to the native packer) against the Python packer: same arrays bit for
    bit, same skeleton dicts and per-camera indices - golden frames plus ragged ones (missing joints, unseen persons,
    joint-less skeletons), one frame at a time and as a batch; inline-list payloads fall back to pack_frames."""
    cfg, npz, meta = helpers.load_golden(config)
    frames = [meta['frames'][t] for t in meta['cases']]
    frames += [helpers.synth.make_frame(cfg, 900 + i, 3, drop_joint_p=0.4, drop_view_p=0.3, rand_conf=True, keep_empty=True) for i in range(12)]
    keys = ('sk_xy', 'sk_vp', 'sk_mask', 'sk_cam', 'head_off', 'node_off')
    for batch in [[f] for f in frames] + [frames]:
        a, b = pack.pack_frames(batch, cfg), pack.pack_frames_fast(batch, cfg)
        assert all(np.array_equal(getattr(a, k), getattr(b, k)) for k in keys)
        assert (a.max_heads, a.max_enodes, a.n_frames) == (b.max_heads, b.max_enodes, b.n_frames)
        assert a.skeletons == b.skeletons and a.skeleton_index == b.skeleton_index
    c = pack.pack_frames_fast(frames, cfg, keep_json=False)
    assert c.skeletons is None and np.array_equal(c.sk_xy, a.sk_xy)
    inline = [{cam: [json.loads(p[0])] + list(p[1:]) for cam, p in f.items()} for f in frames[:3]]
    d, e = pack.pack_frames_fast(inline, cfg), pack.pack_frames(inline, cfg)
    assert all(np.array_equal(getattr(d, k), getattr(e, k)) for k in keys)


def test_literals_locale_and_unterminated_spans():
    """true / false in the valid / prob slots are 1.0 / 0.0 like the Python packer's numpy stores; null is an error (the
    reference raises a TypeError on it); NaN / Infinity pass through; numbers parse the same under a decimal-comma
    LC_NUMERIC; and the parser never reads past json + len (the C ABI takes a span, not a C string)."""
    import ctypes as C
    import locale
    cfg, npz, meta = helpers.load_golden('panoptic')
    text = '[{"trackera": ["[{\\"3\\": [3, 100.5, 200.25, true, false], \\"4\\": [4, 1e400, -1e400, 1, 1]}]", 0.0]}]'
    got = pack.pack_json(text, cfg)
    want = pack.pack_frames(json.loads(text), cfg)
    same(got, want)
    assert got.sk_vp[0, 3].tolist() == [1.0, 0.0] and np.isinf(got.sk_xy[0, 4]).all()
    nan_text = '[{"trackera": [[{"3": [3, NaN, -Infinity, 1, 1]}], 0.0]}]'
    g2 = pack.pack_json(nan_text, cfg)
    assert np.isnan(g2.sk_xy[0, 3, 0]) and g2.sk_xy[0, 3, 1] == -np.inf
    with pytest.raises(libmod.B200PoseError):
        pack.pack_json('[{"trackera": [[{"3": [3, null, 1.0, 1, 1]}], 0.0]}]', cfg)
    # 17-significant-digit numbers take the strtod fallback: same bits under a decimal-comma locale
    hard = '[{"trackera": [[{"3": [3, 123.45678901234567891, 0.30000000000000004441, 1, 1]}], 0.0]}]'
    base = pack.pack_json(hard, cfg).sk_xy.copy()
    old = locale.setlocale(locale.LC_NUMERIC)
    try:
        for name in ('de_DE.UTF-8', 'fr_FR.UTF-8', 'de_DE', 'C.UTF-8'):
            try:
                locale.setlocale(locale.LC_NUMERIC, name)
                break
            except locale.Error:
                continue
        assert np.array_equal(pack.pack_json(hard, cfg).sk_xy, base)
    finally:
        locale.setlocale(locale.LC_NUMERIC, old)
    assert base[0, 3, 0] == 123.45678901234567891 and base[0, 3, 1] == 0.30000000000000004441
    # the span ends in the middle of a number, with more digits (and no NUL) behind it in memory: a clean parse error
    L = libmod.lib()
    blob = b'[{"trackera": [[{"3": [3, 1.5, 2.5, 1, 0.123456789012345678' + b'9' * 64
    buf = C.create_string_buffer(blob, len(blob))
    names = (C.c_char_p * 1)(b'trackera'); idx = (C.c_int32 * 1)(0)
    h = C.c_void_p()
    rc = L.b200pose_pack_json(buf, len(blob) - 64, 1, names, idx, 1, C.byref(h))
    assert rc != 0 and h.value is None


def test_packed_batch_plan_is_validated():
    """A batch whose declared max_heads / max_enodes do not cover its frames is rejected on the host (a kernel would skip
    the frames that exceed the plan)."""
    cfg, npz, meta = helpers.load_golden('panoptic')
    pm = importlib.import_module('3d_multi_pose_estimator_b200.pipeline')
    frames = [meta['frames']['p4a'], meta['frames']['p3']]
    pb = pack.pack_frames(frames, cfg)
    pb.validate()
    pb.max_heads -= 1
    with pytest.raises(ValueError):
        pm.HostBatch(pb, pinned=False)
    pb.max_heads += 1
    pb.max_enodes = 3
    with pytest.raises(ValueError):
        pb.validate()
    assert pm.person_capacity(20, 2) == 10 and pm.person_capacity(20, 1) == 20 and pm.person_capacity(20, 0) == 20 and pm.person_capacity(0, 2) == 1
