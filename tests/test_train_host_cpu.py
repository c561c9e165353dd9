"""CPU: host logic of the training step that needs no device - the flat parameter layout (3d_multi_pose_estimator_b200/train.py)
against the reference's state-dict keys and shapes (gat2.py:25-48 via weights.make_gat_state), and the refusals."""
import importlib

import numpy as np
import pytest

import helpers

train_mod = importlib.import_module('3d_multi_pose_estimator_b200.train')


def test_param_layout_covers_the_reference_state_dict():
    state = helpers.weights_mod.make_gat_state(902, 0, True)
    shapes = {k: tuple(v.shape) for k, v in state.items()}
    layers, slots, n = train_mod.param_layout(shapes)
    assert sorted(slots) == sorted(shapes)
    assert [(l['din'], l['heads'], l['dim']) for l in layers] == helpers.weights_mod.gat_layer_dims(902)
    spans = []
    for key, (off, rows, cols, ld) in slots.items():
        assert off % 4 == 0 and ld % 4 == 0 and ld >= cols and ld - cols < 4, key          # 16-byte aligned rows (TMA store pitch)
        assert rows * cols == int(np.prod(shapes[key])), key
        spans.append((off, off + rows * ld))
    spans.sort()
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))                                # contiguous, no overlap
    n_params = sum(int(np.prod(s)) for s in shapes.values())
    assert n_params == 1961439 and n_params <= n <= n_params + 4 * sum(r for _, r, _, _ in slots.values())
    for l, lay in enumerate(layers):
        assert lay['n2'] == lay['hd'] + 2 * lay['heads'] and lay['ldz'] % 4 == 0 and lay['ldz'] >= lay['n2']


def test_param_layout_refusals():
    no_bias = {k: tuple(v.shape) for k, v in helpers.weights_mod.make_gat_state(902, 0, False).items()}
    with pytest.raises(NotImplementedError):
        train_mod.param_layout(no_bias)
    residual = {k: tuple(v.shape) for k, v in helpers.weights_mod.make_gat_state(902, 0, True, residual=True).items()}
    with pytest.raises(ValueError):
        train_mod.param_layout(residual)                                   # res_fc weights in a model not declared residual
    layers, slots, n = train_mod.param_layout(residual, residual=True)
    assert [l['res'] for l in layers] == [None, 'fc', 'fc', 'fc', 'fc'] and sorted(slots) == sorted(residual)
    ident = {k: tuple(v.shape) for k, v in helpers.weights_mod.make_gat_state(902, 0, True, [40, 40], [1, 4], residual=True).items()}
    layers, slots, n = train_mod.param_layout(ident, residual=True)
    assert [l['res'] for l in layers] == [None, 'identity', 'fc'] and sorted(slots) == sorted(ident)
    with pytest.raises(ValueError):                                        # in_dim != out_dim and no res_fc weights
        train_mod.param_layout({k: v for k, v in residual.items() if 'res_fc' not in k}, residual=True)
    bad = {k: tuple(v.shape) for k, v in helpers.weights_mod.make_gat_state(902, 0, True).items()}
    bad['layers.1.fc2.weight'] = (399, 400)
    with pytest.raises(ValueError):
        train_mod.param_layout(bad)
