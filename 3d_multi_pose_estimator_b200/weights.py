"""Random-init weights of the reference architectures, reproducible from a seed.

Checkpoints are not available offline, so tests and the benchmark use the reference constructors'
own initialisation. The tensors are created in exactly the order the reference constructors create
them (skeleton_matching/gat2.py:25-48,101-135; utils/mlp.py:8-28), so for a given torch version and
`torch.manual_seed` value they are bit-identical to what the reference would hold; the golden fixtures
store checksums to prove it.

State-dict key names are the reference's: GAT `layers.{l}.{attn_l,attn_r,fc1.weight,fc1.bias,
fc2.weight,fc2.bias}`; MLP `layers.{1,3,...,17}.{weight,bias}`.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn as nn

# shipped hyper-parameters: skeleton_matching/train_skeleton_matching.py:40-56,148-149
GAT_HIDDEN = (40, 40, 40, 30)
GAT_HEADS = (10, 10, 8, 5)
GAT_LAYERS = 5
GAT_ALPHA = 0.15
GAT_ACT_SLOPE = 0.01           # nn.LeakyReLU() default
MLP_HIDDEN = (3072, 3072, 2048, 2048, 1024, 1024, 1024, 1024)   # utils/mlp.py:11-27
MLP_SLOPE = 0.1


def gat_layer_dims(in_dim: int, num_hidden: Sequence[int] = GAT_HIDDEN, heads: Sequence[int] = GAT_HEADS,
                   num_classes: int = 1):
    """[(in_dim, heads, out_dim)] per layer as GAT2.__init__ wires them (gat2.py:101-135)."""
    dims = [(in_dim, heads[0], num_hidden[0])]
    for l in range(1, len(num_hidden)):
        dims.append((num_hidden[l - 1] * heads[l - 1], heads[l], num_hidden[l]))
    dims.append((num_hidden[-1] * heads[-1], 1, num_classes))
    return dims


def make_gat_state(in_dim: int, seed: int = 0, bias: bool = True, num_hidden: Sequence[int] = GAT_HIDDEN,
                   heads: Sequence[int] = GAT_HEADS, residual: bool = False) -> Dict[str, torch.Tensor]:
    """residual=True adds `layers.{l}.res_fc.*` for every layer after the first whose in_dim != out_dim, created after the
    layer's attention vectors as gat2.py:43-46 does."""
    torch.manual_seed(seed)
    state = {}
    for l, (din, h, dout) in enumerate(gat_layer_dims(in_dim, num_hidden, heads)):
        fc1 = nn.Linear(din, din, bias=bias)
        fc2 = nn.Linear(din, h * dout, bias=bias)
        attn_l = torch.empty(h, dout, 1)
        attn_r = torch.empty(h, dout, 1)
        nn.init.xavier_normal_(fc1.weight.data, gain=1.414)
        nn.init.xavier_normal_(fc2.weight.data, gain=1.414)
        nn.init.xavier_normal_(attn_l, gain=1.414)
        nn.init.xavier_normal_(attn_r, gain=1.414)
        state['layers.%d.attn_l' % l] = attn_l
        state['layers.%d.attn_r' % l] = attn_r
        state['layers.%d.fc1.weight' % l] = fc1.weight.data
        state['layers.%d.fc2.weight' % l] = fc2.weight.data
        if bias:
            state['layers.%d.fc1.bias' % l] = fc1.bias.data
            state['layers.%d.fc2.bias' % l] = fc2.bias.data
        if residual and l > 0 and din != dout:
            res_fc = nn.Linear(din, h * dout, bias=bias)
            nn.init.xavier_normal_(res_fc.weight.data, gain=1.414)
            state['layers.%d.res_fc.weight' % l] = res_fc.weight.data
            if bias:
                state['layers.%d.res_fc.bias' % l] = res_fc.bias.data
    return state


def make_mlp_state(in_dim: int, out_dim: int = 54, seed: int = 1) -> Dict[str, torch.Tensor]:
    torch.manual_seed(seed)
    dims = (in_dim,) + MLP_HIDDEN + (out_dim,)
    state = {}
    for i in range(9):
        lin = nn.Linear(dims[i], dims[i + 1])
        state['layers.%d.weight' % (2 * i + 1)] = lin.weight.data
        state['layers.%d.bias' % (2 * i + 1)] = lin.bias.data
    return state


def state_checksum(state) -> Dict[str, list]:
    out = {}
    for k, v in state.items():
        a = v.detach().double().cpu().numpy().ravel()
        out[k] = [float(a.sum()), float(abs(a).sum()), float(a[0]), float(a[-1])]
    return out
