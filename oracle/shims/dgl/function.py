"""Test-infrastructure shim (NOT product code): dgl.function.u_mul_e / sum."""
import torch


def u_mul_e(u, e, out):
    def msg(g):
        return {out: g.ndata[u][g._src] * g.edata[e]}
    return msg


def sum(m, out):  # noqa: A001 - mirrors dgl.function.sum
    def red(g, msgs):
        x = msgs[m]
        acc = torch.zeros((g._n,) + tuple(x.shape[1:]), dtype=x.dtype)
        acc.index_add_(0, g._dst, x)
        g.ndata[out] = acc
    return red
