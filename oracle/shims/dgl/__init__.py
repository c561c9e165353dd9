"""Test-infrastructure shim (NOT product code): the subset of DGL the reference's
inference path touches (gat2.py:12-14,59-66,84; graph_generator.py:8-11,867-870;
skeleton_matching_utils.py:26), restated on plain torch following DGL's documented
semantics: edge ids = insertion order; edge_softmax normalises over in-edges of the
destination; update_all(u_mul_e, sum) sums messages per destination. Parity with the real
DGL is unpinned (DGL is absent from this image and from the reference's requirements)."""
import torch
from . import function, ops, data  # noqa: F401


class _EdgeBatch(object):
    def __init__(self, g):
        self.src = {k: v[g._src] for k, v in g.ndata.items()}
        self.dst = {k: v[g._dst] for k, v in g.ndata.items()}
        self.data = g.edata


class DGLGraph(object):
    def __init__(self, src, dst, num_nodes, idtype):
        self._src = torch.as_tensor(src, dtype=torch.int64)
        self._dst = torch.as_tensor(dst, dtype=torch.int64)
        self._n = int(num_nodes)
        self._idtype = idtype
        self.ndata = {}
        self.edata = {}

    def edges(self):
        return self._src.to(self._idtype), self._dst.to(self._idtype)

    def nodes(self):
        return torch.arange(self._n, dtype=self._idtype)

    def number_of_nodes(self):
        return self._n

    num_nodes = number_of_nodes

    def number_of_edges(self):
        return int(self._src.shape[0])

    num_edges = number_of_edges

    def to(self, device):
        return self

    def apply_edges(self, udf):
        self.edata.update(udf(_EdgeBatch(self)))

    def update_all(self, msg, red):
        m = msg(self)
        red(self, m)


def graph(data, num_nodes=None, idtype=torch.int64):
    src, dst = data
    return DGLGraph(src, dst, num_nodes, idtype)


def save_graphs(path, graphs, labels=None):
    """The reference caches a non-test dataset next to the script (graph_generator.py:884-896). The cache only saves a
    rebuild on the next run - it changes no result - so the shim writes nothing (has_cache() then stays False)."""
    return None


def load_graphs(path):
    raise NotImplementedError("dgl shim: graph cache I/O is out of scope")


def batch(graphs):
    """dgl.batch: block-diagonal union; node/edge ids shifted graph by graph, ndata/edata concatenated."""
    src, dst, off = [], [], 0
    for g in graphs:
        src.append(g._src + off)
        dst.append(g._dst + off)
        off += g._n
    out = DGLGraph(torch.cat(src), torch.cat(dst), off, graphs[0]._idtype)
    for k in graphs[0].ndata:
        out.ndata[k] = torch.cat([g.ndata[k] for g in graphs], dim=0)
    for k in graphs[0].edata:
        out.edata[k] = torch.cat([g.edata[k] for g in graphs], dim=0)
    return out
