"""Drop-in for the one DGL entry point the reference's drivers call themselves: `dgl.batch`
(skeleton_matching/train_skeleton_matching.py:80, test/sm_metrics_without_gt.py:60).

The graphs of the B200 graph_generator drop-in are not DGL objects, so their block-diagonal union has to come from
here: `batch(graphs)` concatenates the members' packed skeletons and edge-node lists on the device, shifts ids graph by
graph like DGL does, and builds the union's edges / CSR with b200pose_build_graph_pairs. The result is a B200Graph with
`batch_size`, `batch_num_nodes()`, `.ndata['h']`, `.edata`, `.edges()`, `.nodes()`, `.to()`, which the B200 GAT2 runs in
one launch sequence (a frame batch IS a block-diagonal graph batch on this path).

Nothing else of DGL lives here: with this directory first on sys.path, `import dgl` resolves to this module, and any other
attribute raises with a pointer to this note rather than failing later inside a kernel.
"""
import torch

import _b200pose_runtime as rt

__version__ = '0.0+b200pose'


def batch(graphs, ndata=None, edata=None):
    graphs = list(graphs)
    if not graphs:
        raise ValueError('dgl.batch: empty list of graphs')
    for g in graphs:
        if not hasattr(g, '_b200'):
            raise TypeError('dgl.batch (B200 drop-in) needs graphs built by the B200 graph_generator drop-in')
    from graph_generator import B200Graph
    ctx = rt.context()
    members = [g._b200 for g in graphs]
    db, pairs = rt.training_graphs.batch_device(members, ctx.device)
    arrays = ctx.build_graph_pairs(db, pairs, with_coo=True)
    feats = torch.cat([g.ndata['h'] for g in graphs], dim=0)
    return B200Graph(db, arrays, feats)


def __getattr__(name):
    if name == 'DGLGraph':
        from graph_generator import B200Graph
        return B200Graph
    raise AttributeError('the B200 drop-in provides only dgl.batch (see 3d_multi_pose_estimator_b200/shadow/dgl/__init__.py); '
                         'dgl.%s is not available' % name)
