"""Drop-in for the reference's utils/skeleton_matching_utils.py (get_person_proposal_from_network_output :12-132).

Threshold, score-ordered greedy camera-exclusive merge and connected components run in the one-warp-per-frame
clustering kernel (b200pose_cluster), which reproduces the reference's output bit for bit, including the
CPython-set iteration order its result depends on. The return value has the reference's shape: a list (component
order) of `{camera_name: head_id or None}` over `parameters.used_cameras_skeleton_matching`.
"""
from collections import namedtuple

import torch

import _b200pose_runtime as rt

Matching = namedtuple('Matching', 'id nodes score')


def get_person_proposal_from_network_output(outputs, subgraph, indices, nodes_camera, jsons_for_head=None,
                                            CLASSIFICATION_THRESHOLD=0.5):
    if not hasattr(subgraph, '_b200'):
        raise TypeError('get_person_proposal_from_network_output needs a graph built by the B200 graph_generator drop-in')
    ctx = rt.context()
    lv = getattr(subgraph, '_live', None)
    if lv is not None and lv.fresh() and type(outputs) is not list and getattr(lv, 'scores_ptr', None) == outputs.data_ptr() \
            and float(CLASSIFICATION_THRESHOLD) == ctx.threshold:
        rows = lv.proposals()                  # clustered inside the frame's submission, on exactly these scores
        if rows is not None:
            names = ctx.cfg.used_sm_names
            return [{names[s]: (h if h >= 0 else None) for s, h in enumerate(row)} for row in rows.tolist()]
    rt.sync_live()
    db, arrays = subgraph._b200
    if type(outputs) is list:                       # the reference accepts a python list as well (:27-30)
        scores = torch.tensor(outputs, dtype=torch.float32, device=ctx.device)
    else:
        scores = outputs.detach().to(ctx.device, torch.float32)
    if db.n_frames != 1:
        raise NotImplementedError('get_person_proposal_from_network_output works on one graph, not on a dgl.batch of %d' % db.n_frames)
    scores = scores.reshape(-1).contiguous()
    if scores.numel() != db.n_nodes:
        raise ValueError('expected one score per graph node (%d), got %d' % (db.n_nodes, scores.numel()))
    if nodes_camera is not None and len(nodes_camera) != db.n_nodes:
        raise ValueError('nodes_camera does not belong to this graph')
    person_heads, n_persons = ctx.cluster(db, arrays, scores, threshold=float(CLASSIFICATION_THRESHOLD))
    n = int(n_persons[0].item())
    rows = person_heads[:n].cpu().tolist()
    names = ctx.cfg.used_sm_names
    return [{names[s]: (h if h >= 0 else None) for s, h in enumerate(row)} for row in rows]
