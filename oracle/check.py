"""TEST INFRASTRUCTURE - not product code. Shared parity checks (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).

Two things live here:

* `explain_assignment_mismatch`: the attribution SURVEY.md 7-2 asks for. Person assignment is bit-exact GIVEN the scores;
  end to end it can only differ from the reference where the reference's own ranking of matchings is decided by a score
  gap below the score tolerance (adjacent sorted scores differ by as little as 1 ulp). A frame whose GPU assignment differs
  from the reference's is accepted only if (a) the GPU scores are within the tolerance of the reference's, (b) the oracle's
  clustering run on the GPU's scores reproduces the GPU's answer exactly, and (c) the reference's scores really hold such a
  near-tie: two matchings adjacent in its sorted order closer than the tolerance allows to separate, or a score within
  the tolerance of the threshold. Anything else raises.
* `check_frames`: whole-path comparison of a few frames of a batch result with the oracle (scores, assignment on the
  GPU's scores, 3D joints).
"""
from __future__ import annotations

import numpy as np

from . import pose_oracle as O

SCORE_RTOL = 1e-4        # BASELINE.json north_star: edge scores within 1e-4 relative
JOINT_TOL_M = 0.5e-3     # 3D joints within 0.5 mm


def near_tie(ref_scores_enodes: np.ndarray, thr: float = 0.5, rtol: float = SCORE_RTOL):
    """Does the reference's score vector contain a decision a perturbation of `rtol` relative can flip?
    Returns a description or None."""
    s = np.asarray(ref_scores_enodes, dtype=np.float64)
    if s.size == 0:
        return None
    at_thr = np.abs(s - thr) <= rtol * np.abs(s)
    if at_thr.any():
        return 'score %.9g within %.0e of the threshold' % (float(s[np.argmax(at_thr)]), rtol)
    above = np.sort(s[s > thr])[::-1]
    if above.size >= 2:
        gap = above[:-1] - above[1:]
        lim = rtol * (np.abs(above[:-1]) + np.abs(above[1:]))          # each of the two may move by rtol * |s|
        k = int(np.argmin(gap - lim))
        if gap[k] <= lim[k]:
            return 'sorted matchings %.9g / %.9g are %.2e apart (tolerance %.2e)' % (above[k], above[k + 1], gap[k], lim[k])
    return None


def explain_assignment_mismatch(ref_scores, gpu_scores, og, cfg, gpu_props, thr=0.5, rtol=SCORE_RTOL) -> str:
    """og: oracle graph of the frame (pose_oracle.build_graph). Raises AssertionError unless the mismatch is fully explained;
    returns the explanation."""
    idx = og['indices']
    ref_e, gpu_e = np.asarray(ref_scores)[idx], np.asarray(gpu_scores)[idx]
    rel = np.abs(gpu_e - ref_e) / np.abs(ref_e)
    assert rel.max() <= rtol, 'scores off by %g relative' % rel.max()
    again = O.cluster(np.asarray(gpu_scores), og['pairs'], og['nodes_camera'][:og['n_heads']], cfg.V_sm, og['n_heads'], thr,
                      cfg.min_number_of_views)
    assert np.array_equal(again, gpu_props), 'the clustering of the GPU scores is not the GPU assignment: clustering bug'
    why = near_tie(ref_e, thr, rtol)
    assert why is not None, 'assignment differs although no reference score gap is below the tolerance'
    return why


def check_frames(cfg, frames, gat_w, mlp_w, frame_scores, frame_props, frame_joints, ref_scores=None, ref_props=None):
    """frames: reference frame dicts (empty cameras already dropped); gat_w / mlp_w: numpy state dicts; frame_scores[b]: the
    GPU's N_b scores; frame_props[b]: [P_b, V_sm] head ids (frame-local, -1 = none); frame_joints[b]: [P_b, 54] metres or None.
    Returns dict(worst_score_rel, worst_joint_mm, frames, persons, explained)."""
    tabs = O.CameraTables(cfg)
    worst_s = worst_j = 0.0
    persons_total = 0
    for b, f in enumerate(frames):
        og = O.build_graph(f, tabs)
        if og is None:
            assert len(frame_props[b]) == 0, 'frame %d: persons although the reference builds no graph' % b
            continue
        ref = O.gat_forward(gat_w, og['feats'], og['src'], og['dst'])
        idx = og['indices']
        got = np.asarray(frame_scores[b])
        assert got.shape[0] == og['n_nodes'], 'frame %d: node count' % b
        rel = np.abs(got[idx] - ref[idx]) / np.abs(ref[idx])
        worst_s = max(worst_s, float(rel.max()))
        assert rel.max() <= SCORE_RTOL, 'frame %d: edge scores off by %g relative' % (b, rel.max())
        props = O.cluster(got, og['pairs'], og['nodes_camera'][:og['n_heads']], cfg.V_sm, og['n_heads'], 0.5, cfg.min_number_of_views)
        assert np.array_equal(np.asarray(frame_props[b]).reshape(-1, cfg.V_sm), props), \
            'frame %d: person assignment differs from the oracle clustering of the same scores' % b
        if frame_joints is None or len(props) == 0:
            continue
        want = []
        for person in props:
            sm = cfg.used_sm_names                                     # metrics_from_model.py:248-252: used_cameras order
            pd = {c: og['heads_json'][int(person[sm.index(c)])] for c in cfg.used_pe_names if c in sm and person[sm.index(c)] >= 0}
            x = O.encode_person(pd, tabs)
            want.append(None if x is None else O.mlp_forward(mlp_w, x[None])[0] * np.float32(10.))
        for p, w in enumerate(want):
            if w is None:
                continue
            d = float(np.abs(np.asarray(frame_joints[b][p]) - w).max())
            worst_j = max(worst_j, d)
            assert d <= JOINT_TOL_M, 'frame %d person %d: 3D joints off by %.4f mm' % (b, p, d * 1e3)
        persons_total += len(props)
    return dict(worst_score_rel=worst_s, worst_joint_mm=worst_j * 1e3, frames=len(frames), persons=persons_total)
