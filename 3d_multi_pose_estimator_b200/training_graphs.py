"""Training-side graphs on the B200 path (the optimisation step itself is train.py): the sample synthesis of
`MergedMultipleHumansDataset.process_training` (reference skeleton_matching/graph_generator.py:516-560, 672-810, with the
view augmentation of utils/data_augmentation.py:14-89) and the block-diagonal batching the training / validation drivers
do with `dgl.batch` (train_skeleton_matching.py:67-84, sm_metrics_without_gt.py:46-64).

The order-defining parts (which sample goes into which graph, which head is the person and which is spurious, the order
of the edge-nodes and their labels) are host-side list logic, as in the reference; they produce the packed skeletons
and an explicit edge-node list. Everything with arithmetic in it - the five edges per edge-node and the CSR
(b200pose_build_graph_pairs), the node features (b200pose_node_features), the GAT forward and the clustering - runs on
the device over that list.
"""
from __future__ import annotations

import dataclasses
import itertools
import json
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .pack import PackedBatch, pack_frames


def augment_views(samples: List[dict], used_cameras: Sequence[str], min_views: int = 1) -> List[dict]:
    """Every sample cut down to the used cameras that detected something, followed by its camera subsets with at least
    `min_views` cameras (data_augmentation.py:50-89; subsets in itertools.product order over the used-camera list)."""
    used = list(used_cameras)
    out = []
    for sample in samples:
        seen = [c for c in sample if c in used and json.loads(sample[c][0])]
        if not seen:
            continue
        have = tuple(1 if c in seen else 0 for c in used)
        out.append({c: sample[c] for c in seen})
        for keep in itertools.product((0, 1), repeat=len(used)):
            if keep == have or sum(keep) < min_views or any(k > h for k, h in zip(keep, have)):
                continue
            out.append({c: sample[c] for c in seen if keep[used.index(c)]})
    return out


def load_inputs(files: List[List[dict]], mode: str, used_cameras: Sequence[str], rnd) -> Tuple[list, list]:
    """Per file: augmentation unless the mode is a test mode, then the shuffled index list the sampler pops from
    (graph_generator.py:526-538). `rnd`: the `random` module, as the reference uses the global generator."""
    inputs, indices = [], []
    for samples in files:
        if mode not in ('test', 'test_generated'):
            samples = augment_views(samples, used_cameras, 2)
        order = list(range(len(samples)))
        if mode != 'test':
            rnd.shuffle(order)
        inputs.append(samples)
        indices.append(order)
    return inputs, indices


def sample_sets(inputs, indices, probabilities, limit: int, rnd):
    """Tuples of single-person samples that share a graph (sample_and_remove, graph_generator.py:675-697)."""
    prob = np.array(probabilities)
    for _ in range(limit):
        if not any(len(l) for l in inputs):
            return
        n_people = rnd.randint(1, len(inputs))
        chosen = np.argpartition(prob, -n_people)[-n_people:]
        views = []
        for f in chosen:
            if not indices[f]:
                return
            views.append(inputs[f][indices[f].pop()])
        if views:
            yield views


def edge_node_list(pb: PackedBatch) -> Tuple[np.ndarray, np.ndarray]:
    """Edge-nodes and labels of the graph whose heads are all skeletons of `pb`, every frame of `pb` being one
    single-person sample (graph_generator.py:722-800). Per sample and camera the skeleton with most joints (first on
    ties) is the person, the rest are spurious; edge-nodes are ORDERED pairs of heads from different cameras: a person's
    own heads (label 1), then its heads against every other person's, then against the spurious heads, person after
    person, and finally spurious against spurious (labels 0)."""
    cams = pb.sk_cam
    joints = np.array([bin(int(m)).count('1') for m in pb.sk_mask], dtype=np.int64)
    people, spurious = [], []
    for f in range(pb.n_frames):
        h0, h1 = int(pb.head_off[f]), int(pb.head_off[f + 1])
        person = []
        h = h0
        while h < h1:                                           # runs of equal camera = the cameras of the sample, in order
            e = h
            while e < h1 and cams[e] == cams[h]:
                e += 1
            good = h + int(np.argmax(joints[h:e]))              # first maximum, like max(enumerate(..), key=..)
            person.append(good)
            spurious += [x for x in range(h, e) if x != good]
            h = e
        people.append(person)
    pairs, labels = [], []

    def cross(list1, list2, label):
        for a in list1:
            for b in list2:
                if cams[a] != cams[b]:
                    pairs.append((a, b)); labels.append(label)

    for ip, person in enumerate(people):
        cross(person, person, 1.)
        for io, other in enumerate(people):
            if io != ip:
                cross(person, other, 0.)
        cross(person, spurious, 0.)
    cross(spurious, spurious, 0.)
    return np.array(pairs, dtype=np.int32).reshape(-1, 2), np.array(labels, dtype=np.float64).reshape(-1, 1)


def as_single_graph(pb: PackedBatch, n_enodes: int, max_in_degree: int) -> PackedBatch:
    """The skeletons of all frames of `pb` as the heads of ONE graph with `n_enodes` edge-nodes."""
    H = pb.n_heads
    return dataclasses.replace(pb, n_frames=1, head_off=np.array([0, H], dtype=np.int32),
                               node_off=np.array([0, H + n_enodes], dtype=np.int32),
                               max_heads=max(H, max_in_degree), max_enodes=n_enodes, skeletons=None, skeleton_index=None)


def max_in_degree(pairs: np.ndarray, n_heads: int) -> int:
    if len(pairs) == 0:
        return 1
    return 1 + int(np.bincount(pairs.ravel(), minlength=max(n_heads, 1)).max())


def training_graph_inputs(multi_person: List[Dict[str, list]], cfg):
    """(PackedBatch of one graph, pairs, labels) for one tuple of single-person samples; None when the tuple yields no
    edge-node (graph_generator.py:802)."""
    pb = pack_frames(multi_person, cfg, keep_json=False)
    pairs, labels = edge_node_list(pb)
    if len(pairs) == 0:
        return None
    return as_single_graph(pb, len(pairs), max_in_degree(pairs, pb.n_heads)), pairs, labels


def batch_device(members, device):
    """dgl.batch over graphs resident on the device: `members` = [(DeviceBatch, GraphArrays)]. Returns the merged
    DeviceBatch and the concatenated graph-local edge-node list; heads / nodes / edges keep their member order, which is
    the id shift dgl.batch applies."""
    from .pipeline import DeviceBatch
    cat = lambda name: torch.cat([getattr(db, name)[: db.n_heads] for db, _ in members])
    head_off, node_off, pairs = [0], [0], []
    for db, arrays in members:
        if db.host_offsets is not None:
            ho, no = db.host_offsets
        else:
            ho = db.head_off[: db.n_frames + 1].cpu().numpy().astype(np.int64)
            no = db.node_off[: db.n_frames + 1].cpu().numpy().astype(np.int64)
        head_off += list(head_off[-1] + ho[1:])
        node_off += list(node_off[-1] + no[1:])
        pairs.append(arrays.pairs[: db.n_enodes])
    i32 = lambda a: torch.tensor(a, dtype=torch.int32, device=device)
    merged = DeviceBatch(len(head_off) - 1, int(head_off[-1]), int(node_off[-1]),
                         max(db.max_heads for db, _ in members), max(db.max_enodes for db, _ in members),
                         cat('sk_xy'), cat('sk_vp'), cat('sk_mask'), cat('sk_cam'), i32(head_off), i32(node_off),
                         host_offsets=(np.array(head_off, dtype=np.int64), np.array(node_off, dtype=np.int64)))
    return merged, torch.cat(pairs) if pairs else torch.empty((0, 2), dtype=torch.int32, device=device)


def batch_packed(graphs: List[Tuple[PackedBatch, np.ndarray]]) -> Tuple[PackedBatch, np.ndarray]:
    """The same union on the host, over packed graphs that have not been uploaded yet: one PackedBatch whose "frames" are
    the graphs, and the concatenated graph-local edge-node list - the ingest format of a validation batch."""
    head_off, node_off = [0], [0]
    for pb, _ in graphs:
        head_off += list(head_off[-1] + np.asarray(pb.head_off[1:], dtype=np.int64))
        node_off += list(node_off[-1] + np.asarray(pb.node_off[1:], dtype=np.int64))
    cat = lambda name: np.concatenate([getattr(pb, name) for pb, _ in graphs])
    merged = PackedBatch(n_frames=len(head_off) - 1, sk_xy=cat('sk_xy'), sk_vp=cat('sk_vp'), sk_mask=cat('sk_mask'), sk_cam=cat('sk_cam'),
                         head_off=np.array(head_off, dtype=np.int32), node_off=np.array(node_off, dtype=np.int32),
                         max_heads=max(pb.max_heads for pb, _ in graphs), max_enodes=max(pb.max_enodes for pb, _ in graphs))
    return merged, np.concatenate([p for _, p in graphs]).astype(np.int32).reshape(-1, 2)
