"""Drop-in for the reference's skeleton_matching/graph_generator.py.

Same public names: `MergedMultipleHumansDataset` (:516-916), `HumanGraphFromView` (:214-508), `graphData` (:21).
`MergedMultipleHumansDataset(frame_dict | [json paths], mode='test', alt='3')` builds, on the GPU and without
DGL, exactly the graph `process_test` builds (:813-876): heads in frame-dict camera order, one edge-node per
cross-camera skeleton pair wired with 5 directed edges, alternative-'3' node features bit-exact in fp32. Any other
mode ('train', 'dev', 'test_generated', ...) over a list of single-person files runs `process_training` (:672-810):
the same sample tuples for the same `random` seed, the same edge-node order and labels, graphs built on the GPU from
the explicit edge-node list (forward only - the B200 GAT2 has no backward). The graph object it hands out offers the
slice of the DGL API the reference's callers use (`.to`, `.ndata['h']`, `.edata`, `.edges()`, `.nodes()`,
`.number_of_nodes()`), carries the CSR the B200 GAT2 / clustering drop-ins consume, and is what the `dgl.batch` of
the sibling `dgl` drop-in module merges.

Out of scope (raise NotImplementedError): graph alternatives '1' and '2' (unused by the shipped configuration,
parameters.py:76) and the DGL cache files (every construction processes its inputs).
"""
import json
import random
import sys
from collections import namedtuple

import numpy as np
import torch
import torch as th

import _b200pose_runtime as rt

graphData = namedtuple('graphData', ['src_nodes', 'dst_nodes', 'n_nodes', 'features', 'edge_types', 'edge_norms'])

if th.cuda.is_available() is True:
    device = th.device('cuda')
else:
    device = th.device('cpu')

_COCO18 = ["nose", "left_eye", "right_eye", "left_ear", "right_ear", "left_shoulder", "right_shoulder", "left_elbow",
           "right_elbow", "left_wrist", "right_wrist", "left_hip", "right_hip", "left_knee", "right_knee", "left_ankle",
           "right_ankle", "neck"]
JOINTS_TYPES = {str(i): n for i, n in enumerate(_COCO18)}
NODE_TYPES_ONE_HOT = ['head', 'edge_node'] + _COCO18
JOINT_METRIC_FEATURES = ['i_coordinate', 'j_coordinate', 'valid2D', 'probability']
OTHER_FEATURES = ['n_joints']
_PER_JOINT = {'2': ['_i', '_j', '_valid', '_prob'],
              '3': ['_i', '_j', '_valid', '_prob', '_line_pX', '_line_pY', '_line_pZ', '_line_vX', '_line_vY', '_line_vZ']}
_FEATURES = {}
# every alternative of the reference keeps the same three relation types for the matching graph (:196-211)
RELATIONS = {'2': ['h_h', 'link', 'link_link'], '3': ['h_h', 'link', 'link_link']}


def _features(alt):
    """Feature-name vocabulary of an alternative (graph_generator.py:116-140); only its length and the positions
    of 'head' / 'edge_node' matter to the callers (test/metrics_from_model.py:47)."""
    if alt not in _FEATURES:
        cams = list(rt.config().used_sm_names)
        if alt == '1':
            _FEATURES[alt] = NODE_TYPES_ONE_HOT + cams + JOINT_METRIC_FEATURES + OTHER_FEATURES
        elif alt in _PER_JOINT:
            _FEATURES[alt] = ['head', 'edge_node'] + [c + '_' + p + s for c in cams for p in _COCO18 for s in _PER_JOINT[alt]]
        else:
            raise KeyError(alt)
    return _FEATURES[alt]


def _frame_to_device(frame, keep_json=True):
    ctx = rt.context()
    pb = rt.pack.pack_frames([frame], ctx.cfg, keep_json=keep_json)
    return ctx, pb, rt.pipeline.HostBatch(pb).to_device(ctx.device)


class B200Graph:
    """What `dgl.graph((src, dst), num_nodes, idtype=int32)` is to the reference's callers
    (graph_generator.py:867-870), backed by device tensors produced by b200pose_build_graph /
    b200pose_node_features."""

    def __init__(self, db, arrays, feats):
        self._b200 = (db, arrays)
        self._n = db.n_nodes
        self._e = db.n_edges
        self.batch_size = db.n_frames
        dev = feats.device
        # edge ids: per graph H self loops ('h_h'), then per edge-node four 'link' edges and one 'link_link' (:627-656)
        if db.n_frames == 1:
            H = db.n_heads
            rel = torch.ones(self._e, dtype=torch.int64, device=dev)
            rel[:H] = 0
            rel[H + 4::5] = 2
            self._node_shift = None
        else:                                                   # dgl.batch: graphs one after the other
            ho, no = db.head_off[: db.n_frames + 1].long(), db.node_off[: db.n_frames + 1].long()
            Hs = ho[1:] - ho[:-1]
            Es = Hs + 5 * (no[1:] - no[:-1] - Hs)
            graph_of = torch.repeat_interleave(torch.arange(db.n_frames, device=dev), Es)
            first = torch.cumsum(Es, 0) - Es
            local = torch.arange(self._e, device=dev) - first[graph_of]
            k = local - Hs[graph_of]
            rel = torch.where(k < 0, 0, torch.where(k % 5 == 4, 2, 1))
            self._node_shift = no[:-1][graph_of].to(torch.int32)
        self.ndata = {'h': feats}
        self.edata = {'rel_type': rel, 'norm': torch.ones((self._e, 1), dtype=torch.float32, device=dev)}

    def to(self, device, **kw):
        return self                                             # tensors already live on the CUDA device

    def edges(self):
        _, arrays = self._b200
        src, dst = arrays.src[: self._e], arrays.dst[: self._e]
        if self._node_shift is not None:                        # the kernels keep graph-local ids; DGL's are batch-global
            src, dst = src + self._node_shift, dst + self._node_shift
        return src, dst

    def batch_num_nodes(self):
        db, _ = self._b200
        no = db.node_off[: db.n_frames + 1].long()
        return no[1:] - no[:-1]

    def nodes(self):
        return torch.arange(self._n, dtype=torch.int32, device=self.ndata['h'].device)

    def number_of_nodes(self):
        return self._n

    num_nodes = number_of_nodes

    def number_of_edges(self):
        return self._e

    num_edges = number_of_edges

    @property
    def device(self):
        return self.ndata['h'].device


class HumanGraphFromView:
    joints = JOINTS_TYPES

    def __init__(self, data, camera, alt):
        self.labels = None
        self.num_rels = -1
        self.camera = camera
        cfg = rt.config()
        self.camera_idx = cfg.used_sm_names.index(self.camera)
        if alt != '3':
            if alt in ('1', '2'):
                raise NotImplementedError('graph alternative %s is not part of the B200 path (parameters.graph_alternative is 3)' % alt)
            print(f'Unknown network alternative {alt}')
            sys.exit(-1)
        # one head node with a self loop (initializeWithAlternative3, :444-508)
        self.num_joints = len([j for j in data if j != "ID"])
        self.n_nodes = 1
        self.src_nodes, self.dst_nodes = [0], [0]
        self.edge_types = [RELATIONS['3'].index('h_h')]
        self.edge_norms = [[1.]]
        if self.num_joints:
            ctx, pb, db = _frame_to_device({camera: [[data]]}, keep_json=False)
            self.features = ctx.node_features_f32(db)[:1].clone()
        else:
            self.features = torch.zeros((1, len(_features('3'))), dtype=torch.float32)
            self.features[0, 0] = 1.
        self.cam_from_root = torch.from_numpy(cfg.centre32(cfg.used_sm[self.camera_idx]))

    @staticmethod
    def get_node_types_one_hot():
        return NODE_TYPES_ONE_HOT

    @staticmethod
    def get_cam_types():
        return rt.config().used_sm_names

    @staticmethod
    def get_all_features(alt='1'):
        return _features(alt)

    @staticmethod
    def get_joint_metric_features():
        return JOINT_METRIC_FEATURES

    @staticmethod
    def get_other_features():
        return OTHER_FEATURES

    @staticmethod
    def get_rels(alt='1'):
        if alt not in RELATIONS:
            raise NotImplementedError('relation vocabulary of alternative %s is not part of the B200 path' % alt)
        return RELATIONS[alt]


class MergedMultipleHumansDataset:
    path_save = 'cache/'

    def __init__(self, paths, probabilities=[1.], limit='100000000', alt=None, mode='train', force_reload=False,
                 verbose=True, debug=False, raw_dir='.'):
        if alt is None:
            print('Alt is None')
            sys.exit(-1)
        if alt != '3':
            raise NotImplementedError('graph alternative %s is not part of the B200 path' % alt)
        self.inputs = []
        self.inputs_indices = []
        if type(paths) == list:
            files = []
            for path in paths:
                print('PATH', path)
                files.append(json.loads(open(path, "rb").read()))
            # augmentation (train modes) and the shuffled index lists, drawing from the global `random` like the reference
            self.inputs, self.inputs_indices = rt.training_graphs.load_inputs(files, mode, rt.config().used_pe_names, random)
        elif type(paths) == dict:
            self.inputs.append(paths)
            self.inputs_indices.append(list(range(len(paths))))
        else:
            raise Exception('Unhandled type for MergedMultipleHumansDataset')
        self.name = "MergedMultipleHumansDataset"
        self.probabilities = probabilities
        self.mode = mode
        self.alt = alt
        self.graphs = []
        self.labels = []
        self.data = dict()
        self.data['edge_nodes_indices'] = []
        self.data['nodes_camera'] = []
        self.debug = debug
        self.force_reload = force_reload or mode == 'test'
        self.device = device
        self.limit = limit
        self.verbose = verbose
        self.jsons_for_head = dict()
        self.skeleton_index = dict()
        self.process()

    def process(self):
        if self.mode != 'test':
            self.process_training()
        else:
            self.process_test()

    def process_training(self):
        """graph_generator.py:672-810: one graph per tuple of single-person samples. The tuples, the person / spurious
        split, the edge-node order and the labels come from the host-side list logic of training_graphs; edges, CSR
        and features are built on the device."""
        ctx = rt.context()
        cfg = rt.config()
        names = cfg.used_sm_names
        idx = 0
        for multi_person in rt.training_graphs.sample_sets(self.inputs, self.inputs_indices, self.probabilities,
                                                           int(self.limit), random):
            if idx % 1000 == 0:
                print(idx)
            idx += 1
            built = rt.training_graphs.training_graph_inputs(multi_person, cfg)
            if built is None:                                   # no cross-camera pair: no graph (:802)
                continue
            pb, pairs, labels = built
            db = rt.pipeline.HostBatch(pb).to_device(ctx.device)
            arrays = ctx.build_graph_pairs(db, pairs, with_coo=True)
            feats = ctx.node_features_f32(db)
            H, N = db.n_heads, db.n_nodes
            self.graphs.append(B200Graph(db, arrays, feats))
            self.labels.append(th.from_numpy(labels))
            self.data['edge_nodes_indices'].append(th.arange(H, N, dtype=th.int64).unsqueeze(1))
            self.data['nodes_camera'].append([names[cfg.used_sm.index(int(c))] for c in pb.sk_cam] + [''] * (N - H))

    def process_test(self):
        assert len(self.inputs) == 1, "For testing, please provide __ONE__ single JSON file"
        iterate_over = self.inputs[0] if type(self.inputs[0]) == list else self.inputs
        names = rt.config().used_sm_names
        idx = 0
        for json_view in iterate_over:
            if idx % 1000 == 0 and idx > 0:
                print(idx)
            if idx == self.limit:
                break
            idx += 1
            ctx, pb, db = _frame_to_device(json_view)
            # like the reference, the head bookkeeping always reflects the last frame seen (:573-605)
            self.jsons_for_head = dict(enumerate(pb.skeletons[0]))
            self.skeleton_index = dict(enumerate(pb.skeleton_index[0]))
            H, N = db.n_heads, db.n_nodes
            if N == H:                                          # no cross-camera pair: no graph (:866)
                continue
            arrays = ctx.build_graph(db, with_coo=True)
            feats = ctx.node_features_f32(db)
            self.graphs.append(B200Graph(db, arrays, feats))
            self.labels.append(th.zeros((N - H, 1), dtype=th.float64))
            self.data['edge_nodes_indices'].append(th.arange(H, N, dtype=th.int64).unsqueeze(1))
            self.data['nodes_camera'].append([names[rt.config().used_sm.index(int(c))] for c in pb.sk_cam] + [''] * (N - H))

    def __getitem__(self, idx):
        return self.graphs[idx], self.labels[idx], self.data['edge_nodes_indices'][idx], self.data['nodes_camera'][idx]

    def __len__(self):
        return len(self.graphs)

    def has_cache(self):                                        # no DGL cache files: every construction processes its inputs
        return False

    def save(self):                                             # (the reference itself skips the cache in test mode, :885)
        return

    def download(self):
        pass
