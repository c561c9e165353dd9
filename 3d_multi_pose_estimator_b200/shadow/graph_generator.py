"""Drop-in for the reference's skeleton_matching/graph_generator.py.

Same public names: `MergedMultipleHumansDataset` (:516-916), `HumanGraphFromView` (:214-508), `graphData` (:21).
`MergedMultipleHumansDataset(frame_dict | [json paths], mode='test', alt='3')` builds, on the GPU and without
DGL, exactly the graph `process_test` builds (:813-876): heads in frame-dict camera order, one edge-node per
cross-camera skeleton pair wired with 5 directed edges, alternative-'3' node features bit-exact in fp32. Any other
mode ('train', 'dev', 'test_generated', ...) over a list of single-person files runs `process_training` (:672-810):
the same sample tuples for the same `random` seed, the same edge-node order and labels, graphs built on the GPU from
the explicit edge-node list (the drop-in GAT2 differentiates through them: 3d_multi_pose_estimator_b200/train.py). The graph object it hands out offers the
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


class _Lazy(dict):
    """A dict whose values are produced on first access (the feature matrix of a live frame is 650 KB that the driver
    loop only hands back to the model; the edge attributes are never read by it at all)."""

    def __init__(self, factories):
        dict.__init__(self)
        self._factories = dict(factories)

    def __missing__(self, key):
        if key not in self._factories:
            raise KeyError(key)
        self[key] = v = self._factories[key]()
        return v

    def _all(self):
        for k in self._factories:
            self[k]
        return self

    def __contains__(self, key):
        return key in self._factories or dict.__contains__(self, key)

    def __iter__(self):
        return dict.__iter__(self._all())

    def __len__(self):
        return dict.__len__(self._all())

    def keys(self):
        return dict.keys(self._all())

    def items(self):
        return dict.items(self._all())

    def values(self):
        return dict.values(self._all())


class LazyHeads(dict):
    """`jsons_for_head` of a live frame: head id -> skeleton dict, parsed camera by camera when a head is asked for (the
    packed arrays came from the native packer, so no Python object exists for a skeleton until the driver wants one)."""

    def __init__(self, frame, head_cam, head_idx):
        dict.__init__(self)
        self._frame, self._head_cam, self._head_idx, self._parsed = frame, head_cam, head_idx, {}

    def __missing__(self, head):
        if not (0 <= head < len(self._head_cam)):
            raise KeyError(head)
        cam = self._head_cam[head]
        lst = self._parsed.get(cam)
        if lst is None:
            payload = self._frame[cam][0]
            lst = self._parsed[cam] = json.loads(payload) if isinstance(payload, str) else payload
        self[head] = sk = lst[self._head_idx[head]]
        return sk

    def _all(self):
        for h in range(len(self._head_cam)):
            self[h]
        return self

    def __contains__(self, head):
        return isinstance(head, int) and 0 <= head < len(self._head_cam)

    def __iter__(self):
        return dict.__iter__(self._all())

    def __len__(self):
        return len(self._head_cam)

    def keys(self):
        return dict.keys(self._all())

    def items(self):
        return dict.items(self._all())

    def values(self):
        return dict.values(self._all())


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


class _LiveNodeData(_Lazy):
    """ndata of a live frame: 'h' aliases the feature buffer of the frame's submission while that is the latest of its
    shape, and is recomputed from the packed frame if somebody still reads it after the buffer was reused."""

    def __init__(self, graph):
        _Lazy.__init__(self, {'h': graph._features})
        self._graph = graph
        self._aliased = False

    def __getitem__(self, key):
        if key == 'h' and self._aliased and not self._graph._live.fresh():
            dict.pop(self, 'h', None)
            self._aliased = False
        return _Lazy.__getitem__(self, key)


class LiveGraph(B200Graph):
    """The graph of a frame that was submitted as a whole (3d_multi_pose_estimator_b200/live.py): same interface, but the
    feature matrix and the edge attributes are only materialised if somebody reads them, and `_live` lets GAT2 /
    get_person_proposal_from_network_output / PoseEstimatorDataset / PoseEstimatorMLP answer from the submission."""

    def __init__(self, handle, frame):
        self._live = handle
        self._frame = frame
        pb = handle.pb
        self._n, self._e = pb.n_nodes, pb.n_edges
        self.batch_size = 1
        self._node_shift = None
        self._own = None
        self.ndata = _LiveNodeData(self)
        self.edata = _Lazy({'rel_type': self._rel_type, 'norm': self._norm})

    @property
    def _b200(self):
        if self._live.fresh():
            return self._live.ent.db, self._live.ent.arrays
        if self._own is None:                                   # the submission's buffers have been reused: rebuild from the packed frame
            ctx = rt.context()
            rt.sync_live()
            db = rt.pipeline.HostBatch(self._live.pb, pinned=False).to_device(ctx.device)
            self._own = (db, ctx.build_graph(db, with_coo=True))
        return self._own

    def _features(self):
        f = self._live.features()                               # materialised inside the submission (valid while it is the latest of its shape)
        if f is not None:
            self.ndata._aliased = True
            return f
        ctx = rt.context()
        db, _ = self._b200
        return ctx.node_features_f32(db)

    def _rel_type(self):
        H = self._live.pb.n_heads
        rel = torch.ones(self._e, dtype=torch.int64, device=rt.context().device)
        rel[:H] = 0
        rel[H + 4::5] = 2
        return rel

    def _norm(self):
        return torch.ones((self._e, 1), dtype=torch.float32, device=rt.context().device)

    def edges(self):
        if self._live.fresh():
            torch.cuda.current_stream(rt.context().device).wait_event(self._live.ent.ev_a)
        return B200Graph.edges(self)

    def nodes(self):
        return torch.arange(self._n, dtype=torch.int32, device=rt.context().device)

    @property
    def device(self):
        return rt.context().device


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
            lv = rt.live() if type(json_view) is dict else None
            if lv is not None:
                # the models the driver runs are known: submit the whole frame (graph, GAT, clustering, encoder, MLP) now;
                # the driver's later calls are answered from the submission
                cfg = rt.config()
                pb = rt.pack.pack_frames_fast([json_view], cfg, keep_json=False)
                handle = lv.submit(pb) if pb.skeleton_index is not None else None
                if handle is not None:
                    head_cam = [cfg.camera_names[int(c)] for c in pb.sk_cam]
                    handle.frame, handle.head_cam, handle.head_idx = json_view, head_cam, list(pb.skeleton_index[0])
                    self.jsons_for_head = LazyHeads(json_view, head_cam, handle.head_idx)
                    self.skeleton_index = dict(enumerate(handle.head_idx))
                    H, N = pb.n_heads, pb.n_nodes
                    self.graphs.append(LiveGraph(handle, json_view))
                    self.labels.append(th.zeros((N - H, 1), dtype=th.float64))
                    self.data['edge_nodes_indices'].append(th.arange(H, N, dtype=th.int64).unsqueeze(1))
                    self.data['nodes_camera'].append(head_cam + [''] * (N - H))
                    continue
            rt.sync_live()
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
