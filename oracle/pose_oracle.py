"""TEST INFRASTRUCTURE - not product code.

CPU restatement (numpy / plain Python) of the reference's per-frame inference path, used ONLY as
the parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs. The product path (3d_multi_pose_estimator_b200) never imports this module.

Pinning: every function below is checked against tests/golden/*.npz, which were produced by running
the unmodified reference here (tests/golden/make_golden.py). What is NOT pinned by anything the
reference ships: the behaviour of the un-vendored third-party libraries it calls (DGL edge_softmax /
update_all, pytransform3d inverse, cv2, networkx) - the goldens pin them to the versions in the build
container (cv2 4.13.0, networkx 3.6.1, CPython 3.12.3) and to the DGL shim's documented semantics.

Each function cites the reference file:line it follows (paths relative to the reference root).
"""
from __future__ import annotations

import itertools
import json
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32
N_JOINTS = 18


# ----------------------------------------------------------------------------------------------
# fp32 helpers reproducing the arithmetic order of torch's CPU matmul for tiny operands:
# sequential FMA over k with the first product rounded (verified bit-exact against the goldens).
# ----------------------------------------------------------------------------------------------
def _fma32(a, b, c):
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def _dot_fma32(row, vec):
    acc = f32(f32(row[0]) * f32(vec[0]))
    for k in range(1, len(row)):
        acc = _fma32(f32(row[k]), f32(vec[k]), acc)
    return acc


def _matvec_fma32(M, v):
    return np.array([_dot_fma32(M[i], v) for i in range(M.shape[0])], dtype=f32)


# ----------------------------------------------------------------------------------------------
# Stage 1: graph build  (skeleton_matching/graph_generator.py)
# ----------------------------------------------------------------------------------------------
class CameraTables:
    """Module-import tables of graph_generator.py:32-52 and pose_estimator_dataset_from_json.py:28-47."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.sm_names = cfg.used_sm_names
        self.pe_names = cfg.used_pe_names
        self.ti32 = {n: cfg.T_cam2root32(i) for i, n in enumerate(cfg.camera_names)}      # dataset.py:38-41, by camera name
        # graph_generator.py:38-52 appends its tables walking parameters.camera_names and keeping the cameras that are in
        # used_cameras_skeleton_matching, but HumanGraphFromView reads them at used_cameras_skeleton_matching.index(camera)
        # (:232-233, :488-489): slot s of the used list gets the tables of the s-th used camera in camera_names order.
        # The same camera whenever the two orders agree (every shipped configuration); pinned by the 'pansub' goldens.
        in_rig_order = [n for n in cfg.camera_names if n in self.sm_names]
        table_cam = {n: cfg.camera_names.index(in_rig_order[s]) for s, n in enumerate(self.sm_names)}
        self.kinv32 = {n: cfg.Kinv32(table_cam[n]) for n in self.sm_names}
        self.sm_ti32 = {n: cfg.T_cam2root32(table_cam[n]) for n in self.sm_names}
        self.centre32 = {n: cfg.centre32(table_cam[n]) for n in self.sm_names}
        self.K64 = {n: cfg.K64_from32(i) for i, n in enumerate(cfg.camera_names)}
        self.dist64 = {n: cfg.dist64(i) for i, n in enumerate(cfg.camera_names)}
        self.P64 = {n: cfg.P64(i) for i, n in enumerate(cfg.camera_names)}


def head_feature_row(skeleton: dict, camera: str, tabs: CameraTables) -> Tuple[np.ndarray, int]:
    """HumanGraphFromView.initializeWithAlternative3 (graph_generator.py:444-508): one head row."""
    cfg = tabs.cfg
    F = cfg.n_features_sm
    row = np.zeros(F, dtype=f32)
    row[0] = 1.0                                                     # :471 'head'
    c = tabs.sm_names.index(camera)                                  # :232
    W, H = cfg.image_width, cfg.image_height
    centre = tabs.centre32[camera]
    num_joints = 0
    for j, values in skeleton.items():                               # :493
        if j == "ID":
            continue
        x, y = values[1], values[2]
        ray_cam = _matvec_fma32(tabs.kinv32[camera], [f32(x), f32(y), f32(1.0)])       # :488
        ray = _matvec_fma32(tabs.sm_ti32[camera], [ray_cam[0], ray_cam[1], ray_cam[2], f32(0.0)])  # :489
        base = 2 + 180 * c + 10 * int(j)                             # FEATURES['3'] order :128-140
        row[base + 0] = (x - W / 2) / (W / 2)                        # :496 python float64 -> fp32 store
        row[base + 1] = (H / 2 - y) / (H / 2)                        # :497
        row[base + 2] = values[3]
        row[base + 3] = values[4]
        row[base + 4:base + 7] = centre[0:3]                         # :500-502
        row[base + 7:base + 10] = ray[0:3]                           # :503-505
        num_joints += 1
    return row, num_joints


def build_graph(frame: Dict[str, list], tabs: CameraTables) -> Optional[dict]:
    """MergedMultipleHumansDataset.process_test for one frame (graph_generator.py:813-876,
    573-605, 627-656). `frame` = {camera: [json_string, ...]}. Returns None when no edge-node exists
    (the reference then appends no graph, :866)."""
    cfg = tabs.cfg
    rows, nodes_camera, heads_json, skeleton_index = [], [], {}, {}
    view_heads: Dict[str, List[int]] = {}
    head_id = 0
    for camera in frame:                                              # :583 frame-dict order
        if camera not in tabs.sm_names:
            continue
        view_heads[camera] = []
        for idx, skeleton in enumerate(json.loads(frame[camera][0])):
            row, nj = head_feature_row(skeleton, camera, tabs)
            if nj == 0:                                               # :590
                continue
            rows.append(row)
            view_heads[camera].append(head_id)
            heads_json[head_id] = skeleton
            skeleton_index[head_id] = idx
            nodes_camera.append(tabs.sm_names.index(camera))
            head_id += 1
    H = head_id
    src = list(range(H))                                              # self loops :474-475
    dst = list(range(H))
    pairs = []
    items = list(view_heads.items())
    id_node = H
    for h_index, (cam1, heads1) in enumerate(items):                  # :854-864
        for cam2, heads2 in items[h_index + 1:]:
            for h1 in heads1:
                for h2 in heads2:
                    src += [h1, id_node, h2, id_node, id_node]        # :632-649
                    dst += [id_node, h1, id_node, h2, id_node]
                    pairs.append((h1, h2))
                    id_node += 1
    M = len(pairs)
    if M == 0:
        return None
    F = cfg.n_features_sm
    feats = np.zeros((H + M, F), dtype=f32)
    if H:
        feats[:H] = np.stack(rows)
    feats[H:, 1] = 1.0                                                # :630 'edge_node'
    return dict(src=np.array(src, dtype=np.int32), dst=np.array(dst, dtype=np.int32), n_nodes=H + M, n_heads=H,
                feats=feats, indices=np.arange(H, H + M, dtype=np.int64),
                nodes_camera=np.array(nodes_camera + [-1] * M, dtype=np.int32),
                pairs=np.array(pairs, dtype=np.int32).reshape(-1, 2), heads_json=heads_json,
                skeleton_index=skeleton_index,
                rel_type=np.array([0] * H + [1, 1, 1, 1, 2] * M, dtype=np.int64))   # RELATIONS['3'] sorted :205-211


# ----------------------------------------------------------------------------------------------
# Stage 1, training side: sample synthesis + graph of process_training (graph_generator.py:516-560, 672-810;
# utils/data_augmentation.py:14-89). Forward-only scope: the graphs, labels and index lists the training /
# validation drivers batch with dgl.batch (train_skeleton_matching.py:67-84, sm_metrics_without_gt.py:46-64).
# ----------------------------------------------------------------------------------------------
def augment_views(json_data: List[dict], used_cameras: Sequence[str], min_views: int = 1) -> List[dict]:
    """add_data_to_json (data_augmentation.py:50-89): every sample restricted to the used cameras that saw
    something, followed by each proper camera subset of it with at least min_views cameras, subsets in
    itertools.product(range(2)) order over the used-camera list (permutations_generator, :14-27)."""
    import itertools
    out = []
    for data in json_data:
        flags = [0] * len(used_cameras)
        base = {}
        for c in data:                                                        # dict order of the sample survives
            if c in used_cameras and json.loads(data[c][0]):
                flags[list(used_cameras).index(c)] = 1
                base[c] = data[c]
        if sum(flags) == 0:
            continue
        out.append(dict(base))
        for comb in itertools.product(range(2), repeat=len(flags)):
            if any(f - c < 0 for f, c in zip(flags, comb)) or sum(comb) < min_views or tuple(flags) == comb:
                continue
            out.append({c: base[c] for c in base if comb[list(used_cameras).index(c)]})
    return out


def load_training_inputs(files_json: List[List[dict]], mode: str, used_cameras: Sequence[str], rnd) -> Tuple[list, list]:
    """MergedMultipleHumansDataset.__init__, list branch (:526-538): per file, optional augmentation
    (mode not test / test_generated), index list, random.shuffle of it (mode != 'test'). `rnd` is the
    `random` module (or a random.Random) - the reference draws from the global one."""
    inputs, indices = [], []
    for data in files_json:
        if mode != 'test' and mode != 'test_generated':
            data = augment_views(data, used_cameras, 2)
        idx = list(range(len(data)))
        if mode != 'test':
            rnd.shuffle(idx)
        inputs.append(data)
        indices.append(idx)
    return inputs, indices


def training_samples(inputs, inputs_indices, probabilities, limit: int, rnd):
    """sample_and_remove (:675-697): up to `limit` tuples of single-person samples; the num_people files with
    the largest probabilities (np.argpartition order) each give up the sample at the end of their shuffled list."""
    for _ in range(limit):
        if all(len(l) == 0 for l in inputs):
            break
        num_people = rnd.randint(1, len(inputs))
        max_indx = np.argpartition(np.array(probabilities), -num_people)[-num_people:]
        views = []
        for index in max_indx:
            if not inputs_indices[index]:
                return                                                        # IndexError in the reference (:691-692)
            views.append(inputs[index][inputs_indices[index].pop()])
        if views:
            yield views


def build_training_graph(multi_person: List[Dict[str, list]], tabs: CameraTables) -> Optional[dict]:
    """One iteration of process_training (:699-810): heads of every single-person sample (the skeleton with
    most joints per camera is the person, the others are spurious), then edge-nodes in the order
    true links of person 0, its false links to the other people, its false links to spurious heads, person 1 ...,
    finally spurious x spurious - ordered pairs, so (a, b) and (b, a) both exist."""
    cfg = tabs.cfg
    rows, nodes_camera = [], []
    people, spurious = [], []
    total = 0
    for sample in multi_person:
        view_heads, view_joints = {}, {}
        head_id = 0
        for camera in sample:                                                 # load_people_view_graph (:573-605)
            if camera not in tabs.sm_names:
                continue
            view_heads[camera], view_joints[camera] = [], []
            for skeleton in json.loads(sample[camera][0]):
                row, nj = head_feature_row(skeleton, camera, tabs)
                if nj == 0:
                    continue
                rows.append(row)
                view_heads[camera].append(head_id)
                view_joints[camera].append(nj)
                nodes_camera.append(tabs.sm_names.index(camera))
                head_id += 1
        person = []
        for camera in sample:                                                 # :722-730
            if camera in tabs.sm_names and view_joints[camera]:
                heads_cam, joints_cam = view_heads[camera], view_joints[camera]
                good = max(enumerate(joints_cam), key=lambda x: x[1])[0]
                spurious += [(x + total, camera) for x in heads_cam if x != heads_cam[good]]
                person.append((heads_cam[good] + total, camera))
        people.append(person)
        total += head_id
    H = total
    pairs, labels = [], []
    for ip, person in enumerate(people):
        for h1, c1 in person:                                                 # :755-764
            for h2, c2 in person:
                if c1 != c2:
                    pairs.append((h1, h2)); labels.append(1.)
        for io, other in enumerate(people):                                   # :766-778
            if io == ip:
                continue
            for h1, c1 in person:
                for h2, c2 in other:
                    if c1 != c2:
                        pairs.append((h1, h2)); labels.append(0.)
        for h1, c1 in person:                                                 # :780-789
            for h2, c2 in spurious:
                if c1 != c2:
                    pairs.append((h1, h2)); labels.append(0.)
    for h1, c1 in spurious:                                                   # :791-800
        for h2, c2 in spurious:
            if c1 != c2:
                pairs.append((h1, h2)); labels.append(0.)
    M = len(pairs)
    if M == 0:
        return None                                                           # :802
    src, dst = list(range(H)), list(range(H))
    for k, (h1, h2) in enumerate(pairs):
        e = H + k
        src += [h1, e, h2, e, e]
        dst += [e, h1, e, h2, e]
    feats = np.zeros((H + M, cfg.n_features_sm), dtype=f32)
    feats[:H] = np.stack(rows)
    feats[H:, 1] = 1.0
    return dict(src=np.array(src, dtype=np.int32), dst=np.array(dst, dtype=np.int32), n_nodes=H + M, n_heads=H, feats=feats,
                labels=np.array(labels, dtype=np.float64).reshape(-1, 1), indices=np.arange(H, H + M, dtype=np.int64),
                nodes_camera=np.array(nodes_camera + [-1] * M, dtype=np.int32), pairs=np.array(pairs, dtype=np.int32).reshape(-1, 2),
                rel_type=np.array([0] * H + [1, 1, 1, 1, 2] * M, dtype=np.int64))


def batch_graphs(graphs: List[dict]) -> dict:
    """dgl.batch (train_skeleton_matching.py:80): block-diagonal union, node and edge ids shifted graph by graph."""
    src, dst, feats, off = [], [], [], 0
    for g in graphs:
        src.append(g['src'] + off); dst.append(g['dst'] + off); feats.append(g['feats'])
        off += g['n_nodes']
    return dict(src=np.concatenate(src), dst=np.concatenate(dst), feats=np.concatenate(feats), n_nodes=off)


# ----------------------------------------------------------------------------------------------
# Stage 2a: GAT forward  (skeleton_matching/gat2.py)
# ----------------------------------------------------------------------------------------------
def leaky(x, slope):
    return np.where(x >= 0, x, x * f32(slope)).astype(f32)


def gat_layer(x, src, dst, W1, b1, W2, b2, attn_l, attn_r, heads, alpha, res=None):
    """GraphAttention2.forward (gat2.py:50-76) with DGL's edge_softmax(norm_by='dst') and
    update_all(u_mul_e, sum) semantics (gat2.py:61-66, 78-88); dropout is the identity at inference.
    res: None (residual=False), (W, b) = res_fc, or 'identity' (in_dim == out_dim), gat2.py:70-75."""
    n = x.shape[0]
    ft1 = x @ W1.T + b1                                               # :53
    h2 = leaky(ft1, alpha)                                            # :54
    ft2 = (h2 @ W2.T + b2).astype(f32).reshape(n, heads, -1)          # :55
    a1 = np.einsum('nhd,hd->nh', ft2, attn_l[:, :, 0]).astype(f32)    # :57
    a2 = np.einsum('nhd,hd->nh', ft2, attn_r[:, :, 0]).astype(f32)    # :58
    e = leaky(a1[src] + a2[dst], alpha)                               # :80
    mx = np.full((n, heads), -np.inf, dtype=f32)
    np.maximum.at(mx, dst, e)
    ex = np.exp(e - mx[dst]).astype(f32)
    den = np.zeros((n, heads), dtype=f32)
    np.add.at(den, dst, ex)
    a = (ex / den[dst]).astype(f32)                                   # :84
    out = np.zeros_like(ft2)
    np.add.at(out, dst, ft2[src] * a[:, :, None])                     # :66
    if res is not None:
        if isinstance(res, str):
            resval = x.astype(f32)[:, None, :]                            # :74
        else:
            resval = (x @ res[0].T + res[1]).astype(f32).reshape(n, heads, -1)   # :72
        out = (resval + out).astype(f32)                                  # :75
    return out


def gat_forward(weights: dict, feats, src, dst, heads=(10, 10, 8, 5, 1), alpha=0.15, act_slope=0.01,
                return_layers=False, residual=False):
    """GAT2.forward (gat2.py:137-149): LeakyReLU(0.01) between layers, sigmoid at the end
    (train_skeleton_matching.py:54,148-149). weights: reference state_dict as numpy arrays.
    residual=True: every layer after the first adds res_fc(h) when its weights exist, h itself otherwise (gat2.py:43-48)."""
    h = feats.astype(f32)
    layers = []
    L = len(heads)
    for l in range(L):
        p = lambda k: weights['layers.%d.%s' % (l, k)]
        res = None
        if residual and l > 0:
            res = (p('res_fc.weight'), p('res_fc.bias')) if ('layers.%d.res_fc.weight' % l) in weights else 'identity'
        out = gat_layer(h, src, dst, p('fc1.weight'), p('fc1.bias'), p('fc2.weight'), p('fc2.bias'),
                        p('attn_l'), p('attn_r'), heads[l], alpha, res)
        layers.append(out)
        if l < L - 1:
            h = leaky(out.reshape(out.shape[0], -1), act_slope)       # :141-142
    logits = layers[-1].reshape(-1)
    scores = (1.0 / (1.0 + np.exp(-logits.astype(np.float64)))).astype(f32)
    return (scores, layers) if return_layers else scores


# ----------------------------------------------------------------------------------------------
# Stage 2b: person proposals  (utils/skeleton_matching_utils.py:12-132), array form.
# CPython set layout is emulated explicitly so the CUDA kernel can be a transliteration.
# ----------------------------------------------------------------------------------------------
LINEAR_PROBES = 9
PERTURB_SHIFT = 5


class IntSet:
    """CPython 3.12 set of small non-negative ints (hash(i) == i): Objects/setobject.c
    set_add_entry / set_insert_clean / set_table_resize. Only insert + iterate are needed."""

    def __init__(self):
        self.mask = 7
        self.table = [-1] * 8
        self.fill = 0

    @staticmethod
    def _probe_insert(table, mask, key):
        perturb = key
        i = key & mask
        while True:
            if table[i] < 0:
                table[i] = key
                return True
            if table[i] == key:
                return False
            if i + LINEAR_PROBES <= mask:
                for j in range(1, LINEAR_PROBES + 1):
                    if table[i + j] < 0:
                        table[i + j] = key
                        return True
                    if table[i + j] == key:
                        return False
            perturb >>= PERTURB_SHIFT
            i = (i * 5 + 1 + perturb) & mask

    def add(self, key):
        if not self._probe_insert(self.table, self.mask, key):
            return
        self.fill += 1
        if self.fill * 5 < self.mask * 3:
            return
        minused = self.fill * 4 if self.fill <= 50000 else self.fill * 2
        newsize = 8
        while newsize <= minused:
            newsize <<= 1
        old = self.table
        self.table = [-1] * newsize
        self.mask = newsize - 1
        for k in old:                      # re-insert in old-table order
            if k >= 0:
                self._probe_insert(self.table, self.mask, k)

    def __iter__(self):
        return (k for k in self.table if k >= 0)


def pair_order(h1, h2):
    """list({h1, h2}) where the set was built as set([h1]); add(h2) (skeleton_matching_utils.py:49-55)."""
    s = IntSet()
    s.add(h1)
    s.add(h2)
    return list(s)


def cluster(scores: Sequence[float], pairs: np.ndarray, head_cam: Sequence[int], n_cams: int,
            n_heads: int, thr: float = 0.5, min_views: int = 2) -> np.ndarray:
    """get_person_proposal_from_network_output on a graph produced by build_graph().

    scores: N values (fp32 semantics: compared as python floats converted from fp32, strict >, :52)
    pairs[k] = (h1, h2) of edge-node n_heads + k, in edge-node order.
    Returns int32 [persons, n_cams] with -1 for None, persons in component order (:117-130).
    """
    H = n_heads
    # ---- edge walk (:32-55). Edge order per edge-node: (e->h1) then (e->h2); only those count.
    first_seen: List[int] = []
    seen = [False] * H
    matchings = []
    for k, (h1, h2) in enumerate(pairs):
        h1 = int(h1); h2 = int(h2)
        for h in (h1, h2):
            if not seen[h]:
                seen[h] = True
                first_seen.append(h)
        s = float(f32(scores[H + k]))
        if s > thr:
            matchings.append((k, pair_order(h1, h2), s))
    # ---- greedy merge in score order, stable (:60-108). cams as bitmasks (membership only).
    order = sorted(range(len(matchings)), key=lambda i: -matchings[i][2])   # stable: ties keep edge-node order
    linked = [1 << int(head_cam[h]) for h in range(H)]          # heads_linked_in_cameras (:46)
    group = [-1] * H                                            # human_index
    cams_for: Dict[int, int] = {}
    adj: List[List[int]] = [[] for _ in range(H)]
    cur = 0
    for i in order:
        a, b = matchings[i][1]
        ca, cb = 1 << int(head_cam[a]), 1 << int(head_cam[b])
        if (ca & linked[b]) or (cb & linked[a]):                # :67
            continue
        if group[a] >= 0 and (cb & cams_for[group[a]]):         # :70-72
            continue
        if group[b] >= 0 and (ca & cams_for[group[b]]):         # :73-75
            continue
        if group[a] < 0 and group[b] < 0:                       # :77-83
            group[a] = group[b] = cur
            cams_for[cur] = ca | cb
            cur += 1
        elif group[a] >= 0 and group[b] < 0:                    # :84-86
            group[b] = group[a]
            cams_for[group[a]] |= cb
        elif group[b] >= 0 and group[a] < 0:                    # :87-89
            group[a] = group[b]
            cams_for[group[b]] |= ca
        else:                                                   # :90-104
            if cams_for[group[b]] & cams_for[group[a]]:
                continue
            new, old = group[a], group[b]
            if new != old:
                for n in range(H):
                    if group[n] == old:
                        group[n] = new
                del cams_for[old]                               # absorbed cameras are forgotten (quirk)
            # new == old cannot pass the intersection test (a group always has >=1 camera)
        adj[a].append(b)                                        # :106
        adj[b].append(a)
        linked[a] |= cb                                         # :107-108
        linked[b] |= ca
    # ---- connected components in node-insertion order, networkx 3.x _plain_bfs (:118)
    out = []
    done = [False] * H
    for v in first_seen:
        if done[v]:
            continue
        comp = IntSet()
        comp.add(v)
        done[v] = True
        level = [v]
        size = 1
        while level:
            nxt = []
            for x in level:
                for w in adj[x]:
                    if not done[w]:
                        done[w] = True
                        comp.add(w)
                        size += 1
                        nxt.append(w)
            level = nxt
        if size < min_views:                                    # :120
            continue
        person = [-1] * n_cams
        for x in comp:                                          # set iteration order (:127-128)
            person[int(head_cam[x])] = x
        out.append(person)
    return np.array(out, dtype=np.int32).reshape(-1, n_cams)


# ----------------------------------------------------------------------------------------------
# Stage 3: MLP-input encoder, pairwise DLT, triangulation baseline, MLP
# ----------------------------------------------------------------------------------------------
def undistort_point(u, v, K, dist):
    """cv2.undistortPoints(pt, K, dist) without R/P: 5 fixed-point iterations of the Brown model
    inverse in float64 (OpenCV undistort.dispatch.cpp cvUndistortPointsInternal, default criteria
    COUNT=5). dist = [k1,k2,p1,p2,k3]."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    k1, k2, p1, p2, k3 = dist
    x = (u - cx) * (1.0 / fx)
    y = (v - cy) * (1.0 / fy)
    x0, y0 = x, y
    for _ in range(5):
        r2 = x * x + y * y
        icdist = 1.0 / (1 + ((k3 * r2 + k2) * r2 + k1) * r2)
        if icdist < 0:
            return x0, y0
        dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        x = (x0 - dx) * icdist
        y = (y0 - dy) * icdist
    return x, y


def triangulate_pair(P1, P2, pt1, pt2):
    """cv2.triangulatePoints for one point: null vector of the 4x4 DLT matrix (OpenCV
    triangulate.cpp: rows x*P[2]-P[0], y*P[2]-P[1] per view; SVD, last row of Vt), dehomogenised
    (pose_estimator_dataset_from_json.py:95-96, pose_estimator_utils.py:66-67)."""
    A = np.empty((4, 4))
    A[0] = pt1[0] * P1[2] - P1[0]
    A[1] = pt1[1] * P1[2] - P1[1]
    A[2] = pt2[0] * P2[2] - P2[0]
    A[3] = pt2[1] * P2[2] - P2[1]
    _, _, Vt = np.linalg.svd(A)
    X = Vt[3]
    return X[0:3] / X[3]


def hint_from_triangulation(person: Dict[str, dict], tabs: CameraTables) -> Dict[int, np.ndarray]:
    """get_3D_from_triangulation (pose_estimator_dataset_from_json.py:63-101).
    person = {camera: skeleton} in `used_cameras` order, one skeleton per camera."""
    points_2D: Dict[str, Dict[str, Tuple[float, float]]] = {}
    for cam, sk in person.items():
        if cam not in tabs.pe_names:
            continue
        for j, pos in sk.items():
            if j == "ID":
                continue
            if pos[0] > 0.:                                          # :75 joint id > 0 (quirk)
                points_2D.setdefault(j, {})[cam] = (pos[1], pos[2])
    result = {}
    for idx_i in range(N_JOINTS):
        idx = str(idx_i)
        if idx in points_2D and len(points_2D[idx]) > 1:
            cams = list(points_2D[idx].keys())
            acc = np.zeros(3)
            n = 0
            for i, k in itertools.combinations(range(len(cams)), 2):
                c1, c2 = cams[i], cams[k]
                u1 = undistort_point(*points_2D[idx][c1], tabs.K64[c1], tabs.dist64[c1])
                u2 = undistort_point(*points_2D[idx][c2], tabs.K64[c2], tabs.dist64[c2])
                acc += triangulate_pair(tabs.P64[c1], tabs.P64[c2], u1, u2)
                n += 1
            result[idx_i] = acc / n                                  # :100 plain mean
    return result


def encode_person(person: Dict[str, dict], tabs: CameraTables) -> Optional[np.ndarray]:
    """PoseEstimatorDataset.__init__ dict branch (pose_estimator_dataset_from_json.py:237-289).
    Returns the fp32 vector of length 252*len(used_cameras), or None when sum|v| <= 1 (:287)."""
    cfg = tabs.cfg
    tri = hint_from_triangulation(person, tabs)
    out = np.zeros(cfg.mlp_in, dtype=f32)
    W2, H2 = cfg.image_width / 2, cfg.image_height / 2
    for cam, sk in person.items():
        if cam not in tabs.pe_names:
            continue
        off = tabs.pe_names.index(cam) * (N_JOINTS * 14)
        Ti = tabs.ti32[cam]
        centre = (Ti[:, 3] / f32(10.)).astype(f32)                    # :247 (T . [0,0,0,1] = last column)
        for j, values in sk.items():
            if j == "ID":
                continue
            jo = off + int(j) * 14
            ux, uy = undistort_point(values[1], values[2], tabs.K64[cam], tabs.dist64[cam])   # :261
            ray = _matvec_fma32(Ti, [f32(ux), f32(uy), f32(1.0), f32(0.0)])                   # :262-264
            ray = (ray / f32(10.)).astype(f32)
            out[jo + 0] = values[3]                                   # :271
            out[jo + 1] = (values[1] - W2) / W2                       # :259-260,272 (float64 -> fp32)
            out[jo + 2] = (values[2] - H2) / H2
            out[jo + 3] = values[4]
            out[jo + 4:jo + 7] = centre[0:3]
            out[jo + 7:jo + 10] = ray[0:3]
    for c_index in range(cfg.V_pe):                                   # :280-285 (all used cameras: quirk)
        off = c_index * (N_JOINTS * 14)
        for j, X in tri.items():
            jo = off + j * 14
            out[jo + 10] = 1.
            out[jo + 11:jo + 14] = X / 10.
    if float(np.sum(np.abs(out))) > 1:                                # :287
        return out
    return None


def triangulate_baseline(person: Dict[str, dict], tabs: CameraTables, median_axis: int):
    """triangulate() (utils/pose_estimator_utils.py:52-75) fed as metrics_from_triangulation.py:237-249
    does: every joint of every matched skeleton, no validity filter. Returns ([18,3] float64, mask)."""
    points_2D: Dict[str, Dict[str, Tuple[float, float]]] = {}
    for cam, sk in person.items():
        for j, values in sk.items():
            points_2D.setdefault(j, {})[cam] = (values[1], values[2])
    res = np.zeros((N_JOINTS, 3))
    mask = np.zeros(N_JOINTS, dtype=np.uint8)
    for idx_i in range(N_JOINTS):
        idx = str(idx_i)
        if idx in points_2D and len(points_2D[idx]) > 1:
            cams = list(points_2D[idx].keys())
            pts = []
            for i, k in itertools.combinations(range(len(cams)), 2):
                c1, c2 = cams[i], cams[k]
                u1 = undistort_point(*points_2D[idx][c1], tabs.K64[c1], tabs.dist64[c1])
                u2 = undistort_point(*points_2D[idx][c2], tabs.K64[c2], tabs.dist64[c2])
                pts.append(triangulate_pair(tabs.P64[c1], tabs.P64[c2], u1, u2))
            pts = np.array(pts)
            d = pts[:, median_axis]
            med = np.sort(d)[d.shape[0] // 2]                         # :71 upper median
            keep = np.abs(d - med) < 0.05                             # :72-73
            res[idx_i] = pts[keep].mean(axis=0)                       # :74
            mask[idx_i] = 1
    return res, mask


def mlp_forward(weights: dict, x: np.ndarray) -> np.ndarray:
    """PoseEstimatorMLP.forward (utils/mlp.py:8-31): 9 Linear, LeakyReLU(0.1) between."""
    h = x.astype(f32)
    for i, l in enumerate(range(1, 18, 2)):
        h = h @ weights['layers.%d.weight' % l].T + weights['layers.%d.bias' % l]
        if i < 8:
            h = leaky(h, 0.1)
    return h.astype(f32)


# ----------------------------------------------------------------------------------------------
# One frame end to end (the per-frame glue of test/metrics_from_model.py:178-300)
# ----------------------------------------------------------------------------------------------
def infer_frame(frame, tabs: CameraTables, gat_w: dict, mlp_w: dict, thr: float = 0.5):
    processed = {c: [frame[c][0]] for c in frame if json.loads(frame[c][0])}     # :182-191
    g = build_graph(processed, tabs)
    if g is None:
        return None
    scores = gat_forward(gat_w, g['feats'], g['src'], g['dst'])
    props = cluster(scores, g['pairs'], g['nodes_camera'][:g['n_heads']], tabs.cfg.V_sm, g['n_heads'], thr,
                    tabs.cfg.min_number_of_views)
    mlp_in = []
    for person in props:
        p = {}
        for cam in tabs.pe_names:                                                # :248-252
            if cam in tabs.sm_names and person[tabs.sm_names.index(cam)] >= 0:
                p[cam] = g['heads_json'][int(person[tabs.sm_names.index(cam)])]
        v = encode_person(p, tabs)
        if v is not None:
            mlp_in.append(v)
    joints = mlp_forward(mlp_w, np.stack(mlp_in)) * f32(10.) if mlp_in else np.zeros((0, 54), f32)   # :280-282
    return dict(graph=g, scores=scores, proposals=props, mlp_in=mlp_in, joints=joints)
