"""Drop-in for the reference's skeleton_matching/gat2.py (GraphAttention2 :17-88, GAT2 :90-154).

Same constructors, parameter names (state_dict keys `layers.{l}.{attn_l,attn_r,fc1.*,fc2.*[,res_fc.*]}`) and
initialisation order; `forward` runs on the B200 kernels (split-bf16 tcgen05 projections + the fused
edge-softmax/aggregation kernel) instead of DGL; under autograd (grad enabled, trainable parameters) `loss.backward()` runs
the hand-written backward of csrc/train.cu. Dropout must be 0. Residual layers (gat2.py:70-75, not the
shipped configuration) add res_fc(h) - or h itself when the widths agree - inside the gather aggregation kernel.
"""
import torch
import torch.nn as nn

import _b200pose_runtime as rt


class _GatTrainFunction(torch.autograd.Function):
    """GAT2.forward under autograd: the forward keeps its activations inside a GatGrad (3d_multi_pose_estimator_b200/train.py),
    `loss.backward()` (train_skeleton_matching.py:181) lands here and the parameter gradients come from csrc/train.cu + the
    tensor-core GEMM instead of torch's autograd graph of gat2.py:50-88."""

    @staticmethod
    def forward(ctx, inputs, holder, *params):
        net, db, arrays, x0, sigmoid, names = holder
        scores = net.forward(db, arrays, x0)
        if not sigmoid:
            raise NotImplementedError('B200 GAT2 training path: final_activation must be nn.Sigmoid (train_skeleton_matching.py:33)')
        out = scores.clone()
        ctx.holder = holder
        ctx.save_for_backward(out)
        return out.reshape(-1, 1, 1)

    @staticmethod
    def backward(ctx, dout):
        net = ctx.holder[0]
        (scores,) = ctx.saved_tensors
        n = scores.shape[0]
        if net.cache is None or net.cache['N'] != n or net.cache_token is not ctx.holder:
            raise RuntimeError('B200 GAT2: backward() after another forward of the same model (the drop-in keeps one set of activations)')
        dl = net.buf.f('dlogit', n, 1)
        d = dout.reshape(-1).contiguous().float()
        rt.pipeline.check(net.L.b200pose_sigmoid_bwd(rt.pipeline.ptr(scores), rt.pipeline.ptr(d), n, rt.pipeline.ptr(dl), net.pipe._stream()),
                          'sigmoid_bwd')
        net.backward(dl)
        grads = net.grads()
        return (None, None) + tuple(grads[k].reshape(shape).clone() for k, shape in ctx.holder[5])


class GraphAttention2(nn.Module):
    def __init__(self, g, in_dim, out_dim, num_heads, feat_drop, attn_drop, alpha, residual=False, name=None, bias=False):
        super(GraphAttention2, self).__init__()
        self.g = g
        self.num_heads = num_heads
        self.name = name
        self.fc1 = nn.Linear(in_dim, in_dim, bias=bias)
        self.fc2 = nn.Linear(in_dim, num_heads * out_dim, bias=bias)
        self.feat_drop_p = feat_drop
        self.attn_drop_p = attn_drop
        self.attn_l = nn.Parameter(torch.Tensor(size=(num_heads, out_dim, 1)))
        self.attn_r = nn.Parameter(torch.Tensor(size=(num_heads, out_dim, 1)))
        nn.init.xavier_normal_(self.fc1.weight.data, gain=1.414)
        nn.init.xavier_normal_(self.fc2.weight.data, gain=1.414)
        nn.init.xavier_normal_(self.attn_l.data, gain=1.414)
        nn.init.xavier_normal_(self.attn_r.data, gain=1.414)
        self.alpha = alpha
        self.leaky_relu = nn.LeakyReLU(alpha)
        self.residual = residual
        if residual:
            if in_dim != out_dim:
                self.res_fc = nn.Linear(in_dim, num_heads * out_dim, bias=bias)
                nn.init.xavier_normal_(self.res_fc.weight.data, gain=1.414)
            else:
                self.res_fc = None

    def forward(self, inputs):
        raise NotImplementedError('GraphAttention2 layers are executed by GAT2.forward on the B200 path')


class GAT2(nn.Module):
    def __init__(self, g, num_layers, in_dim, num_classes, num_hidden, heads, activation, final_activation, feat_drop,
                 attn_drop, alpha, residual, bias=False):
        super(GAT2, self).__init__()
        self.g = g
        self.num_layers = num_layers
        self.layers = nn.ModuleList()
        self.activation = activation
        self.final_activation = final_activation
        self.alpha = alpha
        self.layers.append(GraphAttention2(g, in_dim, num_hidden[0], heads[0], feat_drop, attn_drop, alpha, False, '0', bias))
        for l in range(1, num_layers - 1):
            self.layers.append(GraphAttention2(g, num_hidden[l - 1] * heads[l - 1], num_hidden[l], heads[l], feat_drop,
                                               attn_drop, alpha, residual, str(l), bias))
        self.layers.append(GraphAttention2(g, num_hidden[-1] * heads[-1], num_classes, 1, feat_drop, attn_drop, alpha,
                                           residual, 'X', bias))
        self._prepared = None
        self._prepared_key = None

    def set_g(self, g):
        self.g = g
        for l in range(self.num_layers):
            self.layers[l].g = g

    def _check_supported(self):
        for lyr in self.layers:
            if lyr.feat_drop_p or lyr.attn_drop_p:
                raise NotImplementedError('B200 GAT2 is inference-only: feat_drop/attn_drop must be 0')
        if not isinstance(self.activation, nn.LeakyReLU):
            raise NotImplementedError('B200 GAT2: the inter-layer activation must be nn.LeakyReLU')
        if self.final_activation is not None and not isinstance(self.final_activation, nn.Sigmoid):
            raise NotImplementedError('B200 GAT2: final_activation must be nn.Sigmoid or None')
        if self.layers[-1].fc2.out_features != 1:
            raise NotImplementedError('B200 GAT2: num_classes must be 1')

    def _weights(self, ctx):
        plist = self.__dict__.get('_plist')
        if plist is None:                       # the module-tree walk of parameters() is most of a call: done once
            plist = list(self.parameters())     # (.to() and load_state_dict keep the Parameter objects)
            self.__dict__['_plist'] = plist
        key = tuple((p.data_ptr(), p._version) for p in plist)
        if self._prepared is None or key != self._prepared_key:
            rt.sync_live()
            self._prepared = ctx.prepare_gat({k: v for k, v in self.state_dict().items()},
                                              residual=any(lyr.residual for lyr in self.layers))
            self._prepared_key = key
        return self._prepared

    def _forward_train(self, inputs, g, ctx):
        """Grad mode with trainable parameters (the loop of train_skeleton_matching.py:163-184)."""
        if inputs.requires_grad:
            raise NotImplementedError('B200 GAT2: gradients with respect to the input features are not computed (the reference trains the '
                                      'parameters only, train_skeleton_matching.py:150)')
        named = list(self.named_parameters())
        key = tuple((p.data_ptr(), p._version) for _, p in named)
        net = self.__dict__.get('_grad_net')
        rt.sync_live()
        if net is None:
            train_mod = rt.importlib.import_module('3d_multi_pose_estimator_b200.train')
            net = train_mod.GatGrad(ctx, {k: v for k, v in self.state_dict().items()}, self.alpha, self.activation.negative_slope,
                                    residual=any(lyr.residual for lyr in self.layers))
            self.__dict__['_grad_net'] = net
        elif key != self.__dict__.get('_grad_key'):
            net.load_state({k: v for k, v in self.state_dict().items()})       # optimizer.step() changed them
        self.__dict__['_grad_key'] = key
        db, arrays = g._b200
        ctx.wait_graph(arrays)
        x0 = rt.pipeline.Planes.from_f32(inputs.to(ctx.device).float(), ctx._stream())
        holder = (net, db, arrays, x0, self.final_activation is not None, [(k, tuple(p.shape)) for k, p in named])
        net.cache_token = holder
        return _GatTrainFunction.apply(inputs, holder, *[p for _, p in named])

    def forward(self, inputs, g):
        self.set_g(g)
        self._check_supported()
        ctx = rt.context()
        if not hasattr(g, '_b200'):
            raise TypeError('GAT2.forward needs a graph built by the B200 graph_generator drop-in')
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._forward_train(inputs, g, ctx)
        layers = self._weights(ctx)
        # the whole-frame submissions of the dataset drop-in use these weights from now on - when the model is the shipped
        # kind (the pipeline's activation constants) - and this call is answered from the submission of its own graph
        P = rt.pipeline
        shipped = (not any(lyr.residual for lyr in self.layers) and abs(self.alpha - P.GAT_ALPHA) < 1e-12 and abs(self.activation.negative_slope - P.GAT_ACT_SLOPE) < 1e-12
                   and self.final_activation is not None)
        if shipped:
            rt.note_model('gat', self, self._prepared_key, layers)
        lv = getattr(g, '_live', None)
        if lv is not None and shipped and lv.fresh() and lv.key[0] == id(self) and lv.key[1] == self._prepared_key \
                and dict.__contains__(g.ndata, 'h') and inputs.data_ptr() == g.ndata['h'].data_ptr():
            out = lv.scores()
            if out is not None:
                lv.scores_ptr = out.data_ptr()
                return out.reshape(-1, 1, 1)
        rt.sync_live()
        db, arrays = g._b200
        feats = g.ndata['h']
        if inputs.data_ptr() == feats.data_ptr() or (inputs.shape == feats.shape and inputs.device == feats.device
                                                     and torch.equal(inputs, feats)):
            x0, dense = None, False                # the graph's own features: layer 0 runs on the S+1 compact rows
        else:
            x0 = rt.pipeline.Planes.from_f32(inputs.to(ctx.device).float(), ctx._stream())
            dense = True
        out = ctx.gat_forward(db, arrays, x0=x0, dense_rows=dense, layers=layers, alpha=self.alpha,
                              act_slope=self.activation.negative_slope,
                              final_sigmoid=self.final_activation is not None)
        return out.clone().reshape(-1, 1, 1)
