"""Drop-in for the reference's utils/mlp.py (PoseEstimatorMLP :3-31): same module tree, so checkpoints with keys
`layers.{1,3,...,17}.{weight,bias}` load unchanged; forward runs the 9 projections on the B200 tensor cores."""
import torch
from torch import nn

import _b200pose_runtime as rt

HIDDEN_WIDTHS = (3072, 3072, 2048, 2048, 1024, 1024, 1024, 1024)     # the reference's layer widths (utils/mlp.py:11-27)


class PoseEstimatorMLP(nn.Module):
    def __init__(self, input_dimensions, output_dimensions):
        super().__init__()
        print('MLP input size', input_dimensions)
        self.negative_slope = 0.1                                   # LeakyReLU slope of every hidden layer (utils/mlp.py:7)
        widths = (input_dimensions,) + HIDDEN_WIDTHS + (output_dimensions,)
        modules = [nn.Flatten()]                                    # index 0, so the projections sit at 1, 3, ..., 17
        for i in range(len(widths) - 1):
            modules.append(nn.Linear(widths[i], widths[i + 1]))
            if i < len(widths) - 2:
                modules.append(nn.LeakyReLU(negative_slope=self.negative_slope))
        self.layers = nn.Sequential(*modules)
        self._prepared = None
        self._prepared_key = None

    def forward(self, x):
        ctx = rt.context()
        plist = self.__dict__.get('_plist')
        if plist is None:                       # the module-tree walk of parameters() is done once
            plist = list(self.parameters())
            self.__dict__['_plist'] = plist
        key = tuple((p.data_ptr(), p._version) for p in plist)
        if self._prepared is None or key != self._prepared_key:
            rt.sync_live()
            self._prepared = ctx.prepare_mlp({k: v for k, v in self.state_dict().items()})
            self._prepared_key = key
        if abs(self.negative_slope - rt.pipeline.MLP_SLOPE) < 1e-12:
            rt.note_model('mlp', self, key, self._prepared)      # the dataset drop-in submits whole frames with these weights from now on
        x2 = x.reshape(x.shape[0], -1).to(ctx.device).float()
        # the rows are the MLP inputs of the proposals of the frame submitted as a whole, in the order its persons were asked
        # for: the submission already ran these 9 projections (checked by value, not assumed)
        lv = rt.last_live()
        served = getattr(lv, 'served', None) if lv is not None else None
        if served and lv.fresh() and lv.key[2] == (id(self), key) and len(served) == x2.shape[0]:
            mine = lv.enc_device(max(served) + 1)
            mine = mine if served == list(range(len(served))) else mine[served]
            torch.cuda.current_stream(ctx.device).wait_event(lv.ent.ev_b)
            if mine.shape == x2.shape and torch.equal(mine, x2):
                out = lv.mlp_out_raw(max(served) + 1)
                if out is not None:
                    lv.served = []
                    out = out if served == list(range(len(served))) else out[served]
                    return out.to(x.device)
        rt.sync_live()
        planes = rt.pipeline.Planes.from_f32(x2, ctx._stream())
        out = ctx.mlp_forward(planes, x2.shape[0], scale=1.0, layers=self._prepared, slope=self.negative_slope)
        return out.clone().to(x.device)
