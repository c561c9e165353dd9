"""Drop-in for the reference's utils/mlp.py (PoseEstimatorMLP :3-31): same module tree, so checkpoints with keys
`layers.{1,3,...,17}.{weight,bias}` load unchanged; forward runs the 9 projections on the B200 tensor cores."""
from torch import nn

import _b200pose_runtime as rt


class PoseEstimatorMLP(nn.Module):
    def __init__(self, input_dimensions, output_dimensions):
        super().__init__()
        print('MLP input size', input_dimensions)
        negative_slope = 0.1
        self.negative_slope = negative_slope
        self.layers = nn.Sequential(
            nn.Flatten(),
            nn.Linear(input_dimensions, 3072), nn.LeakyReLU(negative_slope=negative_slope),
            nn.Linear(3072, 3072), nn.LeakyReLU(negative_slope=negative_slope),
            nn.Linear(3072, 2048), nn.LeakyReLU(negative_slope=negative_slope),
            nn.Linear(2048, 2048), nn.LeakyReLU(negative_slope=negative_slope),
            nn.Linear(2048, 1024), nn.LeakyReLU(negative_slope=negative_slope),
            nn.Linear(1024, 1024), nn.LeakyReLU(negative_slope=negative_slope),
            nn.Linear(1024, 1024), nn.LeakyReLU(negative_slope=negative_slope),
            nn.Linear(1024, 1024), nn.LeakyReLU(negative_slope=negative_slope),
            nn.Linear(1024, output_dimensions),
        )
        self._prepared = None
        self._prepared_key = None

    def forward(self, x):
        ctx = rt.context()
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._prepared is None or key != self._prepared_key:
            self._prepared = ctx.prepare_mlp({k: v for k, v in self.state_dict().items()})
            self._prepared_key = key
        x2 = x.reshape(x.shape[0], -1).to(ctx.device).float()
        planes = rt.pipeline.Planes.from_f32(x2, ctx._stream())
        out = ctx.mlp_forward(planes, x2.shape[0], scale=1.0, layers=self._prepared, slope=self.negative_slope)
        return out.clone().to(x.device)
