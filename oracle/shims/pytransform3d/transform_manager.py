"""Test-infrastructure shim (NOT product code).

Stand-in for pytransform3d.transform_manager.TransformManager (pinned 1.9.1 in the
reference's requirements.txt:5). The reference only unpickles the object and calls
get_transform(a, b) (graph_generator.py:32,43,46; pose_estimator_dataset_from_json.py:28,38,41).
The pickled `transforms` dict stores ('root', cam) float64 4x4 matrices; the opposite
direction is the matrix inverse (pytransform3d's invert_transform) -- parity unpinned.
"""
import numpy as np


class TransformManager(object):
    def __setstate__(self, state):
        self.__dict__.update(state)

    def get_transform(self, a, b):
        if (a, b) in self.transforms:
            return self.transforms[(a, b)]
        return np.linalg.inv(self.transforms[(b, a)])
