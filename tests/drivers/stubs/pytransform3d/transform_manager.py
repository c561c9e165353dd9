"""Headless stand-in for pytransform3d.transform_manager.TransformManager (the reference pins 1.9.1, requirements.txt:5):
what the reference's scripts use of it - unpickling and get_transform(a, b), the inverse direction being the matrix
inverse (pytransform3d's invert_transform)."""
import numpy as np


class TransformManager(object):
    def __init__(self):
        self.transforms = {}

    def __setstate__(self, state):
        self.__dict__.update(state)

    def add_transform(self, a, b, A2B):
        self.transforms[(a, b)] = np.asarray(A2B, dtype=np.float64)

    def get_transform(self, a, b):
        if (a, b) in self.transforms:
            return self.transforms[(a, b)]
        return np.linalg.inv(self.transforms[(b, a)])
