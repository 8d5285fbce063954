"""Rectangle obstacles (reference env/obstacle.py:1-45): API surface only.

The constraint flags used for training come from the CUDA step kernel (csrc/env.cu nav_obstacle), which
applies the same closed-interval test in fp64; these classes keep `env.obstacle(state)` callable."""
import numpy as np


class Obstacle(object):
    def __init__(self, boundsx, boundsy, penalty=100):
        self.boundsx = boundsx
        self.boundsy = boundsy
        self.penalty = 1

    def __call__(self, x):
        return (self.boundsx[0] <= x[0] <= self.boundsx[1] and self.boundsy[0] <= x[1] <= self.boundsy[1]) * self.penalty


class ComplexObstacle(Obstacle):
    def __init__(self, bounds):
        self.obs = [Obstacle(bx, by) for bx, by in bounds]

    def __call__(self, x):
        return np.max([o(x) for o in self.obs])
