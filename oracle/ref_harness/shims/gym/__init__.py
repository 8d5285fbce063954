"""Minimal functional stand-in for `gym` (absent from this image, no network).

TEST INFRASTRUCTURE ONLY: lets the unmodified reference tree under /root/reference
import and run on CPU so golden vectors can be dumped.  gym is
unpinned in the reference (install.sh:10); Box.sample()/seed() streams here are
*a* choice, not the reference's stream, so random start actions are treated as
inputs in parity mode.
"""
import importlib
import numpy as np
from . import spaces, utils, envs  # noqa: F401
from .envs.registration import register, make  # noqa: F401


class Env(object):
    action_space = None
    observation_space = None

    def seed(self, seed=None):
        return [seed]

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError
