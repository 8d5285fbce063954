"""Helpers of the PETS ensemble definition (the names the reference's config/utils.py:6-27 exports: `swish`,
`truncated_normal`, `get_affine_params`).

The initial weights are the reference's: scipy truncnorm draws on the numpy GLOBAL generator, clipped at two standard
deviations, with std 1 / (2 sqrt(fan_in)); `np.random.seed` in experiment.py:91 therefore pins them.  Biases start at
zero.  Shapes are [ensemble, fan_in, fan_out] and [ensemble, 1, fan_out] (batched `bmm` layout of config/maze.py:71-96).
"""
import math

import torch
from scipy import stats

_CLIP = 2.0   # truncation of the normal, in standard deviations


def swish(x):
    """x * sigmoid(x) (config/utils.py:6-7); the device kernels use the same expression (csrc/mpc.cu)."""
    return torch.sigmoid(x) * x


def truncated_normal(size, std):
    draws = stats.truncnorm.rvs(-_CLIP, _CLIP, size=size)
    return torch.as_tensor(draws * std).to(torch.float32)


def get_affine_params(ensemble_size, in_features, out_features):
    std = 1.0 / (2.0 * math.sqrt(in_features))
    weight = torch.nn.Parameter(truncated_normal((ensemble_size, in_features, out_features), std))
    bias = torch.nn.Parameter(torch.zeros((ensemble_size, 1, out_features), dtype=torch.float32))
    return weight, bias
