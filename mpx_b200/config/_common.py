"""Shared helpers for the config modules (numpy stands in for the reference's jax.numpy)."""
import os

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")


def block_diag(*blocks):
    n = sum(b.shape[0] for b in blocks)
    W = np.zeros((n, n))
    i = 0
    for b in blocks:
        k = b.shape[0]
        W[i:i + k, i:i + k] = b
        i += k
    return W


def table(name):
    """Path of the model table that replaces the reference's MJCF path (meshes are not needed)."""
    return os.path.join(DATA_DIR, name + ".json")
