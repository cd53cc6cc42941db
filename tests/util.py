"""Shared helpers for the tests: golden fixture loading and deterministic weights."""
import json
import os

import numpy as np
import torch

from oracle import oa_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["cfg"] = json.loads(str(d["cfg"]))
    return d


def dyn_state_dict(g, dtype=torch.float32):
    shapes = oa_ref.dynamics_param_shapes(g["cfg"], [int(x) for x in g["node_nfs"]], int(g["condition_nf"]))
    return oa_ref.make_state_dict(shapes, int(g["seed"]), g["cfg"], dtype=dtype, prefix_model="model.")


def leftnet_state_dict(g, dtype=torch.float32):
    return oa_ref.make_state_dict(oa_ref.leftnet_param_shapes(g["cfg"]), int(g["seed"]), g["cfg"], dtype=dtype)


def rel_err_floor(a, b, floor):
    """max|a-b| / max(max|b|, floor): for outputs that can legitimately be ~0 (a position update of a graph with almost no
    active edge is 1e-7 of the positions; its relative error against itself is rounding noise over nothing)."""
    a, b = torch.as_tensor(a, dtype=torch.float64).detach(), torch.as_tensor(b, dtype=torch.float64).detach()
    return float((a - b).abs().max() / b.abs().max().clamp(min=floor))


def rel_err(a, b):
    """max|a-b| / max|b| — the parity figure SURVEY §8d asks for."""
    a, b = torch.as_tensor(a, dtype=torch.float64).detach(), torch.as_tensor(b, dtype=torch.float64).detach()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
