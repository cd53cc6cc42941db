"""Pins the oracle's DIFFERENTIABILITY: gradients of the l2 training objective from torch autograd through the CPU
restatement (oracle/oa_ref.py::train_loss_l2) against golden gradients of the UNMODIFIED reference's autograd
(oracle/gen_golden.py::case_train_grad, fp64).  This is the checker the CUDA backward (BASELINE
config 5) is tested against; the reference's own fp32-vs-fp64 gradient gap is printed next to it."""
import json

import numpy as np
import pytest
import torch

from oracle import oa_ref
from tests.util import load_golden


def _loss_and_grads(g, dtype=torch.float64):
    cfg = g["cfg"]
    sizes = torch.tensor(g["sizes"])
    shapes = oa_ref.dynamics_param_shapes(cfg, [9, 9, 9], 1)
    sd = oa_ref.make_state_dict(shapes, int(g["seed"]), cfg, dtype=dtype, prefix_model="model.")
    buffers = {k for k in sd if k.endswith(("radial_emb.means", "radial_emb.betas"))}
    for k, v in sd.items():
        if k not in buffers:
            v.requires_grad_(True)
    masks = [oa_ref.get_mask_for_frag(sizes) for _ in range(3)]
    xh = [torch.cat([torch.from_numpy(g[f"pos{f}"]), torch.from_numpy(g[f"one_hot{f}"]), torch.from_numpy(g[f"charge{f}"])], dim=1).to(dtype)
          for f in range(3)]
    eps = [torch.from_numpy(g[f"noise0_{f}"]).to(dtype) for f in range(3)]
    gamma = oa_ref.gamma_table("polynomial_2", int(g["T"]), 1e-5).to(dtype)
    loss = oa_ref.train_loss_l2(sd, cfg, gamma, xh, masks, sizes, torch.from_numpy(g["cond"]).to(dtype),
                                torch.from_numpy(g["t_int"]).to(dtype), eps, scales=tuple(float(x) for x in g["scales"]))
    loss.backward()
    return float(loss.detach()), {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sd.items() if k not in buffers}


def test_small_config_full_gradients():
    g = load_golden("grad_small_train")
    assert int(g["n_draws"]) == 1 and bool((g["t_int"] > 0).all())
    loss, grads = _loss_and_grads(g)
    assert abs(loss - float(g["loss"])) < 1e-8 * abs(float(g["loss"]))
    names = json.loads(str(g["param_names"]))
    worst, worst_ref, n_nonzero = 0.0, 0.0, 0
    for pn in names:
        ref = torch.from_numpy(g[f"grad/{pn}"])
        scale = float(ref.abs().max())
        if scale == 0.0:
            assert float(grads[pn].abs().max()) == 0.0, pn  # unused parameters (decoders under pos_only, distance_embedding, ...)
            continue
        n_nonzero += 1
        worst = max(worst, float((grads[pn] - ref).abs().max()) / scale)
        worst_ref = max(worst_ref, float((torch.from_numpy(g[f"grad_f32/{pn}"]).double() - ref).abs().max()) / scale)
    print(f"oracle autograd vs reference autograd (fp64): worst rel err {worst:.2e} over {n_nonzero} parameters; "
          f"reference fp32 vs fp64: {worst_ref:.2e}")
    assert worst < 1e-5 and n_nonzero > 80  # the degenerate legacy node frame amplifies 1e-10 input differences


def test_trained_config_gradient_checksums():
    g = load_golden("grad_trained_train_b3")
    loss, grads = _loss_and_grads(g)
    assert abs(loss - float(g["loss"])) < 1e-8 * abs(float(g["loss"]))
    names = json.loads(str(g["param_names"]))
    gen = torch.Generator().manual_seed(int(g["seed"]) + 99)
    worst = 0.0
    for pn in names:
        direction = torch.randn(grads[pn].shape, generator=gen, dtype=torch.float64)
        gn, gp = float(g[f"gnorm/{pn}"]), float(g[f"gproj/{pn}"])
        if gn == 0.0:
            assert float(grads[pn].norm()) == 0.0, pn
            continue
        worst = max(worst, abs(float(grads[pn].norm()) - gn) / gn, abs(float((grads[pn] * direction).sum()) - gp) / gn)
    print(f"trained config: worst |norm| / projection deviation {worst:.2e}")
    assert worst < 1e-4
