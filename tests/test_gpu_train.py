"""GPU check of the TRAINING path (BASELINE config 5): `compute_loss(...).mean().backward()` through
`oard_forward_train` / `oard_backward` against golden gradients of the UNMODIFIED reference's autograd
(oracle/gen_golden.py::case_train_grad, fp64; inputs and random draws replayed).

The kernels behind these entry points are the same source as the host-emulation build that tests/test_train_emu.py
validates on the CPU (forward 2e-7, every parameter gradient < 1e-5 of the fp64 oracle).  Measured on the B200: worst
parameter-gradient error 9.3e-6 / 1.1e-6 of max|grad| on the two small fixtures.  Tolerance: 2e-4 of max|grad| per parameter
(the reference's own fp32 autograd is up to 7e-2 away from its fp64 run on these fixtures)."""
import json
import os

import numpy as np
import pytest
import torch

import oareactdiff_b200 as ob
from tests.test_gpu_loss import _ReplayDraws
from tests.test_gpu_parity import DEV, make_dynamics
from tests.util import dyn_state_dict, load_golden

pytestmark = pytest.mark.gpu


def _setup(name):
    g = load_golden(name)
    g["node_nfs"], g["condition_nf"] = np.array([9, 9, 9]), np.int64(1)
    dyn = make_dynamics(g["cfg"], dyn_state_dict(g))
    dyn.model.enable_training_path = True
    sched = ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", int(g["T"]), 1e-5), norm_values=(1.0, 1.0, 1.0))
    ddpm = _ReplayDraws(dynamics=dyn, schdule=sched, normalizer=ob.Normalizer(), pos_only=True).to(DEV)
    ddpm.train(True)
    sizes = torch.tensor(g["sizes"])
    reps = [{"size": sizes.clone().to(DEV), "pos": torch.from_numpy(g[f"pos{f}"]).to(DEV),
             "one_hot": torch.from_numpy(g[f"one_hot{f}"]).to(DEV), "charge": torch.from_numpy(g[f"charge{f}"]).to(DEV),
             "mask": ob.get_mask_for_frag(sizes).to(DEV)} for f in range(3)]
    noises = [[torch.from_numpy(g[f"noise{d}_{f}"]) for f in range(3)] for d in range(int(g["n_draws"]))]
    ddpm.set_draws(torch.from_numpy(g["t_int"]).float(), noises)
    return g, ddpm, reps, torch.from_numpy(g["cond"]).to(DEV)


def _loss_backward(name):
    g, ddpm, reps, cond = _setup(name)
    nll, _ = ddpm.compute_loss((reps, cond), scales=tuple(float(x) for x in g["scales"]), training=True)
    loss = nll.mean()
    loss.backward()
    return g, ddpm, float(loss)


@pytest.mark.parametrize("name", ["grad_small_train", "grad_small_train_t0"])
def test_training_step_gradients_small_config(name):
    g, ddpm, loss = _loss_backward(name)
    assert abs(loss - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    worst, n = 0.0, 0
    for pn, prm in ddpm.dynamics.named_parameters():
        ref = torch.from_numpy(g[f"grad/{pn}"])
        scale = float(ref.abs().max())
        if scale == 0.0:
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0, pn
            continue
        assert prm.grad is not None, pn
        n += 1
        worst = max(worst, float((prm.grad.cpu().double() - ref).abs().max()) / scale)
    print(f"{name}: loss {loss:.6f}, worst parameter-gradient error {worst:.2e} over {n} parameters")
    assert worst < 2e-4 and n > 80


def test_training_step_gradient_checksums_trained_config():
    g, ddpm, loss = _loss_backward("grad_trained_train_b3")
    assert abs(loss - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    gen = torch.Generator().manual_seed(int(g["seed"]) + 99)
    worst = 0.0
    for pn in json.loads(str(g["param_names"])):
        prm = dict(ddpm.dynamics.named_parameters())[pn]
        direction = torch.randn(prm.shape, generator=gen, dtype=torch.float64)
        gn, gp = float(g[f"gnorm/{pn}"]), float(g[f"gproj/{pn}"])
        if gn == 0.0:
            continue
        gr = prm.grad.cpu().double()
        worst = max(worst, abs(float(gr.norm()) - gn) / gn, abs(float((gr * direction).sum()) - gp) / gn)
    print(f"grad_trained_train_b3: worst |norm| / projection deviation {worst:.2e}")
    assert worst < 2e-4  # measured 3.0e-6 on the B200


def test_one_sgd_step_reduces_the_loss():
    g, ddpm, loss0 = _loss_backward("grad_small_train")
    with torch.no_grad():
        for prm in ddpm.dynamics.parameters():
            if prm.grad is not None:
                prm -= 0.05 * prm.grad
    ddpm.set_draws(torch.from_numpy(g["t_int"]).float(), ddpm._noises)
    _, _, reps, cond = _setup("grad_small_train")
    nll, _ = ddpm.compute_loss((reps, cond), scales=tuple(float(x) for x in g["scales"]), training=True)
    assert float(nll.mean()) < loss0
