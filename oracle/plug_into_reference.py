"""ORACLE tooling (test infrastructure, build container only): the plugin seam, exercised with the UNMODIFIED reference.

`LEFTNetB200` is handed to the reference's own `EGNNDynamics(model=<class>)` (dynamics/_base.py:21,62-64), the reference's
state dict is strict-loaded, and the reference's own `EnVariationalDiffusion.sample()` drives it for a whole trajectory,
compared with the golden trajectory of the reference's own LEFTNet.  There is no GPU in the build container, so the fp64
oracle stands in for the CUDA engine behind `LEFTNetB200.forward` (as in tests/test_reference_suite_cpu.py): what this proves
is the SEAM — constructor kwargs incl. the injected `act_fn` / `in_node_nf` (_base.py:47-50), parameter names, forward
signature and return convention — not the kernels.

    python oracle/plug_into_reference.py        # prints one JSON line
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)

from oa_reactdiff.diffusion._normalizer import Normalizer  # noqa: E402
from oa_reactdiff.diffusion._schedule import DiffSchedule, PredefinedNoiseSchedule  # noqa: E402
from oa_reactdiff.diffusion.en_diffusion import EnVariationalDiffusion  # noqa: E402
from oa_reactdiff.dynamics import EGNNDynamics  # noqa: E402

import oareactdiff_b200 as ob  # noqa: E402
from oracle import oa_ref  # noqa: E402
from tests.test_reference_suite_cpu import _oracle_forward  # noqa: E402
from tests.util import load_golden, rel_err  # noqa: E402


def main():
    ob.LEFTNetB200.forward = _oracle_forward
    g = load_golden("sample_small_T10")
    cfg, seed, T = g["cfg"], int(g["seed"]), int(g["T"])
    sizes = [int(x) for x in g["sizes"]]
    dyn = EGNNDynamics(model_config=dict(cfg), fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0, condition_nf=1,
                       pos_dim=3, update_pocket_coords=True, condition_time=True, edge_cutoff=None, model=ob.LEFTNetB200,
                       device=torch.device("cpu"))
    sd = oa_ref.make_state_dict(oa_ref.dynamics_param_shapes(cfg, [9, 9, 9], 1), seed, cfg, prefix_model="model.")
    res = dyn.load_state_dict(sd, strict=True)
    sched = DiffSchedule(PredefinedNoiseSchedule("polynomial_2", T, 1e-5), norm_values=(1.0, 1.0, 1.0))
    ddpm = EnVariationalDiffusion(dynamics=dyn, schdule=sched, normalizer=Normalizer(), pos_only=True)
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, seed)
    torch.manual_seed(seed)
    out, _ = ddpm.sample(len(sizes), nodes, cond, h0=h0)
    errs = [rel_err(out[0][f][:, :3], g[f"out{f}"][:, :3]) for f in range(3)]
    h_equal = all(bool((out[0][f][:, 3:].numpy() == g[f"out{f}"][:, 3:]).all()) for f in range(3))
    print(json.dumps({"model_class": type(dyn.model).__name__, "missing": list(res.missing_keys), "unexpected": list(res.unexpected_keys),
                      "trajectory_rel_err": errs, "h_equal": h_equal}))


if __name__ == "__main__":
    main()
