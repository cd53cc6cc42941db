"""The evaluation-side callers of the path, end to end, against the UNMODIFIED reference (oracle/gen_golden.py::
case_eval_pipeline): packed `ProcessedTS1x.batch()` -> `set_new_schedule` -> `inplaint_batch` -> `samples_to_pos_charge`
(evaluate/utils.py:14-63, 91-110; dataset/transition1x.py).  CPU: the fp64 oracle stands in for the CUDA engine behind
`LEFTNetB200.forward`; the reference's noise stream is reproduced draw for draw."""
import copy
import json
import os

import numpy as np
import pytest
import torch

import oareactdiff_b200 as ob
from oareactdiff_b200 import data as D
from oracle import oa_ref
from tests.test_reference_suite_cpu import _oracle_forward
from tests.util import rel_err

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "eval_pipeline_small.npz"), allow_pickle=False)


class _Trainer:  # what the helpers need of the reference's LightningModule: `.ddpm` and `.to()`
    def __init__(self, ddpm):
        self.ddpm = ddpm

    def to(self, device):
        self.ddpm.to(device)
        return self


@pytest.mark.parametrize("producer", ["packed", "collate"])
@pytest.mark.parametrize("wrapped", [True, False])
def test_dataset_batch_to_inpainted_positions(monkeypatch, producer, wrapped):
    monkeypatch.setattr(ob.LEFTNetB200, "forward", _oracle_forward)
    monkeypatch.setattr(ob.EGNNDynamics, "fused_ok", lambda self, device: False)
    cfg, seed = json.loads(str(G["cfg"])), int(G["seed"])
    ds = D.ProcessedTS1x(copy.deepcopy(json.loads(str(G["raw"]))), **json.loads(str(G["kw"])))
    idxs = [int(i) for i in G["idxs"]]
    batch = ds.batch(idxs) if producer == "packed" else D.ProcessedTS1x.collate_fn([ds[i] for i in idxs])
    sd = oa_ref.make_state_dict(oa_ref.dynamics_param_shapes(cfg, [9, 9, 9], 1), seed, cfg, prefix_model="model.")
    dyn = ob.EGNNDynamics(model_config=cfg, fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0, condition_nf=1,
                          model=ob.LEFTNetB200, device=torch.device("cpu"))
    dyn.load_state_dict(sd, strict=True)
    ddpm = ob.EnVariationalDiffusion(dynamics=dyn, schdule=ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", int(G["T0"]), 1e-5),
                                                                            norm_values=(1.0, 1.0, 1.0)),
                                     normalizer=ob.Normalizer(), pos_only=True)
    target = _Trainer(ddpm) if wrapped else ddpm
    back = ob.set_new_schedule(target, timesteps=int(G["T_new"]), device=torch.device("cpu"), noise_schedule="polynomial_3")
    assert back is target and ddpm.T == int(G["T_new"])
    assert np.array_equal(ddpm.schedule.gamma_module.gamma.detach().numpy(), G["gamma_new"])
    torch.manual_seed(seed)
    out, xh_fixed, fragments_nodes = ob.inplaint_batch(batch, target, resamplings=2, jump_length=2, frag_fixed=[0, 2])
    assert ddpm.n_evals == sum(ob.get_repaint_schedule(2, 2, int(G["T_new"]))) + 1
    for f in range(3):
        assert np.allclose(xh_fixed[f].numpy(), G[f"xh_fixed{f}"], rtol=0, atol=1e-6)  # centred in place by inpaint, like the reference
        assert rel_err(out[f][:, :3], G[f"out{f}"][:, :3]) < 1e-4
        assert np.array_equal(out[f][:, 3:].numpy(), G[f"out{f}"][:, 3:])
    pos, z, natoms = ob.samples_to_pos_charge(out, fragments_nodes)
    assert natoms == [int(n) for n in G["natoms"]] and set(pos) == {"reactant", "transition_state", "product"}
    for k, v in pos.items():
        assert len(v) == len(natoms)
        for i, a in enumerate(v):
            assert a.shape == G[f"pos/{k}/{i}"].shape and rel_err(a, G[f"pos/{k}/{i}"]) < 1e-4
    for i, a in enumerate(z):
        assert np.array_equal(a, G[f"z/{i}"])
