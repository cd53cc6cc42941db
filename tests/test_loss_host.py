"""Host-side check (no GPU) of the loss arithmetic in `EnVariationalDiffusion.forward` / `compute_loss`: with the
denoiser replaced by a stub that returns the reference's golden network output, every loss term must match the
golden terms of the unmodified reference to fp32 round-off (en_diffusion.py:56-248, 340-454)."""
import numpy as np
import pytest
import torch
from torch import nn

import oareactdiff_b200 as ob
from tests.util import load_golden


class _StubDynamics(nn.Module):
    pos_dim, node_nfs, fragment_names = 3, [9, 9, 9], ["R", "TS", "P"]

    def __init__(self, outs):
        super().__init__()
        self.outs, self.k = outs, 0

    def forward(self, xh, edge_index, t, conditions, n_frag_switch, combined_mask, edge_attr=None):
        out = [o.clone() for o in self.outs[self.k]]
        self.k += 1
        return out, None


class _Replay(ob.EnVariationalDiffusion):
    def _draw_t_int(self, num_sample, device):
        return self._t_int.view(num_sample, 1)

    def sample_combined_position_feature_noise(self, masks):
        self._k += 1
        return [n.clone() for n in self._noises[self._k - 1]]


@pytest.mark.parametrize("name", ["loss_small_train", "loss_trained_train_b4"])
def test_loss_terms_arithmetic_vs_reference_golden(name):
    g = load_golden(name)
    sizes = torch.tensor(g["sizes"])
    dyn = _StubDynamics([[torch.from_numpy(g[f"net_eps_xh{f}_f32"]) for f in range(3)]])
    sched = ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", int(g["T"]), 1e-5), norm_values=(1.0, 1.0, 1.0))
    ddpm = _Replay(dynamics=dyn, schdule=sched, normalizer=ob.Normalizer(), pos_only=True)
    ddpm.train(True)
    ddpm._t_int, ddpm._k = torch.from_numpy(g["t_int"]).float(), 0
    ddpm._noises = [[torch.from_numpy(g[f"noise{d}_{f}"]) for f in range(3)] for d in range(int(g["n_draws"]))]
    reps = [{"size": sizes.clone(), "pos": torch.from_numpy(g[f"pos{f}"]), "one_hot": torch.from_numpy(g[f"one_hot{f}"]),
             "charge": torch.from_numpy(g[f"charge{f}"]), "mask": ob.get_mask_for_frag(sizes)} for f in range(3)]
    lt = ddpm.forward(reps, torch.from_numpy(g["cond"]))
    for k in ("error_t", "loss_0_x", "loss_0_cat", "loss_0_charge", "eps_xh"):
        for f in range(3):
            assert np.allclose(lt[k][f].numpy(), g[f"{k}{f}_f32"], rtol=2e-5, atol=1e-6), (k, f)
    for k in ("SNR_weight", "neg_log_constants", "kl_prior", "t_int"):
        assert np.allclose(lt[k].numpy(), g[k + "_f32"], rtol=1e-6, atol=1e-7), k
    assert abs(float(lt["delta_log_px"]) - float(g["delta_log_px_f32"])) < 1e-9
