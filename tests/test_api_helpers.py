"""Module-level helpers a caller of the reference may import next to the samplers (schedules `_schedule.py:9-74`,
`diffusion/_utils.py`, `get_inner_edge_index`, `gaussian_KL`, `kl_prior`, `inpaint_fixed`, `build_encoders_decoders`):
same names, same results as the UNMODIFIED reference (fixture: oracle/gen_golden.py::case_api_helpers)."""
import inspect
import os

import numpy as np
import torch

import oareactdiff_b200 as ob
from oareactdiff_b200 import diffusion as D
from oareactdiff_b200 import schedule as S

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "api_helpers.npz"))


def test_alpha2_schedules_bit_exact():
    for T in (10, 100, 1000):
        assert np.array_equal(S.polynomial_schedule(T, s=1e-5, power=2.0), G[f"poly2_{T}"])
        assert np.array_equal(S.polynomial_schedule(T), G[f"poly3_{T}"])
        assert np.array_equal(S.cosine_beta_schedule(T), G[f"cos_{T}"])
        assert np.array_equal(S.cosine_beta_schedule(T, raise_to_power=2.0), G[f"cos2_{T}"])
        assert np.array_equal(S.ccosine_schedule(T, start=0.1, end=0.9, tau=1.5), G[f"ccos_{T}"])
        assert np.array_equal(S.linear_schedule(T), G[f"lin_{T}"])
        for fam in ("polynomial_2", "cosine", "cosine_2", "csin_0.1_0.9_2", "linear"):
            got = S.PredefinedNoiseSchedule(fam, T, 1e-5).gamma.detach().numpy()
            assert got.dtype == np.float32 and np.array_equal(got, G[f"gamma_{fam}_{T}"]), (fam, T)
    assert np.array_equal(S.clip_noise_schedule(G["clip_in"], clip_value=0.2), G["clip_out"])


def test_diffusion_utils_match_reference():
    idx, x = torch.from_numpy(G["u_idx"]), torch.from_numpy(G["u_x"])
    # segment count inferred like torch_scatter (max + 1) and supplied (the samplers' sync-free form): same numbers
    for n_seg in (None, 6):
        np.testing.assert_allclose(D.remove_mean_batch(x, idx, n_seg).numpy(), G["u_remove_mean"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(D.sum_except_batch(x, idx, dim_size=7).numpy(), G["u_sum_except_batch"], rtol=0, atol=1e-6)
    assert np.array_equal(D.cdf_standard_gaussian(x).numpy(), G["u_cdf"])
    assert np.array_equal(D.num_nodes_to_batch_mask(4, torch.tensor([2, 0, 3, 1]), torch.device("cpu")).numpy(), G["u_batch_mask"])
    assert np.array_equal(D.num_nodes_to_batch_mask(3, 2, torch.device("cpu")).numpy(), G["u_batch_mask_int"])
    torch.manual_seed(123)  # same generator draws as the reference: one randn of the full size
    got = D.sample_center_gravity_zero_gaussian_batch([7, 3], [idx[:4], idx[4:]])
    np.testing.assert_allclose(got.numpy(), G["u_cog_noise"], rtol=0, atol=1e-6)
    D.assert_mean_zero_with_mask(got, idx)
    torch.manual_seed(124)
    assert np.array_equal(D.sample_gaussian((5, 2), torch.device("cpu")).numpy(), G["u_gauss"])
    try:
        D.assert_mean_zero_with_mask(x + 1.0, idx)
    except AssertionError as e:
        assert "Mean is not zero" in str(e)
    else:
        raise AssertionError("a shifted cloud must trip the centre-of-mass check")


def test_graph_and_kl_helpers():
    assert np.array_equal(ob.get_inner_edge_index(torch.from_numpy(G["inner_in"])).numpy(), G["inner_out"])
    q = torch.from_numpy(G["kl_in"])
    got = ob.EnVariationalDiffusion.gaussian_KL(q, q + 0.5, 2 * q + 0.1, 3.0)
    np.testing.assert_allclose(got.numpy(), G["kl_out"], rtol=1e-6, atol=1e-7)
    assert ob.EnVariationalDiffusion.kl_prior(None) is NotImplementedError  # the reference returns the class (en_diffusion.py:319-320)


def test_inpaint_fixed_and_builder_signatures():
    a, b = inspect.signature(ob.EnVariationalDiffusion.inpaint), inspect.signature(ob.EnVariationalDiffusion.inpaint_fixed)
    assert list(a.parameters) == list(b.parameters)
    assert [p.default for p in a.parameters.values()] == [p.default for p in b.parameters.values()]
    assert list(inspect.signature(ob.EGNNDynamics.build_encoders_decoders).parameters) == ["self", "enfoce_name_encoding", "source"]
    cfg = dict(cutoff=5.0, num_layers=1, hidden_channels=32, num_radial=16, in_hidden_channels=8, reflect_equiv=True,
               legacy=True, update=True, object_aware=True)
    dyn = ob.EGNNDynamics(model_config=cfg, fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0, condition_nf=1,
                          model=ob.LEFTNetB200, device=torch.device("cpu"), enforce_same_encoding=[1, 2])
    assert dyn.encoders[1] is dyn.encoders[0] and dyn.decoders[2] is dyn.decoders[0]
    assert dyn.edge_encoder is None and dyn.edge_decoder is None
    src = {"encoders": dyn.encoders.state_dict(), "decoders": dyn.decoders.state_dict()}
    dyn.build_encoders_decoders(None, src)  # rebuild + load, as BaseDynamics.__init__ does with `source`
    assert dyn.encoders[1] is not dyn.encoders[0]
    for k, v in src["encoders"].items():
        assert torch.equal(dyn.encoders.state_dict()[k], v)
