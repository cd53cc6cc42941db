"""CPU tests: pin oracle/oa_ref.py against (i) the reference's integer known-answer tests and
(ii) outputs of the unmodified reference frozen in tests/golden (oracle/gen_golden.py)."""
import numpy as np
import pytest
import torch
from torch import tensor

from oracle import oa_ref
from tests.util import dyn_state_dict, leftnet_state_dict, load_golden, rel_err


# ---- integer KATs restated from the reference's tests/utils/test_graph_tools.py:14-63
def test_kat_mask_for_frag():
    assert oa_ref.get_mask_for_frag(tensor([2, 0, 3])).tolist() == [0, 0, 2, 2, 2]


def test_kat_n_frag_switch():
    res = oa_ref.get_n_frag_switch([tensor([2, 0]), tensor([1, 3]), tensor([3, 2])])
    assert res.tolist() == [0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 2]


def test_kat_subgraph_mask():
    ei = tensor([[0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1]])
    assert oa_ref.get_subgraph_mask(ei, tensor([0, 0, 1])).tolist() == [1, 0, 1, 0, 0, 0]


def test_kat_complete_generation():
    # tests/utils/test_graph_tools.py:36-63 and tests/dynamics/test_egnn_dynamics.py:142-153
    frags = [tensor([2, 0]), tensor([2, 3]), tensor([1, 2])]
    masks = [oa_ref.get_mask_for_frag(n) for n in frags]
    nfs = oa_ref.get_n_frag_switch(frags)
    assert nfs.tolist() == [0, 0, 1, 1, 1, 1, 1, 2, 2, 2]
    cm = torch.cat(masks)
    assert cm.tolist() == [0, 0, 0, 0, 1, 1, 1, 0, 1, 1]
    assert oa_ref.get_edges_index(cm).shape == (2, 50)
    ei = oa_ref.get_edges_index(cm, remove_self_edge=True)
    assert ei.shape == (2, 40)
    assert int(oa_ref.get_subgraph_mask(ei, nfs).sum()) == 2 + 2 + 6 + 2


def test_repaint_schedule_identity():
    # _schedule.py:209-213: sum(out) - (len(out)-1)*jump_length == timesteps
    for r, j, T in [(1, 1, 1000), (5, 5, 150), (5, 5, 1000), (2, 3, 12)]:
        s = oa_ref.get_repaint_schedule(r, j, T)
        assert sum(s) - (len(s) - 1) * j == T
    assert sum(oa_ref.get_repaint_schedule(1, 1, 1000)) + 1 == 1001
    assert sum(oa_ref.get_repaint_schedule(5, 5, 150)) + 1 == 731


# ---- golden vectors from the unmodified reference
@pytest.mark.parametrize("name", ["leftnet_small_full", "leftnet_small_cut", "leftnet_small_split", "leftnet_small_noreflect"])
def test_leftnet_forward_matches_reference_fp64(name):
    g = load_golden(name)
    sd = leftnet_state_dict(g, torch.float64)
    h, pos = torch.from_numpy(g["h"]), torch.from_numpy(g["pos"])
    ho, dpos = oa_ref.leftnet_forward(sd, g["cfg"], h, pos, torch.from_numpy(g["edge_index"]),
                                      torch.from_numpy(g["subgraph_mask"]))
    assert rel_err(ho, g["h_out_f64"]) < 1e-11
    assert rel_err(dpos, g["dpos_f64"]) < 1e-9  # reference returns (pos+dpos)-pos


@pytest.mark.parametrize("name", ["dyn_small_ragged", "dyn_trained_cfg1", "dyn_trained_b4", "dyn_trained_b3_far"])
def test_dynamics_matches_reference_fp64(name):
    g = load_golden(name)
    sd = dyn_state_dict(g, torch.float64)
    nf = len(g["node_nfs"])
    xh = [torch.from_numpy(g[f"xh{f}"]) for f in range(nf)]
    dbg = {}
    out = oa_ref.dynamics_forward(sd, g["cfg"], xh, torch.from_numpy(g["edge_index"]), torch.from_numpy(g["t"]),
                                  torch.from_numpy(g["cond"]), torch.from_numpy(g["n_frag_switch"]),
                                  torch.from_numpy(g["combined_mask"]), condition_nf=int(g["condition_nf"]), dbg=dbg)
    # integer artefacts bit-exact
    assert np.array_equal(dbg["mask"].numpy().astype(np.int64), g["mask"])
    assert np.array_equal(dbg["group"].numpy(), g["group"])
    for f in range(nf):
        if g[f"out{f}_f64"].size:
            assert rel_err(out[f], g[f"out{f}_f64"]) < 1e-9, f


def test_graph_construction_bit_exact():
    g = load_golden("dyn_trained_b4")
    frags = [torch.from_numpy(x) for x in g["fragments_nodes"]]
    masks = [oa_ref.get_mask_for_frag(n) for n in frags]
    cm = torch.cat(masks)
    assert np.array_equal(cm.numpy(), g["combined_mask"])
    assert np.array_equal(oa_ref.get_edges_index(cm, remove_self_edge=True).numpy(), g["edge_index"])
    assert np.array_equal(oa_ref.get_n_frag_switch(frags).numpy(), g["n_frag_switch"])


@pytest.mark.parametrize("name", ["sample_small_T10", "sample_trained_cfg1_T10"])
def test_sample_matches_reference_fp32(name):
    g = load_golden(name)
    sizes = [int(x) for x in g["sizes"]]
    shapes = oa_ref.dynamics_param_shapes(g["cfg"], [9, 9, 9], 1)
    sd = oa_ref.make_state_dict(shapes, int(g["seed"]), g["cfg"], prefix_model="model.")
    gamma = oa_ref.gamma_table("polynomial_2", int(g["T"]), 1e-5)
    assert torch.equal(gamma, torch.from_numpy(g["gamma"]))
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, int(g["seed"]))
    smp = oa_ref.Sampler(sd, g["cfg"], gamma)
    torch.manual_seed(int(g["seed"]))
    out, _ = smp.sample(len(sizes), nodes, cond, h0)
    assert smp.n_evals == int(g["T"]) + 1
    for f in range(3):
        # same ops, same RNG stream, same CPU => agreement far below the fp32 trajectory noise
        assert rel_err(out[f][:, :3], g[f"out{f}"][:, :3]) < 2e-4
        assert np.array_equal(out[f][:, 3:].numpy(), g[f"out{f}"][:, 3:])


def test_inpaint_matches_reference_fp32():
    g = load_golden("inpaint_small_T12_r2_j3")
    sizes = [int(x) for x in g["sizes"]]
    shapes = oa_ref.dynamics_param_shapes(g["cfg"], [9, 9, 9], 1)
    sd = oa_ref.make_state_dict(shapes, int(g["seed"]), g["cfg"], prefix_model="model.")
    gamma = oa_ref.gamma_table("polynomial_2", int(g["T"]), 1e-5)
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, int(g["seed"]))
    smp = oa_ref.Sampler(sd, g["cfg"], gamma)
    xh_fixed = [torch.from_numpy(g[f"xh_fixed{f}"]) for f in range(3)]
    torch.manual_seed(int(g["seed"]))
    out, _ = smp.inpaint(len(sizes), nodes, cond, xh_fixed, [0, 2], int(g["resamplings"]), int(g["jump_length"]))
    assert smp.n_evals == sum(oa_ref.get_repaint_schedule(2, 3, 12)) + 1
    for f in range(3):
        assert rel_err(out[f][:, :3], g[f"out{f}"][:, :3]) < 2e-4
