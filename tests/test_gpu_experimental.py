"""GPU checks of the less-travelled paths: `reflect_equiv=False`, sparse / shuffled / split graphs over the option grid on which
the oracle is pinned to the unmodified reference, the device-resident path on ragged fragments with an empty one, and
parity on real Transition1x geometries."""
import os

import numpy as np
import pytest
import torch

from oracle import oa_ref
from tests.test_gpu_parity import DEV, _oracle_inputs, make_dynamics

pytestmark = pytest.mark.gpu


def test_leftnet_reflect_equiv_false_vs_reference_golden():
    """`reflect_equiv=False` (the reference's own tests build such a LEFTNet: tests/model/test_equiv.py:40-41): no abs on the
    frame's cross row and the `x * edge_cross` term in the messages (leftnet.py:268-269) — the node-per-block message kernel
    (k_equi_reduce) instead of the group-staged one.  Golden from the unmodified reference; the oracle is pinned on it in
    tests/test_oracle.py."""
    from tests.test_gpu_parity import REL_TOL, make_leftnet
    from tests.util import leftnet_state_dict, load_golden, rel_err
    g = load_golden("leftnet_small_noreflect")
    assert g["cfg"]["reflect_equiv"] is False
    m = make_leftnet(g["cfg"], leftnet_state_dict(g))
    h, pos = torch.from_numpy(g["h"]).float().to(DEV), torch.from_numpy(g["pos"]).float().to(DEV)
    ho, po, _ = m(h, pos, torch.from_numpy(g["edge_index"]).to(DEV), subgraph_mask=torch.from_numpy(g["subgraph_mask"]).to(DEV))
    e_h, e_p = rel_err(ho.cpu(), g["h_out_f64"]), rel_err((po - pos).cpu(), g["dpos_f64"])
    print(f"noreflect: h {e_h:.2e} dpos {e_p:.2e}")
    assert e_h < REL_TOL and e_p < REL_TOL


def _grid_graphs():
    def full(n):
        return torch.tensor([[i, j] for i in range(n) for j in range(n) if i != j]).T.contiguous()
    g = torch.Generator().manual_seed(0)
    e8 = full(8)
    return {"complete9": (9, full(9)), "path": (4, torch.tensor([[0, 1, 1, 2, 3, 0], [1, 0, 2, 1, 0, 3]])),
            "two_components": (7, torch.tensor([[0, 1, 1, 2, 0, 3, 4, 6, 4, 5], [1, 0, 2, 1, 3, 0, 6, 4, 5, 4]])),
            "complete8_shuffled": (8, e8[:, torch.randperm(e8.size(1), generator=g)]),
            "two_cliques": (11, torch.cat([full(5), full(6) + 5], dim=1))}


@pytest.mark.parametrize("gname", ["complete9", "path", "two_components", "complete8_shuffled", "two_cliques"])
def test_leftnet_option_grid_vs_oracle(gname):
    """The grid on which oracle/fuzz_oracle_leftnet.py pins the oracle to the unmodified reference (graph shapes, masks,
    cut-offs that split groups, reflect_equiv, object_aware, update, depth), now CUDA vs that oracle (fp64), REL_TOL = 2e-4.
    Graphs with almost no active edge have a position update of ~1e-7 of the positions: dpos is measured against
    max(max|dpos_ref|, 1e-3 max|pos|), like the reference's own tests, which compare pos + dpos."""
    import itertools
    from tests.test_gpu_parity import REL_TOL, make_leftnet
    from tests.util import rel_err, rel_err_floor
    n, ei = _grid_graphs()[gname]
    g = torch.Generator().manual_seed(1)
    worst = 0.0
    for reflect, oa, update, layers, cut, cutoff, scale in itertools.product([True, False], [True, False], [True, False], [1, 3],
                                                                          [None, 3], [20.0, 2.5], [1.0, 3.0]):
        if n < 5 and cut:
            continue
        cfg = dict(cutoff=cutoff, num_layers=layers, hidden_channels=32, num_radial=16, in_hidden_channels=6, reflect_equiv=reflect,
                   legacy=True, update=update, object_aware=oa)
        sd = oa_ref.make_state_dict(oa_ref.leftnet_param_shapes(cfg), 3, cfg, dtype=torch.float64)
        h = torch.rand(n, 6, generator=g, dtype=torch.float64)
        pos = torch.rand(n, 3, generator=g, dtype=torch.float64) * scale
        sub = None
        if cut:
            s = (ei < cut).sum(0)
            sub = ((s == 2) | (s == 0)).long()[:, None]
        ho_ref, dpos_ref = oa_ref.leftnet_forward(sd, cfg, h, pos, ei, sub)
        m = make_leftnet(cfg, {k: v.float() for k, v in sd.items()})
        ho, po, _ = m(h.float().to(DEV), pos.float().to(DEV), ei.to(DEV), subgraph_mask=None if sub is None else sub.to(DEV))
        e = max(rel_err(ho.cpu(), ho_ref), rel_err_floor(po.cpu() - pos.float(), dpos_ref, 1e-3 * float(pos.abs().max())))
        worst = max(worst, e)
        assert e < REL_TOL, (gname, reflect, oa, update, layers, cut, cutoff, scale, e)
    print(f"{gname}: worst rel err over the grid {worst:.2e}")


def test_device_resident_path_on_ragged_fragments_with_an_empty_one():
    """The device-resident dynamics wrapper and reverse step (oard_dyn_forward / oard_reverse_step) on the reference's ragged
    fixture shape — fragments of different sizes, one of them EMPTY in a sample (tests/dynamics/test_egnn_dynamics.py:100-104)
    — with equal node_nf so the fused path is taken: against the torch-composed wrapper around the same kernels, and a short
    sample() against the host fast path."""
    from tests.test_gpu_parity import make_dynamics
    from tests.util import rel_err
    import oareactdiff_b200 as ob
    cfg = dict(oa_ref.TRAINED_CFG, hidden_channels=32, num_radial=16, num_layers=2, cutoff=5.0)
    sd = oa_ref.make_state_dict(oa_ref.dynamics_param_shapes(cfg, [9, 9, 9], 1), 4, cfg, prefix_model="model.")
    dyn = make_dynamics(cfg, sd)
    assert dyn.fused_ok(DEV)
    g = torch.Generator().manual_seed(0)
    nodes = [torch.tensor([2, 0]), torch.tensor([2, 3]), torch.tensor([1, 2])]
    masks = [ob.get_mask_for_frag(n) for n in nodes]
    cm = torch.cat(masks)
    ei, nfs = ob.get_edges_index(cm, remove_self_edge=True), ob.get_n_frag_switch(nodes)
    h0 = [torch.cat([torch.nn.functional.one_hot(torch.randint(0, 5, (int(n.sum()),), generator=g), 5),
                     torch.randint(1, 9, (int(n.sum()), 1), generator=g)], dim=1) for n in nodes]
    xh = [torch.cat([oa_ref.remove_mean_batch(torch.randn(h.size(0), 3, generator=g), m), h.float()], dim=1) for h, m in zip(h0, masks)]
    t, cond = torch.rand(2, 1, generator=g), torch.rand(2, 1, generator=g)
    args = ([x.to(DEV) for x in xh], ei.to(DEV), t.to(DEV), cond.to(DEV), nfs.to(DEV), cm.to(DEV))
    for _ in range(3):
        fused, _ = dyn(*args)
    dyn.use_fused = False
    host, _ = dyn(*args)
    dyn.use_fused = True
    ref = oa_ref.dynamics_forward({k: v.double() for k, v in sd.items()}, cfg, [x.double() for x in xh], ei, t.double(), cond.double(), nfs, cm)
    for f in range(3):
        assert fused[f].shape == xh[f].shape
        assert rel_err(fused[f].cpu(), host[f].cpu()) < 1e-4 and rel_err(fused[f].cpu(), ref[f]) < 1e-3
    sched = ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", 6, 1e-5), norm_values=(1.0, 1.0, 1.0))
    ddpm = ob.EnVariationalDiffusion(dynamics=dyn, schdule=sched, normalizer=ob.Normalizer(), pos_only=True).to(DEV)
    outs = []
    for fused_on in (True, False):
        dyn.use_fused = fused_on
        torch.manual_seed(3)
        out, _ = ddpm.sample(2, [n.to(DEV) for n in nodes], cond.to(DEV), h0=[h.to(DEV) for h in h0])
        outs.append(torch.cat([o[:, :3].cpu() for o in out[0]]))
    assert rel_err(outs[0], outs[1]) < 1e-3


def test_dynamics_on_real_transition1x_geometries_vs_oracle():
    """Trained configuration on REAL Transition1x geometries (fixture tests/golden/t1x_geometries_b512.npz: reactant / TS /
    product of the first reactions the bench draws), noised to a mid-trajectory state: CUDA vs the fp64 oracle."""
    import numpy as np
    from tests.test_gpu_parity import REL_TOL, make_dynamics
    from tests.util import rel_err
    import oareactdiff_b200 as ob
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "t1x_geometries_b512.npz"))
    k = 5
    sizes = [int(v) for v in fx["sizes"][:k]]
    n_at = sum(sizes)
    cfg = dict(oa_ref.TRAINED_CFG)
    sd = oa_ref.make_state_dict(oa_ref.dynamics_param_shapes(cfg, [9, 9, 9], 1), 8, cfg, prefix_model="model.")
    nodes = [torch.tensor(sizes)] * 3
    masks = [ob.get_mask_for_frag(n) for n in nodes]
    cm = torch.cat(masks)
    ei, nfs = ob.get_edges_index(cm, remove_self_edge=True), ob.get_n_frag_switch(nodes)
    z = torch.from_numpy(fx["Z"][:n_at])
    lut = {1: 0, 6: 1, 7: 2, 8: 3, 9: 4}
    h = torch.cat([torch.nn.functional.one_hot(torch.tensor([lut[int(v)] for v in z]), 5).float(), z.float()[:, None]], dim=1)
    g = torch.Generator().manual_seed(2)
    xh = []
    for key in ("reactant", "transition_state", "product"):
        x = torch.from_numpy(fx[key][:n_at])
        x = 0.8 * x + 0.6 * oa_ref.remove_mean_batch(torch.randn(n_at, 3, generator=g), masks[0])  # q(z_t | x) at alpha = 0.8
        xh.append(torch.cat([x, h], dim=1))
    t, cond = torch.full((k, 1), 0.35), torch.zeros(k, 1)
    ref = oa_ref.dynamics_forward({kk: v.double() for kk, v in sd.items()}, cfg, [x.double() for x in xh], ei, t.double(), cond.double(), nfs, cm)
    dyn = make_dynamics(cfg, sd)
    out, _ = dyn([x.to(DEV) for x in xh], ei.to(DEV), t.to(DEV), cond.to(DEV), nfs.to(DEV), cm.to(DEV))
    for f in range(3):
        e = rel_err(out[f].cpu(), ref[f])
        print(f"real geometries frag{f}: {e:.2e}")
        assert e < REL_TOL
