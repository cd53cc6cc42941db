"""The reference's OWN LEFTNet / EGNNDynamics tests, restated for the CUDA path (GPU, through the C ABI):
oa_reactdiff/tests/model/test_equiv.py, tests/model/test_subgraphs.py, tests/dynamics/test_switch_fragments.py and
tests/dynamics/test_egnn_dynamics.py — same fixtures (configs, graphs incl. the hand-written, not source-grouped edge
lists, `init_weights`), same properties.  The reference runs them in float64 with 1e-8 ... 1e-6 relative tolerances; the
kernels compute in fp32, so "equal" is max|a - b| <= TOL * max(1, max|b|) with TOL = 2e-3 (the path's stated fp32
tolerance) and "different" keeps the reference's thresholds.  Inputs are float64 like the reference's (the module returns
the caller's dtype).

It drives sparse / non-complete graphs (node-per-block message kernel), edge lists in arbitrary order,
`reflect_equiv=False` and float64 callers."""
import os

import numpy as np
import pytest
import torch
from torch import nn

import oareactdiff_b200 as ob

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
TOL = 2e-3
F64 = torch.float64

# tests/model/utils.py:24-32
LEFT_CONFIG = dict(pos_require_grad=False, cutoff=20.0, num_layers=6, hidden_channels=32, num_radial=32, in_node_nf=8,
                   reflect_equiv=True)
# tests/dynamics/test_egnn_dynamics.py:50-57, test_switch_fragments.py:34-41
LEFTNET_DYN_CONFIG = dict(pos_require_grad=False, cutoff=5.0, num_layers=2, hidden_channels=32, num_radial=8, in_node_nf=8)


def init_weights(m, gain=1.0):  # tests/model/utils.py:39-49 (gain 0.5 in tests/dynamics/test_egnn_dynamics.py:20-31)
    if isinstance(m, nn.Linear):
        nn.init.xavier_uniform_(m.weight, gain=gain)
        if m.bias is not None:
            nn.init.uniform_(m.bias, -gain, gain)


def generate_full_eij(n):  # tests/model/utils.py:52-59
    return torch.tensor([[i, j] for i in range(n) for j in range(n) if i != j]).T.long().contiguous()


def get_cut_graph_mask(edge_index, n_cut):  # tests/model/utils.py:62-69
    s = (edge_index < n_cut).long().sum(dim=0)
    return ((s == 2) | (s == 0)).long()[:, None]


def com(x):
    return x - x.mean(dim=0)


def rotation(theta, alpha):
    rx = torch.tensor([[1, 0, 0], [0, np.cos(theta), -np.sin(theta)], [0, np.sin(theta), np.cos(theta)]], dtype=F64)
    ry = torch.tensor([[np.cos(alpha), 0, np.sin(alpha)], [0, 1, 0], [-np.sin(alpha), 0, np.cos(alpha)]], dtype=F64)
    return ry @ rx


def leftnet(seed, **over):
    torch.manual_seed(seed)
    m = ob.LEFTNetB200(**dict(LEFT_CONFIG, **over))
    m.apply(init_weights)
    return m.to(DEV)


def run(model, h, pos, ei, edge_attr=None, sub=None):
    ho, po, ea = model.forward(h.to(DEV), pos.to(DEV), ei.to(DEV), None if edge_attr is None else edge_attr.to(DEV),
                               subgraph_mask=None if sub is None else sub.to(DEV))
    assert ea is None and ho.dtype == h.dtype and po.dtype == pos.dtype
    return ho.cpu(), po.cpu()


def same(a, b, tol=TOL):
    return float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


def rel_diff(a, b):  # tests/model/utils.py:35-36 (used by the reference for its "must differ" assertions)
    return float(torch.max(torch.abs(a - b) / (a + b + 1e-6) * 2))


# ------------------------------------------------------------------------------------------- tests/model/test_equiv.py
PATH_EI = torch.tensor([[0, 1, 1, 2, 3, 0], [1, 0, 2, 1, 0, 3]])  # test_equiv.py:30-32 (not grouped by source)


def _equiv_inputs(seed=42, n=4):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 8, generator=g, dtype=F64), torch.rand(n, 3, generator=g, dtype=F64)


def test_equiv_rotation_and_edge_attr_ignored():
    """test_equiv.py:80-99 (rotation), :118-151 (edge_attr zeros / None give the same result)."""
    model, (h, pos), rot = leftnet(1), _equiv_inputs(), rotation(0.4, 0.9)
    ea = torch.rand(PATH_EI.size(1), 5, dtype=F64)
    ho, po = run(model, h, pos, PATH_EI, ea)
    ho_r, po_r = run(model, h, pos @ rot, PATH_EI, ea)
    assert same(ho_r, ho) and same(po_r, po @ rot)
    ho_n, po_n = run(model, h, pos, PATH_EI, None)
    assert torch.equal(ho_n, ho) and torch.equal(po_n, po)
    assert float((po - pos).abs().max()) > 1e-4  # the model does move the atoms


def test_equiv_no_reflection_equivariance():
    """test_equiv.py:170-185: with reflect_equiv=False a mirrored input gives a different position update."""
    model, (h, pos) = leftnet(2, reflect_equiv=False), _equiv_inputs()
    pos_reflect = torch.cat([pos[:, :2], -pos[:, 2:]], dim=1)
    _, po = run(model, h, pos, PATH_EI)
    _, po_m = run(model, h, pos_reflect, PATH_EI)
    assert rel_diff(po, po_m) > 1e-5


def test_equiv_disconnected_components_are_independent():
    """test_equiv.py:188-229: a 4-node and a 3-node component; the first evaluated alone gives the same outputs."""
    model = leftnet(3)
    ei = torch.tensor([[0, 1, 1, 2, 0, 3, 4, 6, 4, 5], [1, 0, 2, 1, 3, 0, 6, 4, 5, 4]])
    h, pos = _equiv_inputs(seed=7, n=7)
    ho, po = run(model, h, pos, ei)
    ho_c, po_c = run(model, h[:4], pos[:4], ei[:, :6])
    assert same(ho[:4], ho_c, 1e-5) and same(po[:4], po_c, 1e-5)


# --------------------------------------------------------------------------------------- tests/model/test_subgraphs.py
class _TwoFragments:
    """test_subgraphs.py:35-50: complete graph over n1 + n2 = 4 + 9 atoms, cross-fragment edges masked."""

    def __init__(self, seed=1234):
        g = torch.Generator().manual_seed(seed)
        self.n1, n2 = 4, 9
        self.ei = generate_full_eij(self.n1 + n2)
        self.h = torch.rand(self.n1 + n2, 8, generator=g, dtype=F64)
        self.pos = torch.cat([com(torch.rand(self.n1, 3, generator=g, dtype=F64)), com(torch.rand(n2, 3, generator=g, dtype=F64))])
        self.sub = get_cut_graph_mask(self.ei, self.n1)
        self.trans = torch.rand(3, generator=g, dtype=F64) * 100
        self.gen = g


def test_subgraphs_object_wise_rotation_and_translation():
    """test_subgraphs.py:88-180: rotating / translating ONE fragment leaves h unchanged, co-rotates that fragment's position
    update and leaves the other fragment's alone (the object-aware property)."""
    fx, model, rot = _TwoFragments(), leftnet(4), rotation(0.9, 0.4)
    n1 = fx.n1
    ho, po = run(model, fx.h, fx.pos, fx.ei, sub=fx.sub)
    pos_rot = torch.cat([com(fx.pos[:n1] @ rot), fx.pos[n1:]])
    ho_r, po_r = run(model, fx.h, pos_rot, fx.ei, sub=fx.sub)
    assert same(ho_r, ho) and same(po_r, torch.cat([po[:n1] @ rot, po[n1:]]))
    pos_tr = torch.cat([com(fx.pos[:n1] + fx.trans), fx.pos[n1:]])
    ho_t, po_t = run(model, fx.h, pos_tr, fx.ei, sub=fx.sub)
    assert same(ho_t, ho) and same(po_t, po)


def test_subgraphs_mask_is_not_a_broken_graph():
    """test_subgraphs.py:182-222: masking the cross-fragment edges differs from deleting them (h still mixes across)."""
    fx, model = _TwoFragments(), leftnet(5)
    ho, po = run(model, fx.h, fx.pos, fx.ei, sub=fx.sub)
    ei_cut = fx.ei[:, fx.sub[:, 0] == 1]
    ho_c, po_c = run(model, fx.h, fx.pos, ei_cut)
    assert rel_diff(ho, ho_c) > 1e-4 and rel_diff(po, po_c) > 1e-4


def test_subgraphs_reflection_and_position_update_are_seen_by_the_other_fragment():
    """test_subgraphs.py:224-292."""
    fx = _TwoFragments()
    n1 = fx.n1
    model = leftnet(6, reflect_equiv=False)
    pos_m = fx.pos.clone()
    pos_m[:n1, 2] = -pos_m[:n1, 2]
    _, po = run(model, fx.h, fx.pos, fx.ei, sub=fx.sub)
    _, po_m = run(model, fx.h, pos_m, fx.ei, sub=fx.sub)
    assert rel_diff(po[n1:], po_m[n1:]) > 1e-7
    model = leftnet(7)
    pos_new = fx.pos.clone()
    pos_new[:n1] = com(torch.rand(n1, 3, generator=fx.gen, dtype=F64) * 30)
    ho, po = run(model, fx.h, fx.pos, fx.ei, sub=fx.sub)
    ho_n, po_n = run(model, fx.h, pos_new, fx.ei, sub=fx.sub)
    assert rel_diff(ho[n1:], ho_n[n1:]) > 1e-4 and rel_diff(po[n1:], po_n[n1:]) > 1e-4


def test_subgraphs_separate_graphs_without_edges():
    """test_subgraphs.py:294-339: with no edge between two parts, moving one does not change the other."""
    model = leftnet(8)
    g = torch.Generator().manual_seed(99)
    n1 = 3
    ei = torch.tensor([[0, 1, 1, 2, 0, 2, 3, 4], [1, 0, 2, 1, 2, 0, 4, 3]])
    h = torch.rand(5, 8, generator=g, dtype=F64)
    pos = torch.cat([com(torch.rand(n1, 3, generator=g, dtype=F64)), com(torch.rand(2, 3, generator=g, dtype=F64))])
    ho, po = run(model, h, pos, ei)
    pos_tr = torch.cat([com(pos[:n1] + torch.rand(3, generator=g, dtype=F64)), pos[n1:]])
    ho_t, po_t = run(model, h, pos_tr, ei)
    assert same(ho_t, ho, 1e-5) and same(po_t, po, 1e-5)
    pos2 = pos.clone()
    pos2[:n1] = com(torch.rand(n1, 3, generator=g, dtype=F64))
    ho_n, po_n = run(model, h, pos2, ei)
    assert same(ho_n[n1:], ho[n1:], 1e-5) and same(po_n[n1:], po[n1:], 1e-5)


# -------------------------------------------------------------------- tests/dynamics/test_switch_fragments.py, test_egnn_dynamics.py
def _dynamics(node_nfs, names, edge_nf, seed, gain=0.5):
    torch.manual_seed(seed)
    dyn = ob.EGNNDynamics(model_config=dict(LEFTNET_DYN_CONFIG), node_nfs=node_nfs, edge_nf=edge_nf, condition_nf=3,
                          fragment_names=names, pos_dim=3, update_pocket_coords=True, condition_time=True, edge_cutoff=None,
                          model=ob.LEFTNetB200, device=DEV)
    dyn.apply(lambda m: init_weights(m, gain))
    return dyn.to(DEV)


def _graph(fragments_nodes):
    masks = [ob.get_mask_for_frag(n) for n in fragments_nodes]
    cm = torch.cat(masks)
    return ob.get_n_frag_switch(fragments_nodes), cm, ob.get_edges_index(cm, remove_self_edge=True)


def _dyn_call(dyn, xh, ei, t, cond, nfs, cm, edge_attr=None):
    out, ea = dyn.forward([x.to(DEV) for x in xh], ei.to(DEV), t.to(DEV), cond.to(DEV), nfs.to(DEV), cm.to(DEV),
                          edge_attr=None if edge_attr is None else edge_attr.to(DEV))
    assert ea is None
    return [o.cpu() for o in out]


@pytest.mark.parametrize("same_encoding", [False, True])
def test_switch_fragments(same_encoding):
    """test_switch_fragments.py:113-205: swapping the two fragments changes the prediction unless they share encoder and
    decoder, in which case the outputs are swapped copies (the reference assigns the shared modules AFTER construction)."""
    dyn = _dynamics([5, 5], ["A", "B"], 4, seed=11)
    if same_encoding:
        dyn.encoders[1] = dyn.encoders[0]
        dyn.decoders[1] = dyn.decoders[0]
    g = torch.Generator().manual_seed(0)
    frags = [torch.tensor([4]), torch.tensor([5])]
    nfs, cm, ei = _graph(frags)
    cond, t = torch.rand(1, 3, generator=g), torch.tensor([0.314])
    xh = [torch.rand(int(frags[i].sum()), 5, generator=g) for i in range(2)]
    out = _dyn_call(dyn, xh, ei, t, cond, nfs, cm)
    nfs2, cm2, ei2 = _graph([frags[1], frags[0]])
    out_sw = _dyn_call(dyn, [xh[1], xh[0]], ei2, t, cond, nfs2, cm2)
    if same_encoding:
        assert same(out_sw[1], out[0], 1e-4) and same(out_sw[0], out[1], 1e-4)
    else:
        assert not torch.allclose(out[0], out_sw[1], rtol=1e-6)


def test_egnn_dynamics_forward_shapes_and_conditioning():
    """test_egnn_dynamics.py:142-232: ragged fragments with an EMPTY one, per-fragment node_nf, a 1-element (even integer)
    t, an edge_attr tensor that LEFTNet ignores; the outputs feed back in; time and condition change the prediction."""
    node_nfs = [4, 5, 6]
    dyn = _dynamics(node_nfs, ["inorg_node", "org_edge", "org_node"], 3, seed=12)
    g = torch.Generator().manual_seed(0)
    frags = [torch.tensor([2, 0]), torch.tensor([2, 3]), torch.tensor([1, 2])]
    nfs, cm, ei = _graph(frags)
    assert nfs.tolist() == [0, 0, 1, 1, 1, 1, 1, 2, 2, 2] and cm.tolist() == [0, 0, 0, 0, 1, 1, 1, 0, 1, 1]
    assert tuple(ei.shape) == (2, 40)
    cond, t = torch.rand(2, 3, generator=g), torch.tensor([0.314])
    xh = [torch.rand(int(frags[i].sum()), node_nfs[i], generator=g) for i in range(3)]
    ea = torch.rand(ei.size(1), 3, generator=g)
    out = _dyn_call(dyn, xh, ei, t, cond, nfs, cm, edge_attr=ea)
    assert [tuple(o.shape) for o in out] == [tuple(x.shape) for x in xh]
    assert all(torch.isfinite(o).all() for o in out)
    _dyn_call(dyn, out, ei, t, cond, nfs, cm)  # the prediction has the layout of the input
    out_t = _dyn_call(dyn, xh, ei, torch.tensor([314]), cond, nfs, cm, edge_attr=ea)
    out_c = _dyn_call(dyn, xh, ei, t, torch.rand(2, 3, generator=g), nfs, cm, edge_attr=ea)
    for ii in range(3):
        assert not torch.allclose(out[ii], out_t[ii], rtol=1e-3)
        assert not torch.allclose(out[ii], out_c[ii], rtol=1e-4)
