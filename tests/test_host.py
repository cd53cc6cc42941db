"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the drop-in modules carry
the reference's state-dict names, host helpers agree with the oracle, and the product fails loudly without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import oareactdiff_b200 as ob
from oareactdiff_b200 import _lib
from oracle import oa_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "oard.h")).read()
    declared = set(re.findall(r"\b(oard_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 14
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert _lib.load().oard_abi_version() == 1


@pytest.mark.parametrize("cfg", [oa_ref.TRAINED_CFG, dict(oa_ref.TRAINED_CFG, hidden_channels=32, num_radial=16, num_layers=2)])
def test_state_dict_names_match_reference(cfg):
    # oracle shapes were strict-loaded into the reference LEFTNet/EGNNDynamics by oracle/gen_golden.py
    model = ob.LEFTNetB200(**cfg)
    want = oa_ref.leftnet_param_shapes(cfg)
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == want
    dyn = ob.EGNNDynamics(model_config=cfg, fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0,
                          condition_nf=1, model=ob.LEFTNetB200, device=torch.device("cpu"))
    want = oa_ref.dynamics_param_shapes(cfg, [9, 9, 9], 1)
    got = {k: tuple(v.shape) for k, v in dyn.state_dict().items()}
    assert got == want
    n_params = sum(p.numel() for p in dyn.parameters())
    if cfg["hidden_channels"] == 196:
        assert n_params == 10_645_719  # SURVEY App. B (probe-verified on the reference)


def test_unsupported_options_raise():
    with pytest.raises(NotImplementedError):
        ob.LEFTNetB200(legacy=False)
    with pytest.raises(NotImplementedError):
        ob.LEFTNetB200(pos_grad=True)


def test_no_cpu_fallback():
    cfg = dict(oa_ref.TRAINED_CFG, hidden_channels=32, num_radial=16, num_layers=1)
    m = ob.LEFTNetB200(**cfg)
    ei = torch.tensor([[0, 1], [1, 0]])
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 8), torch.zeros(2, 3), ei)


def test_graph_tools_kats():
    # reference tests/utils/test_graph_tools.py:14-63
    assert ob.get_mask_for_frag(torch.tensor([2, 0, 3])).tolist() == [0, 0, 2, 2, 2]
    assert ob.get_n_frag_switch([torch.tensor([2, 0]), torch.tensor([1, 3]), torch.tensor([3, 2])]).tolist() == \
        [0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 2]
    ei = torch.tensor([[0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1]])
    assert ob.get_subgraph_mask(ei, torch.tensor([0, 0, 1])).tolist() == [1, 0, 1, 0, 0, 0]
    frags = [torch.tensor([2, 0]), torch.tensor([2, 3]), torch.tensor([1, 2])]
    cm = torch.cat([ob.get_mask_for_frag(n) for n in frags])
    assert ob.get_edges_index(cm).shape == (2, 50)
    ei = ob.get_edges_index(cm, remove_self_edge=True)
    assert ei.shape == (2, 40)
    assert int(ob.get_subgraph_mask(ei, ob.get_n_frag_switch(frags)).sum()) == 12


def test_edges_index_bit_exact_vs_oracle():
    sizes = oa_ref.t1x_sizes(16, seed=3)
    nodes = [torch.tensor(sizes)] * 3
    cm = torch.cat([ob.get_mask_for_frag(n) for n in nodes])
    assert torch.equal(ob.get_edges_index(cm, remove_self_edge=True), oa_ref.get_edges_index(cm, remove_self_edge=True))


def test_schedule_matches_oracle():
    for name, T in [("polynomial_2", 1000), ("cosine", 5000), ("polynomial_2", 10)]:
        assert torch.equal(ob.PredefinedNoiseSchedule(name, T, 1e-5).gamma.data, oa_ref.gamma_table(name, T, 1e-5))
    for r, j, T in [(1, 1, 1000), (5, 5, 150), (5, 5, 1000), (2, 3, 12), (3, 7, 20), (1, 5, 3)]:
        assert ob.get_repaint_schedule(r, j, T) == oa_ref.get_repaint_schedule(r, j, T)


def test_engine_plan_accepts_edges_not_grouped_by_source():
    """The reference's own model tests hand-write edge lists in arbitrary order (tests/model/test_equiv.py:30-32); the shim
    brings them into the source-grouped order oard_plan wants and permutes the per-edge subgraph_mask with them.  Checked
    here against a recording stand-in for the C library (no CUDA needed for this host logic)."""
    from oareactdiff_b200.leftnet import _Engine

    seen = {}

    class FakeLib:
        def oard_plan(self, h, n_nodes, n_edges, ptr):
            buf = (ctypes.c_int64 * (2 * n_edges)).from_address(ptr.value)
            seen["ei"] = np.array(buf, dtype=np.int64).reshape(2, n_edges)
            seen["n"] = n_nodes
            return 0

    eng = object.__new__(_Engine)
    eng.lib, eng.h, eng.device, eng.plan_key, eng.edge_perm = FakeLib(), None, torch.device("cpu"), None, None
    ei = torch.tensor([[0, 1, 1, 2, 3, 0], [1, 0, 2, 1, 0, 3]])
    eng.plan(ei, 4)
    assert seen["n"] == 4 and eng.E == 6
    assert np.array_equal(seen["ei"], [[0, 0, 1, 1, 2, 3], [1, 3, 0, 2, 1, 0]])  # stable: (0,1) stays before (0,3)
    sub = torch.tensor([[1], [0], [1], [1], [0], [1]])  # [E, 1] like the callers pass it
    assert eng.edge_order(sub).tolist() == [1, 1, 0, 1, 1, 0]
    assert eng.edge_order(None) is None
    with pytest.raises(ValueError):
        eng.edge_order(torch.ones(5))
    # grouped by source but targets not ascending inside a row: sorted as well (the group-staged message kernel walks the
    # members of a fragment in row order and addresses them by rank)
    ei3 = torch.tensor([[0, 0, 1, 1, 2, 2], [2, 1, 0, 2, 1, 0]])
    eng.plan(ei3, 3)
    assert np.array_equal(seen["ei"], [[0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1]])
    assert eng.edge_order(torch.tensor([10, 11, 12, 13, 14, 15])).tolist() == [11, 10, 12, 13, 15, 14]
    # already sorted (what get_edges_index produces): passed through untouched, and the mask keeps its storage
    ei2 = ob.get_edges_index(torch.tensor([0, 0, 0, 1, 1]), remove_self_edge=True)
    eng.plan(ei2, 5)
    assert eng.edge_perm is None and np.array_equal(seen["ei"], ei2.numpy())
    flat = torch.ones(ei2.size(1), dtype=torch.int64)
    assert eng.edge_order(flat).data_ptr() == flat.data_ptr()
    del eng.h  # (nothing to destroy)


def test_constructors_accept_the_reference_test_fixtures():
    """Configs copied from the reference's own fixtures: tests/model/utils.py:24-32 (LEFTNet), tests/dynamics/
    test_egnn_dynamics.py:50-62 and test_switch_fragments.py:34-50 (in_node_nf given, in_hidden_channels absent, edge_nf > 0
    with no in_edge_nf)."""
    left_config = dict(pos_require_grad=False, cutoff=20.0, num_layers=6, hidden_channels=32, num_radial=32, in_node_nf=8,
                       reflect_equiv=True)
    m = ob.LEFTNetB200(**left_config)
    assert m.cfg["in_hidden_channels"] == 8 and m.cfg["reflect_equiv"] is True
    assert ob.LEFTNetB200(**dict(left_config, reflect_equiv=False)).cfg["reflect_equiv"] is False
    leftnet_config = dict(pos_require_grad=False, cutoff=5.0, num_layers=2, hidden_channels=32, num_radial=8, in_node_nf=8)
    for node_nfs, names, edge_nf in (([4, 5, 6], ["inorg_node", "org_edge", "org_node"], 3), ([5, 5], ["A", "B"], 4)):
        cfg = dict(leftnet_config)
        dyn = ob.EGNNDynamics(model_config=cfg, node_nfs=node_nfs, edge_nf=edge_nf, condition_nf=3, fragment_names=names,
                              pos_dim=3, update_pocket_coords=True, condition_time=True, edge_cutoff=None,
                              model=ob.LEFTNetB200, device=torch.device("cpu"))
        assert dyn.embed_dim == 8 - 1 - 3 and dyn.edge_encoder is None
        assert [e.mlp[0].linear.in_features for e in dyn.encoders] == [n - 3 for n in node_nfs]
        assert [d.mlp[1].linear.out_features for d in dyn.decoders] == [n - 3 for n in node_nfs]


def _recording_dynamics():
    """EGNNDynamics whose engine is a recording stand-in (real `plan` / `edge_order`, no kernels); -> (dynamics, calls)."""
    from oareactdiff_b200.leftnet import _Engine

    calls = []

    class FakeLib:
        def oard_plan(self, h, n_nodes, n_edges, ptr):
            return 0

    class Rec(_Engine):
        def __init__(self):
            self.lib, self.h, self.device, self.plan_key, self.edge_perm = FakeLib(), None, torch.device("cpu"), None, None
            self.weights_key = self.dyn_weights_key = None
            self.N = self.E = 0

        def __del__(self):
            pass

        def sync_weights(self, module, force=False):
            self.weights_key = 1

        def dyn_sync(self, dynamics, n_frag, node_nf, condition_nf, condition_time):
            self.dyn_weights_key = (n_frag, node_nf, condition_nf, condition_time)

        def dyn_plan(self, nfs, cm, n_samples):
            calls.append(("dyn_plan", nfs.numel(), cm.numel(), n_samples))

        def dyn_forward(self, xh, t, cond, sub, out):
            calls.append(("dyn_forward", tuple(xh.shape), xh.dtype, None if t is None else tuple(t.shape), None if cond is None else tuple(cond.shape),
                          None if sub is None else (sub.dtype, sub.numel(), sub.is_contiguous())))
            return out.zero_()

        def reverse_step(self, z, nx, nh, h0, cond, sub, t, alpha_ts, coef, sigma):
            calls.append(("reverse_step", z.data_ptr(), tuple(nx.shape), nh is None, None if h0 is None else tuple(h0.shape), float(t)))

        def inpaint_step(self, z, nx, nh, h0, cond, sub, t, alpha_ts, coef, sigma, x_fixed, known_bits, kx, kh, alpha_s, sigma_s):
            calls.append(("inpaint_step", z.data_ptr(), tuple(nx.shape), nh is None, None if h0 is None else tuple(h0.shape), float(t),
                          tuple(x_fixed.shape), int(known_bits), kx.data_ptr() != nx.data_ptr(), kh is None, float(alpha_s), float(sigma_s)))

        def jump_back(self, z, nx, nh, alpha_ts, sigma_ts):
            calls.append(("jump_back", z.data_ptr(), tuple(nx.shape), nh is None, float(alpha_ts), float(sigma_ts)))

    cfg = dict(oa_ref.TRAINED_CFG, hidden_channels=32, num_radial=16, num_layers=1)
    dyn = ob.EGNNDynamics(model_config=cfg, fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0, condition_nf=1,
                          model=ob.LEFTNetB200, device=torch.device("cpu"))
    eng = Rec()
    dyn.model.engine = lambda device: eng
    dyn.fused_ok = lambda device: True
    return dyn, calls


def test_fused_dynamics_plumbing_with_a_recording_engine():
    """The Python side of the device-resident path (`EGNNDynamics._forward_fused`, `EnVariationalDiffusion._device_setup` /
    `_device_step`, sample() / inpaint() around them) cannot run without CUDA, but its plumbing can: a recording stand-in for
    the engine checks which tensors reach the C calls — shapes, dtypes, the same-fragment mask in the planned edge order,
    persistence of the state buffer across steps (the step graph is keyed by it), the step order incl. RePaint jump-backs."""
    dyn, calls = _recording_dynamics()
    sizes = [3, 5]
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, 0)
    ddpm = ob.EnVariationalDiffusion(dynamics=dyn, schdule=ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", 5, 1e-5), (1.0, 1.0, 1.0)),
                                     normalizer=ob.Normalizer(), pos_only=True)
    torch.manual_seed(0)
    out, masks = ddpm.sample(len(sizes), nodes, cond, h0=h0)
    N, E = 3 * sum(sizes), sum(3 * n * (3 * n - 1) for n in sizes)
    steps = [c for c in calls if c[0] == "reverse_step"]
    assert len(steps) == 5 and len({c[1] for c in steps}) == 1  # one persistent state buffer: the step graph is keyed by it
    assert steps[0][2] == (N, 3) and steps[0][3] and steps[0][4] == (N, 6) and [round(c[5] * 5) for c in steps] == [5, 4, 3, 2, 1]
    fwd = [c for c in calls if c[0] == "dyn_forward"]
    assert len(fwd) == 1 and fwd[0][1] == (N, 9) and fwd[0][2] == torch.float32 and fwd[0][3] == (2,) and fwd[0][4] == (2, 1)
    assert fwd[0][5] == (torch.int64, E, True)
    assert ("dyn_plan", N, N, 2) in calls and ddpm.n_evals == 6
    assert [tuple(o.shape) for o in out[0]] == [(sum(sizes), 9)] * 3
    # RePaint: one oard_inpaint_step per step (clamped fragments as a bit mask, their own noise buffers), jump-backs in place on
    # the same state buffer
    del calls[:]
    xh_fixed = [torch.cat([torch.randn(h.size(0), 3), h.float()], dim=1) for h in h0]
    ddpm.inpaint(len(sizes), nodes, cond, resamplings=2, jump_length=2, timesteps=4, xh_fixed=xh_fixed, frag_fixed=[0, 2])
    steps = [c for c in calls if c[0] == "inpaint_step"]
    sched = ob.get_repaint_schedule(2, 2, 4)
    assert len(steps) == sum(sched) == ddpm.n_evals - 1 and len({c[1] for c in steps}) == 1
    assert all(c[6] == (N, 9) and c[7] == 0b101 and c[8] and c[9] and 0 < c[10] <= 1 and 0 <= c[11] < 1 for c in steps)
    jumps = [c for c in calls if c[0] == "jump_back"]
    assert len(jumps) == len(sched) - 1 and {c[1] for c in jumps} == {steps[0][1]} and all(0 < c[4] <= 1 for c in jumps)
    assert not [c for c in calls if c[0] == "reverse_step"]
    want, s = [], 3
    for i, n in enumerate(sched):
        for j in range(n):
            want.append(s + 1)
            if j == n - 1 and i < len(sched) - 1:
                s += 2
            s -= 1
    assert [round(c[5] * 4) for c in steps] == want


def test_training_path_autograd_plumbing_with_a_recording_engine():
    """`enable_training_path`: LEFTNetB200.forward as an autograd node over oard_forward_train / oard_backward / oard_get_grad.
    With a stand-in engine that returns recognisable values, check the routing: every parameter (by reference name) receives
    the gradient buffer of ITS name, buffers and positions receive none, the node-feature gradient flows on into the encoders
    of EGNNDynamics, inference mode and torch.no_grad() keep using the inference entry."""
    from oareactdiff_b200.leftnet import _Engine

    calls = []

    class FakeLib:
        def oard_plan(self, h, n_nodes, n_edges, ptr):
            return 0

    class Rec(_Engine):
        def __init__(self, names):
            self.lib, self.h, self.device, self.plan_key, self.edge_perm = FakeLib(), None, torch.device("cpu"), None, None
            self.weights_key, self.names, self.N, self.E = None, names, 0, 0

        def __del__(self):
            pass

        def sync_weights(self, module, force=False):
            self.weights_key = 1

        def forward_train(self, h, pos, sub):
            calls.append(("forward_train", tuple(h.shape), None if sub is None else self.edge_order(sub).numel()))
            return 2.0 * h.detach(), torch.ones_like(pos)

        def forward(self, h, pos, sub):
            calls.append(("forward",))
            return torch.zeros_like(h), torch.zeros_like(pos)

        def backward(self, g_h, g_dpos):
            calls.append(("backward", float(g_h.sum()), float(g_dpos.sum())))
            return 2.0 * g_h

        def get_grad(self, name, like):
            return torch.full(like.shape, float(len(name)))

    cfg = dict(oa_ref.TRAINED_CFG, hidden_channels=32, num_radial=16, num_layers=2)
    dyn = ob.EGNNDynamics(model_config=cfg, fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0, condition_nf=1,
                          model=ob.LEFTNetB200, device=torch.device("cpu"))
    model = dyn.model
    names = [n for n in model._oard_tensors() if not n.startswith(("distance_embedding", "last_layer"))]
    eng = Rec(names)
    model.engine = lambda device: eng
    sizes = [3, 4]
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, 0)
    masks = [ob.get_mask_for_frag(n) for n in nodes]
    cm = torch.cat(masks)
    ei, nfs = ob.get_edges_index(cm, remove_self_edge=True), ob.get_n_frag_switch(nodes)
    xh = [torch.cat([torch.randn(h.size(0), 3), h.float()], dim=1) for h in h0]
    t = torch.rand(len(sizes), 1)
    out, _ = dyn(xh, ei, t, cond, nfs, cm)  # default: inference entry, no graph
    assert calls == [("forward",)]
    # caches keyed by (address, version) keep the keyed tensors alive: a freed block would be handed to the next tensor of
    # that size, and a different graph at the same address must not hit the cache
    assert dyn._graph_refs[0] is ei and dyn._graph_refs[1] is nfs and dyn._graph_refs[2] is cm and eng._plan_ref is ei
    model.enable_training_path = True
    with torch.no_grad():
        dyn(xh, ei, t, cond, nfs, cm)
    assert calls == [("forward",), ("forward",)]
    out, _ = dyn(xh, ei, t, cond, nfs, cm)
    assert calls[-1] == ("forward_train", (cm.numel(), 8), ei.size(1)) and out[0].requires_grad
    sum(o.sum() for o in out).backward()
    assert calls[-1][0] == "backward"
    params = dict(model.named_parameters())
    for n in names:
        if n in params:
            assert params[n].grad is not None and torch.equal(params[n].grad, torch.full_like(params[n], float(len(n)))), n
    for n, p in params.items():
        if n not in names:
            assert p.grad is None, n  # distance_embedding / last_layer: present in the state dict, unused by forward
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in dyn.encoders.parameters())  # through g_h_in
    assert all(p.grad is not None for p in dyn.decoders.parameters())
    # the handle keeps ONE forward's activations: backward through an older forward must fail loudly, not silently use the newer ones
    first, _ = dyn(xh, ei, t, cond, nfs, cm)
    second, _ = dyn(xh, ei, t, cond, nfs, cm)
    with pytest.raises(RuntimeError, match="overwritten by a later forward"):
        sum(o.sum() for o in first).backward()
    sum(o.sum() for o in second).backward()


def test_models_with_live_engines_can_be_copied_and_pickled():
    """The reference's trainer deep-copies the diffusion model before every sampling evaluation (pl_trainer.py:291) and
    `torch.save(model)` pickles it; a live engine (a C handle) must neither break that nor be shared: the copy carries none
    and builds its own at its first forward."""
    import copy
    import ctypes as C
    import io
    import pickle

    from oareactdiff_b200.leftnet import _Engine
    cfg = dict(oa_ref.TRAINED_CFG, hidden_channels=32, num_radial=16, num_layers=1)
    dyn = ob.EGNNDynamics(model_config=cfg, fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0, condition_nf=1,
                          model=ob.LEFTNetB200, device=torch.device("cpu"))
    ddpm = ob.EnVariationalDiffusion(dynamics=dyn, schdule=ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", 5, 1e-5), (1.0, 1.0, 1.0)),
                                     normalizer=ob.Normalizer(), pos_only=True)
    eng = object.__new__(_Engine)  # what a forward on a GPU leaves behind: a ctypes library object and a handle
    eng.lib, eng.h, eng.device = C.CDLL(None), None, torch.device("cpu")
    dyn.model._engines[torch.device("cpu")] = eng
    dyn._nan_gen = torch.Generator()
    dyn._graph, dyn._graph_key = dict(sub=torch.ones(4, 1)), (1, 2)
    ddpm._dev = dict(eng=eng, nx=torch.zeros(3, 3), views=[torch.zeros(2)], cond=None, sub=None, H0=None)

    def via_torch_save(m):
        buf = io.BytesIO()
        torch.save(m, buf)
        buf.seek(0)
        return torch.load(buf, weights_only=False)

    for clone in (copy.deepcopy, lambda m: pickle.loads(pickle.dumps(m)), via_torch_save):
        c = clone(ddpm)
        assert c.dynamics.model._engines == {torch.device("cpu"): None} and c._dev["eng"] is None
        assert all(torch.equal(a, b) and a.data_ptr() != b.data_ptr() for a, b in zip(c.parameters(), ddpm.parameters()))
    assert dyn.model._engines[torch.device("cpu")] is eng  # the original keeps its engine


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No built library -> RuntimeError naming the build command; a library that lacks a declared symbol -> AttributeError."""
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "liboard_b200.so"))
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        _lib.load()
    emu = os.path.join(ROOT, "oareactdiff_b200", "libtrain_emu.so")  # a real shared library without the oard_* entry points
    if os.path.exists(emu):
        monkeypatch.setattr(_lib, "LIB_PATH", emu)
        with pytest.raises(AttributeError):
            _lib.load()


def test_unsupported_activation_is_rejected():
    cfg = dict(oa_ref.TRAINED_CFG, hidden_channels=32, num_radial=16, num_layers=1, act_fn="relu")
    with pytest.raises(NotImplementedError, match="act_fn"):
        ob.EGNNDynamics(model_config=cfg, fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0, condition_nf=1,
                        model=ob.LEFTNetB200, device=torch.device("cpu"))


def test_forward_under_autograd_without_training_path_warns_once():
    """With `enable_training_path` off the inference kernels return detached outputs; a training loop would silently update
    everything but this module, so the first such forward warns (before the engine is even created)."""
    import warnings
    m = ob.LEFTNetB200(cutoff=5.0, num_layers=1, hidden_channels=32, num_radial=16, in_hidden_channels=8)
    h, pos, ei = torch.zeros(3, 8), torch.zeros(3, 3), torch.tensor([[0, 1], [1, 0]])
    with pytest.warns(UserWarning, match="enable_training_path"), pytest.raises(RuntimeError, match="CUDA"):
        m(h, pos, ei)
    with warnings.catch_warnings():
        warnings.simplefilter("error")  # second call: no second warning
        with pytest.raises(RuntimeError, match="CUDA"):
            m(h, pos, ei)
        with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
            ob.LEFTNetB200(cutoff=5.0, num_layers=1, hidden_channels=32, num_radial=16, in_hidden_channels=8)(h, pos, ei)


def test_strict_checkpoint_load_validates_before_copying():
    cfg = dict(oa_ref.TRAINED_CFG, hidden_channels=32, num_radial=16, num_layers=1)
    dyn = ob.EGNNDynamics(model_config=cfg, fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0, condition_nf=1,
                          model=ob.LEFTNetB200, device=torch.device("cpu"))
    ddpm = ob.EnVariationalDiffusion(dynamics=dyn, schdule=ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", 5, 1e-5), (1.0, 1.0, 1.0)),
                                     normalizer=ob.Normalizer(), pos_only=True)
    before = {k: v.clone() for k, v in ddpm.state_dict().items()}
    bad = {"ddpm." + k: torch.full_like(v, 7.0) for k, v in before.items()}
    bad.pop(next(iter(bad)))  # one key missing
    with pytest.raises(RuntimeError, match="missing"):
        ob.load_reference_checkpoint(ddpm, bad, strict=True)
    after = ddpm.state_dict()
    assert all(torch.equal(after[k], before[k]) for k in before)  # nothing was copied
    res = ob.load_reference_checkpoint(ddpm, bad, strict=False)
    assert len(res["missing"]) == 1 and res["loaded"] == len(before) - 1
