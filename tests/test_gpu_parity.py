"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI (ctypes), against the
fp64 oracle on identical inputs, against golden vectors of the unmodified reference, and through size-independent
properties at the full BASELINE size.

Tolerances (stated): integer artefacts (mask, group ids, active list) bit-exact; floats
max|x - ref64| / max|ref64| <= REL_TOL = 2e-4 for one forward and its stages (measured on the B200: 1e-7 ... 4e-5; the
reference's own fp32-vs-fp64 gap is 2e-4 ... 6e-3 on this model), TRAJ_TOL = 1e-3 for the 11-evaluation trajectories.
"""
import numpy as np
import pytest
import torch

import oareactdiff_b200 as ob
from oracle import oa_ref
from tests.util import dyn_state_dict, leftnet_state_dict, load_golden, rel_err

pytestmark = pytest.mark.gpu
REL_TOL = 2e-4
TRAJ_TOL = 1e-3
DEV = torch.device("cuda:0")


def make_leftnet(cfg, sd):
    m = ob.LEFTNetB200(**cfg)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV)


def make_dynamics(cfg, sd, node_nfs=(9, 9, 9), condition_nf=1):
    dyn = ob.EGNNDynamics(model_config=cfg, fragment_names=[f"f{i}" for i in range(len(node_nfs))],
                          node_nfs=list(node_nfs), edge_nf=0, condition_nf=condition_nf, model=ob.LEFTNetB200, device=DEV)
    dyn.load_state_dict(sd, strict=True)
    return dyn.to(DEV)


@pytest.mark.parametrize("name", ["leftnet_small_full", "leftnet_small_cut", "leftnet_small_split"])
def test_leftnet_vs_reference_golden(name):
    g = load_golden(name)
    m = make_leftnet(g["cfg"], leftnet_state_dict(g))
    h, pos = torch.from_numpy(g["h"]).float().to(DEV), torch.from_numpy(g["pos"]).float().to(DEV)
    ho, po, _ = m(h, pos, torch.from_numpy(g["edge_index"]).to(DEV), subgraph_mask=torch.from_numpy(g["subgraph_mask"]).to(DEV))
    dpos = (po - pos).cpu()
    e_h, e_p = rel_err(ho.cpu(), g["h_out_f64"]), rel_err(dpos, g["dpos_f64"])
    ref_h, ref_p = rel_err(g["h_out_f32"], g["h_out_f64"]), rel_err(g["dpos_f32"], g["dpos_f64"])
    print(f"{name}: h {e_h:.2e} (ref fp32 gap {ref_h:.2e})  dpos {e_p:.2e} (ref fp32 gap {ref_p:.2e})")
    assert e_h < REL_TOL and e_p < REL_TOL


@pytest.mark.parametrize("name", ["dyn_small_ragged", "dyn_trained_cfg1", "dyn_trained_b4", "dyn_trained_b3_far"])
def test_dynamics_vs_reference_golden_and_integer_artefacts(name):
    g = load_golden(name)
    nfs_list = [int(x) for x in g["node_nfs"]]
    dyn = make_dynamics(g["cfg"], dyn_state_dict(g), nfs_list, int(g["condition_nf"]))
    eng = dyn.model.engine(DEV)
    eng.set_debug(True)
    xh = [torch.from_numpy(g[f"xh{f}"]).float().to(DEV) for f in range(len(nfs_list))]
    out, _ = dyn(xh, torch.from_numpy(g["edge_index"]).to(DEV), torch.from_numpy(g["t"]).float().to(DEV),
                 torch.from_numpy(g["cond"]).float().to(DEV), torch.from_numpy(g["n_frag_switch"]).to(DEV),
                 torch.from_numpy(g["combined_mask"]).to(DEV))
    # integer artefacts: bit-exact against the unmodified reference
    assert np.array_equal(eng.read("mask", torch.uint8).numpy().astype(np.int64), g["mask"])
    assert np.array_equal(eng.read("group", torch.int32).numpy().astype(np.int64), g["group"])
    n_act = int(eng.read("n_act", torch.int32)[0])
    assert n_act == int(g["mask"].sum())
    assert np.array_equal(eng.read("act_idx", torch.int32).numpy()[:n_act], np.nonzero(g["mask"])[0])
    for f in range(len(nfs_list)):
        if g[f"out{f}_f64"].size:
            e = rel_err(out[f].cpu(), g[f"out{f}_f64"])
            ref = rel_err(g[f"out{f}_f32"], g[f"out{f}_f64"])
            print(f"{name} frag{f}: {e:.2e} (ref fp32 gap {ref:.2e})")
            assert e < REL_TOL
    eng.set_debug(False)


def _oracle_inputs(cfg, sizes, seed, pos_scale=1.5):
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, seed)
    masks = [oa_ref.get_mask_for_frag(n) for n in nodes]
    cm = torch.cat(masks)
    ei = oa_ref.get_edges_index(cm, remove_self_edge=True)
    nfs = oa_ref.get_n_frag_switch(nodes)
    gen = torch.Generator().manual_seed(seed)
    xh = [torch.cat([oa_ref.remove_mean_batch(torch.randn(h.size(0), 3, generator=gen) * pos_scale, m), h], dim=1)
          for h, m in zip(h0, masks)]
    t = torch.rand(len(sizes), 1, generator=gen)
    return nodes, h0, cond, masks, cm, ei, nfs, xh, t


def test_intermediates_vs_oracle_trained_cfg():
    """Stage-by-stage parity of one forward (trained config, ragged B=3) against the fp64 oracle."""
    cfg = dict(oa_ref.TRAINED_CFG)
    sizes = [6, 13, 9]
    nodes, h0, cond, masks, cm, ei, nfs, xh, t = _oracle_inputs(cfg, sizes, seed=5)
    sd = oa_ref.make_state_dict(oa_ref.dynamics_param_shapes(cfg, [9, 9, 9], 1), 5, cfg, prefix_model="model.")
    dbg = {}
    ref = oa_ref.dynamics_forward({k: v.double() for k, v in sd.items()}, cfg, [x.double() for x in xh], ei, t.double(),
                                  cond.double(), nfs, cm, dbg=dbg)
    dyn = make_dynamics(cfg, sd)
    eng = dyn.model.engine(DEV)
    eng.set_debug(True)
    out, _ = dyn([x.to(DEV) for x in xh], ei.to(DEV), t.to(DEV), cond.to(DEV), nfs.to(DEV), cm.to(DEV))
    N, E, H = cm.numel(), ei.size(1), cfg["hidden_channels"]
    D = 3 * H + cfg["num_radial"]
    assert np.array_equal(eng.read("mask", torch.uint8).numpy(), dbg["mask"].numpy().astype(np.uint8))
    assert np.array_equal(eng.read("group", torch.int32).numpy().astype(np.int64), dbg["group"].numpy())
    report = {}
    report["pos_frame"] = rel_err(eng.read("pos_frame").view(N, 3), dbg["pos_frame"])
    geo = eng.read("geo").view(E, 4)
    report["dist"] = rel_err(geo[:, 3], dbg["dist"])
    report["coord_diff"] = rel_err(geo[:, :3], dbg["coord_diff"])
    report["s0"] = rel_err(eng.read("s0").view(N, H), dbg["s0"])
    report["NE1"] = rel_err(eng.read("NE1").view(N, 3, H), dbg["NE1"])
    report["e0"] = rel_err(eng.read("e0").view(E, D), dbg["e0"])
    report["pos_prjt"] = rel_err(eng.read("pos_prjt").view(N, 3), dbg["pos_prjt"])
    for l in range(cfg["num_layers"]):
        report[f"s_msg{l}"] = rel_err(eng.read(f"s_msg{l}").view(N, H), dbg[f"s_msg{l}"])
        report[f"vec_msg{l}"] = rel_err(eng.read(f"vec_msg{l}").view(N, 3, H), dbg[f"vec_msg{l}"])
        report[f"e{l + 1}"] = rel_err(eng.read(f"e{l + 1}").view(E, D), dbg[f"e{l + 1}"])
        report[f"s{l + 1}"] = rel_err(eng.read(f"s{l + 1}").view(N, H), dbg[f"s{l + 1}"])
        report[f"vec{l + 1}"] = rel_err(eng.read(f"vec{l + 1}").view(N, 3, H), dbg[f"vec{l + 1}"])
    for f in range(3):
        report[f"out{f}"] = rel_err(out[f].cpu(), ref[f])
    print("stage parity (max rel err vs fp64 oracle):")
    for k, v in report.items():
        print(f"  {k:12s} {v:.3e}")
    eng.set_debug(False)
    bad = {k: v for k, v in report.items() if not v < REL_TOL}
    assert not bad, bad


def test_determinism_and_launch_count():
    cfg = dict(oa_ref.TRAINED_CFG)
    nodes, h0, cond, masks, cm, ei, nfs, xh, t = _oracle_inputs(cfg, [8, 5], seed=6)
    sd = oa_ref.make_state_dict(oa_ref.dynamics_param_shapes(cfg, [9, 9, 9], 1), 6, cfg, prefix_model="model.")
    dyn = make_dynamics(cfg, sd)
    args = ([x.to(DEV) for x in xh], ei.to(DEV), t.to(DEV), cond.to(DEV), nfs.to(DEV), cm.to(DEV))
    a, _ = dyn(*args)
    b, _ = dyn(*args)
    for x, y in zip(a, b):
        assert torch.equal(x, y)  # fixed-order reductions, no atomics: bitwise reproducible
    assert dyn.model.engine(DEV).launches() > 50


class _CpuNoiseDiffusion(ob.EnVariationalDiffusion):
    """Draw the sampler's noise from the CPU generator in the reference's order so trajectories can be compared
    with golden outputs of the unmodified reference (which ran on CPU)."""

    def sample_combined_position_feature_noise(self, masks):
        out = []
        for ii, mask in enumerate(masks):
            x = torch.randn((len(mask), self.pos_dim))
            x = oa_ref.remove_mean_batch(x, mask.cpu())
            hh = torch.randn((len(mask), self.node_nfs[ii] - self.pos_dim))
            out.append(torch.cat([x, torch.zeros_like(hh)], dim=1).to(mask.device))
        return out


def _make_ddpm(cfg, seed, T, cls=ob.EnVariationalDiffusion):
    sd = oa_ref.make_state_dict(oa_ref.dynamics_param_shapes(cfg, [9, 9, 9], 1), seed, cfg, prefix_model="model.")
    dyn = make_dynamics(cfg, sd)
    sched = ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", T, 1e-5), norm_values=(1.0, 1.0, 1.0))
    return cls(dynamics=dyn, schdule=sched, normalizer=ob.Normalizer(), pos_only=True).to(DEV)


@pytest.mark.parametrize("name", ["sample_small_T10", "sample_trained_cfg1_T10"])
def test_sample_trajectory_vs_reference_golden(name):
    """config 1: full sample() through the public API with the reference's noise stream."""
    g = load_golden(name)
    sizes = [int(x) for x in g["sizes"]]
    ddpm = _make_ddpm(g["cfg"], int(g["seed"]), int(g["T"]), _CpuNoiseDiffusion)
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, int(g["seed"]))
    torch.manual_seed(int(g["seed"]))
    out, masks = ddpm.sample(len(sizes), [n.to(DEV) for n in nodes], cond.to(DEV), h0=[h.to(DEV) for h in h0])
    assert ddpm.n_evals == int(g["T"]) + 1
    for f in range(3):
        e = rel_err(out[0][f][:, :3].cpu(), g[f"out{f}"][:, :3])
        print(f"{name} frag{f}: trajectory rel err {e:.2e}")
        assert e < TRAJ_TOL  # 11 chained evaluations amplify the per-forward gap
        assert np.array_equal(out[0][f][:, 3:].cpu().numpy(), g[f"out{f}"][:, 3:])


@pytest.mark.parametrize("name", ["inpaint_small_T12_r2_j3", "inpaint_trained_b4_T12_r2_j3"])
def test_inpaint_trajectory_vs_reference_golden(name):
    """RePaint (en_diffusion.py:722-883) against trajectories of the UNMODIFIED reference: the small configuration and the
    trained one (B = 4 ragged reactions, T = 12, resamplings 2, jump length 3), reactant and product clamped."""
    g = load_golden(name)
    sizes = [int(x) for x in g["sizes"]]
    ddpm = _make_ddpm(g["cfg"], int(g["seed"]), int(g["T"]), _CpuNoiseDiffusion)
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, int(g["seed"]))
    xh_fixed = [torch.from_numpy(g[f"xh_fixed{f}"]).to(DEV) for f in range(3)]
    torch.manual_seed(int(g["seed"]))
    out, _ = ddpm.inpaint(len(sizes), [n.to(DEV) for n in nodes], cond.to(DEV), resamplings=int(g["resamplings"]),
                          jump_length=int(g["jump_length"]), xh_fixed=xh_fixed, frag_fixed=[0, 2])
    assert ddpm.n_evals == sum(oa_ref.get_repaint_schedule(2, 3, 12)) + 1
    for f in range(3):
        e = rel_err(out[0][f][:, :3].cpu(), g[f"out{f}"][:, :3])
        print(f"{name} frag{f}: {e:.2e}")
        assert e < TRAJ_TOL


def _rot(seed):
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=torch.Generator().manual_seed(seed), dtype=torch.float64))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q.float()


def test_full_size_properties_b64():
    """BASELINE size (B=64 Transition1x-shaped, trained config): size-independent properties of the reference's tests
    (tests/model/test_subgraphs.py:88-180): object-wise rotation/translation equivariance; zero CoM of the score;
    plus agreement with the oracle on a sub-batch (reactions are independent)."""
    cfg = dict(oa_ref.TRAINED_CFG)
    sizes = oa_ref.t1x_sizes(64, seed=0)
    nodes, h0, cond, masks, cm, ei, nfs, xh, t = _oracle_inputs(cfg, sizes, seed=7)
    sd = oa_ref.make_state_dict(oa_ref.dynamics_param_shapes(cfg, [9, 9, 9], 1), 7, cfg, prefix_model="model.")
    dyn = make_dynamics(cfg, sd)
    dargs = (ei.to(DEV), t.to(DEV), cond.to(DEV), nfs.to(DEV), cm.to(DEV))
    base, _ = dyn([x.to(DEV) for x in xh], *dargs)
    assert all(torch.isfinite(b).all() for b in base)
    B = len(sizes)
    for f in range(3):  # CoM of the position score is zero per (fragment, sample)
        com = torch.zeros(B, 3, device=DEV).index_add_(0, masks[f].to(DEV), base[f][:, :3])
        assert com.abs().max() < 1e-4 * max(1.0, float(base[f][:, :3].abs().max()))
    # rotate + translate only the TS fragment of every reaction: h-part invariant, that fragment's score co-rotates,
    # the other fragments' scores are unchanged
    Rm = _rot(3)
    xh2 = [x.clone() for x in xh]
    xh2[1][:, :3] = xh2[1][:, :3] @ Rm + torch.tensor([0.7, -1.1, 0.4])
    rot, _ = dyn([x.to(DEV) for x in xh2], *dargs)
    scale = max(float(b[:, :3].abs().max()) for b in base)
    assert (rot[1][:, :3] - base[1][:, :3] @ Rm.to(DEV)).abs().max() < 2e-3 * scale
    for f in (0, 2):
        assert (rot[f][:, :3] - base[f][:, :3]).abs().max() < 2e-3 * scale
    for f in range(3):
        assert (rot[f][:, 3:] - base[f][:, 3:]).abs().max() < 2e-3 * max(1.0, float(base[f][:, 3:].abs().max()))
    # the WHOLE batch against the fp64 oracle: reactions are independent, so the oracle evaluates it in chunks of 8
    worst = 0.0
    node_off = [torch.cat([torch.zeros(1, dtype=torch.long), n.cumsum(0)]) for n in nodes]
    sd64 = {kk: v.double() for kk, v in sd.items()}
    for k0 in range(0, B, 8):
        k1 = min(B, k0 + 8)
        sub_nodes = [n[k0:k1] for n in nodes]
        sub_masks = [oa_ref.get_mask_for_frag(n) for n in sub_nodes]
        sub_cm = torch.cat(sub_masks)
        sub_xh = [x[int(o[k0]):int(o[k1])] for x, o in zip(xh, node_off)]
        ref = oa_ref.dynamics_forward(sd64, cfg, [x.double() for x in sub_xh], oa_ref.get_edges_index(sub_cm, remove_self_edge=True),
                                      t[k0:k1].double(), cond[k0:k1].double(), oa_ref.get_n_frag_switch(sub_nodes), sub_cm)
        for f in range(3):
            worst = max(worst, rel_err(base[f][int(node_off[f][k0]):int(node_off[f][k1])].cpu(), ref[f]))
    print(f"B=64 whole batch vs fp64 oracle: {worst:.2e}")
    assert worst < REL_TOL


def test_cutoff_boundary_mask_is_bit_exact():
    """The edge mask is an integer contract (leftnet.py:747-753): `dist_raw < cutoff` with the reference's fp32 rounding
    sequence.  Pairs of atoms are placed so that the computed fp32 distance lands within a few ulp of the cutoff (both
    sides, and exactly on it); the CUDA mask and the greedy group labels that follow from it must equal the oracle's
    evaluated in fp32."""
    cutoff = 2.5
    cfg = dict(cutoff=cutoff, num_layers=1, hidden_channels=32, num_radial=16, in_hidden_channels=6, reflect_equiv=True,
               legacy=True, update=True, object_aware=True)
    g = torch.Generator().manual_seed(11)
    n = 96
    pos = torch.zeros(n, 3)
    pos[0::2] = torch.rand(n // 2, 3, generator=g) * 40.0  # pairs far from each other: only the partner is near the cutoff
    u = torch.randn(n // 2, 3, generator=g)
    u = u / u.norm(dim=1, keepdim=True)
    u[:8] = torch.eye(3).repeat(3, 1)[:8]  # axis-aligned pairs: the distance is computed exactly
    k = torch.randint(-4, 5, (n // 2, 1), generator=g).float()
    d = torch.tensor(cutoff) * (1.0 + k * 2.0 ** -23)
    d[:8] = torch.tensor([cutoff, float(np.nextafter(np.float32(cutoff), np.float32(0))),
                          float(np.nextafter(np.float32(cutoff), np.float32(9))), cutoff, cutoff, cutoff, cutoff, cutoff])[:, None]
    pos[1::2] = pos[0::2] + u * d
    ei = torch.tensor([[i, j] for i in range(n) for j in range(n) if i != j]).T.contiguous()
    h = torch.rand(n, 6, generator=g)
    dist = (pos[ei[0]] - pos[ei[1]]).pow(2).sum(dim=-1).sqrt()  # the reference's expression (leftnet.py:747), fp32 on the CPU
    ref_mask = (dist < cutoff)
    near = (dist - cutoff).abs() < 4e-6
    assert int(near.sum()) >= n // 2 and 0 < int((ref_mask & near).sum()) < int(near.sum())  # both sides are populated
    sd = oa_ref.make_state_dict(oa_ref.leftnet_param_shapes(cfg), 3, cfg)
    m = make_leftnet(cfg, sd)
    eng = m.engine(DEV)
    eng.set_debug(True)
    m(h.to(DEV), pos.to(DEV), ei.to(DEV), subgraph_mask=None)
    assert torch.equal(eng.read("mask", torch.uint8).bool(), ref_mask)
    group = oa_ref.assemble_nodemask(ei[:, ref_mask], n)
    assert np.array_equal(eng.read("group", torch.int32).numpy().astype(np.int64), group.numpy())
    eng.set_debug(False)


def test_c_abi_error_codes():
    import ctypes as C
    from oareactdiff_b200 import _lib
    lib = _lib.load()
    cfg = _lib.OardCfg(32, 16, 1, 8, 5.0, 1, 1, 1, 1)
    h = C.c_void_p()
    assert lib.oard_create(C.byref(cfg), 0, C.byref(h)) == 0
    x = torch.zeros(4, 8, device=DEV)
    # forward before commit/plan -> OARD_ESTATE
    assert lib.oard_forward(h, C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()), None, C.c_void_p(x.data_ptr()),
                            C.c_void_p(x.data_ptr()), None) == -3
    bad = torch.tensor([[0, 1], [1, 2]], dtype=torch.int64)  # (1,2) has no transpose
    assert lib.oard_plan(h, 3, 2, C.c_void_p(bad.data_ptr())) == -4
    assert b"symmetric" in lib.oard_last_error()
    unsorted = torch.tensor([[1, 0], [0, 1]], dtype=torch.int64)
    assert lib.oard_plan(h, 2, 2, C.c_void_p(unsorted.data_ptr())) == -4
    legacy_off = _lib.OardCfg(32, 16, 1, 8, 5.0, 1, 0, 1, 1)
    h2 = C.c_void_p()
    assert lib.oard_create(C.byref(legacy_off), 0, C.byref(h2)) == -1
    lib.oard_destroy(h)


def _sample_once(debug_asserts, seed, T=12, inpaint=False, use_fused=True):
    cfg = dict(oa_ref.TRAINED_CFG, hidden_channels=32, num_radial=16, num_layers=2, cutoff=5.0)
    sd = oa_ref.make_state_dict(oa_ref.dynamics_param_shapes(cfg, [9, 9, 9], 1), 3, cfg, prefix_model="model.")
    dyn = make_dynamics(cfg, sd)
    dyn.use_fused = use_fused
    sched = ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", T, 1e-5), norm_values=(1.0, 1.0, 1.0))
    ddpm = ob.EnVariationalDiffusion(dynamics=dyn, schdule=sched, normalizer=ob.Normalizer(), pos_only=True,
                                     debug_asserts=debug_asserts).to(DEV)
    sizes = [5, 9, 4]
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, 1)
    torch.manual_seed(seed)
    if inpaint:
        g = torch.Generator().manual_seed(5)
        xh_fixed = [torch.cat([torch.randn(h.size(0), 3, generator=g), h], dim=1).to(DEV) for h in h0]
        out, _ = ddpm.inpaint(len(sizes), [n.to(DEV) for n in nodes], cond.to(DEV), resamplings=2, jump_length=3,
                              xh_fixed=xh_fixed, frag_fixed=[0, 2])
    else:
        out, _ = ddpm.sample(len(sizes), [n.to(DEV) for n in nodes], cond.to(DEV), h0=[h.to(DEV) for h in h0])
    return [o.cpu() for o in out[0]], ddpm.n_evals


@pytest.mark.parametrize("inpaint", [False, True])
def test_fast_sampler_path_equals_reference_structured_path(inpaint):
    """The tabulated / concatenated reverse loop (default) against the reference-structured per-fragment loop
    (debug_asserts=True forces it) with the same CUDA RNG seed: same draws in the same order, same formulas."""
    fast, n1 = _sample_once(False, 11, inpaint=inpaint)
    slow, n2 = _sample_once(True, 11, inpaint=inpaint)
    assert n1 == n2
    for a, b in zip(fast, slow):
        assert rel_err(a[:, :3], b[:, :3]) < 2e-4, rel_err(a[:, :3], b[:, :3])
        assert torch.equal(a[:, 3:], b[:, 3:])


def _dyn_inputs(name):
    g = load_golden(name)
    nfs_list = [int(x) for x in g["node_nfs"]]
    dyn = make_dynamics(g["cfg"], dyn_state_dict(g), nfs_list, int(g["condition_nf"]))
    xh = [torch.from_numpy(g[f"xh{f}"]).float().to(DEV) for f in range(len(nfs_list))]
    args = (xh, torch.from_numpy(g["edge_index"]).to(DEV), torch.from_numpy(g["t"]).float().to(DEV),
            torch.from_numpy(g["cond"]).float().to(DEV), torch.from_numpy(g["n_frag_switch"]).to(DEV),
            torch.from_numpy(g["combined_mask"]).to(DEV))
    return g, dyn, args


@pytest.mark.parametrize("name", ["dyn_trained_cfg1", "dyn_trained_b4", "dyn_trained_b3_far"])
def test_device_dynamics_equals_host_composed_dynamics(name):
    """oard_dyn_forward (encoders, time/condition channels, NaN guard, CoM removal, decoders as CUDA kernels around the
    LEFTNet graph) against the torch-composed wrapper around the same LEFTNet kernels, and against the reference golden."""
    g, dyn, args = _dyn_inputs(name)
    assert dyn.fused_ok(DEV)
    for _ in range(3):  # eager first call, then the cached forward graph
        fused, _ = dyn(*args)
    dyn.use_fused = False
    host, _ = dyn(*args)
    for f in range(len(fused)):
        e = rel_err(fused[f].cpu(), host[f].cpu())
        e64 = rel_err(fused[f].cpu(), g[f"out{f}_f64"])
        print(f"{name} frag{f}: fused vs host-composed {e:.2e}; vs fp64 reference {e64:.2e}")
        assert e < 1e-4 and e64 < REL_TOL  # encoder/decoder MLPs: in-kernel fma chains vs cuBLAS


def test_device_nan_guard_replaces_output_with_noise():
    """egnn_dynamics.py:138-143: a NaN anywhere in the velocity replaces the whole velocity by noise (then CoM-free)."""
    g, dyn, args = _dyn_inputs("dyn_trained_b4")
    xh = [x.clone() for x in args[0]]
    xh[1][0, 0] = float("nan")
    out, _ = dyn(xh, *args[1:])
    vel = torch.cat([o[:, :3] for o in out])
    assert bool(torch.isfinite(vel).all()) and 0.5 < float(vel.std()) < 1.5


@pytest.mark.parametrize("inpaint", [False, True])
def test_device_reverse_step_equals_host_fast_step(inpaint):
    """oard_reverse_step (one CUDA graph per step) against the host-composed `_fast_step` with the same CUDA RNG seed."""
    devp, n1 = _sample_once(False, 11, inpaint=inpaint)
    hostp, n2 = _sample_once(False, 11, inpaint=inpaint, use_fused=False)
    assert n1 == n2
    for a, b in zip(devp, hostp):
        e = rel_err(a[:, :3], b[:, :3])
        print(f"device step vs host step (inpaint={inpaint}): {e:.2e}")
        assert e < 2e-4, e
        assert torch.equal(a[:, 3:], b[:, 3:])
