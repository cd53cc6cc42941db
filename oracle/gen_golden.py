"""ORACLE tooling (test infrastructure): generate tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference) through oracle/shims in the build container.

    python oracle/gen_golden.py            # writes tests/golden/

The reference cannot travel to the GPU box, so its outputs are frozen here as small fixtures.
Weights are NOT stored: both sides rebuild them with oracle.oa_ref.make_state_dict(seed).
"""
import json
import os
import pickle
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)

from oa_reactdiff.model import LEFTNet  # noqa: E402
from oa_reactdiff.dynamics import EGNNDynamics  # noqa: E402
from oa_reactdiff.diffusion._schedule import DiffSchedule, PredefinedNoiseSchedule  # noqa: E402
from oa_reactdiff.diffusion._normalizer import Normalizer  # noqa: E402
from oa_reactdiff.diffusion.en_diffusion import EnVariationalDiffusion  # noqa: E402
from oa_reactdiff.utils import get_edges_index, get_mask_for_frag, get_n_frag_switch, get_subgraph_mask  # noqa: E402

from oracle import oa_ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

SMALL_CFG = dict(cutoff=5.0, num_layers=2, hidden_channels=32, num_radial=16, in_hidden_channels=8,
                 reflect_equiv=True, legacy=True, update=True, object_aware=True)
TRAINED_CFG = dict(oa_ref.TRAINED_CFG)


def ref_cfg(cfg):
    c = dict(cfg)
    c.update(pos_require_grad=False, pos_grad=False, single_layer_output=True)
    return c


def build_dynamics(cfg, node_nfs, condition_nf, seed, dtype):
    dyn = EGNNDynamics(model_config=ref_cfg(cfg), fragment_names=[f"f{i}" for i in range(len(node_nfs))],
                       node_nfs=list(node_nfs), edge_nf=0, condition_nf=condition_nf, pos_dim=3,
                       update_pocket_coords=True, condition_time=True, model=LEFTNet, device=torch.device("cpu"))
    shapes = oa_ref.dynamics_param_shapes(cfg, node_nfs, condition_nf)
    sd = oa_ref.make_state_dict(shapes, seed, cfg, prefix_model="model.")
    missing, unexpected = dyn.load_state_dict(sd, strict=True), None
    return dyn.to(dtype), sd


def graph(fragments_nodes):
    masks = [get_mask_for_frag(n) for n in fragments_nodes]
    cm = torch.cat(masks)
    return masks, cm, get_edges_index(cm, remove_self_edge=True), get_n_frag_switch(fragments_nodes)


def capture_leftnet_intermediates(model, h, pos, edge_index, sub):
    """Re-run pieces of the reference forward to expose integer artefacts (mask, group ids)."""
    i, j = edge_index
    dist = (pos[i] - pos[j]).pow(2).sum(dim=-1).sqrt()
    mask = (dist < model.cutoff).to(pos.dtype)[:, None] * sub
    ei = edge_index.T[torch.where(mask > 0)[0]].T
    group = model.assemble_nodemask(edge_index=ei, pos=pos)
    return mask.squeeze(-1), group.long()


def case_dynamics(name, cfg, fragments_nodes, node_nfs, condition_nf, seed, pos_scale, t_vec=True):
    g = torch.Generator().manual_seed(seed)
    masks, cm, edge_index, nfs = graph(fragments_nodes)
    B = fragments_nodes[0].numel()
    xh = []
    for f, m in enumerate(masks):
        x = torch.randn(len(m), 3, generator=g, dtype=torch.float64) * pos_scale
        hh = torch.randn(len(m), node_nfs[f] - 3, generator=g, dtype=torch.float64)
        xh.append(torch.cat([x, hh], dim=1))
    t = torch.rand(B, 1, generator=g, dtype=torch.float64) if t_vec else torch.tensor([0.314], dtype=torch.float64)
    cond = torch.rand(B, condition_nf, generator=g, dtype=torch.float64)
    out = {}
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        dyn, sd = build_dynamics(cfg, node_nfs, condition_nf, seed, dt)
        with torch.no_grad():
            res, _ = dyn([x.to(dt) for x in xh], edge_index, t.to(dt), cond.to(dt), nfs, cm)
        for f, r in enumerate(res):
            out[f"out{f}_{tag}"] = r.numpy()
        if dt == torch.float64:
            pos = torch.cat([x[:, :3] for x in xh])
            sub = get_subgraph_mask(edge_index, nfs)[:, None]
            mask, group = capture_leftnet_intermediates(dyn.model, pos, pos, edge_index, sub.to(dt))
            out["mask"] = mask.numpy().astype(np.int64)
            out["group"] = group.numpy()
    out.update({f"xh{f}": x.numpy() for f, x in enumerate(xh)})
    out.update(t=t.numpy(), cond=cond.numpy(), edge_index=edge_index.numpy(), n_frag_switch=nfs.numpy(),
               combined_mask=cm.numpy(), seed=np.int64(seed),
               fragments_nodes=np.stack([n.numpy() for n in fragments_nodes]),
               node_nfs=np.array(node_nfs), condition_nf=np.int64(condition_nf),
               cfg=json.dumps(cfg))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "N", cm.numel(), "E", edge_index.size(1), "active", int(out["mask"].sum()),
          "groups", int(out["group"].max()) + 1)


def case_leftnet(name, cfg, n_nodes, seed, cut=None, pos_scale=1.0):
    """Raw LEFTNet.forward on one complete graph (reference tests/model fixtures style)."""
    g = torch.Generator().manual_seed(seed)
    ii, jj = torch.meshgrid(torch.arange(n_nodes), torch.arange(n_nodes), indexing="ij")
    keep = ii != jj
    edge_index = torch.stack([ii[keep], jj[keep]])
    h = torch.rand(n_nodes, cfg["in_hidden_channels"], generator=g, dtype=torch.float64)
    pos = torch.rand(n_nodes, 3, generator=g, dtype=torch.float64) * pos_scale
    if cut is None:
        sub = torch.ones(edge_index.size(1), 1, dtype=torch.long)
    else:
        s = (edge_index < cut).sum(0)
        sub = ((s == 2) | (s == 0)).long()[:, None]
    shapes = oa_ref.leftnet_param_shapes(cfg)
    sd = oa_ref.make_state_dict(shapes, seed, cfg)
    out = {}
    for dt, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        model = LEFTNet(**ref_cfg(cfg))
        model.load_state_dict(sd, strict=True)
        model = model.to(dt)
        with torch.no_grad():
            ho, po, _ = model(h.to(dt), pos.to(dt), edge_index, subgraph_mask=sub)
        out[f"h_out_{tag}"] = ho.numpy()
        out[f"dpos_{tag}"] = (po - pos.to(dt)).numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), h=h.numpy(), pos=pos.numpy(),
                        edge_index=edge_index.numpy(), subgraph_mask=sub.numpy(), seed=np.int64(seed),
                        cfg=json.dumps(cfg), **out)
    print(name, "ok")


def build_ddpm(cfg, seed, T, dtype=torch.float32):
    dyn, sd = build_dynamics(cfg, [9, 9, 9], 1, seed, dtype)
    sched = DiffSchedule(PredefinedNoiseSchedule("polynomial_2", T, 1e-5), norm_values=(1.0, 1.0, 1.0))
    ddpm = EnVariationalDiffusion(dynamics=dyn, schdule=sched, normalizer=Normalizer(), size_histogram=None,
                                  loss_type="l2", pos_only=True, fixed_idx=None)
    return ddpm, sd


def case_sample(name, cfg, sizes, seed, T):
    ddpm, _ = build_ddpm(cfg, seed, T)
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, seed)
    torch.manual_seed(seed)
    out, masks = ddpm.sample(n_samples=len(sizes), fragments_nodes=nodes, conditions=cond, return_frames=1,
                             timesteps=None, h0=h0)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), sizes=np.array(sizes), seed=np.int64(seed), T=np.int64(T),
                        cfg=json.dumps(cfg), gamma=ddpm.schedule.gamma_module.gamma.detach().numpy(),
                        **{f"out{f}": o.numpy() for f, o in enumerate(out[0])})
    print(name, "ok", [tuple(o.shape) for o in out[0]])


def case_inpaint(name, cfg, sizes, seed, T, resamplings, jump_length):
    ddpm, _ = build_ddpm(cfg, seed, T)
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, seed)
    g = torch.Generator().manual_seed(seed + 1)
    xh_fixed = [torch.cat([torch.randn(h.size(0), 3, generator=g) * 1.5, h], dim=1) for h in h0]
    torch.manual_seed(seed)
    out, masks = ddpm.inpaint(n_samples=len(sizes), fragments_nodes=nodes, conditions=cond, return_frames=1,
                              resamplings=resamplings, jump_length=jump_length, timesteps=None,
                              xh_fixed=[x.clone() for x in xh_fixed], frag_fixed=[0, 2])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), sizes=np.array(sizes), seed=np.int64(seed), T=np.int64(T),
                        resamplings=np.int64(resamplings), jump_length=np.int64(jump_length), cfg=json.dumps(cfg),
                        **{f"xh_fixed{f}": x.numpy() for f, x in enumerate(xh_fixed)},
                        **{f"out{f}": o.numpy() for f, o in enumerate(out[0])})
    print(name, "ok")


def case_train_loss(name, cfg, sizes, seed, T, training):
    """Loss terms of EnVariationalDiffusion.forward (en_diffusion.py:56-248) on one synthetic collate_fn-shaped batch
    (dataset/base_dataset.py:54-88 keys: size, pos, one_hot, charge, mask), from the reference in fp64 (the parity target,
    suffix-less keys) and in fp32 (keys *_f32: the reference's own precision gap).  The random draws of the reference call
    (t_int and every noise sample; identical in both runs) are recorded so that the CUDA path can be fed the same ones."""
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, seed)
    g = torch.Generator().manual_seed(seed + 7)
    reps = []
    for f in range(3):
        m = get_mask_for_frag(nodes[f])
        pos = torch.randn(h0[f].size(0), 3, generator=g) * 1.5
        cnt = torch.zeros(len(sizes)).index_add_(0, m, torch.ones(len(m)))
        pos = pos - (torch.zeros(len(sizes), 3).index_add_(0, m, pos) / cnt[:, None])[m]  # centred per sample (base_dataset.py:215-218)
        reps.append({"size": nodes[f].clone(), "pos": pos, "one_hot": h0[f][:, :5].clone(), "charge": h0[f][:, 5:].clone(), "mask": m})
    out = {f"{k}{f}": r[k].numpy().copy() for f, r in enumerate(reps) for k in ("pos", "one_hot", "charge")}
    all_draws = {}
    for dt, tag in ((torch.float64, ""), (torch.float32, "_f32")):
        ddpm, _ = build_ddpm(cfg, seed, T, dtype=dt)
        ddpm.train(training)
        draws = []
        orig = ddpm.sample_combined_position_feature_noise

        def rec(masks, orig=orig, draws=draws):
            o = orig(masks)
            draws.append([x.clone() for x in o])
            return o
        ddpm.sample_combined_position_feature_noise = rec
        torch.manual_seed(seed)
        with torch.no_grad():
            lt = ddpm.forward([{k: (v.to(dt) if v.is_floating_point() else v.clone()) for k, v in r.items()} for r in reps], cond.to(dt))
        for k in ("error_t", "loss_0_x", "loss_0_cat", "loss_0_charge", "net_eps_xh", "eps_xh"):
            for f in range(3):
                out[f"{k}{f}{tag}"] = lt[k][f].numpy()
        for k in ("SNR_weight", "neg_log_constants", "kl_prior", "t_int"):
            out[f"{k}{tag}"] = lt[k].numpy()
        out[f"delta_log_px{tag}"] = np.float64(lt["delta_log_px"])
        all_draws[tag] = draws
    for d, dr in enumerate(all_draws[""]):
        for f in range(3):
            assert torch.equal(dr[f], all_draws["_f32"][d][f])
            out[f"noise{d}_{f}"] = dr[f].numpy()
    assert np.array_equal(out["t_int"], out["t_int_f32"])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), sizes=np.array(sizes), seed=np.int64(seed), T=np.int64(T),
                        training=np.int64(training), n_draws=np.int64(len(all_draws[""])), cond=cond.numpy(), cfg=json.dumps(cfg), **out)
    print(name, "ok t_int", out["t_int"].tolist(), "error_t0", out["error_t0"].tolist(), "f32", out["error_t0_f32"].tolist())


def case_train_grad(name, cfg, sizes, seed, T, store_full):
    """Gradients of the l2 training objective (DDPMModule.compute_loss with loss_type "l2", trainer/pl_trainer.py:208-282,
    scales [1, 2, 1] of train_ts1x.py:111, mean over the batch) w.r.t. every parameter of the dynamics, from the
    UNMODIFIED reference's autograd in fp64 (and fp32: the reference's own gap).  Inputs and random draws as in
    case_train_loss.  store_full: the gradients themselves; otherwise (large configs) per-parameter L2 norms and the
    projection on a seeded random direction (a checksum that a wrong backward cannot reproduce by accident)."""
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, seed)
    g = torch.Generator().manual_seed(seed + 7)
    reps = []
    for f in range(3):
        m = get_mask_for_frag(nodes[f])
        pos = torch.randn(h0[f].size(0), 3, generator=g) * 1.5
        cnt = torch.zeros(len(sizes)).index_add_(0, m, torch.ones(len(m)))
        pos = pos - (torch.zeros(len(sizes), 3).index_add_(0, m, pos) / cnt[:, None])[m]
        reps.append({"size": nodes[f].clone(), "pos": pos, "one_hot": h0[f][:, :5].clone(), "charge": h0[f][:, 5:].clone(), "mask": m})
    out = {f"{k}{f}": r[k].numpy().copy() for f, r in enumerate(reps) for k in ("pos", "one_hot", "charge")}
    scales = (1.0, 2.0, 1.0)
    all_draws = {}
    for dt, tag in ((torch.float64, ""), (torch.float32, "_f32")):
        ddpm, _ = build_ddpm(cfg, seed, T, dtype=dt)
        ddpm.train(True)
        draws = []
        orig = ddpm.sample_combined_position_feature_noise

        def rec(masks, orig=orig, draws=draws):
            o = orig(masks)
            draws.append([x.clone() for x in o])
            return o
        ddpm.sample_combined_position_feature_noise = rec
        torch.manual_seed(seed)
        lt = ddpm.forward([{k: (v.to(dt) if v.is_floating_point() else v.clone()) for k, v in r.items()} for r in reps], cond.to(dt))
        sz = [r["size"].to(dt) for r in reps]
        loss_t = sum(lt["error_t"][f] / (3 * sz[f]) * scales[f] for f in range(3))
        loss_0 = (sum(lt["loss_0_x"][f] * scales[f] / (3 * sz[f]) for f in range(3)) + sum(lt["loss_0_cat"][f] for f in range(3))
                  + sum(lt["loss_0_charge"][f] for f in range(3)))
        loss = (loss_t + loss_0 + lt["kl_prior"]).mean()
        ddpm.zero_grad()
        loss.backward()
        out[f"loss{tag}"] = np.float64(loss.detach())
        out[f"t_int{tag}"] = lt["t_int"].detach().numpy()
        gen = torch.Generator().manual_seed(seed + 99)
        names = []
        for pn, prm in ddpm.dynamics.named_parameters():
            gr = prm.grad if prm.grad is not None else torch.zeros_like(prm)
            names.append(pn)
            direction = torch.randn(prm.shape, generator=gen, dtype=torch.float64)
            out[f"gnorm{tag}/{pn}"] = np.float64(gr.double().norm())
            out[f"gproj{tag}/{pn}"] = np.float64((gr.double() * direction).sum())
            if store_full:
                out[f"grad{tag}/{pn}"] = gr.detach().numpy()
        all_draws[tag] = draws
    for d, dr in enumerate(all_draws[""]):
        for f in range(3):
            assert torch.equal(dr[f], all_draws["_f32"][d][f])
            out[f"noise{d}_{f}"] = dr[f].numpy()
    assert np.array_equal(out["t_int"], out["t_int_f32"])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), sizes=np.array(sizes), seed=np.int64(seed), T=np.int64(T),
                        n_draws=np.int64(len(all_draws[""])), cond=cond.numpy(), cfg=json.dumps(cfg), scales=np.array(scales),
                        param_names=json.dumps(names), store_full=np.int64(store_full), **out)
    print(name, "ok loss", float(out["loss"]), "f32", float(out["loss_f32"]), "params", len(names))


def case_checkpoint(name="checkpoint_small"):
    """State dict of the unmodified reference's EnVariationalDiffusion in the layout of a DDPMModule Lightning checkpoint
    (`ddpm.` prefix, trainer/pl_trainer.py:77-140) plus keys a real checkpoint carries outside the diffusion model."""
    ddpm, _ = build_ddpm(SMALL_CFG, 71, 20)
    out = {"ddpm." + k: v.detach().numpy() for k, v in ddpm.state_dict().items()}
    out["confidence.head.weight"] = np.ones((2, 3), dtype=np.float32)  # e.g. an auxiliary module of the LightningModule
    np.savez_compressed(os.path.join(OUT, name + ".npz"), cfg=json.dumps(SMALL_CFG), T=np.int64(20), **out)
    print(name, "ok", len(out), "tensors")


def case_api_helpers(name="api_helpers"):
    """Outputs of the unmodified reference's module-level helpers a caller may import next to the samplers: the alpha^2
    schedules (_schedule.py:9-74), the gamma tables of every schedule family, diffusion/_utils.py, get_inner_edge_index
    and EnVariationalDiffusion.gaussian_KL."""
    from oa_reactdiff.diffusion import _schedule as S
    from oa_reactdiff.diffusion import _utils as U
    from oa_reactdiff.utils._graph_tools import get_inner_edge_index
    out = {}
    for T in (10, 100, 1000):
        out[f"poly2_{T}"] = S.polynomial_schedule(T, s=1e-5, power=2.0)
        out[f"poly3_{T}"] = S.polynomial_schedule(T)
        out[f"cos_{T}"] = S.cosine_beta_schedule(T)
        out[f"cos2_{T}"] = S.cosine_beta_schedule(T, raise_to_power=2.0)
        out[f"ccos_{T}"] = S.ccosine_schedule(T, start=0.1, end=0.9, tau=1.5)
        out[f"lin_{T}"] = S.linear_schedule(T)
        for fam in ("polynomial_2", "cosine", "cosine_2", "csin_0.1_0.9_2", "linear"):
            out[f"gamma_{fam}_{T}"] = S.PredefinedNoiseSchedule(fam, T, 1e-5).gamma.detach().numpy()
    rng = np.random.RandomState(5)
    a2 = np.sort(rng.rand(50))[::-1].copy()
    out["clip_in"], out["clip_out"] = a2, S.clip_noise_schedule(a2, clip_value=0.2)
    idx = torch.tensor([0, 0, 2, 2, 2, 3, 5])  # samples 1 and 4 are empty
    x = torch.from_numpy(rng.randn(7, 3)).float()
    out["u_idx"], out["u_x"] = idx.numpy(), x.numpy()
    out["u_remove_mean"] = U.remove_mean_batch(x, idx).numpy()
    out["u_sum_except_batch"] = U.sum_except_batch(x, idx, dim_size=7).numpy()
    out["u_cdf"] = U.cdf_standard_gaussian(x).numpy()
    out["u_batch_mask"] = U.num_nodes_to_batch_mask(4, torch.tensor([2, 0, 3, 1]), torch.device("cpu")).numpy()
    out["u_batch_mask_int"] = U.num_nodes_to_batch_mask(3, 2, torch.device("cpu")).numpy()
    torch.manual_seed(123)
    out["u_cog_noise"] = U.sample_center_gravity_zero_gaussian_batch([7, 3], [idx[:4], idx[4:]]).numpy()
    torch.manual_seed(124)
    out["u_gauss"] = U.sample_gaussian((5, 2), torch.device("cpu")).numpy()
    m = torch.tensor([[0, 1, 0], [1, 1, 0]])
    out["inner_in"], out["inner_out"] = m.numpy(), get_inner_edge_index(m).numpy()
    q = torch.from_numpy(rng.rand(6)).float()
    out["kl_in"] = q.numpy()
    out["kl_out"] = EnVariationalDiffusion.gaussian_KL(q, q + 0.5, 2 * q + 0.1, 3.0).numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "ok", len(out), "arrays")


def synthetic_raw_dataset(seed=7, n=9):
    """A raw Transition1x-style dict (the schema transition1x.py:46-85 reads) with ragged sizes, an excluded multi-fragment
    reaction and a `use_ind` subset."""
    rng = np.random.RandomState(seed)
    sizes = [int(x) for x in rng.randint(3, 9, size=n)]
    raw = {"single_fragment": [1, 1, 0, 1, 1, 1, 1, 0, 1][:n], "use_ind": [0, 1, 3, 5, 6, 8]}
    for k in ("reactant", "transition_state", "product"):
        raw[k] = {"num_atoms": list(sizes), "charges": [], "positions": [], "wB97x_6-31G(d).energy": [float(x) for x in rng.randn(n)]}
    for i, m in enumerate(sizes):
        z = rng.choice([1, 6, 7, 8, 9], size=m).tolist()
        for k in ("reactant", "transition_state", "product"):
            raw[k]["charges"].append(list(z))
            raw[k]["positions"].append((rng.randn(m, 3) * 1.3 + rng.randn(1, 3)).tolist())
    return raw


def case_dataset(name="dataset_small"):
    """ProcessedTS1x + BaseDataset.collate_fn + sampling_tools of the unmodified reference on a synthetic raw dataset
    (dataset/transition1x.py:21-150, base_dataset.py:54-88, utils/sampling_tools.py:64-149)."""
    import copy
    import tempfile
    from oa_reactdiff.dataset.transition1x import ProcessedTS1x
    from oa_reactdiff.utils.sampling_tools import assemble_sample_inputs, write_tmp_xyz
    raw = synthetic_raw_dataset()
    out = {"raw": json.dumps(raw)}
    variants = {"plain": dict(single_frag_only=True, use_by_ind=False, swapping_react_prod=False),
                "swap_useind": dict(single_frag_only=True, use_by_ind=True, swapping_react_prod=True),
                "all_zero": dict(single_frag_only=False, use_by_ind=False, swapping_react_prod=False, zero_charge=True, center=False)}
    out["variants"] = json.dumps(variants)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "raw.pkl")
        for vn, kw in variants.items():
            pickle.dump(copy.deepcopy(raw), open(path, "wb"))
            ds = ProcessedTS1x(path, **kw)
            idxs = [len(ds) - 1, 0, 2] if len(ds) > 2 else list(range(len(ds)))
            reps, cond = ProcessedTS1x.collate_fn([ds[i] for i in idxs])
            out[f"{vn}/len"] = np.int64(len(ds))
            out[f"{vn}/idxs"] = np.array(idxs)
            out[f"{vn}/cond"] = cond.numpy()
            for f, r in enumerate(reps):
                for k, v in r.items():
                    out[f"{vn}/{k}{f}"] = v.numpy()
            one = ds[1]
            for k, v in one.items():
                out[f"{vn}/item1/{k}"] = v.numpy()
        atoms = ["C", "H", "H", "O", "N", "F"]
        for ft in (False, True):
            h0 = assemble_sample_inputs(atoms, device=torch.device("cpu"), n_samples=2, frag_type=ft)
            for f in range(3):
                out[f"h0_ft{int(ft)}_{f}"] = h0[f].numpy()
        g = torch.Generator().manual_seed(3)
        nodes = [torch.tensor([2, 4])] * 3
        samples = [torch.cat([torch.randn(6, 3, generator=g), torch.zeros(6, 5), torch.tensor([[1.], [6.], [7.], [8.], [9.], [1.]])], dim=1)
                   for _ in range(3)]
        write_tmp_xyz(nodes, samples, idx=[0, 1, 2], prefix="gen", localpath=td, ex_ind=3)
        for fn in sorted(os.listdir(td)):
            if fn.endswith(".xyz"):
                out["xyz/" + fn] = open(os.path.join(td, fn)).read()
        for f in range(3):
            out[f"xyz_in{f}"] = samples[f].numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), cfg=json.dumps({}), **out)
    print(name, "ok", [k for k in out if k.endswith("/len")], [int(out[k]) for k in out if k.endswith("/len")])


def _reference_evaluate_utils():
    """oa_reactdiff/evaluate/utils.py, unmodified.  Its module-level imports of the Lightning trainer and the pyscf-based
    geometry tools (absent here, and unused by the three helpers wanted) are satisfied by empty stand-in modules."""
    import types
    for name, attrs in (("oa_reactdiff.trainer.pl_trainer", {"DDPMModule": object}),
                        ("oa_reactdiff.analyze.geomopt", {"calc_deltaE": None, "compute_efh": None})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
            pkg = name.rsplit(".", 1)[0]
            if pkg not in sys.modules:
                sys.modules[pkg] = types.ModuleType(pkg)
                sys.modules[pkg].__path__ = []
    import importlib
    return importlib.import_module("oa_reactdiff.evaluate.utils")


def case_eval_pipeline(name="eval_pipeline_small", seed=81):
    """The evaluation-side callers of the path, end to end with the unmodified reference: ProcessedTS1x batch (dataset/
    transition1x.py) -> set_new_schedule -> inplaint_batch -> samples_to_pos_charge (evaluate/utils.py:14-63, 91-110)."""
    import copy
    import tempfile
    from oa_reactdiff.dataset.transition1x import ProcessedTS1x
    EU = _reference_evaluate_utils()
    raw = synthetic_raw_dataset()
    kw = dict(single_frag_only=True, use_by_ind=False, swapping_react_prod=False)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "raw.pkl")
        pickle.dump(copy.deepcopy(raw), open(path, "wb"))
        ds = ProcessedTS1x(path, **kw)
        idxs = [4, 0, 2]
        batch = ProcessedTS1x.collate_fn([ds[i] for i in idxs])
    ddpm, _ = build_ddpm(SMALL_CFG, seed, 20)

    class Trainer:  # what the helpers need of the LightningModule: `.ddpm` and `.to()`
        def __init__(self, d):
            self.ddpm = d

        def to(self, device):
            self.ddpm.to(device)
            return self

    tr = EU.set_new_schedule(Trainer(ddpm), timesteps=8, device=torch.device("cpu"), noise_schedule="polynomial_3")
    torch.manual_seed(seed)
    out, xh_fixed, fragments_nodes = EU.inplaint_batch(batch, tr, resamplings=2, jump_length=2, frag_fixed=[0, 2])
    pos, z, natoms = EU.samples_to_pos_charge(out, fragments_nodes)
    res = {"raw": json.dumps(raw), "kw": json.dumps(kw), "idxs": np.array(idxs), "seed": np.int64(seed), "T0": np.int64(20),
           "gamma_new": tr.ddpm.schedule.gamma_module.gamma.detach().numpy(), "T_new": np.int64(tr.ddpm.T),
           "natoms": np.array(natoms)}
    for f in range(3):
        res[f"out{f}"] = out[f].numpy()
        res[f"xh_fixed{f}"] = xh_fixed[f].numpy()
    for k, v in pos.items():
        for i, a in enumerate(v):
            res[f"pos/{k}/{i}"] = a
    for i, a in enumerate(z):
        res[f"z/{i}"] = a
    np.savez_compressed(os.path.join(OUT, name + ".npz"), cfg=json.dumps(SMALL_CFG), **res)
    print(name, "ok", natoms)


def t1x_histogram():
    path = "/root/reference/oa_reactdiff/data/transition1x/train.pkl"
    with open(path, "rb") as fh:
        data = pickle.load(fh)
    n_atoms = np.array([len(x) for x in data["transition_state"]["charges"]])
    use = np.array(data["use_ind"]) if "use_ind" in data else np.arange(len(n_atoms))
    n_use = n_atoms[use]
    hist = np.bincount(n_use, minlength=24).tolist()
    charges = np.concatenate([np.asarray(data["transition_state"]["charges"][k]) for k in use[:2000]])
    zs, cnt = np.unique(charges, return_counts=True)
    with open(os.path.join(OUT, "t1x_hist.json"), "w") as fh:
        json.dump({"source": "oa_reactdiff/data/transition1x/train.pkl use_ind", "n_reactions": int(len(n_use)),
                   "min": int(n_use.min()), "max": int(n_use.max()), "mean": float(n_use.mean()),
                   "hist_by_natoms": hist,
                   "element_freq": {str(int(z)): int(c) for z, c in zip(zs, cnt)}}, fh, indent=1)
    print("t1x_hist", len(n_use), n_use.min(), n_use.max(), n_use.mean())


if __name__ == "__main__":
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "train_grad":  # only the gradient fixtures
        case_train_grad("grad_small_train", SMALL_CFG, [5, 3, 4], seed=61, T=20, store_full=True)
        case_train_grad("grad_trained_train_b3", TRAINED_CFG, [4, 9, 6], seed=62, T=100, store_full=False)
        case_train_grad("grad_small_train_t0", SMALL_CFG, [4, 5, 3], seed=77, T=20, store_full=True)  # one sample drawn at t = 0: L0 terms
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "checkpoint":
        case_checkpoint()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "inpaint_trained":  # only the trained-configuration RePaint trajectory
        case_inpaint("inpaint_trained_b4_T12_r2_j3", TRAINED_CFG, [5, 9, 14, 7], seed=44, T=12, resamplings=2, jump_length=3)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "noreflect":  # reflect_equiv=False (reference tests/model/test_equiv.py:40-41)
        case_leftnet("leftnet_small_noreflect", dict(SMALL_CFG, reflect_equiv=False), 11, seed=14, cut=5, pos_scale=3.0)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "eval":  # only the evaluation-pipeline fixture
        case_eval_pipeline()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "api":  # only the module-level helper fixture
        case_api_helpers()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "dataset":  # only the dataset / sampling-tools fixture
        case_dataset()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "train_loss":  # only the loss-term fixtures
        case_train_loss("loss_small_train", SMALL_CFG, [5, 3, 4], seed=51, T=20, training=True)
        case_train_loss("loss_small_eval", SMALL_CFG, [5, 3, 4], seed=52, T=20, training=False)
        case_train_loss("loss_trained_train_b4", TRAINED_CFG, [4, 9, 14, 7], seed=53, T=100, training=True)
        sys.exit(0)
    t1x_histogram()
    # raw LEFTNet.forward (reference tests/model style): full graph, and object-aware cut graph
    case_leftnet("leftnet_small_full", SMALL_CFG, 9, seed=11, pos_scale=3.0)
    case_leftnet("leftnet_small_cut", SMALL_CFG, 10, seed=12, cut=4, pos_scale=3.0)
    case_leftnet("leftnet_small_split", SMALL_CFG, 12, seed=13, cut=5, pos_scale=9.0)  # cutoff splits groups
    case_leftnet("leftnet_small_noreflect", dict(SMALL_CFG, reflect_equiv=False), 11, seed=14, cut=5, pos_scale=3.0)
    # dynamics: reference fixture shape (ragged, an EMPTY fragment, per-fragment node_nf) tests/dynamics/test_egnn_dynamics.py:99-104
    case_dynamics("dyn_small_ragged", dict(SMALL_CFG, in_hidden_channels=8),
                  [torch.tensor([2, 0]), torch.tensor([2, 3]), torch.tensor([1, 2])], [4, 5, 6], 3, seed=21,
                  pos_scale=1.5)
    # config 1: trained cfg, single 12-atom triple
    case_dynamics("dyn_trained_cfg1", TRAINED_CFG, [torch.tensor([12])] * 3, [9, 9, 9], 1, seed=31, pos_scale=1.5)
    # trained cfg, ragged B=4 incl. min/max sizes; pos_scale 4 => some pairs beyond the 10 A cutoff
    case_dynamics("dyn_trained_b4", TRAINED_CFG, [torch.tensor([4, 9, 14, 23])] * 3, [9, 9, 9], 1, seed=32,
                  pos_scale=1.5)
    case_dynamics("dyn_trained_b3_far", TRAINED_CFG, [torch.tensor([5, 17, 11])] * 3, [9, 9, 9], 1, seed=33,
                  pos_scale=4.0)
    # sampler trajectories (fp32, reference RNG order)
    case_sample("sample_small_T10", SMALL_CFG, [5, 3], seed=41, T=10)
    case_sample("sample_trained_cfg1_T10", TRAINED_CFG, [12], seed=42, T=10)
    case_inpaint("inpaint_small_T12_r2_j3", SMALL_CFG, [4, 6], seed=43, T=12, resamplings=2, jump_length=3)
    case_inpaint("inpaint_trained_b4_T12_r2_j3", TRAINED_CFG, [5, 9, 14, 7], seed=44, T=12, resamplings=2, jump_length=3)
    case_train_loss("loss_small_train", SMALL_CFG, [5, 3, 4], seed=51, T=20, training=True)
    case_train_loss("loss_small_eval", SMALL_CFG, [5, 3, 4], seed=52, T=20, training=False)
    case_train_loss("loss_trained_train_b4", TRAINED_CFG, [4, 9, 14, 7], seed=53, T=100, training=True)
    case_dataset()
    case_train_grad("grad_small_train", SMALL_CFG, [5, 3, 4], seed=61, T=20, store_full=True)
    case_train_grad("grad_trained_train_b3", TRAINED_CFG, [4, 9, 6], seed=62, T=100, store_full=False)
    case_train_grad("grad_small_train_t0", SMALL_CFG, [4, 5, 3], seed=77, T=20, store_full=True)
    case_checkpoint()
    case_api_helpers()
    case_eval_pipeline()
