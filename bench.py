#!/usr/bin/env python
"""bench.py — reactions/sec for a full 1000-step reverse diffusion on Transition1x-shaped batches (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on the host cores (oracle port)

One "step" = one full reverse diffusion of a B=64 batch: 1000 reverse steps (denoiser evaluation + posterior sampling)
+ the final p(x|z0) decode = 1001 LEFTNet evaluations, on the states a TRAINED model visits: z_t = alpha_t x + sigma_t eps
re-drawn every step from frozen real Transition1x reactant / TS / product geometries (workloads.replay_trajectory; active
edge fraction ~0.32).  The literal `sample()` with random-init weights is reported beside it (`literal_sample`): there the
positions drift out of the 10 A cutoff and most of the message-passing work is (exactly) skipped, which flatters the number.
Weak scaling: every rank runs its own B=64 batch.  `value` = reactions per second with inputs resident in HBM; `e2e` = the
same with HOST (pinned) inputs copied in and results copied out inside the timed region.  Further legs in the same JSON
line (1 GPU only): `inpaint` (BASELINE config 4: RePaint r=1/j=1 and r=5/j=5) and `train_step` (config 5: B=128 forward +
backward with loss / gradient errors against the oracle and its CPU timing).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reactions/sec full 1000-step reverse diffusion, Transition1x-shaped batch"
UNIT = "reactions/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="reactions per GPU")
    ap.add_argument("--denoise-steps", type=int, default=1000, help="T of the reverse diffusion (BASELINE: 1000)")
    ap.add_argument("--profile-every", type=int, default=29, help="untimed profiling pass: bracket kernels with CUDA events every n-th forward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--geometry", default="real", choices=["real", "synthetic"],
                    help="reference geometries of the trained-model states: frozen REAL Transition1x R/TS/P coordinates "
                         "(tests/golden/t1x_geometries_b512.npz, oracle/gen_t1x_geometries.py; default) or compact synthetic clouds")
    ap.add_argument("--no-literal", action="store_true", help="skip the literal random-weight sample() line")
    ap.add_argument("--no-inpaint", action="store_true", help="skip the TS-inpainting legs (BASELINE config 4)")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg (BASELINE config 5)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="end-to-end steps (default: min(steps, 5), at least 3)")
    ap.add_argument("--train-batch", type=int, default=128)
    ap.add_argument("--train-steps", type=int, default=3)
    ap.add_argument("--train-check", type=int, default=4, help="reactions of the training batch checked against / timed on the oracle")
    ap.add_argument("--cpu-evals", type=int, default=2, help="denoiser evaluations timed for cpu_baseline")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm_gbs=6650.0, tflops=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        pw = []
        for r in self.rows:
            try:
                pw.append(float(r[2]))
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_mhz_min": sm[0] if sm else None, "sm_mhz_p10": sm[len(sm) // 10] if sm else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def load_workloads():
    """oareactdiff_b200/workloads.py by file path: the reference arm must not import the package (which loads the CUDA library)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_oard_workloads", os.path.join(ROOT, "oareactdiff_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def tune_cpu_threads():
    """Eager PyTorch on many-core hosts is often fastest well below the core count (oversubscription of small ops):
    time one small denoiser evaluation at a few thread counts and keep the best, so the CPU baseline is a fair one."""
    import torch
    from oracle import oa_ref
    workloads = load_workloads()
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores})
    cfg = dict(oa_ref.TRAINED_CFG)
    sizes = workloads.t1x_sizes(8, seed=1)
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, 1)
    sd = oa_ref.make_state_dict(oa_ref.dynamics_param_shapes(cfg, [9, 9, 9], 1), 0, cfg, prefix_model="model.")
    smp = oa_ref.Sampler(sd, cfg, oa_ref.gamma_table("polynomial_2", 10, 1e-5))
    masks, cm, ei, nfs = smp._graph(nodes)
    z = [torch.cat([torch.randn(h.size(0), 3), h], dim=1) for h in h0]
    t = torch.full((len(sizes), 1), 0.5)
    best, best_t = cands[0], float("inf")
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            smp._dyn(z, ei, t, cond, nfs, masks)  # warm
            t0 = time.perf_counter()
            smp._dyn(z, ei, t, cond, nfs, masks)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = c, dt
    torch.set_num_threads(best)
    return best


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference algorithm on the host cores: oracle/oa_ref.py (a torch-CPU restatement with the reference's op
    structure; /root/reference itself cannot travel to the GPU box).  One step = ONE reverse-diffusion step (one
    denoiser evaluation + posterior sampling) on the full batch, extrapolated x1001 to a full trajectory: the cost per
    step is constant (dense masked compute over all edges)."""
    if rank != 0:
        return
    import torch
    from oracle import oa_ref
    workloads = load_workloads()
    tune_cpu_threads()
    cfg = dict(oa_ref.TRAINED_CFG)
    T = args.denoise_steps
    sizes = workloads.t1x_sizes(args.batch, seed=0)
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, 0)
    sd = oa_ref.make_state_dict(oa_ref.dynamics_param_shapes(cfg, [9, 9, 9], 1), 0, cfg, prefix_model="model.")
    smp = oa_ref.Sampler(sd, cfg, oa_ref.gamma_table("polynomial_2", T, 1e-5))
    masks, cm, ei, nfs = smp._graph(nodes)
    torch.manual_seed(0)
    z = smp._noise(masks)
    z = [torch.cat([z[i][:, :3], h0[i]], dim=1) for i in range(3)]
    times = []
    with torch.no_grad():
        for it in range(args.warmup + args.steps):
            s = T - 1 - it
            s_arr = torch.full((len(sizes), 1), float(s))
            t0 = time.perf_counter()
            z = smp._p_zs_given_zt(s_arr / T, (s_arr + 1) / T, z, ei, nfs, masks, cond)
            z = [torch.cat([z[i][:, :3], h0[i]], dim=1) for i in range(3)]
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    per_eval = sum(times) / len(times)
    traj = per_eval * (T + 1)
    val = len(sizes) / traj
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": traj * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"batch={args.batch} Transition1x-shaped reactions (<=23 atoms), {T} steps, CPU",
                       "global_batch": args.batch, "denoise_steps": T, "nodes": int(cm.numel()), "edges": int(ei.size(1))},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "host_cores": os.cpu_count(), "kind": "port",
                             "sample": f"{args.steps} reverse steps (1 denoiser evaluation each, {per_eval:.2f} s/eval) on the "
                                       f"full batch, extrapolated x{T + 1}; torch {torch.__version__} CPU fp32"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------- B200 arm
def cpu_baseline(args, T, workloads):
    import torch
    from oracle import oa_ref
    tune_cpu_threads()
    cfg = dict(oa_ref.TRAINED_CFG)
    sizes = workloads.t1x_sizes(args.batch, seed=0)
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, 0)
    sd = oa_ref.make_state_dict(oa_ref.dynamics_param_shapes(cfg, [9, 9, 9], 1), 0, cfg, prefix_model="model.")
    smp = oa_ref.Sampler(sd, cfg, oa_ref.gamma_table("polynomial_2", T, 1e-5))
    torch.manual_seed(0)
    t0 = time.perf_counter()
    smp.sample(len(sizes), nodes, cond, h0, max_steps=max(args.cpu_evals - 1, 1))
    dt = time.perf_counter() - t0
    per_eval = dt / smp.n_evals
    val = len(sizes) / (per_eval * (T + 1))
    return {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "host_cores": os.cpu_count(), "kind": "port",
            "sample": f"{smp.n_evals} denoiser evaluations of the same B={args.batch} batch via oracle Sampler.sample "
                      f"({per_eval:.2f} s/eval; the reference's cost per evaluation does not depend on the geometry: dense "
                      f"masked compute over all edges), extrapolated to {T + 1}"}


def train_leg(args, ob, workloads, dev, cfg, T):
    """BASELINE config 5: one training step (l2 objective: EnVariationalDiffusion.compute_loss -> mean -> backward through
    the CUDA denoiser) on B = 128 noised real-geometry reaction triples; plus, on the first `train_check` reactions, the same
    step against the oracle's fp32 forward + torch autograd on the host cores (timed: the CPU baseline of this leg) and its
    fp64 run (the parity figure: loss and worst parameter-gradient error relative to max|grad|)."""
    import torch
    from oracle import oa_ref  # checker + CPU baseline of this leg

    class ReplayDraws(ob.EnVariationalDiffusion):  # replays given draws so that both sides see the same t and noise
        def set_draws(self, t_int, noises):
            self._t_int, self._noises, self._k = t_int, noises, 0

        def _draw_t_int(self, num_sample, device):
            return self._t_int.to(device).view(num_sample, 1)

        def sample_combined_position_feature_noise(self, masks):
            out = [n.to(masks[0].device) for n in self._noises[self._k]]
            self._k += 1
            return out

    Bt = args.train_batch
    sizes = workloads.t1x_sizes(Bt, seed=0)
    x_ref = workloads.real_geometries(0, Bt, sizes)
    _, h0, _ = workloads.reaction_batch(sizes, seed=0)
    torch.manual_seed(0)
    dyn = ob.EGNNDynamics(model_config=cfg, fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0,
                          condition_nf=1, model=ob.LEFTNetB200, device=dev).to(dev)
    dyn.model.enable_training_path = True
    sched = ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", T, 1e-5), norm_values=(1.0, 1.0, 1.0))
    ddpm = ReplayDraws(dynamics=dyn, schdule=sched, normalizer=ob.Normalizer(), pos_only=True).to(dev)
    ddpm.train(True)
    gen = torch.Generator().manual_seed(7)

    def draws(szs):
        t_int = torch.randint(1, T + 1, (len(szs), 1), generator=gen).float()
        mask = torch.repeat_interleave(torch.arange(len(szs)), torch.tensor(szs))
        noise = []
        for _ in range(3):
            x = torch.randn(sum(szs), 3, generator=gen)
            mean = torch.zeros(len(szs), 3).index_add_(0, mask, x) / torch.tensor(szs, dtype=torch.float32)[:, None]
            noise.append(torch.cat([x - mean[mask], torch.zeros(sum(szs), 6)], dim=1))
        return t_int, noise, mask

    def step(szs, xr, hh, t_int, noise):
        reps, cond = workloads.training_batch(szs, xr, hh, dev)
        ddpm.set_draws(t_int, [noise])
        for prm in dyn.parameters():
            prm.grad = None
        nll, _ = ddpm.compute_loss((reps, cond), scales=(1.0, 2.0, 1.0), training=True)
        loss = nll.mean()
        loss.backward()
        return loss

    t_int, noise, _ = draws(sizes)
    step(sizes, x_ref, h0, t_int, noise)  # warm-up (allocations, first-call attributes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(args.train_steps):
        e0.record()
        loss = step(sizes, x_ref, h0, t_int, noise)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    out = {"workload": f"batch={Bt} noised Transition1x reaction triples (real R/TS/P geometries), l2 objective, forward + backward "
                       f"through encoders / LEFTNet / decoders (exact-fp32 SIMT training kernels, csrc/train_core.h)",
           "ms_per_step": times[len(times) // 2], "ms_min": times[0], "ms_max": times[-1], "steps": len(times),
           "reactions_per_s": Bt / (times[len(times) // 2] * 1e-3), "loss": float(loss.detach()),
           "grads_finite": all(bool(torch.isfinite(p_.grad).all()) for p_ in dyn.parameters() if p_.grad is not None)}
    # ---- parity + CPU baseline on a sub-batch
    k = args.train_check
    if k > 0:
        sub_sizes = sizes[:k]
        n_sub = sum(sub_sizes)
        xr = [x[:n_sub] for x in x_ref]
        hh = [h[:n_sub] for h in h0]
        t_sub, noise_sub, mask_sub = draws(sub_sizes)
        loss_gpu = float(step(sub_sizes, xr, hh, t_sub, noise_sub).detach())
        grads = {n_: p_.grad.detach().cpu().double() for n_, p_ in dyn.named_parameters() if p_.grad is not None}
        sd = {n_: p_.detach().cpu() for n_, p_ in dyn.state_dict().items()}
        gamma = sched.gamma_module.gamma.detach().cpu()
        xh = [torch.cat([x - (torch.zeros(k, 3).index_add_(0, mask_sub, x) / torch.tensor(sub_sizes, dtype=torch.float32)[:, None])[mask_sub], h], dim=1)
              for x, h in zip(xr, hh)]
        res = {}
        for dt_, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            sd_ = {n_: v.detach().to(dt_).clone().requires_grad_(True) if v.is_floating_point() else v for n_, v in sd.items()}
            tune_cpu_threads()
            t0 = time.perf_counter()
            l_ = oa_ref.train_loss_l2(sd_, cfg, gamma.to(dt_), [x.to(dt_) for x in xh], [mask_sub] * 3, torch.tensor(sub_sizes),
                                      torch.zeros(k, 1), t_sub.view(-1), [n_.to(dt_) for n_ in noise_sub])
            l_.backward()
            res[tag] = (float(l_.detach()), {n_: v.grad.double() for n_, v in sd_.items() if getattr(v, "grad", None) is not None},
                        time.perf_counter() - t0)
        l64, g64, _ = res["f64"]
        worst, worst32, n_cmp = 0.0, 0.0, 0
        for n_, gref in g64.items():
            sc = float(gref.abs().max())
            if sc == 0.0 or n_ not in grads:
                continue
            n_cmp += 1
            worst = max(worst, float((grads[n_] - gref).abs().max()) / sc)
            worst32 = max(worst32, float((res["f32"][1][n_] - gref).abs().max()) / sc)
        out["check"] = {"reactions": k, "loss_rel_err": abs(loss_gpu - l64) / abs(l64), "grad_rel_err_worst": worst,
                        "parameters_compared": n_cmp, "oracle_fp32_grad_rel_err_worst": worst32,
                        "against": "oracle fp64 forward + torch autograd on the same draws (pinned on the unmodified reference's "
                                   "gradient goldens, tests/test_oracle_grad.py)"}
        import torch as _t
        out["cpu_baseline"] = {"value": k / res["f32"][2], "unit": "reactions/s", "cores": _t.get_num_threads(), "kind": "port",
                               "sample": f"oracle fp32 forward + autograd backward of {k} of the {Bt} reactions "
                                         f"({res['f32'][2]:.2f} s), cost linear in the reactions"}
    return out


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import oareactdiff_b200 as ob
    from oareactdiff_b200 import parallel, workloads

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T, B = args.denoise_steps, args.batch
    cfg = dict(cutoff=10.0, num_layers=6, hidden_channels=196, num_radial=96, in_hidden_channels=8, reflect_equiv=True,
               legacy=True, update=True, object_aware=True)  # trainer/train_ts1x.py:43-56
    torch.manual_seed(0)
    dyn = ob.EGNNDynamics(model_config=cfg, fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0,
                          condition_nf=1, model=ob.LEFTNetB200, device=dev).to(dev)
    bcast_bytes = parallel.broadcast_module_(dyn, src=0)  # the one collective of the path: rank 0's weights (42.6 MB), NCCL
    sched = ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", T, 1e-5), norm_values=(1.0, 1.0, 1.0))
    ddpm = ob.EnVariationalDiffusion(dynamics=dyn, schdule=sched, normalizer=ob.Normalizer(), pos_only=True).to(dev)
    dyn.model.assume_static_weights = True

    # weak scaling: a global batch of B*world Transition1x-shaped reactions, sharded into contiguous chunks balanced by
    # edge count; no per-step communication (reactions are independent)
    all_sizes = workloads.t1x_sizes(B * world, seed=0)
    lo, hi = parallel.shard_reactions(all_sizes, world)[rank]
    sizes = all_sizes[lo:hi]
    B = len(sizes)
    nodes_h, h0_h, cond_h = workloads.reaction_batch(sizes, seed=rank)
    geometry = args.geometry
    if geometry == "real" and hi > 512:
        geometry = "synthetic"  # the fixture covers 512 reactions
    x_h = workloads.real_geometries(lo, hi, sizes) if geometry == "real" else workloads.synthetic_geometries(sizes, rank)
    pin = lambda t: t.pin_memory()
    nodes_h, h0_h, cond_h, x_h = [pin(x) for x in nodes_h], [pin(x) for x in h0_h], pin(cond_h), [pin(x) for x in x_h]
    nodes_d, h0_d, cond_d, x_d = [x.to(dev) for x in nodes_h], [x.to(dev) for x in h0_h], cond_h.to(dev), [x.to(dev) for x in x_h]
    eng = dyn.model.engine(dev)
    out_host = [torch.empty(h.size(0), 3).pin_memory() for h in h0_h]

    # ---- headline step: a full reverse diffusion on trained-model states built from the reference geometries
    seg_probe = []  # one list of CUDA events (every 100 reverse steps) per resident trajectory

    def step_resident():
        torch.manual_seed(4321 + rank)
        seg_probe.append([])
        return workloads.replay_trajectory(ddpm, B, nodes_d, cond_d, h0_d, x_d, T, probe=seg_probe[-1])

    def step_e2e():  # host (pinned) inputs -> device inside the timed region, result back to the host
        torch.manual_seed(4321 + rank)
        nd = [x.to(dev, non_blocking=True) for x in nodes_h]
        hd = [x.to(dev, non_blocking=True) for x in h0_h]
        xd = [x.to(dev, non_blocking=True) for x in x_h]
        cd = cond_h.to(dev, non_blocking=True)
        pos = workloads.replay_trajectory(ddpm, B, nd, cd, hd, xd, T)
        for dst, src in zip(out_host, pos):
            dst.copy_(src.to(torch.float32), non_blocking=True)
        return pos

    def step_literal():  # the literal sample() with random weights: the trajectory drifts out of the cutoff
        torch.manual_seed(1234 + rank)
        out, _ = ddpm.sample(B, nodes_d, cond_d, h0=h0_d)
        return out[0]

    h2d = sum(x.numel() * x.element_size() for x in nodes_h + h0_h + x_h + [cond_h])
    d2h = sum(x.numel() * x.element_size() for x in out_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = {}

    def timed(fn, k):
        """-> (total ms of k steps, max over ranks; per-step ms of this rank)."""
        barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
        evs[0].record()
        t0 = time.perf_counter()
        for i in range(k):
            fn()
            evs[i + 1].record()
        host_ms[fn.__name__] = (time.perf_counter() - t0) * 1e3 / k  # host enqueue time (no sync inside the loop)
        barrier()
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(k)]
        return parallel.max_over_ranks(evs[0].elapsed_time(evs[k]), dev), per

    def profile_of(prof):
        return {k: dict(ms_per_launch=v["ms"] / max(v["launches"], 1), launches=v["launches"], ms=v["ms"],
                        tflops=(v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 else 0.0,
                        gbs=(v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else 0.0, flops=v["flops"], bytes=v["bytes"])
                for k, v in prof.items() if not k.startswith("_")}

    for _ in range(args.warmup):
        step_resident()
    l0 = eng.total_launches()
    clocks = ClockSampler(local_rank)
    clocks.start()
    del seg_probe[:]
    ms, per_step = timed(step_resident, args.steps)
    clk = clocks.stop()
    # per trajectory: the ten 100-step segments (ms) -> is a slow trajectory uniformly slow (clocks) or slow in bursts (stalls)?
    segments = [[round(evs[i].elapsed_time(evs[i + 1]), 1) for i in range(len(evs) - 1)] for evs in seg_probe[:args.steps]]
    launches = eng.total_launches() - l0
    # per-kernel CUDA-event timing in a SEPARATE, untimed pass over the same step (a profiled forward runs eagerly with an
    # event pair around every kernel: ~10 % slower, so it must not sit inside the timed region)
    eng.set_profile(args.profile_every)
    step_resident()
    torch.cuda.synchronize()
    prof = eng.profile()
    eng.set_profile(0)
    e2e_steps = args.e2e_steps if args.e2e_steps > 0 else max(3, min(args.steps, 5))
    ms_e2e, per_e2e = timed(step_e2e, e2e_steps)
    finite = all(bool(torch.isfinite(o).all()) for o in out_host)

    literal = None
    if not args.no_literal:
        eng.set_profile(args.profile_every)  # (the profile is reset by set_profile: this pass only measures the active fraction)
        step_literal()
        torch.cuda.synchronize()
        prof_lit = eng.profile()
        eng.set_profile(0)
        ms_lit, _ = timed(step_literal, 1)
        af = prof_lit.get("_active_fraction")
        literal = {"value": len(all_sizes) / (ms_lit / 1e3), "unit": UNIT, "ms_per_step": ms_lit, "steps": 1,
                   "active_edge_fraction": (af["flops"] / max(af["launches"], 1)) if af else None,
                   "note": "EnVariationalDiffusion.sample() taken literally with random-init weights: positions drift to "
                           "|x| ~ 1e2 A, the cutoff empties and the active-edge stages shrink (exactly skipped work)"}

    inpaint = None
    if world == 1 and not args.no_inpaint:
        inpaint = {}
        for r_, j_ in ((1, 1), (5, 5)):
            try:
                def step_inpaint():
                    torch.manual_seed(77 + rank)
                    xf = [torch.cat([x, h], dim=1) for x, h in zip(x_d, h0_d)]
                    out, _ = ddpm.inpaint(B, nodes_d, cond_d, resamplings=r_, jump_length=j_, xh_fixed=xf, frag_fixed=[0, 2])
                    return out[0]
                if (r_, j_) == (1, 1):
                    step_inpaint()
                ms_i, _ = timed(step_inpaint, 1)
                inpaint[f"r{r_}_j{j_}"] = {"value": B / (ms_i / 1e3), "unit": UNIT, "ms_per_step": ms_i, "denoiser_evaluations": ddpm.n_evals,
                                           "ms_per_evaluation": ms_i / ddpm.n_evals, "host_enqueue_ms": host_ms.get("step_inpaint")}
            except Exception as ex:  # noqa: BLE001
                inpaint[f"r{r_}_j{j_}"] = {"error": repr(ex)[:300]}
        inpaint["workload"] = (f"TS inpainting (RePaint): reactant + product clamped to real Transition1x geometries, TS resampled; "
                               f"batch={B}, {T} steps, 1xB200 (BASELINE config 4)")

    train = None
    if world == 1 and not args.no_train:
        try:
            del step_inpaint
        except Exception:  # noqa: BLE001
            pass
        try:
            train = train_leg(args, ob, workloads, dev, cfg, T)
        except Exception as ex:  # noqa: BLE001
            train = {"error": repr(ex)[:400]}

    total_reactions = len(all_sizes) * args.steps
    value = total_reactions / (ms / 1e3)
    e2e_v = len(all_sizes) * e2e_steps / (ms_e2e / 1e3)
    if rank != 0:
        return
    pk = peaks()
    af_rep = prof.pop("_active_fraction", None)
    active_fraction = (af_rep["flops"] / max(af_rep["launches"], 1)) if af_rep else None
    kernels = profile_of(prof)
    tot = sum(v["ms"] for v in kernels.values()) or 1.0
    for v in kernels.values():
        v["share"] = v["ms"] / tot
    dom = max(kernels, key=lambda k: kernels[k]["share"]) if kernels else None
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if dom and os.path.exists(tpath):
        tj = json.load(open(tpath))
        t_ = tj.get(dom)
        traffic = t_.get("dram_bytes_per_launch") if isinstance(t_, dict) else t_  # ncu dram read+write bytes of one launch
        traffic_src = tj.get("_source")
    roof = None
    ridge = pk["tflops"] * 1e12 / (pk["hbm_gbs"] * 1e9)  # FLOP per byte where the two roofs meet (bf16 tensor vs HBM)
    mp = kernels.get("k_equi_msg")
    mp_roof = None if not mp else {
        "kernel": "k_equi_tgt (EquiMessage message + aggregation at the target)", "bound": "hbm", "achieved": mp["gbs"],
        "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": mp["gbs"] / pk["hbm_gbs"], "ms_per_launch": mp["ms_per_launch"],
        "share_of_step": mp["share"], "algorithmic_bytes": "E_act * (3H*4 + 24): G row + geometry / index record per active edge",
        "measured_on": "the timed run (trained-model geometry)"}
    if dom:
        kd = kernels[dom]
        # tensor pipe executes 3 bf16 MMAs per algorithmic fp32 product (bf16x3 split): that is what competes with HBM
        intensity = (3.0 * kd["flops"] / kd["bytes"]) if kd["bytes"] > 0 else float("inf")
        common = {"kernel": dom, "traffic": traffic, "traffic_source": traffic_src, "share_of_step": kd["share"],
                  "ms_per_launch": kd["ms_per_launch"], "active_edge_fraction": active_fraction, "message_passing": mp_roof}
        if dom.startswith("gemm") and intensity >= ridge:
            roof = {"bound": "tensor", "achieved": kd["tflops"], "peak": pk["tflops"], "unit": "TFLOP/s",
                    "frac": kd["tflops"] / pk["tflops"], **common, "mma_issued_tflops": 3.0 * kd["tflops"],
                    "note": f"achieved = algorithmic fp32-equivalent flops (2MNK, inactive edges skipped) / CUDA-event launch time; "
                            f"the kernel issues 3 bf16 MMAs per product (bf16x3 split), so its own ceiling is peak/3; "
                            f"peak = bf16 cuBLAS sustained ({pk['src']})"}
        else:
            roof = {"bound": "hbm", "achieved": kd["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": kd["gbs"] / pk["hbm_gbs"], **common,
                    "intensity_flop_per_byte": intensity, "ridge_flop_per_byte": ridge,
                    "note": f"achieved = algorithmic bytes (operand read + residual read + result write, fp32) / CUDA-event "
                            f"launch time; arithmetic intensity (bf16 MMA flops per byte) is below the ridge, so HBM is the "
                            f"bound; peak = copy bandwidth ({pk['src']})"}
    gemm_tab = {k: {"tflops_alg": round(v["tflops"], 1), "frac_of_bf16x3_ceiling": round(3.0 * v["tflops"] / pk["tflops"], 3),
                    "gbs_alg": round(v["gbs"], 1), "frac_hbm": round(v["gbs"] / pk["hbm_gbs"], 3)}
                for k, v in kernels.items() if k.startswith("gemm")}
    srt = sorted(per_step)
    geo_txt = ("frozen REAL Transition1x reactant / TS / product coordinates (tests/golden/t1x_geometries_b512.npz)"
               if geometry == "real" else "compact synthetic clouds of molecular density")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (bf16x3 split on tcgen05, pair16 edge storage; fp64 frames)", "data": "synthetic",
            "config": {"workload": f"batch={args.batch} Transition1x-shaped reactions (<=23 atoms) per GPU, {T} reverse steps + decode "
                                   f"({T + 1} LEFTNet evaluations), trained LEFTNet config (6 layers, H=196, R=96), states of a "
                                   f"trained model: z_t = alpha_t x + sigma_t eps re-drawn every step from {geo_txt}",
                       "geometry": geometry, "active_edge_fraction": active_fraction,
                       "literal_sample": literal,
                       "global_batch": len(all_sizes), "denoise_steps": T, "weights_broadcast_bytes": int(bcast_bytes), "nodes_per_gpu": int(sum(sizes) * 3),
                       "edges_per_gpu": workloads.edge_count(sizes), "parallelism": f"dp{world} (reactions sharded, no "
                       "per-step collective)", "l2": "working set per evaluation (edge state 4*E*684 B = "
                       f"{workloads.edge_count(sizes) * 684 * 4 / 1e6:.0f} MB) exceeds the 126 MB L2; no explicit flush",
                       "weights": "torch.manual_seed(0) default init (checkpoint is a git-LFS pointer)"},
            "clocks": clk, "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_ms,
            "spread": {"ms_per_step_min": srt[0], "ms_per_step_median": srt[len(srt) // 2], "ms_per_step_max": srt[-1],
                       "repeats": len(srt), "resident_ms_per_step": [round(x, 1) for x in per_step],
                       "resident_ms_per_100_reverse_steps": segments,
                       "e2e_ms_per_step": [round(x, 1) for x in per_e2e],
                       "note": "per-step CUDA-event times in launch order; this pool's B200s run this workload at the power cap "
                               "(clocks.reasons), step-to-step differences of a few percent follow the clock"},
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps, "outputs_finite": finite},
            "roofline": roof,
            "literal_sample": literal,
            "inpaint": inpaint,
            "train_step": train,
            "gemm_rooflines": gemm_tab,
            "kernels": {k: {kk: (round(vv, 6) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk not in ("flops", "bytes", "ms")}
                        for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["share"])[:14]}}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args, T, workloads)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line, on the process's real stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: everything libraries print on fd 1 while the benchmark runs (NCCL's version
    # banner at communicator creation, for one) is sent to stderr; emit() writes the result to the saved descriptor
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world == 1 and args.gpus > 1:
            print(f"note: --gpus {args.gpus} requested without torchrun; running 1 rank", file=sys.stderr)
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
