#!/usr/bin/env python
"""bench.py — reactions/sec for a full 1000-step reverse diffusion on Transition1x-shaped batches (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on the host cores (oracle port)

One "step" = one full `EnVariationalDiffusion.sample()` of a B=64 batch: 1000 reverse steps + the final p(x|z0)
decode = 1001 denoiser evaluations (config.workload names it).  Weak scaling: every rank samples its own B=64 batch.
`value` = reactions per second with inputs resident in HBM; `e2e` = the same through the public API with HOST (pinned)
inputs copied in and results copied out inside the timed region.  Prints ONE JSON line on rank 0.
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
# stdout carries exactly ONE JSON line: NCCL's version banner (NCCL_DEBUG=VERSION/INFO prints it on stdout) is silenced
if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "INFO"):
    os.environ["NCCL_DEBUG"] = "WARN"

METRIC = "reactions/sec full 1000-step reverse diffusion, Transition1x-shaped batch"
UNIT = "reactions/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="reactions per GPU")
    ap.add_argument("--denoise-steps", type=int, default=1000, help="T of the reverse diffusion (BASELINE: 1000)")
    ap.add_argument("--profile-every", type=int, default=97, help="bracket kernels with CUDA events every n-th forward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-replay", action="store_true", help="skip the realistic-geometry replay line")
    ap.add_argument("--replay-geometry", default="synthetic", choices=["synthetic", "real"],
                    help="replay states from compact synthetic clouds (default, the measured configuration) or from the frozen "
                         "REAL Transition1x geometries of tests/golden/t1x_geometries_b512.npz (oracle/gen_t1x_geometries.py)")
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
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def tune_cpu_threads():
    """Eager PyTorch on many-core hosts is often fastest well below the core count (oversubscription of small ops):
    time one small denoiser evaluation at a few thread counts and keep the best, so the CPU baseline is a fair one."""
    import torch
    from oracle import oa_ref
    from oareactdiff_b200 import workloads
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
    from oareactdiff_b200 import workloads
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
def cpu_baseline(args, T):
    import torch
    from oracle import oa_ref
    from oareactdiff_b200 import workloads
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
                      f"({per_eval:.2f} s/eval), extrapolated to {T + 1}"}


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import oareactdiff_b200 as ob
    from oareactdiff_b200 import workloads

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
    from oareactdiff_b200 import parallel
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
    pin = lambda t: t.pin_memory()
    nodes_h, h0_h, cond_h = [pin(x) for x in nodes_h], [pin(x) for x in h0_h], pin(cond_h)
    nodes_d, h0_d, cond_d = [x.to(dev) for x in nodes_h], [x.to(dev) for x in h0_h], cond_h.to(dev)
    eng = dyn.model.engine(dev)

    def step_resident():
        torch.manual_seed(1234 + rank)
        out, _ = ddpm.sample(B, nodes_d, cond_d, h0=h0_d)
        return out[0]

    # Replay (SURVEY §8d realism caveat): the same per-step work (denoiser + posterior sampling) on states
    # z_t = alpha_t x + sigma_t eps built from compact synthetic geometries, i.e. what a TRAINED model sees: every
    # same-fragment edge stays inside the 10 A cutoff.  With random weights the literal sample() drifts to |x| ~ 1e2 A.
    gen = torch.Generator().manual_seed(99 + rank)
    geo = []
    for n in sizes:
        r = 1.2 * n ** (1.0 / 3.0)
        pts = torch.randn(n, 3, generator=gen)
        pts = pts / pts.norm(dim=1, keepdim=True) * (torch.rand(n, 1, generator=gen) ** (1 / 3)) * r
        geo.append(pts - pts.mean(0, keepdim=True))
    x_frag = []
    for f in range(3):
        xs = [gp + 0.3 * torch.randn(gp.shape, generator=gen) for gp in geo]
        x_frag.append(torch.cat([x - x.mean(0, keepdim=True) for x in xs]).to(dev))
    if args.replay_geometry == "real":
        # reactant / transition-state / product geometries of real Transition1x reactions with exactly these atom counts
        import numpy as np
        fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", "t1x_geometries_b512.npz"))
        if hi > len(fx["sizes"]) or [int(v) for v in fx["sizes"][lo:hi]] != list(sizes):
            raise SystemExit("--replay-geometry real: the fixture holds the first 512 reactions of t1x_sizes(., seed=0) only")
        off = np.concatenate([[0], np.cumsum(fx["sizes"])])
        x_frag = [torch.from_numpy(fx[k][off[lo]:off[hi]]).to(dev) for k in ("reactant", "transition_state", "product")]
    xh0_d = [torch.cat([x_frag[f], h0_d[f]], dim=1) for f in range(3)]

    replay_Z = torch.empty(sum(h.size(0) for h in h0_d), 9, device=dev)

    def step_replay():
        torch.manual_seed(4321 + rank)
        masks, edge_index, nfs = ddpm._setup(B, nodes_d)
        tab = ddpm._tables(T, dev)
        ddpm._seg_setup(masks)
        X = torch.cat(xh0_d)
        H0 = torch.cat(h0_d)
        on_device = ddpm._device_ok(dev)
        if on_device:
            ddpm._device_setup(replay_Z, masks, edge_index, nfs, cond_d, H0)
        for s_int in reversed(range(T)):
            # state a trained model would see at t = s+1: q(z_t | x); then the usual reverse step (denoiser + posterior)
            Z = tab["alpha"][s_int + 1] * X + tab["sigma_abs"][s_int + 1] * ddpm._noise_cat(masks)
            Z[:, 3:] = H0
            if on_device:  # the device step replays one CUDA graph on a persistent state buffer
                replay_Z.copy_(Z)
                ddpm._device_step(s_int, replay_Z, tab)
            else:
                Z = ddpm._fast_step(s_int, Z, tab, edge_index, nfs, masks, cond_d)
        Z0 = tab["alpha"][0] * X + tab["sigma_abs"][0] * ddpm._noise_cat(masks)
        Z0[:, 3:] = H0
        return ddpm.sample_p_xh_given_z0(ddpm._views(Z0), edge_index, nfs, masks, B, cond_d)[0]

    out_host = [torch.empty(h.size(0), 9).pin_memory() for h in h0_h]

    def step_e2e():
        torch.manual_seed(1234 + rank)
        nd = [x.to(dev, non_blocking=True) for x in nodes_h]
        hd = [x.to(dev, non_blocking=True) for x in h0_h]
        cd = cond_h.to(dev, non_blocking=True)
        out, _ = ddpm.sample(B, nd, cd, h0=hd)
        for dst, src in zip(out_host, out[0]):
            dst.copy_(src.to(torch.float32), non_blocking=True)
        return out[0]

    h2d =sum(x.numel() * x.element_size() for x in nodes_h + h0_h + [cond_h])
    d2h = sum(x.numel() * x.element_size() for x in out_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = {}

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(k):
            fn()
        host_ms[fn.__name__] = (time.perf_counter() - t0) * 1e3 / k  # host enqueue time (no sync inside the loop)
        e1.record()
        barrier()
        return parallel.max_over_ranks(e0.elapsed_time(e1), dev)

    for _ in range(args.warmup):
        step_resident()
    eng.set_profile(args.profile_every)
    l0 = eng.total_launches()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ms = timed(step_resident, args.steps)
    clk = clocks.stop()
    launches = eng.total_launches() - l0
    prof = eng.profile()
    eng.set_profile(0)
    ms_e2e = timed(step_e2e, args.steps)
    finite = all(bool(torch.isfinite(o).all()) for o in out_host)
    replay = None
    if not args.no_replay:
        step_replay()
        eng.set_profile(args.profile_every)
        ms_rp = timed(step_replay, 1)
        prof_rp = eng.profile()
        eng.set_profile(0)
        af = prof_rp.get("_active_fraction")
        replay = {"value": len(all_sizes) / (ms_rp / 1e3), "unit": UNIT, "ms_per_step": ms_rp, "steps": 1,
                  "active_edge_fraction": (af["flops"] / max(af["launches"], 1)) if af else None,
                  "geometry": args.replay_geometry,
                  "note": "same per-step work on z_t = alpha_t x + sigma_t eps from "
                          + ("REAL Transition1x reactant / TS / product geometries (tests/golden/t1x_geometries_b512.npz) "
                             if args.replay_geometry == "real" else "compact synthetic geometries ")
                          + "(what a trained model sees; every same-fragment edge inside the cutoff)",
                  "kernels_ms_per_launch": {k: round(v["ms"] / max(v["launches"], 1), 5) for k, v in
                                            sorted(prof_rp.items(), key=lambda kv: -kv[1]["ms"])[:10] if not k.startswith("_")}}

    total_reactions = len(all_sizes) * args.steps
    value = total_reactions / (ms / 1e3)
    e2e_v = total_reactions / (ms_e2e / 1e3)
    if rank != 0:
        return
    pk = peaks()
    af_lit = prof.pop("_active_fraction", None)
    kernels = {k: dict(ms_per_launch=v["ms"] / max(v["launches"], 1), launches=v["launches"],
                       tflops=(v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 else 0.0,
                       gbs=(v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else 0.0,
                       share=v["ms"]) for k, v in prof.items()}
    tot = sum(v["share"] for v in kernels.values()) or 1.0
    for v in kernels.values():
        v["share"] = v["share"] / tot
    dom = max(kernels, key=lambda k: kernels[k]["share"]) if kernels else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if dom and os.path.exists(tpath):
        t_ = json.load(open(tpath)).get(dom)
        traffic = t_.get("dram_bytes_per_launch") if isinstance(t_, dict) else t_  # ncu dram read+write bytes of one launch
    roof = None
    ridge = pk["tflops"] * 1e12 / (pk["hbm_gbs"] * 1e9)  # FLOP per byte where the two roofs meet (bf16 tensor vs HBM)
    if dom:
        kd = kernels[dom]
        pv = prof[dom]
        # tensor pipe executes 3 bf16 MMAs per algorithmic fp32 product (bf16x3 split): that is what competes with HBM
        intensity = (3.0 * pv["flops"] / pv["bytes"]) if pv["bytes"] > 0 else float("inf")
        if dom.startswith("gemm") and intensity >= ridge:
            roof = {"kernel": dom, "bound": "tensor", "achieved": kd["tflops"], "peak": pk["tflops"], "unit": "TFLOP/s",
                    "frac": kd["tflops"] / pk["tflops"], "traffic": traffic, "share_of_step": kd["share"],
                    "mma_issued_tflops": 3.0 * kd["tflops"],
                    "note": f"achieved = algorithmic fp32-equivalent flops (2MNK, inactive edges skipped) / CUDA-event launch time; "
                            f"the kernel issues 3 bf16 MMAs per product (bf16x3 split), so its own ceiling is peak/3; "
                            f"peak = bf16 cuBLAS sustained ({pk['src']})"}
        else:
            roof = {"kernel": dom, "bound": "hbm", "achieved": kd["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": kd["gbs"] / pk["hbm_gbs"], "traffic": traffic, "share_of_step": kd["share"],
                    "intensity_flop_per_byte": intensity, "ridge_flop_per_byte": ridge,
                    "note": f"achieved = algorithmic bytes (operand read + residual read + result write, fp32) / CUDA-event "
                            f"launch time; arithmetic intensity (bf16 MMA flops per byte) is below the ridge, so HBM is the "
                            f"bound; peak = copy bandwidth ({pk['src']})"}
    # the message-passing kernel's roofline is quoted on the replay (trained-model geometry, active fraction ~0.32): in the
    # literal random-weight trajectory the cutoff empties and the kernel has almost no edges to stream
    mp = kernels.get("k_equi_reduce")
    mp_src = "literal sample()"
    if replay is not None and "k_equi_reduce" in prof_rp and prof_rp["k_equi_reduce"]["ms"] > 0:
        v = prof_rp["k_equi_reduce"]
        mp = dict(gbs=v["bytes"] / (v["ms"] * 1e-3) / 1e9, ms_per_launch=v["ms"] / max(v["launches"], 1))
        mp_src = "replay"
    gemm_tab = {k: {"tflops_alg": round(v["tflops"], 1), "frac_of_bf16x3_ceiling": round(3.0 * v["tflops"] / pk["tflops"], 3),
                    "gbs_alg": round(v["gbs"], 1), "frac_hbm": round(v["gbs"] / pk["hbm_gbs"], 3)}
                for k, v in kernels.items() if k.startswith("gemm")}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"batch={args.batch} Transition1x-shaped reactions (<=23 atoms) per GPU, {T} steps "
                                   f"(sample(): {T + 1} LEFTNet evaluations), trained LEFTNet config (6 layers, H=196, R=96)",
                       "global_batch": len(all_sizes), "denoise_steps": T, "weights_broadcast_bytes": int(bcast_bytes), "nodes_per_gpu": int(sum(sizes) * 3),
                       "edges_per_gpu": workloads.edge_count(sizes), "parallelism": f"dp{world} (reactions sharded, no "
                       "per-step collective)", "l2": "working set per evaluation (edge state 4*E*684 B = "
                       f"{workloads.edge_count(sizes) * 684 * 4 / 1e6:.0f} MB) exceeds the 126 MB L2; no explicit flush",
                       "weights": "torch.manual_seed(0) default init (checkpoint is a git-LFS pointer)"},
            "clocks": clk, "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_ms,
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps, "outputs_finite": finite},
            "active_edge_fraction": (af_lit["flops"] / max(af_lit["launches"], 1)) if af_lit else None,
            "replay": replay,
            "roofline": roof,
            "message_passing_roofline": None if not mp else {
                "kernel": "k_equi_reduce", "bound": "hbm", "achieved": mp["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": mp["gbs"] / pk["hbm_gbs"], "ms_per_launch": mp["ms_per_launch"], "measured_on": mp_src},
            "gemm_rooflines": gemm_tab,
            "kernels": {k: {kk: (round(vv, 6) if isinstance(vv, float) else vv) for kk, vv in v.items()}
                        for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["share"])[:12]}}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args, T)
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
