"""GPU probe (not a test): where does a reverse step's time go?  Times, with CUDA events on the current stream,
(a) the LEFTNet evaluation alone through the C ABI (CUDA-graph replay and eager), (b) EGNNDynamics.forward,
(c) the full reverse step of the sampler, all on the same compact B=64 Transition1x-shaped state.
    python tools/perf_probe.py [reps]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oareactdiff_b200 as ob  # noqa: E402
from oareactdiff_b200 import workloads  # noqa: E402


def timed(fn, reps, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    host = (time.perf_counter() - t0) * 1e3 / reps
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, host


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, T = 64, 1000
    cfg = dict(cutoff=10.0, num_layers=6, hidden_channels=196, num_radial=96, in_hidden_channels=8, reflect_equiv=True,
               legacy=True, update=True, object_aware=True)
    torch.manual_seed(0)
    dyn = ob.EGNNDynamics(model_config=cfg, fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0,
                          condition_nf=1, model=ob.LEFTNetB200, device=dev).to(dev)
    sched = ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", T, 1e-5), norm_values=(1.0, 1.0, 1.0))
    ddpm = ob.EnVariationalDiffusion(dynamics=dyn, schdule=sched, normalizer=ob.Normalizer(), pos_only=True).to(dev)
    dyn.model.assume_static_weights = True
    sizes = workloads.t1x_sizes(B, seed=0)
    nodes, h0, cond = workloads.reaction_batch(sizes, seed=0)
    nodes, h0, cond = [x.to(dev) for x in nodes], [x.to(dev) for x in h0], cond.to(dev)
    masks, edge_index, nfs = ddpm._setup(B, nodes)
    tab = ddpm._tables(T, dev)
    ddpm._seg_setup(masks)
    gen = torch.Generator().manual_seed(1)
    Z = torch.cat([torch.cat([torch.randn(h.size(0), 3, generator=gen) * 1.5, h.cpu()], dim=1) for h in h0]).to(dev)
    H0 = torch.cat(h0)
    eng = dyn.model.engine(dev)
    out = {}

    # (c) full reverse step (what sample() runs per step): device-resident step (one CUDA graph) and the host-composed one
    Zd = Z.clone().contiguous()
    ddpm._device_setup(Zd, masks, edge_index, nfs, cond, H0)

    def dstep():
        Zd.copy_(Z)
        ddpm._device_step(500, Zd, tab)
    out["device_step_ms"], out["device_step_host_ms"] = timed(dstep, reps)
    _, out["device_step_host_unthrottled_ms"] = timed(dstep, 6, warm=0)
    dyn.use_fused = False

    def step():
        z = ddpm._fast_step(500, Z, tab, edge_index, nfs, masks, cond)
        z[:, 3:] = H0
    out["sampler_step_ms"], out["sampler_step_host_ms"] = timed(step, reps)
    # 6 steps (~500 launches) fit in the driver's launch queue, so the host is not throttled by the GPU: pure enqueue cost
    _, out["sampler_step_host_unthrottled_ms"] = timed(step, 6, warm=0)

    # (b) dynamics forward only
    tt = tab["tt"][501].expand(B, 1)

    def dyn_fwd():
        dyn(xh=ddpm._views(Z), edge_index=edge_index, t=tt, conditions=cond, n_frag_switch=nfs, combined_mask=ddpm._combined)
    out["dynamics_ms"], out["dynamics_host_ms"] = timed(dyn_fwd, reps)
    dyn.use_fused = True
    out["dynamics_fused_ms"], out["dynamics_fused_host_ms"] = timed(dyn_fwd, reps)

    # (a) LEFTNet evaluation alone through the C ABI
    N = Z.size(0)
    h_in = torch.randn(N, 8, device=dev)
    pos = Z[:, :3].contiguous()
    sub = dyn._graph_cache(edge_index, nfs, ddpm._combined)["sub"]

    def leftnet():
        eng.forward(h_in, pos, sub)
    out["leftnet_graph_ms"], out["leftnet_graph_host_ms"] = timed(leftnet, reps)
    os.environ["OARD_PROBE"] = "1"
    eng.set_debug(False)
    eng.set_profile(1)  # profile every forward => eager launches with event pairs
    out["leftnet_eager_profiled_ms"], _ = timed(leftnet, max(reps // 5, 5))
    prof = eng.profile()
    eng.set_profile(0)
    tot = sum(v["ms"] for k, v in prof.items() if not k.startswith("_"))
    nfw = max(prof["k_edge_mask"]["launches"], 1)
    out["leftnet_kernel_sum_ms"] = tot / nfw
    out["per_kernel_us"] = {k: round(1e3 * v["ms"] / nfw, 1) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
                            if not k.startswith("_")}
    out["launches_per_forward"] = eng.launches()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
