"""GPU probe: the per-trajectory fixed cost of the device-resident sampler (what a 1000-step trajectory pays once): graph
construction, oard_plan, dynamics plan, capture of the step graph, decode.  Runs short trajectories (T reverse steps) of the
bench batch repeatedly and prints host wall times per phase (device synchronised at every phase boundary) — outliers show
allocator / driver stalls.    python tools/setup_probe.py [n_trajectories] [T]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oareactdiff_b200 as ob  # noqa: E402
from oareactdiff_b200 import workloads  # noqa: E402


def main():
    n_traj = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    T_run = int(sys.argv[2]) if len(sys.argv) > 2 else 20
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
    x_ref = [x.to(dev) for x in workloads.real_geometries(0, B, sizes)]

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    for k in range(n_traj):
        t = [sync()]
        masks, edge_index, nfs = ddpm._setup(B, nodes)
        t.append(sync())
        tab = ddpm._tables(T, dev)
        ddpm._seg_setup(masks)
        H0 = torch.cat(h0).float()
        X = torch.cat([torch.cat([x.float(), h], dim=1) for x, h in zip(x_ref, h0)])
        X[:, :3] = ddpm._remove_mean_cat(X[:, :3])
        Z = torch.empty_like(X)
        t.append(sync())
        ddpm._device_setup(Z, masks, edge_index, nfs, cond, H0)
        t.append(sync())
        steps = []
        for s_int in reversed(range(T - T_run, T)):
            Zt = tab["alpha"][s_int + 1] * X + tab["sigma_abs"][s_int + 1] * ddpm._noise_cat(masks)
            Zt[:, 3:] = H0
            Z.copy_(Zt)
            ddpm._device_step(s_int, Z, tab)
            if len(steps) < 3:
                steps.append(sync())
        t.append(sync())
        Z0 = tab["alpha"][0] * X + tab["sigma_abs"][0] * ddpm._noise_cat(masks)
        Z0[:, 3:] = H0
        ddpm.sample_p_xh_given_z0(ddpm._views(Z0), edge_index, nfs, masks, B, cond)
        t.append(sync())
        ms = [round(1e3 * (b - a), 1) for a, b in zip(t[:-1], t[1:])]
        first = [round(1e3 * (b - a), 1) for a, b in zip([t[3]] + steps[:-1], steps)]
        print(f"traj {k}: graph build {ms[0]}  tables/state {ms[1]}  device_setup (plan) {ms[2]}  {T_run} steps {ms[3]} "
              f"(first three: {first})  decode {ms[4]}  | total {round(sum(ms), 1)} ms", flush=True)


if __name__ == "__main__":
    main()
