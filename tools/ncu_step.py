"""Profiling target (not a test): n device-resident reverse steps of the B=64 Transition1x-shaped batch on the state the
bench times (z_t at t = 501 from the frozen real Transition1x geometries, active fraction 0.317; OARD_GEOM=synthetic: a compact
random state), one CUDA graph per step.
    python tools/ncu_step.py [n_steps] [eager]
Used under `ncu --metrics gpu__time_duration.sum` (launch list) and `ncu --set full -k regex:...` (profiles/)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oareactdiff_b200 as ob  # noqa: E402
from oareactdiff_b200 import workloads  # noqa: E402


def main():
    n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, T = int(os.environ.get("OARD_B", "64")), 1000
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
    if os.environ.get("OARD_GEOM", "real") == "real" and B <= 512:
        # the state the bench times: z_t = alpha_t x + sigma_t eps at t = 501 from the frozen real Transition1x geometries
        x_ref = workloads.real_geometries(0, B, sizes)
        a, sg = tab["alpha"][501], tab["sigma_abs"][501]
        Z0 = torch.cat([torch.cat([a * (x - x.mean(0, keepdim=True) * 0) + sg * torch.randn(x.shape, generator=gen), h.cpu()], dim=1)
                        for x, h in zip(x_ref, h0)]).to(dev)
    else:
        Z0 = torch.cat([torch.cat([torch.randn(h.size(0), 3, generator=gen) * 1.5, h.cpu()], dim=1) for h in h0]).to(dev)
    Z = Z0.clone().contiguous()
    ddpm._device_setup(Z, masks, edge_index, nfs, cond, torch.cat(h0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(n_steps):
        Z.copy_(Z0)
        if i == n_steps - 1:
            e0.record()
        ddpm._device_step(500, Z, tab)
    e1.record()
    torch.cuda.synchronize()
    print(f"last step {e0.elapsed_time(e1):.3f} ms, launches/step {dyn.model.engine(dev).launches()}, finite {bool(torch.isfinite(Z).all())}")


if __name__ == "__main__":
    main()
