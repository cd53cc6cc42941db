"""Data-parallel training probe (SURVEY §8f row 2 / BASELINE config 5 at N GPUs): a few SGD steps of the l2 objective on noised
real-geometry Transition1x triples, forward + backward through the CUDA training path, gradients averaged over the ranks with
`parallel.allreduce_gradients_` (flat buckets, NCCL).  Prints ONE JSON line on rank 0: per-step time (max over ranks, CUDA
events), the share of the gradient all-reduce, and the loss per step (it must go down).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/train_ddp_probe.py --batch 32 --steps 10"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oareactdiff_b200 as ob  # noqa: E402
from oareactdiff_b200 import parallel, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32, help="reactions per rank")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--lr", type=float, default=2e-4)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T = 1000
    cfg = dict(cutoff=10.0, num_layers=6, hidden_channels=196, num_radial=96, in_hidden_channels=8, reflect_equiv=True,
               legacy=True, update=True, object_aware=True)
    torch.manual_seed(0)
    dyn = ob.EGNNDynamics(model_config=cfg, fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0, condition_nf=1,
                          model=ob.LEFTNetB200, device=dev).to(dev)
    parallel.broadcast_module_(dyn, src=0)
    dyn.model.enable_training_path = True
    sched = ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", T, 1e-5), norm_values=(1.0, 1.0, 1.0))
    ddpm = ob.EnVariationalDiffusion(dynamics=dyn, schdule=sched, normalizer=ob.Normalizer(), pos_only=True).to(dev)
    ddpm.train(True)
    all_sizes = workloads.t1x_sizes(args.batch * world, seed=0)
    lo, hi = rank * args.batch, (rank + 1) * args.batch
    sizes = all_sizes[lo:hi]
    x_ref = workloads.real_geometries(lo, hi, sizes)
    _, h0, _ = workloads.reaction_batch(sizes, seed=rank)
    opt = torch.optim.AdamW(dyn.parameters(), lr=args.lr)
    torch.manual_seed(100 + rank)
    losses, t_step, t_comm = [], [], []
    for it in range(args.steps + 1):  # step 0 = warm-up (allocations, first-call attributes)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        reps, cond = workloads.training_batch(sizes, x_ref, h0, dev)
        e[0].record()
        opt.zero_grad(set_to_none=True)
        nll, _ = ddpm.compute_loss((reps, cond), scales=(1.0, 2.0, 1.0), training=True)
        loss = nll.mean()
        loss.backward()
        e[1].record()
        nbytes = parallel.allreduce_gradients_(dyn)
        e[2].record()
        torch.nn.utils.clip_grad_norm_(dyn.parameters(), 10.0)
        opt.step()
        e[3].record()
        torch.cuda.synchronize()
        lt = loss.detach().clone()
        if world > 1:
            dist.all_reduce(lt)
            lt /= world
        if it > 0:
            losses.append(float(lt))
            t_step.append(parallel.max_over_ranks(e[0].elapsed_time(e[3]), dev))
            t_comm.append(parallel.max_over_ranks(e[1].elapsed_time(e[2]), dev))
    if rank == 0:
        t_step.sort(); t_comm.sort()
        print(json.dumps({"n_gpus": world, "reactions_per_rank": args.batch, "steps": args.steps,
                          "ms_per_step_median": t_step[len(t_step) // 2], "ms_allreduce_median": t_comm[len(t_comm) // 2],
                          "allreduce_bytes": int(nbytes), "reactions_per_s": args.batch * world / (t_step[len(t_step) // 2] * 1e-3),
                          "loss_first": losses[0], "loss_last": losses[-1], "losses": [round(v, 5) for v in losses],
                          "loss_decreased": losses[-1] < losses[0]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
