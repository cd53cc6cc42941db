"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: reaction sharding, weight broadcast, max-over-ranks,
bucketed gradient averaging."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oareactdiff_b200 import parallel, workloads


def test_shard_reactions_partition_and_balance():
    sizes = workloads.t1x_sizes(512, seed=0)
    for world in (1, 2, 4, 8):
        sh = parallel.shard_reactions(sizes, world)
        assert len(sh) == world and sh[0][0] == 0 and sh[-1][1] == len(sizes)
        assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        cost = [sum(3 * n * (3 * n - 1) for n in sizes[a:b]) for a, b in sh]
        assert min(cost) > 0 and max(cost) / (sum(cost) / world) < 1.05
    assert parallel.shard_reactions([5, 5], 2) == [(0, 1), (1, 2)]
    assert parallel.shard_reactions([4, 23, 4], 3) == [(0, 1), (1, 2), (2, 3)]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)  # different weights per rank before the broadcast
    m = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.LayerNorm(5))
    nbytes = parallel.broadcast_module_(m, src=0)
    flat = torch.cat([p.data.reshape(-1) for p in m.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    mx = parallel.max_over_ranks(10.0 + rank, "cpu")
    # gradient averaging: rank r has grad = (r + 1) on the first Linear's weight only; tiny buckets force several collectives
    m[0].weight.grad = torch.full_like(m[0].weight, float(rank + 1))
    nred = parallel.allreduce_gradients_(m, bucket_bytes=64)
    ok_grad = bool(torch.allclose(m[0].weight.grad, torch.full_like(m[0].weight, (1 + world) / 2.0))) and \
        all(float(p.grad.abs().max()) == 0.0 for n, p in m.named_parameters() if n != "0.weight")
    if rank == 0:
        out.put((nbytes, bool(all(torch.equal(g, gathered[0]) for g in gathered)), mx, nred, ok_grad))
    dist.destroy_process_group()


def test_broadcast_and_max_over_ranks_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    nbytes, equal, mx, nred, ok_grad = q.get()
    assert nbytes == (7 * 5 + 5 + 5 + 5) * 4 and equal and mx == 11.0
    assert nred == (7 * 5 + 5 + 5 + 5) * 4 and ok_grad
