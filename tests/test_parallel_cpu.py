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


def _sampling_worker(rank, world, port, out):
    """World-2 sharded sampling on the CPU: the fp64 oracle stands in for the CUDA engine behind LEFTNetB200.forward."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oareactdiff_b200 as ob
    from oracle import oa_ref
    from tests.test_reference_suite_cpu import _oracle_forward
    ob.LEFTNetB200.forward = _oracle_forward
    ob.EGNNDynamics.fused_ok = lambda self, device: False
    torch.set_num_threads(2)
    cfg = dict(oa_ref.TRAINED_CFG, hidden_channels=32, num_radial=16, num_layers=1, cutoff=5.0)
    torch.manual_seed(50 + rank)  # different initial weights per rank: the broadcast must make them equal
    dyn = ob.EGNNDynamics(model_config=cfg, fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0, condition_nf=1,
                          model=ob.LEFTNetB200, device=torch.device("cpu"))
    parallel.broadcast_module_(dyn, src=0)
    ddpm = ob.EnVariationalDiffusion(dynamics=dyn, schdule=ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", 4, 1e-5), (1.0, 1.0, 1.0)),
                                     normalizer=ob.Normalizer(), pos_only=True)
    sizes = [6, 3, 4, 5, 3]
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, 3)
    final, (a, b) = parallel.sample_sharded(ddpm, nodes, cond, h0=h0, seed=11)
    # the same chunk, sampled alone with this rank's seed: what the gathered result must contain at this rank's rows
    c = [0] + torch.cumsum(nodes[0], 0).tolist()
    torch.manual_seed(11 + rank)
    alone, _ = ddpm.sample(b - a, [n[a:b] for n in nodes], cond[a:b], h0=[h[c[a]:c[b]] for h in h0])
    ok = all(torch.equal(final[f][c[a]:c[b]], alone[0][f]) for f in range(3))
    shapes = [tuple(x.shape) for x in final]
    h_ok = all(torch.equal(final[f][:, 3:].long(), h0[f]) for f in range(3))  # pos_only: atom types come back in global order
    res = [None] * world
    dist.all_gather_object(res, (rank, (a, b), ok, shapes, h_ok, float(final[0].abs().sum())))
    if rank == 0:
        out.put(res)
    dist.destroy_process_group()


def test_sharded_sampling_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_sampling_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = q.get()
    (r0, s0, ok0, sh0, h0ok, sum0), (r1, s1, ok1, sh1, h1ok, sum1) = res
    assert (r0, r1) == (0, 1) and s0[0] == 0 and s0[1] == s1[0] and s1[1] == 5 and s0[1] in (1, 2, 3)
    assert ok0 and ok1 and h0ok and h1ok
    assert sh0 == sh1 == [(21, 9)] * 3 and sum0 == sum1  # every rank holds the same gathered result


def test_all_gather_rows_and_sharding_without_process_group():
    x = torch.arange(6.0).view(3, 2)
    assert parallel.all_gather_rows(x) is x
    assert parallel.shard_reactions([4, 5], 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]  # more ranks than reactions: empty chunks
