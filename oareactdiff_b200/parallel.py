"""Multi-GPU plumbing for the sampling path: reactions are independent, so a batch shards across ranks with
no per-step collective.  One process per GPU (`torch.distributed`, NCCL on GPUs / gloo in CPU tests); the only
communication is one broadcast of the weights at start-up and small reductions of timings/outputs."""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import nn


def shard_reactions(sizes: Sequence[int], world: int, n_frag: int = 3) -> List[Tuple[int, int]]:
    """Split reactions 0..B-1 into `world` contiguous chunks balanced by edge count sum(3n(3n-1)) (the cost driver),
    never leaving a rank empty when B >= world.  Returns [(start, stop)] per rank."""
    B = len(sizes)
    if B <= world:  # at most one reaction per rank; the last world - B ranks stay empty
        return [(min(k, B), min(k + 1, B)) for k in range(world)]
    cost = [n_frag * n * (n_frag * n - 1) + 1 for n in sizes]
    total = float(sum(cost))
    bounds, acc, r = [0], 0.0, 1
    for i, c in enumerate(cost):
        acc += c
        remaining_ranks = world - r
        remaining_items = B - (i + 1)
        if r < world and (acc >= total * r / world or remaining_items == remaining_ranks) and remaining_items >= remaining_ranks:
            bounds.append(i + 1)
            r += 1
    while len(bounds) < world:
        bounds.append(B)
    bounds.append(B)
    return [(bounds[k], bounds[k + 1]) for k in range(world)]


@torch.no_grad()
def broadcast_module_(module: nn.Module, src: int = 0) -> int:
    """Make every rank's parameters and buffers equal to rank `src`'s with ONE flat broadcast; returns bytes sent."""
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    floats = [t for t in tensors if t.is_floating_point()]
    if not floats or not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in floats])
    dist.broadcast(flat, src)
    o = 0
    for t in floats:
        t.copy_(flat[o:o + t.numel()].view_as(t).to(t.dtype))
        o += t.numel()
    return flat.numel() * 4


@torch.no_grad()
def allreduce_gradients_(module: nn.Module, bucket_bytes: int = 64 << 20) -> int:
    """Data-parallel training (the reference trains under Lightning DDP, trainer/train_ts1x.py:214-232): average the
    parameter gradients over the ranks in flat buckets (one collective per ~64 MB instead of one per tensor: on NVSwitch the
    cost is launch latency, not link count).  Parameters without a gradient on this rank contribute zeros, so every rank
    issues the same collectives.  Returns the bytes reduced."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    params = [p for p in module.parameters() if p.requires_grad]
    total, bucket, size = 0, [], 0

    def flush():
        nonlocal bucket, size, total
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32) for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= world
        o = 0
        for p in bucket:
            g = flat[o:o + p.numel()].view_as(p).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            o += p.numel()
        total += flat.numel() * 4
        bucket, size = [], 0

    for p in params:
        bucket.append(p)
        size += p.numel() * 4
        if size >= bucket_bytes:
            flush()
    flush()
    return total


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([float(value)], device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_gather_rows(x: torch.Tensor) -> torch.Tensor:
    """Concatenate `x` ([n_r, ...], n_r differing between ranks) over the ranks in rank order: sizes first, then one padded
    all_gather (KBs for the samplers' final [N_f, 9] tensors: latency-bound, so one collective instead of `world`)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    world = dist.get_world_size()
    n = torch.tensor([x.size(0)], device=x.device, dtype=torch.int64)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    ns = [int(v) for v in ns]
    pad = torch.zeros((max(max(ns), 1),) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
    pad[:x.size(0)] = x
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:k] for p, k in zip(parts, ns)], dim=0)


@torch.no_grad()
def sample_sharded(ddpm, fragments_nodes: List[torch.Tensor], conditions: torch.Tensor, h0=None, seed: int = 0,
                   gather: bool = True, inpaint_kwargs=None, **sample_kwargs):
    """`ddpm.sample(...)` (or `ddpm.inpaint(...)` when `inpaint_kwargs` — xh_fixed, frag_fixed, resamplings, jump_length — is
    given) of a GLOBAL batch, sharded over the ranks: every rank takes a contiguous, edge-balanced chunk of the
    reactions, seeds its generators with `seed + rank`, runs the unchanged sampler on its chunk — no per-step communication —
    and, with `gather`, the final frames are concatenated over the ranks, which restores the global reaction order per
    fragment.  Arguments are the GLOBAL tensors, identical on every rank (`fragments_nodes[f]` [B], `conditions` [B, c],
    `h0[f]` / `xh_fixed[f]` [N_f, .] in the reference's per-fragment node order).  Returns (final frame per fragment,
    (start, stop) of this rank's reactions)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    sizes = [int(v) for v in fragments_nodes[0].tolist()]
    a, b = shard_reactions(sizes, world, n_frag=len(fragments_nodes))[rank]

    def rows(f):  # node range of reactions [a, b) inside fragment f
        c = torch.cumsum(fragments_nodes[f], 0)
        return (int(c[a - 1]) if a > 0 else 0), (int(c[b - 1]) if b > 0 else 0)

    local_nodes = [n[a:b] for n in fragments_nodes]
    cut = (lambda xs: None if xs is None else [x[rows(f)[0]:rows(f)[1]] for f, x in enumerate(xs)])
    torch.manual_seed(seed + rank)
    if b > a:
        if inpaint_kwargs is not None:
            kw = dict(inpaint_kwargs)
            kw["xh_fixed"] = [x.clone() for x in cut(kw["xh_fixed"])]
            out, _ = ddpm.inpaint(b - a, local_nodes, conditions[a:b], **kw, **sample_kwargs)
        else:
            out, _ = ddpm.sample(b - a, local_nodes, conditions[a:b], h0=cut(h0), **sample_kwargs)
        final = list(out[0])
    else:  # more ranks than reactions: this rank contributes empty frames
        dev = conditions.device
        final = [torch.zeros(0, ddpm.node_nfs[f], device=dev) for f in range(len(fragments_nodes))]
    if gather:
        final = [all_gather_rows(x.contiguous()) for x in final]
    return final, (a, b)
