"""Test-infrastructure shim: minimal `torch_scatter` so the UNMODIFIED reference
(/root/reference) imports in this container.  Not product code; only used by
oracle/gen_golden.py and tests that validate the oracle against the reference.

Semantics follow the published torch_scatter API (unpinned in the reference's
env.yaml:17): out[index[i]] (+)= src[i] along `dim`; mean divides by the
per-slot count clamped to >= 1.
"""
import torch


def _expand_index(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        shape = [1] * src.dim()
        shape[dim] = -1
        index = index.view(shape)
    return index.expand_as(src), dim


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    idx, dim = _expand_index(index.long(), src, dim)
    if out is None:
        if dim_size is None:
            dim_size = int(index.max().item()) + 1 if index.numel() else 0
        shape = list(src.shape)
        shape[dim] = dim_size
        out = torch.zeros(shape, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, idx, src)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    total = scatter_add(src, index, dim, out, dim_size)
    ones = torch.ones_like(src)
    count = scatter_add(ones, index, dim, None, total.shape[dim if dim >= 0 else src.dim() + dim])
    return total / count.clamp(min=1)


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if reduce in ("sum", "add"):
        return scatter_add(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    raise NotImplementedError(reduce)
