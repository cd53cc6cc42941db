"""Graph/index helpers with the reference's names and results (oa_reactdiff/utils/_graph_tools.py:9-96),
built in O(E) on whatever device the inputs live on (the reference materialises an N x N boolean matrix and
builds the sub-graph mask on the CPU every step)."""
from typing import List, Optional

import torch
from torch import Tensor


def get_mask_for_frag(natm: Tensor) -> Tensor:
    """Sample id of every node of one fragment: tensor([2, 0, 3]) -> [0, 0, 2, 2, 2]  (_graph_tools.py:84-96)."""
    return torch.repeat_interleave(torch.arange(natm.size(0), device=natm.device), natm)


def get_n_frag_switch(natm_list: List[Tensor]) -> Tensor:
    """Fragment id of every node (_graph_tools.py:62-81)."""
    if len({int(n.shape[0]) for n in natm_list}) != 1:
        raise AssertionError("Tensor must be the same length for <natom_list>")
    dev = natm_list[0].device
    totals = torch.stack([n.sum() for n in natm_list]).to(dev)
    return torch.repeat_interleave(torch.arange(len(natm_list), device=dev), totals)


def get_edges_index(combined_mask: Tensor, pos: Optional[Tensor] = None, edge_cutoff: Optional[float] = None,
                    remove_self_edge: bool = False) -> Tensor:
    """Complete directed graph among nodes sharing a sample id, sorted by (source, target) — bit-identical to the
    reference's `torch.where(adj)` order (_graph_tools.py:9-36)."""
    cm = combined_mask
    n = cm.numel()
    dev = cm.device
    if n == 0:
        return torch.zeros(2, 0, dtype=torch.long, device=dev)
    order = torch.argsort(cm, stable=True)                    # nodes grouped by sample, ascending id inside
    uniq, inv, counts = torch.unique(cm, sorted=True, return_inverse=True, return_counts=True)
    starts = torch.cumsum(counts, 0) - counts                 # first slot of each sample inside `order`
    deg = counts[inv]                                         # row length of every node
    rows = torch.repeat_interleave(torch.arange(n, device=dev), deg)
    row_start = torch.cumsum(deg, 0) - deg
    within = torch.arange(rows.numel(), device=dev) - row_start[rows]
    cols = order[starts[inv][rows] + within]
    keep = torch.ones_like(rows, dtype=torch.bool)
    if remove_self_edge:
        keep &= rows != cols
    if edge_cutoff is not None:
        keep &= (pos[rows] - pos[cols]).norm(dim=-1) <= edge_cutoff
    return torch.stack([rows[keep], cols[keep]], dim=0)


def get_subgraph_mask(edge_index: Tensor, n_frag_switch: Tensor) -> Tensor:
    """1 for edges whose two ends lie in the same fragment (_graph_tools.py:39-59); stays on the device."""
    return (n_frag_switch[edge_index[0]] == n_frag_switch[edge_index[1]]).long()


def get_inner_edge_index(subgraph_mask: Tensor) -> Tensor:
    """Coordinates of the non-zero entries of a mask, one row per dimension (_graph_tools.py:99-100)."""
    return torch.stack(torch.nonzero(subgraph_mask, as_tuple=True), dim=0)
