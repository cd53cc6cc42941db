"""Synthetic Transition1x-shaped sampling inputs (the role of the reference's `utils/sampling_tools.py:64-108`
`assemble_sample_inputs`): fragments_nodes = [n, n, n] per reaction, h0 = [one-hot(5) | atomic number] drawn from
{H, C, N, O} with Transition1x frequencies, conditions = zeros.  The atom-count histogram is the one of the 9000
`use_ind` reactions of Transition1x (min 4, max 23, mean 13.57) so no dataset is needed at run time."""
from typing import List, Tuple

import numpy as np
import torch

T1X_NATOMS_HIST = [0, 0, 0, 0, 1, 4, 15, 36, 157, 261, 599, 947, 1155, 1207, 1461, 1051, 907, 636, 200, 280, 25, 55, 0, 3]
ELEMENT_Z = np.array([1, 6, 7, 8, 9])
ELEMENT_P = np.array([0.445, 0.290, 0.149, 0.116, 0.0])


def t1x_sizes(batch: int, seed: int = 0) -> List[int]:
    p = np.asarray(T1X_NATOMS_HIST, dtype=np.float64)
    return [int(x) for x in np.random.RandomState(seed).choice(len(p), size=batch, p=p / p.sum())]


def reaction_batch(sizes: List[int], seed: int = 0) -> Tuple[List[torch.Tensor], List[torch.Tensor], torch.Tensor]:
    """-> (fragments_nodes [3 x LongTensor[B]], h0 [3 x FloatTensor[sum n, 6]], conditions FloatTensor[B, 1]) on CPU."""
    rng = np.random.RandomState(seed)
    nodes = torch.tensor(list(sizes), dtype=torch.long)
    types = np.concatenate([rng.choice(5, size=n, p=ELEMENT_P) for n in sizes])
    h = np.zeros((types.size, 6), dtype=np.float32)
    h[np.arange(types.size), types] = 1
    h[:, 5] = ELEMENT_Z[types]
    h0 = torch.from_numpy(h)
    return [nodes, nodes.clone(), nodes.clone()], [h0, h0.clone(), h0.clone()], torch.zeros(len(sizes), 1)


def edge_count(sizes: List[int], n_frag: int = 3) -> int:
    return sum(n_frag * n * (n_frag * n - 1) for n in sizes)


def active_edge_count(sizes: List[int], n_frag: int = 3) -> int:
    """Same-fragment directed edges (the most the cutoff mask can leave active)."""
    return sum(n_frag * n * (n - 1) for n in sizes)
