"""Test-infrastructure shim: the slice of torch_geometric's `MessagePassing`
the reference uses (leftnet.py:72-125,186-289,421-428), so the UNMODIFIED
reference imports here.  Not product code.

Published PyG semantics for flow="source_to_target": for every argument `foo_j`
of `message()` gather kwargs["foo"] at edge_index[0], for `foo_i` at
edge_index[1]; other names are passed through; aggregate at edge_index[1].
"""
import inspect

import torch
from torch_scatter import scatter


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", node_dim=-2, **kwargs):
        super().__init__()
        self.aggr = aggr
        self.node_dim = node_dim

    def jittable(self):
        return self

    def propagate(self, edge_index, size=None, **kwargs):
        names = list(inspect.signature(self.message).parameters)
        dim_size = None
        args = []
        for name in names:
            if name.endswith("_j") or name.endswith("_i"):
                src = kwargs[name[:-2]]
                dim_size = src.size(self.node_dim)
                row = edge_index[0] if name.endswith("_j") else edge_index[1]
                args.append(src.index_select(self.node_dim, row))
            else:
                args.append(kwargs[name])
        out = self.message(*args)
        out = self.aggregate(out, edge_index[1], None, dim_size)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        return scatter(inputs, index, dim=self.node_dim, dim_size=dim_size, reduce=self.aggr)

    def update(self, inputs):
        return inputs
