"""ORACLE tooling (test infrastructure, build container only): this package's graph helpers against the UNMODIFIED
reference's (oa_reactdiff/utils/_graph_tools.py:9-96) on random inputs — unsorted sample ids, empty samples and fragments,
distance cut-off, self edges kept or removed.  Bit-exact or it counts as a mismatch.  Prints one JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)

from oa_reactdiff.utils import get_edges_index, get_mask_for_frag, get_n_frag_switch, get_subgraph_mask  # noqa: E402

import oareactdiff_b200 as ob  # noqa: E402


def main(n_cases=400):
    g = torch.Generator().manual_seed(0)
    bad, checks = [], 0
    for it in range(n_cases):
        n = int(torch.randint(1, 14, (1,), generator=g))
        cm = torch.randint(0, 4, (n,), generator=g)  # sample ids in arbitrary order, some samples empty
        pos = torch.randn(n, 3, generator=g)
        for rs in (False, True):
            for cut in (None, 1.0):
                a = get_edges_index(cm, pos=pos, edge_cutoff=cut, remove_self_edge=rs)
                b = ob.get_edges_index(cm, pos=pos, edge_cutoff=cut, remove_self_edge=rs)
                checks += 1
                if not torch.equal(a, b):
                    bad.append(("edges", it, rs, cut))
        e = ob.get_edges_index(cm, remove_self_edge=True)
        nfs = torch.randint(0, 3, (n,), generator=g)
        if e.size(1):
            checks += 1
            if not torch.equal(get_subgraph_mask(e, nfs), ob.get_subgraph_mask(e, nfs)):
                bad.append(("subgraph_mask", it))
        B = int(torch.randint(1, 5, (1,), generator=g))
        frs = [torch.randint(0, 4, (B,), generator=g) for _ in range(int(torch.randint(1, 4, (1,), generator=g)))]
        checks += 2
        if not torch.equal(get_n_frag_switch(frs), ob.get_n_frag_switch(frs)):
            bad.append(("n_frag_switch", it))
        if not torch.equal(get_mask_for_frag(frs[0]), ob.get_mask_for_frag(frs[0])):
            bad.append(("mask_for_frag", it))
    print(json.dumps({"checks": checks, "mismatches": bad[:10], "n_mismatches": len(bad)}))


if __name__ == "__main__":
    main()
