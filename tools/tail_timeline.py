"""GPU probe: clock64 timeline of one row tile of the fused GCL tail (gcl_tail.cuh, CTA 0, third tile, layer 0).
    python tools/tail_timeline.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oareactdiff_b200 as ob  # noqa: E402
from oareactdiff_b200 import workloads  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    cfg = dict(cutoff=10.0, num_layers=6, hidden_channels=196, num_radial=96, in_hidden_channels=8, reflect_equiv=True,
               legacy=True, update=True, object_aware=True)
    torch.manual_seed(0)
    m = ob.LEFTNetB200(**cfg).to(dev)
    sizes = workloads.t1x_sizes(64, seed=0)
    nodes, h0, cond = workloads.reaction_batch(sizes, seed=0)
    masks = [ob.get_mask_for_frag(n) for n in nodes]
    cm = torch.cat(masks)
    ei = ob.get_edges_index(cm, remove_self_edge=True)
    nfs = ob.get_n_frag_switch(nodes)
    sub = ob.get_subgraph_mask(ei, nfs)
    N = cm.numel()
    g = torch.Generator().manual_seed(1)
    h = torch.randn(N, 8, generator=g).to(dev)
    pos = (torch.randn(N, 3, generator=g) * 1.5).to(dev)
    eng = m.engine(dev)
    for _ in range(3):
        m(h, pos, ei.to(dev), subgraph_mask=sub.to(dev))
    eng.set_debug(True)
    m(h, pos, ei.to(dev), subgraph_mask=sub.to(dev))
    ts = eng.read("tail_ts", torch.int64).view(16, 64)
    eng.set_debug(False)
    t0 = int(ts[ts > 0].min())
    us = lambda v: (int(v) - t0) / 1965.0 if int(v) > 0 else float("nan")  # cycles at ~1.9 GHz -> us (approximate)
    mma = ts[12]
    print("MMA warp: wait tile_done %.2f -> %.2f | layer2 issued %.2f | m_full %.2f" % (us(mma[0]), us(mma[1]), us(mma[2]), us(mma[3])))
    print("  layer 3 column tiles (start after acc3_empty, issued):", " ".join("%.1f-%.1f" % (us(mma[4 + 2 * k]), us(mma[5 + 2 * k])) for k in range(8)))
    for w in (0, 3, 4, 8, 11):
        r = ts[w]
        print("epilogue warp %2d: wait acc2 %.2f -> %.2f | pass1 done %.2f | att/m_full %.2f | pass2 done %.2f | tile done %.2f" %
              (w, us(r[0]), us(r[1]), us(r[2]), us(r[3]), us(r[4]), us(r[40])))
        print("   blocks (acc ready, resid ready, compute done, staging free):", " ".join(
            "[%.1f %.1f %.1f %.1f]" % (us(r[8 + 4 * k]), us(r[9 + 4 * k]), us(r[10 + 4 * k]), us(r[11 + 4 * k])) for k in range(8)))


if __name__ == "__main__":
    main()
