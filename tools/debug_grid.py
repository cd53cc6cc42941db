"""GPU diagnostic: stage-by-stage errors of selected option-grid cases (tests/test_gpu_grid.py) against the fp64 oracle,
next to the fp32 oracle's own gap.  Usage (GPU box): python tools/debug_grid.py [gname ...]"""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oa_ref  # noqa: E402
from tests.test_gpu_experimental import _grid_graphs  # noqa: E402
from tests.test_gpu_parity import DEV, make_leftnet  # noqa: E402
from tests.util import rel_err  # noqa: E402


def main():
    names = sys.argv[1:] or ["path", "two_components"]
    thr = float(os.environ.get("THR", "2e-4"))
    for gname in names:
        n, ei = _grid_graphs()[gname]
        g = torch.Generator().manual_seed(1)
        for reflect, oa, update, layers, cut, cutoff, scale in itertools.product(
                [True, False], [True, False], [True, False], [1, 3], [None, 3], [20.0, 2.5], [1.0, 3.0]):
            if n < 5 and cut:
                continue
            cfg = dict(cutoff=cutoff, num_layers=layers, hidden_channels=32, num_radial=16, in_hidden_channels=6,
                       reflect_equiv=reflect, legacy=True, update=update, object_aware=oa)
            sd = oa_ref.make_state_dict(oa_ref.leftnet_param_shapes(cfg), 3, cfg, dtype=torch.float64)
            h = torch.rand(n, 6, generator=g, dtype=torch.float64)
            pos = torch.rand(n, 3, generator=g, dtype=torch.float64) * scale
            sub = None
            if cut:
                s = (ei < cut).sum(0)
                sub = ((s == 2) | (s == 0)).long()[:, None]
            dbg, dbg32 = {}, {}
            ho_ref, dpos_ref = oa_ref.leftnet_forward(sd, cfg, h, pos, ei, sub, dbg=dbg)
            ho32, dp32 = oa_ref.leftnet_forward({k: v.float() for k, v in sd.items()}, cfg, h.float(), pos.float(), ei, sub, dbg=dbg32)
            m = make_leftnet(cfg, {k: v.float() for k, v in sd.items()})
            eng = m.engine(DEV)
            eng.set_debug(True)
            ho, po, _ = m(h.float().to(DEV), pos.float().to(DEV), ei.to(DEV), subgraph_mask=None if sub is None else sub.to(DEV))
            e = max(rel_err(ho.cpu(), ho_ref), rel_err((po.cpu() - pos.float()), dpos_ref))
            e32 = max(rel_err(ho32, ho_ref), rel_err(dp32, dpos_ref))
            if e < thr:
                continue
            print(f"== {gname} reflect={reflect} oa={oa} update={update} L={layers} cut={cut} cutoff={cutoff} scale={scale}: "
                  f"ours {e:.2e}  oracle-fp32 {e32:.2e}")
            N, E, H = n, ei.size(1), 32
            D = 3 * H + 16
            perm = eng.edge_perm
            def rd(name, shape):
                return eng.read(name).view(*shape)
            def edge_rows(x):
                if perm is None:
                    return x
                out = torch.empty_like(x)
                out[perm.cpu()] = x
                return out
            rows = []
            mask_g = eng.read("mask", torch.uint8)
            rows.append(("mask_equal", float((edge_rows(mask_g).double() - dbg["mask"]).abs().max()), 0.0))
            rows.append(("group_equal", float((eng.read("group", torch.int32).double() - dbg["group"].double()).abs().max()), 0.0))
            stages = [("pos_frame", (N, 3), False), ("s0", (N, H), False), ("NE1", (N, 3, H), False), ("e0", (E, D), True),
                      ("nodeframe", (N, 3, 3), False), ("pos_prjt", (N, 3), False)]
            for l in range(layers):
                stages += [(f"s_msg{l}", (N, H), False), (f"vec_msg{l}", (N, 3, H), False), (f"e{l + 1}", (E, D), True)]
                if update:
                    stages += [(f"s{l + 1}", (N, H), False), (f"vec{l + 1}", (N, 3, H), False)]
            for name, shape, is_edge in stages:
                try:
                    x = rd(name, shape)
                except Exception as ex:  # noqa: BLE001
                    rows.append((name, float("nan"), float("nan")))
                    continue
                if is_edge:
                    x = edge_rows(x)
                rows.append((name, rel_err(x, dbg[name]), rel_err(dbg32[name], dbg[name])))
            for name, a, b in rows:
                print(f"   {name:12s} ours {a:.3e}   oracle-fp32 {b:.3e}")
            eng.set_debug(False)


if __name__ == "__main__":
    main()
