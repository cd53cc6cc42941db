"""ORACLE tooling (test infrastructure, build container only): this package's `EGNNDynamics` (constructor + torch-composed
`forward`) against the UNMODIFIED reference's over its option space — condition_time, condition_nf, 1-D / per-sample t,
equal / per-fragment node_nf, 2 or 3 fragments, an empty fragment, enforce_same_encoding, a `source` hand-off of weights
(dynamics/_base.py:21-132, egnn_dynamics.py:63-182).  Both sides evaluate the denoiser with the reference's own fp32
LEFTNet, so what is compared is the wrapper: encoders, time / condition channels, fragment slicing, centre-of-mass removal,
decoders.  Prints one JSON line."""
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)

from oa_reactdiff.dynamics import EGNNDynamics as RDyn  # noqa: E402
from oa_reactdiff.model import LEFTNet as RLeft  # noqa: E402
from oa_reactdiff.utils import get_edges_index, get_mask_for_frag, get_n_frag_switch  # noqa: E402

import oareactdiff_b200 as ob  # noqa: E402

from oracle.ref_engine import install  # noqa: E402

install()


def main():
    report = []
    g = torch.Generator().manual_seed(0)
    frag_sets = {"3frag_ragged_empty": ([torch.tensor([2, 0]), torch.tensor([2, 3]), torch.tensor([1, 2])], ["a", "b", "c"]),
                 "2frag": ([torch.tensor([4]), torch.tensor([5])], ["A", "B"]),
                 "3frag_equal": ([torch.tensor([3, 5, 4])] * 3, ["R", "TS", "P"])}
    for (fs_name, (frags, names)), cond_time, cond_nf, t_kind, per_frag_nf, same_enc in itertools.product(
            frag_sets.items(), [True, False], [0, 1, 3], ["1d", "per_sample"], [False, True], [None, "share"]):
        if not cond_time and cond_nf == 0:
            continue  # (the reference slices h_final[:, :-0]: an empty tensor; nobody builds that)
        k = len(frags)
        node_nfs = [4 + i for i in range(k)] if per_frag_nf else [6] * k
        in_hidden = 4 + int(cond_time) + cond_nf
        cfg = dict(pos_require_grad=False, cutoff=5.0, num_layers=2, hidden_channels=32, num_radial=8, in_hidden_channels=in_hidden)
        enforce = list(range(1, k)) if (same_enc and not per_frag_nf) else None
        kw = dict(node_nfs=node_nfs, edge_nf=0, condition_nf=cond_nf, fragment_names=names, pos_dim=3, update_pocket_coords=True,
                  condition_time=cond_time, edge_cutoff=None, enforce_same_encoding=enforce)
        torch.manual_seed(3)
        ref = RDyn(model_config=dict(cfg), model=RLeft, device=torch.device("cpu"), **kw)
        src = {"model": ref.model.state_dict(), "encoders": ref.encoders.state_dict(), "decoders": ref.decoders.state_dict()}
        ours = ob.EGNNDynamics(model_config=dict(cfg), model=ob.LEFTNetB200, device=torch.device("cpu"), source=src, **kw)
        assert set(ours.state_dict()) == set(ref.state_dict())
        B = frags[0].numel()
        masks = [get_mask_for_frag(n) for n in frags]
        cm = torch.cat(masks)
        ei = get_edges_index(cm, remove_self_edge=True)
        nfs = get_n_frag_switch(frags)
        xh = [torch.randn(int(frags[i].sum()), node_nfs[i], generator=g) for i in range(k)]
        t = torch.tensor([0.314]) if t_kind == "1d" else torch.rand(B, 1, generator=g)
        cond = torch.rand(B, max(cond_nf, 1), generator=g)[:, :cond_nf] if cond_nf else torch.zeros(B, 0)
        with torch.no_grad():
            a, ea = ref.forward([x.clone() for x in xh], ei, t, cond, nfs, cm, edge_attr=None)
            b, eb = ours.forward([x.clone() for x in xh], ei, t, cond, nfs, cm, edge_attr=None)
        worst = 0.0 if (ea is None and eb is None) else float("inf")
        for x, y in zip(a, b):
            if x.shape != y.shape:
                worst = float("inf")
            elif x.numel():
                worst = max(worst, float((x - y).abs().max() / x.abs().max().clamp(min=1e-12)))
        report.append({"frags": fs_name, "condition_time": cond_time, "condition_nf": cond_nf, "t": t_kind,
                       "per_fragment_nf": per_frag_nf, "shared": enforce, "worst_rel": worst})
    bad = [r for r in report if not r["worst_rel"] < 1e-5]
    print(json.dumps({"cases": len(report), "bad": len(bad), "worst": max(r["worst_rel"] for r in report), "bad_cases": bad[:8]}))


if __name__ == "__main__":
    main()
