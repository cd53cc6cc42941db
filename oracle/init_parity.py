"""ORACLE tooling (build container only): default initialisation of `LEFTNetB200` / `EGNNDynamics` against the UNMODIFIED
reference's constructors at the trained configuration (trainer/train_ts1x.py:43-56).  Training from scratch starts from the
constructor's weights, so the initialisation SCHEME (xavier / kaiming-uniform bounds, zero-filled biases, LayerNorm ones and
zeros, RBF buffers) is part of the drop-in contract even though the random draws differ: per tensor, constants must be equal,
large random tensors must agree in spread and bound.  Prints one JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)

from oa_reactdiff.dynamics import EGNNDynamics as RDyn  # noqa: E402
from oa_reactdiff.model import LEFTNet  # noqa: E402

import oareactdiff_b200 as ob  # noqa: E402

CFG = dict(pos_require_grad=False, cutoff=10.0, num_layers=6, hidden_channels=196, num_radial=96, in_hidden_channels=8,
           reflect_equiv=True, legacy=True, update=True, pos_grad=False, single_layer_output=True, object_aware=True)
KW = dict(fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0, condition_nf=1, pos_dim=3, update_pocket_coords=True,
          condition_time=True, edge_cutoff=None, device=torch.device("cpu"))


def main():
    torch.manual_seed(0)
    ma = RDyn(model_config=dict(CFG), model=LEFTNet, **KW)
    torch.manual_seed(1)
    mb = ob.EGNNDynamics(model_config=dict(CFG), model=ob.LEFTNetB200, **KW)
    a, b = ma.state_dict(), mb.state_dict()
    bad, n_const, n_rand = [], 0, 0
    # parameters() order and requires_grad: optimizer states of a reference checkpoint are stored by POSITION
    pa = [(k, tuple(v.shape), v.requires_grad) for k, v in ma.named_parameters()]
    pb = [(k, tuple(v.shape), v.requires_grad) for k, v in mb.named_parameters()]
    if pa != pb:
        bad.append(("named_parameters order / flags", len(pa), len(pb)))
    if [k for k, _ in ma.named_buffers()] != [k for k, _ in mb.named_buffers()]:
        bad.append(("named_buffers",))
    if list(a) != list(b):
        bad.append(("key order", len(a), len(b)))
    for k in a:
        x, y = a[k].double(), b[k].double()
        if x.shape != y.shape:
            bad.append((k, "shape"))
        elif x.numel() == 1 or float(x.std()) == 0.0 or k.startswith("model.radial_emb."):  # constants and deterministic buffers
            n_const += 1
            if float(x.std()) == 0.0 or k.startswith("model.radial_emb."):
                if not torch.equal(x, y):
                    bad.append((k, "constant differs", float(x.flatten()[0]), float(y.flatten()[0])))
        elif x.numel() >= 1000:
            n_rand += 1
            rs, rm = float(y.std() / x.std()), float(y.abs().max() / x.abs().max())
            if abs(rs - 1) > 0.05 or abs(rm - 1) > 0.05 or abs(float(y.mean())) > 0.05 * float(y.std()) + 1e-3:
                bad.append((k, "spread / bound differs", rs, rm))
    print(json.dumps({"tensors": len(a), "constants_checked": n_const, "random_checked": n_rand, "bad": bad[:10], "n_bad": len(bad)}))


if __name__ == "__main__":
    main()
