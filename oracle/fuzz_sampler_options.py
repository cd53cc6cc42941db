"""ORACLE tooling (test infrastructure, build container only): `EnVariationalDiffusion.sample / inpaint` of this package
against the UNMODIFIED reference over their option space — pos_only, return_frames, timesteps override, fixed_idx; RePaint
resamplings x jump_length, frag_fixed — with the NATIVE random streams (same seed on both sides: the order and number of
draws is part of the check) and both host formulations of this package (tabulated fast path / reference-structured).  Both
sides evaluate the denoiser with the reference's own fp32 LEFTNet, so what is compared is the host logic.  One JSON line."""
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)

from oa_reactdiff.diffusion._normalizer import Normalizer as RN  # noqa: E402
from oa_reactdiff.diffusion._schedule import DiffSchedule as RDS, PredefinedNoiseSchedule as RPS  # noqa: E402
from oa_reactdiff.diffusion.en_diffusion import EnVariationalDiffusion as RDiff  # noqa: E402
from oa_reactdiff.dynamics import EGNNDynamics as RDyn  # noqa: E402
from oa_reactdiff.model import LEFTNet as RLeft  # noqa: E402

import oareactdiff_b200 as ob  # noqa: E402
from oracle import oa_ref  # noqa: E402

from oracle.ref_engine import install  # noqa: E402

install()
CFG = dict(cutoff=5.0, num_layers=2, hidden_channels=32, num_radial=16, in_hidden_channels=8, reflect_equiv=True, legacy=True,
           update=True, object_aware=True)
SEED, SIZES, T = 9, [4, 6, 3], 12
SD = oa_ref.make_state_dict(oa_ref.dynamics_param_shapes(CFG, [9, 9, 9], 1), SEED, CFG, prefix_model="model.")


def build(ref, **kw):
    Dyn, Left, DS, PS, N, Diff = (RDyn, RLeft, RDS, RPS, RN, RDiff) if ref else (
        ob.EGNNDynamics, ob.LEFTNetB200, ob.DiffSchedule, ob.PredefinedNoiseSchedule, ob.Normalizer, ob.EnVariationalDiffusion)
    dyn = Dyn(model_config=dict(CFG), fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0, condition_nf=1, model=Left,
              device=torch.device("cpu"))
    dyn.load_state_dict(SD, strict=True)
    return Diff(dynamics=dyn, schdule=DS(PS("polynomial_2", T, 1e-5), (1.0, 1.0, 1.0)), normalizer=N(), **kw)


def flat(v):
    if torch.is_tensor(v):
        return [v]
    return [t for u in v for t in flat(u)]


def compare(a, b):
    xs, ys = flat(a), flat(b)
    if len(xs) != len(ys):
        return float("inf")
    worst = 0.0
    for x, y in zip(xs, ys):
        if x.shape != y.shape:
            return float("inf")
        if x.numel():
            x, y = x.double(), y.double()
            worst = max(worst, float((x - y).abs().max() / x.abs().max().clamp(min=1e-12)))
    return worst


def main():
    nodes, h0, cond = oa_ref.synthetic_batch(len(SIZES), SIZES, SEED)
    g = torch.Generator().manual_seed(SEED + 1)
    xh_fixed = [torch.cat([torch.randn(h.size(0), 3, generator=g) * 1.5, h.float()], dim=1) for h in h0]
    report = []

    def both(kw, call):
        outs = {}
        for side in ("ref", "fast", "structured"):
            d = build(side == "ref", **kw)
            if side == "structured":
                d._fast_ok = lambda: False
            torch.manual_seed(21)
            with torch.no_grad():
                outs[side] = call(d)
        return max(compare(outs["ref"], outs["fast"]), compare(outs["ref"], outs["structured"]))

    for pos_only, frames, steps, fixed in itertools.product([True, False], [1, 3], [None, 6], [None, [0]]):
        kw = dict(pos_only=pos_only, fixed_idx=fixed)
        e = both(kw, lambda d: d.sample(len(SIZES), nodes, cond, return_frames=frames, timesteps=steps,
                                        h0=[h.clone() for h in h0] if pos_only else None))
        report.append({"api": "sample", "pos_only": pos_only, "return_frames": frames, "timesteps": steps, "fixed_idx": fixed, "worst_rel": e})
    for pos_only, (r, j), frag_fixed, steps in itertools.product([True, False], [(1, 1), (2, 3), (3, 2)], [[0, 2], [1]], [None, 6]):
        kw = dict(pos_only=pos_only)
        e = both(kw, lambda d: d.inpaint(len(SIZES), nodes, cond, return_frames=1, resamplings=r, jump_length=j, timesteps=steps,
                                         xh_fixed=[x.clone() for x in xh_fixed], frag_fixed=frag_fixed))
        report.append({"api": "inpaint", "pos_only": pos_only, "resamplings": r, "jump_length": j, "frag_fixed": frag_fixed,
                       "timesteps": steps, "worst_rel": e})
    bad = [r for r in report if not r["worst_rel"] < 1e-4]
    print(json.dumps({"cases": len(report), "bad": len(bad), "worst": max(r["worst_rel"] for r in report), "bad_cases": bad[:8]}))


if __name__ == "__main__":
    main()
