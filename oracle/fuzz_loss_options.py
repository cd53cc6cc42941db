"""ORACLE tooling (test infrastructure, build container only): `EnVariationalDiffusion.forward` (the loss / NLL terms,
en_diffusion.py:56-248, 340-454) of this package against the UNMODIFIED reference over its option space — loss_type
{l2, vlb} x pos_only x training / eval x fixed_idx — with the NATIVE random streams (same seed on both sides, so the order
and number of draws is part of the check).  To isolate the host logic both sides evaluate the denoiser with the reference's
own fp32 LEFTNet: behind `LEFTNetB200.forward` here, natively there.  Prints one JSON line."""
import copy
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT,"oracle","shims")); sys.path.insert(1,"/root/reference"); sys.path.insert(2,ROOT)
from oa_reactdiff.dynamics import EGNNDynamics as RDyn
from oa_reactdiff.model import LEFTNet as RLeft
from oa_reactdiff.diffusion._schedule import DiffSchedule as RDS, PredefinedNoiseSchedule as RPS
from oa_reactdiff.diffusion._normalizer import Normalizer as RN
from oa_reactdiff.diffusion.en_diffusion import EnVariationalDiffusion as RDiff
import oareactdiff_b200 as ob
from oracle import oa_ref
from oracle.ref_engine import install
install()
cfg = dict(cutoff=5.0, num_layers=2, hidden_channels=32, num_radial=16, in_hidden_channels=8, reflect_equiv=True, legacy=True, update=True, object_aware=True)
seed=5; sizes=[4,6,3]
sd = oa_ref.make_state_dict(oa_ref.dynamics_param_shapes(cfg,[9,9,9],1), seed, cfg, prefix_model="model.")
def build(ref, **kw):
    if ref:
        dyn = RDyn(model_config=dict(cfg), fragment_names=["R","TS","P"], node_nfs=[9,9,9], edge_nf=0, condition_nf=1, model=RLeft, device=torch.device("cpu"))
        dyn.load_state_dict(sd, strict=True)
        return RDiff(dynamics=dyn, schdule=RDS(RPS("polynomial_2",20,1e-5),(1.,1.,1.)), normalizer=RN(), **kw)
    dyn = ob.EGNNDynamics(model_config=dict(cfg), fragment_names=["R","TS","P"], node_nfs=[9,9,9], edge_nf=0, condition_nf=1, model=ob.LEFTNetB200, device=torch.device("cpu"))
    dyn.load_state_dict(sd, strict=True)
    return ob.EnVariationalDiffusion(dynamics=dyn, schdule=ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2",20,1e-5),(1.,1.,1.)), normalizer=ob.Normalizer(), **kw)
nodes,h0,cond = oa_ref.synthetic_batch(len(sizes), sizes, seed)
g=torch.Generator().manual_seed(1)
reps=[{"size":nodes[f].clone(),"pos":oa_ref.remove_mean_batch(torch.randn(h0[f].size(0),3,generator=g), oa_ref.get_mask_for_frag(nodes[f])),
       "one_hot":h0[f][:,:5].float(),"charge":h0[f][:,5:].float(),"mask":oa_ref.get_mask_for_frag(nodes[f])} for f in range(3)]
bad=0
report=[]
for loss_type, pos_only, training, fixed in itertools.product(["l2","vlb"],[True,False],[True,False],[None,[0,2]]):
    kw=dict(loss_type=loss_type,pos_only=pos_only,fixed_idx=fixed)
    outs=[]
    for ref in (True,False):
        d=build(ref,**kw); d.train(training)
        torch.manual_seed(11)
        try:
            with torch.no_grad():
                outs.append(d.forward(copy.deepcopy(reps), cond.clone(), return_pred=False))
        except Exception as e:
            outs.append(e)
    a,b=outs
    if isinstance(a,Exception) or isinstance(b,Exception):
        same = type(a)==type(b)
        report.append({"cfg": kw, "training": training, "exception": [repr(a)[:80], repr(b)[:80]]})
        bad += (not same)
        continue
    if set(a)!=set(b): report.append({"cfg": kw, "training": training, "keys": sorted(set(a)^set(b))}); bad+=1; continue
    worst=0
    def flat(v):
        if torch.is_tensor(v): return [v]
        if isinstance(v,(list,tuple)): return [t for u in v for t in flat(u)]
        return [torch.as_tensor(v)]
    for k in a:
        xs,ys=flat(a[k]),flat(b[k])
        if len(xs)!=len(ys): bad+=1; continue
        for x,y in zip(xs,ys):
            x=x.double(); y=y.double()
            if x.shape!=y.shape: bad+=1; continue
            if x.numel()==0: continue
            e=float((x-y).abs().max()/x.abs().max().clamp(min=1e-12)); worst=max(worst,e)
    report.append({"cfg": kw, "training": training, "worst_rel": worst})
    bad += worst>1e-5
print(json.dumps({"cases": len(report), "bad": int(bad), "worst": max(r.get("worst_rel", 0.0) for r in report), "report": report}))
