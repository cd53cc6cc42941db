"""The differentiable training core (csrc/train_core.h: forward with saved activations + hand-derived backward) in its
HOST-EMULATION build (csrc/train_emu.cpp, plain g++): the same source the device runs, with `par_for` as a loop.
Checked here on the CPU against the fp64 oracle: forward values, the gradient w.r.t. the node-feature input and the
gradient of every parameter (torch autograd through oracle/oa_ref.py::leftnet_forward, itself pinned against the
unmodified reference's autograd in tests/test_oracle_grad.py).  Tolerance: 2e-4 of max|grad| per parameter (fp32 core vs fp64)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import oa_ref

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "oareactdiff_b200", "csrc")
LIB = os.path.join(os.path.dirname(HERE), "oareactdiff_b200", "libtrain_emu.so")


@pytest.fixture(scope="module")
def emu():
    src = [os.path.join(CSRC, f) for f in ("train_emu.cpp", "train_core.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(s) for s in src):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, src[0]], check=True)
    lib = C.CDLL(LIB)
    lib.emu_get_grad.restype = C.c_int
    lib.emu_get_act.restype = C.c_int
    return lib


def _p(t):
    return C.c_void_p(t.data_ptr())


def _case(cfg, sizes, seed, pos_scale):
    """Inputs + oracle geometry (fp64) for one batch; -> dict of torch tensors."""
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, seed)
    masks = [oa_ref.get_mask_for_frag(n) for n in nodes]
    cm = torch.cat(masks)
    ei = oa_ref.get_edges_index(cm, remove_self_edge=True)
    nfs = oa_ref.get_n_frag_switch(nodes)
    sub = oa_ref.get_subgraph_mask(ei, nfs)
    g = torch.Generator().manual_seed(seed)
    N = cm.numel()
    pos = torch.randn(N, 3, generator=g, dtype=torch.float64) * pos_scale
    h = torch.randn(N, cfg["in_hidden_channels"], generator=g, dtype=torch.float64)
    shapes = oa_ref.leftnet_param_shapes(cfg)
    sd = oa_ref.make_state_dict(shapes, seed, cfg, dtype=torch.float64)
    return dict(cfg=cfg, ei=ei, sub=sub, pos=pos, h=h, sd=sd, N=N)


def _geometry(case):
    """The non-differentiable inputs of the core, from the oracle's own functions (leftnet.py:747-834)."""
    cfg, ei, pos, sd, N = case["cfg"], case["ei"], case["pos"], case["sd"], case["N"]
    dbg = {}
    with torch.no_grad():
        oa_ref.leftnet_forward(sd, cfg, case["h"], pos, ei, case["sub"][:, None], dbg=dbg)
    pf, mask = dbg["pos_frame"], dbg["mask"].unsqueeze(-1)
    dist, cdiff, ccross, cvert = oa_ref.scalarization(pf, ei)
    dist = dist * mask.squeeze(-1)
    frame = torch.stack((cdiff * mask, ccross * mask, cvert * mask), dim=1)  # [E, 3(k), 3(xyz)]
    rb = 0.5 * (torch.cos(dist * math.pi / float(cfg["cutoff"])) + 1.0)
    deg = torch.zeros(N, dtype=torch.float64).index_add_(0, ei[0], torch.ones(ei.size(1), dtype=torch.float64))
    act = torch.nonzero(dbg["mask"] > 0).flatten().to(torch.int32).contiguous()
    return dict(frame=frame, rb=rb, rbf=dbg["rbf"], inv_deg=1.0 / deg.clamp(min=1), nodeframe=dbg["nodeframe"], pos_prjt=dbg["pos_prjt"],
                act=act, dbg=dbg)


@pytest.mark.parametrize("cfg_name,sizes,pos_scale", [("small", [4, 3], 1.5), ("small", [5, 2, 3], 3.0), ("mid", [4, 3], 1.5),
                                                      ("small_noreflect", [4, 3], 1.5)])
def test_core_forward_and_backward_vs_oracle(emu, cfg_name, sizes, pos_scale):
    cfg = dict(oa_ref.TRAINED_CFG, hidden_channels=32, num_radial=16, num_layers=2, cutoff=5.0)
    if cfg_name == "small_noreflect":  # reflect_equiv=False: no abs on the frame's cross row, cross term in the messages
        cfg["reflect_equiv"] = False
    if cfg_name == "mid":
        cfg = dict(oa_ref.TRAINED_CFG, hidden_channels=52, num_radial=24, num_layers=3, cutoff=6.0)
    case = _case(cfg, sizes, seed=7 + len(sizes), pos_scale=pos_scale)
    geo = _geometry(case)
    sd, N, E = case["sd"], case["N"], case["ei"].size(1)
    H, R, Cin, L = cfg["hidden_channels"], cfg["num_radial"], cfg["in_hidden_channels"], cfg["num_layers"]
    # ---- oracle: values and gradients of a random linear functional of the outputs (fp64 autograd)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if not k.startswith("radial_emb.")}
    sd_g = {**sd, **params}
    h = case["h"].clone().requires_grad_(True)
    h_out, dpos = oa_ref.leftnet_forward(sd_g, cfg, h, case["pos"], case["ei"], case["sub"][:, None])
    gen = torch.Generator().manual_seed(99)
    A = torch.randn(h_out.shape, generator=gen, dtype=torch.float64)
    B = torch.randn(dpos.shape, generator=gen, dtype=torch.float64)
    ((h_out * A).sum() + (dpos * B).sum()).backward()
    # ---- emulated core (fp32)
    f32 = lambda t: t.detach().to(torch.float32).contiguous()
    emu.emu_reset(N, E, H, R, Cin, L, int(cfg["reflect_equiv"]))
    w32 = {k: f32(v) for k, v in sd.items() if not k.startswith(("radial_emb.", "distance_embedding", "last_layer"))}
    for k, v in w32.items():
        emu.emu_set_weight(k.encode(), _p(v), C.c_long(v.numel()))
    ei32, ej32 = case["ei"][0].to(torch.int32).contiguous(), case["ei"][1].to(torch.int32).contiguous()
    t = {k: f32(geo[k]) for k in ("frame", "rb", "rbf", "inv_deg", "nodeframe", "pos_prjt")}
    h32, A32, B32 = f32(case["h"]), f32(A), f32(B)
    ho, dp, gh = torch.zeros(N, Cin), torch.zeros(N, 3), torch.zeros(N, Cin)
    emu.emu_forward_backward(_p(ei32), _p(ej32), _p(t["frame"]), _p(t["rb"]), _p(t["rbf"]), _p(t["inv_deg"]), _p(t["nodeframe"]),
                             _p(t["pos_prjt"]), _p(geo["act"]), C.c_int(geo["act"].numel()), _p(h32), _p(ho), _p(dp), _p(A32), _p(B32), _p(gh))
    assert 0 < geo["act"].numel() < E  # both active and masked edges are exercised
    rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max().clamp(min=1e-30))
    e_h, e_p = rel(ho, h_out.detach()), rel(dp, dpos.detach())
    print(f"forward: h_out {e_h:.2e} dpos {e_p:.2e}")
    assert e_h < 2e-5 and e_p < 2e-5
    e_gh = rel(gh, h.grad)
    worst, worst_name, n_checked = e_gh, "h_in", 0
    for k, v in w32.items():
        ref = params[k].grad
        out = torch.zeros_like(v)
        assert emu.emu_get_grad(k.encode(), _p(out), C.c_long(out.numel())) == 0, k
        if ref is None or float(ref.abs().max()) == 0.0:
            assert float(out.abs().max()) == 0.0, k
            continue
        n_checked += 1
        e = rel(out, ref)
        if e > worst:
            worst, worst_name = e, k
    print(f"backward: d/dh_in {e_gh:.2e}; worst parameter gradient {worst:.2e} ({worst_name}) over {n_checked} parameters")
    assert worst < 2e-4 and n_checked > 40
