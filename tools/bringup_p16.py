"""Bring-up / timing helper (not a pytest file) for the pair16 tcgen05 GEMM (csrc/gemm_p16.cuh) through
oard_test_gemm_p16.  `python tools/bringup_p16.py` prints accuracy for every mode and a timing table next to the
fp32-A kernel (gemm_tc.cuh) on the B=64 edge shapes."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oareactdiff_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def run_p16(M, N, K, mode=0, out_pair=1, act=1, c2=False, ew=0, reps=1, seed=0):
    """-> dict(rel_err, rel_err_c2, nan, ms).  Reference: fp64 on the fp32 inputs (reps must be 1 for accuracy)."""
    if reps > 1:  # timing only: draw on the device (the CPU generator takes seconds for the edge-level shapes)
        g = torch.Generator(device=dev).manual_seed(seed)
        A = torch.randn(M, K, generator=g, device=dev)
        W = torch.randn(N, K, generator=g, device=dev) / K ** 0.5
        b = torch.randn(N, generator=g, device=dev)
        aux = torch.randn(M, 2 * N if mode == 1 else N, generator=g, device=dev) if mode else None
    else:
        g = torch.Generator(device="cpu").manual_seed(seed)
        A = torch.randn(M, K, generator=g).to(dev)
        W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
        b = torch.randn(N, generator=g).to(dev)
        aux = torch.randn(M, 2 * N if mode == 1 else N, generator=g).to(dev) if mode else None
    Cm = torch.full((M, N), float("nan"), device=dev)
    M3 = (M + 2) // 3
    C2 = torch.full((M3, N), float("nan"), device=dev) if c2 else None
    ms = C.c_float()
    rc = lib.oard_test_gemm_p16(0, M, N, K, _p(A), _p(W), _p(b), _p(Cm), mode, _p(aux), out_pair, act, _p(C2), ew, reps,
                                C.byref(ms), None)
    torch.cuda.synchronize()
    if rc != 0:
        return dict(error=f"rc={rc} {lib.oard_last_error().decode()}")
    out = dict(ms=ms.value)
    if reps == 1:
        ref = A.double() @ W.double().T + b.double()
        if mode == 1:
            ref = ref + aux[:, :N].double() + aux[:, N:].double()
        if act:
            ref = ref * torch.sigmoid(ref)
        if mode == 2:
            ref = ref * aux.double()
        if mode == 3:
            ref = ref + aux.double()
        out["rel_err"] = float((Cm.double() - ref).abs().max() / ref.abs().max())
        out["nan"] = int(torch.isnan(Cm).sum())
        if c2:
            out["rel_err_c2"] = float((C2.double() - ref[0::3]).abs().max() / ref.abs().max())
            out["nan"] += int(torch.isnan(C2).sum())
    return out


def time_tc(M, N, K, mode, reps=10):
    g = torch.Generator(device="cpu").manual_seed(0)
    A = torch.randn(M, K, generator=g).to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    aux = torch.randn(M, 2 * N if mode == 1 else N, generator=g).to(dev) if mode else None
    Cm = aux.clone() if mode == 3 else torch.empty(M, N, device=dev)
    ms = C.c_float()
    rc = lib.oard_test_gemm_ex(0, M, N, K, _p(A), _p(W), _p(b), _p(Cm), 1, 1, 0, mode, _p(Cm if mode == 3 else aux), 0, reps,
                               C.byref(ms), None)
    torch.cuda.synchronize()
    return ms.value if rc == 0 else float("nan")


if __name__ == "__main__":
    for (M, N, K) in [(128, 16, 32), (300, 196, 684), (1000, 684, 196), (257, 588, 588), (20000, 196, 684)]:
        for mode, op in [(0, 1), (0, 0), (1, 1), (2, 0), (3, 1)]:
            for ew in (8, 16):
                print(M, N, K, "mode", mode, "pair" if op else "fp32", "ew", ew, run_p16(M, N, K, mode, op, 1, c2=(mode == 3), ew=ew),
                      flush=True)
    E, Ea = 107790, 34188
    shapes = [("edge1", E, 196, 684, 1, 1), ("edge2", E, 196, 196, 0, 1), ("edge_out", E, 684, 196, 3, 1),
              ("dir0", Ea, 588, 684, 0, 1), ("dir2", Ea, 588, 588, 2, 0)]
    for name, M, N, K, mode, op in shapes:
        row = {"tc_fp32A_us": round(1e3 * time_tc(M, N, K, mode), 1)}
        for ew in (8, 16):
            row[f"p16_ew{ew}_us"] = round(1e3 * run_p16(M, N, K, mode, op, 1, c2=False, ew=ew, reps=10).get("ms", float("nan")), 1)
        print(name, M, N, K, row, flush=True)
