"""Bring-up helper (not a pytest file): exercises oard_test_gemm on the tcgen05 path and prints errors."""
import ctypes as C
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oareactdiff_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")


def run(M, N, K, use_tc, act=0, swap=0, bias=True, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = torch.randn(M, K, generator=g).to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev) if bias else None
    Cm = torch.full((M, N), float("nan"), device=dev)
    rc = lib.oard_test_gemm(0, M, N, K, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()),
                            C.c_void_p(b.data_ptr()) if bias else None, C.c_void_p(Cm.data_ptr()), use_tc, act, swap, None)
    torch.cuda.synchronize()
    if rc != 0:
        return f"rc={rc} {lib.oard_last_error().decode()}"
    ref = A.double() @ W.double().T
    if bias:
        ref = ref + b.double()
    if act:
        ref = ref * torch.sigmoid(ref)
    err = (Cm.double() - ref).abs().max() / ref.abs().max()
    return f"rel_err={float(err):.3e} nan={int(torch.isnan(Cm).sum())}"


def run_mode(M, N, K, mode, ablate=0, seed=0):
    """Numerical check of epilogue modes 1 (two row adds), 2 (multiply), 3 (residual, in place) against fp64."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = torch.randn(M, K, generator=g).to(dev)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    aux = torch.randn(M, 2 * N if mode == 1 else N, generator=g).to(dev)
    Cm = aux.clone() if mode == 3 else torch.full((M, N), float("nan"), device=dev)
    auxp = Cm if mode == 3 else aux  # mode 3: residual aliases the output (in place), like the edge-state update
    ms = C.c_float()
    rc = lib.oard_test_gemm_ex(0, M, N, K, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(b.data_ptr()),
                               C.c_void_p(Cm.data_ptr()), 1, 1, 0, mode, C.c_void_p(auxp.data_ptr()), ablate, 1,
                               C.byref(ms), None)
    torch.cuda.synchronize()
    if rc != 0:
        return f"rc={rc} {lib.oard_last_error().decode()}"
    ref = A.double() @ W.double().T + b.double()
    if mode == 1:
        ref = ref + aux[:, :N].double() + aux[:, N:].double()
    ref = ref * torch.sigmoid(ref)
    if mode == 2:
        ref = ref * aux.double()
    if mode == 3:
        ref = ref + aux.double()
    err = (Cm.double() - ref).abs().max() / ref.abs().max()
    return f"rel_err={float(err):.3e} nan={int(torch.isnan(Cm).sum())}"


if __name__ == "__main__":
    swap = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    for (M, N, K) in [(128, 16, 32), (128, 208, 32), (128, 32, 16), (300, 196, 684), (1000, 684, 196), (257, 588, 588),
                      (5000, 588, 96), (40000, 196, 684)]:
        if len(sys.argv) > 2:
            ab = int(sys.argv[2])
            print(f"M={M} N={N} K={K} ablate={ab}: mode1 {run_mode(M, N, K, 1, ab)} | mode2 {run_mode(M, N, K, 2, ab)} | mode3 {run_mode(M, N, K, 3, ab)}", flush=True)
            continue
        print(f"M={M} N={N} K={K} swap={swap}: simt {run(M, N, K, 0)} | tc {run(M, N, K, 1, swap=swap)} | tc+silu {run(M, N, K, 1, act=1, swap=swap)}", flush=True)
