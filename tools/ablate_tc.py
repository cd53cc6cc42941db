"""Ablation timing of the tcgen05 GEMM (not a pytest file): which role bounds the pipeline?"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oareactdiff_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")

def run(M, N, K, mode, act, ablate, reps=10):
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    aux = torch.randn(M, 2 * N if mode == 1 else N, device=dev) if mode else None
    Cm = torch.empty(M, N, device=dev); ms = C.c_float()
    rc = lib.oard_test_gemm_ex(0, M, N, K, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(b.data_ptr()),
                               C.c_void_p(Cm.data_ptr()), 1, act, 0, mode, C.c_void_p(aux.data_ptr()) if mode else None,
                               ablate, reps, C.byref(ms), None)
    assert rc == 0, lib.oard_last_error()
    return ms.value

if __name__ == "__main__":
    E = 107790; EA = 34188
    shapes = [("edge1", E, 196, 684, 1, 1), ("edge2", E, 196, 196, 0, 1), ("edge_out", E, 684, 196, 3, 1),
              ("dir0", EA, 588, 684, 0, 1), ("dir2", EA, 588, 588, 2, 0), ("node", 2613, 196, 196, 0, 1)]
    names = {0: "full", 16: "full_lsuA", 1: "noAld", 2: "noEpi", 3: "noA+noEpi", 7: "onlyMMA", 15: "empty", 14: "onlyA", 13: "onlyEpi"}
    for nm, M, N, K, mode, act in shapes:
        res = {names[a]: round(run(M, N, K, mode, act, a) * 1e3, 1) for a in names}
        ideal = 2.0 * M * (-(-N // 16) * 16) * (-(-K // 16) * 16) * 3 / 2.25e15 * 1e6
        print(f"{nm:9s} M={M} N={N} K={K}  us: {res}  mma_ideal_us={ideal:.1f}", flush=True)
