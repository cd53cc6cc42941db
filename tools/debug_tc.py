import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oareactdiff_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")

def pattern(M, N, K, mode, ablate):
    g = torch.Generator().manual_seed(0)
    A = torch.randn(M, K, generator=g).to(dev); W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev); b = torch.randn(N, generator=g).to(dev)
    aux = torch.randn(M, 2 * N if mode == 1 else N, generator=g).to(dev)
    Cm = torch.zeros(M, N, device=dev)
    rc = lib.oard_test_gemm_ex(0, M, N, K, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(b.data_ptr()),
                               C.c_void_p(Cm.data_ptr()), 1, 0, 0, mode, C.c_void_p(aux.data_ptr()), ablate, 1, None, None)
    torch.cuda.synchronize()
    ref = A.double() @ W.double().T + b.double()
    if mode == 1: ref = ref + aux[:, :N].double() + aux[:, N:].double()
    if mode == 2: ref = ref * aux.double()
    if mode == 3: ref = ref + aux.double()
    err = (Cm.double() - ref).abs()
    bad_rows = (err.max(dim=1).values > 1e-3 * ref.abs().max()).nonzero().flatten()
    bad_cols = (err.max(dim=0).values > 1e-3 * ref.abs().max()).nonzero().flatten()
    tiles = sorted(set((bad_rows // 128).tolist()))
    print(f"M={M} N={N} K={K} mode={mode} ablate={ablate}: bad rows {bad_rows.numel()} in {len(tiles)} tiles (first {tiles[:12]}, last {tiles[-5:]}), "
          f"bad cols {bad_cols.numel()} (first {bad_cols[:8].tolist()} last {bad_cols[-4:].tolist()})")
    if bad_rows.numel():
        r = int(bad_rows[0]); print("   row", r, "rows-in-tile bad:", sorted(set((bad_rows[bad_rows // 128 == r // 128] % 128).tolist()))[:40])
        # is the wrong value equal to the value of some other row (stale A)? compare against ref of other tiles
        diff = Cm[r].double() - ref[r]
        print("   sample err", diff[:6].tolist())
for extra in (0, 64, 128, 192):
    print("== extra bits", extra)
    pattern(40000, 196, 684, 1, 32 | extra)
    pattern(300, 196, 684, 2, 0 | extra)
