"""GPU unit test of the tcgen05 bf16x3 GEMM against an fp64 matmul (tolerance 3e-5 of max|ref|: the split keeps
~16 mantissa bits per operand) and of the exact-fp32 SIMT GEMM (1e-6)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 16, 32), (300, 196, 684), (1000, 684, 196), (257, 588, 588), (5000, 588, 96),
                                   (33, 32, 112), (20000, 196, 684)])
def test_gemm_tc_and_simt(M, N, K):
    from tools.bringup_tc import run
    tc = run(M, N, K, 1)
    simt = run(M, N, K, 0)
    tcs = run(M, N, K, 1, act=1)
    print(M, N, K, "tc", tc, "simt", simt, "tc+silu", tcs)
    for res, tol in ((tc, 3e-5), (tcs, 3e-5), (simt, 2e-6)):
        assert res.startswith("rel_err="), res
        assert float(res.split()[0].split("=")[1]) < tol and res.endswith("nan=0"), res


@pytest.mark.parametrize("M,N,K", [(300, 196, 684), (1000, 684, 196), (257, 588, 588), (5000, 588, 96), (40000, 196, 684),
                                   (30000, 684, 196)])
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_gemm_tc_epilogue_modes(M, N, K, mode):
    """Epilogue modes (two gathered adds / multiplier / in-place residual) incl. multi-tile-per-CTA sizes, on the TMA paths
    and with the LSU fallbacks forced (ablate bits 16 / 32)."""
    from tools.bringup_tc import run_mode
    for ab in (0, 16, 32, 48):
        res = run_mode(M, N, K, mode, ab)
        assert res.startswith("rel_err=") and float(res.split()[0].split("=")[1]) < 3e-5 and res.endswith("nan=0"), (ab, res)


@pytest.mark.parametrize("M,N,K", [(128, 16, 32), (300, 196, 684), (1000, 684, 196), (257, 588, 588), (5000, 588, 96),
                                   (33, 32, 112), (20000, 196, 684), (30000, 684, 196)])
@pytest.mark.parametrize("mode,out_pair", [(0, 1), (0, 0), (1, 1), (2, 0), (3, 1), (3, 0)])
def test_gemm_p16(M, N, K, mode, out_pair):
    """pair16 GEMM (A operand and optionally C / residual / compact copy stored as split-bf16 pairs in the UMMA operand
    layout) against fp64: 3e-5 of max|ref| like the fp32-A kernel (the operands carry the same 16 mantissa bits; a
    pair16 output adds one 2^-17 rounding)."""
    from tools.bringup_p16 import run_p16
    for ew in ((8, 16) if (mode, out_pair) != (3, 0) else (8,)):
        res = run_p16(M, N, K, mode, out_pair, act=1, c2=(mode in (0, 3)), ew=ew)
        assert "error" not in res, res
        assert res["nan"] == 0 and res["rel_err"] < 3e-5 and res.get("rel_err_c2", 0.0) < 3e-5, (ew, res)


@pytest.mark.parametrize("M,N,K", [(128, 16, 32), (300, 196, 684), (257, 588, 588), (1000, 684, 196), (40000, 196, 684)])
@pytest.mark.parametrize("mode,out_pair,ew", [(0, 1, 16), (0, 0, 16), (1, 1, 16), (2, 0, 8), (3, 1, 16)])
def test_gemm_p16_cta_pairs(M, N, K, mode, out_pair, ew):
    """The same kernel on CTA pairs (tcgen05 cta_group::2: 256-row tiles, each CTA holds half of every weight slab, the leader
    issues the MMAs for both) forced for every shape, incl. ragged last tiles whose second CTA has no valid row (ew + 200)."""
    from tools.bringup_p16 import run_p16
    res = run_p16(M, N, K, mode, out_pair, act=1, c2=(mode in (0, 3)), ew=ew + 200)
    assert "error" not in res, res
    assert res["nan"] == 0 and res["rel_err"] < 3e-5 and res.get("rel_err_c2", 0.0) < 3e-5, res
