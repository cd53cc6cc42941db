"""GPU probe of the CTA-pair (cta_group::2) form of the pair16 GEMM: accuracy on ragged / multi-tile shapes with pairs
forced, then timings single CTAs vs pairs on the B=64 edge shapes.  Usage (GPU box): python tools/pair_probe.py [acc|time]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.bringup_p16 import run_p16  # noqa: E402


def main():
    what = sys.argv[1:] or ["acc", "time"]
    if "acc" in what:
        for (M, N, K) in [(128, 16, 32), (300, 196, 684), (1000, 684, 196), (257, 588, 588), (40000, 196, 684), (50000, 684, 196)]:
            for mode, op, ew in [(0, 1, 16), (0, 0, 16), (1, 1, 16), (2, 0, 8), (3, 1, 16)]:
                r = run_p16(M, N, K, mode, op, 1, c2=(mode in (0, 3)), ew=ew + 200)
                print(M, N, K, "mode", mode, "pair" if op else "fp32", r, flush=True)
    if "ablate" in what:
        # role ablation (results meaningless): 1 no A loads, 4 no weight loads, 2 no epilogue traffic, 8 no MMAs
        E, Ea = 107790, 34188
        shapes = [("dir0", Ea, 588, 684, 0, 1, 16), ("dir2", Ea, 588, 588, 2, 0, 8), ("edge1_plain", E, 196, 684, 0, 1, 16),
                  ("edge2", E, 196, 196, 0, 1, 16), ("edge_out", E, 684, 196, 3, 1, 16)]
        for name, M, N, K, mode, op, ew in shapes:
            for ct in ((1, 2) if K >= 512 else (1,)):  # (with CTA pairs the no-MMA ablations, and any ablation of the in-place residual mode, never finish)
                row = {}
                for ab in ((0, 1, 4, 5, 2, 7, 15, 13, 8) if ct == 1 else (0, 1, 4, 5, 2, 7)):
                    os.environ["OARD_P16_ABLATE"] = str(ab)
                    row[ab] = round(1e3 * run_p16(M, N, K, mode, op, 1, c2=False, ew=ew + 100 * ct, reps=20).get("ms", float("nan")), 1)
                os.environ["OARD_P16_ABLATE"] = "0"
                print(name, "ctas", ct, "us by ablation bits:", row, flush=True)
    if "timeline" in what:
        # clock64 marks of CTA 0 around its 4th tile, single CTAs.  MMA thread: 0 tile start, 1 accumulator free, 2 first weight
        # slab, 3 first A box, 4 last MMA issued, 5 next tile start; epilogue warp 0: 6 waits for the accumulator, 7 has it,
        # 8 first block read from tensor memory, 9 its aux box there, 10 math + staging done, 11 store issued, 12 tile done;
        # A loader: 13 first / 14 last box of the tile requested.
        os.environ["OARD_P16_TS"] = "1"
        E, Ea = 107790, 34188
        for name, M, N, K, mode, op, ew in [("edge_out", E, 684, 196, 3, 1, 16), ("edge1_plain", E, 196, 684, 0, 1, 16)]:
            for ab in (0, 15, 7, 8):
                os.environ["OARD_P16_ABLATE"] = str(ab)
                r = run_p16(M, N, K, mode, op, 1, c2=False, ew=ew + 100, reps=5)
                print(name, "ablate", ab, round(1e3 * r.get("ms", float("nan")), 1), "us", flush=True)
        os.environ["OARD_P16_ABLATE"] = "0"
        del os.environ["OARD_P16_TS"]
    if "time" in what:
        E, Ea = 107790, 34188
        shapes = [("edge1", E, 196, 684, 1, 1, 16), ("edge2", E, 196, 196, 0, 1, 16), ("edge_out", E, 684, 196, 3, 1, 16),
                  ("dir0", Ea, 588, 684, 0, 1, 16), ("dir2", Ea, 588, 588, 2, 0, 8)]
        for name, M, N, K, mode, op, ew in shapes:
            row = {}
            for ct in (1, 2, 1, 2):
                row.setdefault(f"ctas{ct}_us", []).append(
                    round(1e3 * run_p16(M, N, K, mode, op, 1, c2=False, ew=ew + 100 * ct, reps=20).get("ms", float("nan")), 1))
            print(name, M, N, K, row, flush=True)


if __name__ == "__main__":
    main()
