"""GPU probe: where the ~14 us of a node-level tcgen05 GEMM (2 613 rows) go.  Prints, for CTA 0 of the last of `reps`
launches, the kernel span (globaltimer) and clock64 marks relative to kernel entry: set-up done, first A box requested /
landed, first converted A stage, first weight slab, last MMA issued, accumulator complete, last store issued, stores
complete, final barrier, exit.  Usage (GPU box): OARD_TC_TS=1 python tools/tc_timeline.py"""
import os
import sys

os.environ["OARD_TC_TS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.bringup_p16 import time_tc  # noqa: E402

NAMES = ["setup", "A_req", "A_land", "A_conv", "W_land", "mma_last", "acc_full", "store_last", "stores_done", "sync", "exit",
         "block_in_regs", "block_staged"]


def main():
    for bn, shapes in (("32", [("node1", 2613, 196, 196, 3), ("x0", 2613, 196, 196, 0), ("node0", 2613, 196, 392, 0)]),
                       ("96", [("x2", 2613, 588, 196, 0)]), ("64", [("PQ", 2613, 392, 196, 0), ("vec_proj", 7839, 392, 196, 0)])):
        os.environ["OARD_TEST_BN"] = bn
        for name, M, N, K, mode in shapes:
            for reps in (1, 20):
                ms = time_tc(M, N, K, mode, reps=reps)
                print(f"{name} BN={bn} reps={reps}: {1e3 * ms:.1f} us per launch (events)", flush=True)
    print("marks:", " ".join(NAMES))


if __name__ == "__main__":
    main()
