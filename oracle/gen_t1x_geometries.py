"""ORACLE tooling (build container only): a small fixture of REAL Transition1x geometries for the bench's replay line.

SURVEY §8d asks for the replay timing "on z_t = alpha_t x + sigma_t eps built from real Transition1x geometries".  The
dataset ships with the reference (oa_reactdiff/data/transition1x/train.pkl) and cannot travel to the GPU box, so the
geometries of the reactions the bench needs are frozen here: for the atom counts `workloads.t1x_sizes(512, seed=0)` draws (the
weak-scaling batch of 8 GPUs x 64; a prefix serves fewer GPUs), in that order, the next unused `use_ind` reaction with that many
atoms — reactant, transition state and product positions (centred per fragment like base_dataset.py:215-218) and the atomic
numbers.  -> tests/golden/t1x_geometries_b512.npz (~300 KB)

    python oracle/gen_t1x_geometries.py
"""
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oareactdiff_b200 import workloads  # noqa: E402

PATH = "/root/reference/oa_reactdiff/data/transition1x/train.pkl"


def main(n_reactions=512):
    raw = pickle.load(open(PATH, "rb"))
    natoms = np.asarray(raw["reactant"]["num_atoms"])
    by_size = {}
    for i in raw["use_ind"]:
        if raw["single_fragment"][i] == 1:
            by_size.setdefault(int(natoms[i]), []).append(int(i))
    used = {k: 0 for k in by_size}
    sizes = workloads.t1x_sizes(n_reactions, seed=0)
    pos = {k: [] for k in ("reactant", "transition_state", "product")}
    Z, picked = [], []
    for n in sizes:
        pool = by_size[n]
        i = pool[used[n] % len(pool)]  # (sizes rarer than their draw count are reused cyclically)
        used[n] += 1
        picked.append(i)
        Z.append(np.asarray(raw["reactant"]["charges"][i][:n], dtype=np.int64))
        for k in pos:
            p = np.asarray(raw[k]["positions"][i][:n], dtype=np.float32)
            pos[k].append(p - p.mean(axis=0, keepdims=True))
    out = os.path.join(ROOT, "tests", "golden", f"t1x_geometries_b{n_reactions}.npz")
    np.savez_compressed(out, sizes=np.asarray(sizes, dtype=np.int64), raw_index=np.asarray(picked, dtype=np.int64),
                        Z=np.concatenate(Z), **{k: np.concatenate(v) for k, v in pos.items()})
    ext = [float(np.abs(p).max()) for p in pos["transition_state"]]
    print(out, os.path.getsize(out), "bytes; reactions", len(sizes), "atoms", int(sum(sizes)), "max |coordinate| %.2f A" % max(ext))


if __name__ == "__main__":
    main()
