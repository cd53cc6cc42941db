"""ORACLE tooling (build container only): this package's packed `ProcessedTS1x` against the UNMODIFIED reference's on the
REAL Transition1x training file shipped with the reference (oa_reactdiff/data/transition1x/train.pkl, 10 073 reactions), with
the trainer's options (trainer/train_ts1x.py:76-96): length, random items, random collated batches (reference `collate_fn`
vs the packed `batch()` producer), bit for bit.  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)

from oa_reactdiff.dataset.transition1x import ProcessedTS1x as Ref  # noqa: E402

import oareactdiff_b200 as ob  # noqa: E402

PATH = "/root/reference/oa_reactdiff/data/transition1x/train.pkl"


def main():
    bad, checks, lens = [], 0, {}
    rng = np.random.RandomState(0)
    for name, kw in (("trainer", dict(single_frag_only=True, swapping_react_prod=True, use_by_ind=True)),
                     ("plain", dict(single_frag_only=False, swapping_react_prod=False, use_by_ind=False))):
        a, b = Ref(PATH, **kw), ob.ProcessedTS1x(PATH, **kw)
        lens[name] = [len(a), len(b)]
        if len(a) != len(b):
            bad.append((name, "len"))
            continue
        for i in rng.randint(0, len(a), size=150).tolist():
            x, y = a[i], b[i]
            checks += 1
            if set(x) != set(y) or any(x[k].dtype != y[k].dtype or not torch.equal(x[k], y[k]) for k in x):
                bad.append((name, "item", i))
        for _ in range(10):
            idxs = rng.randint(0, len(a), size=int(rng.randint(1, 20))).tolist()
            (ra, ca), (rb, cb) = Ref.collate_fn([a[i] for i in idxs]), b.batch(idxs)
            checks += 1
            ok = torch.equal(ca, cb) and all(set(u) == set(v) and all(u[k].dtype == v[k].dtype and torch.equal(u[k], v[k]) for k in u)
                                            for u, v in zip(ra, rb))
            if not ok:
                bad.append((name, "batch", idxs[:4]))
    print(json.dumps({"checks": checks, "lens": lens, "n_mismatches": len(bad), "mismatches": bad[:8]}))


if __name__ == "__main__":
    main()
