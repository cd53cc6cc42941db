"""Input/output producers against golden outputs of the UNMODIFIED reference
(oracle/gen_golden.py::case_dataset): `ProcessedTS1x.__getitem__`, `collate_fn`, the packed `batch()` producer,
`assemble_sample_inputs`, `write_tmp_xyz`.  Bit-exact (integer / index work and float32 copies)."""
import copy
import json
import os

import numpy as np
import pytest
import torch

from oareactdiff_b200 import data as D

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset_small.npz")


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLDEN, allow_pickle=False)
    return {k: z[k] for k in z.files}


def _variants(gold):
    return json.loads(str(gold["variants"])), json.loads(str(gold["raw"]))


@pytest.mark.parametrize("vn", ["plain", "swap_useind", "all_zero"])
def test_dataset_items_and_collate_match_reference(gold, vn):
    variants, raw = _variants(gold)
    ds = D.ProcessedTS1x(copy.deepcopy(raw), **variants[vn])
    assert len(ds) == int(gold[f"{vn}/len"])
    item = ds[1]
    for k, v in item.items():
        ref = gold[f"{vn}/item1/{k}"]
        assert v.numpy().dtype == ref.dtype and np.array_equal(v.numpy(), ref), (vn, k)
    idxs = [int(i) for i in gold[f"{vn}/idxs"]]
    for producer in ("collate", "packed"):
        reps, cond = D.ProcessedTS1x.collate_fn([ds[i] for i in idxs]) if producer == "collate" else ds.batch(idxs)
        assert np.array_equal(cond.numpy(), gold[f"{vn}/cond"]) and cond.dtype == torch.int64
        assert len(reps) == 3
        for f, r in enumerate(reps):
            assert set(r) == {"size", "pos", "one_hot", "charge", "mask"}
            for k, v in r.items():
                ref = gold[f"{vn}/{k}{f}"]
                assert v.numpy().dtype == ref.dtype, (producer, vn, k, v.dtype, ref.dtype)
                assert np.array_equal(v.numpy(), ref), (producer, vn, k, f)


def test_dataset_edge_cases(gold):
    _, raw = _variants(gold)
    ds = D.ProcessedTS1x(copy.deepcopy(raw))
    reps, cond = ds.batch([])  # empty batch
    assert cond.shape == (0, 1) and all(r["pos"].shape == (0, 3) and r["size"].numel() == 0 for r in reps)
    reps, _ = ds.batch([3, 3])  # repeated reaction
    assert reps[0]["mask"].tolist() == [0] * int(ds.sizes[3]) + [1] * int(ds.sizes[3])
    with pytest.raises(NotImplementedError):
        D.ProcessedTS1x(copy.deepcopy(raw), only_ts=True)
    bad = copy.deepcopy(raw)
    bad["reactant"]["charges"][0][0] = 16  # sulphur is not in ATOM_MAPPING (base_dataset.py:8-15)
    with pytest.raises(KeyError):
        D.ProcessedTS1x(bad)
    with pytest.raises(ValueError):
        D.ProcessedTS1x("reactions.txt")


def test_sampling_tools_match_reference(gold, tmp_path):
    atoms = ["C", "H", "H", "O", "N", "F"]
    for ft in (False, True):
        h0 = D.assemble_sample_inputs(atoms, device=torch.device("cpu"), n_samples=2, frag_type=ft)
        for f in range(3):
            ref = gold[f"h0_ft{int(ft)}_{f}"]
            assert h0[f].numpy().dtype == ref.dtype and np.array_equal(h0[f].numpy(), ref)
    nodes = [torch.tensor([2, 4])] * 3
    samples = [torch.from_numpy(gold[f"xyz_in{f}"]) for f in range(3)]
    D.write_tmp_xyz(nodes, samples, idx=[0, 1, 2], prefix="gen", localpath=str(tmp_path), ex_ind=3)
    names = sorted(k[4:] for k in gold if k.startswith("xyz/"))
    assert sorted(os.listdir(tmp_path)) == names and len(names) == 6
    for fn in names:
        assert open(tmp_path / fn).read() == str(gold["xyz/" + fn]), fn


def test_reference_checkpoint_layout_loads(tmp_path):
    """A DDPMModule-style Lightning checkpoint of the UNMODIFIED reference (`ddpm.` prefix; fixture from
    oracle/gen_golden.py::case_checkpoint) loads key for key into this package's EnVariationalDiffusion."""
    import oareactdiff_b200 as ob
    z = np.load(os.path.join(os.path.dirname(GOLDEN), "checkpoint_small.npz"), allow_pickle=False)
    cfg = json.loads(str(z["cfg"]))
    sd = {k: torch.from_numpy(z[k]) for k in z.files if k not in ("cfg", "T")}
    dyn = ob.EGNNDynamics(model_config=cfg, fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0, condition_nf=1,
                          model=ob.LEFTNetB200, device=torch.device("cpu"))
    sched = ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", int(z["T"]), 1e-5), norm_values=(1.0, 1.0, 1.0))
    ddpm = ob.EnVariationalDiffusion(dynamics=dyn, schdule=sched, normalizer=ob.Normalizer(), pos_only=True)
    path = str(tmp_path / "ckpt.pt")
    torch.save({"state_dict": sd, "epoch": 3, "global_step": 10}, path)
    rep = ob.load_reference_checkpoint(ddpm, path)
    assert rep["missing"] == [] and rep["unexpected"] == [] and rep["ignored"] == ["confidence.head.weight"]
    own = ddpm.state_dict()
    assert rep["loaded"] == len(own) == len(sd) - 1
    for k, v in own.items():
        assert torch.equal(v, sd["ddpm." + k]), k
    # bare state dict (no prefix) and strictness
    rep2 = ob.load_reference_checkpoint(ddpm, {k[5:]: v for k, v in sd.items() if k.startswith("ddpm.")})
    assert rep2["loaded"] == len(own)
    broken = {k: v for k, v in sd.items() if "embedding_out" not in k}
    with pytest.raises(RuntimeError):
        ob.load_reference_checkpoint(ddpm, {"state_dict": broken})
    with pytest.raises(KeyError):
        ob.load_reference_checkpoint(ddpm, {"state_dict": {"foo.bar": torch.zeros(1)}})


def test_lightning_checkpoint_that_pickles_the_reference_model_class(tmp_path):
    """A real checkpoint of the reference carries `hyper_parameters["model"] = <class LEFTNet>` (save_hyperparameters,
    pl_trainer.py:147): a pickled class of a package that is not installed here.  Refused with a pointer by default, read with
    placeholders under allow_pickle=True; only the tensors are used."""
    import sys
    import types

    import oareactdiff_b200 as ob
    z = np.load(os.path.join(os.path.dirname(GOLDEN), "checkpoint_small.npz"), allow_pickle=False)
    cfg, T = json.loads(str(z["cfg"])), int(z["T"])
    sd = {k: torch.from_numpy(z[k]) for k in z.files if k not in ("cfg", "T")}
    mod = types.ModuleType("oa_reactdiff_absent_pkg")

    class LEFTNet:  # stands for oa_reactdiff.model.leftnet.LEFTNet at save time
        pass
    LEFTNet.__module__, LEFTNet.__qualname__ = "oa_reactdiff_absent_pkg", "LEFTNet"
    mod.LEFTNet = LEFTNet
    sys.modules["oa_reactdiff_absent_pkg"] = mod
    path = tmp_path / "ddpm-epoch=1.ckpt"
    try:
        torch.save({"epoch": 1, "global_step": 10, "pytorch-lightning_version": "1.8.6", "state_dict": sd,
                    "hyper_parameters": {"model": LEFTNet, "model_config": cfg, "scales": [1.0, 2.0, 1.0], "probe": LEFTNet()}}, path)
    finally:
        del sys.modules["oa_reactdiff_absent_pkg"]  # ... and is not importable at load time
    dyn = ob.EGNNDynamics(model_config=cfg, fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0, condition_nf=1,
                          model=ob.LEFTNetB200, device=torch.device("cpu"))
    ddpm = ob.EnVariationalDiffusion(dynamics=dyn, schdule=ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", T, 1e-5), (1.0, 1.0, 1.0)),
                                     normalizer=ob.Normalizer(), pos_only=True)
    with pytest.raises(RuntimeError, match="allow_pickle=True"):
        ob.load_reference_checkpoint(ddpm, str(path))
    res = ob.load_reference_checkpoint(ddpm, path, allow_pickle=True)
    assert res["missing"] == [] and res["unexpected"] == [] and res["loaded"] == len(ddpm.state_dict())
    for k, v in ddpm.state_dict().items():
        assert torch.equal(v, sd["ddpm." + k]), k
