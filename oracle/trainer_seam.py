"""ORACLE tooling (test infrastructure, build container only): the plugin class inside the reference's REAL trainer module.

`oa_reactdiff/trainer/pl_trainer.py::DDPMModule` is imported unmodified (Lightning, torchmetrics and the pymatgen-based RMSD
tool are absent here; they are satisfied by inert stand-ins: `LightningModule` = `nn.Module` with no-op `log` /
`save_hyperparameters`) and built twice with the configuration of `trainer/train_ts1x.py:43-121` — once with
`model=LEFTNet`, once with `model=LEFTNetB200` and the first one's weights handed over through `source` — on a batch of the
reference's own dataset object.  Checked:
  * `DDPMModule.compute_loss` (pl_trainer.py:208-282), training and evaluation mode: the two modules agree, and this
    package's `EnVariationalDiffusion.compute_loss` reproduces the trainer's numbers over loss_type x pos_only x mode;
  * `training_step` / `validation_step` run on the plugged module; `eval_inplaint_batch`'s sampling part (deep copy of the
    ddpm, schedule swap, RePaint with r = j = 2 on a short schedule) gives the same samples.
Both sides evaluate the denoiser with the reference's own fp32 LEFTNet (behind `LEFTNetB200.forward` on the plugged side):
the SEAM and the host arithmetic are what is compared.  Prints one JSON line."""
import copy
import itertools
import json
import os
import pickle
import sys
import tempfile
import types

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(1, "/root/reference")
sys.path.insert(2, ROOT)


def _stand_ins():
    import pytorch_lightning as pl

    class LightningModule(nn.Module):
        current_epoch = 0

        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    tm, tmc = types.ModuleType("torchmetrics"), types.ModuleType("torchmetrics.classification")
    for n in ("BinaryAccuracy", "BinaryAUROC", "BinaryF1Score", "BinaryPrecision", "BinaryCohenKappa"):
        setattr(tmc, n, object)
    for n in ("PearsonCorrCoef", "SpearmanCorrCoef", "MeanAbsoluteError"):
        setattr(tm, n, object)
    tm.classification = tmc
    sys.modules["torchmetrics"], sys.modules["torchmetrics.classification"] = tm, tmc
    an, rm = types.ModuleType("oa_reactdiff.analyze"), types.ModuleType("oa_reactdiff.analyze.rmsd")
    an.__path__ = []
    captured = {}

    def batch_rmsd(fragments_nodes, out_samples, xh_fixed, idx=1, threshold=0.5):  # keeps the samples, returns a dummy metric
        captured["out"] = [o.clone() for o in out_samples]
        return [0.0]
    rm.batch_rmsd = batch_rmsd
    sys.modules["oa_reactdiff.analyze"], sys.modules["oa_reactdiff.analyze.rmsd"] = an, rm
    return captured


CAPTURED = _stand_ins()

from oa_reactdiff.model import LEFTNet  # noqa: E402
from oa_reactdiff.trainer.pl_trainer import DDPMModule  # noqa: E402
from oa_reactdiff.dataset.transition1x import ProcessedTS1x  # noqa: E402

import oareactdiff_b200 as ob  # noqa: E402
from oracle.gen_golden import synthetic_raw_dataset  # noqa: E402  (the raw-dataset maker of the fixtures; imports the reference too)

from oracle.ref_engine import install  # noqa: E402

install()

# trainer/train_ts1x.py:43-121, with a narrower network so the check runs in seconds on the CPU
LEFTNET_CONFIG = dict(pos_require_grad=False, cutoff=10.0, num_layers=2, hidden_channels=32, num_radial=16, in_hidden_channels=8,
                      reflect_equiv=True, legacy=True, update=True, pos_grad=False, single_layer_output=True, object_aware=True)
OPTIMIZER_CONFIG = dict(lr=2.5e-4, betas=[0.9, 0.999], weight_decay=0, amsgrad=True)
TRAINING_CONFIG = dict(datadir="unused", remove_h=False, bz=14, num_workers=0, clip_grad=True, gradient_clip_val=None, ema=False,
                       ema_decay=0.999, swapping_react_prod=True, append_frag=False, use_by_ind=True, reflection=False,
                       single_frag_only=True, only_ts=False, lr_schedule_type=None, lr_schedule_config=dict(gamma=0.8, step_size=100))


def module(model, loss_type, pos_only, source=None, timesteps=5000):
    return DDPMModule(dict(LEFTNET_CONFIG), dict(OPTIMIZER_CONFIG), copy.deepcopy(TRAINING_CONFIG), node_nfs=[9] * 3, edge_nf=0,
                      condition_nf=1, fragment_names=["R", "TS", "P"], pos_dim=3, update_pocket_coords=True, condition_time=True,
                      edge_cutoff=None, norm_values=(1.0, 1.0, 1.0), norm_biases=(0.0, 0.0, 0.0), noise_schedule="cosine",
                      timesteps=timesteps, precision=1e-5, loss_type=loss_type, pos_only=pos_only, process_type="TS1x", model=model,
                      enforce_same_encoding=None, scales=[1.0, 2.0, 1.0], source=source, fixed_idx=None, eval_epochs=10)


def rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / a.abs().max().clamp(min=1e-12))


def main():
    raw = synthetic_raw_dataset()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "raw.pkl")
        pickle.dump(raw, open(path, "wb"))
        ds = ProcessedTS1x(path, **TRAINING_CONFIG)  # the trainer's own call shape: func(path, **training_config) (pl_trainer.py:160-163)
        ours_ds = ob.ProcessedTS1x(path, **TRAINING_CONFIG)
    batch = ProcessedTS1x.collate_fn([ds[i] for i in (0, 3, 5, 7)])
    ours_batch = ours_ds.batch([0, 3, 5, 7])
    report, worst = [], 0.0
    for loss_type, pos_only, training in itertools.product(["l2", "vlb"], [True, False], [True, False]):
        torch.manual_seed(1)
        a = module(LEFTNet, loss_type, pos_only)
        src = {"model": a.ddpm.dynamics.model.state_dict(), "encoders": a.ddpm.dynamics.encoders.state_dict(),
               "decoders": a.ddpm.dynamics.decoders.state_dict()}
        b = module(ob.LEFTNetB200, loss_type, pos_only, source=src)
        assert type(b.ddpm.dynamics.model).__name__ == "LEFTNetB200"
        assert set(a.state_dict()) == set(b.state_dict())  # what a Lightning checkpoint of either module holds
        # this package's own stack with the same weights
        dyn = ob.EGNNDynamics(model_config=dict(LEFTNET_CONFIG), fragment_names=["R", "TS", "P"], node_nfs=[9] * 3, edge_nf=0,
                              condition_nf=1, model=ob.LEFTNetB200, device=torch.device("cpu"), source=src)
        mine = ob.EnVariationalDiffusion(dynamics=dyn, schdule=ob.DiffSchedule(ob.PredefinedNoiseSchedule("cosine", 5000, 1e-5), (1.0, 1.0, 1.0)),
                                         normalizer=ob.Normalizer(), loss_type=loss_type, pos_only=pos_only)
        outs = []
        for m in (a, b):
            m.train(training)
            torch.manual_seed(7)
            with torch.no_grad():
                outs.append(m.compute_loss(copy.deepcopy(batch)))
        mine.train(training)
        torch.manual_seed(7)
        with torch.no_grad():
            nll_mine, info_mine = mine.compute_loss(copy.deepcopy(ours_batch), scales=(1.0, 2.0, 1.0), training=training)
        e_plug = rel(outs[0][0], outs[1][0])
        e_mine = rel(outs[0][0], nll_mine)
        e_info = max(abs(float(outs[0][1][k]) - float(info_mine[k])) / max(abs(float(outs[0][1][k])), 1e-12) for k in outs[0][1])
        report.append({"loss_type": loss_type, "pos_only": pos_only, "training": training, "plugged_vs_reference": e_plug,
                       "package_vs_reference": e_mine, "info_rel": float(e_info), "info_keys_equal": set(outs[0][1]) == set(info_mine),
                       "nll_reference": [float(v) for v in outs[0][0]], "finite": bool(torch.isfinite(outs[0][0]).all())})
        worst = max(worst, e_plug, e_mine, e_info)
    # the step functions of the trainer on the plugged module, and the sampling half of eval_inplaint_batch
    torch.manual_seed(1)
    a = module(LEFTNet, "l2", True)
    src = {"model": a.ddpm.dynamics.model.state_dict(), "encoders": a.ddpm.dynamics.encoders.state_dict(),
           "decoders": a.ddpm.dynamics.decoders.state_dict()}
    b = module(ob.LEFTNetB200, "l2", True, source=src)
    steps = {}
    for name, m in (("reference", a), ("plugged", b)):
        m.trainer = types.SimpleNamespace(is_global_zero=True)
        m.train(True)
        torch.manual_seed(9)
        with torch.no_grad():
            tr = m.training_step(copy.deepcopy(batch), 1)
        m.train(False)
        torch.manual_seed(9)
        with torch.no_grad():
            va = m.validation_step(copy.deepcopy(batch), 1)
        m.sampling_schedule = type(m.sampling_schedule)(gamma_module=type(m.sampling_schedule.gamma_module)("polynomial_2", 8, 1e-5),
                                                        norm_values=(1.0, 1.0, 1.0))  # 8 steps instead of 150: seconds, not minutes
        torch.manual_seed(9)
        m.eval_inplaint_batch(copy.deepcopy(batch), resamplings=2, jump_length=2, frag_fixed=[0, 2])
        steps[name] = (float(tr["loss"]), float(va["val-totloss"]), [o.clone() for o in CAPTURED["out"]])
    e_tr = abs(steps["reference"][0] - steps["plugged"][0]) / abs(steps["reference"][0])
    e_va = abs(steps["reference"][1] - steps["plugged"][1]) / abs(steps["reference"][1])
    e_smp = max(rel(x[:, :3], y[:, :3]) for x, y in zip(steps["reference"][2], steps["plugged"][2]))
    worst = max(worst, e_tr, e_va, e_smp)
    print(json.dumps({"compute_loss_cases": len(report), "worst": float(worst), "training_step": float(e_tr),
                      "validation_step": float(e_va), "eval_inpaint_samples": float(e_smp), "report": report}))


if __name__ == "__main__":
    main()
