"""Loading the reference's checkpoints.  `DDPMModule` (trainer/pl_trainer.py:77-140) keeps the diffusion
model as `self.ddpm`, so a Lightning checkpoint is `{"state_dict": {"ddpm.<key>": tensor, ...}, "epoch": ..., ...}` with
`<key>` = the state-dict keys of `EnVariationalDiffusion` — which this repo's modules reproduce one to one
(`dynamics.model.*`, `dynamics.encoders.*`, `dynamics.decoders.*`, `schedule.gamma_module.gamma`).
`demo.py:269` (`DDPMModule.load_from_checkpoint`) is therefore replaced by `load_reference_checkpoint(ddpm, path)`."""
import os
import pickle
import types
from typing import Dict, Union

import torch
from torch import Tensor, nn


class _Absent:
    """Stands in for a pickled object whose class cannot be imported here (see _read)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        pass

    def __call__(self, *a, **k):
        return self


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except Exception:
            return _Absent


_tolerant_pickle = types.SimpleNamespace(Unpickler=_TolerantUnpickler, load=pickle.load, loads=pickle.loads,
                                         __name__="pickle")


def _read(path: str, allow_pickle: bool):
    """`DDPMModule.save_hyperparameters()` (pl_trainer.py:147) puts the constructor arguments into the checkpoint, among them
    `model=<class LEFTNet>`: a Lightning checkpoint of the reference therefore pickles a CLASS of the reference package and
    does not open with `weights_only=True`, nor at all where `oa_reactdiff` is not installed.  With `allow_pickle=True` (a
    file you trust: unpickling can run code) it is read with an unpickler that replaces what cannot be imported by a
    placeholder; only the tensors of `state_dict` are used afterwards."""
    try:
        return torch.load(path, map_location="cpu", weights_only=True)
    except Exception as e:
        if not allow_pickle:
            raise RuntimeError(f"{path}: not loadable with weights_only=True ({type(e).__name__}: {str(e)[:200]}).  Lightning "
                               "checkpoints of the reference pickle its model class in `hyper_parameters`; if you trust the file, "
                               "call load_reference_checkpoint(..., allow_pickle=True).") from e
    return torch.load(path, map_location="cpu", weights_only=False, pickle_module=_tolerant_pickle)


def load_reference_checkpoint(ddpm: nn.Module, ckpt: Union[str, Dict], prefix: str = "ddpm.", strict: bool = True,
                              allow_pickle: bool = False) -> Dict:
    """Copy the weights of a reference checkpoint (path or already-loaded dict; plain state dict or Lightning layout) into
    `ddpm` (an `EnVariationalDiffusion` of this package).  Returns {"loaded": n, "ignored": [keys outside `prefix`],
    "missing": [...], "unexpected": [...]}; with strict=True missing / unexpected keys under the prefix raise like
    `load_state_dict` does.  `allow_pickle`: see `_read`."""
    if isinstance(ckpt, (str, os.PathLike)):
        ckpt = _read(os.fspath(ckpt), allow_pickle)
    sd = ckpt.get("state_dict", ckpt) if isinstance(ckpt, dict) else ckpt
    if not isinstance(sd, dict) or not all(isinstance(v, Tensor) for v in sd.values()):
        raise ValueError("checkpoint does not hold a state dict of tensors")
    own = ddpm.state_dict()
    if prefix and not any(k.startswith(prefix) for k in sd):
        if any(k in own for k in sd):
            prefix = ""  # a bare EnVariationalDiffusion state dict
        else:
            raise KeyError(f"no key starts with '{prefix}' and none matches this module: is this an OA-ReactDiff checkpoint?")
    picked = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    ignored = sorted(k for k in sd if not k.startswith(prefix))
    missing, unexpected = sorted(k for k in own if k not in picked), sorted(k for k in picked if k not in own)
    if strict and (missing or unexpected):  # validated BEFORE anything is copied: a failed strict load leaves the module untouched
        raise RuntimeError(f"checkpoint does not match the module: missing {missing[:5]}{'...' if len(missing) > 5 else ''}, "
                           f"unexpected {unexpected[:5]}{'...' if len(unexpected) > 5 else ''}")
    ddpm.load_state_dict(picked, strict=False)
    return {"loaded": len(picked) - len(unexpected), "ignored": ignored, "missing": missing, "unexpected": unexpected}
