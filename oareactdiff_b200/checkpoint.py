"""Loading the reference's checkpoints.  `DDPMModule` (trainer/pl_trainer.py:77-140) keeps the diffusion
model as `self.ddpm`, so a Lightning checkpoint is `{"state_dict": {"ddpm.<key>": tensor, ...}, "epoch": ..., ...}` with
`<key>` = the state-dict keys of `EnVariationalDiffusion` — which this repo's modules reproduce one to one
(`dynamics.model.*`, `dynamics.encoders.*`, `dynamics.decoders.*`, `schedule.gamma_module.gamma`).
`demo.py:269` (`DDPMModule.load_from_checkpoint`) is therefore replaced by `load_reference_checkpoint(ddpm, path)`."""
from typing import Dict, Union

import torch
from torch import Tensor, nn


def load_reference_checkpoint(ddpm: nn.Module, ckpt: Union[str, Dict], prefix: str = "ddpm.", strict: bool = True) -> Dict:
    """Copy the weights of a reference checkpoint (path or already-loaded dict; plain state dict or Lightning layout) into
    `ddpm` (an `EnVariationalDiffusion` of this package).  Returns {"loaded": n, "ignored": [keys outside `prefix`],
    "missing": [...], "unexpected": [...]}; with strict=True missing / unexpected keys under the prefix raise like
    `load_state_dict` does."""
    if isinstance(ckpt, str):
        ckpt = torch.load(ckpt, map_location="cpu", weights_only=True)
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
    res = ddpm.load_state_dict(picked, strict=False)
    missing, unexpected = list(res.missing_keys), list(res.unexpected_keys)
    if strict and (missing or unexpected):
        raise RuntimeError(f"checkpoint does not match the module: missing {missing[:5]}{'...' if len(missing) > 5 else ''}, "
                           f"unexpected {unexpected[:5]}{'...' if len(unexpected) > 5 else ''}")
    return {"loaded": len(picked) - len(unexpected), "ignored": ignored, "missing": missing, "unexpected": unexpected}
