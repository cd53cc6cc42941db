"""Input/output producers either side of the hot path, with the reference's names and results:

* `ProcessedTS1x` / `BaseDataset.collate_fn`  — oa_reactdiff/dataset/transition1x.py:21-150, base_dataset.py:54-88, 148-218
* `assemble_sample_inputs`, `write_single_xyz`, `write_tmp_xyz` — oa_reactdiff/utils/sampling_tools.py:64-149

The reference keeps one small tensor per reaction per property (5 x 3 x n_reactions Python objects) and collates a batch
with ~15 `torch.cat` over Python lists.  Here the whole dataset is packed ONCE into flat arrays (CSR over reactions:
offsets, atomic numbers, centred positions per fragment), `__getitem__` / `collate_fn` reproduce the reference's
per-sample dicts and batch bit-for-bit, and `batch()` builds the same collated batch with a handful of vectorised
gathers into pinned staging buffers and one host->device copy per tensor — the producer that feeds `compute_loss`,
`sample()` and `inpaint()` without per-sample Python work.
"""
import pickle
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor
from torch.utils.data import Dataset

ATOM_MAPPING = {1: 0, 6: 1, 7: 2, 8: 3, 9: 4}  # base_dataset.py:8-15
n_element = len(ATOM_MAPPING)
FRAG_KEYS = ("reactant", "transition_state", "product")
FRAG_MAPPING = {"reactant": "product", "transition_state": "transition_state", "product": "reactant"}  # transition1x.py:8-12


def _load_raw(path_or_dict):
    if isinstance(path_or_dict, dict):
        return path_or_dict
    p = str(path_or_dict)
    if ".npz" in p:
        with np.load(p, allow_pickle=True) as f:
            return {k: v for k, v in f.items()}
    if ".pkl" in p:
        with open(p, "rb") as fh:
            return pickle.load(fh)
    raise ValueError("data file should be either .npz or .pkl")  # base_dataset.py:37


class ProcessedTS1x(Dataset):
    """Transition1x reactions (reactant, transition state, product share the atom list).  Constructor arguments follow
    transition1x.py:22-39; `npz_path` may also be the already-loaded raw dict.  Options outside the training / sampling
    path (`remove_h`, `append_frag`, `reflection`, `only_ts`, `confidence_model`, `ediff`, `pad_fragments`) raise."""

    def __init__(self, npz_path, center=True, pad_fragments=0, device="cpu", zero_charge=False, remove_h=False,
                 single_frag_only=True, swapping_react_prod=False, append_frag=False, reflection=False, use_by_ind=False,
                 only_ts=False, confidence_model=False, position_key="positions", ediff=None, **kwargs):
        super().__init__()
        unsupported = dict(pad_fragments=pad_fragments, append_frag=append_frag, reflection=reflection, only_ts=only_ts,
                           confidence_model=confidence_model, ediff=ediff is not None)
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(f"ProcessedTS1x (B200 producer): unsupported options {bad}")
        raw = _load_raw(npz_path)
        self.center, self.zero_charge, self.device = center, zero_charge, torch.device(device)
        self.n_fragment = self.n_fragments = 3
        n_raw = len(raw["single_fragment"])
        # reaction selection, in the reference's order: list(set(single_frag) & set(use_inds))  (transition1x.py:52-64)
        single = np.where(np.array(raw["single_fragment"]) == 1)[0] if single_frag_only else np.arange(n_raw)
        use = raw["use_ind"] if use_by_ind else range(n_raw)
        sel = list(set(single).intersection(set(use)))
        order: List[Tuple[str, int]] = []  # (source fragment key, raw index) per output reaction, per fragment
        per_frag = {}
        for k in FRAG_KEYS:
            lst = [(k, int(i)) for i in sel]
            if swapping_react_prod:  # the swapped copies are appended after the originals (transition1x.py:66-74)
                lst += [(FRAG_MAPPING[k], int(i)) for i in sel]
            per_frag[k] = lst
        self.n_samples = len(per_frag["reactant"])
        # ---- pack: sizes / offsets shared by the three fragments, atomic numbers and centred positions per fragment
        sizes = np.array([int(raw[src]["num_atoms"][i]) for src, i in per_frag["reactant"]], dtype=np.int64)
        self.sizes = torch.from_numpy(sizes)
        self.offsets = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64))
        tot = int(sizes.sum())
        self.Z = torch.empty(3, tot, dtype=torch.int64)
        self.pos = torch.empty(3, tot, 3, dtype=torch.float32)
        for f, k in enumerate(FRAG_KEYS):
            pkey = position_key if k != "transition_state" else "positions"  # transition1x.py:105-110
            o = 0
            for src, i in per_frag[k]:
                n = int(raw[src]["num_atoms"][i])
                z = np.asarray(raw[src]["charges"][i][:n])
                p = torch.tensor(np.asarray(raw[src][pkey][i])[:n], dtype=torch.float32)
                if center:
                    p = p - torch.mean(p, dim=0)  # base_dataset.py:215-218
                self.Z[f, o:o + n] = torch.from_numpy(z.astype(np.int64))
                self.pos[f, o:o + n] = p
                o += n
        # dtype of the per-atom "charge" feature: the reference builds it with torch.tensor(raw charges) (base_dataset.py:179-185),
        # so it follows the raw container — int32 for the numpy arrays of the shipped Transition1x files, int64 for Python lists
        self.charge_dtype = torch.int64
        if self.n_samples:
            src, i = per_frag["reactant"][0]
            self.charge_dtype = torch.tensor(raw[src]["charges"][i][:int(raw[src]["num_atoms"][i])]).dtype
        lut = torch.full((int(self.Z.max()) + 1,), -1, dtype=torch.int64)
        for z, c in ATOM_MAPPING.items():
            if z < lut.numel():
                lut[z] = c
        self.cls = lut[self.Z]
        if bool((self.cls < 0).any()):
            raise KeyError("atomic number outside ATOM_MAPPING {1, 6, 7, 8, 9}")
        self._pin: Dict[str, Tensor] = {}

    # ---------------------------------------------------------------- reference-compatible per-sample access
    def __len__(self):
        return self.n_samples

    def __getitem__(self, idx) -> Dict[str, Tensor]:
        a, b = int(self.offsets[idx]), int(self.offsets[idx + 1])
        out = {}
        for f in range(3):
            out[f"size_{f}"] = self.sizes[idx].to(self.device)
            out[f"pos_{f}"] = self.pos[f, a:b].to(self.device)
            out[f"one_hot_{f}"] = F.one_hot(self.cls[f, a:b], num_classes=n_element).to(self.device)
            ch = torch.zeros(b - a, 1, dtype=torch.int64) if self.zero_charge else self.Z[f, a:b].view(-1, 1).to(self.charge_dtype)
            out[f"charge_{f}"] = ch.to(self.device)
            out[f"mask_{f}"] = torch.zeros(b - a, dtype=torch.int64, device=self.device)
        out["condition"] = torch.zeros(1, 1, dtype=torch.int64, device=self.device)
        return out

    @staticmethod
    def collate_fn(batch):
        """base_dataset.py:54-88: list of per-sample dicts -> ([{size, pos, one_hot, charge, mask} per fragment], condition)."""
        n_fragment = len([k for k in batch[0].keys() if "size" in k])
        out = [{} for _ in range(n_fragment)]
        res = {}
        for prop in batch[0].keys():
            if prop in ["condition", "target", "rmsd", "ediff"]:
                res[prop] = torch.cat([x[prop] for x in batch], dim=0)
                continue
            idx = int(prop.split("_")[-1])
            _prop = prop.replace(f"_{idx}", "")
            if "size" in prop:
                out[idx][_prop] = torch.tensor([x[prop] for x in batch], device=batch[0][prop].device)
            elif "mask" in prop:
                out[idx][_prop] = torch.cat([i * torch.ones(len(x[prop]), device=x[prop].device).long()
                                             for i, x in enumerate(batch)], dim=0)
            else:
                out[idx][_prop] = torch.cat([x[prop] for x in batch], dim=0)
        if len(res) == 1:
            return out, res["condition"]
        return out, res

    # ---------------------------------------------------------------- packed batch producer
    def _staging(self, name: str, shape, dtype) -> Tensor:
        n = int(np.prod(shape))
        buf = self._pin.get(name)
        if buf is None or buf.numel() < n or buf.dtype != dtype:
            buf = torch.empty(max(n, 1), dtype=dtype)
            if torch.cuda.is_available():
                buf = buf.pin_memory()
            self._pin[name] = buf
        return buf[:n].view(shape)

    def batch(self, indices: Sequence[int], device=None):
        """Same result as `collate_fn([self[i] for i in indices])`, built from the packed arrays: one gather index for
        the whole batch, pinned staging, one H2D copy per tensor."""
        device = self.device if device is None else torch.device(device)
        idx = torch.as_tensor(list(indices), dtype=torch.int64)
        sz = self.sizes[idx]
        B, n_tot = idx.numel(), int(sz.sum())
        mask = torch.repeat_interleave(torch.arange(B, dtype=torch.int64), sz)
        start = torch.cumsum(sz, 0) - sz
        rows = self.offsets[idx][mask] + (torch.arange(n_tot, dtype=torch.int64) - start[mask])  # packed row of every batch atom
        nb = device.type == "cuda"
        out = []
        for f in range(3):
            pos = self._staging(f"pos{f}", (n_tot, 3), torch.float32)
            torch.index_select(self.pos[f], 0, rows, out=pos)
            oh = self._staging(f"oh{f}", (n_tot, n_element), torch.int64)
            oh.zero_()
            oh.scatter_(1, self.cls[f][rows].view(-1, 1), 1)
            ch = self._staging(f"ch{f}", (n_tot, 1), torch.int64)
            if self.zero_charge:
                ch.zero_()
            else:
                torch.index_select(self.Z[f], 0, rows, out=ch.view(-1))
            if not self.zero_charge and self.charge_dtype != torch.int64:
                ch = ch.to(self.charge_dtype)
            out.append({"size": sz.to(device, non_blocking=nb), "pos": pos.to(device, non_blocking=nb),
                        "one_hot": oh.to(device, non_blocking=nb), "charge": ch.to(device, non_blocking=nb),
                        "mask": mask.to(device, non_blocking=nb)})
        if nb:
            torch.cuda.current_stream(device).synchronize()  # the staging buffers are reused by the next batch
        return out, torch.zeros(B, 1, dtype=torch.int64, device=device)


_DECODER = {"H": [1, 0, 0, 0, 0, 1], "C": [0, 1, 0, 0, 0, 6], "N": [0, 0, 1, 0, 0, 7], "O": [0, 0, 0, 1, 0, 8],
            "F": [0, 0, 0, 0, 1, 9]}


def assemble_sample_inputs(atoms: List, device: torch.device = torch.device("cuda"), n_samples: int = 1,
                           frag_type: bool = False) -> List[Tensor]:
    """h0 = [one-hot(5) | atomic number (| fragment type)] per fragment for `sample()` (sampling_tools.py:64-108)."""
    h0 = []
    for ii in range(3):
        rows = [_DECODER[a] + ([ii % 2] if frag_type else []) for a in atoms]
        h0.append(torch.tensor(rows, device=device).repeat(n_samples, 1))
    return h0


_C2A = {1: "H", 6: "C", 7: "N", 8: "O", 9: "F"}


def write_single_xyz(xyzfile, natoms, out):
    """sampling_tools.py:111-126 (same text, `str(float32)` coordinates)."""
    rows = out[:, :3 + 5 + 1].detach().cpu()
    xyz, z = rows[:, :3].numpy(), rows[:, -1].long().tolist()
    with open(xyzfile, "w") as fo:
        fo.write(str(natoms) + "\n\n")
        for a, x in zip(z, xyz):
            fo.write(f"{_C2A[a]} " + " ".join(str(v) for v in x) + "\n")


def write_tmp_xyz(fragments_nodes, out_samples, idx=[0], prefix="gen", localpath="tmp", ex_ind=0):
    """sampling_tools.py:129-149."""
    typemap = {0: "react", 1: "ts", 2: "prod"}
    for ii in idx:
        start = 0
        for jj, natoms in enumerate(fragments_nodes[0]):
            n = int(natoms)
            write_single_xyz(f"{localpath}/{prefix}_{jj + ex_ind}_{typemap[ii]}.xyz", n, out_samples[ii][start:start + n])
            start += n


# ------------------------------------------------------------------------------------------------ evaluation-side callers
# oa_reactdiff/evaluate/utils.py:14-63, 91-110 — the three helpers the reference's evaluation scripts put between a batch of
# the dataset and `inpaint()`.  `ddpm_trainer` is whatever carries the diffusion model as `.ddpm` (the reference's
# LightningModule); a bare `EnVariationalDiffusion` is accepted as well.
def _ddpm_of(obj):
    return getattr(obj, "ddpm", obj)


def set_new_schedule(ddpm_trainer, timesteps: int = 250, device: torch.device = torch.device("cuda"),
                     noise_schedule: str = "polynomial_2"):
    """Replace the noise schedule of a built model — sampling with fewer steps than it was trained with
    (evaluate/utils.py:14-32; precision 1e-5, the model's own norm_values)."""
    from .schedule import DiffSchedule, PredefinedNoiseSchedule
    ddpm = _ddpm_of(ddpm_trainer)
    ddpm.schedule = DiffSchedule(gamma_module=PredefinedNoiseSchedule(noise_schedule=noise_schedule, timesteps=timesteps,
                                                                       precision=1e-5), norm_values=ddpm.norm_values)
    ddpm.T = timesteps
    return ddpm_trainer.to(device)


def inplaint_batch(batch: List, ddpm_trainer, resamplings: int = 1, jump_length: int = 1, frag_fixed: List = [0, 2]):
    """One collated batch -> RePaint inpainting with the fragments `frag_fixed` clamped to the batch's own geometry
    (evaluate/utils.py:35-63; the name keeps the reference's spelling).  Returns (out_samples[0], xh_fixed, fragments_nodes)."""
    from .normalizer import FEATURE_MAPPING
    representations, conditions = batch
    xh_fixed = [torch.cat([rep[k] for k in FEATURE_MAPPING], dim=1) for rep in representations]
    fragments_nodes = [rep["size"] for rep in representations]
    out_samples, _ = _ddpm_of(ddpm_trainer).inpaint(
        n_samples=representations[0]["size"].size(0), fragments_nodes=fragments_nodes, conditions=conditions, return_frames=1,
        resamplings=resamplings, jump_length=jump_length, timesteps=None, xh_fixed=xh_fixed, frag_fixed=frag_fixed)
    return out_samples[0], xh_fixed, fragments_nodes


def samples_to_pos_charge(out_samples, fragments_nodes):
    """Per-reaction numpy positions of the three fragments, atomic numbers and atom counts from `out_samples[0]`
    (evaluate/utils.py:91-110; all three fragments are split with the first fragment's sizes, as in the reference)."""
    cuts = torch.cumsum(fragments_nodes[0], dim=0).to("cpu")[:-1]
    parts = [torch.tensor_split(out_samples[ii], cuts) for ii in range(3)]
    pos = {name: [x[:, :3].cpu().numpy() for x in parts[ii]] for ii, name in enumerate(FRAG_KEYS)}
    z = [x[:, -1].long().cpu().numpy() for x in parts[0]]
    natoms = [f.cpu().item() for f in fragments_nodes[0]]
    return pos, z, natoms
