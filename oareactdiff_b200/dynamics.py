"""`EGNNDynamics` with the reference's constructor and forward signature (oa_reactdiff/dynamics/egnn_dynamics.py:13-168,
dynamics/_base.py:8-132), restructured so that one call issues no host<->device synchronisation: the same-fragment
edge mask, fragment slices and per-(fragment, sample) segment ids are cached per graph instead of being rebuilt on
the CPU every step, and the NaN guard is evaluated on the device."""
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .graph_tools import get_subgraph_mask
from .leftnet import LEFTNetB200, _MLP


class _EncDec(_MLP):
    """MLP(in -> 2d -> out), SiLU between, no activation on the last layer (_base.py:91-109)."""

    def forward(self, x: Tensor) -> Tensor:
        return self.mlp[1].linear(F.silu(self.mlp[0].linear(x)))


class BaseDynamics(nn.Module):
    def __init__(self, model_config: Dict, fragment_names: List[str], node_nfs: List[int], edge_nf: int,
                 condition_nf: int = 0, pos_dim: int = 3, update_pocket_coords: bool = True,
                 condition_time: bool = True, edge_cutoff: Optional[float] = None, model: nn.Module = LEFTNetB200,
                 device: torch.device = torch.device("cuda"), enforce_same_encoding: Optional[List] = None,
                 source: Optional[Dict] = None) -> None:
        super().__init__()
        assert len(node_nfs) == len(fragment_names)
        for nf in node_nfs:
            assert nf > pos_dim
        model_config = dict(model_config)
        model_config.setdefault("act_fn", "swish")
        if str(model_config["act_fn"]).lower() not in ("swish", "silu"):  # encoders / decoders and the fused kernels are SiLU
            raise NotImplementedError(f"act_fn={model_config['act_fn']!r}: only 'swish' (SiLU), the trained configuration, is implemented")
        if "in_node_nf" not in model_config:  # (the reference's own test configs give in_node_nf and no in_hidden_channels)
            model_config["in_node_nf"] = model_config["in_hidden_channels"]
        if model_config.get("in_edge_nf", 0) > 0:
            raise NotImplementedError("edge attributes (in_edge_nf > 0) are not part of the LEFTNet hot path")
        self.model_config, self.node_nfs, self.edge_nf, self.condition_nf = model_config, node_nfs, edge_nf, condition_nf
        self.fragment_names, self.pos_dim = fragment_names, pos_dim
        self.update_pocket_coords, self.condition_time, self.edge_cutoff, self.device = (
            update_pocket_coords, condition_time, edge_cutoff, device)
        self.model = (model or LEFTNetB200)(**model_config)
        if source is not None:
            self.model.load_state_dict(source["model"])
        self.dist_dim = getattr(self.model, "dist_dim", 0)
        self.embed_dim = model_config["in_node_nf"] - (1 if condition_time else 0) - max(condition_nf, 0)
        self.edge_embed_dim = 0
        assert self.embed_dim > 0
        self.build_encoders_decoders(enforce_same_encoding, source)

    def build_encoders_decoders(self, enfoce_name_encoding: Optional[List] = None, source: Optional[Dict] = None):
        """Per-fragment encoder d -> 2d -> embed_dim and decoder embed_dim -> 2d -> d (_base.py:82-132; the keyword keeps
        the reference's spelling).  No edge encoder / decoder: `in_edge_nf > 0` is rejected by the constructor."""
        self.encoders, self.decoders = nn.ModuleList(), nn.ModuleList()
        for nf in self.node_nfs:
            d = nf - self.pos_dim
            self.encoders.append(_EncDec(d, [2 * d, self.embed_dim]))
            self.decoders.append(_EncDec(self.embed_dim, [2 * d, d]))
        if enfoce_name_encoding is not None:
            for ii in enfoce_name_encoding:
                self.encoders[ii] = self.encoders[0]
                self.decoders[ii] = self.decoders[0]
        if source is not None:
            self.encoders.load_state_dict(source["encoders"])
            self.decoders.load_state_dict(source["decoders"])
        self.edge_encoder, self.edge_decoder = None, None

    def forward(self):
        raise NotImplementedError


class EGNNDynamics(BaseDynamics):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._graph_key = None
        self._graph = None
        self._nan_gen: Optional[torch.Generator] = None

    def _graph_cache(self, edge_index: Tensor, n_frag_switch: Tensor, combined_mask: Tensor):
        key = (edge_index.data_ptr(), edge_index._version, n_frag_switch.data_ptr(), n_frag_switch._version,
               combined_mask.data_ptr(), combined_mask._version)
        if key != self._graph_key:
            sub = get_subgraph_mask(edge_index, n_frag_switch)
            frag_index = self.compute_frag_index(n_frag_switch)
            n_samples = int(combined_mask.max().item()) + 1 if combined_mask.numel() else 0
            seg = n_frag_switch * n_samples + combined_mask
            n_seg = len(self.fragment_names) * n_samples
            cnt = torch.zeros(max(n_seg, 1), device=seg.device).index_add_(0, seg, torch.ones_like(seg, dtype=torch.float32))
            self._graph = dict(sub=sub[:, None], sub_flat=sub.to(torch.int64).contiguous(), frag_index=frag_index, seg=seg,
                               n_seg=max(n_seg, 1), n_samples=n_samples, inv_cnt=(1.0 / cnt.clamp(min=1))[:, None])
            self._graph_key = key
            # the key identifies the inputs by address and version: keep them alive while cached, so that no OTHER tensor can
            # take their address (the caching allocators hand a freed block straight to the next request of that size)
            self._graph_refs = (edge_index, n_frag_switch, combined_mask)
        return self._graph

    # ---- device-resident path: prologue, LEFTNet and epilogue behind the C ABI (oard_dyn_forward / oard_reverse_step)
    def fused_ok(self, device) -> bool:
        if torch.is_grad_enabled() and getattr(self.model, "enable_training_path", False):
            return False  # training: the torch-composed wrapper below carries autograd through encoders / decoders
        return (isinstance(self.model, LEFTNetB200) and torch.device(device).type == "cuda" and self.pos_dim == 3
                and len(set(self.node_nfs)) == 1 and self.node_nfs[0] - 3 <= 16 and self.update_pocket_coords
                and len(self.fragment_names) <= 8 and getattr(self, "use_fused", True))

    def fused_engine(self, device, edge_index: Tensor, n_frag_switch: Tensor, combined_mask: Tensor):
        """The model's engine with weights, plan and dynamics plan current for this graph; returns (engine, graph cache)."""
        g = self._graph_cache(edge_index, n_frag_switch, combined_mask)
        eng = self.model.engine(device)
        if not (self.model.assume_static_weights and eng.weights_key is not None):
            eng.sync_weights(self.model)
        eng.plan(edge_index, combined_mask.numel())
        if g.get("sub_planned_key") != eng.plan_key:  # same-fragment mask in the engine's edge order (same tensor if grouped by source)
            g["sub_planned"], g["sub_planned_key"] = eng.edge_order(g["sub_flat"]), eng.plan_key
        if not (self.model.assume_static_weights and getattr(eng, "dyn_weights_key", None) is not None):
            eng.dyn_sync(self, len(self.fragment_names), self.node_nfs[0], self.condition_nf, self.condition_time)
        eng.dyn_plan(n_frag_switch, combined_mask, g["n_samples"])
        return eng, g

    @torch.no_grad()
    def _forward_fused(self, xh, edge_index, t, conditions, n_frag_switch, combined_mask):
        dev = xh[0].device
        eng, g = self.fused_engine(dev, edge_index, n_frag_switch, combined_mask)
        B = g["n_samples"]
        Z = torch.cat(xh).to(torch.float32).contiguous()
        tb = None
        if self.condition_time:
            tb = (t.reshape(1).expand(B) if t.numel() == 1 else t.reshape(-1)).to(torch.float32).contiguous()
        cond = conditions.to(torch.float32).reshape(B, -1).contiguous() if self.condition_nf > 0 else None
        sub = g["sub_planned"] if self.model.object_aware else None
        eps = eng.dyn_forward(Z, tb, cond, sub, torch.empty_like(Z))
        fi = g["frag_index"]
        return [eps[fi[ii]:fi[ii + 1]] for ii in range(len(self.fragment_names))], None

    def forward(self, xh: List[Tensor], edge_index: Tensor, t: Tensor, conditions: Tensor, n_frag_switch: Tensor,
                combined_mask: Tensor, edge_attr: Optional[Tensor] = None) -> Tuple[List[Tensor], Tensor]:
        """Predict eps for every fragment (egnn_dynamics.py:63-168).  Returns (list of [N_f, node_nf], None)."""
        if edge_attr is None and self.fused_ok(xh[0].device):
            return self._forward_fused(xh, edge_index, t, conditions, n_frag_switch, combined_mask)
        g = self._graph_cache(edge_index, n_frag_switch, combined_mask)
        p = self.pos_dim
        pos = torch.cat([_xh[:, :p] for _xh in xh], dim=0)
        h = torch.cat([self.encoders[ii](xh[ii][:, p:]) for ii in range(len(self.fragment_names))], dim=0)
        condition_dim = 0
        if self.condition_time:
            h_time = t.reshape(1, 1).expand(h.size(0), 1) if t.dim() == 1 else t[combined_mask]
            h = torch.cat([h, h_time.to(h.dtype)], dim=1)
            condition_dim += 1
        if self.condition_nf > 0:
            h = torch.cat([h, conditions[combined_mask].to(h.dtype)], dim=1)
            condition_dim += self.condition_nf
        if not self.update_pocket_coords:
            raise NotImplementedError  # egnn_dynamics.py:125
        h_final, pos_final, _ = self.model(h, pos, edge_index, None, node_mask=None, edge_mask=None,
                                           update_coords_mask=None, subgraph_mask=g["sub"])
        vel = pos_final - pos
        # NaN guard (egnn_dynamics.py:138-143) without a host sync: if anything is NaN the whole output is replaced by
        # noise drawn from a private generator, so the caller's RNG stream is untouched in the normal case.
        bad = torch.isnan(vel).any()
        if self._nan_gen is None or self._nan_gen.device != vel.device:
            self._nan_gen = torch.Generator(device=vel.device)
            self._nan_gen.manual_seed(0)
        vel = torch.where(bad, torch.randn(vel.shape, device=vel.device, generator=self._nan_gen), vel)
        h_final = h_final[:, :-condition_dim] if condition_dim else h_final
        # per-(fragment, sample) centre-of-mass removal in one segmented pass (egnn_dynamics.py:147-160, 267-271)
        mean = torch.zeros(g["n_seg"], p, device=vel.device, dtype=vel.dtype).index_add_(0, g["seg"], vel) * g["inv_cnt"]
        vel = vel - mean[g["seg"]]
        fi = g["frag_index"]
        xh_final = [torch.cat([vel[fi[ii]:fi[ii + 1]], self.decoders[ii](h_final[fi[ii]:fi[ii + 1]])], dim=-1)
                    for ii in range(len(self.fragment_names))]
        return xh_final, None

    @staticmethod
    def compute_frag_index(n_frag_switch: Tensor) -> np.ndarray:
        """egnn_dynamics.py:177-182 (host side; cached per graph by the caller)."""
        counts = torch.bincount(n_frag_switch).cpu().numpy()
        counts = counts[counts > 0] if len(counts) else counts
        return np.concatenate([np.array([0]), np.cumsum(counts)]).astype(np.int64)

    @staticmethod
    def remove_mean_batch(x, indices):
        n = int(indices.max().item()) + 1 if indices.numel() else 0
        tot = torch.zeros(n, x.size(1), device=x.device, dtype=x.dtype).index_add_(0, indices, x)
        cnt = torch.zeros(n, device=x.device, dtype=x.dtype).index_add_(0, indices, torch.ones_like(indices, dtype=x.dtype))
        return x - (tot / cnt.clamp(min=1)[:, None])[indices]
