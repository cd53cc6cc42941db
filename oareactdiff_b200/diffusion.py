"""`EnVariationalDiffusion` sampling API with the reference's names and signatures
(oa_reactdiff/diffusion/en_diffusion.py:26-52, 260-304, 459-718, 722-883, 1050-1074).

Host-side PyTorch, as the north star prescribes: tensor plumbing and the noise schedule live here, the denoiser
evaluated at every step is the CUDA LEFTNet behind `dynamics`.  The reverse loop issues no host<->device sync:
the reference's per-step `assert_mean_zero_with_mask` (4 `.item()` calls per step, _utils.py:15-19) runs only at the
start and end of a trajectory unless `debug_asserts=True`.  The RNG draw order of the reference is kept (per
fragment: positions, then features).  Training `forward()` (loss terms) is a later row of SURVEY §8f.
"""
import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .dynamics import EGNNDynamics
from .graph_tools import get_edges_index, get_mask_for_frag, get_n_frag_switch
from .normalizer import FEATURE_MAPPING, Normalizer
from .schedule import DiffSchedule, get_repaint_schedule


def _n_seg(indices: Tensor, n_seg: Optional[int]) -> int:
    # torch_scatter sizes its output from indices.max() (one host sync); the samplers pass the batch size instead
    return int(indices.max()) + 1 if n_seg is None and indices.numel() else (n_seg or 0)


def remove_mean_batch(x: Tensor, indices: Tensor, n_seg: Optional[int] = None) -> Tensor:
    """x - segment_mean(x)[indices]  (diffusion/_utils.py:9-12); pass the segment count to avoid the host sync."""
    n_seg = _n_seg(indices, n_seg)
    tot = torch.zeros(n_seg, x.size(1), device=x.device, dtype=x.dtype).index_add_(0, indices, x)
    cnt = torch.zeros(n_seg, device=x.device, dtype=x.dtype).index_add_(0, indices, torch.ones_like(indices, dtype=x.dtype))
    return x - (tot / cnt.clamp(min=1)[:, None])[indices]


def assert_mean_zero_with_mask(x: Tensor, node_mask: Tensor, n_seg: Optional[int] = None, eps: float = 1e-10):
    """diffusion/_utils.py:15-19."""
    n_seg = _n_seg(node_mask, n_seg)
    largest = x.abs().max().item()
    err = torch.zeros(n_seg, x.size(1), device=x.device, dtype=x.dtype).index_add_(0, node_mask, x).abs().max().item()
    rel = err / (largest + eps)
    assert rel < 1e-2, f"Mean is not zero, relative_error {rel}"


def sample_center_gravity_zero_gaussian_batch(size: List[int], indices: List[Tensor]) -> Tensor:
    """Standard normal projected on the zero-centre-of-mass subspace of every sample (diffusion/_utils.py:22-32)."""
    assert len(size) == 2
    return remove_mean_batch(torch.randn(size, device=indices[0].device), torch.cat(indices))


def sum_except_batch(x: Tensor, indices: Tensor, dim_size: int) -> Tensor:
    """Per-sample sum over nodes and channels (diffusion/_utils.py:35-36)."""
    v = x.sum(-1)
    return torch.zeros(dim_size, device=x.device, dtype=v.dtype).index_add_(0, indices, v)


def cdf_standard_gaussian(x: Tensor) -> Tensor:
    """diffusion/_utils.py:39-40."""
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2)))


def sample_gaussian(size, device) -> Tensor:
    """diffusion/_utils.py:43-45."""
    return torch.randn(size, device=device)


def num_nodes_to_batch_mask(n_samples: int, num_nodes, device) -> Tensor:
    """Sample id per node from per-sample node counts (diffusion/_utils.py:48-57)."""
    assert isinstance(num_nodes, int) or len(num_nodes) == n_samples
    if isinstance(num_nodes, Tensor):
        num_nodes = num_nodes.to(device)
    return torch.repeat_interleave(torch.arange(n_samples, device=device), num_nodes)


class EnVariationalDiffusion(nn.Module):
    def __init__(self, dynamics: EGNNDynamics, schdule: DiffSchedule, normalizer: Normalizer,
                 size_histogram: Optional[Dict] = None, loss_type: str = "l2", pos_only: bool = False,
                 fixed_idx: Optional[List] = None, debug_asserts: bool = False):
        super().__init__()
        assert loss_type in {"vlb", "l2"}
        self.dynamics, self.schedule, self.normalizer = dynamics, schdule, normalizer
        self.size_histogram, self.loss_type, self.pos_only = size_histogram, loss_type, pos_only
        self.fixed_idx = fixed_idx or []
        self.pos_dim, self.node_nfs, self.fragment_names = dynamics.pos_dim, dynamics.node_nfs, dynamics.fragment_names
        self.T = schdule.gamma_module.timesteps
        self.norm_values, self.norm_biases = normalizer.norm_values, normalizer.norm_biases
        self.debug_asserts = debug_asserts
        self.n_evals = 0  # denoiser evaluations of the last sample()/inpaint() call

    # ---------------------------------------------------------------- training / evaluation loss terms (forward only)
    def _draw_t_int(self, num_sample: int, device) -> Tensor:
        lowest_t = 0 if self.training else 1
        return torch.randint(lowest_t, self.T + 1, size=(num_sample, 1), device=device).float()

    def forward(self, representations: List[Dict], conditions: Tensor, return_pred: bool = False):
        """Loss and NLL terms of one noised batch (en_diffusion.py:56-248), same keys and arithmetic as the reference.
        Values only (no autograd graph) unless the model's `enable_training_path` is set and grad mode is on, in which case
        the terms carry gradients to every parameter of the dynamics (oard_forward_train / oard_backward)."""
        with_grad = torch.is_grad_enabled() and getattr(getattr(self.dynamics, "model", None), "enable_training_path", False)
        with torch.set_grad_enabled(with_grad):
            return self._forward_terms(representations, conditions, return_pred)

    def _forward_terms(self, representations: List[Dict], conditions: Tensor, return_pred: bool = False):
        num_sample = representations[0]["size"].size(0)
        n_nodes = torch.stack([rep["size"] for rep in representations], dim=0).sum(dim=0)
        device = representations[0]["pos"].device
        masks = [rep["mask"] for rep in representations]
        self._combined, self._B = torch.cat(masks), num_sample
        edge_index = get_edges_index(self._combined, remove_self_edge=True)
        nfs = get_n_frag_switch([rep["size"] for rep in representations])
        representations = self.normalizer.normalize(representations)
        delta_log_px = self.delta_log_px(n_nodes.sum())
        t_int = self._draw_t_int(num_sample, device)
        s_int = t_int - 1
        t_is_zero = (t_int == 0).float()
        t_is_not_zero = 1 - t_is_zero
        s, t = s_int / self.T, t_int / self.T
        ref = representations[0]["pos"]
        gamma_s = self.schedule.inflate_batch_array(self.schedule.gamma_module(s), ref)
        gamma_t = self.schedule.inflate_batch_array(self.schedule.gamma_module(t), ref)
        xh = [torch.cat([rep[k] for k in FEATURE_MAPPING], dim=1) for rep in representations]
        z_t, eps_xh = self.noised_representation(xh, masks, gamma_t)
        net_eps_xh = self._dyn(z_t, edge_index, t, conditions, nfs, masks)
        if return_pred:
            return eps_xh, net_eps_xh
        p = self.pos_dim
        if self.pos_only:
            for ii in range(len(masks)):
                net_eps_xh[ii][:, p:] = 0.0
        error_t = [self._sum_except_batch((eps_xh[ii] - net_eps_xh[ii]) ** 2, masks[ii]) for ii in range(len(masks))]
        SNR_weight = (1 - self.schedule.SNR(gamma_s - gamma_t)).squeeze(1)
        neg_log_constants = -self.log_constants_p_x_given_z0(n_nodes, device)
        kl_prior = torch.zeros_like(neg_log_constants)
        if self.training:
            log_p = self.log_pxh_given_z0_without_constants(representations, z_t, eps_xh, net_eps_xh, gamma_t, 1e-10)
            tz = t_is_zero.squeeze()
            loss_0_x, loss_0_cat, loss_0_charge = ([-lp * tz for lp in log_p[k]] for k in range(3))
            error_t = [e * t_is_not_zero.squeeze() for e in error_t]
        else:
            t_zeros = torch.zeros_like(s)
            gamma_0 = self.schedule.inflate_batch_array(self.schedule.gamma_module(t_zeros), ref)
            z_0, eps_0 = self.noised_representation(xh, masks, gamma_0)
            net_eps_0 = self._dyn(z_0, edge_index, t_zeros, conditions, nfs, masks)
            log_p = self.log_pxh_given_z0_without_constants(representations, z_0, eps_0, net_eps_0, gamma_0, 1e-10)
            loss_0_x, loss_0_cat, loss_0_charge = ([-lp for lp in log_p[k]] for k in range(3))
        return {"delta_log_px": delta_log_px, "error_t": error_t, "SNR_weight": SNR_weight, "loss_0_x": loss_0_x,
                "loss_0_cat": loss_0_cat, "loss_0_charge": loss_0_charge, "neg_log_constants": neg_log_constants,
                "kl_prior": kl_prior, "log_pN": torch.zeros_like(kl_prior), "t_int": t_int.squeeze(),
                "net_eps_xh": net_eps_xh, "eps_xh": eps_xh}

    def _sum_except_batch(self, x: Tensor, indices: Tensor) -> Tensor:
        """diffusion/_utils.py:34-35."""
        return torch.zeros(self._B, device=x.device, dtype=x.dtype).index_add_(0, indices, x.sum(-1))

    def delta_log_px(self, num_nodes):
        return -self.subspace_dimensionality(num_nodes) * math.log(self.norm_values[0])  # en_diffusion.py:250-251

    def subspace_dimensionality(self, input_size):
        return (input_size - 1) * self.pos_dim  # en_diffusion.py:253-258

    def log_constants_p_x_given_z0(self, n_nodes: Tensor, device) -> Tensor:
        """en_diffusion.py:306-318."""
        batch_size = len(n_nodes)
        dof = self.subspace_dimensionality(n_nodes).to(device)
        gamma_0 = self.schedule.gamma_module(torch.zeros((batch_size, 1), device=device))
        log_sigma_x = 0.5 * gamma_0.view(batch_size)
        return dof * (-log_sigma_x - 0.5 * math.log(2 * math.pi))

    def kl_prior(self):
        """Placeholder of the reference (en_diffusion.py:319-320): it RETURNS the exception class; forward() uses zeros."""
        return NotImplementedError

    @staticmethod
    def gaussian_KL(q_mu_minus_p_mu_squared, q_sigma, p_sigma, d):
        """KL(q || p) of two isotropic d-dimensional normals from ||mu_q - mu_p||^2 and the two scales
        (en_diffusion.py:322-338)."""
        return d * torch.log(p_sigma / q_sigma) + 0.5 * (d * q_sigma ** 2 + q_mu_minus_p_mu_squared) / (p_sigma ** 2) - 0.5 * d

    def log_pxh_given_z0_without_constants(self, representations, z_t, eps_xh, net_eps_xh, gamma_t, epsilon=1e-10):
        """L0 terms: Gaussian position term, discretised-Gaussian atom-type and charge terms (en_diffusion.py:340-454)."""
        p, n = self.pos_dim, len(representations)
        cdf = lambda x: 0.5 * (1.0 + torch.erf(x / math.sqrt(2)))
        masks = [rep["mask"] for rep in representations]
        log_p_x = [-0.5 * self._sum_except_batch((eps_xh[ii][:, :p] - net_eps_xh[ii][:, :p]) ** 2, masks[ii]) for ii in range(n)]
        z_t = [z[:, :3 + 5 + 1] for z in z_t]
        for rep in representations:
            rep["charge"] = rep["charge"][:, :1]
        sigma_0 = self.schedule.sigma(gamma_t, target_tensor=z_t[0])
        sigma_0_cat = sigma_0 * self.normalizer.norm_values[1]
        atoms = [self.normalizer.unnormalize(rep["one_hot"], 1) for rep in representations]
        centered_atoms = [self.normalizer.unnormalize(z[:, p:-1], 1) - 1 for z in z_t]
        log_ph_cat = [torch.log(cdf((centered_atoms[ii] + 0.5) / sigma_0_cat[masks[ii]])
                                - cdf((centered_atoms[ii] - 0.5) / sigma_0_cat[masks[ii]]) + epsilon) for ii in range(n)]
        log_prob = [lp - torch.logsumexp(lp, dim=1, keepdim=True) for lp in log_ph_cat]
        log_p_hcat = [self._sum_except_batch(log_prob[ii] * atoms[ii], masks[ii]) for ii in range(n)]
        sigma_0_charge = sigma_0 * self.normalizer.norm_values[2]
        charges = [self.normalizer.unnormalize(rep["charge"], 2) for rep in representations]
        est_charges = [self.normalizer.unnormalize(z[:, -1:], 2).long() for z in z_t]
        centered_charges = [charges[ii] - est_charges[ii] for ii in range(n)]
        log_ph_charge = [torch.log(cdf((centered_charges[ii] + 0.5) / sigma_0_charge[masks[ii]])
                                   - cdf((centered_charges[ii] - 0.5) / sigma_0_charge[masks[ii]]) + epsilon) for ii in range(n)]
        log_p_hcharge = [self._sum_except_batch(log_ph_charge[ii], masks[ii]) for ii in range(n)]
        return [log_p_x, log_p_hcat, log_p_hcharge]

    def compute_loss(self, batch, scales=(1.0, 1.0, 1.0), training: Optional[bool] = None):
        """The loss composition of `DDPMModule.compute_loss` (trainer/pl_trainer.py:208-282) on this module's terms:
        returns (nll[B], info).  `scales` are the per-fragment weights (train_ts1x.py:111 uses [1, 2, 1])."""
        representations, conditions = batch
        training = self.training if training is None else training
        sizes = [rep["size"] for rep in representations]
        # forward() keys its t range, t == 0 masking and second (t = 0) denoiser call on self.training: one flag, like the
        # reference (DDPMModule.training is ddpm.training) — an explicit `training` switches the module mode for the call
        was_training = self.training
        if training != was_training:
            self.train(training)
        try:
            lt = self.forward(representations, conditions)
        finally:
            if training != was_training:
                self.train(was_training)
        nfr = len(sizes)
        denoms = [(self.pos_dim if self.pos_only else self.pos_dim + self.node_nfs[ii]) * sizes[ii] for ii in range(nfr)]
        err_norm = [lt["error_t"][ii] / denoms[ii] * scales[ii] for ii in range(nfr)]
        if self.loss_type == "l2" and training:
            loss_t = torch.stack(err_norm, dim=0).sum(dim=0)
            l0x = torch.stack([lt["loss_0_x"][ii] * scales[ii] / (self.pos_dim * sizes[ii]) for ii in range(nfr)], dim=0).sum(dim=0)
            loss_0 = l0x + torch.stack(lt["loss_0_cat"], dim=0).sum(dim=0) + torch.stack(lt["loss_0_charge"], dim=0).sum(dim=0)
        else:
            loss_t = torch.stack([-self.T * 0.5 * lt["SNR_weight"] * e for e in lt["error_t"]], dim=0).sum(dim=0)
            loss_0 = (torch.stack(lt["loss_0_x"], dim=0).sum(dim=0) + torch.stack(lt["loss_0_cat"], dim=0).sum(dim=0)
                      + torch.stack(lt["loss_0_charge"], dim=0).sum(dim=0) + lt["neg_log_constants"])
        nll = loss_t + loss_0 + lt["kl_prior"]
        info = {}
        for ii in range(nfr):
            info[f"error_t_{ii}"] = err_norm[ii].mean() / (scales[ii] + 1e-4)
            info[f"unorm_error_t_{ii}"] = lt["error_t"][ii].mean()
        if not (self.loss_type == "l2" and training):
            nll = nll - lt["delta_log_px"] - lt["log_pN"]
        return nll, info

    # ---------------------------------------------------------------- noise
    def sample_combined_position_feature_noise(self, masks: List[Tensor]) -> List[Tensor]:
        """CoM-free Gaussian for positions, standard normal (zero if pos_only) for features (en_diffusion.py:281-304)."""
        out = []
        for ii, mask in enumerate(masks):
            x = remove_mean_batch(torch.randn((len(mask), self.pos_dim), device=mask.device), mask, self._B)
            hh = torch.randn((len(mask), self.node_nfs[ii] - self.pos_dim), device=mask.device)
            if self.pos_only:
                hh = torch.zeros_like(hh)
            out.append(torch.cat([x, hh], dim=1))
        for idx in self.fixed_idx:
            out[idx] = torch.zeros_like(out[idx])
        return out

    def noised_representation(self, xh: List[Tensor], masks: List[Tensor], gamma_t: Tensor):
        """z_t = alpha_t x + sigma_t eps (en_diffusion.py:260-279)."""
        alpha_t, sigma_t = self.schedule.alpha(gamma_t, xh[0]), self.schedule.sigma(gamma_t, xh[0])
        eps = self.sample_combined_position_feature_noise(masks)
        return [alpha_t[masks[ii]] * xh[ii] + sigma_t[masks[ii]] * eps[ii] for ii in range(len(masks))], eps

    def sample_normal(self, mu: List[Tensor], sigma: Tensor, masks: List[Tensor], fix_noise: bool = False):
        if fix_noise:
            raise NotImplementedError("fix_noise option isn't implemented yet")  # en_diffusion.py:642-644
        eps = self.sample_combined_position_feature_noise(masks)
        return [mu[ii] + sigma[masks[ii]] * eps[ii] for ii in range(len(masks))]

    # ---------------------------------------------------------------- reverse kernels
    def _dyn(self, z, edge_index, t, conditions, n_frag_switch, masks):
        self.n_evals += 1
        eps, _ = self.dynamics(xh=z, edge_index=edge_index, t=t, conditions=conditions, n_frag_switch=n_frag_switch,
                               combined_mask=self._combined, edge_attr=None)
        return eps

    def sample_p_zs_given_zt(self, s, t, zt_xh, edge_index, n_frag_switch, masks, conditions=None, fix_noise=False):
        """One reverse step z_t -> z_s (en_diffusion.py:562-632)."""
        gamma_s, gamma_t = self.schedule.gamma_module(s), self.schedule.gamma_module(t)
        sigma2_ts, sigma_ts, alpha_ts = self.schedule.sigma_and_alpha_t_given_s(gamma_t, gamma_s, zt_xh[0])
        sigma_s = self.schedule.sigma(gamma_s, target_tensor=zt_xh[0])
        sigma_t = self.schedule.sigma(gamma_t, target_tensor=zt_xh[0])
        eps = self._dyn(zt_xh, edge_index, t, conditions, n_frag_switch, masks)
        if self.debug_asserts:
            for zz in (zt_xh, eps):
                assert_mean_zero_with_mask(torch.cat([z[:, :self.pos_dim] for z in zz]), self._combined, self._B)
        coef = sigma2_ts / alpha_ts / sigma_t
        mu = [zt_xh[ii] / alpha_ts[masks[ii]] - eps[ii] * coef[masks[ii]] for ii in range(len(zt_xh))]
        zs = self.sample_normal(mu=mu, sigma=sigma_ts * sigma_s / sigma_t, masks=masks, fix_noise=fix_noise)
        for ii in range(len(masks)):  # project the CoM out again
            zs[ii][:, :self.pos_dim] = remove_mean_batch(zs[ii][:, :self.pos_dim], masks[ii], self._B)
        return zs

    def compute_x_pred(self, net_eps_xh, zt_xh, gamma_t, masks):
        """en_diffusion.py:704-718."""
        sigma_t = self.schedule.sigma(gamma_t, target_tensor=net_eps_xh[0])
        alpha_t = self.schedule.alpha(gamma_t, target_tensor=net_eps_xh[0])
        return [1.0 / alpha_t[masks[ii]] * (zt_xh[ii] - sigma_t[masks[ii]] * net_eps_xh[ii]) for ii in range(len(masks))]

    def sample_p_xh_given_z0(self, z0_xh, edge_index, n_frag_switch, masks, batch_size, conditions=None,
                             fix_noise=False):
        """Final decode x ~ p(x | z0) (en_diffusion.py:649-702)."""
        t0 = torch.zeros(size=(batch_size, 1), device=z0_xh[0].device)
        gamma_0 = self.schedule.gamma_module(t0)
        sigma_x = self.schedule.SNR(-0.5 * gamma_0)
        eps = self._dyn(z0_xh, edge_index, t0, conditions, n_frag_switch, masks)
        mu_x = self.compute_x_pred(eps, z0_xh, gamma_0, masks)
        x0 = self.sample_normal(mu=mu_x, sigma=sigma_x, masks=masks, fix_noise=fix_noise)
        p = self.pos_dim
        # (the reference indexes the normalizer with the FRAGMENT index here, en_diffusion.py:686-697 — identical for the identity
        # normalizer of the shipped configurations; mirrored so that a non-default normalizer gives the reference's numbers)
        pos = [self.normalizer.unnormalize(x0[ii][:, :p], ii) for ii in range(len(masks))]
        cat = [self.normalizer.unnormalize(x0[ii][:, p:-1], ii) for ii in range(len(masks))]
        charge = [torch.round(self.normalizer.unnormalize(x0[ii][:, -1:], ii)).long() for ii in range(len(masks))]
        cat = [F.one_hot(torch.argmax(cat[ii], dim=1), self.node_nfs[ii] - 4).long() for ii in range(len(masks))]
        return pos, cat, charge

    def sample_p_zt_given_zs(self, zs, masks, gamma_t, gamma_s, fix_noise=False):
        """Forward jump z_s -> z_t used by RePaint (en_diffusion.py:1050-1074)."""
        _, sigma_ts, alpha_ts = self.schedule.sigma_and_alpha_t_given_s(gamma_t, gamma_s, zs[0])
        zt = self.sample_normal(mu=[alpha_ts[masks[ii]] * zs[ii] for ii in range(len(masks))], sigma=sigma_ts,
                                masks=masks, fix_noise=fix_noise)
        for ii in range(len(masks)):
            zt[ii][:, :self.pos_dim] = remove_mean_batch(zt[ii][:, :self.pos_dim], masks[ii], self._B)
        return zt

    # ---------------------------------------------------------------- fast reverse step (same arithmetic, fewer launches)
    # sample()/inpaint() always call the reverse kernel with ONE (s, t) pair for the whole batch, so the schedule scalars
    # of every step can be tabulated once (with the same float32 torch formulas as above, evaluated for all steps at
    # once) and the per-fragment lists can live in one concatenated [N, nf] tensor.  ~20 launches per step instead of ~95.
    def _fast_ok(self) -> bool:
        own_noise = type(self).sample_combined_position_feature_noise is EnVariationalDiffusion.sample_combined_position_feature_noise
        return len(set(self.node_nfs)) == 1 and not self.fixed_idx and not self.debug_asserts and own_noise

    def _tables(self, timesteps: int, device):
        gam = self.schedule.gamma_module.gamma  # callers swap `ddpm.schedule` / `ddpm.T` after construction (evaluate/utils.py:28-31)
        key = (timesteps, str(device), id(self.schedule), gam.data_ptr(), gam._version, tuple(self.schedule.norm_values))
        if getattr(self, "_tab_key", None) == key:
            return self._tab
        steps = torch.arange(timesteps + 1, device=device)
        tt = (steps / timesteps).view(-1, 1)                       # t = k / T, float32 like (s_array + 1) / timesteps
        gamma = self.schedule.gamma_module(tt)                     # [T+1, 1]
        g_s, g_t = gamma[:-1], gamma[1:]
        ref = torch.zeros(1, 1, device=device)
        sigma2_ts, sigma_ts, alpha_ts = self.schedule.sigma_and_alpha_t_given_s(g_t, g_s, ref)
        sigma_s, sigma_t = self.schedule.sigma(g_s, ref), self.schedule.sigma(g_t, ref)
        tab = {
            "tt": tt,                                               # device tensor, row k = t value of step k
            "t": tt.flatten().tolist(),
            "inv_alpha_ts": (1.0 / alpha_ts).flatten().tolist(),    # mu = z / alpha_ts - eps * coef  (z / a == z * (1/a) is NOT
            "alpha_ts": alpha_ts.flatten().tolist(),                #   bit-identical, so the division is kept: see _fast_step)
            "coef": (sigma2_ts / alpha_ts / sigma_t).flatten().tolist(),
            "sigma": (sigma_ts * sigma_s / sigma_t).flatten().tolist(),
            "alpha": self.schedule.alpha(gamma, ref).flatten().tolist(),
            "sigma_abs": self.schedule.sigma(gamma, ref).flatten().tolist(),
            "sigma_ts": sigma_ts.flatten().tolist(),
            "gamma": gamma.detach().float().cpu().flatten(),        # gamma(k / timesteps): the jump-back of inpaint() pairs any s < t
        }
        self._tab_key, self._tab, self._tab_refs = key, tab, (self.schedule, gam)  # (identity keys: keep the keyed objects alive)
        return tab

    def _seg_setup(self, masks):
        """Per-(fragment, sample) segment ids of the concatenated node order and their inverse sizes."""
        B = self._B
        seg = torch.cat([m + f * B for f, m in enumerate(masks)])
        cnt = torch.zeros(len(masks) * B, device=seg.device).index_add_(0, seg, torch.ones(seg.numel(), device=seg.device))
        self._seg, self._seg_inv = seg, (1.0 / cnt.clamp(min=1))[:, None]
        self._frag_off = [0]
        for m in masks:
            self._frag_off.append(self._frag_off[-1] + len(m))

    def _remove_mean_cat(self, x: Tensor) -> Tensor:
        mean = torch.zeros(self._seg_inv.size(0), x.size(1), device=x.device, dtype=x.dtype).index_add_(0, self._seg, x)
        return x - (mean * self._seg_inv)[self._seg]

    def _noise_cat(self, masks) -> Tensor:
        """Concatenated CoM-free noise with the reference's draw order (per fragment: positions, then features)."""
        xs, hs = [], []
        for ii, mask in enumerate(masks):
            xs.append(torch.randn((len(mask), self.pos_dim), device=mask.device))
            hs.append(torch.randn((len(mask), self.node_nfs[ii] - self.pos_dim), device=mask.device))
        x = self._remove_mean_cat(torch.cat(xs))
        h = torch.cat(hs)
        if self.pos_only:
            h = torch.zeros_like(h)
        return torch.cat([x, h], dim=1)

    def _views(self, Z: Tensor):
        o = self._frag_off
        return [Z[o[f]:o[f + 1]] for f in range(len(o) - 1)]

    def _fast_step(self, s_int: int, Z: Tensor, tab, edge_index, nfs, masks, conditions) -> Tensor:
        """z_s ~ p(z_s | z_t) for t = s + 1 on the concatenated state (same formulas as sample_p_zs_given_zt)."""
        eps = torch.cat(self._dyn(self._views(Z), edge_index, tab["tt"][s_int + 1].expand(self._B, 1), conditions, nfs, masks))
        mu = Z / tab["alpha_ts"][s_int] - eps * tab["coef"][s_int]
        Zs = mu + tab["sigma"][s_int] * self._noise_cat(masks)
        Zs[:, :self.pos_dim] = self._remove_mean_cat(Zs[:, :self.pos_dim])
        return Zs

    # ---------------------------------------------------------------- device-resident reverse step
    # One reverse step = the caller's six standard-normal draws (reference order, torch generator) + ONE C call that
    # replays the whole step (dynamics prologue, LEFTNet, epilogue, posterior update, CoM projections, h0 overwrite) as a
    # CUDA graph on persistent buffers: oard_reverse_step (include/oard.h).
    def _device_ok(self, device) -> bool:
        return self._fast_ok() and self.pos_dim == 3 and getattr(self.dynamics, "fused_ok", lambda d: False)(device)

    def _device_setup(self, Z: Tensor, masks, edge_index, nfs, conditions, H0: Optional[Tensor]):
        dyn = self.dynamics
        eng, g = dyn.fused_engine(Z.device, edge_index, nfs, self._combined)
        N, p, d = Z.size(0), self.pos_dim, self.node_nfs[0] - self.pos_dim
        nx, nh = torch.empty(N, p, device=Z.device), torch.empty(N, d, device=Z.device)
        o = self._frag_off
        views = []
        for f in range(len(masks)):  # draw order of sample_combined_position_feature_noise: positions, then features
            views += [nx[o[f]:o[f + 1]], nh[o[f]:o[f + 1]]]
        cond = None
        if dyn.condition_nf > 0:
            cond = conditions.to(torch.float32).reshape(self._B, -1).contiguous()
        self._dev = dict(eng=eng, nx=nx, nh=nh, views=views, cond=cond, sub=g["sub_planned"] if dyn.model.object_aware else None,
                         H0=None if H0 is None else H0.to(torch.float32).contiguous())

    def _device_step(self, s_int: int, Z: Tensor, tab) -> Tensor:
        """In place on Z (fp32, contiguous, persistent across the trajectory)."""
        dv = self._dev
        for v in dv["views"]:
            v.normal_()
        self.n_evals += 1
        dv["eng"].reverse_step(Z, dv["nx"], None if self.pos_only else dv["nh"], dv["H0"] if self.pos_only else None,
                               dv["cond"], dv["sub"], tab["t"][s_int + 1], tab["alpha_ts"][s_int], tab["coef"][s_int],
                               tab["sigma"][s_int])
        return Z

    # RePaint on the device: the draw for the clamped fragments, the reverse step, the blend and the jump-back
    def _device_setup_inpaint(self, Xf: Tensor, frag_fixed):
        dv = self._dev
        N, p, d = Xf.size(0), self.pos_dim, self.node_nfs[0] - self.pos_dim
        kx, kh = torch.empty(N, p, device=Xf.device), torch.empty(N, d, device=Xf.device)
        o = self._frag_off
        views = []
        for f in range(len(o) - 1):  # noised_representation draws first (en_diffusion.py:806), same per-fragment order
            views += [kx[o[f]:o[f + 1]], kh[o[f]:o[f + 1]]]
        bits = 0
        for f in frag_fixed:
            bits |= 1 << int(f)
        dv.update(kx=kx, kh=kh, kviews=views, Xf=Xf.to(torch.float32).contiguous(), known_bits=bits)

    def _device_inpaint_step(self, s_int: int, Z: Tensor, tab) -> Tensor:
        dv = self._dev
        for v in dv["kviews"]:  # z_known's noise, then the reverse step's: the reference's draw order
            v.normal_()
        for v in dv["views"]:
            v.normal_()
        self.n_evals += 1
        dv["eng"].inpaint_step(Z, dv["nx"], None if self.pos_only else dv["nh"], dv["H0"] if self.pos_only else None,
                               dv["cond"], dv["sub"], tab["t"][s_int + 1], tab["alpha_ts"][s_int], tab["coef"][s_int],
                               tab["sigma"][s_int], dv["Xf"], dv["known_bits"], dv["kx"], None if self.pos_only else dv["kh"],
                               tab["alpha"][s_int], tab["sigma_abs"][s_int])
        return Z

    def _device_jump_back(self, Z: Tensor, alpha_ts: float, sigma_ts: float) -> Tensor:
        dv = self._dev
        for v in dv["views"]:
            v.normal_()
        dv["eng"].jump_back(Z, dv["nx"], None if self.pos_only else dv["nh"], alpha_ts, sigma_ts)
        return Z

    # ---------------------------------------------------------------- drivers
    def _setup(self, n_samples, fragments_nodes):
        masks = [get_mask_for_frag(n) for n in fragments_nodes]
        self._combined = torch.cat(masks)
        self._B = n_samples
        self.n_evals = 0
        return masks, get_edges_index(self._combined, remove_self_edge=True), get_n_frag_switch(fragments_nodes)

    def _with_h0(self, z, h0):
        return [torch.cat([z[ii][:, :self.pos_dim], h0[ii]], dim=1) for ii in range(len(h0))]

    def _check_com(self, zz):
        assert_mean_zero_with_mask(torch.cat([z[:, :self.pos_dim] for z in zz]), self._combined, self._B)

    @torch.no_grad()
    def sample(self, n_samples: int, fragments_nodes: List[Tensor], conditions: Optional[Tensor] = None,
               return_frames: int = 1, timesteps: Optional[int] = None, h0: Optional[List[Tensor]] = None):
        """Unconditional reverse diffusion (en_diffusion.py:459-560).  Returns (out_samples, fragments_masks) with
        out_samples[0] = list of [N_f, pos | one-hot | charge]."""
        timesteps = self.T if timesteps is None else timesteps
        assert 0 < return_frames <= timesteps and timesteps % return_frames == 0
        assert h0 is not None if self.pos_only else True
        masks, edge_index, nfs = self._setup(n_samples, fragments_nodes)
        z = self.sample_combined_position_feature_noise(masks)
        if self.pos_only:
            z = self._with_h0(z, h0)
        self._check_com(z)
        out_samples = [[torch.zeros((return_frames,) + zz.size(), device=zz.device) for zz in z]
                       for _ in range(return_frames)]
        dev = z[0].device
        if self._fast_ok():
            tab = self._tables(timesteps, dev)
            self._seg_setup(masks)
            Z = torch.cat(z).to(torch.float32).contiguous()
            H0 = torch.cat(h0).to(Z.dtype) if self.pos_only else None
            on_device = self._device_ok(dev)
            if on_device:
                self._device_setup(Z, masks, edge_index, nfs, conditions, H0)
            for s in reversed(range(0, timesteps)):
                if on_device:
                    self._device_step(s, Z, tab)
                else:
                    Z = self._fast_step(s, Z, tab, edge_index, nfs, masks, conditions)
                    if self.pos_only:
                        Z[:, self.pos_dim:] = H0
                if (s * return_frames) % timesteps == 0:
                    out_samples[(s * return_frames) // timesteps] = self.normalizer.unnormalize_z([v.clone() for v in self._views(Z)])
            z = self._views(Z)
        else:
            for s in reversed(range(0, timesteps)):
                s_arr = torch.full((n_samples, 1), fill_value=s, device=dev)
                t_arr = (s_arr + 1) / timesteps
                s_arr = s_arr / timesteps
                z = self.sample_p_zs_given_zt(s=s_arr, t=t_arr, zt_xh=z, edge_index=edge_index, n_frag_switch=nfs,
                                              masks=masks, conditions=conditions, fix_noise=False)
                if self.pos_only:
                    z = self._with_h0(z, h0)
                if (s * return_frames) % timesteps == 0:
                    out_samples[(s * return_frames) // timesteps] = self.normalizer.unnormalize_z(z)
        pos, cat, charge = self.sample_p_xh_given_z0(z, edge_index, nfs, masks, n_samples, conditions)
        if self.pos_only:
            cat = [_h0[:, :-1] for _h0 in h0]
            charge = [_h0[:, -1:] for _h0 in h0]
        self._check_com(pos)
        out_samples[0] = [torch.cat([pos[ii], cat[ii], charge[ii]], dim=1) for ii in range(len(pos))]
        return out_samples, masks

    @torch.no_grad()
    def inpaint(self, n_samples: int, fragments_nodes: List[Tensor], conditions: Optional[Tensor] = None,
                return_frames: int = 1, resamplings: int = 1, jump_length: int = 1, timesteps: Optional[int] = None,
                xh_fixed: Optional[List[Tensor]] = None, frag_fixed: Optional[List] = None):
        """RePaint-style conditional generation with the fragments in `frag_fixed` clamped to `xh_fixed`
        (en_diffusion.py:722-883)."""
        timesteps = self.T if timesteps is None else timesteps
        assert 0 < return_frames <= timesteps and timesteps % return_frames == 0
        assert len(xh_fixed)
        masks, edge_index, nfs = self._setup(n_samples, fragments_nodes)
        p = self.pos_dim
        h0 = [x[:, p:].long() for x in xh_fixed]
        for ii in range(len(xh_fixed)):
            xh_fixed[ii][:, :p] = remove_mean_batch(xh_fixed[ii][:, :p], masks[ii], n_samples)
        self._check_com(xh_fixed)
        z = self.sample_combined_position_feature_noise(masks)
        if self.pos_only:
            z = self._with_h0(z, h0)
        out_samples = [[torch.zeros((return_frames,) + zz.size(), device=zz.device) for zz in z]
                       for _ in range(return_frames)]
        dev = z[0].device
        schedule = get_repaint_schedule(resamplings, jump_length, timesteps)
        s = timesteps - 1
        if self._fast_ok():
            tab = self._tables(timesteps, dev)
            self._seg_setup(masks)
            gamma_cpu = tab["gamma"]  # indexed by step of THIS trajectory (timesteps may differ from the schedule's T)
            Z = torch.cat(z).to(torch.float32).contiguous()
            Xf = torch.cat(xh_fixed).to(torch.float32)
            H0 = torch.cat(h0).to(Z.dtype)
            known = torch.cat([torch.full((len(m), 1), ii in frag_fixed, dtype=torch.bool, device=dev)
                               for ii, m in enumerate(masks)])
            on_device = self._device_ok(dev) and len(masks) <= 8
            if on_device:
                self._device_setup(Z, masks, edge_index, nfs, conditions, H0 if self.pos_only else None)
                self._device_setup_inpaint(Xf, frag_fixed)
            for i, n_denoise_steps in enumerate(schedule):
                for j in range(n_denoise_steps):
                    # known fragments: q(z_s | x) (noised_representation); unknown: reverse step from z_t
                    if on_device:  # both draws, the step, the blend and the h0 overwrite: one CUDA-graph launch
                        self._device_inpaint_step(s, Z, tab)
                    else:
                        Z_known = tab["alpha"][s] * Xf + tab["sigma_abs"][s] * self._noise_cat(masks)
                        Z_unknown = self._fast_step(s, Z, tab, edge_index, nfs, masks, conditions)
                        Z = torch.where(known, Z_known, Z_unknown)
                        if self.pos_only:
                            Z[:, p:] = H0
                    if j == n_denoise_steps - 1 and i < len(schedule) - 1:  # jump back `jump_length` steps
                        t = s + jump_length
                        g_s, g_t = gamma_cpu[s].view(1, 1), gamma_cpu[t].view(1, 1)
                        _, sigma_ts, alpha_ts = self.schedule.sigma_and_alpha_t_given_s(g_t, g_s, g_s)
                        if on_device:
                            self._device_jump_back(Z, float(alpha_ts), float(sigma_ts))
                        else:
                            Zj = float(alpha_ts) * Z + float(sigma_ts) * self._noise_cat(masks)
                            Zj[:, :p] = self._remove_mean_cat(Zj[:, :p])
                            Z = Zj
                        s = t
                    s = s - 1
            z = self._views(Z)
            schedule = []
        for i, n_denoise_steps in enumerate(schedule):
            for j in range(n_denoise_steps):
                s_arr = torch.full((n_samples, 1), fill_value=s, device=dev)
                t_arr = (s_arr + 1) / timesteps
                s_arr = s_arr / timesteps
                gamma_s = self.schedule.inflate_batch_array(self.schedule.gamma_module(s_arr), xh_fixed[0])
                z_known, _ = self.noised_representation(xh_fixed, masks, gamma_s)
                z_unknown = self.sample_p_zs_given_zt(s=s_arr, t=t_arr, zt_xh=z, edge_index=edge_index,
                                                      n_frag_switch=nfs, masks=masks, conditions=conditions)
                if self.pos_only:
                    z_known, z_unknown = self._with_h0(z_known, h0), self._with_h0(z_unknown, h0)
                z = [z_known[ii] if ii in frag_fixed else z_unknown[ii] for ii in range(len(h0))]
                if j == n_denoise_steps - 1 and i < len(schedule) - 1:  # jump back `jump_length` steps
                    t = s + jump_length
                    t_arr = torch.full((n_samples, 1), fill_value=t, device=dev) / timesteps
                    gamma_t = self.schedule.inflate_batch_array(self.schedule.gamma_module(t_arr), xh_fixed[0])
                    z = self.sample_p_zt_given_zs(z, masks, gamma_t, gamma_s)
                    s = t
                s = s - 1
        pos, cat, charge = self.sample_p_xh_given_z0(z, edge_index, nfs, masks, n_samples, conditions)
        if self.pos_only:
            cat = [_h0[:, :-1] for _h0 in h0]
            charge = [_h0[:, -1:] for _h0 in h0]
        self._check_com(pos)
        out_samples[0] = [torch.cat([pos[ii], cat[ii], charge[ii]], dim=1) for ii in range(len(pos))]
        return out_samples, masks

    @torch.no_grad()
    def inpaint_fixed(self, n_samples: int, fragments_nodes: List[Tensor], conditions: Optional[Tensor] = None,
                      return_frames: int = 1, resamplings: int = 1, jump_length: int = 1, timesteps: Optional[int] = None,
                      xh_fixed: Optional[List[Tensor]] = None, frag_fixed: Optional[List] = None):
        """The reference keeps a second entry point whose body is statement for statement that of `inpaint`
        (en_diffusion.py:887-1048 vs :722-883); same here."""
        return self.inpaint(n_samples, fragments_nodes, conditions=conditions, return_frames=return_frames,
                            resamplings=resamplings, jump_length=jump_length, timesteps=timesteps, xh_fixed=xh_fixed,
                            frag_fixed=frag_fixed)
