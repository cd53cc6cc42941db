"""Noise schedules with the reference's names (oa_reactdiff/diffusion/_schedule.py).  gamma(t) = -log(alpha^2/sigma^2)
is a float32 lookup table with T+1 entries; sigma/alpha helpers follow from it."""
from typing import List, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor, nn


def clip_noise_schedule(alphas2: np.ndarray, clip_value: float = 0.001) -> np.ndarray:
    """Clip the per-step ratio alpha_t^2 / alpha_{t-1}^2 into [clip_value, 1] and re-accumulate (_schedule.py:43-57)."""
    ratio = np.concatenate([alphas2[:1], alphas2[1:] / alphas2[:-1]])  # alpha_{-1}^2 = 1
    return np.cumprod(np.clip(ratio, clip_value, 1.0))


def polynomial_schedule(timesteps: int, s: float = 1e-4, power: float = 3.0) -> np.ndarray:
    """alpha_t^2 = (1 - (t/T')^power)^2, ratio-clipped, squeezed into [s, 1 - s]; T+1 entries (_schedule.py:60-74)."""
    t = np.linspace(0, timesteps + 1, timesteps + 1)
    a2 = np.square(1.0 - (t / (timesteps + 1)) ** power)
    return (1 - 2 * s) * clip_noise_schedule(a2, clip_value=0.001) + s


def cosine_beta_schedule(timesteps: int, s: float = 0.008, raise_to_power: float = 1) -> np.ndarray:
    """Cumulative alpha^2 of the cosine schedule of Nichol & Dhariwal; T+1 entries (_schedule.py:9-26)."""
    x = np.linspace(0, timesteps + 2, timesteps + 2)
    f = np.cos((x / (timesteps + 2) + s) / (1 + s) * np.pi / 2) ** 2
    f = f / f[0]
    a2 = np.cumprod(1.0 - np.clip(1.0 - f[1:] / f[:-1], 0.0, 0.999))
    return a2 if raise_to_power == 1 else a2 ** raise_to_power


def ccosine_schedule(timesteps: int, start: float = 0, end: float = 1, tau: float = 1, clip_min: float = 1e-9) -> np.ndarray:
    """Continuous-time cosine schedule on [start, end] with exponent 2 tau; T+1 entries (_schedule.py:29-35)."""
    t = np.linspace(0, 1, timesteps + 1)
    lo, hi = np.cos(start * np.pi / 2) ** (2 * tau), np.cos(end * np.pi / 2) ** (2 * tau)
    out = (hi - np.cos((t * (end - start) + start) * np.pi / 2) ** (2 * tau)) / (hi - lo)
    return np.clip(out, clip_min, 1 - clip_min)


def linear_schedule(timesteps: int, clip_min: float = 1e-9) -> np.ndarray:
    """alpha_t^2 = 1 - t/T, clipped away from {0, 1}; T+1 entries (_schedule.py:38-41)."""
    return np.clip(1 - np.linspace(0, 1, timesteps + 1), clip_min, 1 - clip_min)


class PredefinedNoiseSchedule(nn.Module):
    """_schedule.py:77-129.  `noise_schedule` in {"cosine[_p]", "polynomial_p", "csin_a_b_tau", "linear"}."""

    def __init__(self, noise_schedule: str, timesteps: int, precision: float):
        super().__init__()
        self.timesteps = timesteps
        parts = noise_schedule.split("_")
        if "cosine" in noise_schedule:
            assert len(parts) <= 2
            a2 = cosine_beta_schedule(timesteps, raise_to_power=1.0 if len(parts) == 1 else float(parts[1]))
        elif "polynomial" in noise_schedule:
            assert len(parts) == 2
            a2 = polynomial_schedule(timesteps, s=precision, power=float(parts[1]))
        elif "csin" in noise_schedule:
            assert len(parts) == 4
            a2 = ccosine_schedule(timesteps, start=float(parts[1]), end=float(parts[2]), tau=float(parts[3]))
        elif "linear" in noise_schedule:
            a2 = linear_schedule(timesteps)
        else:
            raise ValueError(noise_schedule)
        gamma = -(np.log(a2) - np.log(1 - a2))
        self.gamma = nn.Parameter(torch.from_numpy(gamma).float(), requires_grad=False)

    def forward(self, t: Tensor) -> Tensor:
        return self.gamma[torch.round(t * self.timesteps).long()]


class DiffSchedule(nn.Module):
    """_schedule.py:132-203."""

    def __init__(self, gamma_module: nn.Module, norm_values: Tuple[float]):
        super().__init__()
        self.gamma_module = gamma_module
        self.norm_values = norm_values
        self.check_issues_norm_values()

    @staticmethod
    def inflate_batch_array(array: Tensor, target: Tensor) -> Tensor:
        return array.view((array.size(0),) + (1,) * (target.dim() - 1))

    def sigma(self, gamma, target_tensor):
        return self.inflate_batch_array(torch.sqrt(torch.sigmoid(gamma)), target_tensor)

    def alpha(self, gamma, target_tensor):
        return self.inflate_batch_array(torch.sqrt(torch.sigmoid(-gamma)), target_tensor)

    @staticmethod
    def SNR(gamma):
        return torch.exp(-gamma)

    def sigma_and_alpha_t_given_s(self, gamma_t: Tensor, gamma_s: Tensor, target_tensor: Tensor):
        sigma2 = self.inflate_batch_array(-torch.expm1(F.softplus(gamma_s) - F.softplus(gamma_t)), target_tensor)
        alpha = torch.exp(0.5 * (F.logsigmoid(-gamma_t) - F.logsigmoid(-gamma_s)))
        return sigma2, torch.sqrt(sigma2), self.inflate_batch_array(alpha, target_tensor)

    def check_issues_norm_values(self, num_stdevs=8):
        zeros = torch.zeros((1, 1))
        sigma_0 = self.sigma(self.gamma_module(zeros), target_tensor=zeros).item()
        if sigma_0 * num_stdevs > 1.0 / self.norm_values[1]:
            raise ValueError(f"Value for normalization value {self.norm_values[1]} probably too large with "
                             f"sigma_0 {sigma_0:.5f} and 1 / norm_value = {1. / self.norm_values[1]}")


def get_repaint_schedule(resamplings: int, jump_length: int, timesteps: int) -> List[int]:
    """RePaint segment lengths, last segment first (_schedule.py:206-232):
    sum(out) - (len(out) - 1) * jump_length == timesteps."""
    segs: List[int] = []
    done = 0
    while done < timesteps:
        step = jump_length if done + jump_length < timesteps else timesteps - done
        if segs:
            segs[-1] += step
        if step == jump_length and done + jump_length < timesteps:
            segs.extend([jump_length] * (resamplings - 1 if segs else resamplings))
        elif not segs:
            segs.append(step)
        done += step
    return segs[::-1]
