"""Feature normalizer (oa_reactdiff/diffusion/_normalizer.py); identity with the default values."""
from typing import Dict, List, Tuple

from torch import Tensor, nn

FEATURE_MAPPING = ["pos", "one_hot", "charge"]


class Normalizer(nn.Module):
    def __init__(self, norm_values: Tuple = (1.0, 1.0, 1.0), norm_biases: Tuple = (0.0, 0.0, 0.0), pos_dim: int = 3):
        super().__init__()
        self.norm_values, self.norm_biases, self.pos_dim = norm_values, norm_biases, pos_dim

    def normalize(self, representations: List[Dict]) -> List[Dict]:
        for rep in representations:
            for k, name in enumerate(FEATURE_MAPPING):
                rep[name] = (rep[name] - self.norm_biases[k]) / self.norm_values[k]
        return representations

    def unnormalize(self, x: Tensor, ind: int) -> Tensor:
        return x * self.norm_values[ind] + self.norm_biases[ind]

    def unnormalize_z(self, z_combined: List[Tensor]) -> List[Tensor]:
        p = self.pos_dim
        for z in z_combined:
            z[:, :p] = self.unnormalize(z[:, :p], 0)
            z[:, p:-1] = self.unnormalize(z[:, p:-1], 1)
            z[:, -1:] = self.unnormalize(z[:, -1:], 2)
        return z_combined
