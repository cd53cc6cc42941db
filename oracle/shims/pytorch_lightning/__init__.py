"""Test-infrastructure shim: `seed_everything` only (reference tests import it)."""
import random

import numpy as np
import torch


def seed_everything(seed=0, workers=False):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed
