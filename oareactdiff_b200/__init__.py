"""oareactdiff_b200 — B200-native drop-in for OA-ReactDiff's denoising hot path.

Public names mirror the reference (`oa_reactdiff.model.LEFTNet`, `oa_reactdiff.dynamics.EGNNDynamics`,
`oa_reactdiff.diffusion.*`, `oa_reactdiff.utils.*`).  All LEFTNet arithmetic runs in liboard_b200.so (CUDA, sm_100a,
C ABI in include/oard.h); importing the package without the built library raises.
"""
from . import _lib

_lib.load()  # fail loudly if the CUDA library has not been built: there is no fallback path

from .leftnet import LEFTNetB200  # noqa: E402
from .dynamics import EGNNDynamics  # noqa: E402
from .diffusion import EnVariationalDiffusion  # noqa: E402
from .schedule import (DiffSchedule, PredefinedNoiseSchedule, get_repaint_schedule, polynomial_schedule,  # noqa: E402
                       cosine_beta_schedule, ccosine_schedule, linear_schedule, clip_noise_schedule)
from .normalizer import Normalizer  # noqa: E402
from .graph_tools import (get_edges_index, get_mask_for_frag, get_n_frag_switch, get_subgraph_mask,  # noqa: E402
                          get_inner_edge_index)

LEFTNet = LEFTNetB200

__all__ = ["LEFTNetB200", "LEFTNet", "EGNNDynamics", "EnVariationalDiffusion", "DiffSchedule",
           "PredefinedNoiseSchedule", "get_repaint_schedule", "Normalizer", "get_edges_index", "get_mask_for_frag",
           "get_n_frag_switch", "get_subgraph_mask", "get_inner_edge_index", "polynomial_schedule", "cosine_beta_schedule",
           "ccosine_schedule", "linear_schedule", "clip_noise_schedule"]
from .data import (ProcessedTS1x, assemble_sample_inputs, write_single_xyz, write_tmp_xyz, set_new_schedule,  # noqa: E402,F401
                   inplaint_batch, samples_to_pos_charge)
from .checkpoint import load_reference_checkpoint  # noqa: E402,F401
