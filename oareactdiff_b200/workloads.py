"""Synthetic Transition1x-shaped sampling inputs (the role of the reference's `utils/sampling_tools.py:64-108`
`assemble_sample_inputs`): fragments_nodes = [n, n, n] per reaction, h0 = [one-hot(5) | atomic number] drawn from
{H, C, N, O} with Transition1x frequencies, conditions = zeros.  The atom-count histogram is the one of the 9000
`use_ind` reactions of Transition1x (min 4, max 23, mean 13.57) so no dataset is needed at run time."""
from typing import List, Tuple

import numpy as np
import torch

T1X_NATOMS_HIST = [0, 0, 0, 0, 1, 4, 15, 36, 157, 261, 599, 947, 1155, 1207, 1461, 1051, 907, 636, 200, 280, 25, 55, 0, 3]
ELEMENT_Z = np.array([1, 6, 7, 8, 9])
ELEMENT_P = np.array([0.445, 0.290, 0.149, 0.116, 0.0])


def t1x_sizes(batch: int, seed: int = 0) -> List[int]:
    p = np.asarray(T1X_NATOMS_HIST, dtype=np.float64)
    return [int(x) for x in np.random.RandomState(seed).choice(len(p), size=batch, p=p / p.sum())]


def reaction_batch(sizes: List[int], seed: int = 0) -> Tuple[List[torch.Tensor], List[torch.Tensor], torch.Tensor]:
    """-> (fragments_nodes [3 x LongTensor[B]], h0 [3 x FloatTensor[sum n, 6]], conditions FloatTensor[B, 1]) on CPU."""
    rng = np.random.RandomState(seed)
    nodes = torch.tensor(list(sizes), dtype=torch.long)
    types = np.concatenate([rng.choice(5, size=n, p=ELEMENT_P) for n in sizes])
    h = np.zeros((types.size, 6), dtype=np.float32)
    h[np.arange(types.size), types] = 1
    h[:, 5] = ELEMENT_Z[types]
    h0 = torch.from_numpy(h)
    return [nodes, nodes.clone(), nodes.clone()], [h0, h0.clone(), h0.clone()], torch.zeros(len(sizes), 1)


def edge_count(sizes: List[int], n_frag: int = 3) -> int:
    return sum(n_frag * n * (n_frag * n - 1) for n in sizes)


def active_edge_count(sizes: List[int], n_frag: int = 3) -> int:
    """Same-fragment directed edges (the most the cutoff mask can leave active)."""
    return sum(n_frag * n * (n - 1) for n in sizes)


# ------------------------------------------------------------------------------------------- geometries and drivers (bench)
def geometry_fixture_path() -> str:
    import os
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "t1x_geometries_b512.npz")


def real_geometries(lo: int, hi: int, sizes: List[int]) -> List[torch.Tensor]:
    """Reactant / transition-state / product coordinates [N_f, 3] of REAL Transition1x reactions whose atom counts are
    t1x_sizes(., seed=0)[lo:hi] (frozen fixture, generator oracle/gen_t1x_geometries.py; 512 reactions)."""
    fx = np.load(geometry_fixture_path())
    if hi > len(fx["sizes"]) or [int(v) for v in fx["sizes"][lo:hi]] != list(sizes):
        raise ValueError("the geometry fixture holds the first 512 reactions of t1x_sizes(., seed=0) only")
    off = np.concatenate([[0], np.cumsum(fx["sizes"])])
    return [torch.from_numpy(np.ascontiguousarray(fx[k][off[lo]:off[hi]])).float() for k in ("reactant", "transition_state", "product")]


def synthetic_geometries(sizes: List[int], seed: int = 0) -> List[torch.Tensor]:
    """Compact random clouds of molecular density, three jittered copies (for batches the real fixture does not cover)."""
    gen = torch.Generator().manual_seed(99 + seed)
    geo = []
    for n in sizes:
        r = 1.2 * n ** (1.0 / 3.0)
        pts = torch.randn(n, 3, generator=gen)
        pts = pts / pts.norm(dim=1, keepdim=True) * (torch.rand(n, 1, generator=gen) ** (1 / 3)) * r
        geo.append(pts - pts.mean(0, keepdim=True))
    out = []
    for _ in range(3):
        xs = [gp + 0.3 * torch.randn(gp.shape, generator=gen) for gp in geo]
        out.append(torch.cat([x - x.mean(0, keepdim=True) for x in xs]))
    return out


@torch.no_grad()
def replay_trajectory(ddpm, n_samples: int, fragments_nodes, conditions, h0, x_ref, timesteps=None, probe=None):
    """A full reverse diffusion (T reverse steps + the p(x | z0) decode = T + 1 denoiser evaluations, posterior sampling,
    CoM projections — the per-step work of `EnVariationalDiffusion.sample`) on the states a TRAINED model would visit: before
    every step the state is re-drawn from q(z_t | x_ref) = alpha_t x_ref + sigma_t eps.  With random weights the literal
    sample() drifts to |x| ~ 1e2 A, the 10 A cutoff empties and the message-passing stages have nothing to do; here every
    same-fragment edge of the reference geometries stays inside the cutoff (active fraction ~0.32 on Transition1x).
    x_ref / h0: per-fragment [N_f, 3] / [N_f, nf - 3] on the device.  Returns the decoded positions per fragment.
    probe: optional list; a CUDA event is recorded into it every 100 reverse steps (bench.py: is a slow trajectory uniformly
    slow — clocks — or slow in bursts — stalls?)."""
    T = ddpm.T if timesteps is None else timesteps
    masks, edge_index, nfs = ddpm._setup(n_samples, fragments_nodes)
    dev = x_ref[0].device
    tab = ddpm._tables(T, dev)
    ddpm._seg_setup(masks)
    H0 = torch.cat(h0).to(torch.float32)
    X = torch.cat([torch.cat([x.to(torch.float32), h], dim=1) for x, h in zip(x_ref, h0)])
    X[:, :3] = ddpm._remove_mean_cat(X[:, :3])
    on_device = ddpm._device_ok(dev)
    Z = torch.empty_like(X)
    if on_device:
        ddpm._device_setup(Z, masks, edge_index, nfs, conditions, H0)
    for s_int in reversed(range(T)):
        if probe is not None and on_device and (T - 1 - s_int) % 100 == 0:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            probe.append(ev)
        Zt = tab["alpha"][s_int + 1] * X + tab["sigma_abs"][s_int + 1] * ddpm._noise_cat(masks)
        Zt[:, 3:] = H0
        if on_device:  # the device step replays one CUDA graph on a persistent state buffer
            Z.copy_(Zt)
            ddpm._device_step(s_int, Z, tab)
        else:
            Z = ddpm._fast_step(s_int, Zt, tab, edge_index, nfs, masks, conditions)
    Z0 = tab["alpha"][0] * X + tab["sigma_abs"][0] * ddpm._noise_cat(masks)
    Z0[:, 3:] = H0
    return ddpm.sample_p_xh_given_z0(ddpm._views(Z0), edge_index, nfs, masks, n_samples, conditions)[0]


def training_batch(sizes: List[int], x_ref, h0, device):
    """(representations, conditions) in the collated form `DDPMModule.training_step` receives (dataset/base_dataset.py +
    collate_fn): per fragment size / pos / one_hot / charge / mask."""
    from .graph_tools import get_mask_for_frag
    n = torch.tensor(list(sizes), dtype=torch.long, device=device)
    mask = get_mask_for_frag(n)
    reps = [{"size": n.clone(), "pos": x.to(device).float(), "one_hot": h[:, :-1].to(device).float(),
             "charge": h[:, -1:].to(device).float(), "mask": mask} for x, h in zip(x_ref, h0)]
    return reps, torch.zeros(len(sizes), 1, device=device)
