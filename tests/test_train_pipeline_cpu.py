"""Config 5 end to end on the CPU: `EnVariationalDiffusion.compute_loss(...).mean().backward()` through this repo's
`EGNNDynamics` (torch-composed wrapper: encoders / decoders / CoM removal under torch autograd) with the LEFTNet replaced by
the HOST-EMULATION build of the training kernels (csrc/train_core.h via csrc/train_emu.cpp) — against golden loss and
gradients of the UNMODIFIED reference's autograd (oracle/gen_golden.py::case_train_grad, fp64), with the reference's random
draws replayed.  Everything except the CUDA launch mapping of `par_for` / `gemm` is the code the device path runs.
Tolerance: 5e-4 of max|grad| per parameter (fp32 here, fp64 golden; the reference's own fp32 autograd is 7e-2 away)."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

import oareactdiff_b200 as ob
from oracle import oa_ref
from tests.test_train_emu import emu, _p  # noqa: F401  (fixture)
from tests.util import dyn_state_dict, load_golden


class _EmuFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lib, cfg, geo, ei, names, h, *tensors):
        N, E = h.size(0), ei.size(1)
        lib.emu_reset(N, E, cfg["hidden_channels"], cfg["num_radial"], cfg["in_hidden_channels"], cfg["num_layers"],
                      int(cfg["reflect_equiv"]))
        keep = [t.detach().to(torch.float32).contiguous() for t in tensors]
        for n, t in zip(names, keep):
            lib.emu_set_weight(n.encode(), _p(t), C.c_long(t.numel()))
        ei32, ej32 = ei[0].to(torch.int32).contiguous(), ei[1].to(torch.int32).contiguous()
        h32 = h.detach().to(torch.float32).contiguous()
        ho, dp = torch.zeros(N, h.size(1)), torch.zeros(N, 3)
        lib.emu_forward_backward(_p(ei32), _p(ej32), _p(geo["frame"]), _p(geo["rb"]), _p(geo["rbf"]), _p(geo["inv_deg"]),
                                 _p(geo["nodeframe"]), _p(geo["pos_prjt"]), _p(geo["act"]), C.c_int(geo["act"].numel()), _p(h32), _p(ho),
                                 _p(dp), None, None, None)
        ctx.saved = (lib, geo, ei32, ej32, h32, keep, names, [t.requires_grad for t in tensors])
        return ho, dp

    @staticmethod
    def backward(ctx, g_h, g_dp):
        lib, geo, ei32, ej32, h32, keep, names, req = ctx.saved
        N = h32.size(0)
        ho, dp, gin = torch.zeros_like(h32), torch.zeros(N, 3), torch.zeros_like(h32)
        gh, gd = g_h.contiguous().float(), g_dp.contiguous().float()
        # the emulation entry runs forward + backward in one call (activations live in its context)
        lib.emu_forward_backward(_p(ei32), _p(ej32), _p(geo["frame"]), _p(geo["rb"]), _p(geo["rbf"]), _p(geo["inv_deg"]),
                                 _p(geo["nodeframe"]), _p(geo["pos_prjt"]), _p(geo["act"]), C.c_int(geo["act"].numel()), _p(h32), _p(ho),
                                 _p(dp), _p(gh), _p(gd), _p(gin))
        grads = []
        for n, t, r in zip(names, keep, req):
            if not r:
                grads.append(None)
                continue
            out = torch.zeros_like(t)
            assert lib.emu_get_grad(n.encode(), _p(out), C.c_long(out.numel())) == 0, n
            grads.append(out)
        return (None, None, None, None, None, gin, *grads)


def _make_emu_leftnet(lib):
    class EmuLEFTNet(ob.LEFTNetB200):
        """LEFTNetB200's parameters and signature; arithmetic = the emulated training core; geometry from the oracle."""

        def forward(self, h, pos, edge_index, edge_attr=None, node_mask=None, edge_mask=None, update_coords_mask=None,
                    subgraph_mask=None):
            cfg = dict(self.cfg)
            sd = {k: v.detach().double() for k, v in self._oard_tensors().items()}
            dbg = {}
            with torch.no_grad():
                oa_ref.leftnet_forward(sd, cfg, h.detach().double(), pos.detach().double(), edge_index, subgraph_mask, dbg=dbg)
                pf, mask = dbg["pos_frame"], dbg["mask"].unsqueeze(-1)
                dist, cdiff, ccross, cvert = oa_ref.scalarization(pf, edge_index)
                dist = dist * mask.squeeze(-1)
                deg = torch.zeros(pos.size(0), dtype=torch.float64).index_add_(0, edge_index[0], torch.ones(edge_index.size(1), dtype=torch.float64))
                f32 = lambda t: t.to(torch.float32).contiguous()
                geo = dict(frame=f32(torch.stack((cdiff * mask, ccross * mask, cvert * mask), dim=1)),
                           rb=f32(0.5 * (torch.cos(dist * math.pi / float(cfg["cutoff"])) + 1.0)), rbf=f32(dbg["rbf"]),
                           inv_deg=f32(1.0 / deg.clamp(min=1)), nodeframe=f32(dbg["nodeframe"]), pos_prjt=f32(dbg["pos_prjt"]),
                           act=torch.nonzero(dbg["mask"] > 0).flatten().to(torch.int32).contiguous())
            names = [n for n in self._oard_tensors() if not n.startswith(("radial_emb.", "distance_embedding", "last_layer"))]
            tensors = [self._oard_tensors()[n] for n in names]
            h_out, dpos = _EmuFn.apply(lib, cfg, geo, edge_index, names, h, *tensors)
            return h_out, pos + dpos, None

    return EmuLEFTNet


class _Replay(ob.EnVariationalDiffusion):
    def _draw_t_int(self, num_sample, device):
        return self._t_int.view(num_sample, 1)

    def sample_combined_position_feature_noise(self, masks):
        self._k += 1
        return [n.clone() for n in self._noises[self._k - 1]]


@pytest.mark.parametrize("name", ["grad_small_train", "grad_small_train_t0"])  # _t0: one sample drawn at t = 0 (L0 terms active)
def test_training_step_loss_and_gradients_vs_reference_golden(emu, name):  # noqa: F811
    g = load_golden(name)
    g["node_nfs"], g["condition_nf"] = np.array([9, 9, 9]), np.int64(1)
    dyn = ob.EGNNDynamics(model_config=g["cfg"], fragment_names=["R", "TS", "P"], node_nfs=[9, 9, 9], edge_nf=0, condition_nf=1,
                          model=_make_emu_leftnet(emu), device=torch.device("cpu"))
    dyn.load_state_dict(dyn_state_dict(g), strict=True)
    dyn.model.enable_training_path = True
    sched = ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_2", int(g["T"]), 1e-5), norm_values=(1.0, 1.0, 1.0))
    ddpm = _Replay(dynamics=dyn, schdule=sched, normalizer=ob.Normalizer(), pos_only=True)
    ddpm.train(True)
    ddpm._t_int = torch.from_numpy(g["t_int"]).float()
    ddpm._noises = [[torch.from_numpy(g[f"noise{d}_{f}"]) for f in range(3)] for d in range(int(g["n_draws"]))]
    ddpm._k = 0
    sizes = torch.tensor(g["sizes"])
    reps = [{"size": sizes.clone(), "pos": torch.from_numpy(g[f"pos{f}"]), "one_hot": torch.from_numpy(g[f"one_hot{f}"]),
             "charge": torch.from_numpy(g[f"charge{f}"]), "mask": ob.get_mask_for_frag(sizes)} for f in range(3)]
    nll, info = ddpm.compute_loss((reps, torch.from_numpy(g["cond"])), scales=tuple(float(x) for x in g["scales"]), training=True)
    loss = nll.mean()
    assert loss.requires_grad
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 2e-5 * abs(float(g["loss"]))
    worst, worst_name, worst_ref, n = 0.0, "", 0.0, 0
    for pn, prm in dyn.named_parameters():
        ref = torch.from_numpy(g[f"grad/{pn}"])
        scale = float(ref.abs().max())
        if scale == 0.0:
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0, pn
            continue
        assert prm.grad is not None, pn
        n += 1
        e = float((prm.grad.double() - ref).abs().max()) / scale
        worst_ref = max(worst_ref, float((torch.from_numpy(g[f"grad_f32/{pn}"]).double() - ref).abs().max()) / scale)
        if e > worst:
            worst, worst_name = e, pn
    print(f"loss {float(loss):.7f} (reference {float(g['loss']):.7f}); worst parameter-gradient error {worst:.2e} ({worst_name}) "
          f"over {n} parameters; the reference's own fp32 autograd: {worst_ref:.2e}")
    assert worst < 5e-4 and n > 80
    # no-grad mode still returns plain values
    ddpm._k = 0
    with torch.no_grad():
        nll2, _ = ddpm.compute_loss((reps, torch.from_numpy(g["cond"])), scales=tuple(float(x) for x in g["scales"]), training=True)
    assert not nll2.requires_grad and torch.allclose(nll2, nll.detach(), rtol=1e-5)
