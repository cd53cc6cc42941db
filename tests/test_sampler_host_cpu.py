"""CPU check of the HOST side of the samplers: `EnVariationalDiffusion.sample / inpaint / inpaint_fixed` and the
torch-composed `EGNNDynamics.forward` of this package, with the fp64 oracle standing in for the CUDA engine behind
`LEFTNetB200.forward`, against golden trajectories of the UNMODIFIED reference (oracle/gen_golden.py::case_sample /
case_inpaint; the reference's noise stream is reproduced draw for draw).  Both host formulations are covered: the tabulated
fast path the product uses and the reference-structured per-fragment path.  The kernels themselves are judged by the GPU
twins of these tests (tests/test_gpu_parity.py), whose bodies are reused here."""
import pytest
import torch

import oareactdiff_b200 as ob
import tests.test_gpu_parity as gp
from oracle import oa_ref
from tests.test_reference_suite_cpu import _oracle_forward


@pytest.fixture()
def oracle_engine(monkeypatch):
    monkeypatch.setattr(gp, "DEV", torch.device("cpu"))
    monkeypatch.setattr(ob.LEFTNetB200, "forward", _oracle_forward)
    monkeypatch.setattr(ob.EGNNDynamics, "fused_ok", lambda self, device: False)
    # the GPU twins draw the noise on the CPU through a subclass, which (by design) switches the tabulated fast path off; on the
    # CPU the package's own noise function already produces the reference's stream, so the plain class is used and the
    # `fast` parameter below really selects between the two host formulations
    monkeypatch.setattr(gp, "_CpuNoiseDiffusion", ob.EnVariationalDiffusion)


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("name", ["sample_small_T10", "sample_trained_cfg1_T10"])
def test_sample_trajectory_host_logic(oracle_engine, monkeypatch, name, fast):
    if not fast:
        monkeypatch.setattr(ob.EnVariationalDiffusion, "_fast_ok", lambda self: False)
    gp.test_sample_trajectory_vs_reference_golden(name)


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("entry,name", [("inpaint", "inpaint_small_T12_r2_j3"), ("inpaint_fixed", "inpaint_small_T12_r2_j3"),
                                        ("inpaint", "inpaint_trained_b4_T12_r2_j3")])
def test_inpaint_trajectory_host_logic(oracle_engine, monkeypatch, fast, entry, name):
    if not fast:
        monkeypatch.setattr(ob.EnVariationalDiffusion, "_fast_ok", lambda self: False)
    if entry == "inpaint_fixed":  # the reference's second entry point with the same body (en_diffusion.py:887-1048)
        real = ob.EnVariationalDiffusion.inpaint
        calls = []

        def via_fixed(self, *a, **k):
            if not calls:  # route the test's call through inpaint_fixed once; it lands in the real inpaint
                calls.append(1)
                return ob.EnVariationalDiffusion.inpaint_fixed(self, *a, **k)
            return real(self, *a, **k)
        monkeypatch.setattr(ob.EnVariationalDiffusion, "inpaint", via_fixed)
    gp.test_inpaint_trajectory_vs_reference_golden(name)


def test_schedule_swapped_after_construction_is_honoured(oracle_engine, monkeypatch):
    """The reference's callers replace `ddpm.schedule` / `ddpm.T` on a built model (evaluate/utils.py:14-31,
    pl_trainer.py:284-293: trained with one schedule, sampled with another): the tabulated fast path must follow — it is
    compared with the reference-structured path, which reads the schedule directly."""
    g = gp.load_golden("sample_small_T10")
    sizes = [int(x) for x in g["sizes"]]
    ddpm = gp._make_ddpm(g["cfg"], int(g["seed"]), 10, gp._CpuNoiseDiffusion)
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, int(g["seed"]))

    def run():
        torch.manual_seed(3)
        out, _ = ddpm.sample(len(sizes), nodes, cond, h0=h0)
        return torch.cat([o[:, :3] for o in out[0]])

    first = run()
    ddpm.schedule = ob.DiffSchedule(ob.PredefinedNoiseSchedule("polynomial_3", 6, 1e-5), norm_values=ddpm.norm_values)
    ddpm.T = 6
    fast = run()
    assert ddpm.n_evals == 7 and not torch.allclose(fast, first)
    monkeypatch.setattr(ob.EnVariationalDiffusion, "_fast_ok", lambda self: False)
    structured = run()
    assert torch.allclose(fast, structured, rtol=0, atol=1e-5 * float(structured.abs().max()))


def test_inpaint_with_fewer_steps_than_the_schedule_jumps_back_correctly(oracle_engine, monkeypatch):
    """`timesteps` below the schedule's T (en_diffusion.py:736) with RePaint jump-backs: the forward jump z_s -> z_t needs
    gamma at s / timesteps and t / timesteps, not at the raw step indices (regression: the tabulated path used the latter)."""
    g = gp.load_golden("inpaint_small_T12_r2_j3")
    sizes = [int(x) for x in g["sizes"]]
    ddpm = gp._make_ddpm(g["cfg"], int(g["seed"]), int(g["T"]), gp._CpuNoiseDiffusion)
    nodes, h0, cond = oa_ref.synthetic_batch(len(sizes), sizes, int(g["seed"]))
    xh_fixed = [torch.from_numpy(g[f"xh_fixed{f}"]) for f in range(3)]

    def run():
        torch.manual_seed(5)
        out, _ = ddpm.inpaint(len(sizes), nodes, cond, resamplings=3, jump_length=2, timesteps=6,
                              xh_fixed=[x.clone() for x in xh_fixed], frag_fixed=[0, 2])
        return torch.cat([o[:, :3] for o in out[0]])

    fast = run()
    monkeypatch.setattr(ob.EnVariationalDiffusion, "_fast_ok", lambda self: False)
    structured = run()
    assert torch.allclose(fast, structured, rtol=0, atol=1e-5 * float(structured.abs().max()))
