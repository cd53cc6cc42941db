"""The plugin seam with the UNMODIFIED reference on the other side (build container only: skipped where /root/reference
is absent): oracle/plug_into_reference.py hands `LEFTNetB200` to the reference's own `EGNNDynamics(model=<class>)` and lets
the reference's own `sample()` drive it.  Run in a subprocess — the shims for torch_scatter / torch_geometric must not leak
into this test session."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/oa_reactdiff"), reason="the reference only exists in the build container")
def test_leftnetb200_plugs_into_the_unmodified_reference():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "plug_into_reference.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["model_class"] == "LEFTNetB200" and out["missing"] == [] and out["unexpected"] == []
    assert out["h_equal"] and max(out["trajectory_rel_err"]) < 1e-4, out


@pytest.mark.skipif(not os.path.isdir("/root/reference/oa_reactdiff"), reason="the reference only exists in the build container")
def test_the_reference_own_test_files_pass_with_the_plugin_class():
    """oracle/run_reference_tests.py: tests/model/test_equiv.py, test_subgraphs.py, tests/dynamics/test_switch_fragments.py and
    test_egnn_dynamics.py of the reference, unmodified, with `oa_reactdiff.model.LEFTNet` replaced by `LEFTNetB200` (oracle as
    engine; the confidence head keeps the reference's class), plus tests/utils/test_graph_tools.py and tests/datasets/
    test_transition1x.py with this package's graph helpers and packed dataset swapped in: the reference's whole suite."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_reference_tests.py")], capture_output=True, text=True,
                       timeout=900, cwd="/tmp")
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["rc"] == 0 and out["failed"] == 0 and out["passed"] >= 24, (out, r.stdout[-3000:])


@pytest.mark.skipif(not os.path.isdir("/root/reference/oa_reactdiff"), reason="the reference only exists in the build container")
def test_graph_helpers_bit_exact_vs_the_unmodified_reference_on_random_inputs():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "fuzz_graph_tools.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["n_mismatches"] == 0 and out["checks"] > 2000, out


@pytest.mark.skipif(not os.path.isdir("/root/reference/oa_reactdiff"), reason="the reference only exists in the build container")
def test_loss_terms_match_the_unmodified_reference_over_its_option_space():
    """oracle/fuzz_loss_options.py: loss_type {l2, vlb} x pos_only x training / eval x fixed_idx, native random streams."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "fuzz_loss_options.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["cases"] == 16 and out["bad"] == 0 and out["worst"] < 1e-5, out


@pytest.mark.skipif(not os.path.isdir("/root/reference/oa_reactdiff"), reason="the reference only exists in the build container")
def test_samplers_match_the_unmodified_reference_over_their_option_space():
    """oracle/fuzz_sampler_options.py: sample (pos_only x return_frames x timesteps x fixed_idx) and inpaint (pos_only x
    resamplings / jump_length x frag_fixed x timesteps), native random streams, both host formulations of this package."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "fuzz_sampler_options.py")], capture_output=True, text=True,
                       timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["cases"] == 40 and out["bad"] == 0 and out["worst"] < 1e-5, out


@pytest.mark.skipif(not os.path.isdir("/root/reference/oa_reactdiff"), reason="the reference only exists in the build container")
def test_dynamics_wrapper_matches_the_unmodified_reference_over_its_option_space():
    """oracle/fuzz_dynamics_options.py: condition_time x condition_nf x t layout x per-fragment node_nf x fragment sets (incl. an
    empty fragment) x shared encoders, weights handed over through `source`."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "fuzz_dynamics_options.py")], capture_output=True, text=True,
                       timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["cases"] >= 100 and out["bad"] == 0 and out["worst"] < 1e-6, out


@pytest.mark.skipif(not os.path.isdir("/root/reference/oa_reactdiff"), reason="the reference only exists in the build container")
def test_oracle_leftnet_matches_the_unmodified_reference_over_graphs_and_options():
    """oracle/fuzz_oracle_leftnet.py: 576 float64 cases (graph shapes incl. sparse / disconnected / shuffled edge lists, masks,
    cut-offs that split groups, reflect_equiv, object_aware, update, depth) — the checker of the kernels pinned on the reference
    itself, beyond the committed golden vectors."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "fuzz_oracle_leftnet.py")], capture_output=True, text=True,
                       timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["cases"] >= 500 and out["bad"] == 0 and out["worst"] < 1e-9, out


@pytest.mark.skipif(not os.path.isdir("/root/reference/oa_reactdiff"), reason="the reference only exists in the build container")
def test_plugin_class_inside_the_reference_real_trainer_module():
    """oracle/trainer_seam.py: the reference's `DDPMModule` (pl_trainer.py, unmodified; Lightning / torchmetrics stand-ins) built
    with `model=LEFTNetB200` from train_ts1x.py's configuration: same checkpoint keys, `compute_loss` over loss_type x pos_only x
    mode equal to the `model=LEFTNet` module AND to this package's own `compute_loss` on the packed batch producer,
    `training_step` / `validation_step` / the sampling half of `eval_inplaint_batch` equal."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "trainer_seam.py")], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["compute_loss_cases"] == 8 and out["worst"] < 1e-6, out
    assert all(c["finite"] and c["info_keys_equal"] for c in out["report"])
    assert max(out["training_step"], out["validation_step"], out["eval_inpaint_samples"]) < 1e-6


@pytest.mark.skipif(not os.path.isfile("/root/reference/oa_reactdiff/data/transition1x/train.pkl"), reason="needs the Transition1x file shipped with the reference")
def test_packed_dataset_bit_exact_on_the_real_transition1x_file():
    """oracle/fuzz_dataset.py: 13 466 (trainer options) / 10 073 reactions; random items and collated batches, dtypes included
    (the shipped file stores atomic numbers as int32 arrays, which the reference's `charge` feature inherits)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "fuzz_dataset.py")], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["n_mismatches"] == 0 and out["checks"] >= 300 and out["lens"]["trainer"] == [13466, 13466], out


@pytest.mark.skipif(not os.path.isdir("/root/reference/oa_reactdiff"), reason="the reference only exists in the build container")
def test_default_initialisation_scheme_matches_the_reference_constructors():
    """oracle/init_parity.py: trained configuration, 246 tensors — constants (zero biases, LayerNorm, RBF buffers) equal, large
    random tensors agree in spread and bound: training from scratch starts from the same distribution."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "init_parity.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["n_bad"] == 0 and out["tensors"] == 246 and out["random_checked"] > 50 and out["constants_checked"] > 20, out
