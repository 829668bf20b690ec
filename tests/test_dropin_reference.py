"""The drop-in boundary exercised with the REFERENCE'S OWN callers: INTEGRATION.md's two-line shim is
installed in memory and the reference's ``core/models_pt.py``, ``utils/backend.py``, ``utils/weight_transfer.py``,
``solvers/solver_pt.py`` and the TorchQuantum half of ``compare_backends.py`` run on top of this repo's module
(tests/harness/reference_dropin.py, in its own process).  Needs a reference checkout: ``/root/reference`` in the
build container, ``baseline/_ref`` next to the repo on a GPU pod; skipped with that reason otherwise."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "harness", "reference_dropin.py")


def _reference_root():
    for cand in (os.environ.get("QON_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "core", "models_pt.py")):
            return cand
    return None


def _run(mode, tmp_path):
    ref = _reference_root()
    if ref is None:
        pytest.skip("no reference checkout (/root/reference or baseline/_ref with core/models_pt.py): "
                    "the reference's own callers cannot be imported here")
    r = subprocess.run([sys.executable, HARNESS, ref, mode, str(tmp_path)], capture_output=True, text=True, timeout=900,
                       cwd=str(tmp_path))
    lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    assert r.returncode == 0 and lines, (r.returncode, r.stdout[-2000:], r.stderr[-4000:])
    return json.loads(lines[-1][len("RESULT "):]), ref


def _check_construction(res, ref):
    assert res["models_pt_file"].startswith(os.path.realpath(ref))                 # the reference's file, not our mirror
    assert res["route_quanonet_tq"] == "pytorch_quantum" and res["route_heaqnn_tq"] == "pytorch_quantum"
    assert res["quanonet_layer_class"] == "quanonet_b200.core.quantum_circuits_tq._TQHEACircuit"
    assert res["quanonet_state_dict_keys"] == ["bias", "branch_freq.bias", "branch_freq.weights",
                                               "quantum_layer.ansatz_weights", "trunk_freq.bias", "trunk_freq.weights"]
    assert res["quanonet_n_params"] == 2401 and res["quanonet_ansatz_shape"] == [120, 3, 5]     # SURVEY §8 header
    assert res["heaqnn_state_dict_keys"] == ["freq.bias", "freq.weights", "quantum_layer.ansatz_weights"]   # no bias (:205-213)
    assert res["heaqnn_block_configs"] == [[3, 2]] * 4
    if "antideriv_bias" in res:
        assert abs(res["antideriv_bias"] - 0.032724) < 1e-6 and res["antideriv_ansatz_shape"] == [10, 3, 2]
    assert "solver_error" not in res, res["solver_error"]
    assert res["solver_model_class"] == "core.models_pt.QuanONetPT"
    assert res["solver_layer_class"] == "quanonet_b200.core.quantum_circuits_tq"


def test_reference_callers_construct_on_the_dropin_module(tmp_path):
    """No GPU needed: routing, construction, state_dict keys, checkpoint loading and PTSolver set-up through the
    reference's own code."""
    res, ref = _run("cpu", tmp_path)
    _check_construction(res, ref)


@pytest.mark.gpu
def test_reference_callers_run_on_the_b200_kernels(tmp_path):
    """The reference's QuanONetPT forward/backward (compare_backends.py shapes, its own tolerances), the shipped
    Antideriv checkpoint through utils/weight_transfer.load_quanonet_pt, and PTSolver.train / evaluate for 2 epochs."""
    res, ref = _run("gpu", tmp_path)
    _check_construction(res, ref)
    atol, atol_grad = res["cb_tolerances"]
    assert res["cb_quanonet_fwd_maxabs"] < atol
    assert res["cb_quanonet_grad_ansatz_maxabs"] < atol_grad and res["cb_quanonet_grad_branch_freq_maxabs"] < atol_grad
    if "antideriv_cos_rel_l2" in res:
        assert abs(res["antideriv_cos_rel_l2"] - 0.026915) < 2e-4
    assert res["solver_device"].startswith("cuda")
    assert len(res["solver_loss_history"]) == 2 and all(v == v and v < 1e3 for v in res["solver_loss_history"])
    assert res["solver_weights_moved"] > 0 and res["solver_best_ckpt_exists"]
