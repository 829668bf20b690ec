"""Drop-in proof with the REFERENCE'S OWN CODE.

Installs the two-line shim of INTEGRATION.md into ``sys.modules['core.quantum_circuits_tq']`` (and lifts the
torchquantum gate of utils/backend.py:26-31), then runs the reference's callers on top of this repo's module:
``core/models_pt.py:103-213`` (QuanONetPT / HEAQNNPT), ``utils/weight_transfer.load_quanonet_pt`` (:101-140),
``utils/backend.check_compatibility`` (:49-129), ``solvers/solver_pt.PTSolver`` (:26-125 construction, :191-274
train, :279-330 evaluate) and the TorchQuantum half of ``compare_backends.py:140-212`` (same shapes and seeds; the
other backends are absent, the fp64 oracle stands in for them at the reference's own tolerances, :26-31).

    python tests/harness/reference_dropin.py <reference_root> cpu|gpu <workdir>      -> one JSON line on stdout

Runs in its own process: PTSolver replaces sys.stdout (solvers/solver_pt.py:41) and the reference's packages
(`core`, `utils`, `solvers`, `data_utils`) are generic top-level names.
"""
import importlib
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def install_shim(ref_root):
    """INTEGRATION.md section 2, applied in memory instead of editing the reference tree."""
    sys.path.insert(0, ref_root)
    if ROOT not in sys.path:
        sys.path.insert(1, ROOT)
    import quanonet_b200.core.quantum_circuits_tq as ours
    shim = types.ModuleType("core.quantum_circuits_tq")
    for name in ("_TQHEACircuit", "_make_block_configs", "_ham_params", "build_quanonet_tq", "build_heaqnn_tq"):
        setattr(shim, name, getattr(ours, name))
    importlib.import_module("core")                       # the reference's package
    sys.modules["core.quantum_circuits_tq"] = shim
    sys.modules["core"].quantum_circuits_tq = shim
    import utils.backend as ref_backend                   # the reference's gate
    from quanonet_b200.utils.backend import backend as b200_backend
    type(ref_backend.backend).is_torchquantum_available = property(lambda self: b200_backend.is_b200_library_built)
    return ours, ref_backend


def main():
    ref_root, mode, workdir = sys.argv[1], sys.argv[2], sys.argv[3]
    real_stdout = sys.stdout
    res = {}
    ours, ref_backend = install_shim(ref_root)
    import numpy as np
    import torch
    import core.models_pt as ref_models
    assert os.path.realpath(ref_models.__file__).startswith(os.path.realpath(ref_root)), ref_models.__file__
    res["models_pt_file"] = os.path.realpath(ref_models.__file__)

    # ---- routing (utils/backend.py:49-129)
    res["route_quanonet_tq"] = ref_backend.backend.check_compatibility("QuanONet", quantum_backend="torchquantum")
    res["route_heaqnn_tq"] = ref_backend.backend.check_compatibility("HEAQNN", quantum_backend="torchquantum")

    # ---- construction through the reference's QuanONetPT / HEAQNNPT (core/models_pt.py:103-213)
    torch.manual_seed(42)
    m = ref_models.QuanONetPT(num_qubits=5, branch_input_size=100, trunk_input_size=2, net_size=(40, 2, 20, 2),
                              scale_coeff=0.1, if_trainable_freq=True, quantum_backend="torchquantum",
                              ham_bound=(-5.0, 5.0))
    res["quanonet_layer_class"] = type(m.quantum_layer).__module__ + "." + type(m.quantum_layer).__name__
    res["quanonet_state_dict_keys"] = sorted(m.state_dict().keys())
    res["quanonet_n_params"] = int(sum(p.numel() for p in m.parameters()))
    res["quanonet_ansatz_shape"] = list(m.quantum_layer.ansatz_weights.shape)
    h = ref_models.HEAQNNPT(num_qubits=3, input_size=12, net_size=(4, 2), scale_coeff=0.1, if_trainable_freq=True,
                            quantum_backend="torchquantum", ham_bound=(-5.0, 5.0))
    res["heaqnn_state_dict_keys"] = sorted(h.state_dict().keys())
    res["heaqnn_block_configs"] = [list(b) for b in h.quantum_layer.block_configs]

    # ---- utils/weight_transfer.load_quanonet_pt (:101-140) on the shipped Antideriv checkpoint
    from utils.weight_transfer import load_quanonet_pt
    npz = os.path.join(ref_root, "pretrained_weights", "Antideriv",
                       "Antideriv_QuanONet_Net5-1-5-1_Q2_TF_S0.001_1000x100_Seed0", "best_model.npz")
    anti = None
    if os.path.exists(npz):
        anti = load_quanonet_pt(npz, quantum_backend="torchquantum", branch_input_size=10, trunk_input_size=1,
                                net_size=(5, 1, 5, 1), num_qubits=2, scale_coeff=0.001, ham_bound=(-5.0, 5.0))
        res["antideriv_bias"] = float(anti.bias.detach().reshape(-1)[0])
        res["antideriv_ansatz_shape"] = list(anti.quantum_layer.ansatz_weights.shape)

    # ---- PTSolver construction (solvers/solver_pt.py:26-125): data, model, optimiser
    cfg = {"model_type": "QuanONet", "operator": "Antideriv", "quantum_backend": "torchquantum", "prefix": os.path.join(workdir, "outputs"),
           "num_qubits": 2, "net_size": [2, 1, 2, 1], "scale_coeff": 0.01, "if_trainable_freq": "true",
           "ham_bound": [-5, 5], "learning_rate": 1e-2, "num_epochs": 2, "batch_size": 100, "num_train": 20, "num_test": 5,
           "num_points": 10, "num_points_0": 10, "num_cal": 100, "train_sample_num": 10, "test_sample_num": 10,
           "seed": 0, "if_save": True, "gpu": 0 if mode == "gpu" else None}
    solver = None
    try:
        from solvers.solver_pt import PTSolver
        if mode != "gpu":
            torch.cuda.is_available = lambda: False           # PTSolver picks its device at :44-49
        solver = PTSolver(cfg)
        res["solver_model_class"] = type(solver.model).__module__ + "." + type(solver.model).__name__
        res["solver_layer_class"] = type(solver.model.quantum_layer).__module__
        res["solver_train_samples"] = int(solver.train_output.shape[0])
    except Exception as e:                                    # pragma: no cover - environment dependent
        res["solver_error"] = f"{type(e).__name__}: {e}"
    finally:
        sys.stdout = real_stdout

    if mode == "gpu":
        from oracle import hea_oracle as orc
        dev = torch.device("cuda:0")
        # ---- TorchQuantum half of compare_backends.py:140-212 (same shapes, seeds and tolerances)
        ATOL_PT, ATOL_GRAD_PT = 1e-4, 1e-4
        rng = np.random.default_rng(0)
        n_q, ns, b_in, t_in, batch = 2, (2, 1, 2, 1), 8, 1, 6
        torch.manual_seed(42)
        mt = ref_models.QuanONetPT(num_qubits=n_q, branch_input_size=b_in, trunk_input_size=t_in, net_size=ns,
                                   scale_coeff=0.1, if_trainable_freq=True, ham_bound=(-5.0, 5.0),
                                   quantum_backend="torchquantum").to(dev)
        mt.eval()
        branch = rng.random((batch, b_in)).astype(np.float32)
        trunk = rng.random((batch, t_in)).astype(np.float32)
        tgt = rng.random((batch, 1)).astype(np.float32)
        tb, tt, ty = (torch.tensor(a, device=dev) for a in (branch, trunk, tgt))
        with torch.no_grad():
            out = mt(tb, tt).cpu().numpy()
        params = {k: v.detach().double().cpu().numpy() for k, v in mt.state_dict().items()}
        ham = orc.ham_from_bound(n_q, -5.0, 5.0)
        e = orc.quanonet_forward(branch, trunk, params, n_q, ns, ham)
        res["cb_quanonet_fwd_maxabs"] = float(np.abs(out[:, 0] - e).max())
        mt.zero_grad()
        ((mt(tb, tt) - ty) ** 2).mean().backward()
        g_tq = mt.quantum_layer.ansatz_weights.grad.detach().double().cpu().numpy()
        blocks = orc.make_block_configs(n_q, ns[2], ns[3], ns[0], ns[1])
        xenc = np.concatenate([orc.tiled_elementwise(trunk, ns[2] * n_q, params["trunk_freq.weights"], params["trunk_freq.bias"]),
                               orc.tiled_elementwise(branch, ns[0] * n_q, params["branch_freq.weights"], params["branch_freq.bias"])], 1)
        gout = 2.0 / batch * (e - tgt[:, 0].astype(np.float64))
        _, gx_ref, gw_ref = orc.hea_forward_backward(xenc, params["quantum_layer.ansatz_weights"], n_q, blocks, ham, gout)
        res["cb_quanonet_grad_ansatz_maxabs"] = float(np.abs(g_tq - gw_ref).max())
        gfw = mt.branch_freq.weights.grad.detach().double().cpu().numpy()
        ub = np.tile(branch.astype(np.float64), (1, int(np.ceil(ns[0] * n_q / b_in))))[:, :ns[0] * n_q]
        res["cb_quanonet_grad_branch_freq_maxabs"] = float(np.abs(gfw - (gx_ref[:, ns[2] * n_q:] * ub).sum(0)).max())
        res["cb_tolerances"] = [ATOL_PT, ATOL_GRAD_PT]

        # ---- the shipped Antideriv checkpoint through the reference's loader: closed forms of ibm_inference.py:177-189
        if anti is not None:
            anti = anti.to(dev)
            xs = np.linspace(0, 1, 100)
            b = np.tile(np.cos(np.pi * np.linspace(0, 1, 10))[None], (100, 1)).astype(np.float32)
            with torch.no_grad():
                pred = anti(torch.tensor(b, device=dev), torch.tensor(xs[:, None].astype(np.float32), device=dev)).cpu().numpy()[:, 0]
            truth = np.sin(np.pi * xs) / np.pi
            res["antideriv_cos_rel_l2"] = float(np.linalg.norm(pred - truth) / np.linalg.norm(truth))

        # ---- PTSolver.train for 2 epochs + evaluate (solvers/solver_pt.py:191-330), on this repo's kernels
        if solver is not None:
            try:
                before = solver.model.quantum_layer.ansatz_weights.detach().clone()
                hist = solver.train()
                sys.stdout = real_stdout
                metrics = solver.evaluate(hist)
                sys.stdout = real_stdout
                res["solver_loss_history"] = [float(v) for v in hist["loss_train"]]
                res["solver_weights_moved"] = float((solver.model.quantum_layer.ansatz_weights.detach() - before).abs().max())
                res["solver_rel_l2"] = float(metrics["rel_l2"])
                res["solver_best_ckpt_exists"] = bool(solver.best_model_path and os.path.exists(solver.best_model_path))
                res["solver_device"] = str(next(solver.model.parameters()).device)
            except BaseException as e:                         # pragma: no cover
                sys.stdout = real_stdout
                res["solver_error"] = f"{type(e).__name__}: {e}"
    sys.stdout = real_stdout
    print("RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
