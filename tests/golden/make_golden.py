#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/.

Runs ONLY in the build container, where the reference checkout exists at /root/reference
(it does not exist on the GPU box, so nothing in tests/ or bench.py reads it at run time).

What comes from the reference itself (imported, never copied):
  * the four shipped checkpoints under ``pretrained_weights/`` (weights = fixture data);
  * its PDE solvers ``data_utils/data_generation.py:224-352`` → ground-truth grids of the
    notebook demo (``visualization.ipynb`` cell 7);
  * its wrapper modules ``core/models_pt.py`` (``QuanONetPT`` / ``HEAQNNPT``; importable without
    any simulator) — run here with the fp64 oracle patched in as the quantum layer, so the golden
    outputs pin the frequency layer / concatenation order / bias handling of the REAL wrapper;
  * ``utils/weight_transfer.ms_npz_to_pt_state_dict`` for the ``.npz`` checkpoint.
What is restated: the circuit arithmetic (oracle/hea_oracle.py), because torchquantum /
mindquantum are not installed and cannot be (no network).  The restatement is pinned by the
published numbers recorded in ``published.json`` (tests/test_oracle_golden.py).

Usage:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(1, REF)

from oracle import hea_oracle as orc  # noqa: E402
from quanonet_b200.checkpoint import load_raw, ms_to_pt_arrays, parse_experiment_dir  # noqa: E402

CKPTS = {
    "Antideriv": "pretrained_weights/Antideriv/Antideriv_QuanONet_Net5-1-5-1_Q2_TF_S0.001_1000x100_Seed0/best_model.npz",
    "Advection": "pretrained_weights/Advection/Advection_QuanONet_Net40-2-20-2_Q5_TF_S0.1_1000x100_Seed0/best_model.ckpt",
    "RDiffusion": "pretrained_weights/RDiffusion/RDiffusion_QuanONet_Net40-2-20-2_Q5_TF_S0.1_1000x100_Seed0/best_model.ckpt",
    "Darcy": "pretrained_weights/Darcy/Darcy_QuanONet_Net40-2-20-2_Q5_TF_S0.1_1000x25_Seed0/best_model.ckpt",
}

# Figure titles of visualization.ipynb (outputs at :162,172,197,207,232,242), "MSE=…"/"MAE=…" with
# the notebook's '.1e' formatting.
PUBLISHED = {
    "Advection/sin2": {"mse": "3.0e-03", "mae": "4.5e-02"},
    "Advection/sin4": {"mse": "1.2e-02", "mae": "8.8e-02"},
    "RDiffusion/sin2": {"mse": "1.0e-04", "mae": "8.1e-03"},
    "RDiffusion/sin4": {"mse": "7.0e-04", "mae": "2.1e-02"},
    "Darcy/sin2": {"mse": "7.6e-04", "mae": "2.1e-02"},
    "Darcy/sin4": {"mse": "9.2e-03", "mae": "7.7e-02"},
}


def weights_fixture():
    out = {}
    for name, rel in CKPTS.items():
        cfg = parse_experiment_dir(rel)
        sd = ms_to_pt_arrays(load_raw(os.path.join(REF, rel)), cfg["net_size"], cfg["num_qubits"])
        for k, v in sd.items():
            out[f"{name}/{k}"] = v
    # cross-check the .npz reader against the reference's own transfer function
    from utils.weight_transfer import ms_npz_to_pt_state_dict as ref_transfer
    ref_sd = ref_transfer(os.path.join(REF, CKPTS["Antideriv"]), net_size=(5, 1, 5, 1), num_qubits=2)
    for k, v in ref_sd.items():
        assert np.array_equal(v.numpy(), out[f"Antideriv/{k}"]), k
    np.savez_compressed(os.path.join(HERE, "pretrained.npz"), **out)
    return out


def params_of(weights, name):
    return {k.split("/", 1)[1]: v for k, v in weights.items() if k.startswith(name + "/")}


def notebook_demo(weights):
    """visualization.ipynb cell 7, MindQuantum replaced by the fp64 oracle."""
    from data_utils.data_generation import solve_advection_pde, solve_darcy_pde, solve_rdiffusion_pde

    solvers = {
        "Advection": (solve_advection_pde, {"c": 1.0}),
        "RDiffusion": (solve_rdiffusion_pde, {"D": 0.01, "k": 0.01}),
        "Darcy": (solve_darcy_pde, {"K": 0.1, "f": -1.0}),
    }
    inputs = {"sin2": lambda x: np.sin(2 * np.pi * x), "sin4": lambda x: np.sin(4 * np.pi * x)}
    fx = {}
    report = {}
    ham = orc.ham_from_bound(5, -5.0, 5.0)
    for op, (solver, args) in solvers.items():
        P = 25 if op == "Darcy" else 100
        x0 = np.linspace(0, 1, 100).astype(np.float32)
        x = np.linspace(0, 1, P).astype(np.float32)
        X, T = np.meshgrid(x, x)
        trunk = np.hstack((X.flatten()[:, None], T.flatten()[:, None])).astype(np.float32)
        params = params_of(weights, op)
        for tag, f in inputs.items():
            u0 = f(x0)
            truth = solver(P, 0.2, u0_cal=f(np.linspace(0, 1, 100).astype(np.float32)), **args)[0].T
            branch = np.tile(u0, (trunk.shape[0], 1)).astype(np.float32)
            pred = orc.quanonet_forward(branch, trunk, params, 5, (40, 2, 20, 2), ham).reshape(P, P)
            diff = truth - pred
            mse, mae = float(np.mean(diff ** 2)), float(np.mean(np.abs(diff)))
            key = f"{op}/{tag}"
            report[key] = {"mse": mse, "mae": mae, "mse_1e": f"{mse:.1e}", "mae_1e": f"{mae:.1e}",
                           "rel_l2": float(np.linalg.norm(diff) / np.linalg.norm(truth))}
            fx[f"{key}/u0"] = u0.astype(np.float32)
            fx[f"{key}/truth"] = truth.astype(np.float64)
            fx[f"{key}/pred_fp64"] = pred.astype(np.float64)
            print(key, report[key], "published", PUBLISHED[key])
            assert report[key]["mse_1e"] == PUBLISHED[key]["mse"], key
            assert report[key]["mae_1e"] == PUBLISHED[key]["mae"], key
    np.savez_compressed(os.path.join(HERE, "notebook_demo.npz"), **fx)
    return report


def antideriv_closed_forms(weights):
    """ibm_inference.py:177-189 closed-form inputs through the Antideriv Q2 weights."""
    params = params_of(weights, "Antideriv")
    ham = orc.ham_from_bound(2, -5.0, 5.0)
    xs = np.linspace(0, 1, 100)
    trunk = xs[:, None].astype(np.float32)
    cases = {
        "cos": (np.cos(np.pi * np.linspace(0, 1, 10)), np.sin(np.pi * xs) / np.pi),
        "lin": (np.linspace(0, 1, 10), xs ** 2 / 2),
    }
    out = {}
    fx = {}
    for tag, (u0, truth) in cases.items():
        branch = np.tile(u0, (100, 1)).astype(np.float32)
        pred = orc.quanonet_forward(branch, trunk, params, 2, (5, 1, 5, 1), ham)
        diff = pred - truth
        out[tag] = {"rel_l2": float(np.linalg.norm(diff) / np.linalg.norm(truth)),
                    "mse": float(np.mean(diff ** 2))}
        fx[f"{tag}/branch"] = branch
        fx[f"{tag}/trunk"] = trunk
        fx[f"{tag}/truth"] = truth
        fx[f"{tag}/pred_fp64"] = pred
        print("Antideriv", tag, out[tag])
    np.savez_compressed(os.path.join(HERE, "antideriv_closed_form.npz"), **fx)
    return out


def wrapper_goldens(weights):
    """Run the REFERENCE's QuanONetPT / HEAQNNPT (core/models_pt.py) with the fp64 oracle patched
    in as the quantum layer, on compare_backends.py-shaped cases, and record outputs + gradients of
    ``((model(x)-tgt)**2).mean()`` (compare_backends.py:188-199,259-268,355-376)."""
    import torch
    import core.models_pt as ref_models
    import core.quantum_circuits_tq as ref_tq

    class _OracleFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, mod):
            ctx.mod = mod
            ctx.save_for_backward(x, w)
            e = orc.hea_forward(x.detach().double().numpy(), w.detach().double().numpy(),
                                mod.n_wires, mod.block_configs, mod._ham)
            return torch.from_numpy(e).reshape(-1, 1)

        @staticmethod
        def backward(ctx, g):
            x, w = ctx.saved_tensors
            mod = ctx.mod
            _, gx, gw = orc.hea_forward_backward(x.detach().double().numpy(), w.detach().double().numpy(),
                                                 mod.n_wires, mod.block_configs, mod._ham,
                                                 grad_out=g.double().numpy().reshape(-1))
            return torch.from_numpy(gx), torch.from_numpy(gw), None

    class _OracleCircuit(torch.nn.Module):
        """Same constructor as the reference's _TQHEACircuit (quantum_circuits_tq.py:39-63), in
        float64, evaluating through the oracle."""

        def __init__(self, n_wires, block_configs, ham_offset=0.0, ham_coeff_per_qubit=0.0, ham_diag=None):
            super().__init__()
            self.n_wires = n_wires
            self.block_configs = block_configs
            s = sum(d for _, d in block_configs)
            self.ansatz_weights = torch.nn.Parameter(torch.empty(s, 3, n_wires))
            torch.nn.init.uniform_(self.ansatz_weights, -np.pi, np.pi)
            if ham_diag is not None:
                self._ham = orc.ham_from_diag(ham_diag, n_wires, order="msb0")
            else:
                self._ham = orc.Ham("pauli", "Z", float(ham_offset), float(ham_coeff_per_qubit))

        def forward(self, x):
            return _OracleFn.apply(x, self.ansatz_weights, self)

    ref_tq._TQHEACircuit = _OracleCircuit  # the reference's builders now produce oracle circuits
    fx = {}

    def record(tag, model, inputs, tgt):
        model = model.double()
        inputs = [torch.tensor(a, dtype=torch.float64) for a in inputs]
        tgt_t = torch.tensor(tgt, dtype=torch.float64)
        out = model(*inputs)
        loss = ((out - tgt_t) ** 2).mean()
        model.zero_grad()
        loss.backward()
        for i, a in enumerate(inputs):
            fx[f"{tag}/in{i}"] = a.numpy().astype(np.float32)
        fx[f"{tag}/tgt"] = tgt.astype(np.float32)
        fx[f"{tag}/out_fp64"] = out.detach().numpy()
        fx[f"{tag}/loss_fp64"] = np.array(loss.item())
        for k, p in model.named_parameters():
            fx[f"{tag}/param/{k}"] = p.detach().numpy().astype(np.float32)
            fx[f"{tag}/grad_fp64/{k}"] = (p.grad.numpy() if p.grad is not None else np.zeros(p.shape))

    rng = np.random.default_rng(0)
    # compare_backends.py:145-160 shape: n=2, net (2,1,2,1), b_in 8, t_in 1, batch 6, TF
    torch.manual_seed(42)
    m = ref_models.QuanONetPT(num_qubits=2, branch_input_size=8, trunk_input_size=1, net_size=(2, 1, 2, 1),
                              scale_coeff=0.1, if_trainable_freq=True, ham_bound=(-5.0, 5.0))
    record("quanonet_q2_tf", m, [rng.random((6, 8)), rng.random((6, 1))], rng.random((6, 1)))
    # compare_backends.py:224-236 shape: HEAQNN n=2 net (2,1,0,0) in 6 batch 6
    torch.manual_seed(42)
    m = ref_models.HEAQNNPT(num_qubits=2, input_size=6, net_size=(2, 1, 0, 0), scale_coeff=0.1,
                            if_trainable_freq=True, ham_bound=(-5.0, 5.0))
    record("heaqnn_q2_tf", m, [rng.random((6, 6))], rng.random((6, 1)))
    # fixed-scale mode, n=3, uneven tiling (in > out for the trunk), non-default bound
    torch.manual_seed(7)
    m = ref_models.QuanONetPT(num_qubits=3, branch_input_size=7, trunk_input_size=4, net_size=(3, 2, 1, 1),
                              scale_coeff=0.7, if_trainable_freq=False, ham_bound=(-2.0, 3.0))
    record("quanonet_q3_fixed", m, [rng.standard_normal((9, 7)), rng.random((9, 4))], rng.standard_normal((9, 1)))
    # explicit diagonal Hamiltonian (scripts/reproduce_hamiltonian.sh:103), TorchQuantum index order
    torch.manual_seed(3)
    m = ref_models.QuanONetPT(num_qubits=2, branch_input_size=5, trunk_input_size=1, net_size=(3, 2, 3, 2),
                              scale_coeff=0.5, if_trainable_freq=True, ham_diag=[-5.0, -2.5, 2.5, 5.0])
    record("quanonet_q2_diag", m, [rng.standard_normal((8, 5)), rng.random((8, 1))], rng.standard_normal((8, 1)))
    # pretrained Antideriv Q2 (compare_backends.py:288-376 shape), batch 16 RNG inputs
    m = ref_models.QuanONetPT(num_qubits=2, branch_input_size=10, trunk_input_size=1, net_size=(5, 1, 5, 1),
                              scale_coeff=0.001, if_trainable_freq=True, ham_bound=(-5.0, 5.0))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params_of(weights, "Antideriv").items()})
    record("antideriv_pretrained", m, [rng.random((16, 10)), rng.random((16, 1))], rng.random((16, 1)))
    # pretrained Advection Q5 Net40-2-20-2, batch 32 (primary config), with trainable-freq grads
    m = ref_models.QuanONetPT(num_qubits=5, branch_input_size=100, trunk_input_size=2, net_size=(40, 2, 20, 2),
                              scale_coeff=0.1, if_trainable_freq=True, ham_bound=(-5.0, 5.0))
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params_of(weights, "Advection").items()})
    record("advection_pretrained", m, [rng.standard_normal((32, 100)), rng.random((32, 2))],
           rng.standard_normal((32, 1)))
    np.savez_compressed(os.path.join(HERE, "wrapper_cases.npz"), **fx)


def raw_circuit_goldens():
    """Bare circuit cases straight on (x, W): every n in 1..8, Pauli X/Y/Z sums, diagonals in both
    index orders, ragged x (fewer columns than encoding gates, quantum_circuits_tq.py:83)."""
    rng = np.random.default_rng(1234)
    fx = {}
    meta = {}

    def case(tag, n, blocks, ham, B, n_cols=None, ham_meta=None):
        E = orc.num_encode_cols(blocks) if n_cols is None else n_cols
        S = orc.num_sublayers(blocks)
        x = rng.uniform(-np.pi, np.pi, (B, E)).astype(np.float32)
        w = rng.uniform(-np.pi, np.pi, (S, 3, n)).astype(np.float32)
        g = rng.standard_normal(B).astype(np.float32)
        e, gx, gw = orc.hea_forward_backward(x, w, n, blocks, ham, grad_out=g)
        fx[f"{tag}/x"], fx[f"{tag}/w"], fx[f"{tag}/g"] = x, w, g
        fx[f"{tag}/e"], fx[f"{tag}/gx"], fx[f"{tag}/gw"] = e, gx, gw
        meta[tag] = {"n": n, "blocks": [list(b) for b in blocks], "ham": ham_meta}

    for n in range(1, 9):
        blocks = orc.make_block_configs(n, 2, 2, 3, 1)
        case(f"z_n{n}", n, blocks, orc.ham_from_bound(n, -5, 5), 5,
             ham_meta={"kind": "pauli", "pauli": "Z", "bound": [-5, 5]})
    for p in "XY":
        for n in (2, 5, 6):
            blocks = orc.make_block_configs(n, 2, 2, 2, 2)
            case(f"{p.lower()}_n{n}", n, blocks, orc.ham_from_bound(n, -3, 7, pauli=p), 4,
                 ham_meta={"kind": "pauli", "pauli": p, "bound": [-3, 7]})
    for n in (2, 4, 5):
        d = rng.uniform(-5, 5, 1 << n)
        blocks = orc.heaqnn_block_configs(n, 4, 2)
        case(f"diag_msb0_n{n}", n, blocks, orc.ham_from_diag(d, n, "msb0"), 4,
             ham_meta={"kind": "diag", "order": "msb0", "diag": d.tolist()})
        case(f"diag_lsb0_n{n}", n, blocks, orc.ham_from_diag(d, n, "lsb0"), 4,
             ham_meta={"kind": "diag", "order": "lsb0", "diag": d.tolist()})
    # ragged: 3 blocks of n=3 → 9 encoding gates, x has only 7 columns
    case("ragged_n3", 3, orc.heaqnn_block_configs(3, 3, 1), orc.ham_from_bound(3, -5, 5), 4, n_cols=7,
         ham_meta={"kind": "pauli", "pauli": "Z", "bound": [-5, 5]})
    # primary config, synthetic
    case("c2_q5_net40", 5, orc.make_block_configs(5, 20, 2, 40, 2), orc.ham_from_bound(5, -5, 5), 16,
         ham_meta={"kind": "pauli", "pauli": "Z", "bound": [-5, 5]})
    np.savez_compressed(os.path.join(HERE, "circuit_cases.npz"), **fx)
    with open(os.path.join(HERE, "circuit_cases.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    w = weights_fixture()
    demo = notebook_demo(w)
    anti = antideriv_closed_forms(w)
    with open(os.path.join(HERE, "published.json"), "w") as f:
        json.dump({"notebook_published": PUBLISHED, "notebook_oracle_fp64": demo,
                   "antideriv_closed_form_oracle_fp64": anti,
                   "survey_probe_expected": {"cos": {"rel_l2": 0.026915, "mse": 3.6332e-05},
                                             "lin": {"rel_l2": 0.088683, "mse": 3.9920e-04}}}, f, indent=1)
    wrapper_goldens(w)
    raw_circuit_goldens()
    print("golden fixtures written to", HERE)
