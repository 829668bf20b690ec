#!/usr/bin/env python
"""Cross-backend consistency check in the style of the reference's compare_backends.py (helpers :40-47,
tolerances :26-31, cases :140-212, :219-281, :288-376): same weights -> compare forward and the gradients of
((model(x) - tgt)**2).mean() between

    B200 (this package, CUDA)  vs  TorchQuantum-faithful complex64 restatement (CPU)  vs  fp64 oracle (CPU).

The oracle side makes this a checker, hence it lives under tests/.  Run:  python tests/harness/compare_backends_b200.py
Exit code 1 iff any comparison FAILs (as the reference's script does).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from oracle import hea_oracle as orc
from oracle.tq_faithful import tq_forward
from quanonet_b200.core.models_pt import HEAQNNPT, QuanONetPT, _tile_to

ATOL_PT, ATOL_MSPT, ATOL_GRAD_PT, ATOL_GRAD_MS = 1e-4, 1e-4, 1e-4, 5e-4      # compare_backends.py:26-31
results = {}
RNG = np.random.default_rng(0)


def _ok(tag, a, b, atol):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    diff = float(np.abs(a - b).max())
    passed = diff <= atol
    print(f"  [{'PASS' if passed else 'FAIL'}]  {tag:<58s}  max_diff={diff:.2e}")
    results[tag] = passed
    return passed


def _enc(layer, u):
    if hasattr(layer, "weights"):
        return _tile_to(u, layer.out_features) * layer.weights + layer.bias
    return _tile_to(u * layer.scale, layer.out_features)


def _cpu_eval(model_cpu, inputs, tgt, fp64):
    """Forward + gradients on the CPU through the TQ-faithful complex64 path or the fp64 oracle."""
    m = model_cpu
    q = m.quantum_layer
    if hasattr(m, "branch_freq"):
        x = torch.cat([_enc(m.trunk_freq, inputs[1]), _enc(m.branch_freq, inputs[0])], dim=1)
    else:
        x = _enc(m.freq, inputs[0])
    if fp64:
        class Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, xx, ww):
                ctx.save_for_backward(xx, ww)
                e = orc.hea_forward(xx.detach().numpy(), ww.detach().numpy(), q.n_wires, q.block_configs,
                                    orc.Ham("pauli", "Z", q.ham_offset, q.ham_coeff))
                return torch.tensor(e, dtype=xx.dtype).reshape(-1, 1)

            @staticmethod
            def backward(ctx, g):
                xx, ww = ctx.saved_tensors
                _, gx, gw = orc.hea_forward_backward(xx.detach().numpy(), ww.detach().numpy(), q.n_wires, q.block_configs,
                                                     orc.Ham("pauli", "Z", q.ham_offset, q.ham_coeff),
                                                     grad_out=g.numpy().reshape(-1))
                return torch.tensor(gx, dtype=xx.dtype), torch.tensor(gw, dtype=ww.dtype)
        out = Fn.apply(x, q.ansatz_weights)
    else:
        out = tq_forward(x, q.ansatz_weights, q.n_wires, q.block_configs, q.ham_offset, q.ham_coeff)
    if hasattr(m, "bias"):
        out = out + m.bias
    m.zero_grad()
    ((out - tgt) ** 2).mean().backward()
    return out.detach().numpy(), {k: p.grad.numpy().copy() for k, p in m.named_parameters()}


def compare(tag, make, inputs, tgt, atol_f, atol_g):
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    m_gpu = make().to(dev)
    sd = {k: v.cpu() for k, v in m_gpu.state_dict().items()}
    gin = tuple(torch.tensor(a, dtype=torch.float32, device=dev) for a in inputs)
    gt = torch.tensor(tgt, dtype=torch.float32, device=dev)
    out = m_gpu(*gin)
    m_gpu.zero_grad()
    ((out - gt) ** 2).mean().backward()
    g_gpu = {k: p.grad.cpu().numpy() for k, p in m_gpu.named_parameters()}
    for name, fp64 in (("TQ-faithful c64", False), ("fp64 oracle", True)):
        mc = make()
        mc.load_state_dict(sd)
        if fp64:
            mc = mc.double()
        dt = torch.float64 if fp64 else torch.float32
        o, g = _cpu_eval(mc, tuple(torch.tensor(a, dtype=dt) for a in inputs), torch.tensor(tgt, dtype=dt), fp64)
        _ok(f"{tag}  B200 == {name}", out.detach().cpu().numpy(), o, atol_f)
        for k in g:
            _ok(f"{tag}  B200 == {name} (grad {k})", g_gpu[k], g[k], atol_g)


def main():
    print("\n--- QuanONet: B200 vs TorchQuantum-faithful vs fp64 oracle (compare_backends.py:140-212 shape) ---")
    compare("QuanONet n=2", lambda: QuanONetPT(2, 8, 1, (2, 1, 2, 1), scale_coeff=0.1, if_trainable_freq=True),
            (RNG.random((6, 8)), RNG.random((6, 1))), RNG.random((6, 1)), ATOL_PT, ATOL_GRAD_PT)
    print("\n--- HEAQNN (compare_backends.py:219-281 shape) ---")
    compare("HEAQNN n=2", lambda: HEAQNNPT(2, 6, (2, 1, 0, 0), scale_coeff=0.1, if_trainable_freq=True),
            (RNG.random((6, 6)),), RNG.random((6, 1)), ATOL_PT, ATOL_GRAD_PT)
    print("\n--- QuanONet pretrained Antideriv Q2 / Advection Q5 (compare_backends.py:288-376 shape) ---")
    z = np.load(os.path.join(ROOT, "tests", "golden", "pretrained.npz"))

    def pre(name, cfg):
        def make():
            m = QuanONetPT(**cfg)
            m.load_state_dict({k.split("/", 1)[1]: torch.tensor(z[k]) for k in z.files if k.startswith(name + "/")})
            return m
        return make
    q2 = dict(num_qubits=2, branch_input_size=10, trunk_input_size=1, net_size=(5, 1, 5, 1), scale_coeff=0.001,
              if_trainable_freq=True)
    compare("Antideriv Q2", pre("Antideriv", q2), (RNG.random((16, 10)), RNG.random((16, 1))), RNG.random((16, 1)),
            ATOL_MSPT, ATOL_GRAD_MS)
    q5 = dict(num_qubits=5, branch_input_size=100, trunk_input_size=2, net_size=(40, 2, 20, 2), scale_coeff=0.1,
              if_trainable_freq=True)
    compare("Advection Q5", pre("Advection", q5), (RNG.standard_normal((16, 100)), RNG.random((16, 2))),
            RNG.standard_normal((16, 1)), ATOL_MSPT, ATOL_GRAD_MS)
    n_fail = sum(1 for v in results.values() if v is False)
    print(f"\n{len(results) - n_fail} PASS, {n_fail} FAIL")
    sys.exit(1 if n_fail else 0)


if __name__ == "__main__":
    main()
