"""Hamiltonian sweep in fp64 (BASELINE config 5; shapes of the reference's scripts/reproduce_hamiltonian.sh:41-104):
Pauli X/Y/Z sums at n = 5 net (20,2,10,2), ham_bound +-1 ... +-10, explicit diagonals at n = 2 net (50,2,50,2).
For every case: fp64 CUDA forward + adjoint gradients vs the fp64 oracle on a 48-sample slice (norm-relative error,
bar 1e-12) and throughput at B samples.

    python tests/harness/sweep_hamiltonian.py [--batch 262144] [--out profiles/hamiltonian.jsonl]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import hea_oracle as orc
from quanonet_b200 import _lib
from quanonet_b200.ops import hea_expval_backward

def rel(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

def run_case(tag, n, net, ham, op_kw, B, dev):
    b_d, b_l, t_d, t_l = net
    blocks = orc.make_block_configs(n, t_d, t_l, b_d, b_l); depths = [d for _, d in blocks]
    g = torch.Generator().manual_seed(hash(tag) % (1 << 31))
    x = ((torch.rand(B, n * len(blocks), generator=g, dtype=torch.float64) * 2 - 1) * np.pi).to(dev)
    w = ((torch.rand(sum(depths), 3, n, generator=g, dtype=torch.float64) * 2 - 1) * np.pi).to(dev)
    go = torch.randn(B, generator=g, dtype=torch.float64).to(dev)
    hd = None if op_kw["ham_diag"] is None else torch.tensor(op_kw["ham_diag"], dtype=torch.float64, device=dev)
    args = (n, depths, hd, op_kw["diag_order"], op_kw["ham_offset"], op_kw["ham_coeff"], op_kw["ham_kind"], True)
    o, gx, gw = hea_expval_backward(go, x, w, *args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3): hea_expval_backward(go, x, w, *args)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    sl = slice(0, 48)
    e, egx, _ = orc.hea_forward_backward(x[sl].cpu().numpy(), w.cpu().numpy(), n, blocks, ham, grad_out=go[sl].cpu().numpy())
    # shared-parameter gradient: compare on the slice alone (its own launch)
    _, _, gws = hea_expval_backward(go[sl].contiguous(), x[sl].contiguous(), w, *args)
    _, _, egw = orc.hea_forward_backward(x[sl].cpu().numpy(), w.cpu().numpy(), n, blocks, ham, grad_out=go[sl].cpu().numpy())
    return {"case": tag, "n": n, "net": list(net), "B": B, "ms": ms, "samples_per_s": B / ms * 1e3,
            "err_out": rel(o[sl, 0].cpu().numpy(), e), "err_grad_x": rel(gx[sl].cpu().numpy(), egx),
            "err_grad_w": rel(gws.cpu().numpy(), egw)}

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=262144)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    rows = []
    for p_, kind in (("Z", _lib.QON_HAM_DIAG), ("X", _lib.QON_HAM_PAULI_X), ("Y", _lib.QON_HAM_PAULI_Y)):
        for bound in (1, 2, 5, 10):
            off, co = orc.ham_params(5, -bound, bound)
            rows.append(run_case(f"pauli{p_}_bound{bound}", 5, (20, 2, 10, 2), orc.ham_from_bound(5, -bound, bound, pauli=p_),
                                 dict(ham_diag=None, diag_order=0, ham_offset=off, ham_coeff=co, ham_kind=kind), a.batch, dev))
            print(json.dumps(rows[-1]), flush=True)
    for d in ([-5, 5, 5, 5], [-5, -5, -5, 5], [-5, 0, 0, 5], [-5, -2.5, 2.5, 5]):
        for order, code in (("msb0", _lib.QON_DIAG_MSB0), ("lsb0", _lib.QON_DIAG_LSB0)):
            rows.append(run_case(f"diag{d}_{order}", 2, (50, 2, 50, 2), orc.ham_from_diag(d, 2, order),
                                 dict(ham_diag=d, diag_order=code, ham_offset=0.0, ham_coeff=0.0, ham_kind=_lib.QON_HAM_DIAG),
                                 a.batch, dev))
            print(json.dumps(rows[-1]), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            for r in rows: f.write(json.dumps(r) + "\n")
    worst = max(max(r["err_out"], r["err_grad_x"], r["err_grad_w"]) for r in rows)
    print(f"worst norm-relative error over {len(rows)} cases: {worst:.2e}")
    sys.exit(0 if worst < 1e-12 else 1)

if __name__ == "__main__":
    main()
