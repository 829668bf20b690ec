"""CPU tests: the oracle against the golden vectors / the reference's published numbers."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_circuit_cases, oracle_ham, rel_l2
from oracle import hea_oracle as orc
from oracle.tq_faithful import tq_forward_backward


def _params(name):
    z = np.load(os.path.join(GOLDEN, "pretrained.npz"))
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}


def test_published_notebook_numbers_from_stored_grids():
    """The twelve MSE/MAE figures printed in visualization.ipynb (figure titles) follow from the
    stored truth grids (reference PDE solvers) and the oracle's fp64 predictions."""
    z = np.load(os.path.join(GOLDEN, "notebook_demo.npz"))
    pub = json.load(open(os.path.join(GOLDEN, "published.json")))["notebook_published"]
    assert len(pub) == 6
    for key, exp in pub.items():
        diff = z[f"{key}/truth"] - z[f"{key}/pred_fp64"]
        assert f"{np.mean(diff ** 2):.1e}" == exp["mse"], key
        assert f"{np.mean(np.abs(diff)):.1e}" == exp["mae"], key


@pytest.mark.parametrize("op", ["Advection", "RDiffusion", "Darcy"])
def test_oracle_recomputes_notebook_predictions(op):
    """Re-run the oracle on a strided subset of each demo grid; must equal the stored predictions."""
    z = np.load(os.path.join(GOLDEN, "notebook_demo.npz"))
    P = 25 if op == "Darcy" else 100
    xs = np.linspace(0, 1, P).astype(np.float32)
    X, T = np.meshgrid(xs, xs)
    trunk = np.hstack((X.flatten()[:, None], T.flatten()[:, None])).astype(np.float32)
    sel = np.arange(0, P * P, 37 if P == 100 else 5)
    key = f"{op}/sin2"
    branch = np.tile(z[f"{key}/u0"], (len(sel), 1))
    pred = orc.quanonet_forward(branch, trunk[sel], _params(op), 5, (40, 2, 20, 2), orc.ham_from_bound(5))
    assert np.abs(pred - z[f"{key}/pred_fp64"].reshape(-1)[sel]).max() < 1e-12


def test_survey_point_values():
    """Point values the survey's independent fp64 restatement recorded.  That
    probe did not round its inputs to float32 the way the notebook does, hence 1e-6 and not 1e-12."""
    z = np.load(os.path.join(GOLDEN, "notebook_demo.npz"))
    p = z["Advection/sin2/pred_fp64"]
    assert abs(p[0, 0] - (-0.0372213004)) < 1e-6
    assert abs(p[50, 50] - 0.0647315409) < 1e-6
    assert abs(p.reshape(-1)[1234] - 0.9708587997) < 1e-6
    assert abs(p.reshape(-1)[9999] - 0.1118678526) < 1e-6
    assert abs(z["Darcy/sin2/pred_fp64"][12, 12] - (-0.7047581092)) < 1e-6
    assert abs(z["RDiffusion/sin4/pred_fp64"][50, 50] - 0.0210927628) < 1e-6


def test_antideriv_closed_forms():
    """ibm_inference.py:177-189 closed forms through the shipped Antideriv Q2 weights: Rel-L2 0.026915
    for u0=cos(pi x), 0.088683 for u0=x (a reversed CNOT ring gives 6.33 / 4.16)."""
    z = np.load(os.path.join(GOLDEN, "antideriv_closed_form.npz"))
    exp = json.load(open(os.path.join(GOLDEN, "published.json")))["survey_probe_expected"]
    for tag in ("cos", "lin"):
        pred = orc.quanonet_forward(z[f"{tag}/branch"], z[f"{tag}/trunk"], _params("Antideriv"), 2, (5, 1, 5, 1),
                                    orc.ham_from_bound(2))
        assert abs(rel_l2(pred, z[f"{tag}/truth"]) - exp[tag]["rel_l2"]) < 1e-6
        assert abs(np.mean((pred - z[f"{tag}/truth"]) ** 2) - exp[tag]["mse"]) < 1e-6


def test_oracle_matches_circuit_goldens():
    z, meta = load_circuit_cases()
    for tag, m in meta.items():
        if m["n"] > 6:
            continue
        blocks = [tuple(b) for b in m["blocks"]]
        e, gx, gw = orc.hea_forward_backward(z[f"{tag}/x"], z[f"{tag}/w"], m["n"], blocks, oracle_ham(m["ham"], m["n"]),
                                             grad_out=z[f"{tag}/g"])
        assert np.abs(e - z[f"{tag}/e"]).max() < 1e-12 and np.abs(gx - z[f"{tag}/gx"]).max() < 1e-12
        assert np.abs(gw - z[f"{tag}/gw"]).max() < 1e-11


def test_adjoint_gradient_vs_finite_differences_and_state_rewind():
    rng = np.random.default_rng(0)
    for n, pauli in ((3, "Z"), (2, "X"), (3, "Y")):
        blocks = orc.make_block_configs(n, 2, 2, 1, 1)
        ham = orc.ham_from_bound(n, -2, 3, pauli=pauli)
        x = rng.uniform(-3, 3, (3, orc.num_encode_cols(blocks)))
        w = rng.uniform(-3, 3, (orc.num_sublayers(blocks), 3, n))
        g = rng.standard_normal(3)
        e, gx, gw, psi = orc.hea_forward_backward(x, w, n, blocks, ham, g, return_state=True)
        assert abs(psi[:, 0] - 1).max() < 1e-12 and abs(psi[:, 1:]).max() < 1e-12
        f = lambda xx, ww: float((g * orc.hea_forward(xx, ww, n, blocks, ham)).sum())
        eps = 1e-6
        for idx in list(np.ndindex(*w.shape))[::5]:
            wp, wm = w.copy(), w.copy()
            wp[idx] += eps
            wm[idx] -= eps
            assert abs((f(x, wp) - f(x, wm)) / (2 * eps) - gw[idx]) < 1e-7
        for idx in list(np.ndindex(*x.shape))[::4]:
            xp, xm = x.copy(), x.copy()
            xp[idx] += eps
            xm[idx] -= eps
            assert abs((f(xp, w) - f(xm, w)) / (2 * eps) - gx[idx]) < 1e-7


def test_norm_and_spectrum_bounds():
    rng = np.random.default_rng(1)
    n = 5
    blocks = orc.make_block_configs(n, 3, 2, 4, 2)
    x = rng.uniform(-np.pi, np.pi, (8, n * 7))
    w = rng.uniform(-np.pi, np.pi, (14, 3, n))
    psi = orc.hea_state(x, w, n, blocks)
    assert np.abs((np.abs(psi) ** 2).sum(1) - 1).max() < 1e-13
    e = orc.hea_forward(x, w, n, blocks, orc.ham_from_bound(n, -5, 5))
    assert np.all(np.abs(e) <= 5 + 1e-12)


def test_tq_faithful_complex64_agrees_with_oracle():
    """The TorchQuantum-faithful complex64 restatement (op order, dtypes and autograd of
    core/quantum_circuits_tq.py:65-127) agrees with the fp64 oracle, incl. TQ's MSB-first diagonal."""
    z, meta = load_circuit_cases()
    for tag in ("z_n1", "z_n3", "z_n5", "diag_msb0_n4", "ragged_n3"):
        m = meta[tag]
        n = m["n"]
        blocks = [tuple(b) for b in m["blocks"]]
        h = m["ham"]
        kw = dict(ham_diag=h["diag"]) if h["kind"] == "diag" else dict(
            zip(("ham_offset", "ham_coeff"), orc.ham_params(n, *h["bound"])))
        out, gx, gw = tq_forward_backward(torch.tensor(z[f"{tag}/x"]), torch.tensor(z[f"{tag}/w"]), n, blocks,
                                          torch.tensor(z[f"{tag}/g"]), **kw)
        assert rel_l2(out.numpy()[:, 0], z[f"{tag}/e"]) < 1e-5
        assert rel_l2(gx.numpy(), z[f"{tag}/gx"]) < 1e-5
        assert rel_l2(gw.numpy(), z[f"{tag}/gw"]) < 1e-5


def test_diag_index_orders_differ_unless_symmetric():
    n = 2
    d = np.array([-5.0, -2.5, 2.5, 5.0])   # scripts/reproduce_hamiltonian.sh:103 — NOT bit-reversal symmetric
    assert not np.allclose(orc.diag_msb0_to_lsb0(d, n), d)
    sym = np.array([-5.0, 0.0, 0.0, 5.0])
    assert np.allclose(orc.diag_msb0_to_lsb0(sym, n), sym)
    # sum-Z as an explicit diagonal equals the Pauli form in either order
    rng = np.random.default_rng(2)
    blocks = orc.heaqnn_block_configs(3, 2, 2)
    x, w = rng.uniform(-3, 3, (4, 6)), rng.uniform(-3, 3, (4, 3, 3))
    zs = orc.z_sum_diag(3)
    a = orc.hea_forward(x, w, 3, blocks, orc.Ham("pauli", "Z", 0.5, 1.5))
    for order in ("lsb0", "msb0"):
        b = orc.hea_forward(x, w, 3, blocks, orc.ham_from_diag(0.5 + 1.5 * zs, 3, order))
        assert np.abs(a - b).max() < 1e-13


def test_oracle_properties_on_random_circuits():
    """Randomised (hypothesis) properties of the oracle itself: the expectation value lies in the spectrum of H,
    the adjoint gradient is linear in the upstream gradient and matches a central finite difference on a
    randomly chosen shared angle and a randomly chosen encoding angle."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None)
    @given(n=st.integers(1, 4), depths=st.lists(st.integers(1, 3), min_size=1, max_size=3),
           kind=st.sampled_from(["X", "Y", "Z", "diag"]), seed=st.integers(0, 10_000))
    def check(n, depths, kind, seed):
        rng = np.random.default_rng(seed)
        blocks = [(n, d) for d in depths]
        K, S, B = len(depths), sum(depths), 3
        x = rng.uniform(-np.pi, np.pi, (B, n * K))
        w = rng.uniform(-np.pi, np.pi, (S, 3, n))
        if kind == "diag":
            d = rng.uniform(-3, 3, 1 << n)
            ham, lo, hi = orc.ham_from_diag(d, n), d.min(), d.max()
        else:
            off, co = 0.4, 0.9
            ham, lo, hi = orc.Ham("pauli", kind, off, co), off - n * abs(co), off + n * abs(co)
        g = rng.standard_normal(B)
        e, gx, gw = orc.hea_forward_backward(x, w, n, blocks, ham, g)
        assert np.all(e >= lo - 1e-9) and np.all(e <= hi + 1e-9)
        _, gx2, gw2 = orc.hea_forward_backward(x, w, n, blocks, ham, 2.5 * g)
        assert np.allclose(gx2, 2.5 * gx, atol=1e-12) and np.allclose(gw2, 2.5 * gw, atol=1e-12)
        h = 1e-6
        s_, g_, q_ = rng.integers(S), rng.integers(3), rng.integers(n)
        wp, wm = w.copy(), w.copy()
        wp[s_, g_, q_] += h; wm[s_, g_, q_] -= h
        fd = (g * (orc.hea_forward(x, wp, n, blocks, ham) - orc.hea_forward(x, wm, n, blocks, ham))).sum() / (2 * h)
        assert abs(fd - gw[s_, g_, q_]) < 1e-6 * max(1.0, abs(fd))
        b_, c_ = rng.integers(B), rng.integers(n * K)
        xp, xm = x.copy(), x.copy()
        xp[b_, c_] += h; xm[b_, c_] -= h
        fdx = g[b_] * (orc.hea_forward(xp, w, n, blocks, ham)[b_] - orc.hea_forward(xm, w, n, blocks, ham)[b_]) / (2 * h)
        assert abs(fdx - gx[b_, c_]) < 1e-6 * max(1.0, abs(fdx))

    check()
