"""CPU emulation (exact complex arithmetic) of the tensor-core tier's ADJOINT sweep, against the fp64 oracle:
per-sublayer un-apply matrices, Pauli-string moments on post-ring cuts (computational / Hadamard basis),
encoding-angle gradients as Z-type moments in the Hadamard basis, finalize formulas."""
import os, sys, re
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import hea_oracle as orc
import tc_emulate as emu

n, N = 5, 32


def load_strings():
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "quanonet_b200", "csrc", "tc_strings.cuh")).read()
    out = {}
    for name in ("kTcStrComp", "kTcStrHad"):
        body = src[src.index("TcString " + name):]
        body = body[:body.index("};")]
        out[name] = [tuple(int(v) for v in m) for m in re.findall(r"\{(\d+), (\d+), (\d+)\}", body)]
        assert len(out[name]) == 15
    return out


STR = load_strings()


def fwht(v):
    v = v.copy()
    for q in range(n):
        for z in range(N):
            if z & (1 << q):
                continue
            z1 = z | (1 << q)
            a, b = v[z], v[z1]
            v[z], v[z1] = (a + b) / np.sqrt(2), (a - b) / np.sqrt(2)
    return v


def rev_matrix(w, s, last, first_in_block, input_had):
    """Composite C_s = [H if first-in-block] (R_s^+ Ring^+) ... (R_last^+ Ring^+) [H if the block's output cut is held
    in the Hadamard basis]: takes the block's OUTPUT cut to the cut after sublayer s-1."""
    G = np.zeros((N, N), complex)
    for j in range(N):
        v = np.zeros(N, complex); v[j] = 1
        if input_had:
            v = fwht(v)
        for ss in range(last, s - 1, -1):
            for i in reversed(range(n)):
                c = (i + 1) % n
                for z in range(N):
                    if ((z >> c) & 1) and not ((z >> i) & 1):
                        z1 = z | (1 << i)
                        v[z], v[z1] = v[z1], v[z]
            for q in range(n):
                al, be = emu.su2(w[ss, 0, q], w[ss, 1, q], w[ss, 2, q])
                for z in range(N):
                    if z & (1 << q):
                        continue
                    z1 = z | (1 << q)
                    x0, x1 = v[z], v[z1]
                    v[z] = np.conj(al) * x0 + np.conj(be) * x1
                    v[z1] = -be * x0 + al * x1
        if first_in_block:
            v = fwht(v)
        G[:, j] = v
    return G


def moments(psi, lam, table):
    m = np.zeros(15)
    for t, (mx, mz, k) in enumerate(table):
        acc = 0.0
        for zp in range(N):
            sgn = (-1) ** bin((zp ^ mx) & mz).count("1")
            acc += np.imag(np.conj(lam[zp]) * (1j ** k) * sgn * psi[zp ^ mx])
        m[t] = acc
    return m


def tc_backward(x, w, depths, hdiag, gout):
    B, K, S = x.shape[0], len(depths), sum(depths)
    Ms, s0 = [], 0
    for k, d in enumerate(depths):
        Ms.append(emu.block_matrix(w, s0, d, k == K - 1)); s0 += d
    first = np.zeros(S, bool); blk = np.zeros(S, int); last = np.zeros(S, bool)
    s = 0
    for k, d in enumerate(depths):
        first[s] = True; blk[s:s + d] = k; last[s + d - 1] = True; s += d
    lastof = np.zeros(S, int)
    s = 0
    for k, d in enumerate(depths):
        lastof[s:s + d] = s + d - 1; s += d
    Gs = [rev_matrix(w, s, lastof[s], True, blk[s] < K - 1) for s in range(S)]
    out = np.zeros(B); gx = np.zeros((B, n * K)); mom = np.zeros((S, 15))
    for b in range(B):
        amp = np.full(N, 1 / np.sqrt(N), complex)
        phs = []
        for k in range(K):
            th = x[b, k * n:(k + 1) * n]
            ph = np.array([np.prod([np.exp((-1j if not (z >> q) & 1 else 1j) * th[q] / 2) for q in range(n)]) for z in range(N)])
            phs.append(ph)
            amp = Ms[k] @ (amp * ph)
        out[b] = np.sum(hdiag * np.abs(amp) ** 2)
        psi, lam = amp, gout[b] * hdiag * amp
        for s in reversed(range(S)):
            had = not (s == S - 1)
            if last[s]:
                opsi, olam = psi, lam          # the block's output cut: operand of every GEMM of the block
            mom[s] += moments(psi, lam, STR["kTcStrHad"] if had else STR["kTcStrComp"])
            psi, lam = Gs[s] @ opsi, Gs[s] @ olam
            if first[s]:
                k = blk[s]
                wz = np.imag(np.conj(lam) * psi)
                for q in range(n):
                    gx[b, k * n + q] = sum((1 - 2 * ((z >> q) & 1)) * wz[z] for z in range(N))
                psi, lam = np.conj(phs[k]) * psi, np.conj(phs[k]) * lam
    gw = np.zeros_like(w)
    for s in range(S):
        for q in range(n):
            mX, mY, mZ = mom[s, 3 * q:3 * q + 3]
            bb, cc = w[s, 1, q], w[s, 2, q]
            gw[s, 0, q] = np.cos(bb) * mY - np.sin(bb) * (np.cos(cc) * mX - np.sin(cc) * mZ)
            gw[s, 1, q] = np.cos(cc) * mZ + np.sin(cc) * mX
            gw[s, 2, q] = mY
    return out, gx, gw


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    for depths in ([1], [2], [1, 2], [2, 1, 3], [2] * 5):
        K, S = len(depths), sum(depths)
        B = 4
        x = rng.uniform(-np.pi, np.pi, (B, n * K)); w = rng.uniform(-np.pi, np.pi, (S, 3, n)); g = rng.normal(size=B)
        hd = np.array([n - 2 * bin(z).count("1") for z in range(N)], float)
        o_ref, gx_ref, gw_ref = orc.hea_forward_backward(x, w, n, [(n, d) for d in depths], orc.ham_from_bound(n), g)
        o, gx, gw = tc_backward(x, w, depths, hd, g)
        rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
        print(f"depths {depths}: out {rel(o, o_ref):.1e}  grad_x {rel(gx, gx_ref):.1e}  grad_w {rel(gw, gw_ref):.1e}")
