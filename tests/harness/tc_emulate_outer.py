"""CPU emulation of the tensor-core tier's GEMM-form weight gradients (csrc/hea_tc3.cuh), against the fp64 oracle.

The sublayers of a block are sample-independent, so every Pauli moment the finalize kernel needs inside block k follows
from ONE batch-summed outer product per block, taken at the block's output cut:
    Y_k = sum_b g_b |psi_b><lam_b|        (a (64 x B) . (B x 64) real GEMM on the tensor cores, f16 hi/lo x 3 products)
    cut after the rotations of sublayer s:  Y <- T Y T^+  with the sample-independent T of the sublayers behind it,
    moment of P_q there:  Im <lam|P_q|psi> summed over the batch  =  Im tr(P_q Y).
Emulated here: the real 64 x 64 form of the outer product and how Y is read from it, the per-step power-of-two scale
E >= max|g| that keeps g_b lam_b inside the f16 range, fp32 tile accumulators summed in fp64, and the conjugation chain
of the moment kernel (same order of operations as tc_moment_kernel)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import hea_oracle as orc
import tc_emulate as emu
import tc_emulate_bwd as emub

n, N = 5, 32
SA = 32768.0


def split16(v):
    h = v.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    h = h.view(np.float32)
    l = (v.astype(np.float32) - h).astype(np.float16).astype(np.float32)
    return h.astype(np.float16).astype(np.float32), l


def real_rows(v):
    """(B, 32) complex -> (B, 64) real rows [re0, im0, re1, im1, ...] (the operand rows of the kernels)"""
    out = np.zeros((v.shape[0], 64), np.float32)
    out[:, 0::2], out[:, 1::2] = v.real, v.imag
    return out


def outer_fixed(psi, lam, gt, exact):
    """D[m][n'] = sum_b lam~[b][m] psi~[b][n'] per 128-sample tile (fp32 accumulator), tiles summed in fp64."""
    B = psi.shape[0]
    acc = np.zeros((64, 64))
    for t0 in range(0, B, 128):
        P = real_rows(psi[t0:t0 + 128] * SA)
        L = real_rows(lam[t0:t0 + 128] * gt[t0:t0 + 128, None] * SA)
        if exact:
            D = L.astype(np.float64).T @ P.astype(np.float64)
        else:
            Ph, Pl = split16(P); Lh, Ll = split16(L)
            D = (Lh.T @ Ph + Lh.T @ Pl + Ll.T @ Ph).astype(np.float32)
        acc += D.astype(np.float64)
    return acc


def y_from_d(acc, scale):
    """Y[j][i] = sum_b g_b psi_j conj(lam_i):  Re = D[2i][2j] + D[2i+1][2j+1],  Im = D[2i][2j+1] - D[2i+1][2j]"""
    D = acc.astype(np.float64) * scale
    Y = np.zeros((N, N), complex)
    for i in range(N):
        for j in range(N):
            Y[j, i] = (D[2 * i, 2 * j] + D[2 * i + 1, 2 * j + 1]) + 1j * (D[2 * i, 2 * j + 1] - D[2 * i + 1, 2 * j])
    return Y


def conj_1q(Y, q, U):
    """Y <- U_q Y U_q^+ for a 2x2 U on qubit q (rows transform like psi, columns like conj(lam))"""
    Y = Y.copy()
    for z in range(N):
        if z & (1 << q):
            continue
        z1 = z | (1 << q)
        r0, r1 = Y[z].copy(), Y[z1].copy()
        Y[z], Y[z1] = U[0, 0] * r0 + U[0, 1] * r1, U[1, 0] * r0 + U[1, 1] * r1
    Uc = np.conj(U)
    for z in range(N):
        if z & (1 << q):
            continue
        z1 = z | (1 << q)
        c0, c1 = Y[:, z].copy(), Y[:, z1].copy()
        Y[:, z], Y[:, z1] = Uc[0, 0] * c0 + Uc[0, 1] * c1, Uc[1, 0] * c0 + Uc[1, 1] * c1
    return Y


def unring(Y):
    """Y <- Ring^+ Y Ring: the CNOTs (control (i+1)%n -> target i) in reverse order, on rows and columns"""
    perm = np.arange(N)
    for i in reversed(range(n)):
        c = (i + 1) % n
        perm = np.array([z ^ (1 << i) if (z >> c) & 1 else z for z in perm])
    # after the sweep perm[z] = image of z under Ring^+ (a permutation): new[perm[z]] = old[z]
    out = np.zeros_like(Y)
    out[np.ix_(perm, perm)] = Y
    return out


def block_moments(Y, w, s0, d, had):
    """moments [d][15] of block sublayers s0 .. s0+d-1 from the block-output outer product Y (mirror of tc_moment_kernel)"""
    Hd = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    if had:
        for q in range(n):
            Y = conj_1q(Y, q, Hd)
    mom = np.zeros((d, 15))
    for s in range(s0 + d - 1, s0 - 1, -1):
        Y = unring(Y)
        for q in range(n):
            bq = 1 << q
            mx = sum(Y[z ^ bq, z] for z in range(N))                                   # tr(X_q Y)
            my = sum((1j if (z >> q) & 1 else -1j) * Y[z ^ bq, z] for z in range(N))   # tr(Y_q Y)
            mz = sum((1 - 2 * ((z >> q) & 1)) * Y[z, z] for z in range(N))             # tr(Z_q Y)
            mom[s - s0, 3 * q:3 * q + 3] = [mx.imag, my.imag, mz.imag]
        for q in range(n):
            al, be = emu.su2(w[s, 0, q], w[s, 1, q], w[s, 2, q])
            Ud = np.array([[np.conj(al), np.conj(be)], [-be, al]])                    # U^+
            Y = conj_1q(Y, q, Ud)
    return mom


def tc_backward_outer(x, w, depths, hdiag, gout, exact=False):
    B, K, S = x.shape[0], len(depths), sum(depths)
    Ms, s0s, s0 = [], [], 0
    for k, d in enumerate(depths):
        Ms.append(emu.block_matrix(w, s0, d, k == K - 1)); s0s.append(s0); s0 += d
    hmax = np.max(np.abs(hdiag))
    # forward
    amp = np.full((B, N), 1 / np.sqrt(N), complex)
    phs = []
    for k in range(K):
        th = x[:, k * n:(k + 1) * n]
        ph = np.ones((B, N), complex)
        for q in range(n):
            bit = (np.arange(N) >> q) & 1
            ph *= np.exp(np.where(bit[None, :] == 0, -1j, 1j) * th[:, q:q + 1] / 2)
        phs.append(ph)
        amp = (amp * ph) @ Ms[k].T
    out = np.sum(hdiag[None, :] * np.abs(amp) ** 2, axis=1)
    gmax = np.float32(np.max(np.abs(gout)))
    E = float(np.uint32((np.float32(gmax).view(np.uint32) & np.uint32(0x7F800000)) + np.uint32(0x00800000)).view(np.float32)) if gmax > 0 else 1.0
    gt = (gout / E)
    psi, lam = amp, (hdiag[None, :] / hmax) * amp
    gx = np.zeros((B, n * K)); mom = np.zeros((S, 15))
    for k in reversed(range(K)):
        acc = outer_fixed(psi, lam, gt, exact)
        Y = y_from_d(acc, E * hmax / (SA * SA))
        mom[s0s[k]:s0s[k] + depths[k]] = block_moments(Y, w, s0s[k], depths[k], k < K - 1)
        Minv = np.conj(Ms[k].T)
        psi, lam = psi @ Minv.T, lam @ Minv.T
        wz = np.imag(np.conj(lam) * psi)
        for q in range(n):
            gx[:, k * n + q] = gout * hmax * np.sum((1 - 2 * ((np.arange(N) >> q) & 1))[None, :] * wz, axis=1)
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
    for depths, B in (([1], 5), ([2], 130), ([1, 2], 7), ([2, 1, 3], 300), ([2] * 5, 200)):
        K, S = len(depths), sum(depths)
        x = rng.uniform(-np.pi, np.pi, (B, n * K)); w = rng.uniform(-np.pi, np.pi, (S, 3, n)); g = rng.normal(size=B) * 1e-3
        hd = np.array([n - 2 * bin(z).count("1") for z in range(N)], float)
        o_ref, gx_ref, gw_ref = orc.hea_forward_backward(x, w, n, [(n, d) for d in depths], orc.ham_from_bound(n), g)
        rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
        for exact in (True, False):
            o, gx, gw = tc_backward_outer(x, w, depths, hd, g, exact)
            print(f"depths {depths} B={B} exact={exact}: out {rel(o, o_ref):.1e}  grad_x {rel(gx, gx_ref):.1e}  grad_w {rel(gw, gw_ref):.1e}")
