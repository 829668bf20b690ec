"""CPU emulation of the tensor-core tier's arithmetic (csrc/hea_tc.cuh), checked against the fp64 oracle.

Validates, without a GPU: the Hadamard-basis reformulation (RX layer -> diagonal phases, H folded into the
block matrices), the real 64x64 form and its K-major image offsets, the conjugate-symmetric phase table, and
the f16 hi/lo split with three products (an estimate of the precision the kernel should reach).
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import hea_oracle as orc

n, N = 5, 32
SA, SB = 32768.0, 4096.0


def su2(a, b, c):
    ca, sa, cb, sb, cc, sc = np.cos(a / 2), np.sin(a / 2), np.cos(b / 2), np.sin(b / 2), np.cos(c / 2), np.sin(c / 2)
    al = cb * (cc * ca - sc * sa) + 1j * (-sb * (cc * ca + sc * sa))
    be = cb * (sc * ca + cc * sa) + 1j * (sb * (cc * sa - sc * ca))
    return al, be


def block_matrix(w, s0, d, last):
    M = np.zeros((N, N), complex)
    for j in range(N):
        v = np.array([(-1.0) ** bin(z & j).count("1") for z in range(N)], complex) / np.sqrt(N)
        for s in range(s0, s0 + d):
            for q in range(n):
                al, be = su2(w[s, 0, q], w[s, 1, q], w[s, 2, q])
                for z in range(N):
                    if z & (1 << q):
                        continue
                    z1 = z | (1 << q)
                    x0, x1 = v[z], v[z1]
                    v[z] = al * x0 - np.conj(be) * x1
                    v[z1] = be * x0 + np.conj(al) * x1
            for i in range(n):
                c = (i + 1) % n
                for z in range(N):
                    if ((z >> c) & 1) and not ((z >> i) & 1):
                        z1 = z | (1 << i)
                        v[z], v[z1] = v[z1], v[z]
        if not last:
            for q in range(n):
                for z in range(N):
                    if z & (1 << q):
                        continue
                    z1 = z | (1 << q)
                    x0, x1 = v[z], v[z1]
                    v[z], v[z1] = (x0 + x1) / np.sqrt(2), (x0 - x1) / np.sqrt(2)
        M[:, j] = v
    return M


def b_offset(nn, kk):
    return (nn >> 3) * 1024 + (kk >> 3) * 128 + (nn & 7) * 16 + (kk & 7) * 2


def image(M):
    """f16 hi/lo images in the K-major no-swizzle layout, then read back as Bt[nn][kk] (checks the offsets)."""
    hi = np.zeros(4096, np.float16)
    lo = np.zeros(4096, np.float16)
    seen = set()
    def put(nn, kk, v):
        vs = v * SB
        h = np.float16(vs)
        l = np.float16(vs - float(h))
        o = b_offset(nn, kk) >> 1
        assert o not in seen
        seen.add(o)
        hi[o], lo[o] = h, l
    for i in range(N):
        for j in range(N):
            re, im = M[i, j].real, M[i, j].imag
            put(2 * i, 2 * j, re); put(2 * i, 2 * j + 1, -im); put(2 * i + 1, 2 * j, im); put(2 * i + 1, 2 * j + 1, re)
    assert len(seen) == 4096
    Bh = np.zeros((64, 64)); Bl = np.zeros((64, 64))
    for nn in range(64):
        for kk in range(64):
            o = b_offset(nn, kk) >> 1
            Bh[nn, kk], Bl[nn, kk] = float(hi[o]), float(lo[o])
    return Bh, Bl


def phase_table(th, scale):
    s, c = np.sin(th / 2).astype(np.float32), np.cos(th / 2).astype(np.float32)
    def cm(a, b): return np.complex64(a) * np.complex64(b)
    b0 = cm(c[0] - 1j * s[0], c[1] - 1j * s[1]); b1 = cm(c[0] + 1j * s[0], c[1] - 1j * s[1])
    e2 = c[2] - 1j * s[2]
    l = [cm(b0, e2), cm(b1, e2), cm(np.conj(b1), e2), cm(np.conj(b0), e2)]
    c4, s4 = c[4] * scale, s[4] * scale
    h0 = cm(c[3] - 1j * s[3], c4 - 1j * s4); h1 = cm(c[3] + 1j * s[3], c4 - 1j * s4)
    p = []
    for z in range(16):
        j = z & 7
        x = l[j] if j < 4 else np.conj(l[7 - j])
        p.append(cm(x, h0 if z < 8 else h1))
    return p


def split16(x):
    bits = x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    hi = bits.view(np.float32)
    lo = (x.astype(np.float32) - hi).astype(np.float16).astype(np.float32)
    return hi.astype(np.float16).astype(np.float32), lo


def tc_forward(x, w, depths, hdiag, exact=False):
    B = x.shape[0]
    K = len(depths)
    s0 = 0
    imgs = []
    for k, d in enumerate(depths):
        M = block_matrix(w, s0, d, k == K - 1)
        s0 += d
        imgs.append(image(M) if not exact else M)
    out = np.zeros(B)
    for b in range(B):
        st = np.zeros(64, np.float32)
        st[0::2] = np.float32(SA / np.sqrt(32))
        for k in range(K):
            p = phase_table(x[b, k * n:(k + 1) * n].astype(np.float32), 1.0 if k == 0 else 1.0 / SB)
            amp = (st[0::2] + 1j * st[1::2]).astype(np.complex64)
            for z in range(32):
                amp[z] = amp[z] * (p[z] if z < 16 else np.conj(p[31 - z]))
            if exact:
                newamp = imgs[k] @ amp.astype(complex) * SB
                st = np.zeros(64); st[0::2] = newamp.real; st[1::2] = newamp.imag
                continue
            a = np.zeros(64, np.float32); a[0::2] = amp.real; a[1::2] = amp.imag
            ah, al = split16(a)
            Bh, Bl = imgs[k]
            d = (Bh @ ah.astype(np.float64) + Bl @ ah.astype(np.float64) + Bh @ al.astype(np.float64))
            st = d.astype(np.float32)
        re, im = st[0::2].astype(np.float64), st[1::2].astype(np.float64)
        out[b] = np.sum(hdiag * (re * re + im * im)) / (SA * SB) ** 2
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for depths in ([1], [2, 1], [2] * 6, [2] * 60):
        K, S = len(depths), sum(depths)
        B = 16 if K < 60 else 8
        x = rng.uniform(-np.pi, np.pi, (B, n * K))
        w = rng.uniform(-np.pi, np.pi, (S, 3, n))
        blocks = [(n, d) for d in depths]
        ham = orc.ham_from_bound(n)
        ref = orc.hea_forward(x, w, n, blocks, ham)
        hd = np.array([n - 2 * bin(z).count("1") for z in range(N)], float)
        ex = tc_forward(x, w, depths, hd, exact=True)
        em = tc_forward(x, w, depths, hd)
        e1 = np.linalg.norm(ex - ref) / np.linalg.norm(ref)
        e2 = np.linalg.norm(em - ref) / np.linalg.norm(ref)
        print(f"depths K={K}: exact-matrix formulation rel-L2 {e1:.2e}; f16x3 emulation rel-L2 {e2:.2e}")
