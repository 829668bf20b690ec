"""GPU check of the tensor-core tier's GEMM-form weight gradients (csrc/hea_tc3.cuh): raw outer-product accumulator of
the first tile vs its exact value, gradients vs the fp64 oracle / the FFMA2 register kernel / the per-sublayer-moment
tensor-core kernel, bitwise determinism, the fused encoding + MSE training step (mode 5) and timings.
    python tests/harness/tc_check_outer.py [quick]
"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))
from oracle import hea_oracle as orc
from quanonet_b200 import _lib
from quanonet_b200.ops import _backward_impl, encoded_mse_step
import tc_emulate as emu

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
lib = _lib.load()
dev = torch.device("cuda:0")
n = 5
results = {}
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, float) - np.asarray(b, float)) / max(np.linalg.norm(np.asarray(b, float)), 1e-300))
TIERS = {"ffma2": 0, "outer": 1, "strings": 2}


def cfg(tc, dbg=None, err=None):
    lib.qon_tensor_tier(int(tc), 0, None if dbg is None else dbg.data_ptr(), None if err is None else err.data_ptr())


def expected_outer(x, w, depths, hdiag, g):
    """D[m][n'] of the LAST block for the first tile (exact arithmetic), and E"""
    B, K = x.shape[0], len(depths)
    Ms, s0 = [], 0
    for k, d in enumerate(depths):
        Ms.append(emu.block_matrix(w, s0, d, k == K - 1)); s0 += d
    amp = np.full((B, 32), 1 / np.sqrt(32), complex)
    for k in range(K):
        th = x[:, k * n:(k + 1) * n]
        ph = np.ones((B, 32), complex)
        for q in range(n):
            bit = (np.arange(32) >> q) & 1
            ph *= np.exp(np.where(bit[None, :] == 0, -1j, 1j) * th[:, q:q + 1] / 2)
        amp = (amp * ph) @ Ms[k].T
    gmax = np.float32(np.max(np.abs(g.astype(np.float32))))
    E = float(np.uint32((gmax.view(np.uint32) & np.uint32(0x7F800000)) + np.uint32(0x00800000)).view(np.float32))
    hmax = np.max(np.abs(hdiag))
    psi = amp[:128] * emu.SA
    lam = amp[:128] * (hdiag / hmax)[None, :] * (g[:128] / E)[:, None] * emu.SA
    il = lambda v: np.stack([v.real, v.imag], -1).reshape(v.shape[0], -1)
    return il(lam).T @ il(psi), E


def case(depths, B, seed, need_gx=True, dump=False, gscale=1.0):
    rng = np.random.default_rng(seed)
    K, S = len(depths), sum(depths)
    x = rng.uniform(-np.pi, np.pi, (B, n * K)); w = rng.uniform(-np.pi, np.pi, (S, 3, n)); g = rng.normal(size=B) * gscale
    nref = min(B, 256)
    blocks = [(n, d) for d in depths]
    xt = torch.tensor(x, dtype=torch.float32, device=dev); wt = torch.tensor(w, dtype=torch.float32, device=dev)
    gt = torch.tensor(g, dtype=torch.float32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    dbg = torch.zeros(128 * 32, dtype=torch.float32, device=dev) if dump else None
    full, pref = {}, {}
    for name, tc in TIERS.items():
        cfg(tc, dbg if name == "outer" else None, err if tc else None)
        full[name] = [None if t is None else t.double().cpu().numpy() for t in _backward_impl(gt, xt, wt, n, list(depths), None, 0, 0.0, 1.0, 0, need_gx)]
        pref[name] = [None if t is None else t.double().cpu().numpy() for t in _backward_impl(gt[:nref].contiguous(), xt[:nref].contiguous(), wt, n, list(depths), None, 0, 0.0, 1.0, 0, need_gx)]
        torch.cuda.synchronize()
    cfg(1, None, err)
    again = _backward_impl(gt, xt, wt, n, list(depths), None, 0, 0.0, 1.0, 0, need_gx)
    torch.cuda.synchronize()
    o_ref, gx_ref, gw_ref = orc.hea_forward_backward(x[:nref], w, n, blocks, orc.ham_from_bound(n), g[:nref])
    r = dict(err=int(err.item()), deterministic=bool(np.array_equal(again[2].double().cpu().numpy(), full["outer"][2])))
    for name in TIERS:
        r[f"{name}_out"] = rel(pref[name][0][:, 0], o_ref)
        r[f"{name}_gw"] = rel(pref[name][2], gw_ref)
        if need_gx:
            r[f"{name}_gx"] = rel(pref[name][1], gx_ref)
    r["outer_vs_ffma2_gw"] = rel(full["outer"][2], full["ffma2"][2])
    r["strings_vs_ffma2_gw"] = rel(full["strings"][2], full["ffma2"][2])
    if need_gx:
        r["outer_vs_ffma2_gx"] = rel(full["outer"][1], full["ffma2"][1])
    print(f"K={K} S={S} B={B} gx={need_gx} gscale={gscale:g}: " + "  ".join(f"{k} {v:.2e}" if isinstance(v, float) else f"{k} {v}" for k, v in r.items()), flush=True)
    results[f"grad_K{K}_S{S}_B{B}_gx{int(need_gx)}_{gscale:g}"] = r
    if dump:
        hd = np.array([n - 2 * bin(z).count("1") for z in range(32)], float)
        D, E = expected_outer(x, w, depths, hd, g)
        raw = dbg.cpu().numpy().reshape(4, 32, 32)        # [warp = TMEM subpartition][thread][register of the 16x256b.x8 load]
        got = np.zeros((64, 64))                          # natural order: [lam component 2 i + c][psi component]
        for wq in range(4):
            for T in range(32):
                for gI in range(8):
                    i, cp = 8 * wq + T // 4, T % 4
                    for e in range(4):
                        got[2 * i + (e >> 1), 8 * gI + 2 * cp + (e & 1)] = raw[wq, T, 4 * gI + e]
        np.savez(os.path.join(OUT, f"tc_outer_dbg_K{K}.npz"), raw=raw, expected=D, E=E)
        print(f"   outer-product accumulator, first tile, last block (16x256b fragment, planar lam rows): vs exact {rel(got, D):.2e}", flush=True)
        results[f"outer_raw_K{K}"] = rel(got, D)
    return r


def mse_case(B, seed, net=(40, 2, 20, 2), time_it=True):
    """the bench's training step: fused encoding + MSE + adjoint gradients (mode 5)"""
    bd, bl, td, tl = net
    depths = [tl] * td + [bl] * bd
    K, S = len(depths), sum(depths)
    g = torch.Generator().manual_seed(seed)
    branch = torch.randn(B, 100, generator=g).to(dev); trunk = torch.rand(B, 2, generator=g).to(dev)
    y = torch.randn(B, generator=g).to(dev)
    fw = (torch.randn(n * K, generator=g) * 0.3).to(dev); fb = ((torch.rand(n * K, generator=g) * 2 - 1) * np.pi).to(dev)
    w = ((torch.rand(S, 3, n, generator=g) * 2 - 1) * np.pi).to(dev)
    bias = torch.tensor([0.05], device=dev)
    outs = {}
    for name, tc in TIERS.items():
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        cfg(tc, None, err if tc else None)
        gw, gfw, gfb, sums = encoded_mse_step(trunk, branch, fw, fb, td, w, y, bias, 2.0 / B, n, depths, None, 0, 0.0, 1.0, 0, True)
        torch.cuda.synchronize()
        outs[name] = [t.double().cpu().numpy() for t in (gw, gfw, gfb, sums)] + [int(err.item())]
    for name in ("outer", "strings"):
        r = {k: rel(outs[name][i], outs["ffma2"][i]) for i, k in enumerate(("gw", "gfw", "gfb", "sums"))}
        r["err"] = outs[name][4]
        print(f"mse step B={B}: {name} vs ffma2 " + "  ".join(f"{k} {v:.2e}" if isinstance(v, float) else f"{k} {v}" for k, v in r.items()), flush=True)
        results[f"mse_{name}_B{B}"] = r
    if not time_it:
        return
    res = {}
    for name, tc in TIERS.items():
        cfg(tc)
        for _ in range(3):
            encoded_mse_step(trunk, branch, fw, fb, td, w, y, bias, 2.0 / B, n, depths, None, 0, 0.0, 1.0, 0, True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            encoded_mse_step(trunk, branch, fw, fb, td, w, y, bias, 2.0 / B, n, depths, None, 0, 0.0, 1.0, 0, True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res[name] = dict(ms=ms, samples_per_s=B / ms * 1e3)
        print(f"   training-step kernels B={B}: {name}: {ms:.3f} ms  {B / ms * 1e3:.3e} samples/s", flush=True)
    results[f"bench_mse_B{B}"] = res


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    case([1], 128, 1, dump=True)
    case([2, 1], 300, 2, dump=True)
    if not quick:
        case([2] * 6, 1000, 3, gscale=1e-4)
        case([1, 3, 2] * 4, 3000, 4, need_gx=False, gscale=37.0)
        case([2] * 60, 2000, 5)
        mse_case(20000, 7)
        mse_case(1_000_000, 8)
    cfg(1)
    json.dump(results, open(os.path.join(OUT, "tc_check_outer.json"), "w"), indent=1)
