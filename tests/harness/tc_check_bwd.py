"""GPU check of the tensor-core tier's fused forward + adjoint backward (csrc/hea_tc2.cuh): gradients vs the fp64
oracle and vs the FFMA2 register kernel, step-by-step state dump vs the exact emulation on a small case, the fused
encoding + MSE training step (mode 5), and timings at B = 1M.
    python tests/harness/tc_check_bwd.py
"""
import ctypes, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))
from oracle import hea_oracle as orc
from quanonet_b200 import _lib
from quanonet_b200.ops import _backward_impl, _forward_impl, encoded_mse_step
import tc_emulate as emu
import tc_emulate_bwd as emub

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
lib = _lib.load()
dev = torch.device("cuda:0")
n = 5
results = {}
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, float) - np.asarray(b, float)) / max(np.linalg.norm(np.asarray(b, float)), 1e-300))


def cfg(tc, dbg=None, err=None):
    lib.qon_tensor_tier(int(tc), 0, None if dbg is None else dbg.data_ptr(), None if err is None else err.data_ptr())


def case(depths, B, seed, need_gx=True, dump=False):
    rng = np.random.default_rng(seed)
    K, S = len(depths), sum(depths)
    x = rng.uniform(-np.pi, np.pi, (B, n * K)); w = rng.uniform(-np.pi, np.pi, (S, 3, n)); g = rng.normal(size=B)
    nref = min(B, 256)
    blocks = [(n, d) for d in depths]
    # oracle on the first nref rows (gradient of the shared weights from those rows only: compare via a second run)
    xt = torch.tensor(x, dtype=torch.float32, device=dev); wt = torch.tensor(w, dtype=torch.float32, device=dev)
    gt = torch.tensor(g, dtype=torch.float32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    dbg = torch.zeros((K + S) * 128 * 128, dtype=torch.float32, device=dev) if dump else None
    cfg(1, dbg, err)
    o_tc, gx_tc, gw_tc = _backward_impl(gt, xt, wt, n, list(depths), None, 0, 0.0, 1.0, 0, need_gx)
    torch.cuda.synchronize()
    cfg(0)
    o_rg, gx_rg, gw_rg = _backward_impl(gt, xt, wt, n, list(depths), None, 0, 0.0, 1.0, 0, need_gx)
    torch.cuda.synchronize()
    # oracle on a prefix (its weight gradient needs the same rows -> rerun TC on the prefix)
    o_ref, gx_ref, gw_ref = orc.hea_forward_backward(x[:nref], w, n, blocks, orc.ham_from_bound(n), g[:nref])
    cfg(1, None, err)
    o_p, gx_p, gw_p = _backward_impl(gt[:nref].contiguous(), xt[:nref].contiguous(), wt, n, list(depths), None, 0, 0.0, 1.0, 0, need_gx)
    torch.cuda.synchronize()
    cfg(0)
    o_q, gx_q, gw_q = _backward_impl(gt[:nref].contiguous(), xt[:nref].contiguous(), wt, n, list(depths), None, 0, 0.0, 1.0, 0, need_gx)
    torch.cuda.synchronize()
    r = dict(
        tc_out=rel(o_p.cpu().numpy()[:, 0], o_ref), tc_gw=rel(gw_p.cpu().numpy(), gw_ref),
        ffma2_out=rel(o_q.cpu().numpy()[:, 0], o_ref), ffma2_gw=rel(gw_q.cpu().numpy(), gw_ref),
        tc_vs_ffma2_out=rel(o_tc.cpu().numpy(), o_rg.cpu().numpy()), tc_vs_ffma2_gw=rel(gw_tc.cpu().numpy(), gw_rg.cpu().numpy()),
        err=int(err.item()))
    if need_gx:
        r.update(tc_gx=rel(gx_p.cpu().numpy(), gx_ref), ffma2_gx=rel(gx_q.cpu().numpy(), gx_ref),
                 tc_vs_ffma2_gx=rel(gx_tc.cpu().numpy(), gx_rg.cpu().numpy()))
    print(f"K={K} S={S} B={B} gx={need_gx}: " + "  ".join(f"{k} {v:.2e}" if isinstance(v, float) else f"{k} {v}" for k, v in r.items()), flush=True)
    results[f"grad_K{K}_S{S}_B{B}_gx{int(need_gx)}"] = r
    if dump:
        d = dbg.cpu().numpy().reshape(K + S, 128, 128)
        np.savez(os.path.join(OUT, f"tc_bwd_dbg_K{K}_S{S}.npz"), dbg=d, x=x, w=w, g=g, depths=np.array(depths))
        # exact emulation of the same steps for the first 8 samples
        hd = np.array([n - 2 * bin(z).count("1") for z in range(32)], float)
        exp = expected_steps(x[:8], w, depths, hd)
        for step in range(K + S):
            a = d[step, :8, :64]; e = exp[step][0]
            line = f"   step {step}: psi {np.linalg.norm(a / emu.SA - e) / np.linalg.norm(e):.2e}"
            if step >= K:
                al = d[step, :8, 64:]; el = exp[step][1]
                # lam_hat is normalised: compare directions
                na, ne = al / np.linalg.norm(al, axis=1, keepdims=True), el / np.linalg.norm(el, axis=1, keepdims=True)
                line += f"  lam (direction) {np.linalg.norm(na - ne) / np.linalg.norm(ne):.2e}"
            print(line, flush=True)
    return r


def expected_steps(x, w, depths, hdiag):
    """states (as interleaved re/im rows) after every GEMM of the kernel's step sequence, exact arithmetic"""
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
    Gs = [emub.rev_matrix(w, s, lastof[s], True, blk[s] < K - 1) for s in range(S)]
    steps = [[np.zeros((B, 64)), np.zeros((B, 64))] for _ in range(K + S)]
    il = lambda v: np.stack([v.real, v.imag], -1).reshape(-1)
    for b in range(B):
        amp = np.full(32, 1 / np.sqrt(32), complex); phs = []
        for k in range(K):
            th = x[b, k * n:(k + 1) * n]
            ph = np.array([np.prod([np.exp((-1j if not (z >> q) & 1 else 1j) * th[q] / 2) for q in range(n)]) for z in range(32)])
            phs.append(ph); amp = Ms[k] @ (amp * ph); steps[k][0][b] = il(amp)
        psi, lam = amp, hdiag * amp
        st = K
        for s in reversed(range(S)):
            if last[s]:
                opsi, olam = psi, lam
            psi, lam = Gs[s] @ opsi, Gs[s] @ olam
            steps[st][0][b], steps[st][1][b] = il(psi), il(lam)
            st += 1
            if first[s]:
                psi, lam = np.conj(phs[blk[s]]) * psi, np.conj(phs[blk[s]]) * lam
    return steps


def mse_case(B, seed, net=(40, 2, 20, 2)):
    """the bench's training step: fused encoding + MSE + adjoint gradients (mode 5), TC vs FFMA2"""
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
    for name, tc in (("ffma2", 0), ("tc", 1)):
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        cfg(tc, None, err)
        gw, gfw, gfb, sums = encoded_mse_step(trunk, branch, fw, fb, td, w, y, bias, 2.0 / B, n, depths, None, 0, 0.0, 1.0, 0, True)
        torch.cuda.synchronize()
        outs[name] = [t.double().cpu().numpy() for t in (gw, gfw, gfb, sums)] + [int(err.item())]
    r = {k: rel(outs["tc"][i], outs["ffma2"][i]) for i, k in enumerate(("gw", "gfw", "gfb", "sums"))}
    r["err"] = outs["tc"][4]
    print(f"mse step B={B}: tc vs ffma2 " + "  ".join(f"{k} {v:.2e}" if isinstance(v, float) else f"{k} {v}" for k, v in r.items()), flush=True)
    results[f"mse_B{B}"] = r
    # timing
    res = {}
    for name, tc in (("ffma2", 0), ("tc", 1)):
        cfg(tc)
        for _ in range(2):
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
    case([1], 128, 1, dump=True)
    case([2, 1], 128, 2, dump=True)
    case([2] * 6, 1000, 3)
    case([1, 3, 2] * 4, 3000, 4, need_gx=False)
    case([2] * 60, 2000, 5)
    mse_case(20000, 7)
    mse_case(1_000_000, 8)
    json.dump(results, open(os.path.join(OUT, "tc_check_bwd.json"), "w"), indent=1)
