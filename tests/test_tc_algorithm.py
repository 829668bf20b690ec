"""CPU checks of the ALGORITHM the tensor-core tier implements (csrc/hea_tc.cuh, hea_tc2.cuh), against the fp64 oracle:
the Hadamard-basis reformulation with pre-fused block unitaries, the split-f16 three-product GEMM (precision), the
composite un-apply matrices and the Pauli-string moments of the adjoint sweep.  The kernels themselves are checked on
the GPU (tests/test_gpu_parity.py::test_tensor_tier_*); these run in the build container."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))


def test_pauli_string_table_is_current():
    """csrc/tc_strings.cuh is generated (and verified against dense 32x32 matrices) by scripts/gen_tc_strings.py."""
    path = os.path.join(ROOT, "quanonet_b200", "csrc", "tc_strings.cuh")
    before = open(path).read()
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_tc_strings.py")], check=True, capture_output=True)
    assert open(path).read() == before


def test_forward_reformulation_and_split_precision():
    from oracle import hea_oracle as orc
    import tc_emulate as emu
    rng = np.random.default_rng(3)
    n, depths = 5, [2, 1, 2]
    K, S, B = len(depths), sum(depths), 6
    x = rng.uniform(-np.pi, np.pi, (B, n * K))
    w = rng.uniform(-np.pi, np.pi, (S, 3, n))
    ref = orc.hea_forward(x, w, n, [(n, d) for d in depths], orc.ham_from_bound(n))
    hd = np.array([n - 2 * bin(z).count("1") for z in range(32)], float)
    exact = emu.tc_forward(x, w, depths, hd, exact=True)        # block matrices in fp64, phases in fp32
    split = emu.tc_forward(x, w, depths, hd)                    # + f16 hi/lo operands, three products
    nrm = np.linalg.norm(ref)
    assert np.linalg.norm(exact - ref) / nrm < 1e-6
    assert np.linalg.norm(split - ref) / nrm < 2e-6             # the kernels' bar is 1e-5


def test_adjoint_sweep_composites_and_string_moments():
    from oracle import hea_oracle as orc
    import tc_emulate_bwd as emub
    rng = np.random.default_rng(4)
    n = 5
    for depths in ([1], [2, 1], [1, 3]):
        K, S, B = len(depths), sum(depths), 3
        x = rng.uniform(-np.pi, np.pi, (B, n * K))
        w = rng.uniform(-np.pi, np.pi, (S, 3, n))
        g = rng.normal(size=B)
        hd = np.array([n - 2 * bin(z).count("1") for z in range(32)], float)
        o_ref, gx_ref, gw_ref = orc.hea_forward_backward(x, w, n, [(n, d) for d in depths], orc.ham_from_bound(n), g)
        o, gx, gw = emub.tc_backward(x, w, depths, hd, g)
        rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
        assert rel(o, o_ref) < 1e-12 and rel(gx, gx_ref) < 1e-12 and rel(gw, gw_ref) < 1e-12, depths


def test_gemm_form_weight_gradients():
    """csrc/hea_tc3.cuh: one batch-summed outer product per block (real 64 x 64 form, f16 hi/lo operands, power-of-two
    gradient scale, fp32 tile accumulators) and the conjugation chain of tc_moment_kernel, against the fp64 oracle."""
    from oracle import hea_oracle as orc
    import tc_emulate_outer as emo
    rng = np.random.default_rng(9)
    n = 5
    for depths, B, gs in (([1], 5, 1.0), ([2, 1], 130, 1e-4), ([1, 3, 2], 300, 50.0)):
        K, S = len(depths), sum(depths)
        x = rng.uniform(-np.pi, np.pi, (B, n * K))
        w = rng.uniform(-np.pi, np.pi, (S, 3, n))
        g = rng.normal(size=B) * gs
        hd = np.array([n - 2 * bin(z).count("1") for z in range(32)], float)
        o_ref, gx_ref, gw_ref = orc.hea_forward_backward(x, w, n, [(n, d) for d in depths], orc.ham_from_bound(n), g)
        rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
        o, gx, gw = emo.tc_backward_outer(x, w, depths, hd, g, exact=True)
        assert rel(o, o_ref) < 1e-12 and rel(gx, gx_ref) < 1e-12 and rel(gw, gw_ref) < 5e-7, depths     # fp32 operand rows
        o, gx, gw = emo.tc_backward_outer(x, w, depths, hd, g, exact=False)
        assert rel(gw, gw_ref) < 2e-6, depths                                                             # the kernels' bar is 1e-5
