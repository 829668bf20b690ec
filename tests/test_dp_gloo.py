"""CPU test of the N>1 path: world_size-2 gloo run of DataParallelTrainer.  The kernel call is
replaced by an oracle-backed stand-in (``kernel_fn``) so that the sharding, the global-batch loss
normalisation, the flat-buffer all-reduce and the replicated optimiser step are what is tested."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import hea_oracle as orc


def _oracle_kernel(x, w, y, bias, grad_scale, qlayer, depths, need_gx):
    n = qlayer.n_wires
    blocks = [(n, d) for d in depths]
    ham = orc.Ham("pauli", "Z", qlayer.ham_offset, qlayer.ham_coeff)
    xn, wn = x.double().numpy(), w.detach().double().numpy()
    e = orc.hea_forward(xn, wn, n, blocks, ham)
    b = float(bias.detach()) if bias is not None else 0.0
    g = grad_scale * (e + b - y.double().numpy())
    _, gx, gw = orc.hea_forward_backward(xn, wn, n, blocks, ham, grad_out=g)
    t = lambda a: torch.tensor(a, dtype=x.dtype)
    return t(e).reshape(-1, 1), t(g), (t(gx) if need_gx else torch.empty(0)), t(gw)


def _make(seed=0, kind="quanonet"):
    from quanonet_b200.core.models_pt import HEAQNNPT, QuanONetPT
    torch.manual_seed(seed)
    if kind == "heaqnn":        # no bias parameter: the flat gradient layout carries a spare slot instead
        m = HEAQNNPT(2, 5, (3, 2), scale_coeff=0.3, if_trainable_freq=True).double()
        with torch.no_grad():
            m.freq.bias.uniform_(-1, 1)
        return m
    m = QuanONetPT(2, 4, 1, (2, 1, 2, 2), scale_coeff=0.3, if_trainable_freq=True).double()
    with torch.no_grad():
        m.bias.fill_(0.1)
        m.branch_freq.bias.uniform_(-1, 1)
    return m


def _data(B=12, kind="quanonet"):
    g = torch.Generator().manual_seed(5)
    if kind == "heaqnn":
        return (torch.randn(B, 5, generator=g, dtype=torch.float64), None,
                torch.randn(B, 1, generator=g, dtype=torch.float64))
    return (torch.randn(B, 4, generator=g, dtype=torch.float64), torch.rand(B, 1, generator=g, dtype=torch.float64),
            torch.randn(B, 1, generator=g, dtype=torch.float64))


def _inputs(branch, trunk, sl=slice(None)):
    return (branch[sl],) if trunk is None else (branch[sl], trunk[sl])


def _worker(rank, world, port, out_dir, kind="quanonet"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quanonet_b200.train import DataParallelTrainer
    model = _make(seed=rank, kind=kind)  # different init per rank: the trainer must broadcast rank 0's
    tr = DataParallelTrainer(model, lr=1e-2, optimizer="adam", kernel_fn=_oracle_kernel)
    branch, trunk, y = _data(kind=kind)
    sl = slice(rank * 6, (rank + 1) * 6)
    loss = tr.step(_inputs(branch, trunk, sl), y[sl])
    torch.save({"loss": loss.item(), "grad": tr.flat_grad.clone(),
                "params": {k: v.clone() for k, v in model.state_dict().items()}}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("kind", ["quanonet", "heaqnn"])
def test_two_rank_gloo_matches_single_process(tmp_path, kind):
    from quanonet_b200.train import DataParallelTrainer
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), kind), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    # replicas stay identical
    assert r0["loss"] == pytest.approx(r1["loss"], rel=1e-12)
    assert torch.equal(r0["grad"], r1["grad"])
    for k in r0["params"]:
        assert torch.equal(r0["params"][k], r1["params"][k]), k
    # and equal the single-process step on the whole batch
    model = _make(seed=0, kind=kind)
    tr = DataParallelTrainer(model, lr=1e-2, optimizer="adam", kernel_fn=_oracle_kernel)
    branch, trunk, y = _data(kind=kind)
    loss = tr.step(_inputs(branch, trunk), y)
    assert loss.item() == pytest.approx(r0["loss"], rel=1e-10)
    assert torch.allclose(tr.flat_grad, r0["grad"], rtol=1e-9, atol=1e-12)
    for k, v in model.state_dict().items():
        assert torch.allclose(v, r0["params"][k], rtol=1e-9, atol=1e-12), k


def test_fused_step_gradients_equal_autograd_of_the_reference_sequence():
    """compute_grads == autograd through cat(freq layers) -> circuit -> +bias -> MSELoss."""
    from quanonet_b200.train import DataParallelTrainer
    model = _make(seed=3)
    tr = DataParallelTrainer(model, lr=1e-2, optimizer="sgd", kernel_fn=_oracle_kernel)
    branch, trunk, y = _data(8)
    loss = tr.compute_grads((branch, trunk), y)
    got = {k: p.grad.clone() for k, p in model.named_parameters()}

    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w):
            ctx.save_for_backward(x, w)
            q = model.quantum_layer
            e = orc.hea_forward(x.detach().numpy(), w.detach().numpy(), 2, q.block_configs,
                                orc.Ham("pauli", "Z", q.ham_offset, q.ham_coeff))
            return torch.tensor(e).reshape(-1, 1)

        @staticmethod
        def backward(ctx, g):
            x, w = ctx.saved_tensors
            q = model.quantum_layer
            _, gx, gw = orc.hea_forward_backward(x.detach().numpy(), w.detach().numpy(), 2, q.block_configs,
                                                 orc.Ham("pauli", "Z", q.ham_offset, q.ham_coeff),
                                                 grad_out=g.numpy().reshape(-1))
            return torch.tensor(gx), torch.tensor(gw)

    ref = _make(seed=3)
    x = torch.cat([ref.trunk_freq(trunk), ref.branch_freq(branch)], dim=1)
    pred = _Fn.apply(x, ref.quantum_layer.ansatz_weights) + ref.bias
    ref_loss = torch.nn.functional.mse_loss(pred, y)
    ref_loss.backward()
    assert loss.item() == pytest.approx(ref_loss.item(), rel=1e-12)
    for k, p in ref.named_parameters():
        assert torch.allclose(got[k], p.grad, rtol=1e-10, atol=1e-13), k
