"""torchrun target: the data-parallel training step with the exchange fused into the finalize kernel
(qon_encoded_mse_step_dp) vs the same step with a separate peer all-reduce vs NCCL.
usage: torchrun --nproc-per-node N scripts/dp_fused_exchange_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from quanonet_b200.train import DataParallelTrainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

def trainer(mode):
    os.environ["QON_COLLECTIVE"] = "nccl" if mode == "nccl" else "peer"
    tr = DataParallelTrainer(bench.make_model(dev, seed=0), lr=1e-3, optimizer_kwargs={"capturable": True})
    if mode == "peer":
        tr._fused_exchange = False
    return tr

for B in (100, 5000, 20000):          # latency tier, latency tier, throughput tier
    fused, peer, nccl = trainer("fused"), trainer("peer"), trainer("nccl")
    assert fused._fused_exchange and not peer._fused_exchange
    worst = 0.0
    for it in range(20):
        branch, trunk, y = bench.synth_batch(B, seed=1000 * it + rank, device=dev)
        lf = fused.step((branch, trunk), y); lp = peer.step((branch, trunk), y); ln = nccl.step((branch, trunk), y)
        assert torch.equal(fused.flat_grad, peer.flat_grad), (B, it, (fused.flat_grad - peer.flat_grad).abs().max())
        assert torch.equal(fused.flat_param, peer.flat_param)
        worst = max(worst, float(((fused.flat_grad - nccl.flat_grad).norm() / nccl.flat_grad.norm())))
        chk = fused.flat_param.clone(); dist.broadcast(chk, src=0)
        assert torch.equal(chk, fused.flat_param), "replicas diverged"
    assert worst < 1e-5, worst
    assert not fused._all_reduce.timed_out()

    def timeit(tr, n=200):
        branch, trunk, y = bench.synth_batch(B, seed=7 + rank, device=dev)
        for _ in range(10): tr.step((branch, trunk), y)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n): tr.step((branch, trunk), y)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    tf, tp, tn = timeit(fused), timeit(peer), timeit(nccl)
    if rank == 0:
        print(f"world={world} B={B}/gpu: eager step fused-exchange {tf:.1f} us | separate peer all-reduce {tp:.1f} us | "
              f"NCCL {tn:.1f} us; grads bit-equal (fused vs peer), rel diff vs NCCL {worst:.1e}", flush=True)

# HEAQNN with fixed frequencies (no frequency gradients, no bias): the flat layout is [ansatz | spare | sse | pad]
from quanonet_b200.core.models_pt import HEAQNNPT
def heaqnn(mode):
    os.environ["QON_COLLECTIVE"] = "nccl" if mode == "nccl" else "peer"
    torch.manual_seed(3)
    m = HEAQNNPT(4, 9, (6, 2), scale_coeff=0.3, if_trainable_freq=False).to(dev)
    tr = DataParallelTrainer(m, lr=1e-3)
    if mode == "peer":
        tr._fused_exchange = False
    return tr
fused, peer, nccl = heaqnn("fused"), heaqnn("peer"), heaqnn("nccl")
assert fused._fused_exchange
for it in range(10):
    gen = torch.Generator().manual_seed(50 * it + rank)
    u = torch.randn(300, 9, generator=gen).to(dev); y = torch.randn(300, 1, generator=gen).to(dev)
    fused.step((u,), y); peer.step((u,), y); nccl.step((u,), y)
    assert torch.equal(fused.flat_grad, peer.flat_grad)
    assert torch.allclose(fused.flat_grad, nccl.flat_grad, rtol=1e-5, atol=1e-7)
    assert torch.equal(fused.flat_param, peer.flat_param)
if rank == 0:
    print(f"world={world} HEAQNN fixed-frequency: fused exchange == separate peer all-reduce (bitwise), ~= NCCL", flush=True)

# n = 7: fused encoding + fused exchange on the wide latency tier (decided per batch size)
from quanonet_b200.core.models_pt import QuanONetPT
def q7(mode):
    os.environ["QON_COLLECTIVE"] = "nccl" if mode == "nccl" else "peer"
    torch.manual_seed(5)
    m = QuanONetPT(7, 9, 2, (3, 2, 2, 1), scale_coeff=0.3, if_trainable_freq=True).to(dev)
    tr = DataParallelTrainer(m, lr=1e-3)
    if mode == "peer":
        tr._fused_exchange = False
    return tr
fused, peer, nccl = q7("fused"), q7("peer"), q7("nccl")
assert fused._fused_exchange and fused._enc_by_batch is not None
for it in range(10):
    gen = torch.Generator().manual_seed(70 * it + rank)
    b = torch.randn(200, 9, generator=gen).to(dev); t = torch.rand(200, 2, generator=gen).to(dev)
    y = torch.randn(200, 1, generator=gen).to(dev)
    fused.step((b, t), y); peer.step((b, t), y); nccl.step((b, t), y)
    assert torch.equal(fused.flat_grad, peer.flat_grad)
    assert torch.allclose(fused.flat_grad, nccl.flat_grad, rtol=1e-5, atol=1e-7)
    assert torch.equal(fused.flat_param, peer.flat_param)
assert fused._enc_by_batch == {200: True}
if rank == 0:
    print(f"world={world} QuanONet Q7 (wide latency tier): fused exchange == separate peer all-reduce (bitwise), ~= NCCL", flush=True)
dist.barrier(); dist.destroy_process_group()
