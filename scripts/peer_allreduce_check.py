"""torchrun target: peer-memory all-reduce (csrc/qon_peer.cuh) vs NCCL — equality and latency.
usage: torchrun --nproc-per-node N scripts/peer_allreduce_check.py [iters]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from quanonet_b200.comm import PeerAllReduce

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
L = 2402
par = PeerAllReduce(L, dev)
g = torch.Generator(device=dev).manual_seed(1234 + rank)
bad = 0
for it in range(iters):
    v = torch.randn(L, generator=g, device=dev)
    ref = v.clone(); dist.all_reduce(ref)
    out = par(v.clone())
    if it % 7 == 0:                      # let ranks drift apart now and then
        torch.cuda._sleep(int(2e5 * (rank + 1)))
    # NCCL's summation order differs from rank order: compare with a tolerance, and bitwise across ranks
    if not torch.allclose(out, ref, rtol=1e-5, atol=1e-5): bad += 1
    chk = out.clone(); dist.broadcast(chk, src=0)
    if not torch.equal(chk, out): bad += 1
torch.cuda.synchronize()
assert not par.timed_out(), "a peer wait timed out"
assert bad == 0, f"{bad} mismatches"

def timeit(fn, n=500):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
v = torch.randn(L, device=dev)
t_peer = timeit(lambda: par(v))
t_nccl = timeit(lambda: dist.all_reduce(v))
# graph replay of the peer kernel (device-side epoch counter keeps it replayable)
gr = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s): par(v)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(gr): par(v)
t_graph = timeit(gr.replay)
assert not par.timed_out()
if rank == 0:
    print(f"world={world} len={L}: peer {t_peer:.1f} us | peer (graph replay) {t_graph:.1f} us | nccl {t_nccl:.1f} us; {iters} iterations equal", flush=True)
dist.barrier(); dist.destroy_process_group()
