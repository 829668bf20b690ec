"""torchrun target: soak test of the fused finalize + exchange step — many small steps with random per-rank delays
(host sleeps and device spins), replicas must stay bit-identical and no peer wait may time out.
usage: torchrun --nproc-per-node N scripts/dp_soak.py [steps]"""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from quanonet_b200.train import DataParallelTrainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
tr = DataParallelTrainer(bench.make_model(dev, seed=0), lr=1e-4)
assert tr._fused_exchange
rng = random.Random(100 + rank)
batches = [bench.synth_batch(100, seed=17 * i + rank, device=dev) for i in range(8)]
t0 = time.time()
for it in range(steps):
    branch, trunk, y = batches[it % 8]
    r = rng.random()
    if r < 0.01:
        time.sleep(rng.random() * 0.02)                 # host hiccup on this rank only
    elif r < 0.05:
        torch.cuda._sleep(int(rng.random() * 2e6))      # device-side delay on this rank only
    loss = tr.step((branch, trunk), y)
    if it % 5000 == 4999:
        chk = tr.flat_param.clone(); dist.broadcast(chk, src=0)
        assert torch.equal(chk, tr.flat_param), f"replicas diverged at step {it}"
        assert torch.isfinite(loss).item()
torch.cuda.synchronize()
assert not tr._all_reduce.timed_out()
chk = tr.flat_param.clone(); dist.broadcast(chk, src=0)
assert torch.equal(chk, tr.flat_param)
if rank == 0:
    print(f"world={world}: {steps} fused-exchange steps with random rank delays in {time.time() - t0:.1f} s; "
          f"replicas bit-identical, no timeouts, final loss {float(loss):.4f}", flush=True)
dist.barrier(); dist.destroy_process_group()
