"""torchrun target: B200Solver under data parallelism — the training step replayed as a CUDA graph (peer-memory
exchange inside the graph, tensor learning rate read at replay time) must reproduce the eager data-parallel loop, on
every rank, with an lr scheduler; evaluate() reloads the best checkpoint on every rank.
usage: torchrun --nproc-per-node N scripts/dp_solver_graph_check.py"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from quanonet_b200.solvers.solver_pt import B200Solver

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rng = np.random.default_rng(3)
branch = rng.standard_normal((192, 6)); t = rng.random((192, 2)); y = np.sin(branch[:, :1] + t[:, :1])
data = {"train_branch_input": branch, "train_trunk_input": t, "train_output": y,
        "test_branch_input": branch[:32], "test_trunk_input": t[:32], "test_output": y[:32]}
tmp = tempfile.mkdtemp() if rank == 0 else None
box = [tmp]
dist.broadcast_object_list(box, src=0)
out = {}
for mode in (True, False):
    cfg = {"model_type": "QuanONet", "num_qubits": 5, "net_size": [2, 2, 2, 1], "scale_coeff": 0.4, "if_trainable_freq": "true",
           "learning_rate": 0.01, "num_epochs": 6, "batch_size": 48, "seed": 2, "cuda_graph": mode, "lr_scheduler": "step",
           "lr_scheduler_kwargs": {"step_size": 2, "gamma": 0.5}, "output_dir": os.path.join(box[0], f"run_{int(mode)}")}
    torch.manual_seed(1)
    s = B200Solver(cfg, data, device=f"cuda:{local}")
    assert s.use_graph == mode, (s.use_graph, mode, type(s.trainer._all_reduce))
    hist = s.train()["loss_train"]
    params = torch.cat([v.reshape(-1).float() for v in s.model.state_dict().values()])
    chk = params.clone(); dist.broadcast(chk, src=0)
    assert torch.equal(chk, params), "replicas diverged"
    metrics = s.evaluate()
    m = torch.tensor([metrics["rel_l2"]], device=dev, dtype=torch.float64); m0 = m.clone(); dist.broadcast(m0, src=0)
    assert torch.equal(m, m0), "evaluate() differs across ranks"
    out[mode] = (hist, params)
    if rank == 0:
        assert os.path.exists(os.path.join(cfg["output_dir"], "best_model.pt")) and os.path.exists(os.path.join(cfg["output_dir"], "final_model.npz"))
assert np.allclose(out[True][0], out[False][0], rtol=2e-4), (out[True][0], out[False][0])
assert torch.allclose(out[True][1], out[False][1], rtol=2e-4, atol=2e-5)
if rank == 0:
    print(f"world={world}: CUDA-graph replay of the data-parallel step == eager (6 epochs, step lr schedule); loss {out[True][0][0]:.5f} -> {out[True][0][-1]:.5f}", flush=True)
dist.barrier(); dist.destroy_process_group()
