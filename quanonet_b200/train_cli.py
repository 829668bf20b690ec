"""Training launcher for the B200 path: ``python -m quanonet_b200.train_cli --config cfg.json --data data.npz``
or, batch-sharded over the GPUs of one node,

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        -m quanonet_b200.train_cli --config cfg.json --data data.npz

Plays the role of the reference's ``main.py`` for the PyTorch quantum route (parse → route → seed → device →
solver.train → evaluate, ``main.py:16-122``) without pinning ``CUDA_VISIBLE_DEVICES="0"`` (``main.py:52``):
every rank takes the GPU ``LOCAL_RANK`` names.  ``--config`` is a JSON dict with the reference's keys
(``utils/common.py:97-152``: model_type, num_qubits, net_size, scale_coeff, if_trainable_freq, ham_bound,
ham_diag, ham_pauli, learning_rate, num_epochs, batch_size, optimizer, lr_scheduler, seed, …); ``--data`` is an
``.npz`` with the arrays of ``DataManager.get_data()`` (``data_utils/data_manager.py:74-106``; data generation
itself is outside this package).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

from .solvers.solver_pt import B200Solver
from .utils.backend import backend


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--config", required=True)
    ap.add_argument("--data", required=True)
    ap.add_argument("--output_dir", default=None)
    ap.add_argument("--num_epochs", type=int, default=None)
    ap.add_argument("--batch_size", type=int, default=None)
    args = ap.parse_args(argv)
    with open(args.config) as f:
        cfg = json.load(f)
    for k in ("output_dir", "num_epochs", "batch_size"):
        if getattr(args, k) is not None:
            cfg[k] = getattr(args, k)
    route = backend.check_compatibility(cfg.get("model_type", "QuanONet"), cfg.get("quantum_backend", "torchquantum"))
    if route != "pytorch_quantum":
        raise SystemExit(f"route {route!r} is outside quanonet_b200")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("quanonet_b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    seed = int(cfg.get("seed", cfg.get("random_seed", 0)))
    np.random.seed(seed)
    torch.manual_seed(seed)                                   # identical initial parameters on every rank
    with np.load(args.data) as z:
        data = {k: z[k] for k in z.files}
    solver = B200Solver(cfg, data, device=f"cuda:{local}")
    history = solver.train()
    metrics = solver.evaluate()
    if int(os.environ.get("RANK", "0")) == 0:
        out = {"metrics": metrics, "final_train_loss": history["loss_train"][-1], "epochs": len(history["loss_train"])}
        print(json.dumps(out))
        if cfg.get("output_dir"):
            with open(os.path.join(cfg["output_dir"], "metric.json"), "w") as f:
                json.dump(out, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
