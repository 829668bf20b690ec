#!/bin/bash
# torchrun check of the training launcher at N ranks: B200Solver (CUDA-graph replay of the step, fused exchange)
# usage: scripts/dp_solver_check.sh N
set -e
N=${1:-2}
cd "$(dirname "$0")/.."
T=$(mktemp -d)
python - "$T" <<'PY'
import sys, json, numpy as np
t = sys.argv[1]
rng = np.random.default_rng(4)
branch = rng.standard_normal((1600, 5)); tt = rng.random((1600, 1)); y = np.cos(branch[:, :1]) * tt
np.savez(t + "/data.npz", train_branch_input=branch, train_trunk_input=tt, train_output=y,
         test_branch_input=branch[:40], test_trunk_input=tt[:40], test_output=y[:40])
cfg = {"model_type": "QuanONet", "num_qubits": 3, "net_size": [2, 1, 2, 1], "scale_coeff": 0.3,
       "if_trainable_freq": "true", "learning_rate": 0.02, "num_epochs": 5, "batch_size": 200,
       "output_dir": t + "/Demo_QuanONet_Net2-1-2-1_Q3_TF_S0.3_1600x1_Seed0", "quantum_backend": "torchquantum"}
json.dump(cfg, open(t + "/cfg.json", "w"))
PY
PYTHONPATH=$PWD timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port 29571 -m quanonet_b200.train_cli --config $T/cfg.json --data $T/data.npz 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -4
