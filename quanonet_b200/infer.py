"""Inference API with the reference's surface (``infer.py``: ``load_model`` :180-232, ``predict`` :235-291,
``evaluate`` :294-302), for QuanONet / HEAQNN checkpoints on the B200 kernels.

Differences from the reference, on purpose:
* ``.npz`` and MindSpore ``.ckpt`` checkpoints are routed to the PyTorch model (the reference sends both to
  MindSpore at ``infer.py:99-104``, which leaves its own npz→PT branch ``:213-226`` unreachable); the three
  shipped Q5 ``.ckpt`` files load through ``quanonet_b200.checkpoint.read_mindspore_ckpt``.
* ``predict`` runs large batches (default 262,144 rows instead of 128): one kernel launch per batch and no
  encoding matrix in memory (fused-encoding forward).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch

from .checkpoint import load_raw, ms_to_pt_arrays, parse_experiment_dir
from .core.models_pt import HEAQNNPT, QuanONetPT


def _resolve_config(ckpt_path: str, overrides: dict) -> dict:
    try:
        cfg = parse_experiment_dir(ckpt_path)
    except ValueError:
        cfg = {}
    cfg.update({k: v for k, v in overrides.items() if v is not None})
    for key in ("model_type", "net_size", "num_qubits"):
        if key not in cfg:
            raise ValueError(f"cannot determine '{key}' from {ckpt_path!r}; pass it as a keyword argument")
    cfg.setdefault("if_trainable_freq", True)
    cfg.setdefault("scale_coeff", 0.01)
    cfg.setdefault("ham_bound", (-5.0, 5.0))
    cfg.setdefault("ham_diag", None)
    cfg["_backend"] = "pytorch"
    cfg["quantum_backend"] = "torchquantum"
    return cfg


def build_model(cfg: dict, branch_in: Optional[int] = None, trunk_in: Optional[int] = None):
    """QuanONetPT / HEAQNNPT from a config dict (keys as in the reference's ``_build_pt_model``, ``infer.py:151-175``)."""
    kw = dict(num_qubits=int(cfg["num_qubits"]), net_size=tuple(cfg["net_size"]),
              scale_coeff=float(cfg["scale_coeff"]), if_trainable_freq=bool(cfg["if_trainable_freq"]),
              ham_bound=tuple(cfg["ham_bound"]), ham_diag=cfg.get("ham_diag"))
    for extra in ("ham_pauli", "diag_order"):
        if cfg.get(extra) is not None:
            kw[extra] = cfg[extra]
    if cfg["model_type"] == "QuanONet":
        if branch_in is None or trunk_in is None:
            raise ValueError("QuanONet needs branch_in and trunk_in")
        return QuanONetPT(branch_input_size=branch_in, trunk_input_size=trunk_in, **kw)
    if cfg["model_type"] == "HEAQNN":
        if branch_in is None:
            raise ValueError("HEAQNN needs branch_in (its input size)")
        return HEAQNNPT(input_size=branch_in + (trunk_in or 0), **kw)
    raise ValueError(f"model_type {cfg['model_type']!r} is outside quanonet_b200 (QuanONet / HEAQNN only)")


def load_model(ckpt_path: str, branch_in: Optional[int] = None, trunk_in: Optional[int] = None,
               device: Optional[str] = None, dtype=torch.float32, **overrides) -> Tuple[torch.nn.Module, dict]:
    """Build the model described by the checkpoint's experiment directory name (``utils/logger.py:55-118``)
    and load its weights.  Accepts ``.pt`` (PyTorch state_dict), ``.npz`` (PyTorch- or MindSpore-named) and
    MindSpore ``.ckpt``.  When the frequency-layer sizes are not given they are inferred from the checkpoint
    only if that is unambiguous; otherwise pass ``branch_in`` / ``trunk_in`` as the reference's callers do
    (``visualization.ipynb`` cell 7: ``load_model(path, branch_in=100, trunk_in=2)``)."""
    cfg = _resolve_config(ckpt_path, overrides)
    ext = os.path.splitext(ckpt_path)[1].lower()
    if ext == ".pt":
        sd = torch.load(ckpt_path, map_location="cpu")
        sd = {k: v for k, v in sd.items()}
    else:
        arrays = ms_to_pt_arrays(load_raw(ckpt_path), cfg["net_size"], cfg["num_qubits"], cfg["if_trainable_freq"],
                                 cfg["model_type"])
        sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in arrays.items()}
    model = build_model(cfg, branch_in, trunk_in)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [k for k in missing if not k.endswith("ham_diag")]
    if missing or unexpected:
        raise RuntimeError(f"checkpoint does not match the model: missing {missing}, unexpected {list(unexpected)}")
    dev = torch.device(device) if device else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    model = model.to(device=dev, dtype=dtype).eval()
    cfg["branch_in"], cfg["trunk_in"] = branch_in, trunk_in
    return model, cfg


@torch.no_grad()
def predict(model, branch_input: np.ndarray, trunk_input: Optional[np.ndarray] = None, cfg: Optional[dict] = None,
            batch_size: int = 262_144) -> np.ndarray:
    """Batched inference; returns ``(N, 1)`` float32 (same contract as the reference's ``predict``)."""
    p = next(model.parameters())
    has_trunk = trunk_input is not None and (cfg or {}).get("model_type", "QuanONet") == "QuanONet"
    n = branch_input.shape[0]
    out = np.empty((n, 1), dtype=np.float32)
    for s in range(0, n, batch_size):
        b = torch.as_tensor(np.ascontiguousarray(branch_input[s:s + batch_size]), dtype=p.dtype).to(p.device)
        if has_trunk:
            t = torch.as_tensor(np.ascontiguousarray(trunk_input[s:s + batch_size]), dtype=p.dtype).to(p.device)
            y = model(b, t)
        else:
            y = model(b)
        out[s:s + batch_size] = y.float().cpu().numpy()
    return out


def evaluate(y_pred: np.ndarray, y_true: np.ndarray) -> dict:
    """Rel-L2, MSE, MAE (reference ``infer.py:294-302``)."""
    y_pred = np.asarray(y_pred, dtype=np.float64).reshape(-1)
    y_true = np.asarray(y_true, dtype=np.float64).reshape(-1)
    diff = y_pred - y_true
    return {"rel_l2": float(np.linalg.norm(diff) / (np.linalg.norm(y_true) + 1e-8)),
            "mse": float(np.mean(diff ** 2)), "mae": float(np.mean(np.abs(diff)))}


def main(argv=None):
    """CLI in the shape of the reference's (``infer.py:333-427``): checkpoint + test data in, metrics out.
    ``--data`` is an ``.npz`` with ``test_branch_input`` / ``test_trunk_input`` (or ``test_input``) and optionally
    ``test_output``; auto-generating data is outside this package."""
    import argparse
    import json

    ap = argparse.ArgumentParser(description="QuanONet / HEAQNN inference on the B200 kernels")
    ap.add_argument("--ckpt", required=True)
    ap.add_argument("--data", required=True)
    ap.add_argument("--output", default=None, help="save predictions to this .npy")
    ap.add_argument("--batch_size", type=int, default=262_144)
    args = ap.parse_args(argv)
    with np.load(args.data) as z:
        branch = z["test_branch_input"] if "test_branch_input" in z.files else z["test_input"]
        trunk = z["test_trunk_input"] if "test_trunk_input" in z.files else None
        truth = z["test_output"] if "test_output" in z.files else None
    model, cfg = load_model(args.ckpt, branch_in=branch.shape[1], trunk_in=None if trunk is None else trunk.shape[1])
    pred = predict(model, branch, trunk, cfg, batch_size=args.batch_size)
    print(f"Output: {pred.shape}")
    if truth is not None:
        print(json.dumps(evaluate(pred, np.asarray(truth).reshape(len(truth), -1)[:, :1])))
    if args.output:
        np.save(args.output, pred)
    return 0


if __name__ == "__main__":
    import sys
    sys.exit(main())
