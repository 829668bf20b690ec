"""Training / evaluation loop for QuanONetPT / HEAQNNPT with the reference solver's config keys and
per-epoch behaviour (``solvers/solver_pt.py``: model creation :86-125, optimiser/scheduler :149-189,
``train`` :191-274, ``evaluate`` :279-329), restructured for the B200 path:

* the dataset is copied to the GPU once and every epoch's permutation is drawn on the device — no
  per-batch ``torch.tensor(arr, device=…)`` H2D copy (reference ``:130-138,227-228``);
* one fused kernel pass per batch through ``DataParallelTrainer`` — no ``.item()`` syncs per batch
  (reference ``:238-241``); losses are accumulated on the device and read once per epoch;
* under ``torch.distributed`` every rank trains on its shard of each batch (same global batch order on
  all ranks; gradients all-reduced), so N GPUs run the reference's batch size N times faster.

Data generation, experiment directories and TensorBoard logging are outside this package: ``data_dict`` uses the
keys of the reference's ``DataManager.get_data`` (``data_utils/data_manager.py:74-106``) and checkpoints are
written next to ``config['output_dir']`` as ``best_model.pt`` + ``best_model.npz`` exactly like ``:250-257``.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist

from ..core.models_pt import HEAQNNPT, QuanONetPT
from ..train import DataParallelTrainer


class B200Solver:
    def __init__(self, config: dict, data_dict: Dict[str, np.ndarray], device: Optional[str] = None):
        self.config = config
        self.model_type = config.get("model_type", "QuanONet")
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank() if self.distributed else 0
        self.world = dist.get_world_size() if self.distributed else 1
        self.device = torch.device(device) if device else torch.device("cuda" if torch.cuda.is_available() else "cpu")
        t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).to(self.device)
        d = data_dict
        if self.model_type == "HEAQNN":
            self.train_in = (t(d["train_input"]),)
            self.test_in = (t(d["test_input"]),)
        else:
            self.train_in = (t(d["train_branch_input"]), t(d["train_trunk_input"]))
            self.test_in = (t(d["test_branch_input"]), t(d["test_trunk_input"]))
        self.train_out = t(d["train_output"]).reshape(len(d["train_output"]), -1)[:, :1]
        self.test_out = t(d["test_output"]).reshape(len(d["test_output"]), -1)[:, :1]
        self.model = self._create_model().to(self.device)
        opt_name = str(config.get("optimizer", "adam")).lower()
        if opt_name == "lbfgs":
            raise NotImplementedError("LBFGS requires a closure; not supported by the current training loop.")
        # single-process CUDA runs replay each full-size batch as ONE CUDA graph (static batch buffers): at the
        # reference's default batch_size = 100 a step is launch-latency bound (~0.6 ms eager, ~0.33 ms replayed)
        self.use_graph = (bool(config.get("cuda_graph", True)) and self.device.type == "cuda"
                          and opt_name in ("adam", "adamw", "sgd"))
        opt_kw = dict(config.get("optimizer_kwargs", {}))
        lr = config["learning_rate"]
        self._lr_in_graph = False
        if self.use_graph and opt_name in ("adam", "adamw"):
            opt_kw.setdefault("capturable", True)
            # a TENSOR learning rate is read at replay time: a scheduler then never forces a re-capture
            lr = torch.tensor(float(lr), dtype=torch.float32, device=self.device)
            self._lr_in_graph = True
        self.trainer = DataParallelTrainer(self.model, lr=lr, optimizer=opt_name, optimizer_kwargs=opt_kw)
        if self.world > 1:
            # replaying the step under torch.distributed needs an exchange that is itself graph-replayable and a batch
            # that splits evenly: the library's peer-memory exchange (csrc/qon_peer.cuh keeps its epoch on the device)
            from ..comm import PeerAllReduce
            self.use_graph = self.use_graph and isinstance(self.trainer._all_reduce, PeerAllReduce)
        self.scheduler = self._build_scheduler()
        self.best_loss = float("inf")
        self.best_model_path = None

    # reference solvers/solver_pt.py:86-125
    def _create_model(self):
        c = self.config
        kw = dict(num_qubits=int(c["num_qubits"]), net_size=tuple(c.get("net_size", [20, 2, 10, 2])),
                  scale_coeff=float(c.get("scale_coeff", 0.01)),
                  if_trainable_freq=str(c.get("if_trainable_freq", "true")).lower() == "true",
                  ham_bound=tuple(c.get("ham_bound", [-5, 5])), ham_diag=c.get("ham_diag", None))
        if c.get("ham_pauli") and c.get("ham_diag") is None:
            kw["ham_pauli"] = c["ham_pauli"]          # honoured here; the reference's PTSolver drops it (:88-93)
        if self.model_type == "QuanONet":
            return QuanONetPT(branch_input_size=self.train_in[0].shape[1], trunk_input_size=self.train_in[1].shape[1],
                              **kw)
        if self.model_type == "HEAQNN":
            return HEAQNNPT(input_size=self.train_in[0].shape[1], **kw)
        raise ValueError(f"B200Solver does not support model_type='{self.model_type}'")

    # reference solvers/solver_pt.py:165-187
    def _build_scheduler(self):
        name = str(self.config.get("lr_scheduler", "none")).lower()
        kw = self.config.get("lr_scheduler_kwargs", {})
        opt = self.trainer.optimizer
        epochs = self.config.get("num_epochs", 1000)
        if name == "cosine":
            return torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=kw.get("T_max", epochs),
                                                              eta_min=kw.get("eta_min", 0.0))
        if name == "step":
            return torch.optim.lr_scheduler.StepLR(opt, step_size=kw.get("step_size", 100), gamma=kw.get("gamma", 0.5))
        if name == "exponential":
            return torch.optim.lr_scheduler.ExponentialLR(opt, gamma=kw.get("gamma", 0.99))
        return None

    def train(self):
        n = self.train_out.shape[0]
        bs = min(int(self.config.get("batch_size", 100)), n)
        epochs = int(self.config["num_epochs"])
        num_batches = max(1, int(np.ceil(n / bs)))
        out_dir = self.config.get("output_dir")
        if out_dir:
            if self.rank == 0:
                os.makedirs(out_dir, exist_ok=True)
            self.best_model_path = os.path.join(out_dir, "best_model.pt")     # every rank knows it; rank 0 writes it
        history = {"loss_train": [], "loss_test": [], "rel_l2_train": []}
        gen = torch.Generator(device=self.device)
        gen.manual_seed(int(self.config.get("seed", 0)))       # same permutation on every rank
        graph = None
        use_graph = self.use_graph and bs % self.world == 0
        sb = bs // self.world if use_graph else bs                # this rank's shard of a full batch
        if use_graph:
            static_in = tuple(torch.empty((sb,) + tuple(a.shape[1:]), dtype=a.dtype, device=self.device)
                              for a in self.train_in)
            static_y = torch.empty((sb, 1), dtype=self.train_out.dtype, device=self.device)

        def capture():
            """(Re)capture one full-batch training step on the static buffers."""
            for dst, src in zip(static_in, self.train_in):
                dst.copy_(src[:sb])
            static_y.copy_(self.train_out[:sb])
            saved = {k: v.clone() for k, v in self.model.state_dict().items()}
            opt = self.trainer.optimizer
            params = [p for g_ in opt.param_groups for p in g_["params"]]
            opt_saved = [{k: v.clone() for k, v in opt.state.get(p, {}).items() if torch.is_tensor(v)} for p in params]
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):                            # warm-up outside capture (allocator, lazy init)
                    self.trainer.step(static_in, static_y, global_batch=bs)
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss_static = self.trainer.step(static_in, static_y, global_batch=bs)
            # warm-up / capture must not change the run; restore IN PLACE (the graph holds these addresses)
            self.model.load_state_dict(saved)
            for p, old in zip(params, opt_saved):
                for k, v in opt.state.get(p, {}).items():
                    if torch.is_tensor(v):
                        if k in old:
                            v.copy_(old[k])
                        else:
                            v.zero_()
            return g, loss_static

        for epoch in range(epochs):
            self.model.train()
            if use_graph and (graph is None or (self.scheduler is not None and not self._lr_in_graph)):
                graph, loss_static = capture()                # a float lr is baked into the graph: recapture when it moves
            perm = torch.randperm(n, device=self.device, generator=gen)
            loss_sum = torch.zeros((), device=self.device)
            sse = torch.zeros((), device=self.device)
            for i in range(num_batches):
                idx = perm[i * bs:(i + 1) * bs]
                gb = idx.numel()
                if graph is not None and gb == bs:
                    ridx = idx[self.rank::self.world] if self.world > 1 else idx
                    for dst, src in zip(static_in, self.train_in):
                        torch.index_select(src, 0, ridx, out=dst)
                    torch.index_select(self.train_out, 0, ridx, out=static_y)
                    graph.replay()
                    loss = loss_static
                else:
                    idx = idx[self.rank::self.world]             # this rank's shard of the batch
                    loss = self.trainer.step(tuple(a[idx] for a in self.train_in), self.train_out[idx],
                                             global_batch=gb)
                loss_sum += loss
                sse += loss * gb
            avg = float(loss_sum) / num_batches                   # one host sync per epoch
            if self.trainer.exchange_timed_out():
                raise RuntimeError("the peer-memory all-reduce gave up waiting for a rank (~30 s); gradients of that "
                                   "step were NaN-poisoned — check that every rank is alive, or set QON_COLLECTIVE=nccl")
            rel = float(torch.sqrt(sse) / (torch.linalg.norm(self.train_out) + 1e-8))
            history["loss_train"].append(avg)
            history["rel_l2_train"].append(rel)
            if avg < self.best_loss:
                self.best_loss = avg
                if self.best_model_path and self.rank == 0 and self.config.get("if_save", True):
                    sd = self.model.state_dict()
                    torch.save(sd, self.best_model_path)
                    np.savez(self.best_model_path.replace(".pt", ".npz"), **{k: v.cpu().numpy() for k, v in sd.items()})
            if self.scheduler is not None:
                self.scheduler.step()
        # final checkpoint, as the reference writes it (solvers/solver_pt.py:262-270)
        if out_dir and self.rank == 0 and self.config.get("if_save", True):
            sd = self.model.state_dict()
            final = os.path.join(out_dir, "final_model.pt")
            torch.save(sd, final)
            np.savez(final.replace(".pt", ".npz"), **{k: v.cpu().numpy() for k, v in sd.items()})
        return history

    @torch.no_grad()
    def evaluate(self):
        # EVERY rank reloads the best weights (rank 0 wrote them): replicas stay identical, so metrics agree across
        # ranks and a later train() call is still data-parallel on identical replicas
        if self.world > 1:
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            dist.barrier()
        if self.best_model_path and os.path.exists(self.best_model_path):
            self.model.load_state_dict(torch.load(self.best_model_path, map_location=self.device))
        self.model.eval()
        pred = self.model(*self.test_in)
        diff = (pred - self.test_out).double()
        return {"rel_l2": float(torch.linalg.norm(diff) / (torch.linalg.norm(self.test_out.double()) + 1e-8)),
                "MSE": float((diff ** 2).mean()), "MAE": float(diff.abs().mean()), "Max": float(diff.abs().max())}
