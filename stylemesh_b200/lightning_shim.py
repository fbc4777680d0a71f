"""Minimal stand-ins for the pytorch_lightning names the reference imports (model/model.py:1, model/optimize.py:10,
data/abstract_dataset.py:18).  pytorch_lightning is not installable offline; only what the texture-optimisation
flow touches is provided: LightningModule hooks, LightningDataModule, and a Trainer with
add_argparse_args / from_argparse_args / fit / logger.{save_dir,version,experiment}.

The Trainer drives one process per GPU (RANK / LOCAL_RANK / WORLD_SIZE from the environment, NCCL backend): views
are sharded round-robin over ranks and the model's fused optimizer all-reduces the texture gradient once per step.
"""
from __future__ import annotations

import json
import os
from argparse import ArgumentParser
from typing import Any, Optional

import torch
import torch.nn as nn


class _ScalarLog:
    """`logger.experiment`: accepts the SummaryWriter calls the reference makes and keeps scalars lazily (device
    tensors are not synchronised until flush())."""

    def __init__(self, path: Optional[str] = None):
        self.path = path
        self._pending = []

    def add_scalar(self, tag, value, step=None):
        self._pending.append((tag, value, step))
        if len(self._pending) >= 4096:
            self.flush()

    def add_scalars(self, tag, values, step=None):
        for k, v in values.items():
            self.add_scalar(f"{tag}/{k}", v, step)

    def add_image(self, *a, **k):
        pass

    def flush(self):
        if not self.path:
            self._pending.clear()
            return
        os.makedirs(os.path.dirname(self.path), exist_ok=True)
        with open(self.path, "a") as fh:
            for tag, value, step in self._pending:
                v = float(value.detach().reshape(-1)[0]) if isinstance(value, torch.Tensor) else float(value)
                fh.write(json.dumps({"tag": tag, "value": v, "step": step}) + "\n")
        self._pending.clear()


class _Logger:
    def __init__(self, save_dir: str = ".", version: Optional[int] = None):
        self.save_dir = save_dir
        root = os.path.join(save_dir, "lightning_logs")
        if version is None:
            version = 0
            if os.path.isdir(root):
                olds = [int(d.split("_")[1]) for d in os.listdir(root) if d.startswith("version_") and
                        d.split("_")[1].isdigit()]
                version = max(olds) + 1 if olds else 0
        self.version = version
        self.log_dir = os.path.join(root, f"version_{version}")
        self.experiment = _ScalarLog(os.path.join(self.log_dir, "scalars.jsonl"))


class LightningModule(nn.Module):
    """nn.Module plus the attributes/hooks the reference's module uses."""

    def __init__(self):
        super().__init__()
        self.current_epoch = 0
        self.logger = _Logger.__new__(_Logger)
        self.logger.save_dir, self.logger.version, self.logger.log_dir = ".", 0, "."
        self.logger.experiment = _ScalarLog(None)
        self.trainer = None
        self.hparams = {}

    def save_hyperparameters(self, *args, **kwargs):
        import inspect
        frame = inspect.currentframe().f_back
        try:
            local = frame.f_locals
            self.hparams = {k: v for k, v in local.items() if k not in ("self", "__class__") and not k.startswith("_")}
        finally:
            del frame

    # hooks (no-ops by default)
    def on_train_epoch_start(self): ...
    def on_train_epoch_end(self): ...
    def on_validation_epoch_start(self): ...
    def on_validation_epoch_end(self): ...
    def on_epoch_end(self): ...


class LightningDataModule:
    def prepare_data(self): ...
    def setup(self, stage=None): ...
    def train_dataloader(self): raise NotImplementedError
    def val_dataloader(self): return None


def _to_device(obj: Any, device, _top: bool = True):
    if isinstance(obj, torch.Tensor):
        return obj.to(device, non_blocking=True)
    if isinstance(obj, (list, tuple)):
        moved = [_to_device(o, device, False) for o in obj]
        if _top and len(obj) == 13 and isinstance(obj[8], torch.Tensor):
            moved[8] = obj[8]          # the dataset index stays on the host: it keys the per-view plan cache
        return type(obj)(moved)
    return obj


class Trainer:
    _FLAGS = [("--gpus", int, 1), ("--max_epochs", int, 1), ("--default_root_dir", str, "."),
              ("--limit_train_batches", int, -1), ("--limit_val_batches", int, -1), ("--num_sanity_val_steps", int, 0),
              ("--log_every_n_steps", int, 50), ("--resume_from_checkpoint", str, None), ("--profiler", str, None)]

    def __init__(self, gpus=1, max_epochs=1, default_root_dir=".", limit_train_batches=-1, limit_val_batches=-1,
                 **_ignored):
        self.gpus, self.max_epochs = gpus, max_epochs
        self.limit_train_batches, self.limit_val_batches = limit_train_batches, limit_val_batches
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.logger = _Logger(default_root_dir)
        self.global_step = 0

    @classmethod
    def add_argparse_args(cls, parser: ArgumentParser) -> ArgumentParser:
        for flag, typ, default in cls._FLAGS:
            parser.add_argument(flag, type=typ, default=default)
        return parser

    @classmethod
    def from_argparse_args(cls, args, **kwargs):
        known = {f.lstrip("-"): getattr(args, f.lstrip("-"), d) for f, _, d in cls._FLAGS}
        known.update(kwargs)
        return cls(**known)

    def _setup_distributed(self):
        import torch.distributed as dist
        if self.world_size > 1 and not dist.is_initialized():
            dist.init_process_group(backend="nccl" if torch.cuda.is_available() else "gloo")

    def _my_batches(self, loader):
        """(batch_idx, batch) of this rank: view sharding, rank r owns views r, r+N, ...  Every rank must take the same
        number of optimiser steps (the gradient exchange is collective), so an incomplete last group of fewer than
        world_size batches is dropped (Lightning's DistributedSampler pads it by repeating samples instead)."""
        limit = None
        if self.world_size > 1 and hasattr(loader, "__len__"):
            n = len(loader)
            if 0 <= self.limit_train_batches < n:
                n = self.limit_train_batches
            limit = (n // self.world_size) * self.world_size
        for batch_idx, batch in enumerate(loader):
            if 0 <= self.limit_train_batches <= batch_idx or (limit is not None and batch_idx >= limit):
                break
            if batch_idx % self.world_size != self.rank:
                continue
            yield batch_idx, batch

    def fit(self, model: LightningModule, datamodule: LightningDataModule):
        if not torch.cuda.is_available():
            raise RuntimeError("stylemesh_b200 Trainer needs a CUDA device (no CPU path)")
        torch.cuda.set_device(self.local_rank)
        device = torch.device("cuda", self.local_rank)
        self._setup_distributed()
        model.to(device)
        model.trainer, model.logger = self, self.logger
        if hasattr(model, "cache_view_plans"):
            model.cache_view_plans = True      # views repeat (RepeatingSampler): build each view's mask plan once
        (optimizer,), schedulers = model.configure_optimizers()
        train_loader = datamodule.train_dataloader()
        val_loader = datamodule.val_dataloader()
        from .staging import BatchStager
        stager = BatchStager(device)
        for epoch in range(self.max_epochs):
            model.current_epoch = epoch
            model.on_train_epoch_start()
            model.train()
            # the H2D copy of the NEXT view is issued on a copy stream before this view's kernels are launched
            # (stylemesh_b200/staging.py); Lightning's own loop copies on the compute stream right before the step
            mine = self._my_batches(train_loader)
            nxt = next(mine, None)
            ticket = stager.stage(nxt[1]) if nxt is not None else None
            while nxt is not None:
                batch_idx, cur_ticket = nxt[0], ticket
                nxt = next(mine, None)
                ticket = stager.stage(nxt[1]) if nxt is not None else None
                batch = stager.acquire(cur_ticket)
                optimizer.zero_grad()
                out = model.training_step(batch, batch_idx)
                out["loss"].backward()
                optimizer.step()
                stager.release(cur_ticket)
                self.global_step += 1
            model.on_train_epoch_end()
            if val_loader is not None:
                model.on_validation_epoch_start()
                model.eval()
                with torch.no_grad():
                    for batch_idx, batch in enumerate(val_loader):
                        if 0 <= self.limit_val_batches <= batch_idx:
                            break
                        model.validation_step(_to_device(batch, device), batch_idx)
                model.on_validation_epoch_end()
            model.on_epoch_end()
            for s in schedulers:
                s.step()
            self.logger.experiment.flush()
        return model
