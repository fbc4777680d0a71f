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


def next_log_version(save_dir: str) -> int:
    root = os.path.join(save_dir, "lightning_logs")
    if not os.path.isdir(root):
        return 0
    olds = [int(d.split("_")[1]) for d in os.listdir(root) if d.startswith("version_") and d.split("_")[1].isdigit()]
    return max(olds) + 1 if olds else 0


class _Logger:
    """lightning_logs/version_N like the TensorBoardLogger the reference gets from Lightning (optimize.py:31).  Only
    rank 0 writes; the other ranks share its version number and keep a sink."""

    def __init__(self, save_dir: str = ".", version: Optional[int] = None, writer: bool = True):
        self.save_dir = save_dir
        if version is None:
            version = next_log_version(save_dir)
        self.version = version
        self.log_dir = os.path.join(save_dir, "lightning_logs", f"version_{version}")
        self.experiment = _ScalarLog(os.path.join(self.log_dir, "scalars.jsonl") if writer else None)


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
                 resume_from_checkpoint=None, **_ignored):
        self.gpus, self.max_epochs = gpus, max_epochs
        self.limit_train_batches, self.limit_val_batches = limit_train_batches, limit_val_batches
        self.resume_from_checkpoint = resume_from_checkpoint
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        if isinstance(gpus, int) and gpus > 1 and self.world_size == 1:
            import warnings
            warnings.warn(f"--gpus {gpus} but WORLD_SIZE is 1: this Trainer runs one process per GPU - launch it with "
                          f"`python -m torch.distributed.run --nproc-per-node {gpus} -m model.optimize ...`; "
                          f"continuing on ONE GPU")
        if resume_from_checkpoint and not os.path.isfile(resume_from_checkpoint):
            raise FileNotFoundError(f"--resume_from_checkpoint {resume_from_checkpoint}: no such file")
        # rank 0 picks the log version; the others follow it (each picking its own would race on the directory list)
        version = None
        if self.world_size > 1:
            self._setup_distributed()
            import torch.distributed as dist
            box = [next_log_version(default_root_dir) if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            version = box[0]
        self.logger = _Logger(default_root_dir, version, writer=self.rank == 0)
        self.global_step = 0
        self.start_epoch = 0

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
            if torch.cuda.is_available():
                torch.cuda.set_device(self.local_rank)
                dist.init_process_group(backend="nccl", device_id=torch.device("cuda", self.local_rank))
            else:
                dist.init_process_group(backend="gloo")

    # ---- checkpoints (Lightning's default ModelCheckpoint: lightning_logs/version_N/checkpoints/epoch=E-step=S.ckpt
    #      every epoch, latest only; the reference relies on it implicitly, model.py:69-72 + optimize.py:30,241) -------
    def checkpoint_dir(self) -> str:
        return os.path.join(self.logger.log_dir, "checkpoints")

    def save_checkpoint(self, model, optimizer, schedulers, epoch: int) -> Optional[str]:
        """Collective (the sharded Adam moments are gathered); rank 0 writes.  Keys follow Lightning's .ckpt layout:
        `state_dict` (the texture layers - the frozen VGG is reloaded from vgg_gatys_model_path), `optimizer_states`
        (torch.optim.Adam layout), `lr_schedulers`, `epoch` (the next epoch to run), `global_step`."""
        opt_state = optimizer.state_dict()
        if self.rank != 0:
            return None
        tex = {k: v.detach().cpu().clone() for k, v in model.state_dict().items() if k.startswith("texture.")}
        ckpt = {"epoch": epoch + 1, "global_step": self.global_step, "state_dict": tex,
                "optimizer_states": [opt_state], "lr_schedulers": [s.state_dict() for s in schedulers],
                "hyper_parameters": {k: v for k, v in getattr(model, "hparams", {}).items()
                                     if isinstance(v, (int, float, str, bool, list, dict, type(None)))},
                "gram_cache": {k: [g.detach().cpu() for g in v] for k, v in
                               getattr(getattr(model, "vgg_loss", None), "gram_cache", {}).items()},
                "stylemesh_b200": {"world_size": self.world_size}}
        d = self.checkpoint_dir()
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, f"epoch={epoch}-step={self.global_step}.ckpt")
        tmp = path + ".tmp"
        torch.save(ckpt, tmp)
        os.replace(tmp, path)
        for old in os.listdir(d):
            if old.endswith(".ckpt") and os.path.join(d, old) != path:
                os.remove(os.path.join(d, old))
        return path

    def load_checkpoint(self, path: str, model, optimizer, schedulers) -> None:
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        own = model.state_dict()
        missing = [k for k in own if k.startswith("texture.") and k not in ckpt["state_dict"]]
        if missing:
            raise KeyError(f"{path}: texture entries missing from the checkpoint: {missing}")
        with torch.no_grad():
            for k, v in ckpt["state_dict"].items():
                if k in own and k.startswith("texture."):
                    if own[k].shape != v.shape:
                        raise ValueError(f"{path}: {k} is {tuple(v.shape)}, the model has {tuple(own[k].shape)}")
                    own[k].copy_(v.to(own[k].device))
        if ckpt.get("optimizer_states"):
            optimizer.load_state_dict(ckpt["optimizer_states"][0])
        for s, sd in zip(schedulers, ckpt.get("lr_schedulers", [])):
            s.load_state_dict(sd)
        for g in optimizer.param_groups:                   # StepLR keeps the decayed lr in its own state
            if schedulers and getattr(schedulers[0], "_last_lr", None):
                g["lr"] = schedulers[0]._last_lr[0]
        cache = ckpt.get("gram_cache") or {}
        if cache and hasattr(getattr(model, "vgg_loss", None), "gram_cache"):
            dev = next(model.parameters()).device
            model.vgg_loss.gram_cache = {k: [g.to(dev) for g in v] for k, v in cache.items()}
        self.start_epoch = int(ckpt.get("epoch", 0))
        self.global_step = int(ckpt.get("global_step", 0))

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
        dm_args = getattr(datamodule, "args", None)
        if getattr(dm_args, "batch_size", 1) != 1:
            raise ValueError("batch_size must be 1: one view per step and rank (the reference's masked_features breaks "
                             "for batch_size > 1, content_and_style_losses.py:137); use more GPUs for more views per step")
        if hasattr(getattr(model, "vgg_loss", None), "cache_content_targets") and \
                getattr(dm_args, "sampler_mode", None) == "repeat" and getattr(dm_args, "index_repeat", 1) > 1:
            # VGG(target) is constant per view (cs:294) and the repeat sampler shows every view index_repeat times in a
            # row (abstract_dataset.py:498-512): compute it on the first visit only
            model.vgg_loss.cache_content_targets = True
        if self.rank != 0 and hasattr(model, "save_texture"):
            model.save_texture = False         # rank 0 exports the (replicated) texture
        (optimizer,), schedulers = model.configure_optimizers()
        self.optimizers, self.lr_schedulers = [optimizer], list(schedulers)
        if self.resume_from_checkpoint:
            self.load_checkpoint(self.resume_from_checkpoint, model, optimizer, schedulers)
        train_loader = datamodule.train_dataloader()
        val_loader = datamodule.val_dataloader()
        from .staging import BatchStager
        stager = BatchStager(device)
        for epoch in range(self.start_epoch, self.max_epochs):
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
            if val_loader is not None and self.rank == 0:      # replicas are identical: rank 0 validates and logs
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
            self.save_checkpoint(model, optimizer, schedulers, epoch)
            self.logger.experiment.flush()
            if self.world_size > 1:
                # nobody enters the next epoch's first exchange (a device-side spin-wait with a watchdog) while rank 0
                # still validates / writes files
                torch.cuda.synchronize()
                torch.distributed.barrier()
        return model
