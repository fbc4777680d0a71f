"""TextureOptimizationStyleTransferPipeline with the reference's constructor and step API (model/model.py:16-401),
running the whole per-view step on the sm_100a kernels:

    sample (uv_sample_fwd) -> VGG forward (tcgen05 implicit-GEMM) -> masked Gram / content losses
    -> hand-written backward -> UV scatter-add with the angle/depth hooks folded in -> [NCCL all-reduce]
    -> fused clamp + regulariser + Adam.

The step does not go through torch autograd: `training_step` has already deposited the texture gradient in a
persistent flat buffer when it returns; the returned loss carries a no-op grad_fn so that a Lightning-style loop
(`loss.backward(); optimizer.step()`) keeps working unchanged.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .. import engine as _eng
from ..lightning_shim import LightningModule
from ..view_cache import ViewLRU
from .. import nvtx
from .losses.content_and_style_losses import ContentAndStyleLoss, layer_hw, loss_plan_from_counts
from .losses.rgb_transform import post
from .texture.texture import HierarchicalNeuralTexture, NeuralTexture, to_image


class _GradAlreadyDeposited(torch.autograd.Function):
    """Gives the step's total loss a grad_fn whose backward does nothing (the fused step already wrote .grad)."""

    @staticmethod
    def forward(ctx, value, anchor):
        return value.clone()

    @staticmethod
    def backward(ctx, g):
        return None, None


class FusedTextureAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8, weight_decay=0) for the texture layers (model.py:391-395)
    as ONE fused kernel over all layers that also applies the clamp of NeuralTexture.normalize(), adds the regulariser
    gradient, scales by 1/world_size after the all-reduce and zeroes the gradient buffer for the next step."""

    def __init__(self, pipeline, params, lr):
        super().__init__(params, dict(lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0))
        self._pipeline = pipeline
        self._steps = 0

    def zero_grad(self, set_to_none: bool = True):
        pass            # the Adam kernel leaves the flat gradient buffer zeroed; .grad stays a view of it

    def _full_moments(self):
        """(exp_avg, exp_avg_sq) of the whole flat buffer.  The fused multi-GPU kernel shards the moments (rank r only
        ever touches slice r of its buffers, the rest stays 0), so the full state is the sum over ranks - collective."""
        st = self._pipeline._ensure_fused_state()
        m, v = st["exp_avg"].clone(), st["exp_avg_sq"].clone()
        if st.get("peer") and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            torch.distributed.all_reduce(m)
            torch.distributed.all_reduce(v)
        return m, v

    def state_dict(self):
        """torch.optim.Adam's layout: state[i] = {step, exp_avg, exp_avg_sq} per texture layer (model.py:391-395), so
        a checkpoint written here loads into the reference's optimizer and vice versa.  Collective when sharded."""
        st = self._pipeline._ensure_fused_state()
        m, v = self._full_moments()
        mods = self._pipeline._layer_modules()
        state = {i: {"step": torch.tensor(float(self._steps)),
                     "exp_avg": m[a:b].view_as(mod.data).detach().cpu().clone(),
                     "exp_avg_sq": v[a:b].view_as(mod.data).detach().cpu().clone()}
                 for i, (mod, (a, b)) in enumerate(zip(mods, st["spans"]))}
        groups = [{k: val for k, val in g.items() if k != "params"} for g in self.param_groups]
        groups[0]["params"] = list(range(len(mods)))
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        st = self._pipeline._ensure_fused_state()
        mods = self._pipeline._layer_modules()
        if len(sd["state"]) not in (0, len(mods)):
            raise ValueError(f"optimizer state has {len(sd['state'])} entries for {len(mods)} texture layers")
        sharded = bool(st.get("peer")) and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
        steps = 0
        with torch.no_grad():
            st["exp_avg"].zero_()
            st["exp_avg_sq"].zero_()
            for i, (mod, (a, b)) in enumerate(zip(mods, st["spans"])):
                ent = sd["state"].get(i, sd["state"].get(str(i)))
                if ent is None:
                    continue
                st["exp_avg"][a:b].copy_(ent["exp_avg"].reshape(-1).to(st["exp_avg"].device, torch.float32))
                st["exp_avg_sq"][a:b].copy_(ent["exp_avg_sq"].reshape(-1).to(st["exp_avg"].device, torch.float32))
                steps = max(steps, int(float(ent["step"])))
            if sharded:            # keep only this rank's slice (same split as dist_adam_kernel: float4 granules)
                world, rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
                n4 = st["param"].numel() // 4
                per = (n4 + world - 1) // world
                lo, hi = min(rank * per, n4) * 4, min((rank + 1) * per, n4) * 4
                for buf in (st["exp_avg"], st["exp_avg_sq"]):
                    buf[:lo].zero_()
                    buf[hi:].zero_()
        self._steps = steps
        for g, new in zip(self.param_groups, sd.get("param_groups", [])):
            for k, val in new.items():
                if k != "params":
                    g[k] = val

    @torch.no_grad()
    def step(self, closure=None):
        with nvtx.range("optimizer"):
            return self._step()

    def _step(self):
        pl = self._pipeline
        st = pl._ensure_fused_state()
        self._steps += 1
        group = self.param_groups[0]
        b1, b2 = group["betas"]
        spans = st["spans"]             # all layers in one launch: the flat buffers are contiguous, padding stays 0
        world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size()
            if world > 1 and st.get("peer"):
                # the one exchange of the step fused with the optimiser: reduce-scatter of the gradient,
                # Adam on this rank's slice (sharded moments), all-gather of the new texels - one kernel over NVLink
                pr = st["peer"]["ptrs"]
                st["peer"]["epoch"] = st["peer"].get("epoch", 0) + 1      # monotonic for the life of the flag buffers
                _eng.dist_adam_step(torch.distributed.get_rank(), world, pr["grad"], pr["param"], pr["flags"],
                                    st["exp_avg"], st["exp_avg_sq"], st["param"].numel(), [a for a, _ in spans],
                                    [pl._reg_grad_coef(l) for l in range(len(spans))], group["lr"], b1, b2,
                                    group["eps"], self._steps, epoch=st["peer"]["epoch"])
                return None
            if world > 1:
                torch.distributed.all_reduce(st["grad"])          # fallback: NCCL all-reduce, then the local Adam
        _eng.adam_step_segments(st["param"], st["grad"], st["exp_avg"], st["exp_avg_sq"], [a for a, _ in spans],
                                [pl._reg_grad_coef(l) for l in range(len(spans))], group["lr"], b1, b2, group["eps"],
                                self._steps, grad_scale=1.0 / world)
        return None


class TextureOptimizationStyleTransferPipeline(LightningModule):
    states = ["train", "val"]
    loss_types = ["tex_reg", "content", "style", "total"]
    default_loss_weights = {l: 0.0 for l in loss_types}

    def __init__(self, W, H,
                 hierarchical_texture=True, hierarchical_layers=4, random_texture_init=False,
                 style_image=None,
                 style_layers=ContentAndStyleLoss.style_layers, content_layers=ContentAndStyleLoss.content_layers,
                 style_weights=ContentAndStyleLoss.style_weights, content_weights=ContentAndStyleLoss.content_weights,
                 vgg_gatys_model_path=None, use_angle_weight=True, use_depth_scaling=True,
                 style_pyramid_mode="single", gram_mode="current", angle_threshold=60,
                 log_images_nth=-1, save_texture=True, texture_dir="", texture_prefix="",
                 learning_rate=1e-3, decay_gamma=0.1, decay_step_size=30,
                 loss_weights=default_loss_weights, tex_reg_weights=None, extra_args={}):
        super().__init__()
        self.hparams = {k: v for k, v in locals().items() if k not in ("self", "__class__", "style_image")}

        # ---- texture (model.py:76-92) ----
        self.hierarchical_texture = hierarchical_texture
        self.hierarchical_layers = hierarchical_layers
        self.C = 3
        if hierarchical_texture:
            self.texture = HierarchicalNeuralTexture(W, H, self.C, hierarchical_layers, random_texture_init)
        else:
            self.texture = NeuralTexture(W, H, self.C, random_texture_init)
        self.tex_reg_weights = tex_reg_weights
        if hierarchical_texture and not tex_reg_weights:
            self.tex_reg_weights = [pow(2, hierarchical_layers - i - 1) for i in range(hierarchical_layers)]
            self.tex_reg_weights[-1] = 0
            print(f"No tex_reg_weights specified. Setting them to {self.tex_reg_weights}")
        if hierarchical_texture and hierarchical_layers != len(self.tex_reg_weights):
            raise ValueError(
                f"Have {hierarchical_layers} texture layers, but only {len(self.tex_reg_weights)} weights specified")

        # ---- losses (model.py:97-111) ----
        self.loss_history = {loss: {k: [] for k in self.states} for loss in self.loss_types}
        self.loss_weights = dict(loss_weights) if loss_weights else {}
        for loss in self.loss_history.keys():
            if loss not in self.loss_weights:
                self.loss_weights[loss] = self.default_loss_weights[loss]
                print(f"No weight specified for the '{loss}' loss. Setting it to {self.loss_weights[loss]}")
        self.vgg_gatys_model_path = vgg_gatys_model_path
        self.vgg_loss = ContentAndStyleLoss(vgg_gatys_model_path, style_layers, content_layers, style_weights,
                                            content_weights, angle_threshold=angle_threshold,
                                            style_pyramid_mode=style_pyramid_mode, gram_mode=gram_mode)

        # ---- misc (model.py:116-141) ----
        if style_image is None:
            raise ValueError("style_image is required")
        self.style_image = style_image
        self.orig_style_image = style_image.clone()
        self.angle_threshold = angle_threshold
        self.style_pyramid_mode = style_pyramid_mode
        self.gram_mode = gram_mode
        self.use_angle_weight = use_angle_weight
        self.use_depth_scaling = use_depth_scaling
        self.learning_rate = learning_rate
        self.decay_gamma = decay_gamma
        self.decay_step_size = decay_step_size
        self.log_images_nth = log_images_nth
        self.save_texture = save_texture
        self.texture_prefix = texture_prefix
        self.texture_dir = texture_dir
        self.batches_per_epoch = {k: 0 for k in self.states}
        self.train_epoch_end = False
        self.val_epoch_end = False

        # ---- B200 step state ----
        self._fused: Optional[dict] = None
        self._loss_buf: Optional[torch.Tensor] = None
        self.cache_view_plans = False            # opt-in for resident views (bench `value` leg, repeated views)
        self._plan_cache = ViewLRU(16 << 30)     # per-view masks / hooks / counts, least recently used view evicted
        self._style_ready = False

    # ------------------------------------------------------------------------------------------------------
    # texture parameter plumbing
    # ------------------------------------------------------------------------------------------------------
    def _layer_modules(self) -> List[NeuralTexture]:
        return list(self.texture.layers) if self.hierarchical_texture else [self.texture]

    def _ensure_fused_state(self) -> dict:
        """Re-home every layer Parameter (and its .grad) as a view of one flat buffer so that the gradient
        all-reduce is a single call and Adam streams contiguous memory.  Re-done if the module was moved."""
        mods = self._layer_modules()
        dev = mods[0].data.device
        if self._fused is not None and self._fused["param"].device == dev and all(
                m.data.data_ptr() == self._fused["param"][a:b].data_ptr() for m, (a, b) in
                zip(mods, self._fused["spans"])):
            return self._fused
        _eng.require_cuda_device(dev)
        sizes = [m.data.numel() for m in mods]
        spans, off = [], 0
        for n in sizes:
            spans.append((off, off + n))
            off += (n + 63) // 64 * 64                     # keep every layer 256-byte aligned
        peer = self._alloc_peer_buffers(off, dev)          # N > 1: parameters / gradient in NVLink peer memory
        flat = peer["param"] if peer else torch.zeros(off, device=dev, dtype=torch.float32)
        grad = peer["grad"] if peer else torch.zeros_like(flat)
        st = {"param": flat, "grad": grad, "exp_avg": torch.zeros_like(flat),
              "exp_avg_sq": torch.zeros_like(flat), "spans": spans, "peer": peer}
        with torch.no_grad():
            for m, (a, b) in zip(mods, spans):
                flat[a:b].copy_(m.data.detach().reshape(-1).to(torch.float32))
                m.data.data = flat[a:b].view_as(m.data)
                m.data.grad = st["grad"][a:b].view_as(m.data)
        dist = torch.distributed
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(flat, 0)                        # replicas start identical (DDP broadcasts rank 0's weights)
            if peer:
                peer["handles"][0].barrier()
        self._fused = st
        return st

    @staticmethod
    def _alloc_peer_buffers(numel: int, dev) -> Optional[dict]:
        """Flat parameter / gradient / flag buffers that every rank of the node can address (symmetric memory over
        NVLink), for the fused reduce-scatter + Adam + all-gather kernel (smb_dist_adam_step).  None: single process, a
        non-NCCL backend, SMB_DIST_ADAM=0, or symmetric memory unavailable -> plain all_reduce + local Adam."""
        import os
        dist = torch.distributed
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
            return None
        if os.environ.get("SMB_DIST_ADAM", "1") == "0" or dist.get_backend() != "nccl" or dist.get_world_size() > 16:
            return None
        bufs, err = None, None
        try:
            import torch.distributed._symmetric_memory as symm
            bufs, handles = {}, []
            for name, n, dt in (("param", numel, torch.float32), ("grad", numel, torch.float32),
                                ("flags", 64, torch.int32)):
                t = symm.empty(n, dtype=dt, device=dev)
                handles.append(symm.rendezvous(t, dist.group.WORLD))
                t.zero_()
                bufs[name] = t
            bufs["handles"] = handles
            bufs["ptrs"] = {name: [int(p) for p in h.buffer_ptrs] for name, h in zip(("param", "grad", "flags"), handles)}
        except Exception as e:                             # pragma: no cover - depends on the box
            bufs, err = None, e
        # every rank must take the SAME path: a rank in the all_reduce fallback would leave the others spinning in
        # dist_adam_kernel until its watchdog traps
        ok = torch.tensor([1 if bufs is not None else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            import warnings
            warnings.warn(f"symmetric memory unavailable on at least one rank (here: {err!r}); every rank falls back to "
                          f"NCCL all_reduce + local Adam")
            return None
        return bufs

    def _layer_tensors(self) -> List[torch.Tensor]:
        return [m.data.detach() for m in self._layer_modules()]

    def _grad_tensors(self) -> List[torch.Tensor]:
        st = self._ensure_fused_state()
        return [st["grad"][a:b].view_as(m.data) for m, (a, b) in zip(self._layer_modules(), st["spans"])]

    def _reg_weight(self, l: int) -> float:
        lam = self.loss_weights.get("tex_reg", 0.0)
        if lam <= 0 or not self.hierarchical_texture:       # model.py:264-267 and tex_reg_loss() :163-171
            return 0.0
        return float(lam * self.tex_reg_weights[l])

    def _reg_grad_coef(self, l: int) -> float:
        n = self._layer_modules()[l].data.numel()
        return 2.0 * self._reg_weight(l) / n

    # ------------------------------------------------------------------------------------------------------
    # reference-compatible pieces
    # ------------------------------------------------------------------------------------------------------
    def _ensure_style_targets(self, like: torch.Tensor):
        """model.py:149-153 — lazy style-target initialisation on the first batch."""
        if self._style_ready:
            return
        if self.style_image.dim() != 4:
            self.style_image = self.style_image.unsqueeze(0)
        self.style_image = self.style_image.to(like.device, like.dtype)
        self.vgg_loss.set_style_image(self.style_image)
        self._style_ready = True

    def forward(self, x):
        """model.py:143-161 — sampled prediction per UV pyramid level (autograd-visible module path)."""
        image, uv_map = x[0], x[9]
        self._ensure_style_targets(image)
        return [self.texture(v) for v in uv_map]

    def tex_reg_loss(self):
        if self.hierarchical_texture:
            return self.texture.regularizer(self.tex_reg_weights)
        return torch.zeros(1).type_as(self.texture.data)

    def update_batch_count(self, batch_idx, state):
        self.batches_per_epoch[state] = max(self.batches_per_epoch[state], batch_idx + 1)

    # ------------------------------------------------------------------------------------------------------
    # per-view plan: masks, hooks, factors (model.py:188-257 + cs:146-217)
    # ------------------------------------------------------------------------------------------------------
    @staticmethod
    def _erode(x):
        """model.py:204-208 — keep x where the zero-padded 3x3 box mean is exactly 1 (smb_view_erode3x3)."""
        return _eng.view_erode3x3(x.to(torch.float32).contiguous())

    def build_view_plan(self, batch) -> dict:
        """Everything about a view's masks that the step needs (model.py:188-257 + cs:146-217): per kept pyramid level
        the angle hook (bilinear), the depth-interpolation hook (nearest of the eroded level weights), the per-layer
        row masks with the angle pass / fail split and their pixel counts.  1 + L launches of the mask-pyramid kernels
        (csrc/mask_plan_kernels.cu) and ONE host read-back of the counts; cached per view by fused_view_step."""
        (rgb, _, _, _, _, rounded, other, interp_w, _, uvs, mask, angle_guidance, angle_degrees) = batch
        sizes = [(int(u.shape[1]), int(u.shape[2])) for u in uvs]
        L = len(sizes)
        Hr, Wr = int(mask.shape[-2]), int(mask.shape[-1])
        style_on = self.loss_weights.get("style", 0.0) != 0.0
        layer_names = (self.vgg_loss.style_layers if style_on else []) + self.vgg_loss.content_layers
        split = self.vgg_loss.style_pyramid_mode == "multi" and style_on
        convs = [_eng.layer_index(n) for n in layer_names]
        if self.use_depth_scaling:                                                       # model.py:210-251
            lvl_mask, lvl_weight = _eng.view_level_masks(mask.reshape(Hr, Wr), rounded.reshape(Hr, Wr),
                                                         other.reshape(Hr, Wr), interp_w.reshape(Hr, Wr).float(), L)
            todo = list(range(L))
        else:                                                                            # model.py:253-254
            lvl_mask, lvl_weight = mask.reshape(1, Hr, Wr).float().contiguous(), None
            todo = [L - 1]                               # every other level has an all-zero mask and is dropped (:256)
        guidance = angle_guidance.reshape(Hr, Wr).float().contiguous() if self.use_angle_weight else None
        degrees = angle_degrees.reshape(Hr, Wr).float().contiguous() if split else None
        stride = 1 + 3 * len(layer_names)
        counts = torch.zeros(max(len(todo), 1) * stride, device=mask.device, dtype=torch.int32)
        per_level = {}
        for j, i in enumerate(todo):
            per_level[i] = _eng.view_level_plan(
                lvl_mask[i if self.use_depth_scaling else 0], lvl_weight[i] if lvl_weight is not None else None, guidance,
                degrees, float(self.vgg_loss.angle_threshold), sizes[i], [layer_hw(c, *sizes[i]) for c in convs],
                counts[j * stride:(j + 1) * stride])
        host = counts.tolist()                                                           # the one host sync of the plan
        cnt = {i: host[j * stride:(j + 1) * stride] for j, i in enumerate(todo)}
        keep = [i for i in todo if cnt[i][0] > 0]                                        # model.py:256-257
        plan = loss_plan_from_counts([sizes[i] for i in keep], [per_level[i] for i in keep], [cnt[i] for i in keep],
                                     layer_names, split)
        hooks0 = {i: per_level[i]["hook0"] for i in keep if self.use_angle_weight}       # model.py:195-202
        dweights = [per_level[i]["hook1"] if i in per_level else None for i in range(L)]  # model.py:247-251
        return {"keep": keep, "sizes": sizes, "hook0": hooks0, "hook1": dweights, "plan": plan,
                "layer_names": layer_names}

    # ------------------------------------------------------------------------------------------------------
    # the fused step
    # ------------------------------------------------------------------------------------------------------
    def fused_view_step(self, batch, want_grads: bool = True) -> torch.Tensor:
        """Runs one view: returns loss_buf = [style*w, content*w, tex_reg*w, total] (device tensor, reused next
        call) and, if want_grads, ACCUMULATES the texture gradient into the flat gradient buffer."""
        rgb, uvs = batch[0], batch[9]
        idx = batch[8]
        self._ensure_style_targets(rgb)
        self._ensure_fused_state()
        if self._loss_buf is None or self._loss_buf.device != rgb.device:
            self._loss_buf = torch.zeros(4, device=rgb.device, dtype=torch.float32)
        buf = self._loss_buf
        buf.zero_()
        key = None
        if self.cache_view_plans:
            key = int(idx.reshape(-1)[0]) if isinstance(idx, torch.Tensor) else int(idx)
        vp = self._plan_cache.get(key) if key is not None else None
        if vp is None:
            vp = self.build_view_plan(batch)
            if key is not None:
                self._plan_cache.put(key, vp)
        keep = vp["keep"]
        layers = self._layer_tensors()
        with nvtx.range("sample"):
            preds = [_eng.uv_sample_fwd(layers, uvs[i][0]) for i in keep]
        w_style = float(self.loss_weights.get("style", 0.0))
        w_content = float(self.loss_weights.get("content", 0.0))
        loss = self.vgg_loss
        saved_style_layers = loss.style_layers
        if w_style == 0.0:
            loss.style_layers = []                       # 0 * style_loss: skip the work (and VGG beyond r42)
            loss.layers = loss.content_layers
        try:
            with nvtx.range("content_targets"):
                tgts = loss.content_targets(rgb, [vp["sizes"][i] for i in keep], cache_key=key)
            with nvtx.range("vgg_loss"):
                grads = loss.fused_loss_and_grads(preds, vp["plan"], tgts, w_style, w_content, buf,
                                                  want_grads=want_grads)
        finally:
            loss.style_layers = saved_style_layers
            loss.layers = saved_style_layers + loss.content_layers
        if want_grads:
            with nvtx.range("scatter"):
                gl = self._grad_tensors()
                for j, i in enumerate(keep):
                    _eng.uv_scatter_bwd(gl, uvs[i][0], grads[j], vp["hook0"].get(i), vp["hook1"][i])
        mods = self._layer_modules()                                                     # model.py:264-267
        reg = [self._reg_weight(l) / m.data.numel() for l, m in enumerate(mods)]
        if any(c > 0 for c in reg):
            with nvtx.range("regulariser"):
                st = self._ensure_fused_state()
                _eng.texreg_value_segments(st["param"], [a for a, _ in st["spans"]], reg, buf[2:3])
        torch.sum(buf[0:3], dim=0, keepdim=True, out=buf[3:4])                           # model.py:270
        return buf

    def forward_with_loss(self, batch, batch_idx, state):
        log_idx = batch_idx + self.current_epoch * self.batches_per_epoch[state]
        self.update_batch_count(batch_idx, state)
        buf = self.fused_view_step(batch, want_grads=(state == "train")).clone()
        named = {"style": buf[0:1], "content": buf[1:2], "tex_reg": buf[2:3], "total": buf[3:4]}
        for loss_type, value in named.items():                        # model.py:277-282, without the host syncs
            self.loss_history[loss_type][state].append(value)
            self.logger.experiment.add_scalar(f"Batch/Loss/{state}/{loss_type}", value, log_idx)
        anchor = self._layer_modules()[0].data
        total = _GradAlreadyDeposited.apply(named["total"], anchor) if state == "train" else named["total"]
        return {"loss": total}

    def training_step(self, batch, batch_idx, optimizer_idx=0):
        return self.forward_with_loss(batch, batch_idx, "train")

    def validation_step(self, batch, batch_idx):
        return self.forward_with_loss(batch, batch_idx, "val")

    # ------------------------------------------------------------------------------------------------------
    # epoch bookkeeping (model.py:329-385)
    # ------------------------------------------------------------------------------------------------------
    def reset_loss_count(self, state):
        for loss_type in self.loss_history.keys():
            self.loss_history[loss_type][state].clear()

    def compute_mean_loss(self, state):
        for loss_type, loss in self.loss_history.items():
            if isinstance(state, list):
                means = {s: torch.stack(loss[s]).mean() for s in state if loss[s]}
                if means:
                    self.logger.experiment.add_scalars(f"Loss/{'-'.join(state)}/{loss_type}", means, self.current_epoch)
            elif loss[state]:
                self.logger.experiment.add_scalar(f"Loss/{state}/{loss_type}", torch.stack(loss[state]).mean(),
                                                  self.current_epoch)

    def on_train_epoch_start(self) -> None:
        self.train_epoch_end = False
        self.val_epoch_end = False
        self.reset_loss_count("train")

    def on_validation_epoch_start(self) -> None:
        self.val_epoch_end = False
        self.reset_loss_count("val")

    def on_train_epoch_end(self) -> None:
        self.train_epoch_end = True

    def on_validation_epoch_end(self) -> None:
        self.val_epoch_end = True

    def on_epoch_end(self) -> None:
        if not self.train_epoch_end:
            return
        self.compute_mean_loss("train")
        self.compute_mean_loss("val")
        self.compute_mean_loss(["train", "val"])
        if self.save_texture and self.texture_dir:
            with torch.no_grad():
                self.texture.save_layers(self.texture_dir, f"{self.texture_prefix}epoch_{self.current_epoch}",
                                         normalize_transform=post())
                self.texture.save_image(self.texture_dir, f"{self.texture_prefix}epoch_{self.current_epoch}_",
                                        normalize_transform=post())

    def configure_optimizers(self):
        """model.py:387-401 — Adam on the texture + StepLR (stepped once per epoch by the trainer)."""
        params = [{"params": [m.data for m in self._layer_modules()], "weight_decay": 0.0,
                   "lr": self.learning_rate}]
        optimizer = FusedTextureAdam(self, params, lr=self.learning_rate)
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer=optimizer, gamma=self.decay_gamma,
                                                    step_size=self.decay_step_size)
        return [optimizer], [scheduler]


def to_tensor_image(t, idx=0):
    from torchvision.transforms import ToTensor
    if t.dim() == 4:
        return torch.stack([to_tensor_image(t[b], idx) for b in range(t.shape[0])], dim=0)
    return ToTensor()(to_image(t, idx, normalize_transform=post()))


def find_pyramid_size(pyramid, sample):
    for i, p in enumerate(pyramid):
        if p.shape[2] == sample.shape[2]:
            return i, p
    return 0, pyramid[-1][0]
