"""CLI with the flag surface of the reference's `python -m model.optimize` (model/optimize.py:238-290), so that the
scripts/train/optimize_texture_*.sh presets run unchanged on the B200 path.

In scope: flag parsing, model construction, the training loop (lightning_shim.Trainer), texture export, and
`--dataset scannet|matterport`: one scene / house region in the reference's directory layout, prepared once on the
device and kept resident (stylemesh_b200/data).  `--dataset synthetic` generates seeded views in memory;
other datasets plug in through `register_datamodule`.
`--renderer_mipmap cuda` replaces the post-run OpenGL mip-map render of the styled views (optimize.py:181-208) by a
trilinear lookup of the views' own (u, v, LOD) maps on the device (stylemesh_b200.export.MipPreview).
Out of scope (SURVEY §2 #6,#9-#11): the multi-scene dataset classes, the video and the LPIPS / reprojection evaluation.
"""
from __future__ import annotations

import os
from argparse import ArgumentParser
from os.path import join
from typing import Callable, Dict

import torch

from .. import synthetic as syn
from ..lightning_shim import LightningDataModule, Trainer
from .losses.content_and_style_losses import ContentAndStyleLoss
from .losses.rgb_transform import pre
from .model import TextureOptimizationStyleTransferPipeline

_DATAMODULES: Dict[str, Callable] = {}


def register_datamodule(name: str, factory: Callable) -> None:
    """factory(args, transforms: dict) -> LightningDataModule yielding the reference's 13-tuples."""
    _DATAMODULES[name] = factory


class SyntheticSceneDataModule(LightningDataModule):
    """Seeded in-memory stand-in for {ScanNet,Matterport}_Single_Scene_DataModule: `max_images` views, sequential
    train/val split, 'repeat' sampler (each view index_repeat times in a row, data/abstract_dataset.py:498-505)."""

    def __init__(self, args):
        self.args = args
        self.train_indices, self.val_indices = [], []
        self.selected_scene = ""

    def setup(self, stage=None):
        a = self.args
        n = a.max_images if a.max_images > 0 else 8
        n_train = max(1, int(round(n * a.train_split)))
        self.train_indices = list(range(n_train))
        self.val_indices = list(range(n_train, n))
        h = a.resize_size
        w = int(round(h * 4 / 3))
        sizes = syn.pyramid_sizes((h, w), a.pyramid_levels) if a.pyramid_levels > 1 else [(h, w)]
        self._views = {i: syn.make_view(1000 + i, (h, w), sizes) for i in range(n)}

    def _loader(self, indices, repeat):
        order = [i for i in indices for _ in range(repeat)]
        return [self._views[i].as_batch() for i in order]

    def train_dataloader(self):
        return self._loader(self.train_indices, max(1, self.args.index_repeat))

    def val_dataloader(self):
        return self._loader(self.val_indices, 1) if self.val_indices else None


def _load_style_image(path: str) -> torch.Tensor:
    """model/optimize.py:117-126 — open, cap the long side at 2048, ToTensor, pre()."""
    if path.startswith("synthetic:"):
        _, h, w = (path.split(":") + ["600", "468"])[:3]
        return syn.make_style_image(7, int(h), int(w))
    import PIL
    from PIL import Image
    from torchvision.transforms import Resize, ToTensor
    PIL.Image.MAX_IMAGE_PIXELS = 933120000
    img = Image.open(path).convert("RGB")
    if img.size[0] > 2048 or img.size[1] > 2048:
        img = Resize(2048)(img)
    return pre()(ToTensor()(img))


def main(args):
    trainer = Trainer.from_argparse_args(args)
    log_dir = join(trainer.logger.save_dir, f"lightning_logs/version_{trainer.logger.version}")
    os.makedirs(log_dir, exist_ok=True)

    if args.dataset == "synthetic":
        dm = SyntheticSceneDataModule(args)
    elif args.dataset in _DATAMODULES:
        dm = _DATAMODULES[args.dataset](args, {"rgb_pre": pre()})
    elif args.dataset == "scannet":                    # optimize.py:44-63 on the GPU-resident view store
        from ..data.scannet_scene import ScanNetViewStoreDataModule
        dm = ScanNetViewStoreDataModule(args)
    elif args.dataset == "matterport":                 # optimize.py:65-88
        from ..data.matterport_scene import MatterportViewStoreDataModule
        dm = MatterportViewStoreDataModule(args)
    else:
        raise ValueError(f"Unsupported dataset: {args.dataset} (file loaders are outside the B200 hot path; register "
                         f"one with stylemesh_b200.model.optimize.register_datamodule or use --dataset synthetic)")
    dm.prepare_data()
    dm.setup()

    if args.loss_weights:                                          # optimize.py:104-108
        args.loss_weights = {l[0]: float(l[1]) for l in args.loss_weights}
    if args.tex_reg_weights:                                       # optimize.py:110-115
        d = {int(w[0]): float(w[1]) for w in args.tex_reg_weights}
        args.tex_reg_weights = [d[i] for i in range(len(d))]

    vgg_path = args.vgg_gatys_model_path
    if vgg_path.startswith("synthetic:"):                          # seeded He-init weights (no vgg_conv.pth offline)
        vgg_path = join(log_dir, "vgg_synthetic.pth")
        torch.save(syn.make_vgg_state_dict(int(args.vgg_gatys_model_path.split(":")[1] or 0)), vgg_path)

    model = TextureOptimizationStyleTransferPipeline(
        W=args.texture_size[0], H=args.texture_size[1],
        hierarchical_texture=args.hierarchical, hierarchical_layers=args.hierarchical_layers,
        random_texture_init=args.random_texture_init,
        style_image=_load_style_image(args.style_image_path),
        style_layers=args.style_layers, content_layers=args.content_layers,
        style_weights=args.style_weights, content_weights=args.content_weights,
        vgg_gatys_model_path=vgg_path,
        use_angle_weight=not args.no_angle_weight, use_depth_scaling=not args.no_depth_scaling,
        angle_threshold=args.angle_threshold, style_pyramid_mode=args.style_pyramid_mode, gram_mode=args.gram_mode,
        learning_rate=args.learning_rate, tex_reg_weights=args.tex_reg_weights, decay_gamma=args.decay_gamma,
        decay_step_size=args.decay_step_size, loss_weights=args.loss_weights,
        extra_args={**vars(args), "indices": {"train": dm.train_indices, "val": dm.val_indices}},
        log_images_nth=args.log_images_nth, save_texture=args.save_texture, texture_dir=log_dir)

    trainer.fit(model, dm)
    if args.renderer_mipmap == "cuda" and trainer.rank == 0:       # optimize.py:167-208 without the OpenGL renderer
        render_previews(model, dm, join(log_dir, "preview"), max_views=args.preview_views,
                        lod_bias=args.preview_lod_bias)
    return model


def render_previews(model, dm, out_dir: str, max_views: int = 16, lod_bias: float = 0.0):
    """The reference's post-run step renders the scene with the final texture through its OpenGL renderer
    (GL_LINEAR_MIPMAP_LINEAR, model/optimize.py:181-208).  Headless equivalent: every view's UV map already stores
    (u, v, mip LOD) per pixel (render_uv/shader/uvmap.frag:8-13), so the styled view is a trilinear lookup into the
    texture's mip chain (stylemesh_b200.export.MipPreview) - validation views first, `max_views` at most."""
    import numpy as np
    from .. import export
    os.makedirs(out_dir, exist_ok=True)
    mips = export.MipPreview(model.texture)
    dev = mips.levels[0].device
    order = list(getattr(dm, "val_indices", [])) + list(getattr(dm, "train_indices", []))
    written = []
    for i in order[:max(0, int(max_views))]:
        scene = getattr(dm, "scene", None)
        if scene is not None:                                        # the renderer's (H, W, 3) [u, v, lod] of the finest level
            uv = torch.from_numpy(np.ascontiguousarray(np.load(scene.uv_levels[-1][i]), dtype=np.float32)).to(dev)
        else:                                                        # in-memory views: grids in [-1, 1], no LOD channel
            view = dm._views[i] if hasattr(dm, "_views") else dm.store[i]
            grid = view.uvs[-1][0] if hasattr(view, "uvs") else view[9][-1][0]
            uv = export.grid_to_uv(grid.to(dev))
        path = join(out_dir, f"{i:05d}.jpg")
        export.save_rgb8(mips.render(uv, lod_bias), path)
        written.append(path)
    return written


def build_parser() -> ArgumentParser:
    parser = ArgumentParser()
    parser = Trainer.add_argparse_args(parser)
    parser.add_argument('--root_path', default="/path/to/datasets/scannet")
    parser.add_argument('--dataset', default="scannet",
                        choices=["icl", "scannet", "vase", "3dfuture", "matterport", "synthetic"])
    parser.add_argument('--matterport_region_index', default=0, type=int)
    parser.add_argument('--train_split', default=0.8, type=float)
    parser.add_argument('--val_split', default=0.2, type=float)
    parser.add_argument('--split_mode', default="sequential", type=str)
    parser.add_argument('--scene', default="")
    parser.add_argument('--max_images', default=-1, type=int)
    parser.add_argument('--min_images', default=1000, type=int)
    parser.add_argument('--resize_size', default=256, type=int)
    parser.add_argument('--texture_size', default="512,512", type=lambda s: [int(f) for f in s.split(",")])
    parser.add_argument('--hierarchical', default=False, action="store_true")
    parser.add_argument('--hierarchical_layers', default=4, type=int)
    parser.add_argument('--random_texture_init', default=False, action="store_true")
    parser.add_argument('--batch_size', default=1, type=int)
    parser.add_argument('--learning_rate', default=1, type=float)
    parser.add_argument("--loss_weight", action='append', type=lambda kv: kv.split("="), dest='loss_weights')
    parser.add_argument("--tex_reg_weight", action='append', type=lambda kv: kv.split("="), dest='tex_reg_weights')
    parser.add_argument('--decay_gamma', default=0.1, type=float)
    parser.add_argument('--decay_step_size', default=30, type=int)
    parser.add_argument('--num_workers', default=4, type=int)
    parser.add_argument('--log_images_nth', default=-1, type=int)
    parser.add_argument('--save_texture', default=False, action="store_true")
    parser.add_argument('--shuffle', default=False, action="store_true")
    parser.add_argument('--sampler_mode', default="repeat", type=str)
    parser.add_argument('--index_repeat', default=1, type=int)
    parser.add_argument('--vgg_gatys_model_path', default="/path/to/models/vgg_conv.pth", type=str)
    parser.add_argument('--style_image_path', required=True, type=str)
    parser.add_argument('--style_layers', type=lambda s: s.split(","), default=ContentAndStyleLoss.style_layers)
    parser.add_argument('--content_layers', type=lambda s: s.split(","), default=ContentAndStyleLoss.content_layers)
    parser.add_argument('--style_weights', type=lambda s: [float(f) for f in s.split(",")],
                        default=ContentAndStyleLoss.style_weights)
    parser.add_argument('--content_weights', type=lambda s: [float(f) for f in s.split(",")],
                        default=ContentAndStyleLoss.content_weights)
    parser.add_argument('--no_angle_weight', default=False, action="store_true")
    parser.add_argument('--no_depth_scaling', default=False, action="store_true")
    parser.add_argument('--angle_threshold', default=60.0, type=float)
    parser.add_argument('--pyramid_levels', default=8, type=int)
    parser.add_argument('--min_pyramid_depth', default=0.25, type=float)
    parser.add_argument('--min_pyramid_height', default=32, type=int)
    parser.add_argument('--style_pyramid_mode', default='single', choices=ContentAndStyleLoss.style_pyramid_modes)
    parser.add_argument('--gram_mode', default='current', choices=ContentAndStyleLoss.gram_modes)
    parser.add_argument('--renderer_mipmap', default=None, type=str,
                        help="the reference passes the path of its OpenGL renderer here; 'cuda' renders the styled "
                             "views headlessly from the views' own (u, v, LOD) maps after training")
    parser.add_argument('--preview_views', default=16, type=int)
    parser.add_argument('--preview_lod_bias', default=0.0, type=float)
    return parser


if __name__ == '__main__':
    main(build_parser().parse_args())
