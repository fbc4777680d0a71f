"""Host logic of the view store and the ScanNet scene reader on CPU (the engine's view_* kernels are replaced by the
numpy oracle through tests/fake_engine.py): file discovery / ordering, level filtering, tuple layout, sampler order —
checked against the 13-tuples the real reference dataset class produced (tests/golden/view_prep.npz)."""
import numpy as np
import pytest
import torch

import fake_engine
import view_scene_util as vsu


@pytest.fixture(scope="module")
def gold():
    return np.load(vsu.GOLD)


def test_scene_reader_discovers_what_the_reference_does(gold, tmp_path):
    from stylemesh_b200.data.scannet_scene import ScanNetScene
    root = vsu.write_scene(gold, tmp_path)
    sc = ScanNetScene(f"{root}/train/images/{vsu.SCENE}", pyramid_levels=3, min_pyramid_height=32)
    assert len(sc) == 3 and not sc.rendered_depth
    assert sc.levels == gold["levels"].tolist() and sc.all_levels == gold["all_levels"].tolist()
    raw, size_wh = sc.load_raw(1, 30)
    assert size_wh == (40, 30) and raw.rgb.shape == (30, 40, 3) and raw.depth.dtype == np.uint16
    assert [u.shape for u in raw.uv_pyramid] == [(32, 42, 3), (48, 64, 3), (64, 85, 3)]
    assert np.array_equal(raw.depth, gold["depth_mm_1"]) and raw.index == 1
    # incomplete scenes are rejected loudly (the reference skips them silently)
    import os
    os.remove(f"{root}/train/images/{vsu.SCENE}/uv/2.angle.npy")
    with pytest.raises(ValueError, match="incompletely"):
        ScanNetScene(f"{root}/train/images/{vsu.SCENE}", pyramid_levels=3, min_pyramid_height=32)


def test_view_store_tuples_match_the_reference(gold, tmp_path, monkeypatch):
    fake_engine.install(monkeypatch)
    from stylemesh_b200.data.scannet_scene import ScanNetScene, load_scene_into_store
    root = vsu.write_scene(gold, tmp_path)
    sc = ScanNetScene(f"{root}/train/images/{vsu.SCENE}", pyramid_levels=3, min_pyramid_height=32)
    store = load_scene_into_store(sc, "cpu", 30, min_pyramid_depth=1.0)
    assert len(store) == 3
    for i in range(3):
        vsu.check_view_against_golden(store[i], gold, i)
    b = store[0]
    assert b[0].shape == (1, 3, 30, 40) and b[3].shape == (1, 1, 30, 40) and b[10].shape == (1, 30, 40)
    assert b[9][2].shape == (1, 64, 85, 2) and b[8].shape == (1,)
    order = [int(v[8][0]) for v in store.batches([2, 0], index_repeat=3)]      # RepeatingSampler order
    assert order == [2, 2, 2, 0, 0, 0]


def test_resample_tables_equal_the_oracle():
    """The product's own table builders (it must not import the oracle) against the pinned oracle."""
    from oracle import view_prep_oracle as vo
    from stylemesh_b200.data import resample as rs
    for src, dst in [(48, 30), (64, 40), (60, 32), (80, 42), (33, 64), (47, 85), (7, 7), (5, 11), (1, 4)]:
        o1, a1 = rs.cv2_linear_table(src, dst)
        o2, a2 = vo.cv2_linear_table(src, dst)
        assert np.array_equal(o1, o2) and np.array_equal(a1, a2)
        assert np.array_equal(rs.cv2_nearest_table(src, dst), vo.cv2_nearest_table(src, dst))
        assert np.array_equal(rs.pil_nearest_table(src, dst), vo.pil_nearest_table(src, dst))


def test_datamodule_split_and_sampler(gold, tmp_path, monkeypatch):
    fake_engine.install(monkeypatch)
    from stylemesh_b200.data.scannet_scene import ScanNetViewStoreDataModule
    from stylemesh_b200.model.optimize import build_parser
    root = vsu.write_scene(gold, tmp_path)
    args = build_parser().parse_args(["--style_image_path", "synthetic:64:48", "--root_path", root, "--scene", vsu.SCENE,
                                      "--resize_size", "30", "--pyramid_levels", "3", "--min_pyramid_height", "32",
                                      "--min_pyramid_depth", "1.0", "--min_images", "1", "--max_images", "10",
                                      "--train_split", "0.7", "--index_repeat", "2", "--sampler_mode", "repeat"])
    dm = ScanNetViewStoreDataModule(args, device=torch.device("cpu"))
    dm.setup()
    assert dm.train_indices == [0, 1] and dm.val_indices == [2]
    assert [int(b[8][0]) for b in dm.train_dataloader()] == [0, 0, 1, 1]
    assert [int(b[8][0]) for b in dm.val_dataloader()] == [2]
    args.sampler_mode = "random"                     # SubsetRandomSampler: a permutation of the train indices per epoch
    torch.manual_seed(0)
    loader = dm.train_dataloader()
    epochs = [sorted(int(b[8][0]) for b in loader) for _ in range(3)]
    assert len(loader) == 2 and all(e == [0, 1] for e in epochs)


def test_matterport_region_matches_the_reference(tmp_path, monkeypatch):
    """Matterport layout (house / rendered / region_k, hash_iC_Y file names, depth in 1/4000 m, mask without the depth
    factor) against the 13-tuples of the reference's Matterport_Single_House_Dataset."""
    fake_engine.install(monkeypatch)
    from stylemesh_b200.data.matterport_scene import MatterportRegion
    from stylemesh_b200.data.scene_base import load_scene_into_store
    gold = np.load(vsu.MP_GOLD)
    root = vsu.write_matterport(gold, tmp_path)
    reg = MatterportRegion(f"{root}/v1/scans/{vsu.MP_HOUSE}", region_index=0, pyramid_levels=3, min_pyramid_height=32)
    assert len(reg) == 3 and not reg.rendered_depth and not reg.mask_uses_depth and reg.depth_divisor == 4000.0
    assert reg.levels == gold["levels"].tolist() and reg.all_levels == gold["all_levels"].tolist()
    assert [p.split("/")[-1].split(".")[0] for p in reg.colors] == [str(n) for n in gold["names"]]
    store = load_scene_into_store(reg, "cpu", 30, min_pyramid_depth=1.0)
    for i in range(3):
        vsu.check_view_against_golden(store[i], gold, i)


def test_trainer_gives_every_rank_the_same_number_of_steps():
    """View sharding (rank r owns batches r, r+N, ...): an incomplete last group is dropped so that the collective
    gradient exchange never waits for a rank that has run out of views."""
    from stylemesh_b200.lightning_shim import Trainer
    loader = list(range(11))
    counts = []
    for rank in range(4):
        t = Trainer.__new__(Trainer)
        t.world_size, t.rank, t.limit_train_batches = 4, rank, -1
        mine = [i for i, _ in t._my_batches(loader)]
        counts.append(len(mine))
        assert mine == [rank, rank + 4]
    assert counts == [2, 2, 2, 2]
    t = Trainer.__new__(Trainer)
    t.world_size, t.rank, t.limit_train_batches = 1, 0, 5
    assert [i for i, _ in t._my_batches(loader)] == [0, 1, 2, 3, 4]


def test_rendered_depth_scene_matches_the_reference(tmp_path, monkeypatch):
    """Empty depth/ folder -> the renderer's float32 depth maps (scannet_dataset.py:117-144, 303-304); numpy keeps
    float32 arithmetic for uv_height there, and so do the oracle and the kernel."""
    fake_engine.install(monkeypatch)
    from stylemesh_b200.data.scannet_scene import ScanNetScene, load_scene_into_store
    gold = np.load(vsu.RD_GOLD)
    root = vsu.write_rendered_depth_scene(gold, tmp_path)
    sc = ScanNetScene(f"{root}/train/images/{vsu.RD_SCENE}", pyramid_levels=2, min_pyramid_height=32)
    assert sc.rendered_depth and len(sc) == 2 and sc.levels == [32.0, 48.0]
    raw, _ = sc.load_raw(0, 30)
    assert raw.depth.dtype == np.float32 and raw.depth_divisor == 1.0
    store = load_scene_into_store(sc, "cpu", 30, min_pyramid_depth=1.0)
    for i in range(2):
        vsu.check_view_against_golden(store[i], gold, i)
