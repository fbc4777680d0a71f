"""Preparation throughput of the view store: upload + smb_view_* kernels per 640x480 view with the
ScanNet 4-level UV pyramid.  Not a test; prints one JSON line.  python tools/view_store_bench.py"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylemesh_b200.data import RawView, ViewStore

rng = np.random.default_rng(0)
H, W = 480, 640
sizes = [(256, 341), (432, 576), (608, 811), (784, 1045)]
raws = []
for i in range(8):
    depth = rng.integers(300, 4000, size=(480, 640), dtype=np.uint16)
    uvs = [rng.random((h, w, 3)).astype(np.float32) for h, w in sizes]
    raws.append(RawView(rgb=rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8), uv_pyramid=uvs,
                        angle=rng.random((480, 640, 3)).astype(np.float32), depth=depth, depth_divisor=1000.0))
store = ViewStore("cuda", [256, 432, 608, 784], 0.25, (W, H))
store.add(raws[0])
torch.cuda.synchronize()
t0 = time.perf_counter()
for r in raws:
    store.add(r)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / len(raws)
raw_bytes = sum(a.nbytes for a in [raws[0].rgb, raws[0].angle, raws[0].depth] + raws[0].uv_pyramid)
print(json.dumps({"prepare_ms_per_view": dt * 1e3, "views_per_s": 1.0 / dt, "raw_bytes_per_view": raw_bytes,
                  "resident_bytes_per_view": store.bytes_resident // len(store), "views": len(store)}))
