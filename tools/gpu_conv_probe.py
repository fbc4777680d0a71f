"""Timing probe of the VGG forward (igemm launches only) on a 480x640 image; used with SMB_IGEMM_DEBUG=0/1/2 and
SMB_IGEMM_MAX_BN to find out what bounds the conv kernel.  Not a test."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylemesh_b200 import engine as eng, synthetic as syn

e = eng.VGGEngine(syn.make_vgg_state_dict(0))
H, W = [int(x) for x in os.environ.get("PROBE_HW", "480x640").split("x")]
img = (torch.rand(3, H, W) * 255 - 120).cuda()
slot = e.begin(H, W)
for _ in range(3):
    e.forward(slot, img, 12)
e.set_timing(True)
for _ in range(5):
    e.forward(slot, img, 12)
t = e.read_timing()
per = {k: round(v["ms"] / 5, 4) for k, v in t.items() if v["launches"]}
print(json.dumps({"hw": [H, W], "per_class_ms": per, "dbg": os.environ.get("SMB_IGEMM_DEBUG", "0"), "max_bn": os.environ.get("SMB_IGEMM_MAX_BN", "128"),
                  "igemm_fwd_ms": t["igemm_conv_fwd"]["ms"] / 5, "tflops_alg": t["igemm_conv_fwd"]["flops"] / 5 / (t["igemm_conv_fwd"]["ms"] / 5 * 1e-3) / 1e12}))
