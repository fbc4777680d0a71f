"""Per-layer timing probe of the VGG forward (igemm launches only) on a 480x640 image; used with
SMB_CONV_IMPL / SMB_IGEMM_DEBUG / SMB_IGEMM_MAX_BN to find out what bounds the conv kernel.  Not a test.
Layer i's time = class total of forward(.., last=i) minus forward(.., last=i-1)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylemesh_b200 import engine as eng, synthetic as syn

NAMES = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv3_4", "conv4_1", "conv4_2",
         "conv4_3", "conv4_4", "conv5_1"]
CIN = [3, 64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512]
COUT = [64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512, 512]
DIV = [1, 1, 2, 2, 4, 4, 4, 4, 8, 8, 8, 8, 16]
REPS = int(os.environ.get("PROBE_REPS", "5"))

e = eng.VGGEngine(syn.make_vgg_state_dict(0))
H, W = [int(x) for x in os.environ.get("PROBE_HW", "480x640").split("x")]
img = (torch.rand(3, H, W) * 255 - 120).cuda()
slot = e.begin(H, W)
for _ in range(3):
    e.forward(slot, img, 12)


def timed(last):
    e.set_timing(True)
    for _ in range(REPS):
        e.forward(slot, img, last)
    t = e.read_timing()
    e.set_timing(False)
    return t["igemm_conv_fwd"]["ms"] / REPS


cum = [0.0] + [timed(i) for i in range(1, 13)]
layers = {}
for i in range(1, 13):
    ms = cum[i] - cum[i - 1]
    gf = 2.0 * 9 * CIN[i] * COUT[i] * (H // DIV[i]) * (W // DIV[i]) / 1e9
    layers[NAMES[i]] = {"us": round(ms * 1e3, 1), "tflops_alg": round(gf / ms, 1) if ms > 0 else None}
print(json.dumps({"hw": [H, W], "impl": os.environ.get("SMB_CONV_IMPL", "tc"), "dbg": os.environ.get("SMB_IGEMM_DEBUG", "0"),
                  "max_bn": os.environ.get("SMB_IGEMM_MAX_BN", "128"), "igemm_fwd_ms": round(cum[12], 4), "layers": layers}))
