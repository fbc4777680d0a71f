"""Per-CTA timeline of ONE stream-K conv launch (the last igemm launch of forward(.., last=LAYER)); see the TR_*
slots in csrc/tc_igemm_v2.cu.  Not a test.  PROBE_LAYER=5 (conv3_2) by default."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylemesh_b200 import engine as eng, synthetic as syn, _abi

SLOTS = ["gt_in", "gt_out", "clk_in", "clk_prologue", "clk_tma_end", "clk_mma_first", "clk_mma_end", "clk_epi_first",
         "clk_epi_end", "w_flags", "w_tmem_full", "w_full", "w_tmem_empty", "w_empty", "clk_out", "smid"]
layer = int(os.environ.get("PROBE_LAYER", "5"))
H, W = [int(x) for x in os.environ.get("PROBE_HW", "480x640").split("x")]
e = eng.VGGEngine(syn.make_vgg_state_dict(0))
img = (torch.rand(3, H, W) * 255 - 120).cuda()
slot = e.begin(H, W)
for _ in range(3):
    e.forward(slot, img, layer)
buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")      # uint64 slots read back as int64
lib = _abi.load()
torch.cuda.synchronize()
lib.smb_debug_set_igemm_trace(_abi.ptr(buf))
e.forward(slot, img, layer)
torch.cuda.synchronize()
lib.smb_debug_set_igemm_trace(None)
t = buf.cpu().view(148, 16)
t = t[t[:, 0] != 0]
gt0 = int(t[:, 0].min())
rows = []
for r in t.tolist():
    d = dict(zip(SLOTS, r))
    c0 = d["clk_in"]
    rows.append({"smid": d["smid"], "start_ns": d["gt_in"] - gt0, "end_ns": d["gt_out"] - gt0,
                 "cycles": d["clk_out"] - c0, "prologue": d["clk_prologue"] - c0, "mma_first": d["clk_mma_first"] - c0,
                 "tma_end": d["clk_tma_end"] - c0, "mma_end": d["clk_mma_end"] - c0, "epi_first": d["clk_epi_first"] - c0,
                 "epi_end": d["clk_epi_end"] - c0, "w_flags": d["w_flags"], "w_tmem_full": d["w_tmem_full"],
                 "w_full": d["w_full"] & 0xffffffff, "w_full_halo": (d["w_full"] >> 32) & 0xffffffff,
                 "w_tmem_empty": d["w_tmem_empty"], "w_empty": d["w_empty"]})
import statistics as st
summ = {k: {"min": min(r[k] for r in rows), "med": st.median(r[k] for r in rows), "max": max(r[k] for r in rows)}
        for k in rows[0] if k != "smid"}
print(json.dumps({"layer": layer, "hw": [H, W], "impl": os.environ.get("SMB_CONV_IMPL", "tc"),
                  "dbg": os.environ.get("SMB_IGEMM_DEBUG", "0"), "ctas": len(rows), "summary": summ, "first_ctas": rows[:6]}))
