"""CPU emulation for DESIGN §7 item 7: would Winograd F(2x2, 3x3) keep the precision budget of the 3 x bf16 split?
One 256 -> 256 channel layer at 60 x 80, post-ReLU input, He-normal weights, relative L2 error against float64 of the
direct convolution and of the Winograd form with fp32 / 3 x bf16 / 1 x bf16 GEMMs (fp32 accumulate, round-to-nearest:
the tensor core truncates, §3).  python tools/winograd_precision.py
"""
import torch, math
torch.manual_seed(0)
torch.set_num_threads(8)
C, K, H, W = 256, 256, 60, 80
x = torch.relu(torch.randn(C, H, W, dtype=torch.float64))            # post-ReLU activations
w = torch.randn(K, C, 3, 3, dtype=torch.float64) * math.sqrt(2.0 / (9 * C))
ref = torch.nn.functional.conv2d(x[None], w, padding=1)[0]

def split(t32):
    hi = t32.to(torch.bfloat16).to(torch.float32)
    lo = (t32 - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo

def mm3(a32, b32):                      # a: (..., M, C), b: (..., C, N): hi*hi + lo*hi + hi*lo, fp32 accumulate
    ah, al = split(a32); bh, bl = split(b32)
    return ah @ bh + (al @ bh + ah @ bl)

rel = lambda a: float((a.double() - ref).norm() / ref.norm())
# direct conv, 3 x bf16 products (im2col)
xp = torch.nn.functional.unfold(x.float()[None], 3, padding=1)[0]     # (C*9, P)
d3 = mm3(w.float().reshape(K, -1), xp).reshape(K, H, W)
d1 = (w.float().reshape(K, -1).to(torch.bfloat16).float() @ xp.to(torch.bfloat16).float()).reshape(K, H, W)
dfp32 = (w.float().reshape(K, -1) @ xp).reshape(K, H, W)
print("direct fp32            ", rel(dfp32))
print("direct 3 x bf16 split  ", rel(d3))
print("direct 1 x bf16        ", rel(d1))
# Winograd F(2x2, 3x3)
Bt = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float32)
G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float32)
At = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float32)
xpad = torch.nn.functional.pad(x.float(), (1, 1, 1, 1))
tiles = xpad.unfold(1, 4, 2).unfold(2, 4, 2)                           # (C, H/2, W/2, 4, 4)
V = Bt @ tiles @ Bt.T                                                  # input transform in fp32
U = G @ w.float() @ G.T                                                # (K, C, 4, 4) in fp32
Vm = V.permute(3, 4, 0, 1, 2).reshape(16, C, -1)                       # (16, C, T)
Um = U.permute(2, 3, 0, 1).reshape(16, K, C)
for name, M in (("winograd fp32 GEMMs   ", Um @ Vm), ("winograd 3 x bf16     ", mm3(Um, Vm)),
                ("winograd 1 x bf16     ", Um.to(torch.bfloat16).float() @ Vm.to(torch.bfloat16).float())):
    Mt = M.reshape(4, 4, K, H // 2, W // 2).permute(2, 3, 4, 0, 1)
    Y = At @ Mt @ At.T                                                 # (K, H/2, W/2, 2, 2)
    y = Y.permute(0, 1, 3, 2, 4).reshape(K, H, W)
    print(name, rel(y))
