"""N-rank check of smb_dist_adam_step (run under torchrun on N GPUs of one box; tests/test_gpu_dist_adam.py spawns it):
the fused reduce-scatter + Adam + all-gather kernel against NCCL all_reduce + the single-GPU Adam kernel on the same
per-rank gradients.  Prints one JSON line on rank 0 and exits non-zero on any mismatch."""
import json, os, sys
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylemesh_b200 import engine as eng

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

sizes = [3 * 257 * 255, 3 * 128 * 128, 3 * 64 * 64, 3 * 33 * 31]        # ragged layers, 64-float aligned segments
begin, off = [], 0
for n in sizes:
    begin.append(off)
    off += (n + 63) // 64 * 64
coefs = [2.0 * 5e3 * w / n for w, n in zip([8.0, 4.0, 2.0, 0.0], sizes)]


def alloc(n, dt):
    t = symm.empty(n, dtype=dt, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD)
    t.zero_()
    return t, h


param, hp = alloc(off, torch.float32)
grad, hg = alloc(off, torch.float32)
flags, hf = alloc(64, torch.int32)
m, v = torch.zeros(off, device=dev), torch.zeros(off, device=dev)
g0 = torch.Generator(device="cpu").manual_seed(1)
init = torch.zeros(off)
for a, n in zip(begin, sizes):
    init[a:a + n] = torch.rand(n, generator=g0) * 400 - 200
param.copy_(init)
ref_p, ref_m, ref_v = init.clone().to(dev), torch.zeros(off, device=dev), torch.zeros(off, device=dev)
hp.barrier()

ok, worst = True, 0.0
for step in range(1, 6):
    gr = torch.zeros(off)
    gen = torch.Generator(device="cpu").manual_seed(100 * step + rank)          # different views on every rank
    for a, n in zip(begin, sizes):
        gr[a:a + n] = torch.randn(n, generator=gen) * (1.0 if step % 2 else 1e-3)
        gr[a:a + n][torch.rand(n, generator=gen) < 0.5] = 0.0                    # untouched texels
    grad.copy_(gr)
    ref_g = gr.to(dev)
    dist.all_reduce(ref_g)
    eng.adam_step_segments(ref_p, ref_g, ref_m, ref_v, begin, coefs, 1.0, 0.9, 0.999, 1e-8, step, grad_scale=1.0 / world)
    eng.dist_adam_step(rank, world, [int(x) for x in hg.buffer_ptrs], [int(x) for x in hp.buffer_ptrs],
                       [int(x) for x in hf.buffer_ptrs], m, v, off, begin, coefs, 1.0, 0.9, 0.999, 1e-8, step, epoch=step)
    torch.cuda.synchronize()
    # every replica bit-identical (one writer per texel) ...
    mine = param.clone()
    allp = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allp, mine)
    same = all(torch.equal(allp[0], t) for t in allp)
    # ... the gradient buffer is zero again, and the result equals all_reduce + local Adam (bit for bit at N = 2:
    # a + b is commutative; for N > 2 the summation order differs from NCCL's ring, compare to fp32 rounding)
    zero = float(grad.abs().max()) == 0.0
    diff = float((param - ref_p).abs().max())
    worst = max(worst, diff)
    tol = 0.0 if world == 2 else 2e-3
    ok = ok and same and zero and diff <= tol
    hp.barrier()
res = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "ok": bool(res.item() == 1.0), "max_abs_diff_vs_allreduce_path": worst}), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if res.item() == 1.0 else 1)
