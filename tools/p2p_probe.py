"""Probe (2+ ranks under torchrun): can a rank map a peer's CUDA buffer (torch IPC handles / symmetric memory) and
read / write it from a kernel?  Prints one JSON line per rank.  Not a test."""
import json, os, sys
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = {"rank": rank, "world": world}
try:
    out["can_access_peer"] = [bool(torch.cuda.can_device_access_peer(local, p)) for p in range(world) if p != local]
except Exception as e:
    out["can_access_peer"] = repr(e)

# --- torch's CUDA IPC handles -----------------------------------------------------------------------------------
try:
    buf = torch.full((1 << 20,), float(rank + 1), device=dev)
    handle = buf.untyped_storage()._share_cuda_()
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    peers = {}
    for p in range(world):
        if p == rank:
            continue
        st = torch.UntypedStorage._new_shared_cuda(*handles[p])
        t = torch.empty(0, dtype=torch.float32, device=st.device).set_(st)
        peers[p] = t
    torch.cuda.synchronize()
    dist.barrier()
    out["ipc_read"] = {p: float(t[:8].sum().item()) for p, t in peers.items()}      # expect 8 * (p + 1)
    out["ipc_peer_device"] = {p: str(t.device) for p, t in peers.items()}
    for p, t in peers.items():                                                       # remote write
        t[rank * 16:(rank + 1) * 16] = 100.0 + rank
    torch.cuda.synchronize()
    dist.barrier()
    out["ipc_after_remote_write"] = float(buf[: 16 * world].sum().item())
    # bandwidth of a plain peer read (copy kernel)
    for p, t in peers.items():
        dst = torch.empty_like(t)
        big = t
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dst.copy_(big)
        e1.record()
        torch.cuda.synchronize()
        out.setdefault("ipc_read_GBps", {})[p] = 20 * big.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
        break
except Exception as e:
    out["ipc_error"] = repr(e)[:300]

# --- symmetric memory (torch.distributed._symmetric_memory) ---------------------------------------------------------
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
    out["symm_read"] = float(peer[:8].sum().item())
    out["symm_has_multicast"] = bool(getattr(hdl, "multicast_ptr", 0))
except Exception as e:
    out["symm_error"] = repr(e)[:300]

print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
