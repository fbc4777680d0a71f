"""Host -> device staging of step inputs, overlapped with the previous step's kernels.

The reference relies on Lightning to move each 13-tuple (data/abstract_dataset.py:329-342) to the GPU right before
`training_step` (model/model.py:346): the copy sits on the compute stream, so every step pays PCIe time (16 MB per
640x480 view) before its first kernel.  `BatchStager` keeps two device-side slots and a copy stream:

    t_next = stager.stage(batch_0)
    for i in range(n):
        t, t_next = t_next, (stager.stage(batch_{i+1}) if i + 1 < n else None)   # next copy is issued first
        batch = stager.acquire(t)          # compute stream waits for the copy of THIS step only
        ... run the step on `batch` ...
        stager.release(t)                  # the slot may be overwritten once these kernels are done

Inputs may be plain (nested) tensors — one cudaMemcpyAsync per tensor, asynchronous when the source is pinned — or a
`PackedBatch` (all tensors collated into ONE pinned buffer, one copy per step; what a pinning DataLoader worker
would hand over).  Element 8 of a 13-tuple (the dataset index) stays on the host: it keys the per-view plan cache.
"""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

import torch

_ALIGN = 256


def _is_view_tuple(obj) -> bool:
    return isinstance(obj, (list, tuple)) and len(obj) == 13 and isinstance(obj[8], torch.Tensor)


def _flatten(obj, out: List[torch.Tensor], keep_host: List[bool], _top: bool = True):
    """Returns a structure mirror where every tensor is replaced by its index into `out`."""
    if isinstance(obj, torch.Tensor):
        out.append(obj)
        keep_host.append(False)
        return ("t", len(out) - 1)
    if isinstance(obj, (list, tuple)):
        kids = []
        for j, o in enumerate(obj):
            node = _flatten(o, out, keep_host, False)
            if _top and j == 8 and _is_view_tuple(obj) and node[0] == "t":
                keep_host[node[1]] = True
            kids.append(node)
        return ("l" if isinstance(obj, list) else "u", kids)
    return ("c", obj)


def _rebuild(node, leaves: List[Any]):
    kind, val = node
    if kind == "t":
        return leaves[val]
    if kind == "c":
        return val
    seq = [_rebuild(k, leaves) for k in val]
    return seq if kind == "l" else tuple(seq)


class PackedBatch:
    """A step's tensors collated into one pinned byte buffer (built once per view, e.g. by the data loader)."""

    def __init__(self, batch, pin: bool = True):
        tensors: List[torch.Tensor] = []
        keep: List[bool] = []
        self.tree = _flatten(batch, tensors, keep)
        self.meta: List[Optional[Tuple[int, torch.Size, torch.dtype]]] = []
        self.host_leaves: List[Optional[torch.Tensor]] = []
        off = 0
        for t, k in zip(tensors, keep):
            if k:
                self.meta.append(None)
                self.host_leaves.append(t)
                continue
            nbytes = t.numel() * t.element_size()
            self.meta.append((off, t.shape, t.dtype))
            self.host_leaves.append(None)
            off = (off + nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        self.nbytes = off
        self.payload_bytes = int(sum(t.numel() * t.element_size() for t, k in zip(tensors, keep) if not k))
        self.host = torch.empty(max(off, 1), dtype=torch.uint8)
        if pin:                                   # pin=False only for host-side tests on a box without a GPU
            self.host = self.host.pin_memory()
        for t, m in zip(tensors, self.meta):
            if m is None:
                continue
            o, _, _ = m
            n = t.numel() * t.element_size()
            if n:
                self.host[o:o + n].copy_(t.detach().contiguous().reshape(-1).view(torch.uint8))

    def views_of(self, dev_buf: torch.Tensor):
        leaves = []
        for m, h in zip(self.meta, self.host_leaves):
            if m is None:
                leaves.append(h)
                continue
            o, shape, dtype = m
            n = int(torch.Size(shape).numel()) * torch.empty((), dtype=dtype).element_size()
            leaves.append(dev_buf[o:o + n].view(dtype).view(shape))
        return _rebuild(self.tree, leaves)


class _Slot:
    def __init__(self):
        self.flat: Optional[torch.Tensor] = None        # device byte buffer for PackedBatch inputs
        self.tensors: List[Optional[torch.Tensor]] = []  # device tensors for plain inputs
        self.ready = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.used = False


class BatchStager:
    def __init__(self, device, slots: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("BatchStager stages onto a CUDA device")
        self.copy_stream = torch.cuda.Stream(self.device)
        self._slots = [_Slot() for _ in range(max(2, slots))]
        self._next = 0

    def stage(self, batch):
        """Issue the H2D copy of `batch` (nested tensors or a PackedBatch) on the copy stream; returns a ticket."""
        if not isinstance(batch, PackedBatch) and self._resident(batch):
            return None, batch                     # e.g. a ViewStore batch: already in HBM, nothing to copy or order
        slot = self._slots[self._next]
        self._next = (self._next + 1) % len(self._slots)
        with torch.cuda.stream(self.copy_stream):
            if slot.used:
                self.copy_stream.wait_event(slot.done)      # the step that read this slot has finished
            if isinstance(batch, PackedBatch):
                if slot.flat is None or slot.flat.numel() < batch.nbytes:
                    slot.flat = torch.empty(max(batch.nbytes, 1), dtype=torch.uint8, device=self.device)
                slot.flat[:batch.nbytes].copy_(batch.host[:batch.nbytes], non_blocking=True)
                staged = batch.views_of(slot.flat)
            else:
                tensors: List[torch.Tensor] = []
                keep: List[bool] = []
                tree = _flatten(batch, tensors, keep)
                if len(slot.tensors) != len(tensors):
                    slot.tensors = [None] * len(tensors)
                leaves = []
                for j, (t, k) in enumerate(zip(tensors, keep)):
                    if k:
                        leaves.append(t)
                        continue
                    d = slot.tensors[j]
                    if d is None or d.shape != t.shape or d.dtype != t.dtype:
                        d = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                        slot.tensors[j] = d
                    d.copy_(t, non_blocking=True)
                    leaves.append(d)
                staged = _rebuild(tree, leaves)
            slot.ready.record(self.copy_stream)
        slot.used = True
        return slot, staged

    def _resident(self, batch) -> bool:
        tensors: List[torch.Tensor] = []
        keep: List[bool] = []
        _flatten(batch, tensors, keep)
        return bool(tensors) and all(k or (t.is_cuda and t.device == self.device) for t, k in zip(tensors, keep))

    def acquire(self, ticket):
        """Make the current stream wait for the ticket's copy; returns the device-side batch."""
        slot, staged = ticket
        if slot is not None:
            torch.cuda.current_stream(self.device).wait_event(slot.ready)
        return staged

    def release(self, ticket) -> None:
        """Call after the step's kernels have been enqueued on the current stream."""
        slot, _ = ticket
        if slot is not None:
            slot.done.record(torch.cuda.current_stream(self.device))
