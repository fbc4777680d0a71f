"""Byte-bounded LRU for per-view device data (mask plans, content-target features).

Views repeat (RepeatingSampler: each view index_repeat = 20..100 times in a row, data/abstract_dataset.py:498-512, and
again every epoch), so everything that is constant per view is built once and kept in HBM - but a 1000-view scene with a
4-level pyramid is tens of GB of masks and features, so the caches are bounded by bytes and evict the least recently
used view."""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Optional

import torch


def tensor_bytes(obj: Any) -> int:
    """bytes of every tensor reachable through dicts / lists / tuples / objects with __dict__ (shared storage is
    counted once per tensor object)."""
    seen, total, stack = set(), 0, [obj]
    while stack:
        o = stack.pop()
        if isinstance(o, torch.Tensor):
            if id(o) not in seen:
                seen.add(id(o))
                total += o.numel() * o.element_size()
        elif isinstance(o, dict):
            stack.extend(o.values())
        elif isinstance(o, (list, tuple)):
            stack.extend(o)
        elif hasattr(o, "__dict__") and not isinstance(o, type):
            stack.extend(vars(o).values())
    return total


class ViewLRU:
    def __init__(self, max_bytes: int):
        self.max_bytes = int(max_bytes)
        self._items: "OrderedDict[Any, tuple]" = OrderedDict()
        self.bytes = 0
        self.hits = 0
        self.misses = 0
        self.evictions = 0

    def __len__(self):
        return len(self._items)

    def __contains__(self, key):
        return key in self._items

    def get(self, key, default=None) -> Optional[Any]:
        item = self._items.get(key)
        if item is None:
            self.misses += 1
            return default
        self._items.move_to_end(key)
        self.hits += 1
        return item[0]

    def put(self, key, value) -> None:
        if key in self._items:
            self.bytes -= self._items.pop(key)[1]
        n = tensor_bytes(value)
        self._items[key] = (value, n)
        self.bytes += n
        while self.bytes > self.max_bytes and len(self._items) > 1:      # the newest entry always stays
            _, (_, freed) = self._items.popitem(last=False)
            self.bytes -= freed
            self.evictions += 1

    def clear(self) -> None:
        self._items.clear()
        self.bytes = 0
