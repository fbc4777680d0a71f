"""NVTX ranges around the phases of a step (SMB_NVTX=1), for nsys / ncu --nvtx timelines: `sample`, `content_targets`,
`vgg_loss[level i]`, `scatter`, `regulariser`, `optimizer`.  Off by default: a push/pop pair costs ~1 us of host time and
a step issues ~15 of them."""
from __future__ import annotations

import contextlib
import os

_ON = os.environ.get("SMB_NVTX", "0") not in ("", "0")


@contextlib.contextmanager
def _real(name: str):
    import torch
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


_NULL = contextlib.nullcontext()


def range(name: str):          # noqa: A001 - mirrors torch.cuda.nvtx.range
    return _real(name) if _ON else _NULL
