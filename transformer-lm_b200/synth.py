"""Synthetic TinyStories- / OWT-shaped corpora (bench and test infrastructure).
Thin front-end over bpe_synth_host / bpe_synth_dev (csrc/synth_gen.h): the bytes are a pure function
of (shape, seed, n) and identical on CPU and GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

SHAPES = {"tinystories": _lib.SYNTH_TINYSTORIES, "tiny": _lib.SYNTH_TINYSTORIES, "owt": _lib.SYNTH_OWT}
BLOCK = 4096
# seeds fixed by SURVEY 8d
SEED_TINY_TRAIN, SEED_OWT_TRAIN, SEED_OWT_ENCODE = 1234, 4321, 4322


def synth_host(shape: str, seed: int, n: int, out: np.ndarray | None = None, first_block: int = 0) -> np.ndarray:
    a = out if out is not None else np.empty(n, dtype=np.uint8)
    rc = _lib.lib().bpe_synth_host_at(SHAPES[shape], seed, first_block, _lib.ptr(a) if n else None, n)
    if rc != 0:
        raise _lib.BpeError(rc, "bpe_synth_host")
    return a[:n]


def synth_device(shape: str, seed: int, n: int, device_ptr: int, *, ctx=None, first_block: int = 0) -> None:
    ctx = ctx or _lib.default_context()
    ctx.check(_lib.lib().bpe_synth_dev_at(ctx.handle, SHAPES[shape], seed, first_block, C.c_void_p(device_ptr), n))
