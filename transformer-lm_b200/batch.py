"""Training-batch feeder: mirrors load_batch of the reference (models/util.py:37-57) with the token array resident in HBM.

    load_batch(dataset, batch_size, context_length, device, generator=None) -> (inputs, targets)   # torch.long, on `device`

The reference draws `batch_size` start indices with torch.randint on the host and copies the windows row by row in Python;
here the same indices (same generator, same values) go to `bpe_batch_windows_dev`, which gathers both windows of every row
in one kernel from the device copy of the token array.  The array is uploaded once per dataset object (uint16 / int32, the
encoder's output formats; anything else is converted to int32) and kept until the dataset is garbage collected."""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib

_resident: dict[tuple, tuple] = {}               # (id(dataset), device index) -> (weakref, device tensor, dtype code)
_contexts: dict[int, list] = {}                  # device index -> [Context, cuda stream handle it is on]


def _device_tokens(dataset, device):
    """Device copy of the token array, uploaded once per (dataset object, device) and kept until the dataset is collected.
    `device` is a torch.device with an explicit index.  (An in-place change of the host array is not seen: pass a new object.)"""
    import torch
    key = (id(dataset), device.index)
    hit = _resident.get(key)
    if hit is not None and hit[0]() is dataset:
        return hit[1], hit[2]
    if isinstance(dataset, torch.Tensor):
        arr = dataset
        if arr.dtype not in (torch.uint16, torch.int32):
            arr = arr.to(torch.int32)
        dev = arr.to(device).contiguous()
    else:
        arr = np.asarray(dataset)
        if arr.dtype not in (np.uint16, np.int32):
            if arr.size and (arr.min() < 0 or arr.max() >= 2**31):
                raise ValueError("token ids must fit int32")
            arr = arr.astype(np.int32)
        dev = torch.from_numpy(np.ascontiguousarray(arr)).to(device)
    code = _lib.DTYPE_U16 if dev.dtype == torch.uint16 else _lib.DTYPE_I32
    try:
        ref = weakref.ref(dataset, lambda _r, k=key: _resident.pop(k, None))
    except TypeError:                            # not weak-referenceable: do not cache
        return dev, code
    _resident[key] = (ref, dev, code)
    return dev, code


def load_batch(dataset, batch_size: int, context_length: int, device: str, generator=None, *, ctx=None):
    """Same signature, same random draws and same result as the reference's load_batch (models/util.py:37-57)."""
    import torch
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.BpeError(_lib.ERR_NO_DEVICE, "load_batch gathers on a B200: device must be a cuda device (there is no CPU path)")
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    dev = torch.device("cuda", index)            # torch.device("cuda") != torch.device("cuda:0"): always carry the index
    # The gather runs on torch's CURRENT stream: x / y come from torch's caching allocator and the token array may just have
    # been copied by torch, so the kernel must be ordered with torch's queued work (a private stream would race with it).
    stream = torch.cuda.current_stream(dev).cuda_stream
    if ctx is None:
        slot = _contexts.get(index)
        if slot is None:
            slot = _contexts[index] = [_lib.Context(index), None]
        ctx = slot[0]
        if slot[1] != stream:
            ctx.use_stream(stream)
            slot[1] = stream
    else:
        ctx.use_stream(stream)
    n = len(dataset)
    limit = n - context_length                   # models/util.py:48
    start_idx = torch.randint(limit, (batch_size,), generator=generator)   # host generator, like the reference (:49)
    tokens, code = _device_tokens(dataset, dev)
    x = torch.empty((batch_size, context_length), dtype=torch.long, device=dev)
    y = torch.empty((batch_size, context_length), dtype=torch.long, device=dev)
    starts = np.ascontiguousarray(start_idx.numpy().astype(np.int64))
    with ctx.lock:
        ctx.check(_lib.lib().bpe_batch_windows_dev(ctx.handle, C.c_void_p(tokens.data_ptr()), code, n, _lib.ptr(starts), batch_size,
                                                   context_length, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr())))
    return x, y
