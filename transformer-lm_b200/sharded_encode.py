"""Multi-GPU bulk encode: one process per GPU, the text cut into one shard per rank at EXACT boundaries, no data collective.

The reference encodes in one process (models/tokenizer/encode.py:18-38).  SURVEY 8(e): `Tokenizer.encode` splits the text
on the special tokens first (tokenizer.py:63-66, 113), so the start of a special-token occurrence is a cut that changes
nothing on either side; between specials a lone U+0020 between two ASCII non-space bytes is exact as well.
Rank r encodes the bytes [cut_r, cut_(r+1)) with the single-GPU encoder; only the token counts are exchanged (an
all-gather of one int64 per rank) to place every rank's ids in the global array.

The host logic (cuts, offsets, assembling the output file) talks to the device through `tokenizer.encode_to_numpy`, so the
tests run it under `gloo` on CPU with a checker-backed tokenizer and under `nccl` on the GPUs with the product's.
"""
from __future__ import annotations

import os
import re
from typing import Callable, List, Tuple

import numpy as np

_PLAIN_SPACE = re.compile(rb"[\x21-\x7e] (?=[\x21-\x7e])")
CUT_WINDOW = 8 << 20


def _overlapped_from_left(buf: bytes, specials: List[bytes], p: int) -> bool:
    """Does a special-token occurrence that starts before p extend past p?"""
    for t in specials:
        for q in range(max(p - len(t) + 1, 0), p):
            if buf.startswith(t, q):
                return True
    return False


def first_exact_cut(peek: Callable[[int, int], bytes], size: int, specials: List[bytes], nominal: int, window: int = CUT_WINDOW) -> int:
    """Smallest exact cut position >= nominal (`size` when nominal >= size): the start of a special-token occurrence that no
    other occurrence overlaps from the left, else a lone space between two ASCII non-space bytes.  `peek(lo, hi)` returns the
    text bytes [lo, hi).  Raises when the text has no exact boundary within `window` bytes, doubling the window up to the end."""
    if nominal <= 0:
        return 0
    if nominal >= size:
        return size
    specials = [s for s in specials if s]
    margin = max([len(s) for s in specials] + [1])
    while True:
        lo, hi = max(0, nominal - margin), min(size, nominal + window + margin)
        buf = bytes(peek(lo, hi))
        at = nominal - lo
        best = None
        for s in specials:
            p = buf.find(s, at)
            while p >= 0 and p < at + window:
                if not _overlapped_from_left(buf, specials, p):
                    best = p if best is None else min(best, p)
                    break
                p = buf.find(s, p + 1)
        if best is None:
            for m in _PLAIN_SPACE.finditer(buf, max(at - 1, 0), min(len(buf), at + window)):
                p = m.start() + 1
                if p >= at and not _overlapped_from_left(buf, specials, p) and not _overlapped_from_left(buf, specials, p + 1):
                    best = p
                    break
        if best is not None:
            return lo + best
        if hi >= size:
            raise RuntimeError("no exact cut point after byte %d: the text cannot be sharded there" % nominal)
        window *= 2


def shard_cuts(peek, size: int, specials: List[bytes], world: int) -> List[int]:
    """world + 1 cut positions: rank r owns [cuts[r], cuts[r + 1]).  Every rank computes the same list."""
    cuts = [0]
    for r in range(1, world):
        try:
            c = first_exact_cut(peek, size, specials, size * r // world)
        except RuntimeError:
            c = size                             # no exact boundary from there on: the rest stays in one piece (later ranks idle)
        cuts.append(max(cuts[-1], c))
    cuts.append(size)
    return cuts


def _dist_info(group):
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return None, 0, 1
    return dist, dist.get_rank(group), dist.get_world_size(group)


def _gather_counts(dist, group, n_local: int, device=None) -> List[int]:
    import torch
    world = dist.get_world_size(group)
    t = torch.tensor([n_local], dtype=torch.int64, device=device)
    out = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
    return [int(x) for x in out.cpu()]


def _collective_device(tokenizer):
    import torch
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        ctx = getattr(tokenizer, "_tok_ctx", None) or getattr(tokenizer, "_ctx", None)
        return torch.device("cuda", ctx.device if ctx is not None else torch.cuda.current_device())
    return None


def encode_sharded(tokenizer, data, dtype=np.uint16, group=None) -> Tuple[np.ndarray, int, int]:
    """Every rank passes the SAME text (bytes / uint8 array).  Returns (ids of this rank's shard, global index of its first
    id, total number of ids): the concatenation of the ranks' arrays in rank order equals Tokenizer.encode of the text."""
    dist, rank, world = _dist_info(group)
    arr = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.asarray(data, dtype=np.uint8)
    specials = [s.encode("utf-8") for s in tokenizer.special_tokens]
    cuts = shard_cuts(lambda lo, hi: arr[lo:hi].tobytes(), arr.size, specials, world)
    ids = tokenizer.encode_to_numpy(arr[cuts[rank]: cuts[rank + 1]], dtype)
    if world == 1:
        return ids, 0, int(ids.size)
    counts = _gather_counts(dist, group, int(ids.size), _collective_device(tokenizer))
    return ids, sum(counts[:rank]), sum(counts)


def encode_file_sharded(tokenizer, input_path, output_path, dtype=np.uint16, piece_bytes: int = 1 << 30, group=None) -> int:
    """encode_file over the ranks of an initialised torch.distributed group: rank r reads and encodes only its byte range
    of the file and writes its ids at their global offset of `output_path` (which every rank must be able to open: a shared
    file system, or one node).  Returns the total token count on every rank."""
    from .encode_file import encode_file
    dist, rank, world = _dist_info(group)
    if world == 1:
        return encode_file(tokenizer, input_path, output_path, dtype, piece_bytes)
    dtype = np.dtype(dtype)
    size = os.path.getsize(input_path)
    specials = [s.encode("utf-8") for s in tokenizer.special_tokens]
    with open(input_path, "rb") as f:
        def peek(lo, hi):
            f.seek(lo)
            return f.read(hi - lo)
        cuts = shard_cuts(peek, size, specials, world)
    part = "%s.part%d" % (output_path, rank)
    n_local = encode_file(tokenizer, input_path, part, dtype, piece_bytes, byte_range=(cuts[rank], cuts[rank + 1]))
    counts = _gather_counts(dist, group, n_local, _collective_device(tokenizer))
    total, first = sum(counts), sum(counts[:rank])
    if rank == 0:
        with open(output_path, "wb") as out:
            out.truncate(total * dtype.itemsize)
    dist.barrier(group=group)
    with open(part, "rb") as src, open(output_path, "r+b") as out:
        out.seek(first * dtype.itemsize)
        while True:
            block = src.read(64 << 20)
            if not block:
                break
            out.write(block)
    os.remove(part)
    dist.barrier(group=group)
    return total
