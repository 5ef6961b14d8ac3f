"""GPT-2 pretokenizer front-end (bpe_pretokenize): replaces regex.finditer with the GPT-2 pattern,
models/tokenizer/train.py:143-146,21-23 and models/tokenizer/tokenizer.py:63-90 of the reference."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def pretoken_starts(data, special_tokens: list[str] | None = None, *, ctx=None) -> np.ndarray:
    """Ascending byte offsets at which pretokens start.  With special_tokens the text is first split on them
    (leftmost, longest first) and every special occurrence is reported as one pretoken."""
    ctx = ctx or _lib.default_context()
    L = _lib.lib()
    arr = _lib.as_u8(data)
    specials = sorted(set(special_tokens or []), key=len, reverse=True)
    sp_blob, sp_offs = _lib.pack_blobs([s.encode("utf-8") for s in specials])
    n_out = C.c_uint64(0)
    cap = arr.size
    out = np.zeros(max(cap, 1), dtype=np.uint64)
    rc = L.bpe_pretokenize(ctx.handle, _lib.ptr(arr) if arr.size else None, arr.size, _lib.ptr(sp_blob), _lib.ptr(sp_offs),
                           len(specials), _lib.ptr(out), cap, C.byref(n_out))
    if rc == _lib.ERR_UTF8:
        bytes(arr).decode("utf-8")
    ctx.check(rc)
    return out[: n_out.value]


def pretokens(text: str, special_tokens: list[str] | None = None, *, ctx=None) -> list[str]:
    data = text.encode("utf-8")
    st = pretoken_starts(data, special_tokens, ctx=ctx).tolist() + [len(data)]
    return [data[st[i]:st[i + 1]].decode("utf-8") for i in range(len(st) - 1)]


def utf8_validate(data, *, ctx=None) -> int:
    """-1 when valid, else the offset of the first ill-formed sequence (UnicodeDecodeError.start)."""
    ctx = ctx or _lib.default_context()
    arr = _lib.as_u8(data)
    rc = _lib.lib().bpe_utf8_validate(ctx.handle, _lib.ptr(arr) if arr.size else None, arr.size)
    if rc == _lib.ERR_UTF8:
        return int(_lib.lib().bpe_last_error_detail(ctx.handle))
    ctx.check(rc)
    return -1
