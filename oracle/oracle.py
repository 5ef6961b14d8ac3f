"""ctypes wrapper around the CPU oracle (oracle/bpe_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of bpe_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  The product package never does.

API mirrors the reference (paths relative to /root/reference):
  train_bpe(input_path, vocab_size, special_tokens)        models/tokenizer/train.py:142-231
  OracleTokenizer(vocab, merges, special_tokens)            models/tokenizer/tokenizer.py:11-167
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib
import subprocess
from typing import Iterable, Iterator

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libbpe_oracle.so"
_lib = None


def build(force: bool = False) -> pathlib.Path:
    src = _HERE / "bpe_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-B", "libbpe_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        u8p, u32p, u64p, i32p, i64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_int32, C.c_int64))
        L.orc_cp_class.restype = C.c_int
        L.orc_cp_class.argtypes = [C.c_uint32]
        L.orc_utf8_validate.restype = C.c_int64
        L.orc_utf8_validate.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_universal_newlines.restype = C.c_uint64
        L.orc_universal_newlines.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_pretokenize.restype = C.c_int64
        L.orc_pretokenize.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_count_pretokens.restype = C.c_int64
        L.orc_count_pretokens.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, u64p]
        L.orc_train_bpe.restype = C.c_int
        L.orc_train_bpe.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_void_p, C.POINTER(C.c_int), i64p]
        L.orc_tok_create.restype = C.c_void_p
        L.orc_tok_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_tok_destroy.restype = None
        L.orc_tok_destroy.argtypes = [C.c_void_p]
        L.orc_encode.restype = C.c_int
        L.orc_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, u64p, u64p, u32p]
        L.orc_decode_bytes.restype = C.c_int
        L.orc_decode_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, u64p, u64p]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _buf(b: bytes) -> np.ndarray:
    a = np.frombuffer(b, dtype=np.uint8)
    return a if a.size else np.zeros(1, dtype=np.uint8)


def _pack_specials(specials: list[bytes]):
    offs = np.zeros(len(specials) + 1, dtype=np.uint32)
    for i, s in enumerate(specials):
        offs[i + 1] = offs[i] + len(s)
    return _buf(b"".join(specials)), offs


def cp_class(cp: int) -> int:
    return lib().orc_cp_class(cp)


def utf8_validate(data: bytes) -> int:
    a = _buf(data)
    return lib().orc_utf8_validate(_ptr(a), len(data))


def pretokenize(data: bytes) -> list[int]:
    """Byte offsets of every GPT-2 pretoken start in `data` (valid UTF-8)."""
    a = _buf(data)
    n = lib().orc_pretokenize(_ptr(a), len(data), None, 0)
    if n < 0:
        raise ValueError("invalid utf-8")
    out = np.zeros(max(n, 1), dtype=np.uint64)
    lib().orc_pretokenize(_ptr(a), len(data), _ptr(out), n)
    return out[:n].tolist()


def pretokens(text: str) -> list[str]:
    data = text.encode("utf-8")
    st = pretokenize(data) + [len(data)]
    return [data[st[i]:st[i + 1]].decode("utf-8") for i in range(len(st) - 1)]


def count_pretokens(data: bytes, special_tokens: list[str] = ()) -> dict[bytes, int]:
    """extract_subword_frequencies on already newline-translated bytes (train.py:16-28)."""
    a = _buf(data)
    sb, so = _pack_specials([s.encode("utf-8") for s in special_tokens])
    nb = C.c_uint64(0)
    n = lib().orc_count_pretokens(_ptr(a), len(data), _ptr(sb), _ptr(so), len(special_tokens), None, None, None, C.byref(nb))
    if n < 0:
        raise ValueError("invalid utf-8")
    blob = np.zeros(max(nb.value, 1), dtype=np.uint8)
    offs = np.zeros(n + 1, dtype=np.uint64)
    cnts = np.zeros(max(n, 1), dtype=np.int64)
    lib().orc_count_pretokens(_ptr(a), len(data), _ptr(sb), _ptr(so), len(special_tokens),
                              _ptr(blob), _ptr(offs), _ptr(cnts), C.byref(nb))
    raw = blob.tobytes()
    return {raw[int(offs[i]):int(offs[i + 1])]: int(cnts[i]) for i in range(n)}


def _ref_vocab(special_tokens: list[str]) -> dict[int, bytes]:
    """Vocab.__init__ (models/tokenizer/vocab.py:2-13): specials, then 256 bytes, deduped by value."""
    idx_to_token: dict[int, bytes] = {}
    seen = set()
    for tok in [s.encode("utf-8") for s in special_tokens] + [bytes([i]) for i in range(256)]:
        if tok in seen:
            continue
        seen.add(tok)
        idx_to_token[len(idx_to_token)] = tok
    return idx_to_token


def train_bpe_on_bytes(data: bytes, vocab_size: int, special_tokens: list[str] = []):
    vocab = _ref_vocab(list(special_tokens))
    n_merges = max(vocab_size - len(vocab), 0)          # range(vocab_size - len(vocab)), train.py:183
    a = _buf(data)
    sb, so = _pack_specials([s.encode("utf-8") for s in special_tokens])
    pairs = np.zeros((max(n_merges, 1), 2), dtype=np.int32)
    n_done = C.c_int(0)
    err = C.c_int64(-1)
    rc = lib().orc_train_bpe(_ptr(a), len(data), _ptr(sb), _ptr(so), len(special_tokens), n_merges,
                             _ptr(pairs), C.byref(n_done), C.byref(err))
    if rc == -1:
        data.decode("utf-8")                               # raises the UnicodeDecodeError the reference raises
        raise AssertionError("oracle flagged invalid UTF-8 at %d but CPython accepted it" % err.value)
    sym = [bytes([i]) for i in range(256)]
    merges = []
    seen = set(vocab.values())
    for k in range(n_done.value):
        x, y = sym[pairs[k, 0]], sym[pairs[k, 1]]
        merges.append((x, y))
        sym.append(x + y)
        if x + y not in seen:                              # Vocab.add_token dedupe (vocab.py:28-34)
            seen.add(x + y)
            vocab[len(vocab)] = x + y
    return vocab, merges


def train_bpe(input_path, vocab_size: int, special_tokens: list[str] = []):
    with open(input_path, "rb") as f:
        data = f.read()
    return train_bpe_on_bytes(data, vocab_size, special_tokens)


class OracleTokenizer:
    """Restates Tokenizer (tokenizer.py:11-167) on top of the C core."""

    def __init__(self, vocab: dict[int, bytes], merges: list[tuple[bytes, bytes]], special_tokens=[]):
        self.vocab = vocab
        self.vocab_inv = {v: k for k, v in vocab.items()}
        self.merges = merges
        self.special_tokens = list(set(special_tokens or []))
        self.special_tokens.sort(key=len, reverse=True)
        for token in self.special_tokens:                 # tokenizer.py:35-38 (quirk kept, SURVEY A-12)
            tb = token.encode("utf-8")
            if tb not in self.vocab_inv:
                self.vocab[tb] = len(self.vocab)
                self.vocab_inv[tb] = len(self.vocab) - 1
        self._h = None
        self._build()

    def _build(self):
        items = [(k, v) for k, v in self.vocab.items() if isinstance(k, int) and isinstance(v, bytes)]
        # entries of vocab_inv that did not come from int->bytes items (the A-12 quirk) still resolve on encode
        inv_items = list(self.vocab_inv.items())
        vb = b"".join(k for k, _ in inv_items)
        voffs = np.zeros(len(inv_items) + 1, dtype=np.uint64)
        np.cumsum([len(k) for k, _ in inv_items], out=voffs[1:]) if inv_items else None
        vids = np.array([v for _, v in inv_items] or [0], dtype=np.int64)
        mb = b"".join(a + b for a, b in self.merges)
        moffs = np.zeros(2 * len(self.merges) + 1, dtype=np.uint64)
        if self.merges:
            lens = np.array([l for a, b in self.merges for l in (len(a), len(b))], dtype=np.uint64)
            np.cumsum(lens, out=moffs[1:])
        sb, so = _pack_specials([s.encode("utf-8") for s in self.special_tokens])
        self._keep = (_buf(vb), voffs, vids, _buf(mb), moffs, sb, so)
        self._h = lib().orc_tok_create(_ptr(self._keep[0]), _ptr(voffs), _ptr(vids), len(inv_items),
                                       _ptr(self._keep[3]), _ptr(moffs), len(self.merges),
                                       _ptr(sb), _ptr(so), len(self.special_tokens))
        self._decode_vocab = dict(items)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_tok_destroy(self._h)
            self._h = None

    def encode_bytes(self, data: bytes) -> np.ndarray:
        a = _buf(data)
        n_out = C.c_uint64(0)
        koff = C.c_uint64(0)
        klen = C.c_uint32(0)
        rc = lib().orc_encode(self._h, _ptr(a), len(data), None, 0, C.byref(n_out), C.byref(koff), C.byref(klen))
        if rc == -2:
            raise KeyError(data[koff.value:koff.value + klen.value])
        if rc == -1:
            raise ValueError("invalid utf-8")
        out = np.zeros(max(n_out.value, 1), dtype=np.int64)
        lib().orc_encode(self._h, _ptr(a), len(data), _ptr(out), n_out.value, C.byref(n_out), C.byref(koff), C.byref(klen))
        return out[:n_out.value]

    def encode(self, text: str) -> list[int]:
        return self.encode_bytes(text.encode("utf-8")).tolist()

    def encode_iterable(self, iterable: Iterable[str]) -> Iterator[int]:
        while True:                                        # tokenizer.py:140-153
            text = ""
            for line in iterable:
                text += line
                if len(text) >= 1024 * 1024 * 2:
                    break
            if not text:
                break
            yield from self.encode(text)

    def decode(self, ids: list[int]) -> str:
        return b"".join([self.vocab[i] for i in ids]).decode("utf-8", errors="replace")
