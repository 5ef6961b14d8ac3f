"""Bulk encode driver: text file -> raw little-endian token array on disk.

Replaces models/tokenizer/encode.py:18-38 of the reference.  The reference reads the file in text mode
in 1 Mi-character pieces, calls encode_iterable([piece]) -- which never terminates on a list (SURVEY A-14) -- and
torch.saves a uint16 array as .pt, while its trainer memory-maps a raw .bin (train.py:230-232).  This driver defines
the output as Tokenizer.encode(whole text, read in text mode): strict UTF-8, universal newlines (SURVEY A-2), and writes
the raw `.bin` the trainer's np.memmap(dtype=np.uint16) reads (optionally the reference's .pt as well).

The file is streamed in pieces cut at exact boundaries -- an occurrence of a special token (Tokenizer.segment splits
there first, tokenizer.py:63-66) or a lone U+0020 between two ASCII non-space bytes -- so a file of any
size needs bounded host memory; every piece goes through bpe_encode, which overlaps upload, kernels and download, while
a reader thread fetches the next block and a writer thread stores the previous piece's ids.  `distributed=True` shards the
file over the ranks of a torch.distributed group (sharded_encode.py).
"""
from __future__ import annotations

import os
import re

import numpy as np

from . import _lib

_PLAIN_SPACE = re.compile(rb"[\x21-\x7e] [\x21-\x7e]")


def _overlapped_from_left(buf: bytes, specials: list[bytes], p: int) -> bool:
    """Does a special-token occurrence that starts before p extend past p?"""
    for t in specials:
        for q in range(max(p - len(t) + 1, 0), p):
            if buf.startswith(t, q):
                return True
    return False


def _last_exact_cut(buf: bytes, specials: list[bytes], lo: int) -> int:
    """Largest exact cut position >= lo in buf (0 = none): the start of a special-token occurrence that no other special
    occurrence overlaps from the left, else a lone space between two ASCII non-space bytes."""
    best = 0
    for s in specials:
        p = buf.rfind(s, lo)
        while p > 0:
            if not _overlapped_from_left(buf, specials, p):
                best = max(best, p)
                break
            p = buf.rfind(s, lo, p + len(s) - 1)
    if best:
        return best
    for m in reversed(list(_PLAIN_SPACE.finditer(buf, max(lo, len(buf) - (8 << 20))))):
        p = m.start() + 1
        if not _overlapped_from_left(buf, specials, p):
            return p
    return 0


def encode_file(tokenizer, input_path, output_path, dtype=np.uint16, piece_bytes: int = 1 << 30, save_pt: str | None = None,
                byte_range: tuple[int, int] | None = None, distributed: bool = False, group=None) -> int:
    """Encode `input_path` with `tokenizer` into raw little-endian `dtype` ids at `output_path`.  Returns the token count.

    Three stages run concurrently: a reader thread pulls the next block from the file, the calling thread
    cuts and encodes the current piece (bpe_encode, which itself overlaps upload, kernels and download), a writer thread
    appends the ids of the previous piece to the output.  byte_range = (lo, hi) restricts the work to those bytes of the file
    (lo and hi must be exact cut positions: sharded_encode.py); distributed=True shards the file over the ranks of the
    initialised torch.distributed group."""
    if distributed:
        from .sharded_encode import encode_file_sharded
        return encode_file_sharded(tokenizer, input_path, output_path, dtype, piece_bytes, group)
    import queue
    import threading
    dtype = np.dtype(dtype)
    specials = [s.encode("utf-8") for s in tokenizer.special_tokens]
    lo, hi = byte_range if byte_range is not None else (0, os.path.getsize(input_path))
    blocks: queue.Queue = queue.Queue(maxsize=2)
    results: queue.Queue = queue.Queue(maxsize=2)
    failure: list = []

    def reader():
        try:
            with open(input_path, "rb", buffering=0) as f:
                f.seek(lo)
                left = hi - lo
                while left > 0 and not failure:
                    want = min(piece_bytes, left)
                    block = bytearray(want)
                    mv, got = memoryview(block), 0
                    while got < want:
                        k = f.readinto(mv[got:])
                        if not k:
                            break
                        got += k
                    left -= got
                    blocks.put(block[:got] if got < want else block)
                    if got < want:
                        break
        except BaseException as e:            # surfaced in the calling thread
            failure.append(e)
        blocks.put(None)

    def writer():
        try:
            with open(output_path, "wb") as out:
                while True:
                    ids = results.get()
                    if ids is None:
                        return
                    ids.astype(dtype.newbyteorder("<"), copy=False).tofile(out)
        except BaseException as e:
            failure.append(e)
            while results.get() is not None:
                pass

    rt, wt = threading.Thread(target=reader, daemon=True), threading.Thread(target=writer, daemon=True)
    rt.start(); wt.start()
    total = 0
    offset = lo                                  # bytes of the (untranslated) file consumed so far: for error offsets
    pin = None
    try:
        carry = b""
        eof = False
        while not eof:
            block = blocks.get()
            if failure:
                raise failure[0]
            eof = block is None
            buf = carry if eof else (carry + block if carry else block)
            if not buf:
                break
            hold = b""
            if not eof and buf.endswith(b"\r"):              # the "\n" of a "\r\n" may start the next block
                buf, hold = buf[:-1], b"\r"
            if b"\r" in buf:
                buf = buf.replace(b"\r\n", b"\n").replace(b"\r", b"\n")       # text-mode read of the reference (A-2)
            if eof:
                piece, carry = buf, b""
            else:
                cut = _last_exact_cut(buf, specials, len(buf) // 2)
                if cut == 0:
                    carry = buf + hold                        # no exact boundary yet: read more
                    if len(carry) > 8 * piece_bytes:
                        raise RuntimeError("no exact cut point within %d bytes" % len(carry))
                    continue
                piece, carry = buf[:cut], buf[cut:] + hold
            if not piece:
                continue
            if getattr(tokenizer, "host_only", False):       # (checker-backed tokenizers of the CPU tests: no page-locked staging)
                staged = np.frombuffer(piece, dtype=np.uint8)
            else:
                if pin is None or pin.nbytes < len(piece):
                    if pin is not None:
                        pin.free()
                    pin = _lib.PinnedBuffer(max(len(piece), piece_bytes + (piece_bytes >> 2)))
                pin.array[: len(piece)] = np.frombuffer(piece, dtype=np.uint8)
                staged = pin.array[: len(piece)]
            try:
                ids = tokenizer.encode_to_numpy(staged, dtype)
            except UnicodeDecodeError as e:
                raise UnicodeDecodeError(e.encoding, e.object[max(e.start - 8, 0): e.end + 8], min(e.start, 8), min(e.start, 8) + (e.end - e.start),
                                         e.reason + " (near byte %d of the file)" % (offset + e.start)) from None
            results.put(ids)
            if failure:
                raise failure[0]
            total += ids.size
            offset += len(piece)
    except BaseException:
        failure.append(RuntimeError("encode_file aborted"))   # stops the reader
        while rt.is_alive():
            try:
                blocks.get(timeout=0.05)
            except queue.Empty:
                pass
        raise
    finally:
        results.put(None)
        wt.join()
        rt.join(timeout=5)
        if pin is not None:
            pin.free()
    if failure:
        raise failure[0]
    if save_pt:
        import torch
        torch.save(np.fromfile(output_path, dtype=dtype), save_pt, pickle_protocol=4)       # models/tokenizer/encode.py:37-38
    return total


# file names of the reference's driver (models/tokenizer/encode.py:8-15)
FNAME = {"tiny/train": "TinyStoriesV2-GPT4-train.txt", "tiny/valid": "TinyStoriesV2-GPT4-valid.txt", "owt/train": "owt_train.txt",
         "owt/valid": "owt_valid.txt", "corpus/train": "corpus.en", "corpus/valid": "corpus.en"}


def main(dataset: str, split: str, input_file: str | None = None, output: str | None = None, tokenizer_dir: str = "data/tokenizer"):
    from .tokenizer import Tokenizer
    if input_file is None:
        input_file = "tests/fixtures/corpus.en" if dataset == "corpus" else "/data/" + FNAME[dataset + "/" + split]
    tok = Tokenizer.from_files(os.path.join(tokenizer_dir, dataset + "-vocab.pkl"), os.path.join(tokenizer_dir, dataset + "-merges.pkl"),
                               special_tokens=["<|endoftext|>"])
    output = output or os.path.join(tokenizer_dir, "%s-tokens-%s.bin" % (dataset, split))
    n = encode_file(tok, input_file, output, np.uint16, save_pt=os.path.splitext(output)[0] + ".pt")
    print("%d tokens -> %s" % (n, output))


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", type=str, required=True)
    ap.add_argument("--split", type=str, required=True)
    ap.add_argument("--input", type=str, default=None)
    ap.add_argument("--output", type=str, default=None)
    ap.add_argument("--tokenizer-dir", type=str, default="data/tokenizer")
    a = ap.parse_args()
    main(a.dataset, a.split, a.input, a.output, a.tokenizer_dir)
