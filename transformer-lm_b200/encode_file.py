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

import errno
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


def encode_file(tokenizer, input_path, output_path, dtype=np.uint16, piece_bytes: int = 64 << 20, save_pt: str | None = None,
                byte_range: tuple[int, int] | None = None, distributed: bool = False, group=None) -> int:
    """Encode `input_path` with `tokenizer` into raw little-endian `dtype` ids at `output_path`.  Returns the token count.

    Three stages run concurrently: a reader thread pulls the next block from the file straight into page-locked
    memory, the calling thread cuts and encodes the current piece (bpe_encode: upload, kernels, download into page-locked memory), a
    writer thread appends the ids of the previous piece to the output.  byte_range = (lo, hi) restricts the work to those bytes of
    the file (lo and hi must be exact cut positions: sharded_encode.py); distributed=True shards the file over the ranks of the
    initialised torch.distributed group."""
    if distributed:
        from .sharded_encode import encode_file_sharded
        return encode_file_sharded(tokenizer, input_path, output_path, dtype, piece_bytes, group)
    if getattr(tokenizer, "host_only", False):               # (checker-backed tokenizers of the CPU tests: no page-locked staging)
        total = _encode_file_simple(tokenizer, input_path, output_path, dtype, piece_bytes, byte_range)
    else:
        total = _encode_file_pinned(tokenizer, input_path, output_path, dtype, piece_bytes, byte_range)
    if save_pt:
        import torch
        torch.save(np.fromfile(output_path, dtype=np.dtype(dtype)), save_pt, pickle_protocol=4)       # models/tokenizer/encode.py:37-38
    return total


_IO_THREADS = int(os.environ.get("BPE_IO_THREADS", "8"))
_PIN_CACHE: dict = {}            # page-locked staging buffers, kept between calls (pinning costs ~0.3 s per GB)


def _pinned(role: str, nbytes: int):
    b = _PIN_CACHE.get(role)
    if b is None or b.nbytes < nbytes:
        if b is not None:
            b.free()
        b = _PIN_CACHE[role] = _lib.PinnedBuffer(nbytes)
    return b


def release_buffers() -> None:
    """Free the page-locked staging buffers encode_file keeps between calls."""
    for b in _PIN_CACHE.values():
        b.free()
    _PIN_CACHE.clear()


def _find_cut(mv: memoryview, n: int, specials: list[bytes]) -> int:
    """Largest exact cut position in the second half of mv[:n] (0 = none); looks at the last few MB first."""
    lo = n // 2
    ctx = max([len(t) for t in specials] + [2])
    for window in (4 << 20, n - lo):
        start = max(lo, n - window)
        c0 = max(0, start - ctx)
        cut = _last_exact_cut(bytes(mv[c0:n]), specials, start - c0)
        if cut:
            return c0 + cut
        if start == lo:
            break
    return 0


def _encode_file_pinned(tokenizer, input_path, output_path, dtype, piece_bytes, byte_range) -> int:
    """The device path of encode_file: no intermediate Python byte strings.  Two page-locked input buffers (the reader thread's
    readinto target; a piece = [bytes carried over from the previous block][new block] up to the last exact cut) and two page-locked
    output buffers (bpe_encode downloads into them, the writer thread writes them out)."""
    import queue
    import threading
    dtype = np.dtype(dtype)
    specials = [s.encode("utf-8") for s in tokenizer.special_tokens]
    lo, hi = byte_range if byte_range is not None else (0, os.path.getsize(input_path))
    span = max(hi - lo, 0)
    P = max(1, min(int(piece_bytes), max(span, 1)))
    in_cap = 2 * P + 64
    inb = [_pinned("in%d" % i, in_cap).array for i in range(2)]
    outb = [_pinned("out%d" % i, in_cap * dtype.itemsize).array for i in range(2)]
    out_views = [b[: in_cap * dtype.itemsize].view(dtype) for b in outb]
    requests: queue.Queue = queue.Queue()
    filled: queue.Queue = queue.Queue()
    results: queue.Queue = queue.Queue()
    free_out: queue.Queue = queue.Queue()
    for i in range(2):
        free_out.put(i)
    failure: list = []

    # A block is read, and a piece's ids are written, by _IO_THREADS threads at once (positional reads / writes of slices: one
    # thread copies ~4 GB/s between the page cache and user memory, which is less than the GPU encodes and the PCIe link carries).
    import concurrent.futures as cf
    io_pool = cf.ThreadPoolExecutor(max_workers=_IO_THREADS)

    def pread_slice(fd, mv, pos):
        got = 0
        while got < len(mv):
            k = os.preadv(fd, [mv[got:]], pos + got)
            if k <= 0:
                break
            got += k
        return got

    def pwrite_slice(fd, mv, pos):
        put = 0
        while put < len(mv):
            put += os.pwrite(fd, mv[put:], pos + put)

    def slices(n):
        step = max(1 << 20, -(-n // _IO_THREADS))
        return [(o, min(o + step, n)) for o in range(0, n, step)]

    def reader():
        try:
            fd = os.open(os.fspath(input_path), os.O_RDONLY)
            try:
                pos = lo
                while True:
                    req = requests.get()
                    if req is None:
                        return
                    idx, off, want = req
                    mv = memoryview(inb[idx])[off: off + want]
                    parts = [(a, b, io_pool.submit(pread_slice, fd, mv[a:b], pos + a)) for a, b in slices(want)]
                    got = 0
                    for a, b, fut in parts:                  # (a short slice can only be the one that met the end of the file)
                        k = fut.result()
                        if got == a:
                            got += k
                    pos += got
                    filled.put(got)
            finally:
                os.close(fd)
        except BaseException as e:            # surfaced in the calling thread
            failure.append(e)
            filled.put(0)

    def writer():
        # write() calls on ONE file serialise on its inode lock however many threads issue them (measured: 3.7 GB/s into tmpfs with
        # one or four threads, and the whole pipeline waited for it).  The output is therefore a sparse file of the largest possible
        # size (a token covers at least one byte), mapped into memory, filled by parallel copies and cut to its real size at the end;
        # where that is not possible (no mmap on the file system) positional writes do the job.
        import mmap
        try:
            fd = os.open(os.fspath(output_path), os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o666)
            mm = out_arr = None
            try:
                cap = max(span, 1) * dtype.itemsize
                try:
                    os.ftruncate(fd, cap)
                    mm = mmap.mmap(fd, cap)
                    out_arr = np.frombuffer(mm, dtype=np.uint8)
                except (OSError, ValueError):
                    mm = out_arr = None
                    os.ftruncate(fd, 0)
                pos = 0
                while True:
                    item = results.get()
                    if item is None:
                        break
                    idx, n_tok, ids = item
                    arr = ids.astype(dtype.newbyteorder("<"), copy=False) if ids is not None else out_views[idx][:n_tok]   # (little-endian host: the raw buffer is the file format)
                    src = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
                    if out_arr is not None:
                        # a store into a mapped hole of a full file system is a SIGBUS, not an OSError: look before writing
                        try:
                            vfs = os.fstatvfs(fd)
                            room = vfs.f_bavail * vfs.f_frsize if vfs.f_blocks else None
                        except OSError:
                            room = None
                        if room is not None and room < src.size:
                            raise OSError(errno.ENOSPC, "no room for %d more bytes of token ids" % src.size, os.fspath(output_path))
                        futs = [io_pool.submit(np.copyto, out_arr[pos + a: pos + b], src[a:b]) for a, b in slices(src.size)]
                    else:
                        mv = memoryview(src)
                        futs = [io_pool.submit(pwrite_slice, fd, mv[a:b], pos + a) for a, b in slices(src.size)]
                    for fut in futs:
                        fut.result()
                    pos += src.size
                    if ids is None:
                        free_out.put(idx)
                if mm is not None:
                    out_arr = None
                    mm.close()
                    mm = None
                    os.ftruncate(fd, pos)
            finally:
                if mm is not None:
                    out_arr = None
                    mm.close()
                os.close(fd)
        except BaseException as e:
            failure.append(e)
            while results.get() is not None:
                pass

    rt, wt = threading.Thread(target=reader, daemon=True), threading.Thread(target=writer, daemon=True)
    rt.start(); wt.start()
    total = 0
    offset = lo                                  # bytes of the (untranslated) file consumed so far: for error offsets
    rest_from = None                             # no exact cut in a whole piece: the rest of the file goes through the simple path
    try:
        k, carry_len, left = 0, 0, span
        want = min(P, left)
        requests.put((0, 0, want))
        prof = os.environ.get("BPE_ENC_PROFILE") is not None
        import time as _t
        acc = {"wait_read": 0.0, "cut": 0.0, "wait_out": 0.0, "encode": 0.0, "pieces": 0}
        while True:
            t0 = _t.perf_counter()
            got = filled.get()
            acc["wait_read"] += _t.perf_counter() - t0
            if failure:
                raise failure[0]
            cur = inb[k % 2]
            n = carry_len + got
            left -= got
            eof = left <= 0 or got < want
            hold = 1 if (not eof and n and cur[n - 1] == 13) else 0     # the "\n" of a "\r\n" may start the next block
            n_eff = n - hold
            t0 = _t.perf_counter()
            if eof:
                cut = n_eff
            else:
                cut = _find_cut(memoryview(cur), n_eff, specials)
                if cut == 0:
                    rest_from = offset
                    break
            next_carry = n - cut
            if not eof:
                inb[(k + 1) % 2][:next_carry] = cur[cut:n]
                want = min(P, left)
                requests.put(((k + 1) % 2, next_carry, want))
            acc["cut"] += _t.perf_counter() - t0
            if cut:
                try:
                    t0 = _t.perf_counter()
                    while True:                              # (a writer that failed hands no buffer back: do not wait for one forever)
                        try:
                            idx = free_out.get(timeout=0.2)
                            break
                        except queue.Empty:
                            if failure:
                                raise failure[0]
                    t1 = _t.perf_counter()
                    n_tok = tokenizer.encode_into(cur[:cut], out_views[idx])
                    acc["wait_out"] += t1 - t0; acc["encode"] += _t.perf_counter() - t1; acc["pieces"] += 1
                    if tokenizer.last_saw_cr:                # text-mode read of the reference (A-2): the device saw a carriage return
                        free_out.put(idx)                    # (rare: translate on the host and encode the piece again)
                        piece = bytes(memoryview(cur)[:cut]).replace(b"\r\n", b"\n").replace(b"\r", b"\n")
                        ids = tokenizer.encode_to_numpy(piece, dtype)
                        results.put((-1, ids.size, ids))
                        n_tok = ids.size
                    else:
                        results.put((idx, n_tok, None))
                except UnicodeDecodeError as e:
                    raise UnicodeDecodeError(e.encoding, e.object[max(e.start - 8, 0): e.end + 8], min(e.start, 8), min(e.start, 8) + (e.end - e.start),
                                             e.reason + " (near byte %d of the file)" % (offset + e.start)) from None
                if failure:
                    raise failure[0]
                total += n_tok
                offset += cut
            if eof:
                break
            carry_len = next_carry
            k += 1
    finally:
        requests.put(None)
        results.put(None)
        t0 = _t.perf_counter()
        wt.join()
        rt.join(timeout=5)
        io_pool.shutdown(wait=True)
        if prof:
            import sys
            acc["drain"] = _t.perf_counter() - t0
            print("  [encode_file: %s]" % ", ".join("%s %.1f ms" % (k_, v * 1e3) if k_ != "pieces" else "pieces %d" % v for k_, v in acc.items()), file=sys.stderr)
    if failure:
        raise failure[0]
    if rest_from is not None:
        part = os.fspath(output_path) + ".rest"
        try:
            total += _encode_file_simple(tokenizer, input_path, part, dtype, max(piece_bytes, 1 << 20), (rest_from, hi))
            with open(output_path, "ab") as out, open(part, "rb") as src:
                while True:
                    blk = src.read(64 << 20)
                    if not blk:
                        break
                    out.write(blk)
        finally:
            if os.path.exists(part):
                os.remove(part)
    return total


def _encode_file_simple(tokenizer, input_path, output_path, dtype, piece_bytes, byte_range) -> int:
    """The plain implementation (Python byte strings): checker-backed tokenizers of the CPU tests, and files in which a whole piece
    holds no exact cut."""
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
