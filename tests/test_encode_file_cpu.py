"""Host logic of the bulk encode driver's device path (transformer_lm_b200/encode_file.py:_encode_file_pinned) on CPU: reader thread,
carry-over between blocks, exact cuts, the carriage-return path, the writer thread and its failure handling.  The page-locked buffers
are replaced by numpy arrays and the device by a checker-backed tokenizer (the oracle); the .bin must hold the oracle's ids of the
whole text read in text mode (what models/tokenizer/encode.py:18-38 of the reference encodes)."""
import errno
import os
import random

import numpy as np
import pytest

import _bootstrap  # noqa: F401
from oracle import oracle
from tests.common import FIXTURES_PATH
from transformer_lm_b200 import encode_file as ef

EOT = "<|endoftext|>"


class FakePinned:
    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.array = np.zeros(self.nbytes, dtype=np.uint8)

    def free(self):
        self.array = None


class PipelineTokenizer:
    """What _encode_file_pinned needs of a Tokenizer: special_tokens, encode_into + last_saw_cr, encode_to_numpy."""

    def __init__(self, vocab, merges, specials):
        self.special_tokens = sorted(set(specials), key=len, reverse=True)
        self._o = oracle.OracleTokenizer(dict(vocab), list(merges), specials)
        self.pieces = 0

    def encode_to_numpy(self, data, dtype=np.int32):
        return self._o.encode_bytes(bytes(data)).astype(dtype)

    def encode_into(self, data, out):
        raw = bytes(data)
        self.last_saw_cr = b"\r" in raw                       # (bpe_tok_saw_cr: the flags kernel sees every byte)
        ids = self._o.encode_bytes(raw)
        out[: ids.size] = ids
        self.pieces += 1
        return int(ids.size)


@pytest.fixture()
def fake_pinned(monkeypatch):
    ef.release_buffers()
    monkeypatch.setattr(ef._lib, "PinnedBuffer", FakePinned)
    yield
    ef.release_buffers()


def _trained():
    return oracle.train_bpe_on_bytes((FIXTURES_PATH / "corpus.en").read_bytes()[:60000], 400, [EOT])


def _texts():
    rnd = random.Random(11)
    body = (FIXTURES_PATH / "tinystories_sample.txt").read_bytes()
    words = [b"alpha", b"beta", b" ", b"  ", b"\n", b"gamma's", EOT.encode(), "naïve".encode(), b"12", b" x"]
    return {
        "stories": body * 3,
        "lone_spaces": b" ".join(rnd.choice(words[:2]) for _ in range(3000)),
        "random": b"".join(rnd.choice(words) for _ in range(4000)),
        "crlf": (body.replace(b"\n", b"\r\n") + b"\r" + EOT.encode()) * 2,
        "cr_at_block_end": (b"ab cd\r\n" * 700),
        "no_cut": b"x" * 20000 + b" tail",                     # a whole piece without an exact cut: the rest takes the simple path
        "tiny": b"ab",
        "empty": b"",
    }


@pytest.mark.parametrize("specials", [[EOT], []])
def test_pinned_pipeline_equals_the_oracle_on_the_text_mode_read(tmp_path, fake_pinned, specials):
    vocab, merges = _trained()
    for name, text in _texts().items():
        tok = PipelineTokenizer(vocab, merges, specials)
        src, dst = tmp_path / (name + ".txt"), tmp_path / (name + ".bin")
        src.write_bytes(text)
        for piece, dtype in ((3000, np.uint16), (1 << 20, np.int32)):
            n = ef.encode_file(tok, src, dst, dtype, piece_bytes=piece)
            want = tok._o.encode_bytes(text.replace(b"\r\n", b"\n").replace(b"\r", b"\n"))
            got = np.fromfile(dst, dtype=dtype)
            assert n == want.size == got.size and (got == want).all(), (name, piece)
            assert os.path.getsize(dst) == want.size * np.dtype(dtype).itemsize, name      # the sparse file is cut to its real size
        if name == "stories":
            assert tok.pieces > 3                                  # really streamed in pieces


def test_byte_range_encodes_only_that_range(tmp_path, fake_pinned):
    vocab, merges = _trained()
    tok = PipelineTokenizer(vocab, merges, [EOT])
    text = (FIXTURES_PATH / "tinystories_sample.txt").read_bytes() * 2
    lo, hi = text.find(EOT.encode(), 5000), text.rfind(EOT.encode())
    src, dst = tmp_path / "t.txt", tmp_path / "t.bin"
    src.write_bytes(text)
    n = ef.encode_file(tok, src, dst, np.uint16, piece_bytes=4096, byte_range=(lo, hi))
    want = tok._o.encode_bytes(text[lo:hi])
    assert n == want.size and (np.fromfile(dst, dtype=np.uint16) == want).all()


def test_writer_failures_reach_the_caller(tmp_path, fake_pinned, monkeypatch):
    vocab, merges = _trained()
    tok = PipelineTokenizer(vocab, merges, [EOT])
    src = tmp_path / "t.txt"
    src.write_bytes((FIXTURES_PATH / "tinystories_sample.txt").read_bytes() * 3)
    with pytest.raises(OSError):                                   # the output cannot be created
        ef.encode_file(tok, src, tmp_path / "missing_dir" / "t.bin", np.uint16, piece_bytes=3000)

    import mmap
    probe = tmp_path / "probe"
    probe.write_bytes(b"\0" * 4096)
    try:                                                           # (the check guards the memory-mapped output; positional writes fail by themselves)
        with open(probe, "r+b") as f:
            mmap.mmap(f.fileno(), 4096).close()
    except (OSError, ValueError):
        return

    class Full:
        f_bavail, f_frsize, f_blocks = 0, 4096, 1000

    monkeypatch.setattr(ef.os, "fstatvfs", lambda fd: Full)      # a full file system: an error, not a SIGBUS in the mapped output
    with pytest.raises(OSError) as e:
        ef.encode_file(tok, src, tmp_path / "t.bin", np.uint16, piece_bytes=3000)
    assert e.value.errno == errno.ENOSPC


def test_invalid_utf8_reports_the_file_offset(tmp_path, fake_pinned):
    vocab, merges = _trained()
    tok = PipelineTokenizer(vocab, merges, [EOT])

    def strict(data, out, _inner=tok.encode_into):
        bytes(data).decode("utf-8")                                # (the device raises the reference's UnicodeDecodeError)
        return _inner(data, out)
    tok.encode_into = strict
    body = (FIXTURES_PATH / "tinystories_sample.txt").read_bytes()
    bad = len(body) + 1234
    text = bytearray(body * 2)
    text[bad] = 0xFF
    src = tmp_path / "bad.txt"
    src.write_bytes(bytes(text))
    with pytest.raises(UnicodeDecodeError) as e:
        ef.encode_file(tok, src, tmp_path / "bad.bin", np.uint16, piece_bytes=3000)
    assert "near byte %d of the file" % bad in e.value.reason
