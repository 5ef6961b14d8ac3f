"""Host logic of the streamed ingest behind train_bpe(path) (transformer_lm_b200/train.py:_train_bpe_streamed) on
CPU: chunk cuts on code-point boundaries, halos, the ring of chunk buffers filled by reader threads, block-aligned O_DIRECT reads and
their skew, error offsets, and the fall-back signals.  numpy arrays stand in for the page-locked buffers and the checker-backed
counter (the oracle) for the device: the merged table must equal the oracle's count of the whole file."""
import os

import numpy as np
import pytest

import _bootstrap  # noqa: F401
from oracle import oracle
from tests.common import FIXTURES_PATH
from tests.helpers.fake_counter import OracleCounter
from transformer_lm_b200 import sharded, train


class AlignedBuffer:
    """Page-aligned like cudaMallocHost memory (O_DIRECT wants aligned addresses)."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        raw = np.zeros(self.nbytes + 4096, dtype=np.uint8)
        skip = -raw.ctypes.data % 4096
        self._raw, self.array = raw, raw[skip: skip + self.nbytes]

    def free(self):
        self.array = self._raw = None


class TableCounter(OracleCounter):
    chunks = 0

    def __init__(self, ctx=None):
        super().__init__()

    def add(self, *a, **k):
        TableCounter.chunks += 1
        return super().add(*a, **k)

    def finish(self, vocab_size, special_tokens, return_stats=False):
        return dict(self.table)


@pytest.fixture()
def small_chunks(monkeypatch):
    train.release_buffers()
    monkeypatch.setattr(train._lib, "PinnedBuffer", AlignedBuffer)
    monkeypatch.setattr(sharded, "DeviceCounter", TableCounter)
    monkeypatch.setattr(train, "_STREAM_CHUNK", 8192)
    monkeypatch.setattr(train, "_STREAM_HALO_RIGHT", 512)
    monkeypatch.setattr(train, "_READ_THREADS", 3)
    TableCounter.chunks = 0
    yield
    train.release_buffers()


def _corpus():
    body = (FIXTURES_PATH / "corpus.en").read_bytes().replace(b"\r", b"")[:90000]
    return body + "é🙃中 naïve ".encode() * 400 + body[:7777]


@pytest.mark.parametrize("direct", [False, True])
def test_streamed_table_equals_the_count_of_the_whole_file(tmp_path, small_chunks, direct):
    data = _corpus()
    path = tmp_path / "c.txt"                                      # (tmp_path is on a disk file system here; tmpfs refuses O_DIRECT
    path.write_bytes(data)                                         #  and _open_for_read then reads buffered: both are fine)
    table = train._train_bpe_streamed(path, len(data), 1000, ["<|endoftext|>"], direct_io=direct)
    assert table == oracle.count_pretokens(data, [])
    assert TableCounter.chunks >= len(data) // 8192 - 1            # really chunked
    # a second call reuses the chunk buffers it kept
    kept = [id(b) for b in train._PIN_CACHE]
    assert train._train_bpe_streamed(path, len(data), 1000, [], direct_io=direct) == table
    assert [id(b) for b in train._PIN_CACHE] == kept


def test_streamed_reports_the_reference_decode_error_and_falls_back_on_cr(tmp_path, small_chunks):
    data = bytearray(_corpus())
    bad = 50001
    while (data[bad] & 0xC0) == 0x80 or data[bad] >= 0x80:
        bad += 1
    data[bad] = 0xFF
    path = tmp_path / "bad.txt"
    path.write_bytes(bytes(data))
    with pytest.raises(UnicodeDecodeError) as e:
        train._train_bpe_streamed(path, len(data), 1000, [])
    with pytest.raises(UnicodeDecodeError) as want:
        bytes(data).decode("utf-8")
    assert e.value.reason == want.value.reason and e.value.object[e.value.start] == 0xFF
    data[bad] = 0x0D                                               # a carriage return: the caller must take the one-piece path
    path.write_bytes(bytes(data))
    assert train._train_bpe_streamed(path, len(data), 1000, []) is None
    long_word = b"x" * 5000                                        # a pretoken longer than the halo: one-piece path as well
    path.write_bytes(_corpus()[:8000] + long_word + _corpus()[:20000])
    assert train._train_bpe_streamed(path, os.path.getsize(path), 1000, []) is None
