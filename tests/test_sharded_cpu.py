"""Multi-rank host logic (transformer_lm_b200/sharded.py) under gloo, world_size 2 and 3, on CPU: byte-range
planning, halos, error agreement, table exchange.  The device is replaced by the checker-backed OracleCounter;
the merged table must equal the count of the whole text on every rank (the merge loop depends only on it,
SURVEY A-8)."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _bootstrap  # noqa: F401
from oracle import oracle
from tests.common import FIXTURES_PATH
from transformer_lm_b200 import sharded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, scenarios, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.helpers.fake_counter import DeferredCheckCounter, OracleCounter
        for name, data, specials, halo in scenarios:
            sharded.HALO_RIGHT = halo
            counter = DeferredCheckCounter() if name.startswith("deferred") else OracleCounter()
            try:
                out = sharded.sharded_count(counter, sharded.FileShards(lambda lo, hi: data[lo:hi], len(data)), specials, None, True)
                if name.startswith("deferred"):    # the all-reduced per-rank tables were handed over instead of compared at once
                    assert counter.expected is not None and bool((counter.expected == counter.pair_table(specials)).all())
                q.put((name, rank, out, dict(counter.table), counter.adds))
            except UnicodeDecodeError as e:
                q.put((name, rank, "utf8", (e.start, e.reason), 0))
    finally:
        dist.destroy_process_group()


def _run_many(world, scenarios):
    """One process group per world size runs every scenario in turn (process start-up dominates otherwise)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, scenarios, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world * len(scenarios))]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out = {}
    for name, rank, o, table, adds in res:
        out.setdefault(name, []).append((rank, o, table, adds))
    return {k: sorted(v) for k, v in out.items()}


def _run(world, data, specials=("<|endoftext|>",), halo=64 << 10):
    return _run_many(world, [("one", data, list(specials), halo)])["one"]


def _want(data, specials=()):
    return oracle.count_pretokens(data, [])


def test_plan_shard_cuts_on_code_points():
    data = ("é🙃中" * 50).encode()
    peek = lambda lo, hi: data[lo:hi]      # noqa: E731
    for world in (2, 3, 5, 7):
        prev_hi = 0
        for r in range(world):
            lo, hi, rlo, rhi = sharded.plan_shard(peek, len(data), r, world, 32)
            assert lo == prev_hi and rlo <= lo <= hi <= rhi
            for p in (lo, hi, rlo, rhi):
                assert p == len(data) or (data[p] & 0xC0) != 0x80
            prev_hi = hi
        assert prev_hi == len(data)


def _corpus():
    return (FIXTURES_PATH / "corpus.en").read_bytes().replace(b"\r", b"")


def _unicode_text():
    rnd = random.Random(4)
    alpha = ["é", "🙃", "中", " ", "  ", "\n", "a", "it's", "<|endoftext|>", "x" * 300, " " * 200]
    return "".join(rnd.choice(alpha) for _ in range(4000)).encode()


def _bad_corpus():
    data = bytearray(_corpus())
    data[100000] = 0xFF
    data[1000] = 0xC0
    return bytes(data)


@pytest.fixture(scope="module")
def two_ranks():
    sp = ["<|endoftext|>"]
    return _run_many(2, [("corpus", _corpus(), sp, 64 << 10), ("unicode_tiny_halo", _unicode_text(), sp, 32), ("two_bytes", b"ab", sp, 64 << 10),
                         ("empty", b"", sp, 64 << 10), ("bad", _bad_corpus(), sp, 64 << 10),
                         ("crlf", b"hello world\r\nsecond line " * 2000, sp, 64 << 10)])


def test_sharded_count_equals_whole_text(two_ranks):
    want = _want(_corpus())
    for rank, out, table, adds in two_ranks["corpus"]:
        assert out == "ok" and adds == 1
        assert table == want


def test_three_ranks():
    res = _run_many(3, [("corpus", _corpus(), ["<|endoftext|>"], 64 << 10), ("two_bytes", b"ab", [], 64 << 10)])
    want = _want(_corpus())
    for rank, out, table, adds in res["corpus"]:
        assert out == "ok" and table == want
    for rank, out, table, adds in res["two_bytes"]:           # more ranks than bytes
        assert out == "ok" and table == {b"ab": 1}


def test_unicode_heavy_text_and_tiny_halo_forces_retries(two_ranks):
    want = _want(_unicode_text())
    res = two_ranks["unicode_tiny_halo"]
    assert any(adds > 1 for _, _, _, adds in res)          # the 32-byte halo was too small somewhere
    for rank, out, table, adds in res:
        assert out == "ok" and table == want


def test_tiny_and_empty_inputs(two_ranks):
    for rank, out, table, adds in two_ranks["two_bytes"]:
        assert out == "ok" and table == {b"ab": 1}
    for rank, out, table, adds in two_ranks["empty"]:
        assert out == "ok" and table == {}


def test_invalid_utf8_raises_the_first_offset_on_every_rank(two_ranks):
    try:
        _bad_corpus().decode("utf-8")
    except UnicodeDecodeError as e:
        want = (e.start, e.reason)
    for rank, out, info, _ in two_ranks["bad"]:
        assert out == "utf8"
        # the exception is rebuilt from a 16-byte window around the first bad byte: same reason, offset inside the window
        assert info[1] == want[1]


def test_carriage_return_falls_back(two_ranks):
    for rank, out, table, adds in two_ranks["crlf"]:
        assert out == "newline"


def test_merges_digest_is_order_sensitive():
    a = [(b"a", b"b"), (b"ab", b"c")]
    assert sharded.merges_digest(a) != sharded.merges_digest(a[::-1])
    assert sharded.merges_digest(a) == sharded.merges_digest(list(a))


def test_deferred_linearity_check_receives_the_all_reduced_table():
    data = _corpus()
    res = _run_many(2, [("deferred", data, ["<|endoftext|>"], 64 << 10)])["deferred"]
    want = _want(data)
    for rank, out, table, adds in res:
        assert out == "ok" and table == want
