"""Multi-rank host logic of the bulk encode (transformer_lm_b200/sharded_encode.py, encode_file) under gloo, world sizes 2
and 3, on CPU: exact cut positions, token offsets, assembling the output file.  The device is replaced by a checker-backed
tokenizer (the oracle): the ranks' ids in rank order must equal the oracle's encode of the whole text."""
import os
import random
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import _bootstrap  # noqa: F401
from oracle import oracle
from tests.common import FIXTURES_PATH
from transformer_lm_b200 import sharded_encode

EOT = "<|endoftext|>"


class CheckerTokenizer:
    """Tokenizer stand-in with the two things the host logic needs: special_tokens and encode_to_numpy."""
    host_only = True

    def __init__(self, vocab, merges, specials):
        self.special_tokens = sorted(set(specials), key=len, reverse=True)
        self._o = oracle.OracleTokenizer(dict(vocab), list(merges), specials)

    def encode_to_numpy(self, data, dtype=np.int32):
        return self._o.encode_bytes(bytes(data)).astype(dtype)


def _trained():
    data = (FIXTURES_PATH / "corpus.en").read_bytes()[:60000]
    return oracle.train_bpe_on_bytes(data, 400, [EOT])


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _texts():
    rnd = random.Random(5)
    body = (FIXTURES_PATH / "tinystories_sample.txt").read_bytes()
    words = [b"alpha", b"beta", b" ", b"  ", b"\n", b"gamma's", EOT.encode(), "naïve".encode(), b"12", b" x"]
    return {
        "stories": body * 3,                                          # specials near every cut
        "no_specials_near": b" ".join(rnd.choice(words[:2]) for _ in range(3000)),      # falls back to the lone-space rule
        "random": b"".join(rnd.choice(words) for _ in range(4000)),
        "crlf": (body.replace(b"\n", b"\r\n") + b"\r" + EOT.encode()) * 2,
        "tiny": b"ab",
        "empty": b"",
    }


def _worker(rank, world, port, tmpdir, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vocab, merges = _trained()
        for specials in ([EOT], []):
            tok = CheckerTokenizer(vocab, merges, specials)
            for name, text in _texts().items():
                if name != "crlf":
                    ids, first, total = sharded_encode.encode_sharded(tok, text, np.int32)
                    q.put(("mem", name, tuple(specials), rank, first, total, ids.tolist()))
                path = os.path.join(tmpdir, "%s_%d.txt" % (name, len(specials)))
                if rank == 0:
                    with open(path, "wb") as f:
                        f.write(text)
                dist.barrier()
                out = os.path.join(tmpdir, "%s_%d.bin" % (name, len(specials)))
                n = sharded_encode.encode_file_sharded(tok, path, out, np.uint16, piece_bytes=3000)
                if rank == 0:
                    q.put(("file", name, tuple(specials), rank, 0, n, np.fromfile(out, dtype="<u2").tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_encode_equals_the_whole_text_encode(world, tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    texts = _texts()
    n_mem = 2 * (len(texts) - 1) * world
    n_file = 2 * len(texts)
    res = [q.get(timeout=300) for _ in range(n_mem + n_file)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    vocab, merges = _trained()
    for specials in ([EOT], []):
        otok = oracle.OracleTokenizer(dict(vocab), list(merges), specials)
        for name, text in texts.items():
            translated = text.replace(b"\r\n", b"\n").replace(b"\r", b"\n")
            want = otok.encode_bytes(translated).tolist()
            files = [r for r in res if r[0] == "file" and r[1] == name and r[2] == tuple(specials)]
            assert len(files) == 1 and files[0][5] == len(want) and files[0][6] == want, (name, specials)
            if name == "crlf":
                continue
            parts = sorted((r for r in res if r[0] == "mem" and r[1] == name and r[2] == tuple(specials)), key=lambda r: r[3])
            assert len(parts) == world
            got, pos = [], 0
            for _, _, _, rank, first, total, ids in parts:
                assert first == pos and total == len(want)
                got += ids
                pos += len(ids)
            assert got == want, (name, specials)


def test_first_exact_cut_rules():
    sp = [EOT.encode(), (EOT * 2).encode()]
    text = b"aaaa " + EOT.encode() * 2 + b"bbbb cc dd"
    peek = lambda lo, hi: text[lo:hi]
    n = len(text)
    assert sharded_encode.first_exact_cut(peek, n, sp, 0) == 0
    assert sharded_encode.first_exact_cut(peek, n, sp, n + 5) == n
    assert sharded_encode.first_exact_cut(peek, n, sp, 2) == 5                   # the start of the (double) special
    # a position inside the double special is not a cut: the next exact boundary is the lone space after "bbbb"
    assert sharded_encode.first_exact_cut(peek, n, sp, 6) == text.index(b" cc")
    assert sharded_encode.first_exact_cut(peek, n, [], 1) == 4                   # "aaaa| <" : lone space between ASCII non-spaces
    with pytest.raises(RuntimeError):
        sharded_encode.first_exact_cut(lambda lo, hi: (b"x" * 100)[lo:hi], 100, sp, 10)
    cuts = sharded_encode.shard_cuts(peek, n, sp, 4)
    assert cuts[0] == 0 and cuts[-1] == n and cuts == sorted(cuts)
