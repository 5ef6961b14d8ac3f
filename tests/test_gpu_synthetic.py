"""Parity on the bench's own input distributions: synthetic TinyStories- and OWT-shaped text (Unicode
whitespace, CJK, emoji, URLs, contractions, documents separated by <|endoftext|>) against the CPU oracle at sizes it
finishes in seconds, and size-independent properties at larger sizes (encode -> decode round trip, token-count additivity)."""
import numpy as np
import pytest

import _bootstrap  # noqa: F401
from oracle import oracle
from tests.adapters import get_tokenizer
from transformer_lm_b200.synth import synth_host

pytestmark = pytest.mark.gpu
EOT = "<|endoftext|>"


def _train(data, vocab_size, specials, **kw):
    from models.tokenizer.train import train_bpe_on_bytes
    return train_bpe_on_bytes(data, vocab_size, specials, **kw)


@pytest.mark.parametrize("shape,seed,mb,vocab", [("tinystories", 1234, 24, 700), ("owt", 4321, 12, 600)])
def test_train_on_synthetic_shapes_matches_oracle(shape, seed, mb, vocab):
    data = synth_host(shape, seed, mb << 20).tobytes()
    want = oracle.train_bpe_on_bytes(data, vocab, [EOT])
    got = _train(data, vocab, [EOT])
    assert got[1] == want[1] and got[0] == want[0]


@pytest.mark.parametrize("shape,seed", [("owt", 4321), ("tinystories", 1234)])
def test_train_64MB_slice_vocab_1000_matches_oracle(shape, seed):
    """the 64 MB prefix of each bench corpus (same generator and seed) at vocab 1000, exact vocab and merges
    against the oracle (about 10 s of oracle time per shape)."""
    data = synth_host(shape, seed, 64 << 20).tobytes()
    want = oracle.train_bpe_on_bytes(data, 1000, [EOT])
    got = _train(data, 1000, [EOT])
    assert got[1] == want[1] and got[0] == want[0]


@pytest.mark.parametrize("grid", ["2", "3", "17"])
def test_merge_loop_result_does_not_depend_on_the_grid(monkeypatch, grid):
    # Small grids put many 64-block chunks on every warp: more than the shared-memory cache of block maxima holds
    # (MG_CACHE_ITERS) and, in the first step, more dirty blocks than a CTA's list holds (MG_LIST_CAP) -- the overflow
    # paths of k_merge_loop.  The result must equal the default grid's (which the oracle pins at smaller sizes) and, on
    # a slice the oracle finishes, the oracle's.
    data = synth_host("owt", 4321, 64 << 20)
    want = _train(data, 2500, [EOT])
    small = synth_host("owt", 4321, 12 << 20).tobytes()
    want_small = oracle.train_bpe_on_bytes(small, 600, [EOT])
    monkeypatch.setenv("BPE_MERGE_G", grid)
    got = _train(data, 2500, [EOT])
    assert got[1] == want[1] and got[0] == want[0]
    got_small = _train(small, 600, [EOT])
    assert got_small[1] == want_small[1] and got_small[0] == want_small[0]


@pytest.fixture(scope="module")
def owt_tokenizer():
    data = synth_host("owt", 4321, 64 << 20)
    vocab, merges = _train(data, 8000, [EOT])
    return vocab, merges


def test_encode_synthetic_owt_matches_oracle(owt_tokenizer):
    vocab, merges = owt_tokenizer
    tok = get_tokenizer(dict(vocab), list(merges), [EOT])
    otok = oracle.OracleTokenizer(dict(vocab), list(merges), [EOT])
    data = synth_host("owt", 4322, 6 << 20).tobytes()
    got = tok.encode_to_numpy(data, np.int32)
    want = otok.encode_bytes(data)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_encode_16MB_with_the_32k_vocab_matches_oracle():
    """a 16 MB slice of the bench's encode text with a 32 000-entry vocab trained on the GPU on 1 GB of the
    bench's train corpus (text generated in HBM), ids bit-exact against the oracle's Tokenizer.encode."""
    import torch
    from transformer_lm_b200 import _lib
    from transformer_lm_b200.synth import synth_device
    ctx = _lib.default_context()
    n = (1 << 30) // 4096 * 4096
    t = torch.empty(n, dtype=torch.uint8, device="cuda")
    synth_device("owt", 4321, n, t.data_ptr(), ctx=ctx)
    vocab, merges = _train(None, 32000, [EOT], device_ptr=t.data_ptr(), n_bytes=n)
    del t
    assert len(merges) == 32000 - 257
    tok = get_tokenizer(dict(vocab), list(merges), [EOT])
    otok = oracle.OracleTokenizer(dict(vocab), list(merges), [EOT])
    data = synth_host("owt", 4322, 16 << 20).tobytes()
    got = tok.encode_to_numpy(data, np.uint16)
    want = otok.encode_bytes(data)
    assert got.shape == want.shape and np.array_equal(got.astype(np.int64), want)
    assert tok.decode_bytes(got.astype(np.int64)) == data


def test_roundtrip_and_additivity_at_scale(owt_tokenizer):
    """256 MB: decode(encode(x)) == x byte for byte, and the token count of a text equals the sum over its documents'
    halves when it is cut at a special token (the encoder splits there first, tokenizer.py:63-66)."""
    vocab, merges = owt_tokenizer
    tok = get_tokenizer(dict(vocab), list(merges), [EOT])
    n = 256 << 20
    host = synth_host("owt", 4322, n)
    ids = tok.encode_to_numpy(host, np.uint16)
    assert ids.max() < len(vocab)
    back = tok.decode_bytes(ids[: 40_000_000].astype(np.int64))
    assert back == host[: len(back)].tobytes()
    raw = host.tobytes()
    cut = raw.find(EOT.encode(), n // 2)
    a = tok.encode_to_numpy(host[:cut], np.uint16)
    b = tok.encode_to_numpy(host[cut:], np.uint16)
    assert a.size + b.size == ids.size
    assert np.array_equal(ids[: a.size], a) and np.array_equal(ids[a.size:], b)
    # uint16 ids reinterpret as the raw little-endian .bin the reference's trainer memory-maps
    assert ids.dtype == np.uint16 and ids.tobytes()[:2] == int(ids[0]).to_bytes(2, "little")
