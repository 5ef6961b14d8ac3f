"""Synthetic corpora (bench / test infrastructure): deterministic, valid UTF-8, no carriage returns,
block-addressable (a rank can generate just its shard), identical on CPU and GPU."""
import numpy as np
import pytest

import _bootstrap  # noqa: F401
from oracle import oracle
from transformer_lm_b200.synth import BLOCK, synth_host


@pytest.mark.parametrize("shape,seed", [("owt", 4321), ("owt", 4322), ("tinystories", 1234)])
def test_host_corpus_is_valid_deterministic_and_block_addressable(shape, seed):
    n = 64 * BLOCK
    a = synth_host(shape, seed, n).tobytes()
    assert a == synth_host(shape, seed, n).tobytes()
    text = a.decode("utf-8")                         # strict: raises on invalid UTF-8
    assert "\r" not in text and "<|endoftext|>" in text
    assert synth_host(shape, seed, 16 * BLOCK).tobytes() == a[: 16 * BLOCK]          # prefix property
    assert synth_host(shape, seed, 8 * BLOCK, first_block=40).tobytes() == a[40 * BLOCK: 48 * BLOCK]
    assert synth_host(shape, seed + 1, n).tobytes() != a
    for b in range(0, 64, 7):                        # every block is valid UTF-8 on its own (shards cut at block seams)
        a[b * BLOCK:(b + 1) * BLOCK].decode("utf-8")


def test_shapes_have_the_documented_statistics():
    owt = synth_host("owt", 4321, 4 << 20).tobytes()
    tiny = synth_host("tinystories", 1234, 4 << 20).tobytes()
    co, ct = oracle.count_pretokens(owt, []), oracle.count_pretokens(tiny, [])
    bo, bt = len(owt) / sum(co.values()), len(tiny) / sum(ct.values())
    assert 4.0 < bo < 5.2 and 3.3 < bt < 4.5         # bytes per pretoken
    assert len(co) > 4 * len(ct)                      # OWT shape has the long tail
    non_ascii = sum(1 for b in owt if b >= 128) / len(owt)
    assert 0.005 < non_ascii < 0.03
    assert any(len(w) > 12 for w in co) and b" the" not in co      # synthetic lexicon, not English


@pytest.mark.gpu
def test_device_generator_matches_host():
    import torch
    from transformer_lm_b200.synth import synth_device
    for shape, seed in (("owt", 4321), ("tinystories", 1234)):
        n = 300 * BLOCK + 123
        t = torch.empty(n, dtype=torch.uint8, device="cuda")
        synth_device(shape, seed, n, t.data_ptr())
        assert np.array_equal(t.cpu().numpy(), synth_host(shape, seed, n))
        t2 = torch.empty(5 * BLOCK, dtype=torch.uint8, device="cuda")
        synth_device(shape, seed, 5 * BLOCK, t2.data_ptr(), first_block=77)
        assert np.array_equal(t2.cpu().numpy(), synth_host(shape, seed, 82 * BLOCK)[77 * BLOCK:])
