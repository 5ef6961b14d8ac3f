"""GPU parity of BPE training (bpe_train through the reference-compatible adapter) against the reference's
golden snapshot, golden vectors generated from the reference, and the CPU oracle on seeded corpora."""
import random
import time

import pytest

import _bootstrap  # noqa: F401
from oracle import oracle
from tests.adapters import run_train_bpe
from tests.common import FIXTURES_PATH, load_golden, load_reference_train_snapshot

pytestmark = pytest.mark.gpu


def _train_bytes(data, vocab_size, specials, **kw):
    from models.tokenizer.train import train_bpe_on_bytes
    return train_bpe_on_bytes(data, vocab_size, specials, **kw)


def test_train_bpe_reference_snapshot():
    """The reference's golden test (tests/test_train_bpe.py:28-65): corpus.en, vocab 500."""
    vocab, merges = run_train_bpe(FIXTURES_PATH / "corpus.en", 500, ["<|endoftext|>"])
    ref_vocab, ref_merges = load_reference_train_snapshot()
    assert merges == ref_merges
    assert set(vocab.keys()) == set(ref_vocab.keys())
    assert set(vocab.values()) == set(ref_vocab.values())


def test_train_bpe_speed():
    """tests/test_train_bpe.py:9-25 of the reference: < 1.5 s (CUDA context already created by earlier tests
    or not -- the budget includes it)."""
    t0 = time.time()
    run_train_bpe(FIXTURES_PATH / "corpus.en", 500, ["<|endoftext|>"])
    assert time.time() - t0 < 1.5


def _golden_input(case):
    src = case["source"]
    if src.startswith("fixture:"):
        return (FIXTURES_PATH / src.split(":", 1)[1]).read_bytes()
    return bytes.fromhex(case["input_hex"])


@pytest.mark.parametrize("name", sorted(load_golden("train_golden.json").keys()))
def test_train_golden(name):
    case = load_golden("train_golden.json")[name]
    data = _golden_input(case)
    if "error" in case:
        with pytest.raises(UnicodeDecodeError):
            _train_bytes(data, case["vocab_size"], case["special_tokens"])
        return
    vocab, merges = _train_bytes(data, case["vocab_size"], case["special_tokens"])
    assert [[a.hex(), b.hex()] for a, b in merges] == case["merges"]
    assert {str(k): v.hex() for k, v in vocab.items()} == case["vocab"]


def _fuzz_corpus(seed, n_words, alphabet_words):
    r = random.Random(seed)
    out = []
    for _ in range(n_words):
        out.append(r.choice(alphabet_words))
        out.append(r.choice([" ", " ", " ", "\n", "  ", ", ", ". ", ""]))
    return "".join(out).encode("utf-8")


WORDS = ["the", "there", "then", "other", "a", "an", "and", "band", "hand", "it's", "don't", "we'll", "naïve", "日本", "日本語",
         "🙃", "🙃🙃", "1", "12", "123", "2024", "x_y", "--", "----", "http://a.b/c", "aaaa", "aaa", "aa", "abab", "ababab", "Zürich"]


@pytest.mark.parametrize("seed,n_words,vocab_size", [(1, 50, 400), (2, 500, 600), (3, 5000, 900), (4, 20000, 1500), (5, 12, 2000)])
def test_train_differential_vs_oracle(seed, n_words, vocab_size):
    data = _fuzz_corpus(seed, n_words, WORDS)
    want = oracle.train_bpe_on_bytes(data, vocab_size, ["<|endoftext|>"])
    got = _train_bytes(data, vocab_size, ["<|endoftext|>"])
    assert got[1] == want[1]
    assert got[0] == want[0]


def test_train_corpus_vocab_3000_vs_oracle():
    data = (FIXTURES_PATH / "corpus.en").read_bytes()
    want = oracle.train_bpe_on_bytes(data, 3000, ["<|endoftext|>"])
    got = _train_bytes(data, 3000, ["<|endoftext|>"])
    assert got[1] == want[1]
    assert got[0] == want[0]


def test_train_is_deterministic():
    data = _fuzz_corpus(7, 8000, WORDS)
    a = _train_bytes(data, 1200, [])
    b = _train_bytes(data, 1200, [])
    assert a == b


def test_train_long_pretokens_and_crlf():
    data = (b"x" * 5000 + b" " + b"ab" * 3000 + b"\r\n" + b"x" * 5000 + b"\r" + b"ab" * 3000 + b" 0123456789abcdef 0123456789abcdef ") * 3
    want = oracle.train_bpe_on_bytes(data, 400, [])
    got = _train_bytes(data, 400, [])
    assert got == want


def test_train_stats_are_reported():
    data = (FIXTURES_PATH / "corpus.en").read_bytes()
    vocab, merges, stats = _train_bytes(data, 500, ["<|endoftext|>"], return_stats=True)
    assert stats["n_bytes"] == len(data.replace(b"\r\n", b"\n"))
    assert stats["n_pretokens"] == len(oracle.pretokenize(data))
    assert len(merges) == 243 and stats["duplicate_tokens"] == 0
