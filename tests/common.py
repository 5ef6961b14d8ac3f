"""Shared helpers for the test-suite (fixtures, GPT-2 byte<->unicode table, golden loaders)."""
from __future__ import annotations

import json
import pathlib
from functools import lru_cache

TESTS_PATH = pathlib.Path(__file__).resolve().parent
FIXTURES_PATH = TESTS_PATH / "fixtures"
GOLDEN_PATH = TESTS_PATH / "golden"
REPO_ROOT = TESTS_PATH.parent


@lru_cache()
def gpt2_bytes_to_unicode() -> dict[int, str]:
    """GPT-2's printable stand-in for each byte value (the encoding used by gpt2_vocab.json,
    gpt2_merges.txt and train-bpe-reference-*.{json,txt}): printable Latin-1 bytes map to
    themselves, the other 68 bytes map to U+0100, U+0101, ... in increasing byte order."""
    keep = [b for b in range(256) if (33 <= b <= 126) or (161 <= b <= 172) or (174 <= b <= 255)]
    table = {b: chr(b) for b in keep}
    shift = 0
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + shift)
            shift += 1
    return table


@lru_cache()
def gpt2_unicode_to_byte() -> dict[str, int]:
    return {c: b for b, c in gpt2_bytes_to_unicode().items()}


def gpt2_decode(s: str) -> bytes:
    dec = gpt2_unicode_to_byte()
    return bytes(dec[c] for c in s)


def load_gpt2_fixture():
    """(vocab: dict[int, bytes], merges: list[tuple[bytes, bytes]]) from the GPT-2 fixtures."""
    with open(FIXTURES_PATH / "gpt2_vocab.json", encoding="utf-8") as f:
        raw = json.load(f)
    vocab = {idx: gpt2_decode(tok) for tok, idx in raw.items()}
    merges = []
    with open(FIXTURES_PATH / "gpt2_merges.txt", encoding="utf-8") as f:
        for line in f:
            parts = line.rstrip().split(" ")
            if line.rstrip() and len(parts) == 2:
                merges.append((gpt2_decode(parts[0]), gpt2_decode(parts[1])))
    return vocab, merges


def load_reference_train_snapshot():
    """The reference's golden snapshot for corpus.en / vocab 500 (tests/test_train_bpe.py:28-65)."""
    with open(FIXTURES_PATH / "train-bpe-reference-merges.txt", encoding="utf-8") as f:
        merges = [tuple(gpt2_decode(t) for t in line.rstrip().split(" ")) for line in f]
    with open(FIXTURES_PATH / "train-bpe-reference-vocab.json", encoding="utf-8") as f:
        vocab = {idx: gpt2_decode(tok) for tok, idx in json.load(f).items()}
    return vocab, merges


@lru_cache()
def load_golden(name: str):
    with open(GOLDEN_PATH / name, encoding="utf-8") as f:
        return json.load(f)
