"""The plugin seam of the reference's test-suite (tests/adapters.py:580-643 there): the two functions its
tokenizer tests call.  They import through the reference-compatible `models.tokenizer.*` path."""
from __future__ import annotations

import os
from typing import Optional


def get_tokenizer(vocab: dict[int, bytes], merges: list[tuple[bytes, bytes]], special_tokens: Optional[list[str]] = None):
    from models.tokenizer.tokenizer import Tokenizer

    return Tokenizer(vocab, merges, special_tokens)


def run_train_bpe(input_path: str | os.PathLike, vocab_size: int, special_tokens: list[str], **kwargs):
    from models.tokenizer.train import train_bpe

    vocab, merges = train_bpe(input_path, vocab_size, special_tokens)
    return vocab, merges


def run_get_batch(dataset, batch_size: int, context_length: int, device: str):
    """tests/adapters.py:398-424 of the reference."""
    from models.util import load_batch

    return load_batch(dataset, batch_size, context_length, device)
