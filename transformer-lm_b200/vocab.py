"""id -> bytes table.  Mirrors models/tokenizer/vocab.py:1-43 of the reference (same names and
semantics; membership is tracked in a set so add_token is O(1) instead of a scan over the values)."""
from __future__ import annotations


class Vocab:
    def __init__(self, special_tokens: list[str] = []) -> None:
        self.idx_to_token: dict[int, bytes] = {}
        self._present: set[bytes] = set()
        # ids 0..S-1 = specials in the order given, then the 256 byte values (vocab.py:5-10);
        # a byte string that is already present is skipped (vocab.py:28-34, SURVEY A-5)
        for token in special_tokens:
            self.add_token(token.encode("utf-8"))
        for i in range(256):
            self.add_token(bytes([i]))
        self.unk_idx: int = 0

    @classmethod
    def from_dict(cls, vocab: dict[int, bytes], special_tokens: list[str] = []) -> "Vocab":
        instance = cls(special_tokens)
        instance.idx_to_token = vocab
        instance._present = set(vocab.values())
        return instance

    def __len__(self) -> int:
        return len(self.idx_to_token)

    def __getitem__(self, idx: int) -> bytes:
        return self.idx_to_token.get(idx, self.idx_to_token[self.unk_idx])

    def add_token(self, token: bytes) -> None:
        if token in self._present:
            return
        self._present.add(token)
        self.idx_to_token[len(self.idx_to_token)] = token

    def get_inv(self) -> dict[bytes, int]:
        return {v: k for k, v in self.idx_to_token.items()}

    def get_idx_to_token(self) -> dict[int, bytes]:
        return self.idx_to_token

    def set_unk_idx(self, unk_idx: int) -> None:
        self.unk_idx = unk_idx
