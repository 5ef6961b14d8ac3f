"""Drop-in for the reference's models/tokenizer/vocab.py (Vocab, 1-43)."""
from transformer_lm_b200.vocab import Vocab  # noqa: F401
