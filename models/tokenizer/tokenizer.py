"""Drop-in for the reference's models/tokenizer/tokenizer.py (Tokenizer, 11-167)."""
from transformer_lm_b200.tokenizer import Tokenizer  # noqa: F401
