"""Drop-in for the reference's models/tokenizer/train.py (train_bpe, 142-231)."""
from transformer_lm_b200.train import train_bpe, train_bpe_on_bytes  # noqa: F401
