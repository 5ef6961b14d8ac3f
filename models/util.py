"""Reference-compatible import path `models.util.load_batch` (models/util.py:37-57 of the reference); the gather runs on the
GPU (transformer-lm_b200/batch.py -> bpe_batch_windows_dev).  Checkpoint helpers of that module are out of scope."""
from transformer_lm_b200.batch import load_batch  # noqa: F401
