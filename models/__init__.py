"""Reference-compatible import path: `models.tokenizer.{train,tokenizer,vocab}` (tests/adapters.py:603,639 of
the reference import exactly these).  The implementation lives in transformer-lm_b200/."""
import pathlib
import sys

_root = str(pathlib.Path(__file__).resolve().parent.parent)
if _root not in sys.path:
    sys.path.insert(0, _root)
import _bootstrap  # noqa: E402,F401  (registers transformer_lm_b200)
