"""transformer-lm_b200: the byte-level BPE hot path of gashon/transformer-lm as sm_100a CUDA kernels.

The directory name contains a hyphen (it mirrors the reference's repository name), so the package is
imported under the alias ``transformer_lm_b200`` -- ``models/__init__.py`` and ``_bootstrap.py`` at the
repository root register it in ``sys.modules``.

Host-side mirror of the reference interface (paths relative to the reference repository root):
    vocab.Vocab                 models/tokenizer/vocab.py:1-43
    train.train_bpe             models/tokenizer/train.py:142-231
    tokenizer.Tokenizer         models/tokenizer/tokenizer.py:11-167
All compute goes through the C ABI of libbpe_sm100.so (include/bpe_sm100.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
