"""CPU-only checks of the C-ABI library: it builds, loads, and exports every symbol the header declares.
No compute is attempted without a GPU; the product must fail loudly instead of falling back."""
import ctypes
import re

import pytest

import _bootstrap  # noqa: F401
from tests.common import REPO_ROOT
from transformer_lm_b200 import _build, _lib


def _declared_functions():
    text = (REPO_ROOT / "include" / "bpe_sm100.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bpe_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = _build.build()
    dll = ctypes.CDLL(str(path))
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(dll, name), "libbpe_sm100.so does not export %s" % name
    for name in _lib.EXPORTS:
        assert hasattr(dll, name), name
    assert dll.bpe_version() >= 100


def test_header_and_binding_list_agree():
    declared = set(_declared_functions())
    bound = set(_lib.EXPORTS)
    assert declared == bound


def test_unicode_table_source_is_stamped():
    L = _lib.lib()
    import regex
    assert L.bpe_unicode_table_source().decode() == "regex-" + regex.__version__


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.BpeError):
        _lib.Context(0)
    from models.tokenizer.train import train_bpe
    from tests.common import FIXTURES_PATH
    with pytest.raises(_lib.BpeError):
        train_bpe(FIXTURES_PATH / "corpus.en", 300, [])


def test_product_never_imports_oracle():
    pkg = REPO_ROOT / "transformer-lm_b200"
    offenders = []
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")) + \
            list((REPO_ROOT / "models").rglob("*.py")):
        src = p.read_text(errors="replace")
        if re.search(r"^\s*(from|import)\s+oracle\b|bpe_oracle|libbpe_oracle", src, flags=re.M):
            offenders.append(str(p))
    assert not offenders, offenders
