"""Registers the hyphen-named package directory `transformer-lm_b200/` as `transformer_lm_b200`."""
import importlib.util
import pathlib
import sys

_ROOT = pathlib.Path(__file__).resolve().parent
_PKG_DIR = _ROOT / "transformer-lm_b200"
ALIAS = "transformer_lm_b200"


def ensure():
    if ALIAS in sys.modules:
        return sys.modules[ALIAS]
    spec = importlib.util.spec_from_file_location(ALIAS, _PKG_DIR / "__init__.py",
                                                  submodule_search_locations=[str(_PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[ALIAS] = mod
    spec.loader.exec_module(mod)
    return mod


ensure()
