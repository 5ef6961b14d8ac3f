"""Host-only synthetic corpus generator (bench / test infrastructure).

oracle/libsynth_host.so is transformer-lm_b200/csrc/synth_gen.h compiled with g++: the CPU arms of bench.py
(--impl reference, cpu_baseline) take their input from here, so those processes never load the product's CUDA library.
The bytes equal bpe_synth_host / bpe_synth_dev of the product library (tests/test_synth.py)."""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libsynth_host.so"
_lib = None
SHAPES = {"tinystories": 0, "tiny": 0, "owt": 1}


def build(force: bool = False) -> pathlib.Path:
    src = _HERE / "synth_host.cpp"
    hdr = _HERE.parent / "transformer-lm_b200" / "csrc" / "synth_gen.h"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.check_call(["make", "-C", str(_HERE), "-B", "libsynth_host.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def synth_host(shape: str, seed: int, n: int, first_block: int = 0) -> np.ndarray:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.synth_host_at.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64]
    a = np.empty(max(n, 1), dtype=np.uint8)
    if _lib.synth_host_at(SHAPES[shape], seed, first_block, a.ctypes.data_as(C.c_void_p), n) != 0:
        raise ValueError("synth_host_at(%r)" % shape)
    return a[:n]
