"""GPU parity: device pretokenizer / UTF-8 validator vs the CPU oracle (and `regex` where installed)."""
import random
import re

import numpy as np
import pytest

import _bootstrap  # noqa: F401
from oracle import oracle
from tests.common import FIXTURES_PATH

pytestmark = pytest.mark.gpu

ALPHA = list("'sdmtlvre ab1.\n\t") + [" ", " ", "\x85", "\x1c", "中", "\U0001F643", "é", "١", "<", "|", ">",
                                       "'ll", "'ve", "'re", "  ", "<|endoftext|>", "　", "²", "́"]


def _starts(data, specials=None):
    from transformer_lm_b200.pretok import pretoken_starts
    return pretoken_starts(data, specials).tolist()


def _expected_with_specials(data: bytes, specials):
    sp = sorted(set(specials), key=len, reverse=True)
    if not sp:
        return oracle.pretokenize(data)
    rgx = b"(" + b"|".join(re.escape(s.encode()) for s in sp) + b")"
    out, pos = [], 0
    for i, seg in enumerate(re.split(rgx, data)):
        if seg == b"":
            continue
        if i % 2 == 1:
            out.append(pos)
        else:
            out.extend(pos + s for s in oracle.pretokenize(seg))
        pos += len(seg)
    return out


@pytest.mark.parametrize("name", ["corpus.en", "address.txt", "german.txt", "tinystories_sample.txt"])
def test_fixture_boundaries(name):
    data = (FIXTURES_PATH / name).read_bytes().replace(b"\r", b"")
    assert _starts(data) == oracle.pretokenize(data)


def test_fuzz_boundaries_one_big_buffer():
    rnd = random.Random(5)
    # many adversarial snippets concatenated: exercises tile and chunk seams at every alignment
    text = "".join(rnd.choice(ALPHA) for _ in range(400000))
    data = text.encode("utf-8")
    got = _starts(data)
    want = oracle.pretokenize(data)
    assert got == want


def test_fuzz_boundaries_small_strings():
    rnd = random.Random(6)
    for _ in range(300):
        s = "".join(rnd.choice(ALPHA) for _ in range(rnd.randint(0, 40)))
        data = s.encode("utf-8")
        assert _starts(data) == oracle.pretokenize(data), repr(s)


def test_every_offset_alignment():
    base = "it's they'll  we've\n\nI'm 🙃 x'llama 12 3.4  \n  end"
    for pad in range(0, 40):
        s = "a" * pad + " " + base
        data = s.encode("utf-8")
        assert _starts(data) == oracle.pretokenize(data), pad


def test_long_runs_cross_tiles():
    data = (b" " * 10000 + b"x" * 9000 + b"\n" * 5000 + "é".encode() * 3000 + b"1" * 4097 + b"!" * 8192 + b"   y")
    assert _starts(data) == oracle.pretokenize(data)


def test_empty_and_tiny():
    assert _starts(b"") == []
    assert _starts(b"a") == [0]
    assert _starts(" ".encode()) == [0]


@pytest.mark.parametrize("specials", [["<|endoftext|>"], ["<|endoftext|>", "<|endoftext|><|endoftext|>"], ["aa"], ["aa", "aaa", "b"],
                                      ["s", "'s"], ["ab", "ba"]])
def test_special_token_split(specials):
    rnd = random.Random(9)
    alpha = ALPHA + ["<|endoftext|>", "<|endoftext|>", "a", "aa", "b", "ab"]
    for trial in range(40):
        s = "".join(rnd.choice(alpha) for _ in range(rnd.randint(0, 300)))
        data = s.encode("utf-8")
        assert _starts(data, specials) == _expected_with_specials(data, specials), (trial, repr(s))


def test_special_pathological_chain():
    data = b"a" * 5000 + b"b" + b"a" * 33
    for sp in (["aa"], ["aaa", "aa"], ["aaaa", "a"]):
        assert _starts(data, sp) == _expected_with_specials(data, sp)


def test_utf8_validation_matches_cpython():
    from transformer_lm_b200.pretok import utf8_validate
    rnd = random.Random(3)
    pieces = [b"a", b"\xc3\xa9", b"\xe2\x82\xac", b"\xf0\x9f\x99\x83", b"\xff", b"\xc0\x80", b"\xed\xa0\x80", b"\xf4\x90\x80\x80",
              b"\xe2\x82", b"\x80", b"\xf0\x9f", b"\xe0\x9f\xbf", b"\xf0\x8f\xbf\xbf", b"\xc2", b"hello world " * 3, b"\xfe", b"\xf8\x88\x80\x80\x80"]
    for _ in range(300):
        data = b"".join(rnd.choice(pieces) for _ in range(rnd.randint(0, 12)))
        try:
            data.decode("utf-8")
            want = -1
        except UnicodeDecodeError as e:
            want = e.start
        assert utf8_validate(data) == want, data
    big = bytearray(("héllo wörld 🙃 " * 20000).encode())
    assert utf8_validate(bytes(big)) == -1
    big[123457] = 0xFF
    try:
        bytes(big).decode("utf-8")
    except UnicodeDecodeError as e:
        assert utf8_validate(bytes(big)) == e.start


def test_every_code_point_goes_through_the_device_class_table():
    """Every scalar value U+0000..U+10FFFF (surrogates excluded) between an ASCII letter and a digit: the device's
    two-level class table in shared memory (pretok.cuh cp_class_smem) against the oracle's table walk, so all 684 L
    ranges / 146 N ranges / 25 whitespace code points are pinned on the GPU too, in all four UTF-8 lengths."""
    cps = [cp for cp in range(0x110000) if not 0xD800 <= cp <= 0xDFFF]
    for lo in range(0, len(cps), 1 << 18):
        text = "".join("a" + chr(cp) + "1 " for cp in cps[lo: lo + (1 << 18)])
        data = text.encode("utf-8")
        got = np.asarray(_starts(data), dtype=np.int64)
        want = np.asarray(oracle.pretokenize(data), dtype=np.int64)
        assert got.shape == want.shape and np.array_equal(got, want), "code points %#x.." % cps[lo]
    # and back to back without separators (class transitions between neighbouring code points)
    text = "".join(chr(cp) for cp in cps[:: 7])
    data = text.encode("utf-8")
    assert _starts(data) == oracle.pretokenize(data)
