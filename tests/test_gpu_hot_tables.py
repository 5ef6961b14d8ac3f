"""The hot tables and the batch machinery of the count stage (train.cu) and of the encoder's lookup (encode.cu) against the oracle.

At their production thresholds these paths need hundreds of megabytes of text (the hot table of the count stage is built after the
first 256 MB batch, the encoder samples a batch of >= 8 M pretokens), which the oracle cannot follow.  The library reads its test
knobs from the environment once per process, so every case runs in a fresh interpreter with small batches (a few MB) and low
thresholds: many batches, the hot table built and rebuilt, the shared-memory cache image, the KeyError position recovered from the
ordinal in a late batch -- all on inputs the oracle finishes in seconds."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code: str, env: dict):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], cwd=ROOT, env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout, r.stderr


TRAIN_CODE = """
    import _bootstrap
    from oracle import oracle
    from transformer_lm_b200.synth import synth_host
    from models.tokenizer.train import train_bpe_on_bytes
    EOT = "<|endoftext|>"
    for shape, seed, mb, vocab in [("owt", 4321, 24, 700), ("tinystories", 1234, 16, 500)]:
        data = synth_host(shape, seed, mb << 20).tobytes()
        want = oracle.train_bpe_on_bytes(data, vocab, [EOT])
        got = train_bpe_on_bytes(data, vocab, [EOT], return_stats=True)
        assert got[1] == want[1] and got[0] == want[0], shape
        print(shape, "ok", got[2]["n_pretokens"], got[2]["n_unique"])
"""


@pytest.mark.parametrize("hot_max", ["20000", "100000", "0"])
def test_count_stage_with_hot_table_matches_oracle(hot_max):
    # 2 MB batches: a dozen batches per corpus; the hot table is built after ~0.2 M pretokens and rebuilt at 4x and 16x that
    out, err = _run(TRAIN_CODE, {"BPE_COUNT_BATCH_KB": "2048", "BPE_COUNT_HOT_AFTER": "200000", "BPE_COUNT_HOT_TEST": "1",
                                 "BPE_COUNT_HOT_MAX": hot_max, "BPE_COUNT_PROFILE": "1"})
    assert out.count(" ok ") == 2
    assert (err.count("[hot table:") >= 4) == (hot_max != "0"), err[-2000:]     # built and rebuilt for both corpora


ENCODE_CODE = """
    import numpy as np
    import _bootstrap
    from oracle import oracle
    from transformer_lm_b200.synth import synth_host
    from models.tokenizer.train import train_bpe_on_bytes
    from tests.adapters import get_tokenizer
    EOT = "<|endoftext|>"
    vocab, merges = train_bpe_on_bytes(synth_host("owt", 4321, 32 << 20), 3000, [EOT])
    tok = get_tokenizer(vocab, merges, [EOT])
    otok = oracle.OracleTokenizer(dict(vocab), list(merges), [EOT])
    text = synth_host("owt", 4322, 6 << 20).tobytes()
    text = text[: text.rfind(EOT.encode())]
    want = otok.encode_bytes(text)
    for rep in range(2):                                       # second pass: everything cached, hot table and image in place
        got = tok.encode_to_numpy(text, np.int32)
        assert len(got) == len(want) and bool((got == want).all()), rep
    assert tok.decode(got.tolist()).encode("utf-8") == text
    # a KeyError in a late batch: the offending pretoken is found through its ordinal
    vocab2 = {i: bytes([i]) for i in range(256)}
    tok2 = get_tokenizer(vocab2, [(b"a", b"b")], [])           # b"ab" is not in the vocab
    body = ("x y zz " * 600000).encode()                       # ~4 MB: several batches of 512 KB
    try:
        tok2.encode_to_numpy(body + b"q ab ab" + body, np.int32)
        raise SystemExit("no KeyError")
    except KeyError as e:
        assert e.args == (b"ab",), e.args
    assert len(tok2.encode_to_numpy(body, np.int32)) == len(body)
    print("encode ok", len(want))
"""


@pytest.mark.parametrize("hot_max", ["1048576", "20000", "0"])
def test_encoder_lookup_with_hot_table_matches_oracle(hot_max):
    out, err = _run(ENCODE_CODE, {"BPE_ENC_BATCH_KB": "512", "BPE_ENC_HOT_MIN": "50000", "BPE_ENC_HOT_TEST": "1", "BPE_ENC_HOT_MAX": hot_max,
                                  "BPE_ENC_PROFILE": "1"})
    assert "encode ok" in out
    assert ("[encoder hot table:" in err) == (hot_max != "0"), err[-2000:]
