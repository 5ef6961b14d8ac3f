"""GPU parity of Tokenizer.encode / encode_iterable / decode (bpe_tok_create, bpe_encode, bpe_decode through the
reference-compatible adapter) against: the reference's own tests (tests/test_tokenizer.py there: roundtrips,
tiktoken GPT-2 equality, overlapping specials), golden vectors generated from the unmodified reference
(tests/golden/encode_golden.json), and the CPU oracle on seeded fuzz inputs."""
import random

import numpy as np
import pytest

import _bootstrap  # noqa: F401
from oracle import oracle
from tests.adapters import get_tokenizer
from tests.common import FIXTURES_PATH, load_golden, load_gpt2_fixture
from tests.test_oracle_pins import GPT2_PAT, _golden_tokenizer

pytestmark = pytest.mark.gpu

EOT = "<|endoftext|>"


@pytest.fixture(scope="module")
def gpt2():
    vocab, merges = load_gpt2_fixture()
    return get_tokenizer(vocab, merges, [EOT])


@pytest.fixture(scope="module")
def gpt2_plain():
    vocab, merges = load_gpt2_fixture()
    return get_tokenizer(vocab, merges, None)


@pytest.fixture(scope="module")
def tiktoken_gpt2():
    tiktoken = pytest.importorskip("tiktoken")
    vocab, _ = load_gpt2_fixture()
    ranks = {v: k for k, v in vocab.items() if v != EOT.encode()}
    return tiktoken.Encoding("gpt2-local", pat_str=GPT2_PAT, mergeable_ranks=ranks, special_tokens={EOT: 50256})


# ---- the reference's roundtrip tests (tests/test_tokenizer.py:82-90,112-120,142-150,169-177,198-206) ----
@pytest.mark.parametrize("text", ["", "s", "🙃", "Hello, how are you?", "Héllò hôw are ü? 🙃"])
def test_roundtrip_strings(gpt2_plain, text):
    ids = gpt2_plain.encode(text)
    assert gpt2_plain.decode(ids) == text


@pytest.mark.parametrize("text", ["", "s", "🙃", "Hello, how are you?", "Héllò hôw are ü? 🙃"])
def test_strings_match_tiktoken(gpt2_plain, tiktoken_gpt2, text):
    assert gpt2_plain.encode(text) == tiktoken_gpt2.encode(text)


def test_known_ids(gpt2_plain):
    assert gpt2_plain.encode("Hello, how are you?") == [15496, 11, 703, 389, 345, 30]     # test_tokenizer.py:192
    assert gpt2_plain.encode("🙃") == [8582, 247, 225]
    assert gpt2_plain.encode("s") == [82]
    assert [gpt2_plain.decode([i]) for i in gpt2_plain.encode("Hello, how are you?")] == ["Hello", ",", " how", " are", " you", "?"]


def test_roundtrip_unicode_string_with_special_tokens(gpt2):
    text = "Héllò hôw <|endoftext|><|endoftext|> are ü? 🙃<|endoftext|>"      # test_tokenizer.py:224-235
    ids = gpt2.encode(text)
    pieces = [gpt2.decode([x]) for x in ids]
    assert pieces.count(EOT) == 3
    assert gpt2.decode(ids) == text


def test_special_tokens_match_tiktoken(gpt2, tiktoken_gpt2):
    text = "Héllò hôw <|endoftext|><|endoftext|> are ü? 🙃<|endoftext|>"      # test_tokenizer.py:238-252
    assert gpt2.encode(text) == tiktoken_gpt2.encode(text, allowed_special={EOT})


def test_overlapping_special_tokens():
    vocab, merges = load_gpt2_fixture()                                       # test_tokenizer.py:255-269
    vocab[len(vocab)] = b"<|endoftext|><|endoftext|>"
    tok = get_tokenizer(vocab, merges, [EOT, EOT + EOT])
    text = "Hello, how <|endoftext|><|endoftext|> are you?<|endoftext|>"
    ids = tok.encode(text)
    pieces = [tok.decode([x]) for x in ids]
    assert pieces.count(EOT) == 1
    assert pieces.count(EOT + EOT) == 1
    assert tok.decode(ids) == text


@pytest.mark.parametrize("name", ["address.txt", "german.txt", "tinystories_sample.txt", "corpus.en"])
def test_fixture_files_roundtrip_and_match_tiktoken(gpt2, tiktoken_gpt2, name):
    text = (FIXTURES_PATH / name).read_text(encoding="utf-8")                  # test_tokenizer.py:272-357
    ids = gpt2.encode(text)
    assert ids == tiktoken_gpt2.encode(text, allowed_special={EOT})
    assert gpt2.decode(ids) == text


def test_encode_iterable_tinystories(gpt2, tiktoken_gpt2):
    path = FIXTURES_PATH / "tinystories_sample.txt"                           # test_tokenizer.py:360-392
    with open(path, encoding="utf-8") as f:
        ids = list(gpt2.encode_iterable(f))
    text = path.read_text(encoding="utf-8")
    assert gpt2.decode(ids) == text
    assert ids == tiktoken_gpt2.encode(text, allowed_special={EOT})


def test_encode_iterable_is_lazy_and_chunked(gpt2):
    # chunks are cut between items once >= 2 Mi characters are buffered and encoded independently (SURVEY A-13)
    line = "x" * 1023 + "\n"
    lines = [line] * 2048 + ["\n\nFoo"]         # the first 2048 items are exactly 2 Mi characters
    it = iter(lines)
    gen = gpt2.encode_iterable(it)
    first = next(gen)
    assert isinstance(first, int)
    assert next(it, None) is not None          # the generator has not drained the iterable yet
    it2 = iter(lines)
    got = list(gpt2.encode_iterable(it2))
    otok = oracle.OracleTokenizer(*load_gpt2_fixture(), [EOT])
    assert got == list(otok.encode_iterable(iter(lines)))
    assert got != gpt2.encode("".join(lines))  # "...\n" | "\n\nFoo" splits differently from the concatenation


def test_encode_iterable_look_ahead_keeps_order_and_errors(gpt2):
    # chunk k + 1 is encoded on a worker thread while the caller consumes chunk k.  Same ids, same chunk
    # cuts as the sequential loop over several chunks; the iterable is read at most one chunk ahead of the ids consumed.
    rng = random.Random(5)
    words = ["alpha", " beta", "\n", " 12", "gamma!", " δ", "  ", "\n\n", EOT]
    lines = ["".join(rng.choice(words) for _ in range(rng.randint(50, 400))) + "\n" for _ in range(9000)]
    assert sum(map(len, lines)) > 3 * 2 * 1024 * 1024                       # at least four chunks
    otok = oracle.OracleTokenizer(*load_gpt2_fixture(), [EOT])
    want = list(otok.encode_iterable(iter(lines)))
    pulled = [0]

    def counting():
        for line in lines:
            pulled[0] += 1
            yield line
    gen = gpt2.encode_iterable(counting())
    got = [next(gen)]
    first_chunk_items = pulled[0]
    assert first_chunk_items < len(lines) // 2                               # nothing read ahead before the first id
    got.append(next(gen))
    assert first_chunk_items < pulled[0] < 2.5 * first_chunk_items            # exactly one chunk ahead afterwards
    mid = gpt2.encode("interleaved call on the same context")               # the caller may use the tokenizer meanwhile
    assert mid == otok.encode("interleaved call on the same context")
    got.extend(gen)
    assert got == want
    # an error of a later chunk surfaces after every id of the earlier chunks, like in the sequential loop
    vocab = {i: bytes([i]) for i in range(256)}
    tok = get_tokenizer(vocab, [(b"a", b"b")], [])                           # b"ab" is not in the vocab
    good = ["x" * 1023 + "\n"] * 2048
    seen = []
    with pytest.raises(KeyError) as ei:
        for t in tok.encode_iterable(iter(good + ["ab\n"])):
            seen.append(t)
    assert ei.value.args == (b"ab",)
    assert len(seen) == 2048 * 1024


# ---- golden vectors produced by the unmodified reference -------------------------------------------------
@pytest.mark.parametrize("name", sorted(load_golden("encode_golden.json").keys()))
def test_encode_golden(name):
    entry = load_golden("encode_golden.json")[name]
    vocab, merges = _golden_tokenizer(entry)
    tok = get_tokenizer(vocab, merges, entry["special_tokens"])
    for text, want in zip(entry["texts"], entry["results"]):
        if "error" in want:
            with pytest.raises(KeyError) as ei:
                tok.encode(text)
            assert ei.value.args[0].hex() == want["arg"]
        else:
            ids = tok.encode(text)
            assert ids == want["ids"], repr(text[:80])
            assert tok.decode(ids) == want["decoded"]
    if "iterable_ids" in entry:
        with open(FIXTURES_PATH / entry["iterable_source"], encoding="utf-8") as f:
            assert list(tok.encode_iterable(f)) == entry["iterable_ids"]


# ---- differential vs the oracle -------------------------------------------------------------------------
ALPHA = list("'sdmtlvre ab1.\n\t") + [" ", " ", "\x85", "中", "\U0001F643", "é", "١", "<", "|", ">", "'ll", "'ve", "'re", "  ",
                                       EOT, "the", " the", " and", "ing", "tion", "http://", "0", "00", "aaaa", "\n\n"]


def _trained(vocab_size, specials):
    tv = load_golden("train_golden.json")
    name = "corpus_1000_eot"
    vocab = {int(k): bytes.fromhex(v) for k, v in tv[name]["vocab"].items()}
    merges = [(bytes.fromhex(a), bytes.fromhex(b)) for a, b in tv[name]["merges"]]
    return vocab, merges


@pytest.mark.parametrize("which", ["gpt2", "corpus1000"])
def test_fuzz_vs_oracle(which):
    vocab, merges = load_gpt2_fixture() if which == "gpt2" else _trained(1000, [EOT])
    tok = get_tokenizer(dict(vocab), list(merges), [EOT])
    otok = oracle.OracleTokenizer(dict(vocab), list(merges), [EOT])
    rnd = random.Random(11)
    for trial in range(60):
        s = "".join(rnd.choice(ALPHA) for _ in range(rnd.randint(0, 400)))
        assert tok.encode(s) == otok.encode(s), (trial, repr(s))
    big = "".join(rnd.choice(ALPHA) for _ in range(300000))
    ids = tok.encode(big)
    assert ids == otok.encode(big)
    assert tok.decode(ids) == big


def test_long_pretokens_take_the_chunked_path(gpt2):
    otok = oracle.OracleTokenizer(*load_gpt2_fixture(), [EOT])
    rnd = random.Random(5)
    parts = ["a" * 33, "a" * 1000, " " * 777 + "x", "ab" * 500, "=" * 4097, "\n" * 100,
             "".join(rnd.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(3000)),
             "".join(rnd.choice("0123456789") for _ in range(2000)),
             "é" * 700, "🙃" * 300, "the" * 400, "x" * 32, "y" * 31, "z" * 64, "w" * 65]
    for p in parts:
        assert gpt2.encode(p) == otok.encode(p), p[:20]
    text = " ".join(parts)
    ids = gpt2.encode(text)
    assert ids == otok.encode(text)
    assert gpt2.decode(ids) == text


def test_cache_is_consistent_across_calls_and_resets(gpt2):
    text = (FIXTURES_PATH / "corpus.en").read_text(encoding="utf-8")
    a = gpt2.encode(text)
    b = gpt2.encode(text)                      # second call: every pretoken is a cache hit
    assert a == b
    assert gpt2.last_stats["cache_new_unique"] == 0
    from transformer_lm_b200 import _lib
    _lib.lib().bpe_tok_cache_reset(gpt2._device_tok())
    assert gpt2.encode(text) == a
    assert gpt2.last_stats["cache_new_unique"] > 0


def test_uint16_and_int32_outputs_agree(gpt2):
    data = (FIXTURES_PATH / "german.txt").read_bytes()
    a = gpt2.encode_to_numpy(data, np.uint16)
    b = gpt2.encode_to_numpy(data, np.int32)
    assert a.dtype == np.uint16 and b.dtype == np.int32
    assert a.astype(np.int64).tolist() == b.astype(np.int64).tolist()


def test_uint16_output_refused_when_ids_do_not_fit():
    vocab = {i: bytes([i]) for i in range(256)}
    vocab[70000] = b"ab"
    tok = get_tokenizer(vocab, [(b"a", b"b")], [])
    assert tok.encode("abc") == [70000, 99]
    from transformer_lm_b200 import _lib
    with pytest.raises(_lib.BpeError):
        tok.encode_to_numpy(b"abc", np.uint16)


def test_key_errors_like_the_reference():
    vocab = {i: bytes([i]) for i in range(256)}
    tok = get_tokenizer(vocab, [(b"a", b"b"), (b"ab", b"c")], [])         # b"ab", b"abc" are not in the vocab
    otok = oracle.OracleTokenizer(dict(vocab), [(b"a", b"b"), (b"ab", b"c")], [])
    for text in ["xx abc", "zab", "q abq abc"]:
        with pytest.raises(KeyError) as e1:
            tok.encode(text)
        with pytest.raises(KeyError) as e2:
            otok.encode(text)
        assert e1.value.args == e2.value.args
    assert tok.encode("xyz") == [120, 121, 122]
    with pytest.raises(KeyError):
        tok.decode([1, 2, 999999])
    with pytest.raises(KeyError):
        tok.decode([-5])


def test_duplicate_merges_and_unreachable_merges():
    vocab = {i: bytes([i]) for i in range(256)}
    for t in [b"ab", b"abc", b"bc", b"zz"]:
        vocab[len(vocab)] = t
    merges = [(b"a", b"b"), (b"b", b"c"), (b"ab", b"c"), (b"a", b"b"), (b"q", b"zz"), (b"zzz", b"z"), (b"a", b"bc")]
    tok = get_tokenizer(dict(vocab), list(merges), [])
    otok = oracle.OracleTokenizer(dict(vocab), list(merges), [])
    for text in ["abc", "ababc abcabc", "bcbc", "aabcc", "zzzz qzz"]:
        assert tok.encode(text) == otok.encode(text), text


def test_special_missing_from_vocab_quirk():
    vocab = {i: bytes([i]) for i in range(256)}                          # SURVEY A-12
    tok = get_tokenizer(dict(vocab), [], ["<|x|>"])
    otok = oracle.OracleTokenizer(dict(vocab), [], ["<|x|>"])
    ids = tok.encode("a<|x|>b")
    assert ids == otok.encode("a<|x|>b") == [97, 256, 98]
    with pytest.raises(KeyError):
        tok.decode(ids)


def test_save_and_from_files_roundtrip(tmp_path, gpt2):
    from models.tokenizer.tokenizer import Tokenizer
    vocab, merges = _trained(1000, [EOT])
    tok = Tokenizer(vocab, merges, [EOT])
    tok.save(str(tmp_path), prefix="corpus")
    tok2 = Tokenizer.from_files(str(tmp_path / "corpus-vocab.pkl"), str(tmp_path / "corpus-merges.pkl"), [EOT])
    text = (FIXTURES_PATH / "address.txt").read_text(encoding="utf-8")
    assert tok2.encode(text) == tok.encode(text)
    import pickle
    assert pickle.load(open(tmp_path / "corpus-merges.pkl", "rb")) == merges


def test_train_from_file_then_encode():
    from models.tokenizer.tokenizer import Tokenizer
    tok = Tokenizer.train_from_file(FIXTURES_PATH / "corpus.en", 600, [EOT])
    ov, om = oracle.train_bpe(FIXTURES_PATH / "corpus.en", 600, [EOT])
    assert tok.merges == om and tok.vocab == ov
    text = (FIXTURES_PATH / "tinystories_sample.txt").read_text(encoding="utf-8")
    assert tok.encode(text) == oracle.OracleTokenizer(ov, om, [EOT]).encode(text)


def test_helper_methods_segment_match_pretokenize(gpt2):
    import regex
    text = "Hello<|endoftext|>world's  end<|endoftext|>"
    assert gpt2.segment(text) == regex.split("(" + regex.escape(EOT) + ")", text)
    assert gpt2.match("it's  fine\n") == regex.findall(GPT2_PAT, "it's  fine\n")
    assert gpt2.pretokenize(gpt2.segment(text)) == ["Hello", EOT, "world", "'s", " ", " end", EOT]
    assert gpt2.merge([b"a", b"a", b"a", b"b"], (b"a", b"a"), b"aa") == [b"aa", b"a", b"b"]


@pytest.mark.parametrize("specials", [[EOT], []])
def test_pipelined_host_encode_equals_single_pass(specials):
    """bpe_encode on host memory cuts big inputs at exact boundaries (special tokens, or a lone space between ASCII
    non-space bytes) and double-buffers the chunks; bpe_encode_dev encodes the same text in one piece."""
    import ctypes as C
    import torch
    from transformer_lm_b200 import _lib
    from transformer_lm_b200.synth import synth_host
    vocab, merges = _trained(1000, [EOT])
    tok = get_tokenizer(dict(vocab), list(merges), specials)
    n = 300 << 20
    host = synth_host("owt", 4322, n)
    got = tok.encode_to_numpy(host, np.uint16)
    h, ctx, L = tok._device_tok(), tok._tok_ctx, _lib.lib()
    dev = torch.from_numpy(host).cuda()
    out = torch.empty(n, dtype=torch.uint16, device="cuda")
    n_out = C.c_uint64(0)
    ctx.check(L.bpe_encode_dev(h, C.c_void_p(dev.data_ptr()), n, _lib.DTYPE_U16, C.c_void_p(out.data_ptr()), n, C.byref(n_out), None))
    want = out[: n_out.value].cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got, want)
    # and the first MB against the oracle (cut at a document boundary)
    piece = host[: 1 << 20].tobytes()
    piece = piece[: piece.rfind(b"<|endoftext|>")]
    assert tok.encode_to_numpy(piece, np.int32).tolist() == oracle.OracleTokenizer(dict(vocab), list(merges), specials).encode_bytes(piece).tolist()


def test_bulk_encode_driver_streams_a_file(tmp_path):
    """encode_file (replaces models/tokenizer/encode.py): pieces cut at exact boundaries, text-mode newline handling,
    raw little-endian uint16 output that np.memmap reads (train.py:230-232 of the reference)."""
    from models.tokenizer.encode import encode_file
    vocab, merges = _trained(1000, [EOT])
    tok = get_tokenizer(dict(vocab), list(merges), [EOT])
    body = (FIXTURES_PATH / "tinystories_sample.txt").read_bytes() + (FIXTURES_PATH / "corpus.en").read_bytes()[:40000]
    raw = (body + b"\r\nline two\rline three\r\n" + EOT.encode()) * 12 + b"tail without newline"
    src = tmp_path / "in.txt"
    src.write_bytes(raw)
    with open(src, "r", encoding="utf-8") as f:
        text = f.read()                                   # what the reference's text-mode read sees
    want = oracle.OracleTokenizer(dict(vocab), list(merges), [EOT]).encode(text)   # the oracle, not the GPU's own encode
    assert tok.encode(text) == want
    for piece in (1 << 30, 70000, 9000):                  # one piece, several, many (forces carries and "\r" holds)
        dst = tmp_path / ("out_%d.bin" % piece)
        n = encode_file(tok, src, dst, np.uint16, piece_bytes=piece)
        got = np.memmap(dst, dtype=np.uint16, mode="r")
        assert n == len(want) == got.size
        assert got.tolist() == want
    # no specials in the tokenizer: pieces are cut at lone spaces
    tok2 = get_tokenizer(dict(vocab), list(merges), [])
    dst = tmp_path / "out_nosp.bin"
    encode_file(tok2, src, dst, np.int32, piece_bytes=5000)
    assert np.fromfile(dst, dtype="<i4").tolist() == oracle.OracleTokenizer(dict(vocab), list(merges), []).encode(text)
