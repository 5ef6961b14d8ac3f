"""Host-side logic that needs no GPU: the Tokenizer constructor's reference quirks, the integer-symbol tables handed
to bpe_tok_create, Vocab semantics, and the loud failure of every compute entry point without a device."""
import pytest

import _bootstrap  # noqa: F401
from tests.common import load_gpt2_fixture
from transformer_lm_b200 import _lib
from transformer_lm_b200.tokenizer import Tokenizer
from transformer_lm_b200.vocab import Vocab


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


def test_constructor_mirrors_reference_quirks():
    vocab = {i: bytes([i]) for i in range(256)}
    tok = Tokenizer(vocab, [], ["<|a|>", "<|longer|>", "<|a|>"])
    assert tok.vocab is vocab                                      # keeps and mutates the caller's dict (tokenizer.py:18,37)
    assert tok.special_tokens == ["<|longer|>", "<|a|>"]            # deduped, longest first (29-30)
    assert tok.vocab_inv[b"<|longer|>"] == 256 and tok.vocab_inv[b"<|a|>"] == 257
    assert vocab[b"<|longer|>"] == 256                              # key/value swapped like the reference (A-12)
    assert Tokenizer(dict(vocab), [], None).special_tokens == []


def test_symbol_tables_dedupe_and_rank_override():
    vocab = {i: bytes([i]) for i in range(256)}
    vocab[256] = b"ab"; vocab[300] = b"abc"
    merges = [(b"a", b"b"), (b"b", b"c"), (b"ab", b"c"), (b"a", b"b"), (b"zz", b"q"), (b"a", b"bc")]
    tok = Tokenizer(vocab, merges, [])
    pairs, result, sym_bytes, sym_to_id = tok._tables()
    assert sym_bytes[:256] == [bytes([i]) for i in range(256)]
    assert sym_bytes[256:] == [b"ab", b"bc", b"abc", b"zzq"]       # distinct products in first-seen order; abc only once
    assert pairs[0].tolist() == [-1, -1]                            # superseded by the later duplicate (last index wins)
    assert pairs[3].tolist() == [97, 98] and result[3] == 256
    assert pairs[2].tolist() == [256, 99] and result[2] == 258
    assert pairs[4].tolist() == [-1, -1]                            # b"zz" is not a symbol: can never be adjacent
    assert pairs[5].tolist() == [97, 257] and result[5] == 258      # a second way to spell b"abc": same symbol
    assert sym_to_id[256] == 256 and sym_to_id[258] == 300 and sym_to_id[257] == -1 and sym_to_id[259] == -1


def test_symbol_tables_scale_to_gpt2():
    vocab, merges = load_gpt2_fixture()
    tok = Tokenizer(vocab, merges, ["<|endoftext|>"])
    pairs, result, sym_bytes, sym_to_id = tok._tables()
    assert len(sym_bytes) == 256 + len(merges) and (pairs >= 0).all() and (sym_to_id >= 0).all()


def test_vocab_semantics():
    v = Vocab(["<|endoftext|>", "a"])
    assert v.idx_to_token[0] == b"<|endoftext|>" and v.idx_to_token[1] == b"a"
    assert len(v) == 257                                            # b"a" appears once: the byte 0x61 is skipped (A-5)
    n = len(v)
    v.add_token(b"a")
    assert len(v) == n
    v.add_token(b"zz")
    assert v.idx_to_token[n] == b"zz" and v.get_inv()[b"zz"] == n


def test_merge_helper():
    tok = Tokenizer({i: bytes([i]) for i in range(256)}, [], [])
    assert tok.merge([b"a", b"a", b"a", b"b", b"a", b"a"], (b"a", b"a"), b"aa") == [b"aa", b"a", b"b", b"aa"]
    assert tok.merge([], (b"a", b"b"), b"ab") == []


def test_compute_fails_loudly_without_a_device():
    if not _no_gpu():
        pytest.skip("GPU present")
    tok = Tokenizer({i: bytes([i]) for i in range(256)}, [], [])
    with pytest.raises(_lib.BpeError):
        tok.encode("hello")
    with pytest.raises(_lib.BpeError):
        tok.decode([104])
    with pytest.raises(_lib.BpeError):
        from transformer_lm_b200.pretok import pretoken_starts
        pretoken_starts(b"hello world")


def test_load_batch_shim_has_the_reference_signature_and_no_cpu_path():
    # models/util.py:37-43 of the reference: load_batch(dataset, batch_size, context_length, device, generator=None)
    import inspect
    import numpy as np
    from models.util import load_batch
    from transformer_lm_b200 import _lib
    params = list(inspect.signature(load_batch).parameters)
    assert params[:5] == ["dataset", "batch_size", "context_length", "device", "generator"]
    with pytest.raises(_lib.BpeError):          # the gather runs on a B200 or not at all
        load_batch(np.arange(100, dtype=np.uint16), 2, 8, "cpu")


def test_merge_builder_in_pieces_equals_one_piece_and_dedupes_like_the_reference():
    """MergeBuilder (what the live follower of the merge loop feeds): symbol-id pairs -> (vocab, merges) of the reference, in one
    piece or in many, including a merge whose bytes already exist (a special token's bytes, or the same bytes made twice: Vocab.add_token
    skips them, models/tokenizer/vocab.py:28-34)."""
    import numpy as np
    from transformer_lm_b200.train import MergeBuilder, merges_to_python
    from transformer_lm_b200.vocab import Vocab
    rng = np.random.default_rng(7)
    n = 3000
    pairs = np.zeros((n, 2), dtype=np.int32)
    for i in range(n):
        pairs[i] = rng.integers(0, min(256 + i, 900), 2)
    pairs[5] = (ord("<"), ord("|"))                       # b"<|" is a special token below: already present
    pairs[9] = pairs[7]                                   # the same bytes twice
    specials = ["<|", "<|endoftext|>"]

    def reference_way():
        vocab = Vocab(special_tokens=list(specials))
        sym = [bytes([i]) for i in range(256)]
        merges = []
        for a, b in pairs.tolist():
            merges.append((sym[a], sym[b]))
            sym.append(sym[a] + sym[b])
            vocab.add_token(sym[a] + sym[b])
        return vocab.get_idx_to_token(), merges

    want = reference_way()
    assert merges_to_python(Vocab(special_tokens=list(specials)), pairs, n) == want
    b = MergeBuilder(Vocab(special_tokens=list(specials)))
    cuts = [0, 1, 2, 6, 7, 10, 700, 701, 2999, 3000]
    for lo, hi in zip(cuts, cuts[1:]):
        b.feed(pairs[lo:hi])
    assert b.n_fed == n and b.result() == want
    assert len(want[0]) < 256 + len(specials) + n         # the duplicates were skipped


def test_numa_binding_is_best_effort():
    """_lib.bind_to_gpu_numa_node never raises: without a GPU / sysfs it reports why it left the process where it was."""
    import os
    from transformer_lm_b200 import _lib
    before = os.sched_getaffinity(0)
    info = _lib.bind_to_gpu_numa_node(0)
    assert isinstance(info, dict) and "bound" in info
    if not info["bound"]:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
