"""GPU parity on the inputs of the live reference differential (tests/helpers/live_shapes.py): the oracle agrees with the unmodified
reference on exactly these seeded corpora and texts in the build container (tests/test_oracle_vs_reference_live.py); here the product
(bpe_train / bpe_tok_create / bpe_encode / bpe_decode through the C ABI) must agree with the oracle on them -- merges, vocab,
UnicodeDecodeError positions, ids, KeyError arguments, the decode quirk of specials missing from the vocab (SURVEY A-12)."""
import os

import pytest

import _bootstrap  # noqa: F401
from oracle import oracle
from tests.adapters import get_tokenizer
from tests.helpers import live_shapes

pytestmark = pytest.mark.gpu
SEEDS = [int(k) for k in os.environ.get("BPE_LIVE_GPU_SEEDS", "0,1").split(",")]      # (longer hunts: more seeds)


@pytest.mark.parametrize("seed", SEEDS)
def test_train_bpe_on_the_live_shapes(seed):
    from models.tokenizer.train import train_bpe_on_bytes
    n_err = 0
    for data, vocab_size, specials in live_shapes.train_cases(seed):
        try:
            want = oracle.train_bpe_on_bytes(data, vocab_size, specials)
        except UnicodeDecodeError as e:
            n_err += 1
            with pytest.raises(UnicodeDecodeError) as got:
                train_bpe_on_bytes(data, vocab_size, specials)
            assert (got.value.start, got.value.end, got.value.reason) == (e.start, e.end, e.reason), data
            continue
        vocab, merges = train_bpe_on_bytes(data, vocab_size, specials)
        assert merges == want[1], (data, vocab_size, specials)
        assert vocab == want[0], (data, vocab_size, specials)
    assert n_err >= 1


@pytest.mark.parametrize("seed", SEEDS)
def test_tokenizer_on_the_live_shapes(seed):
    set_ups, texts = live_shapes.encode_cases(seed)
    n_key = n_dec = 0
    for vocab, merges, specials in set_ups:
        tok = get_tokenizer(dict(vocab), list(merges), list(specials))
        otok = oracle.OracleTokenizer(dict(vocab), list(merges), list(specials))
        for text in texts:
            try:
                want = otok.encode(text)
            except KeyError as e:
                n_key += 1
                with pytest.raises(KeyError) as got:
                    tok.encode(text)
                assert got.value.args[0] == e.args[0], text
                continue
            assert tok.encode(text) == want, (specials, text)
            try:
                back = otok.decode(want)
            except KeyError as e:
                n_dec += 1
                with pytest.raises(KeyError) as got:
                    tok.decode(want)
                assert got.value.args[0] == e.args[0], text
                continue
            assert tok.decode(want) == back, (specials, text)
    assert n_key >= 1 and n_dec >= 1
