"""Seeded inputs shared by the live reference differential (tests/test_oracle_vs_reference_live.py, CPU) and its GPU counterpart
(tests/test_gpu_live_shapes.py): small corpora and texts mixing scripts, contractions, digits, odd white space (NBSP, thin space, CR /
CRLF), special tokens that are ordinary strings, and -- for training -- invalid UTF-8."""
import random

from oracle import oracle
from tests.common import FIXTURES_PATH

WORDS = ["the", "a", "cat", "it's", "they'll", "we've", "don't", "I'm", "naïve", "café", "日本語", "テスト", "привет", "🙃", "👍🏽", "3.14",
         "2024", "1,000", "http://x.y/z?q=1", "a_b", "--", "...", "!?", "<|endoftext|>", "<|pad|>", "x²", "١٢٣", "é", "aaa", "abab", "he"]
SEPS = [" ", " ", " ", "  ", "\n", "\n\n", "\t", " \n", "   ", "", " ", " ", "\r\n", "\r"]


def text(r, n_words, crlf):
    seps = SEPS if crlf else SEPS[:-2]
    out = []
    for _ in range(n_words):
        w = r.choice(WORDS)
        out.append(w.capitalize() if r.random() < 0.2 else w)
        out.append(r.choice(seps))
        if r.random() < 0.05:
            out.append(r.choice([".", ",", "'", "\"", "'s", "'re", "'"]))
    return "".join(out)



def train_cases(seed):
    """[(data, vocab_size, special_tokens)]: 28 corpora; every fourth one has a (usually) invalid byte planted."""
    r = random.Random(986 + seed)
    cases = []
    for k in range(28):
        flavour = k % 4
        if flavour == 0:
            data = "".join(r.choice("ab c") for _ in range(r.randint(1, 300))).encode()
        elif flavour == 3:
            data = bytearray(text(r, r.randint(5, 200), False).encode())
            data[r.randrange(len(data))] = r.choice([0xFF, 0xC0, 0xE2, 0x80, 0xF5])
            data = bytes(data)
        else:
            data = text(r, r.randint(1, 700), crlf=flavour == 2).encode()
        specials = r.choice([[], ["<|endoftext|>"], ["<|endoftext|>", "<|pad|>"], ["he"], [" the", "<|endoftext|>", "<|endoftext|>"]])
        vocab_size = r.choice([0, 257, 270, 300, 400, 600, 2000])
        cases.append((data, vocab_size, specials))
    return cases


def encode_cases(seed):
    """([(vocab, merges, special_tokens)], texts): three trained tokenizers (one with specials the vocab lacks, A-12) and one with
    holes in its vocab (KeyErrors, tokenizer.py:120,135)."""
    r = random.Random(4711 + seed)
    corpus = (FIXTURES_PATH / "corpus.en").read_bytes()[:80000] + text(r, 2000, False).encode()
    set_ups = []
    for vocab_size, specials in ((700, ["<|endoftext|>"]), (400, []), (900, ["<|endoftext|>", "<|endoftext|><|endoftext|>", "<|pad|>"])):
        vocab, merges = oracle.train_bpe_on_bytes(corpus, vocab_size, specials[:1])
        set_ups.append((vocab, merges, specials))
    holed = {k: v for k, v in set_ups[0][0].items() if v not in (b" the", b"e", b"\xf0")}
    set_ups.append((holed, set_ups[0][1], ["<|endoftext|>"]))
    texts = ["", " ", "a", "🙃"] + [text(r, r.randint(1, 300), crlf=k % 3 == 0) for k in range(30)]
    return set_ups, texts
