#!/usr/bin/env python3
"""Generate tests/golden/*.json by running the UNMODIFIED reference (/root/reference) here.

The reference is pure Python and cannot travel to the GPU box, so its outputs on a fixed set of
inputs are committed as golden vectors.  Run from the repo root in the build container:

    python tools/make_golden.py

The reference is executed in a subprocess (cwd=/root/reference) because its package is called
`models`, like ours.  Inputs are deterministic (fixed seeds / fixtures); bytes are stored as hex.
"""
import json
import pathlib
import random
import subprocess
import sys
import tempfile

ROOT = pathlib.Path(__file__).resolve().parent.parent
REF = pathlib.Path("/root/reference")
FIX = ROOT / "tests" / "fixtures"
OUT = ROOT / "tests" / "golden"

WORKER = r'''
import json, sys, logging
sys.path.insert(0, "/root/reference")
logging.disable(logging.CRITICAL)
import tqdm
from models.tokenizer import train as T
T.tqdm = lambda x, *a, **k: x
from models.tokenizer.tokenizer import Tokenizer
job = json.load(open(sys.argv[1]))
out = {}
if job["kind"] == "train":
    try:
        vocab, merges = T.train_bpe(job["path"], job["vocab_size"], job["special_tokens"])
        out = {"vocab": {str(k): v.hex() for k, v in vocab.items()},
               "merges": [[a.hex(), b.hex()] for a, b in merges]}
    except Exception as e:
        out = {"error": type(e).__name__}
elif job["kind"] == "encode":
    vocab = {int(k): bytes.fromhex(v) for k, v in job["vocab"].items()}
    merges = [(bytes.fromhex(a), bytes.fromhex(b)) for a, b in job["merges"]]
    tok = Tokenizer(vocab, merges, job["special_tokens"])
    res = []
    for text in job["texts"]:
        try:
            ids = tok.encode(text)
            res.append({"ids": ids, "decoded": tok.decode(ids)})
        except KeyError as e:
            res.append({"error": "KeyError", "arg": e.args[0].hex() if isinstance(e.args[0], bytes) else repr(e.args[0])})
    out = {"results": res}
    if job.get("iterable_lines") is not None:
        out["iterable_ids"] = list(tok.encode_iterable(iter(job["iterable_lines"])))
json.dump(out, open(sys.argv[2], "w"))
'''


def run_ref(job):
    with tempfile.TemporaryDirectory() as td:
        td = pathlib.Path(td)
        (td / "worker.py").write_text(WORKER)
        (td / "job.json").write_text(json.dumps(job))
        subprocess.check_call([sys.executable, str(td / "worker.py"), str(td / "job.json"), str(td / "out.json")],
                              cwd=str(REF), env={"PYTHONDONTWRITEBYTECODE": "1", "PATH": "/usr/bin:/bin"})
        return json.loads((td / "out.json").read_text())


def fuzz_text(seed, n_words, flavor):
    """Small deterministic corpora exercising the parity hazards in SURVEY Appendix A."""
    r = random.Random(seed)
    if flavor == "tiny_alpha":
        return "".join(r.choice("ab ") for _ in range(n_words))
    words = ["the", "a", "cat", "dog", "sat", "on", "mat", "it's", "they'll", "we've", "don't", "I'm",
             "naïve", "café", "Zürich", "日本語", "テスト", "привет", "мир", "🙃", "👍🏽", "3.14", "2024", "1,000",
             "http://x.y/z?q=1", "a_b", "--", "...", "!?", "<|endoftext|>", "x²", "١٢٣", "é"]
    seps = [" ", " ", " ", " ", "  ", "\n", "\n\n", "\t", " \n", " ", " ", "   ", ""]
    if flavor == "crlf":
        seps = [" ", "\r\n", "\r", "\n", " \r\n ", "\r\r\n"]
    out = []
    for _ in range(n_words):
        w = r.choice(words)
        if r.random() < 0.2:
            w = w.capitalize()
        out.append(w)
        out.append(r.choice(seps))
        if r.random() < 0.05:
            out.append(r.choice([".", ",", "!", "?", ";", "'", "\"", "'s", "'re"]))
    return "".join(out)


TRAIN_CASES = [
    # name, text-source, vocab_size, specials
    ("corpus_500_eot", "fixture:corpus.en", 500, ["<|endoftext|>"]),
    ("corpus_1000_eot", "fixture:corpus.en", 1000, ["<|endoftext|>"]),
    ("corpus_300_special_he", "fixture:corpus.en", 300, ["he"]),
    ("corpus_300_special_the", "fixture:corpus.en", 300, [" the"]),
    ("corpus_260_two_specials", "fixture:corpus.en", 300, ["<|endoftext|>", "<|pad|>", "<|endoftext|>"]),
    ("tinystories_400", "fixture:tinystories_sample.txt", 400, ["<|endoftext|>"]),
    ("german_330", "fixture:german.txt", 330, []),
    ("zero_phase_aaaa", "literal:aaaa abab aaaa", 268, []),
    ("exhaust_tiny", "literal:ab ab abc", 300, []),
    ("empty_file", "literal:", 300, ["<|endoftext|>"]),
    ("single_char", "literal:a", 300, []),
    ("no_merges", "literal:hello hello", 256, []),
    ("fewer_than_base", "literal:hello hello", 10, ["<|endoftext|>"]),
    ("tiny_alpha_s1", "fuzz:1:400:tiny_alpha", 290, []),
    ("tiny_alpha_s2", "fuzz:2:60:tiny_alpha", 300, []),
    ("mixed_s3", "fuzz:3:3000:mixed", 700, ["<|endoftext|>"]),
    ("mixed_s4_exhaust", "fuzz:4:40:mixed", 1200, ["<|endoftext|>"]),
    ("crlf_s5", "fuzz:5:500:crlf", 400, []),
    ("invalid_utf8", "hex:6162ff6364", 300, []),
    ("truncated_utf8", "hex:6162e282", 300, []),
]

ENCODE_TEXTS = [
    "", "s", "🙃", "Hello, how are you?", "Héllò hôw are ü? 🙃",
    "Héllò hôw <|endoftext|><|endoftext|> are ü? 🙃<|endoftext|>",
    "Hello, how <|endoftext|><|endoftext|> are you?<|endoftext|>",
    "it's they'll we've I'm don't 'sabc x'llama !'s  's \n's",
    "a  b   c \n d\n\n\ne \t f  \n  g   ",
    "trailing spaces   ", "   leading", "\n\n", " ", "  ",
    "1234567890 3.14159 1,000,000 ١٢٣ x²",
    "日本語のテキスト、テスト。 Привет мир! naïve café é",
    "<|endoftext|>", "<|endoftext|><|endoftext|><|endoftext|>", "<|endoftext", "a<|endoftext|>'s",
    "nbsp\u00a0here\u2009thin\u3000wide\x85nel\x1cfs",
]


def source_bytes(src):
    kind, _, rest = src.partition(":")
    if kind == "fixture":
        return (FIX / rest).read_bytes()
    if kind == "literal":
        return rest.encode("utf-8")
    if kind == "hex":
        return bytes.fromhex(rest)
    if kind == "fuzz":
        seed, n, flavor = rest.split(":")
        return fuzz_text(int(seed), int(n), flavor).encode("utf-8")
    raise ValueError(src)


def gpt2_vocab_merges():
    sys.path.insert(0, str(ROOT))
    from tests.common import load_gpt2_fixture
    return load_gpt2_fixture()


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    # ---- training goldens ------------------------------------------------
    train = {}
    with tempfile.TemporaryDirectory() as td:
        for name, src, vs, sp in TRAIN_CASES:
            data = source_bytes(src)
            p = pathlib.Path(td) / (name + ".txt")
            p.write_bytes(data)
            res = run_ref({"kind": "train", "path": str(p), "vocab_size": vs, "special_tokens": sp})
            entry = {"source": src, "vocab_size": vs, "special_tokens": sp}
            if not src.startswith("fixture:"):
                entry["input_hex"] = data.hex()
            entry.update(res)
            train[name] = entry
            print("train", name, "merges=%s" % (len(res.get("merges", [])) if "merges" in res else res))
    (OUT / "train_golden.json").write_text(json.dumps(train, indent=0, sort_keys=True))

    # ---- encode goldens ---------------------------------------------------
    enc = {}
    vocab, merges = gpt2_vocab_merges()
    vocab_eot = dict(vocab)
    texts = list(ENCODE_TEXTS)
    for f in ["address.txt", "german.txt", "tinystories_sample.txt"]:
        texts.append((FIX / f).read_text(encoding="utf-8"))
    texts.append((FIX / "corpus.en").read_text(encoding="utf-8")[:20000])
    texts.append(fuzz_text(11, 800, "mixed"))
    job = {"kind": "encode", "vocab": {str(k): v.hex() for k, v in vocab_eot.items()},
           "merges": [[a.hex(), b.hex()] for a, b in merges], "special_tokens": ["<|endoftext|>"], "texts": texts,
           "iterable_lines": (FIX / "tinystories_sample.txt").read_text(encoding="utf-8").splitlines(keepends=True)}
    res = run_ref(job)
    enc["gpt2_eot"] = {"tokenizer": "gpt2_fixture", "special_tokens": ["<|endoftext|>"], "texts": texts,
                       "results": res["results"], "iterable_source": "tinystories_sample.txt",
                       "iterable_ids": res["iterable_ids"]}
    print("encode gpt2_eot", len(texts), "texts")
    # overlapping specials (tests/test_tokenizer.py:255-269)
    sp2 = ["<|endoftext|>", "<|endoftext|><|endoftext|>"]
    v2 = dict(vocab)
    v2[len(v2)] = sp2[1].encode()
    t2 = ["Hello, how <|endoftext|><|endoftext|> are you?<|endoftext|>", "<|endoftext|><|endoftext|><|endoftext|>"]
    res = run_ref({"kind": "encode", "vocab": {str(k): v.hex() for k, v in v2.items()},
                   "merges": [[a.hex(), b.hex()] for a, b in merges], "special_tokens": sp2, "texts": t2})
    enc["gpt2_overlapping_specials"] = {"tokenizer": "gpt2_fixture+double_eot", "special_tokens": sp2, "texts": t2,
                                        "results": res["results"]}
    # no specials at all (special_tokens=None)
    res = run_ref({"kind": "encode", "vocab": {str(k): v.hex() for k, v in vocab.items()},
                   "merges": [[a.hex(), b.hex()] for a, b in merges], "special_tokens": None,
                   "texts": ENCODE_TEXTS})
    enc["gpt2_no_specials"] = {"tokenizer": "gpt2_fixture", "special_tokens": None, "texts": ENCODE_TEXTS,
                               "results": res["results"]}
    # a small trained tokenizer (vocab from the reference itself), incl. a vocab with a hole -> KeyError
    tv = train["corpus_1000_eot"]
    small_vocab = dict(tv["vocab"])
    t3 = ENCODE_TEXTS + [fuzz_text(12, 500, "mixed"), (FIX / "corpus.en").read_text(encoding="utf-8")[5000:9000]]
    res = run_ref({"kind": "encode", "vocab": small_vocab, "merges": tv["merges"], "special_tokens": ["<|endoftext|>"],
                   "texts": t3})
    enc["corpus1000"] = {"tokenizer": "train_golden:corpus_1000_eot", "special_tokens": ["<|endoftext|>"], "texts": t3,
                         "results": res["results"]}
    holed = {k: v for k, v in small_vocab.items() if bytes.fromhex(v) not in (b" the", b"e")}
    res = run_ref({"kind": "encode", "vocab": holed, "merges": tv["merges"], "special_tokens": ["<|endoftext|>"],
                   "texts": ["in the end", "xyz", "qqq e"]})
    enc["corpus1000_holed"] = {"tokenizer": "train_golden:corpus_1000_eot", "drop_tokens_hex": [b" the".hex(), b"e".hex()],
                               "special_tokens": ["<|endoftext|>"], "texts": ["in the end", "xyz", "qqq e"],
                               "results": res["results"]}
    (OUT / "encode_golden.json").write_text(json.dumps(enc, indent=0, sort_keys=True))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
