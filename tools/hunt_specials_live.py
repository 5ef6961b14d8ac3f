"""One-off hunt (build container only, needs /root/reference): Tokenizer.encode of the unmodified reference against the oracle with nasty
special-token sets (regex metacharacters, white space, overlapping and nested specials, specials missing from the vocab).  Run with a fixed
PYTHONHASHSEED (the reference orders equal-length specials by string hash):
    PYTHONHASHSEED=0 python tools/hunt_specials_live.py <first seed> <number of seeds>"""
import sys, random, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import tests.test_oracle_vs_reference_live as T
from tests.helpers import live_shapes
from oracle import oracle
from tests.common import FIXTURES_PATH
import tempfile
tmp = pathlib.Path(tempfile.mkdtemp(prefix='hunt_sp_'))
POOL = ["<|endoftext|>", "<|endoftext|><|endoftext|>", "<|pad|>", "ab", "bc", "abc", "a", " ", "  ", "\n", ".", "(?:x)", "[a-z]+", "a|b", "\\", "^", "$",
        "é", "日本", "🙃", "he", " the", "<|", "|>", "<|a|>", "<|a|><|b|>", "<|b|>", "'s", "1", "12", "\r\n", "\t"]
seed0 = int(sys.argv[1]); n = int(sys.argv[2])
bad = 0
for seed in range(seed0, seed0 + n):
    r = random.Random(seed)
    corpus = (FIXTURES_PATH / "corpus.en").read_bytes()[:30000] + live_shapes.text(r, 500, False).encode()
    vocab, merges = oracle.train_bpe_on_bytes(corpus, r.choice([300, 500, 800]), [])
    set_ups = []
    for _ in range(6):
        sp = r.sample(POOL, r.randint(1, 5))
        if r.random() < 0.3: sp = sp + [sp[0]]
        set_ups.append((dict(vocab), merges, sp))
    texts = []
    for _ in range(25):
        parts = []
        for _ in range(r.randint(1, 60)):
            parts.append(r.choice(POOL + live_shapes.WORDS + live_shapes.SEPS + ["abcabc", "ababc", "bcab", "x"]))
        texts.append("".join(parts))
    jobs = [{"kind": "encode", "vocab": {str(k): v.hex() for k, v in v.items()}, "merges": [[a.hex(), b.hex()] for a, b in m],
             "special_tokens": sp, "texts": texts} for v, m, sp in set_ups]
    ref = T._run_reference(tmp, jobs)
    for (v, m, sp), rr in zip(set_ups, ref):
        tok = oracle.OracleTokenizer(dict(v), list(m), list(sp))
        for text, want in zip(texts, rr):
            try:
                got = {"ids": tok.encode(text)}
            except KeyError as e:
                got = {"error": "KeyError", "arg": e.args[0].hex() if isinstance(e.args[0], bytes) else repr(e.args[0])}
            w = {k: want[k] for k in want if k in ("ids", "error", "arg")}
            if got != w:
                bad += 1
                print("MISMATCH seed", seed, sp, repr(text[:120]), "\n want", str(w)[:300], "\n got ", str(got)[:300])
                break
print("done", seed0, n, "mismatches", bad)
