"""One-off hunt (build container only, needs /root/reference): train_bpe of the unmodified reference against the oracle with special tokens that
are ordinary strings (whole pretokens such as " the", "  ", "12"; single bytes that collide with the base vocab, SURVEY A-5), odd vocab sizes and
CR / CRLF corpora.    PYTHONHASHSEED=0 python tools/hunt_train_live.py <first seed> <number of seeds>     (120 cases run once: identical)"""
import sys, random, pathlib, json, tempfile
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import tests.test_oracle_vs_reference_live as T
from tests.helpers import live_shapes
from oracle import oracle
POOL = ["<|endoftext|>", "<|pad|>", "ab", "bc", "a", " ", "  ", "\n", "\n\n", ".", "é", "日本語", "🙃", "he", " the", " a", "'s", "1", "12", " 12", "\t", "!?", " !?", "...", "aaa", " aaa", "abab"]
seed0 = int(sys.argv[1]); n = int(sys.argv[2])
tmp = pathlib.Path(tempfile.mkdtemp(prefix='hunt_tr_'))
bad = 0
for seed in range(seed0, seed0 + n):
    r = random.Random(seed)
    jobs, inputs = [], []
    for k in range(20):
        if k % 3 == 0:
            data = "".join(r.choice(["a", "b", " ", "ab", "aaa", "\n"]) for _ in range(r.randint(1, 200))).encode()
        else:
            data = live_shapes.text(r, r.randint(1, 400), crlf=k % 3 == 2).encode()
        sp = r.sample(POOL, r.randint(0, 5))
        if sp and r.random() < 0.3: sp = sp + [sp[0]]
        vs = r.choice([0, 256, 258, 262, 300, 500, 3000])
        p = tmp / ("c%d_%d.txt" % (seed, k)); p.write_bytes(data)
        jobs.append({"kind": "train", "path": str(p), "vocab_size": vs, "special_tokens": sp}); inputs.append(data)
    for job, data, ref in zip(jobs, inputs, T._run_reference(tmp, jobs)):
        vocab, merges = oracle.train_bpe_on_bytes(data, job["vocab_size"], job["special_tokens"])
        if [[a.hex(), b.hex()] for a, b in merges] != ref["merges"] or {str(k): v.hex() for k, v in vocab.items()} != ref["vocab"]:
            bad += 1
            print("MISMATCH seed", seed, job["special_tokens"], job["vocab_size"], repr(data[:100]))
print("done", seed0, n, "mismatches", bad)
