#!/usr/bin/env python3
"""Fuzz the oracle's GPT-2 pretokenizer (oracle/bpe_oracle.c: gpt2_match_len) against the
installed `regex` module, which is what the reference calls (train.py:143-146)."""
import pathlib
import random
import sys
import time

import regex

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle  # noqa: E402

PAT = regex.compile(r"""'(?:[sdmt]|ll|ve|re)| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""")

ALPHA = list("'sdmtlvre ab1.\n\t") + [
    "\u00a0", "\u2009", "\x85", "\x1c", "\u4e2d", "\U0001F643", "\u00e9", "\u0661", "<", "|", ">",
    "'ll", "'ve", "'re", "  ", "<|endoftext|>", "\r", "\u3000", "\u00b2", "\u0301",
]


def main(iters=200000, seed=1):
    rnd = random.Random(seed)
    t = time.time()
    bad = 0
    for _ in range(iters):
        s = "".join(rnd.choice(ALPHA) for _ in range(rnd.randint(0, 14)))
        if oracle.pretokens(s) != PAT.findall(s):
            bad += 1
            if bad < 5:
                print(repr(s), oracle.pretokens(s), PAT.findall(s))
    print("mismatches", bad, "of", iters, "in %.1fs" % (time.time() - t))
    for f in ["corpus.en", "address.txt", "german.txt", "tinystories_sample.txt"]:
        s = open(ROOT / "tests" / "fixtures" / f, encoding="utf-8").read()
        print(f, oracle.pretokens(s) == PAT.findall(s))
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(*(int(a) for a in sys.argv[1:])) else 0)
