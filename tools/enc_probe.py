"""Probe: train a vocab on a synthetic OWT-shape slice, bulk-encode a larger slice resident in HBM, print stage times."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import _bootstrap  # noqa: F401,E402
from transformer_lm_b200 import _lib  # noqa: E402
from transformer_lm_b200.synth import synth_device  # noqa: E402
from transformer_lm_b200.tokenizer import Tokenizer  # noqa: E402
from transformer_lm_b200.train import train_bpe_on_bytes  # noqa: E402

train_bytes = int(float(sys.argv[1])) if len(sys.argv) > 1 else 64 << 20
enc_bytes = int(float(sys.argv[2])) if len(sys.argv) > 2 else 512 << 20
vocab_size = int(sys.argv[3]) if len(sys.argv) > 3 else 32000
ctx = _lib.default_context(0)
L = _lib.lib()
t = torch.empty(train_bytes, dtype=torch.uint8, device="cuda")
synth_device("owt", 4321, train_bytes, t.data_ptr(), ctx=ctx)
t0 = time.time()
vocab, merges, st = train_bpe_on_bytes(None, vocab_size, ["<|endoftext|>"], ctx=ctx, return_stats=True, device_ptr=t.data_ptr(), n_bytes=train_bytes)
print("train", round(time.time() - t0, 3), "s", {k: round(v, 2) if isinstance(v, float) else v for k, v in st.items()})
tok = Tokenizer(vocab, merges, ["<|endoftext|>"], ctx=ctx)
h = tok._device_tok()
x = torch.empty(enc_bytes, dtype=torch.uint8, device="cuda")
synth_device("owt", 4322, enc_bytes, x.data_ptr(), ctx=ctx)
out = torch.empty(enc_bytes, dtype=torch.uint16, device="cuda")
for it in range(3):
    n_out = C.c_uint64(0)
    stats = _lib.EncodeStats()
    L.bpe_tok_cache_reset(h)                     # like bench.py: no step reuses cached BPE results
    t0 = time.time()
    rc = L.bpe_encode_dev(h, C.c_void_p(x.data_ptr()), enc_bytes, 0, C.c_void_p(out.data_ptr()), enc_bytes, C.byref(n_out), C.byref(stats))
    ctx.check(rc)
    dt = time.time() - t0
    print("encode iter", it, round(dt * 1e3, 1), "ms", round(enc_bytes / 1e6 / dt, 1), "MB/s", {k: round(v, 2) if isinstance(v, float) else v for k, v in stats.as_dict().items()})
# parity on a 2 MB prefix vs the oracle
from oracle import oracle  # noqa: E402
n = 2 << 20
host = x[:n].cpu().numpy().tobytes()
# cut at a document boundary so the prefix tokenises identically
cut = host.rfind(b"<|endoftext|>")
want = oracle.OracleTokenizer(dict(vocab), list(merges), ["<|endoftext|>"]).encode_bytes(host[:cut])
got = tok.encode_to_numpy(host[:cut], np.int32)
print("parity on", cut, "bytes:", bool((got == want).all()) and len(got) == len(want), len(got))
