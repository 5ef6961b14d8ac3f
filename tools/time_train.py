"""GPU: time train_bpe stages on a synthetic corpus.  usage: python tools/time_train.py <shape> <seed> <bytes> <vocab> [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
shape, seed, n, vocab = sys.argv[1], int(sys.argv[2]), int(float(sys.argv[3])) // 4096 * 4096, int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
import _bootstrap, torch, hashlib
from transformer_lm_b200 import _lib
from transformer_lm_b200.synth import synth_device
from transformer_lm_b200.train import train_bpe_on_bytes
ctx = _lib.default_context(0)
t = torch.empty(n, dtype=torch.uint8, device='cuda')
synth_device(shape, seed, n, t.data_ptr(), ctx=ctx)
for it in range(reps):
    v, m, st = train_bpe_on_bytes(None, vocab, ["<|endoftext|>"], ctx=ctx, return_stats=True, device_ptr=t.data_ptr(), n_bytes=n)
h = hashlib.sha256(b"".join(a + b"\x00" + b + b"\x01" for a, b in m)).hexdigest()[:16]
print(shape, n, vocab, "merges", len(m), "steps", st["merge_steps"], "sha", h, {k: round(x, 1) for k, x in st.items() if k.startswith('ms_')},
      "us/merge %.2f us/step %.2f" % (1e3 * st["ms_merge"] / max(len(m), 1), 1e3 * st["ms_merge"] / max(st["merge_steps"], 1)))
