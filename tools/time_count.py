"""GPU: time the pretokenise + count stages alone (no merges).  usage: python tools/time_count.py <shape> <seed> <bytes> [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
shape, seed, n = sys.argv[1], int(sys.argv[2]), int(float(sys.argv[3])) // 4096 * 4096
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
import _bootstrap, torch
from transformer_lm_b200 import _lib
from transformer_lm_b200.synth import synth_device
from transformer_lm_b200.train import train_bpe_on_bytes
ctx = _lib.default_context(0)
t = torch.empty(n, dtype=torch.uint8, device='cuda')
synth_device(shape, seed, n, t.data_ptr(), ctx=ctx)
for it in range(reps):
    v, m, st = train_bpe_on_bytes(None, 257, ["<|endoftext|>"], ctx=ctx, return_stats=True, device_ptr=t.data_ptr(), n_bytes=n)
print(shape, n, {k: round(x, 1) for k, x in st.items() if k.startswith('ms_')}, "pretokens", st.get("n_pretokens"), "unique", st.get("n_unique"))
