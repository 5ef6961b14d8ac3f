"""GPU: train on a synthetic corpus and dump (a, b, count) per merge (BPE_DUMP_MERGES) for offline batching analysis.
usage: python tools/dump_merges.py <shape> <seed> <bytes> <vocab> <out>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
shape, seed, n, vocab, out = sys.argv[1], int(sys.argv[2]), int(float(sys.argv[3])) // 4096 * 4096, int(sys.argv[4]), sys.argv[5]
os.environ["BPE_DUMP_MERGES"] = out
import _bootstrap, torch
from transformer_lm_b200 import _lib
from transformer_lm_b200.synth import synth_device
from transformer_lm_b200.train import train_bpe_on_bytes
ctx = _lib.default_context(0)
t = torch.empty(n, dtype=torch.uint8, device='cuda')
synth_device(shape, seed, n, t.data_ptr(), ctx=ctx)
v, m, st = train_bpe_on_bytes(None, vocab, ["<|endoftext|>"], ctx=ctx, return_stats=True, device_ptr=t.data_ptr(), n_bytes=n)
print(shape, n, vocab, len(m), {k: round(x, 1) for k, x in st.items() if k.startswith('ms_')})
