import sys, time
sys.path.insert(0, '.')
import _bootstrap, torch, numpy as np
from transformer_lm_b200 import _lib
from transformer_lm_b200.synth import synth_device
from transformer_lm_b200.train import train_bpe_on_bytes, merges_to_python
from transformer_lm_b200.vocab import Vocab
import transformer_lm_b200.train as T
ctx = _lib.default_context(0)
n = int(1.1e10) // 4096 * 4096
t = torch.empty(n, dtype=torch.uint8, device='cuda')
synth_device('owt', 4321, n, t.data_ptr(), ctx=ctx)
tt = []
orig_finish = T.LiveMerges.finish
def timed(self, p, k):
    t0 = time.time(); r = orig_finish(self, p, k); tt.append(time.time() - t0); return r
T.LiveMerges.finish = timed
for it in range(4):
    torch.cuda.synchronize(); t0 = time.time()
    v, m, st = train_bpe_on_bytes(None, 32000, ["<|endoftext|>"], ctx=ctx, return_stats=True, device_ptr=t.data_ptr(), n_bytes=n)
    torch.cuda.synchronize(); dt = time.time() - t0
    print("wall %.1f ms, C stages total %.1f ms, LiveMerges.finish (what is left after the loop) %.1f ms" % (dt * 1e3, st["ms_total"], tt[-1] * 1e3))
