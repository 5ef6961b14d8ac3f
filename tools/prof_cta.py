"""Per-CTA timeline of the merge loop (profile build: BPE_EXTRA_NVCC_FLAGS=-DBPE_MERGE_PROFILE, BPE_CTA_PROFILE=<file>).
For every step: when each CTA started phase 1, arrived at / left barrier 1, arrived at barrier 2."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap, torch, numpy as np
from transformer_lm_b200 import _lib
from transformer_lm_b200.synth import synth_device
from transformer_lm_b200.train import train_bpe_on_bytes
n = int(float(sys.argv[1])) // 4096 * 4096
G = int(os.environ.get("BPE_MERGE_G", "148"))
ctx = _lib.default_context(0)
t = torch.empty(n, dtype=torch.uint8, device='cuda')
synth_device('owt', 4321, n, t.data_ptr(), ctx=ctx)
v, m, st = train_bpe_on_bytes(None, 32000, ["<|endoftext|>"], ctx=ctx, return_stats=True, device_ptr=t.data_ptr(), n_bytes=n)
print("stages", {k: round(x, 1) for k, x in st.items() if k.startswith('ms_')})
a = np.fromfile(os.environ["BPE_CTA_PROFILE"], dtype=np.uint64)
nm = len(m)
a = a[: nm * G * 4].reshape(nm, G, 4).astype(np.int64)
ok = (a > 0).all(axis=(1, 2))
idx = np.nonzero(ok)[0]
print("grid steps with full data: %d (merges %d)" % (ok.sum(), nm))
a = a[idx]
r = np.diff(np.concatenate([idx, [nm]]))           # merges per step
start, arr1, exit1, arr2 = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
nxt = np.roll(start, -1, axis=0)
def rep(name, x, m):
    x = x[m]
    print("  %-42s mean %6.2f us  p50 %6.2f  p90 %6.2f" % (name, x.mean() / 1e3, np.percentile(x, 50) / 1e3, np.percentile(x, 90) / 1e3))
valid = np.ones(len(idx), bool); valid[-1] = False
for name, m in [("merges 100-1000", (idx >= 100) & (idx < 1000)), ("merges 1000-5000", (idx >= 1000) & (idx < 5000)), ("merges 5000-15000", (idx >= 5000) & (idx < 15000)),
                ("merges 15000-", idx >= 15000), ("batch 1 after 5000", (idx >= 5000) & (r == 1)), ("batch 8 after 5000", (idx >= 5000) & (r == 8))]:
    m = m & valid
    if not m.any(): continue
    print("%s: %d steps, %.2f merges per step" % (name, m.sum(), r[m].mean()))
    rep("phase 1: last arrive1 - first start", arr1.max(1) - start.min(1), m)
    rep("phase 1 of the median CTA", np.median(arr1 - start, axis=1), m)
    rep("start skew: last start - first start", start.max(1) - start.min(1), m)
    rep("gather+select: first exit1 - last arrive1", exit1.min(1) - arr1.max(1), m)
    rep("gather+select: last exit1 - last arrive1", exit1.max(1) - arr1.max(1), m)
    rep("apply: last arrive2 - first exit1", arr2.max(1) - exit1.min(1), m)
    rep("apply of the median CTA", np.median(arr2 - exit1, axis=1), m)
    rep("apply of the token CTA", (arr2 - exit1)[:, G - 1], m)
    rep("barrier 2 + status: first next start - last arrive2", nxt.min(1) - arr2.max(1), m)
    rep("whole step", nxt.min(1) - start.min(1), m)
    last2 = arr2[m].argmax(1)
    print("  last CTA at barrier 2: token CTA in %.1f %% of steps, CTA 0 in %.1f %%" % (100 * (last2 == G - 1).mean(), 100 * (last2 == 0).mean()))
