"""GPU, profile build (-DBPE_MERGE_PROFILE): per-phase time of the merge loop on CTA 0.  usage: prof_merge.py <bytes> [shape seed vocab]"""
import sys, ctypes as C, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap, torch, numpy as np
from transformer_lm_b200 import _lib
from transformer_lm_b200.synth import synth_device
from transformer_lm_b200.train import train_bpe_on_bytes
n = int(float(sys.argv[1])) // 4096 * 4096
shape = sys.argv[2] if len(sys.argv) > 2 else 'owt'
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 4321
vocab = int(sys.argv[4]) if len(sys.argv) > 4 else 32000
ctx = _lib.default_context(0)
t = torch.empty(n, dtype=torch.uint8, device='cuda')
synth_device(shape, seed, n, t.data_ptr(), ctx=ctx)
for it in range(2):
    v, m, st = train_bpe_on_bytes(None, vocab, ["<|endoftext|>"], ctx=ctx, return_stats=True, device_ptr=t.data_ptr(), n_bytes=n)
out = (C.c_ulonglong * 32)()
_lib.lib().bpe_debug_merge_profile(out)
p = list(out)
steps = max(p[7], 1)
print("stages", {k: round(x, 1) for k, x in st.items() if k.startswith('ms_')}, "merges", len(m), "steps", st["merge_steps"])
print("per-step us: phase1 %.2f gather+select %.2f apply %.2f sync2(+sort) %.2f | tokenCTA %.2f | records/step %.0f sites/step %.1f dirty_blk/step %.1f steps %d" % (
    p[0]/steps/1e3, p[1]/steps/1e3, p[2]/steps/1e3, p[3]/steps/1e3, p[4]/steps/1e3, p[5]/steps, p[6]/steps, p[9]/steps, steps))
print("slow (byte-wise) token compares per step: %.1f" % (p[8] / steps))
print("detail us/step: " + " ".join("p%d=%.2f" % (i, p[i]/steps/1e3) for i in range(10, 32) if p[i]))
if os.environ.get("BPE_STEP_PROFILE"):
    a = np.fromfile(os.environ["BPE_STEP_PROFILE"], dtype=np.uint32).reshape(-1, 4).astype(np.int64)
    a = a[a[:, 3] > 0]
    ph1, ph2, r = a[:, 0] / 1e3, a[:, 1] / 1e3, a[:, 3]
    tot = ph1 + ph2
    print("step time us: sum %.0f ms; percentiles 10/50/90/99: %s" % (tot.sum() / 1e3, np.percentile(tot, [10, 50, 90, 99]).round(1)))
    idx = np.nonzero(np.fromfile(os.environ["BPE_STEP_PROFILE"], dtype=np.uint32).reshape(-1, 4)[:, 3] > 0)[0]
    for lo, hi in [(0, 100), (100, 1000), (1000, 5000), (5000, 15000), (15000, 1 << 30)]:
        mk = (idx >= lo) & (idx < hi)
        if mk.any(): print("  merges %5d-%5d: %5d steps, %.1f ms total, %.1f us/step (ph1+gather %.1f, apply+sync %.1f), %.1f merges/step" % (
            lo, min(hi, idx.max()), mk.sum(), tot[mk].sum() / 1e3, tot[mk].mean(), ph1[mk].mean(), ph2[mk].mean(), r[mk].mean()))
    for k in range(1, 16):
        mk = r == k
        if mk.any(): print("  batch %d: %6d steps, ph1+gather %.1f us, apply+sync %.1f us" % (k, mk.sum(), ph1[mk].mean(), ph2[mk].mean()))
