import sys, ctypes as C, time
sys.path.insert(0, '/root/repo')
import _bootstrap, torch, numpy as np
from transformer_lm_b200 import _lib
from transformer_lm_b200.synth import synth_device
from transformer_lm_b200.train import train_bpe_on_bytes
n = int(float(sys.argv[1])) // 4096 * 4096
ctx = _lib.default_context(0)
t = torch.empty(n, dtype=torch.uint8, device='cuda')
synth_device('owt', 4321, n, t.data_ptr(), ctx=ctx)
for it in range(2):
    v, m, st = train_bpe_on_bytes(None, 32000, ["<|endoftext|>"], ctx=ctx, return_stats=True, device_ptr=t.data_ptr(), n_bytes=n)
out = (C.c_ulonglong * 32)()
_lib.lib().bpe_debug_merge_profile(out)
p = list(out)
steps = max(p[7], 1)
print("stages", {k: round(x, 1) for k, x in st.items() if k.startswith('ms_')})
print("per-step us: phase1 %.2f sync1 %.2f apply %.2f sync2 %.2f | tokenCTA %.2f | records/step %.0f words/step %.1f dirty_sb(n/a) %.1f dirty_blk/step %.1f steps %d" % (
    p[0]/steps/1e3, p[1]/steps/1e3, p[2]/steps/1e3, p[3]/steps/1e3, p[4]/steps/1e3, p[5]/steps, p[6]/steps, p[8]/steps, p[9]/steps, steps))
print("slow compares per step: %.1f" % (p[8] / steps))
print("tail detail us/step: p1a %.2f p1b %.2f p1c %.2f p2 %.2f | list blocks %.1f superblocks %.1f | index range %.0f | sort: %d sorts, %.2f us each, ctr10-like" % (
    p[16]/steps/1e3, p[17]/steps/1e3, p[18]/steps/1e3, p[19]/steps/1e3, p[20]/steps, p[21]/steps, p[22]/steps, p[24], p[23]/max(p[24],1)/1e3))
print("token CTA us/step: winner %.2f meta %.2f bytes+key %.2f | CTA0 winner phase us: candidate loads %.2f compare+warp_best %.2f winner_range %.2f" % tuple([p[i] / steps / 1e3 for i in (10, 11, 12, 13, 14, 15)]))
print("local compares of 5 candidates: %.2f us" % (p[25] / steps / 1e3))
import os
if os.environ.get("BPE_STEP_PROFILE"):
    a = np.fromfile(os.environ["BPE_STEP_PROFILE"], dtype=np.uint32).reshape(-1, 4).astype(np.int64)
    ph1, ph2 = a[:, 0] / 1e3, a[:, 1] / 1e3
    rec = np.diff(np.concatenate([[0], a[:, 2]])) % (1 << 32)
    wrd = np.diff(np.concatenate([[0], a[:, 3]])) % (1 << 32)
    tot = ph1 + ph2
    print("step time us: sum %.0f ms; percentiles 10/50/90/99: %s" % (tot.sum() / 1e3, np.percentile(tot, [10, 50, 90, 99]).round(1)))
    for lo, hi in [(0, 100), (100, 1000), (1000, 5000), (5000, 15000), (15000, len(tot))]:
        sl = slice(lo, hi)
        print("steps %5d-%5d: ph1 %.1f us ph2 %.1f us  records %.0f words %.0f  (sum %.0f ms)" % (lo, hi, ph1[sl].mean(), ph2[sl].mean(), rec[sl].mean(), wrd[sl].mean(), tot[sl].sum() / 1e3))
    light = wrd <= 8
    print("light steps (<=8 words): %d, mean ph1 %.1f ph2 %.1f ; records there %.0f" % (light.sum(), ph1[light].mean(), ph2[light].mean(), rec[light].mean()))
