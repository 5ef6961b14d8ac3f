"""Offline: how many merges per step would the exact multi-merge rule accept?  Input: BPE_DUMP_MERGES file."""
import sys, numpy as np
a = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, 3)
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = len(a)
s = 0; steps = 0; sizes = []
while s < n:
    best = 1
    toks = set()
    r = 0
    while r < K and s + r < n:
        x, y, c = a[s + r]
        if x == y and r > 0: break
        if x in toks or y in toks: break
        if x >= 256 + s or y >= 256 + s: break     # uses a token made inside this batch
        toks.add(x); toks.add(y)
        r += 1
        if x == y: break
        # strictness: the next merge's count must be smaller than this one's
        if s + r >= n or a[s + r][2] < c: best = r
    sizes.append(best); s += best; steps += 1
sizes = np.array(sizes)
print("merges", n, "steps", steps, "avg batch %.2f" % (n / steps), "hist", np.bincount(sizes, minlength=K + 1)[1:])
# by phase
pos = np.cumsum(sizes) - sizes
for lo, hi in [(0, 1000), (1000, 5000), (5000, 15000), (15000, 32000)]:
    m = (pos >= lo) & (pos < hi)
    if m.any(): print("  merges %5d-%5d: avg batch %.2f" % (lo, hi, sizes[m].sum() / m.sum()))
