"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list: total time and launches per kernel."""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
ui = h.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = r[ki].split("(")[0]
    tot[name] += v; cnt[name] += 1
all_ms = sum(tot.values())
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("| `%s` | %d | %.3f | %.1f %% |" % (k, cnt[k], v, 100 * v / all_ms))
print("| all | %d | %.3f | |" % (sum(cnt.values()), all_ms))
