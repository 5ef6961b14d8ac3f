"""profiles/r2_traffic.json from the round's .ncu-rep captures (read on the CPU box): DRAM bytes (read + written) per captured launch
of every kernel, x the launches a bench step makes (from the launch list of the same command).
usage: python tools/ncu_traffic.py [--append] <launches.csv> <out.json> <rep>:<config json> [<rep>:<config json> ...]
--append keeps the entries <out.json> already holds; a config key "_only" (regular expression on the full kernel name, template
arguments included) restricts a capture file to the kernels it was taken for; "_source" names the script that took it."""
import csv
import json
import re
import subprocess
import sys


def launches_per_name(launch_csv):
    """kernel base name -> launch count in the (ncu --metrics gpu__time_duration.sum --csv) launch list"""
    n = {}
    rows = [r for r in csv.reader(open(launch_csv, errors="replace")) if len(r) > 6]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki = hdr.index("Kernel Name")
    for r in rows:
        if r is hdr or len(r) <= ki:
            continue
        name = re.sub(r"^void ", "", r[ki]).split("(")[0]       # (template arguments kept: k_pretok_flags<0> trains, <1> encodes)
        n[name] = n.get(name, 0) + 1
    return n


def unit_scale(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


def main():
    argv = [a for a in sys.argv[1:] if a != "--append"]
    append = len(argv) != len(sys.argv) - 1
    launch_csv, out_path, specs = argv[0], argv[1], argv[2:]
    per_step_div = 1
    counts = launches_per_name(launch_csv)
    caps = []
    for spec in specs:
        rep, cfg = spec.split(":", 1)
        cfg = json.loads(cfg)
        steps_in_list = cfg.pop("_steps_in_launch_list", 1)      # the launch list covers warm-up + timed steps of bench.py
        only = re.compile(cfg.pop("_only", ""))
        script = cfg.pop("_source", "tools/profile_r2.sh")
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        h, units = rows[0], rows[1]
        ki, ri, wi, ti = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
        seen = {}
        for r in rows[2:]:
            name = re.sub(r"^void ", "", r[ki]).split("(")[0]
            if not only.search(name):
                continue
            b = float(r[ri]) * unit_scale(units[ri]) + float(r[wi]) * unit_scale(units[wi])
            seen.setdefault(name, []).append((b, float(r[ti]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(units[ti], 1.0)))
        for name, lst in seen.items():
            b = sum(x[0] for x in lst) / len(lst)
            ms = sum(x[1] for x in lst) / len(lst)
            per_step = max(1, round(counts.get(name, steps_in_list) / steps_in_list))
            caps.append({"kernel": name.split("<")[0], "config": cfg, "dram_bytes_per_launch": round(b, -3), "launches_per_step": per_step,
                         "dram_bytes_per_step": round(b * per_step, -3), "duration_ms": round(ms, 6), "captured_launches": len(lst),
                         "source": "ncu --set full --clock-control none, %s (round 2, %s)" % (rep.split("/")[-1], script)})
    if append:
        caps = json.load(open(out_path))["captures"] + caps
    json.dump({"captures": caps}, open(out_path, "w"), indent=1)
    print("wrote", out_path, len(caps), "entries")


if __name__ == "__main__":
    main()
