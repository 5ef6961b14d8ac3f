"""Summarise an .ncu-rep (read here, on the CPU box): one row per captured launch with the metrics the roofline uses."""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp_insts"), ("sm__inst_executed.avg.per_cycle_elapsed", "ipc")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    idx = [(h.index(m), label) for m, label in WANT if m in h]
    ki = h.index("Kernel Name")
    print("| kernel | " + " | ".join("%s (%s)" % (label, units[i]) if units[i] else label for i, label in idx) + " |")
    print("|---|" + "---|" * len(idx))
    for r in rows[2:]:
        print("| `%s` | " % r[ki].split("(")[0][:44] + " | ".join(r[i] for i, _ in idx) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
