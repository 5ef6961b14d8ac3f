"""Per-source-line summary of one kernel of an .ncu-rep captured with --import-source on (read on the CPU box).
usage: python tools/ncu_source_lines.py <file.ncu-rep> <kernel regex> [top N]"""
import csv, subprocess, sys

def main(path, kernel, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = []
    ncol = None
    for line in out.splitlines():                # (source text may hold unescaped quotes: split by hand, numbers counted from the right)
        tk = line.strip().strip('"').split('","')
        if tk and tk[0] == "Line No": ncol = len(tk)
        if ncol and len(tk) > ncol and tk[0].isdigit(): tk = [tk[0], '","'.join(tk[1:len(tk) - ncol + 2])] + tk[len(tk) - ncol + 2:]
        rows.append(tk)
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
        if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
        if hdr and len(r) >= len(hdr) - 2 and r[0] not in ("", "Line No") and r[0].isdigit():
            lines.append((cur_file, r))
    ix = {}
    for i, h in enumerate(hdr): ix.setdefault(h, i)
    I = lambda r, h: int(float(r[ix[h]] or 0))
    tot_s = sum(I(r, "# Samples") for _, r in lines) or 1
    tot_i = sum(I(r, "Instructions Executed") for _, r in lines) or 1
    print("total samples %d, warp instructions %d" % (tot_s, tot_i))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for f, r in sorted(lines, key=lambda x: -I(x[1], "# Samples"))[:top]:
        st = sorted(((I(r, h), h[6:]) for h in stalls), reverse=True)[:2]
        print("%5.1f%% smp %5.1f%% inst lanes %4.1f  %s:%s  %s   [%s]" % (
            100.0 * I(r, "# Samples") / tot_s, 100.0 * I(r, "Instructions Executed") / tot_i,
            I(r, "Thread Instructions Executed") / max(1, I(r, "Instructions Executed")), f, r[0], r[1].strip()[:100],
            ", ".join("%s %d" % (n, v) for v, n in st if v)))

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
