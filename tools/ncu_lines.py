"""Instructions and stall samples per CUDA source line of one kernel of an .ncu-rep (needs -lineinfo and --import-source on).
python tools/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def lines(rep, regex):
    txt = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + regex, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    agg, src, fname = defaultdict(lambda: [0, 0]), {}, ""
    hdr = None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            hdr = None
            continue
        if len(r) > 4 and r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            i_ins, i_smp = r.index("Instructions Executed"), r.index("# Samples")
            continue
        if hdr is None or len(r) <= i_ins or not r[0].isdigit():
            continue
        key = (fname, int(r[0]))
        src[key] = r[1].strip()
        agg[key][0] += int(r[i_ins] or 0)
        agg[key][1] += int(r[i_smp] or 0)
    return agg, src


if __name__ == "__main__":
    agg, src = lines(sys.argv[1], sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    print(f"total warp instructions {ti}, samples {ts}")
    for key, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * n / ti:5.1f}% ins {100 * s / ts:5.1f}% smp  {key[0]}:{key[1]:<4d} {src[key][:100]}")
