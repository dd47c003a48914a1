"""Per-CUDA-source-line hot spots from an .ncu-rep (needs -lineinfo + --import-source on).
Usage: ncu_lines.py report.ncu-rep [top_n] [kernel-name regex]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kf = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
txt = subprocess.run(["ncu", "-i", rep] + kf + ["--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = None
hdr = None
out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        # duplicate 'Source' header: first is cuda source, second sass
        idx_samples = hdr.index("# Samples"); idx_inst = hdr.index("Instructions Executed"); idx_thr = hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] == "":  # sass row
        continue
    try:
        out.append((int(r[idx_samples] or 0), int(r[idx_inst] or 0), int(r[idx_thr] or 0), cur_file, r[0], r[1].strip()))
    except ValueError:
        pass
ts = sum(o[0] for o in out) or 1; ti = sum(o[1] for o in out) or 1
print(f"total samples {ts} inst {ti}")
for o in sorted(out, key=lambda o: -o[0])[:topn]:
    print(f"{100*o[0]/ts:5.1f}% smp {100*o[1]/ti:5.1f}% inst thr/inst {o[2]/max(o[1],1):4.1f} | {o[3].split('/')[-1]}:{o[4]} | {o[5][:100]}")
