"""Splits the SASS listing of an ncu --page source CSV at barriers and prints, per segment, executed warp
instructions, stall samples and shared-memory wavefront excess.  Usage: ncu_segments.py file.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
seg = []
cur = {"inst": 0, "samples": 0, "n": 0, "wf": 0, "wf_ideal": 0, "first": None, "ops": {}}
tot_inst = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ci["Source"]].strip()
    inst = int(r[ci["Instructions Executed"]] or 0)
    smp = int(r[ci["# Samples"]] or 0)
    wf = int(r[ci["L1 Wavefronts Shared"]] or 0)
    wfi = int(r[ci["L1 Wavefronts Shared Ideal"]] or 0)
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    cur["inst"] += inst; cur["samples"] += smp; cur["n"] += 1; cur["wf"] += wf; cur["wf_ideal"] += wfi
    cur["ops"][op] = cur["ops"].get(op, 0) + inst
    if cur["first"] is None: cur["first"] = src
    tot_inst += inst
    if src.startswith("BAR") or "BAR.SYNC" in src:
        seg.append(cur)
        cur = {"inst": 0, "samples": 0, "n": 0, "wf": 0, "wf_ideal": 0, "first": None, "ops": {}}
seg.append(cur)
tot_s = sum(s["samples"] for s in seg)
for i, s in enumerate(seg):
    top = sorted(s["ops"].items(), key=lambda kv: -kv[1])[:6]
    print(f"seg {i:2d}: sass={s['n']:4d} inst={s['inst']:11d} ({100*s['inst']/max(tot_inst,1):5.1f}%) samples={s['samples']:6d} ({100*s['samples']/max(tot_s,1):5.1f}%) smem_wf={s['wf']} ideal={s['wf_ideal']}  top={top}")
print("total inst", tot_inst, "samples", tot_s)
