"""Per-kernel summary of an ncu launch list (csv made with --metrics gpu__time_duration.sum,sm__cycles_active.avg,
sm__cycles_elapsed.max,sm__inst_executed.sum).  python tools/launch_summary.py gpurun_out/x.csv [--md]"""
import collections
import csv
import sys


def summarize(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    d = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows[hdr + 1:]:
        if len(r) < 15:
            continue
        name = r[4].split("(")[0].replace("void ", "")
        d[name][r[12]].append(float(r[14].replace(",", "")))
    out = []
    for k, v in d.items():
        t = sum(v["gpu__time_duration.sum"]) / len(v["gpu__time_duration.sum"]) / 1000
        b = sum(v["sm__cycles_active.avg"]) / sum(v["sm__cycles_elapsed.max"])
        inst = sum(v["sm__inst_executed.sum"]) / len(v["sm__inst_executed.sum"]) / 1e6
        out.append((t, k, len(v["gpu__time_duration.sum"]), b, t * b, inst))
    out.sort(reverse=True)
    return out


if __name__ == "__main__":
    out = summarize(sys.argv[1])
    md = "--md" in sys.argv
    tot = sum(o[0] for o in out)
    if md:
        print("| kernel | launches | avg us | share of summed time | SMs busy (active / elapsed cycles) | busy us | warp instructions (M) |")
        print("|---|---|---|---|---|---|---|")
    for t, k, n, b, tb, inst in out:
        print(f"| {k} | {n} | {t:.1f} | {100 * t / tot:.1f}% | {b:.2f} | {tb:.1f} | {inst:.1f} |" if md else
              f"{k:34s} {n:3d} {t:8.1f} {b:5.2f} {tb:7.1f} {inst:7.1f}")
    print(f"\nSum of the kernel durations {tot:.0f} us; sum of (duration x fraction of SMs busy) = {sum(o[4] for o in out):.0f} us.")
