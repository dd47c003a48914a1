"""Rebuilds the summaries under profiles/ from the raw captures in gpurun_out/ (see the gpurun command in the docstring
of each section).  Inputs: gpurun_out/bench_final.json (python bench.py), gpurun_out/launches_final.csv (ncu launch
list), gpurun_out/front_final.ncu-rep (ncu --set full of front_kernel), gpurun_out/variants.jsonl (bench variants).
Usage: python tools/make_profiles.py"""
import collections
import csv
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def launches():
    shutil.copy(os.path.join(G, "launches_final.csv"), os.path.join(P, "r1_launches_final.csv"))
    shutil.copy(os.path.join(G, "bench_final.json"), os.path.join(P, "r1_bench_final.json"))
    for n in (2, 4, 8):  # torchrun runs of the same bench, when they were made
        if os.path.exists(os.path.join(G, f"bench{n}_final.json")):
            shutil.copy(os.path.join(G, f"bench{n}_final.json"), os.path.join(P, f"r1_bench_final_{n}gpu.json"))
    rows = list(csv.reader(open(os.path.join(P, "r1_launches_final.csv"))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    ci = {n: i for i, n in enumerate(h)}
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows[hi + 1:]:
        if len(r) < len(h):
            continue
        k = r[ci["Kernel Name"]].split("(")[0].replace("void ", "")
        agg[k][r[ci["Metric Name"]]].append(float(r[ci["Metric Value"]].replace(",", "")))
    mean = lambda v: sum(v) / len(v)
    tot = sum(mean(v["gpu__time_duration.sum"]) for v in agg.values())
    d = last_json(os.path.join(P, "r1_bench_final.json"))
    md = ["# Round 1, final state: ncu launch list + bench (4K BGR, batch 64, 6 markers/frame)", "",
          "Command: `ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,sm__inst_executed.sum --clock-control none -s 117 -c 100 --csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu`",
          "(per-launch times are cold-cache and serialised: compare shares, not absolutes; raw csv: r1_launches_final.csv; 16 kernels per batch)", "",
          "| kernel | launches | avg us | share of summed time | SMs busy (active / elapsed cycles) | busy us | warp instructions (M) |", "|---|---|---|---|---|---|---|"]
    tb = 0
    for k, v in sorted(agg.items(), key=lambda kv: -mean(kv[1]["gpu__time_duration.sum"])):
        t = mean(v["gpu__time_duration.sum"]) / 1000
        act = sum(v["sm__cycles_active.avg"]) / sum(v["sm__cycles_elapsed.max"])
        inst = mean(v["sm__inst_executed.sum"]) / 1e6
        tb += t * act
        md.append(f"| {k} | {len(v['gpu__time_duration.sum'])} | {t:.1f} | {100 * t * 1000 / tot:.1f}% | {act:.2f} | {t * act:.1f} | {inst:.1f} |")
    st = d["stages_ms_per_step_unoverlapped"]
    md += ["", f"Sum of the kernel durations {tot / 1000:.0f} us; sum of (duration x fraction of SMs busy) = {tb:.0f} us.", "",
           f"Same workload timed with CUDA events by the library, one batch at a time (`bench.py`, not under ncu), ms per 64-frame step: front {st['front']:.3f}, ccl {st['ccl']:.3f}, quad {st['quad']:.3f}, feature {st['feature']:.3f}, decode {st['decode']:.3f}; sum {sum(st.values()):.2f} ms.",
           f"Pipelined ({d['pipelining']}): {d['ms_per_step']:.3f} ms/step = {d['value']:.0f} frames/s -- within a few % of the busy sum above: with several batches in flight the step time is set by how long the kernels keep SMs occupied, not by the critical path, so only instruction / efficiency cuts inside a kernel move it (tail-heavy kernels such as quad_fit overlap with the next batch's dense kernels). `bench.py --timeline` (ctag_stage_timeline_ms) shows the overlap per batch.",
           f"Front kernel: {d['roofline']['achieved']:.0f} GB/s algorithmic = {100 * d['roofline']['frac']:.1f} % of the measured 6552 GB/s.",
           f"e2e (pinned host frames through ctag_detect_batch): {d['e2e']['value']:.0f} frames/s = {d['e2e'].get('ms_per_step', float('nan')):.2f} ms per step (PCIe bound: 1593 MB H2D per step; the same bytes through a plain pinned copy with nothing else running take {d['e2e'].get('h2d_copy_alone_ms_per_step', float('nan')):.2f} ms = {d['e2e'].get('h2d_copy_alone_gbs', float('nan')):.1f} GB/s). CPU baseline: {d['cpu_baseline']['value']:.0f} frames/s on {d['cpu_baseline']['cores']} threads ({d['cpu_baseline']['sample']}).",
           f"Warm single-frame latency (test.bmp through ctag_detect, host frame in, markers out): {d.get('single_frame', {}).get('median_ms', float('nan')):.2f} ms.", "",
           "Round-1 trajectory of the pipelined step (same workload): 6.9 ms (first correct path) -> 1.78 ms -> 1.43 ms -> 1.39 ms: front kernel 3 CTAs/SM + instruction diet + merged vertical/extrema phase (0.61 -> 0.51 ms), sliding tile columns + L2 prefetch + one-row-pair hand-over (0.51 -> 0.466 ms), CCL runs + link deduplication (0.44 -> 0.23 ms), edges kernel occupancy / register-resident silhouettes / trace early-out and the fit kernel's lane refill + size-sorted work list (quad stage 1.38 -> 1.12 ms, fit instructions 136 M -> 79 M), refine ladder (0.37 -> 0.34 ms), decode kernel without divisions / local-memory record (0.25 -> 0.19 ms)."]
    open(os.path.join(P, "r1_launches_final.md"), "w").write("\n".join(md) + "\n")


def front():
    rep = os.path.join(G, "front_final.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(G, "front_final_src.csv"), "w").write(src)
    segs = subprocess.run(["python", os.path.join(ROOT, "tools", "ncu_segments.py"), os.path.join(G, "front_final_src.csv")],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    g = lambda n: (v[h.index(n)], u[h.index(n)])
    tob = lambda val, unit: float(val.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_dynamic",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
    rd, wr = tob(*g("dram__bytes_read.sum")), tob(*g("dram__bytes_write.sum"))
    alg = 4.25 * 3840 * 2160 * 64
    md = ["# Round 1: `ncu --set full` capture of front_bgr_slide_kernel (final state), 64 4K BGR frames per launch", "",
          "Command: `ncu --set full --clock-control none --import-source on -k regex:front_bgr -s 3 -c 1 -o prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --batch 64`",
          "(cold-cache single launch under the profiler; the bench number comes from CUDA events, not from here)", "", "| metric | value | unit |", "|---|---|---|"]
    md += [f"| {w} | {g(w)[0]} | {g(w)[1]} |" for w in want if w in h]
    md += ["", f"DRAM traffic per launch: {rd / 1e6:.1f} MB read + {wr / 1e6:.1f} MB written = {(rd + wr) / 1e6:.1f} MB; algorithmic bytes (3N read + N gray + N/4 binary, N = 3840x2160, 64 frames) = {alg / 1e6:.1f} MB -> traffic / algorithmic = {(rd + wr) / alg:.3f}.",
           f"The staged regions overlap horizontally only (192 full-res columns loaded per 160 owned: 1.2x; vertically a CTA walks down its tile column and loads every row once); L2 absorbs most of that, DRAM reads are {rd / (3 * 3840 * 2160 * 64):.2f}x the input bytes. Writes are slightly below the algorithmic figure because part of the last tiles' output is still in L2 when the kernel ends. No re-read problem to fix.", "",
           "SASS evidence (cuobjdump -sass libctag_b200.so): `UTMALDG.3D` (TMA tile loads), `UTMAPF.L2.3D` (TMA prefetch of the next tile's rows into L2), `SYNCS.PHASECHK.TRANS64.TRYWAIT` (mbarrier), `IDP.2A.*.U16.U8` / `IDP.4A.U8.S8` (dp2a/dp4a stencil arithmetic), `VIMNMX3.U16x2` (column extrema).", "",
           "Per-phase breakdown (SASS split at the CTA barriers, `tools/ncu_segments.py`; executed warp instructions, stall samples, shared-memory wavefronts vs ideal):", "", "```"]
    md += segs.strip().splitlines()
    md += ["```", "", "seg 1 = run / next-tile bookkeeping, seg 2 = reuse from the tile above (10 half-res rows, 2 rows of column extrema) + wait for the staged rows, seg 3 = BGR->gray of the 80 new rows + gray store (incl. the 11 rows owned by the tile below when the run goes on), seg 4-5 = replicate-border patch (edge tiles only), seg 6 = horizontal taps on the 40 new row pairs, seg 7 = vertical taps + rounding + column extrema of the 40 new half-res rows (first row pair from the keep buffer), seg 8 = 5x5 tile extrema, seg 9 = 3x3 dilation + threshold, seg 10 = compare + store.", "",
           "Reading: three CTAs per SM (shared memory 73.4 KB each, 56 registers per thread: both limits sit at 3). Against the non-sliding kernel (320.5 M warp instructions, 0.523 ms) the gray conversion fell from 123 M to 80 M instructions (no vertical halo), the horizontal pass from 63 M to 49 M (the tile below needs exactly one row pair of the tile above, handed over through a 2 x 384 B keep buffer) and the vertical pass from 75 M to 71 M; the kernel issues about 60 % of its slots, barrier and shared-memory (short scoreboard) stalls lead. The wait for the staged rows (seg 2) is the one place where load latency shows: a CTA has one staging area, so its next load can only start once the vertical pass has released it (the rows are prefetched into L2 meanwhile); the other two CTAs of the SM cover most of it. Next levers: the horizontal pass's 24-lane row mapping costs 1.65x the ideal shared-memory wavefronts, and phases D-F keep only 128-200 of the 384 threads busy between their barriers."]
    open(os.path.join(P, "r1_front_kernel_ncu.md"), "w").write("\n".join(md) + "\n")
    json.dump({"4k": {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_total": rd + wr, "frames_per_launch": 64,
                      "source": "profiles/r1_front_kernel_ncu.md"}}, open(os.path.join(P, "front_traffic.json"), "w"), indent=1)


def variants():
    d = last_json(os.path.join(P, "r1_bench_final.json"))
    vs = [json.loads(l) for l in open(os.path.join(G, "variants.jsonl")) if l.strip()][-3:]
    row = lambda name, v: f"| {name} | {v['value']:,.0f} | {v['ms_per_step']:.3f} | {v['roofline']['kernel_ms_per_launch']:.3f} | {v['roofline']['frac']:.3f} | {v['e2e']['value']:,.0f} |"
    md = ["# Round 1: bench variants on one B200 (final state), `python bench.py --steps 60 --warmup 3 --no-cpu <variant>`", "",
          "All numbers: CUDA events, pipelined loop (several batches in flight, one stream each), parity spot check against the oracle = ok in every run.", "",
          "| variant | frames/s (device resident) | ms / step | front kernel ms / launch | front roofline frac | e2e frames/s (pinned host, H2D inside) |", "|---|---|---|---|---|---|",
          row("4K BGR, batch 64 (default, BASELINE config 4/5; 100 steps)", d), row("4K gray, batch 64 (detect's own contract; 1.25 B/px; sliding gray kernel, 4 CTAs/SM)", vs[0]),
          row("1080p BGR, batch 128, six markers per frame", vs[1]), row("1080p gray, batch 128, six markers per frame", vs[2])]
    c3 = os.path.join(G, "bench_1080p_b256_m1.json")
    if os.path.exists(c3):  # BASELINE config 3 to the letter: 256 frames, ONE rendered marker each
        shutil.copy(c3, os.path.join(P, "r1_bench_config3_1080p_b256.json"))
        md.append(row("1080p BGR, batch 256, one marker per frame (BASELINE config 3 as written)", last_json(c3)))
    md += ["", "Unoverlapped stage times (ms per step, batches one at a time):"]
    for name, v in (("4K BGR", d), ("4K gray", vs[0]), ("1080p BGR (128 frames)", vs[1]), ("1080p gray (128 frames)", vs[2])):
        md.append(f"* {name}: " + ", ".join(f"{k} {x:.3f}" for k, x in v["stages_ms_per_step_unoverlapped"].items()))
    md += ["", "Reading: with gray input the fused front end is bound by its stencil arithmetic and shared-memory traffic, not by HBM",
           "(same phases B-F as the BGR kernel on a quarter of the bytes), so its HBM fraction is low although it is faster in",
           "absolute terms; at 1080p the per-frame sparse stages (same 6 markers per frame) dominate. e2e is PCIe bound in every",
           "variant (24.9 MB per 4K BGR frame).", "",
           "Multi-GPU (torchrun, one rank per GPU, weak scaling, frames sharded, no data-path collective; max-over-ranks timing):", "",
           "| GPUs | frames/s (device resident) | ms / step | e2e frames/s | state |", "|---|---|---|---|---|",
           f"| 1 | {d['value']:,.0f} | {d['ms_per_step']:.3f} | {d['e2e']['value']:,.0f} | final |",
           *[f"| {v['n_gpus']} | {v['value']:,.0f} | {v['ms_per_step']:.3f} | {v['e2e']['value']:,.0f} | final ({v['value'] / d['value']:.2f}x / e2e {v['e2e']['value'] / d['e2e']['value']:.2f}x of one GPU) |"
             for v in (last_json(os.path.join(G, n)) for n in ("bench2_final.json", "bench4_final.json") if os.path.exists(os.path.join(G, n)))], "",
           "Result equality across GPUs (`tools/multigpu_check.py` under torchrun, 2 ranks, NCCL gather of the records): 16 frames of the config-2 sequence sharded over 2 B200s, 78 markers -- the gathered list equals the single-GPU list byte for byte.", "",
           "Warm single-frame latency of `ctag_detect` (host gray frame in, markers out, wall clock, `tools/latency.py`): test.bmp 1920x1200 1.35 ms,",
           "synthetic 4K 1.19 ms (quad 0.29-0.66 ms and decode 0.20-0.24 ms dominate: single-lane / single-warp serial parts)."]
    open(os.path.join(P, "r1_bench_variants.md"), "w").write("\n".join(md) + "\n")


if __name__ == "__main__":
    launches()
    front()
    variants()
    print("profiles/ rebuilt")
