#!/usr/bin/env python
"""Benchmark of the CylinderTag detect hot path (BASELINE.json metric: detect frames/s at 1080p & 4K on 1/2/4/8 B200,
HBM GB/s vs peak, CPU reference beside it).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, weak scaling, no data-path collective)

Workload (N = 1 and every N): BASELINE config 5 as SURVEY 8(d) writes it -- a device-resident ring of 64 DISTINCT
3840x2160 BGR frames (the config-4 2f12c frames, 4..8 rendered markers each, seeds 2000..2063; 1.6 GB >> 126 MB L2),
one step = the whole detect path over the 64 frames of the ring (BGR -> gray -> 2x cubic decimation -> adaptive
threshold -> CCL -> quad extraction -> pairing -> edge refinement -> decode).
  value        frames/s with the ring resident in HBM (CUDA events on the detector's stream, max over ranks)
  e2e          the same metric through ctag_detect_batch(is_device=0): pinned HOST frames in, markers out, H2D + D2H
               inside the timed region
  roofline     the dominant dense kernel (fused front end, 4.25 algorithmic bytes per full-res pixel) against the
               measured HBM peak; frac_of_dense_ceiling = whole path against the 5.5 B/px dense ceiling
  variants     (N = 1) config 3 as written (256 frames 1920x1080 BGR, one marker each) and the 4K ring as GRAY frames
               (detect()'s own contract), each with its own roofline
  parity_check every distinct frame of the ring against the reference's own code (oracle/_ref) after the timed region
  cpu_baseline / --impl reference: the reference's corner_detector.cpp + CylinderTag.cpp compiled unmodified
               (oracle/_ref), main.cpp's cvtColor + detect loop, one CylinderTag per host thread, all host threads
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "detect_frames_per_s"
DTYPE = "u8/int32 dense, f32/f64 sparse"


def workload_config():
    """Identical in both arms (the driver compares the dicts)."""
    return {"workload": "BASELINE config 5: ring of 64 distinct 3840x2160 BGR synthetic frames (config-4 2f12c frames, 4-8 rendered "
                        "markers each, seeds 2000-2063; 1593 MB vs 126 MB L2), batch 64 = one pass over the ring, "
                        "detect(adaptiveThresh=5, cornerSubPix=true, dist=5) incl. the caller's BGR2GRAY",
            "resolution": "3840x2160", "channels": 3, "batch": 64, "distinct_frames": 64, "markers_per_frame": "4-8",
            "codebook": "CTag_2f12c (41x12, feature size 2)", "l2_policy": "inputs larger than L2", "frames_per_step_per_gpu": 64}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while a timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev, self.rows, self.proc = dev, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        num = lambda s: s.replace(".", "", 1).isdigit()
        sm = [float(r[0]) for r in self.rows if r and num(r[0])]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and num(r[1])]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and num(r[2])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


def hbm_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def bind_to_gpu_numa_node(local):
    """Pins this process (and therefore the pages of the pinned buffers it allocates next: first touch) to the CPUs of
    the NUMA node the GPU hangs off.  Returns a description, or None when the topology cannot be read."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus  # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        guessed = False
        if node < 0:
            # the container hides the GPU's node: spread the ranks over the online nodes in order (HGX boxes hang GPUs 0-3
            # off socket 0 and 4-7 off socket 1), which at least keeps the pinned buffers of one rank on one node
            nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
            world = int(os.environ.get("WORLD_SIZE", "1"))
            if len(nodes) < 2 or world < 2:
                return {"gpu_bus": bus, "numa_node": node, "bound": False, "nodes_online": len(nodes)}
            node = nodes[min(local * len(nodes) // world, len(nodes) - 1)]
            guessed = True
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return {"gpu_bus": bus, "numa_node": node, "bound": False}
        os.sched_setaffinity(0, allowed)
        return {"gpu_bus": bus, "numa_node": node, "bound": True, "cpus": len(allowed), "node_guessed_from_rank": guessed}
    except Exception as exc:  # pragma: no cover
        return {"bound": False, "error": str(exc)}


def load_ring(kind, rank, world, barrier=None, limit=0):
    """The distinct frames of a workload, rendered once per box (cached under /tmp) with the ranks sharing the work."""
    from cylindertag_b200 import workloads as wl
    jobs = [(4, "2f12c", i) for i in range(wl.CONFIG5_DISTINCT)] if kind == "config5" else [(3, None, i) for i in range(wl.CONFIG3_FRAMES)]
    if limit:
        jobs = jobs[:limit]
    if world > 1:
        mine = [j for k, j in enumerate(jobs) if k % world == rank]
        wl.render_many(mine, workers=max(1, (os.cpu_count() or 8) // world))
        barrier()
    return wl.render_many(jobs, workers=min(16, os.cpu_count() or 8))


# ---- the reference arm / cpu_baseline -----------------------------------------------------------------------------------
def cpu_reference_pass(frames, state, fs, threads):
    """One pass of the reference's own code over `frames` on `threads` host threads.  Returns (frames/s, markers)."""
    from oracle import ref_api
    t = time.perf_counter()
    n_mk = ref_api.detect_batch_bgr(frames, state, fs, 5, True, 5, threads)
    return len(frames) / (time.perf_counter() - t), n_mk


def cpu_sample_text(n, threads):
    return (f"{n} distinct frames of the workload per pass (3840x2160 BGR), the reference's own corner_detector.cpp + CylinderTag.cpp "
            f"compiled unmodified (oracle/_ref: g++ -O3 -DNDEBUG, no -march, OpenCV entry points from oracle/ref_shim), "
            f"main.cpp's cvtColor + detect loop, one CylinderTag per host thread, {threads} thread(s)")


def run_reference_arm(a):
    from cylindertag_b200 import workloads as wl
    cores = os.cpu_count() or 1
    state, fs = wl.codebook("2f12c")
    frames = np.stack(load_ring("config5", 0, 1, limit=a.ring))
    for _ in range(min(a.warmup, 1)):
        cpu_reference_pass(frames[:cores], state, fs, cores)
    t0, fps = time.perf_counter(), []
    for _ in range(max(a.steps, 1)):
        v, _ = cpu_reference_pass(frames, state, fs, cores)
        fps.append(v)
        if time.perf_counter() - t0 > 150:  # keep the whole arm within a few minutes whatever K is
            break
    value = float(np.mean(fps))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": a.gpus, "steps": len(fps),
            "warmup": a.warmup, "ms_per_step": 1000.0 * len(frames) / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": workload_config(),
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": cpu_sample_text(len(frames), cores)},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


# ---- our arm ------------------------------------------------------------------------------------------------------------
class DeviceRun:
    """A batch resident in HBM plus the pipelined / sequential loops over it."""

    def __init__(self, torch, det, dev, frames_np, channels, cap=16):
        self.torch, self.det, self.dev, self.cap = torch, det, dev, cap
        self.n, self.h, self.w = frames_np.shape[:3]
        self.ch = channels
        self.host = torch.empty(frames_np.shape, dtype=torch.uint8).pin_memory()
        self.host.copy_(torch.from_numpy(frames_np))
        self.frames = self.host.to(dev, non_blocking=False)
        self.pitch, self.fstride = self.w * channels, self.w * self.h * channels

    def enqueue(self):
        self.det.enqueue_device(self.frames.data_ptr(), self.n, self.w, self.h, self.pitch, self.fstride, self.ch, 5, True, 5)

    def step(self):
        self.enqueue()
        return self.det.collect(self.cap)

    def warm(self, steps=3):
        """Warm-up: `steps` batches one at a time, then one batch on EVERY workspace of the detector (each slot sizes its
        buffers on first use; that must not happen inside a timed region)."""
        out = None
        for _ in range(steps):
            out = self.step()
        depth = self.det.max_in_flight()
        for _ in range(depth):
            self.enqueue()
        for _ in range(depth):
            self.det.collect(self.cap)
        return out

    def pipelined(self, steps, barrier, seconds=None, timeline=None):
        """`steps` batches (or as many as fit in `seconds`) with up to max_in_flight() batches enqueued ahead.  Returns
        (elapsed ms between CUDA events on the detector's stream, steps run, kernel launches, markers, last result)."""
        torch, det = self.torch, self.det
        depth = det.max_in_flight()
        stream = torch.cuda.ExternalStream(det.stream(), device=self.dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches = n_markers = done = queued = 0
        last = None
        barrier()
        ev0.record(stream)  # every detector stream is idle here: stamped immediately
        t0 = time.perf_counter()
        target = steps if seconds is None else 1 << 30
        while queued < min(depth - 1, target):
            self.enqueue()
            queued += 1
        while done < target:
            if queued < target and (seconds is None or time.perf_counter() - t0 < seconds):
                self.enqueue()
                queued += 1
            if done >= queued:
                break
            last = det.collect(self.cap)
            done += 1
            if timeline is not None:
                timeline.append(det.stage_timeline_ms())
            launches += det.launch_count()
            n_markers += int(last[1].sum())
        ev1.record(stream)  # every collect synchronised its stream: stamped after the last batch finished
        barrier()
        return ev0.elapsed_time(ev1), done, launches, n_markers, last

    def stage_times(self, reps):
        """Per-stage GPU time with batches run ONE AT A TIME (nothing else shares the GPU with the kernel being timed);
        CUDA events recorded by the library on the launching stream."""
        from cylindertag_b200 import _capi
        acc = {k: 0.0 for k in _capi.STAGE_NAMES}
        for _ in range(reps):
            self.step()
            for k, v in self.det.stage_times_ms().items():
                acc[k] += v
        return {k: v / reps for k, v in acc.items()}


def roofline_of(front_ms, w, h, n, ch, traffic=None, traffic_source=None):
    peak, peak_src = hbm_peak()
    alg = (4.25 if ch == 3 else 1.25) * w * h * n
    achieved = alg / (front_ms / 1000.0) / 1e9
    return {"bound": "hbm", "kernel": ("front_bgr_slide_kernel (BGR->gray + " if ch == 3 else "front_gray_slide_kernel (") +
            "2x cubic decimation + tile extrema + adaptive threshold, fused)", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
            "traffic_source": traffic_source, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
            "algorithmic_bytes_per_pixel": 4.25 if ch == 3 else 1.25, "kernel_ms_per_launch": front_ms}


def gray_traffic(n):
    """DRAM bytes of one launch of the gray front kernel: the committed ncu capture (profiles/front_traffic.json)."""
    try:
        td = json.load(open(os.path.join(ROOT, "profiles", "front_traffic.json"))).get("4k_gray")
        if td and td.get("frames_per_launch") == n:
            return td["dram_bytes_total"], f"constant from {td.get('source')}"
    except Exception:
        pass
    return None, None


def measure_traffic(kernel_regex):
    """dram__bytes of ONE launch of the front kernel on this workload, measured now: re-runs this script under ncu with
    --traffic-probe (two batches of the cached ring, the second one captured).  None when ncu is not usable."""
    import shutil
    ncu = shutil.which("ncu") or ("/usr/local/cuda/bin/ncu" if os.path.exists("/usr/local/cuda/bin/ncu") else None)
    if not ncu:
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", f"regex:{kernel_regex}",
           "-s", "1", "-c", "1", "--csv", sys.executable, os.path.abspath(__file__), "--traffic-probe"]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        tot = 0.0
        for row in res.stdout.splitlines():
            cells = [c.strip('"') for c in row.split('","')]
            if len(cells) > 3 and cells[-3].startswith("dram__bytes_"):
                v = float(cells[-1].replace(",", ""))
                unit = cells[-2].lower()
                tot += v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        return (tot, "measured in this run: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on one launch") if tot > 0 \
            else (None, "ncu gave no counters: " + (res.stderr or res.stdout)[-200:].replace("\n", " "))
    except Exception as exc:  # pragma: no cover
        return None, f"ncu failed: {exc}"


def traffic_probe():
    """Child of measure_traffic (runs under ncu): two batches of the cached ring through the device path."""
    import torch
    from cylindertag_b200 import Detector, workloads as wl
    state, fs = wl.codebook("2f12c")
    frames = np.stack(load_ring("config5", 0, 1))
    det = Detector(state=state, feature_size=fs, device=0)
    run = DeviceRun(torch, det, torch.device("cuda", 0), frames, 3)
    run.step()
    run.step()
    det.close()


def check_parity(markers, counts, frames_bgr, state, fs, cores):
    """Every frame of the ring against the reference's own code (oracle/_ref); outside every timed region."""
    try:
        from oracle import ref_api
        if not ref_api.available():
            raise RuntimeError("oracle/_ref not available")
        rc, rm = ref_api.detect_batch_mt(frames_bgr, state, fs, 5, True, 5, threads=cores, cap=markers.shape[1])
        bad, worst, total = [], 0.0, 0
        for f in range(len(frames_bgr)):
            n = int(rc[f][5])
            ok = n == int(counts[f])
            for k in range(n if ok else 0):
                g, r = markers[f][k], rm[f][k]
                nf = int(r["n_features"])
                ok = ok and int(g["marker_id"]) == int(r["marker_id"]) and int(g["n_features"]) == nf and int(g["inverse"]) == int(r["inverse"])
                for name in ("feature_pos", "feature_id", "id_left", "id_right"):
                    ok = ok and list(g[name][:nf]) == list(r[name][:nf])
                if ok:
                    err = float(np.abs(g["corners"][:nf] - r["corners"][:nf]).max())
                    worst = max(worst, err)
                    ok = err <= 1e-3
                total += 1
            if not ok:
                bad.append(f)
        return {"status": "ok" if not bad else "MISMATCH", "checker": "oracle/_ref (reference's corner_detector.cpp + CylinderTag.cpp compiled unmodified)",
                "frames_checked": len(frames_bgr), "markers_checked": total, "worst_corner_error_px": worst, "mismatching_frames": bad[:8]}
    except Exception as exc:  # pragma: no cover
        return {"status": f"error: {exc}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--timeline", default="", help="write the stage timeline of the timed batches (ms, per batch) to this file")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-jpeg", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 4096-distinct-frame run (GPU-rendered frames)")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-traffic", action="store_true")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of the sustained loop after the K timed steps")
    ap.add_argument("--traffic-probe", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--ring", type=int, default=0, help=argparse.SUPPRESS)  # tests: first N frames of the ring (reference arm)
    a = ap.parse_args()
    if a.traffic_probe:
        return traffic_probe()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        if rank == 0:
            run_reference_arm(a)
        return
    a.warmup = max(a.warmup, 3)

    numa = bind_to_gpu_numa_node(local)  # before torch allocates pinned memory
    import torch
    import torch.distributed as dist
    from cylindertag_b200 import Detector, _capi, workloads as wl
    _capi.load()  # fail loudly if the CUDA extension is missing: there is no fallback
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    state, fs = wl.codebook("2f12c")
    ring = np.stack(load_ring("config5", rank, world, barrier))
    if world > 1:  # weak scaling: every rank owns a ring of the same 64 frames, rotated so that the ranks are out of phase
        ring = np.roll(ring, -(rank * (len(ring) // world)), axis=0)
    n, h, w = ring.shape[:3]
    det = Detector(state=state, feature_size=fs, device=local)
    run = DeviceRun(torch, det, dev, ring, 3)

    run.warm(a.warmup)
    depth = det.max_in_flight()
    timeline = [] if a.timeline else None
    sampler = ClockSampler(local).start()
    ms, steps_done, launches, n_markers, last = run.pipelined(a.steps, barrier, timeline=timeline)
    clocks = sampler.stop()
    ms_max = max_over_ranks(ms)
    value = n * a.steps * world / (ms_max / 1000.0)
    markers, counts, info = last
    if a.timeline and rank == 0:
        with open(a.timeline, "w") as f:
            f.write("# batch: start of front, ccl, quad, feature, decode, end (ms since detector creation)\n")
            for i, row in enumerate(timeline):
                f.write(f"{i} " + " ".join(f"{v:.3f}" for v in row) + "\n")

    # a run long enough to see the clocks settle (the K-step region above lasts a fraction of a second)
    sustained = None
    if a.sustain > 0:
        s2 = ClockSampler(local).start()
        ms2, steps2, _, _, _ = run.pipelined(0, barrier, seconds=a.sustain)
        c2 = s2.stop()
        ms2 = max_over_ranks(ms2)
        steps2 = int(-max_over_ranks(-steps2))  # min over ranks
        sustained = {"seconds": ms2 / 1000.0, "steps": steps2, "value": n * steps2 * world / (ms2 / 1000.0), "unit": "frames/s", "clocks": c2}

    stage_acc = run.stage_times(max(3, min(a.steps, 10)))
    quad_counters = det.debug_quad_counters()

    # ---- end to end through the host-buffer C-ABI call (H2D + detect + D2H inside the timed region) ----
    e2e = None
    if not a.no_e2e:
        host_np = run.host.numpy()
        e2e_steps = max(2, min(a.steps, 5))
        det.detect_batch(host_np, 5, True, 5, run.cap)  # warm-up (staging allocation)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            m2, c2_, i2 = det.detect_batch(host_np, 5, True, 5, run.cap)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        d2h = int(c2_.sum()) * _capi.MARKER_DTYPE.itemsize + n * 48
        e2e = {"value": n * e2e_steps * world / dt, "unit": "frames/s", "h2d_bytes_per_step": int(n * run.fstride),
               "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": 1000.0 * dt / e2e_steps,
               "note": "pinned host frames -> ctag_detect_batch(is_device=0); wall clock around synchronous calls, max over ranks",
               "numa": numa}
        # context: the same bytes through a bare pinned-host -> device copy, all ranks copying at the same time
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        run.frames.copy_(run.host, non_blocking=True)
        barrier()
        c0.record()
        for _ in range(3):
            run.frames.copy_(run.host, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        copy_ms = max_over_ranks(c0.elapsed_time(c1) / 3)
        e2e["h2d_copy_alone_ms_per_step"] = copy_ms
        e2e["h2d_copy_alone_gbs_per_gpu"] = n * run.fstride / copy_ms / 1e6
        e2e["h2d_copy_alone_gbs_aggregate"] = world * n * run.fstride / copy_ms / 1e6
        e2e["frac_of_bare_copy"] = copy_ms / e2e["ms_per_step"]

    # ---- compressed ingest: the same ring as JPEG byte strings, decoded on the GPU (nvJPEG) inside ctag_detect_batch_jpeg ----
    e2e_jpeg = None
    if not a.no_e2e and not a.no_jpeg:
        try:
            import cv2
            enc = [cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, 16])[1].reshape(-1) for f in ring]
            pinned = torch.empty(sum(e.size for e in enc), dtype=torch.uint8).pin_memory()  # like the raw e2e: pinned host memory
            jpegs, pos = [], 0
            for e in enc:
                pinned[pos:pos + e.size] = torch.from_numpy(e)
                jpegs.append(pinned[pos:pos + e.size].numpy())
                pos += e.size
            for _ in range(2):
                mj, cj, _ = det.detect_batch_jpeg(jpegs, 5, True, 5, run.cap)  # warm-up: decoder buffers of every workspace used
            barrier()
            js = max(2, min(a.steps, 5))
            t0 = time.perf_counter()
            for _ in range(js):
                mj, cj, _ = det.detect_batch_jpeg(jpegs, 5, True, 5, run.cap)
            torch.cuda.synchronize()
            dtj = max_over_ranks(time.perf_counter() - t0)
            e2e_jpeg = {"value": n * js * world / dtj, "unit": "frames/s", "h2d_bytes_per_step": int(sum(j.size for j in jpegs)),
                        "d2h_bytes_per_step": int(cj.sum()) * _capi.MARKER_DTYPE.itemsize + n * 48, "steps": js, "ms_per_step": 1000.0 * dtj / js,
                        "decoder": det.jpeg_backend(), "jpeg_quality": 90, "restart_interval_mcus": 16, "sampling": "4:2:0",
                        "compressed_mb_per_frame": sum(j.size for j in jpegs) / n / 1e6, "markers_decoded_per_step": int(cj.sum()),
                        "note": "JPEG byte strings in pinned host memory -> ctag_detect_batch_jpeg (decode on the GPU into the BGR staging "
                                "buffer, then the detect path); wall clock around synchronous calls; decoded pixels equal cv::imdecode's and "
                                "detections equal the reference's on them: tests/test_jpeg_gpu.py"}
            # the same ring four times per call (256 frames): the decoder's kernels run one thread per restart interval and
            # like many frames per launch
            big = jpegs * 4
            det.detect_batch_jpeg(big, 5, True, 5, run.cap)
            barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                det.detect_batch_jpeg(big, 5, True, 5, run.cap)
            torch.cuda.synchronize()
            e2e_jpeg["value_256_frames_per_call"] = len(big) * 2 * world / max_over_ranks(time.perf_counter() - t0)
        except Exception as exc:  # pragma: no cover
            e2e_jpeg = {"error": str(exc)}

    # ---- multi-GPU: the sharded result equals the single-GPU list ----
    multi_parity = None
    if world > 1:
        from cylindertag_b200.sharding import frame_shard, gather_detections
        common = np.stack(load_ring("config5", 0, 1))  # the unrotated ring: the same 64 frames on every rank
        s, e = frame_shard(len(common), rank, world)
        m_loc, c_loc, _ = det.detect_batch(common[s:e], 5, True, 5, run.cap)
        allm, allc = gather_detections(m_loc, c_loc, len(common), dist, device=f"cuda:{local}")
        if rank == 0:
            m1, c1_, _ = det.detect_batch(common, 5, True, 5, run.cap)
            same = bool(np.array_equal(allc, c1_) and allm.tobytes() == m1.tobytes())
            multi_parity = {"status": "ok" if same else "MISMATCH", "frames": int(len(common)), "markers": int(c1_.sum()), "ranks": world,
                            "how": "each rank detects its contiguous block of the 64-frame ring, records gathered over NCCL, compared "
                                   "byte for byte with rank 0 detecting all 64 frames alone"}

    # ---- BASELINE config 5 to the letter: 4096 DISTINCT synthetic 4K frames, sharded over the ranks.  No host could render
    # them in a bench run; the library's CUDA renderer (ctag_render_frames) writes them straight into HBM, 256 per rank at a
    # time (rendering is outside the timed regions), and the pipelined detect loop runs over each group of four batches. ----
    sweep = None
    if not a.no_sweep:
        from cylindertag_b200 import synth
        per_rank = 4096 // world
        group_batches = 4
        groups = max(1, per_rank // (group_batches * n))
        bufs = [torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev) for _ in range(group_batches)]
        stream = torch.cuda.ExternalStream(det.stream(), device=dev)
        tot_ms, tot_frames, tot_markers, render_s = 0.0, 0, 0, 0.0
        sweep_parity = None
        for g in range(groups):
            t0 = time.perf_counter()
            for b in range(group_batches):
                first = 100000 + ((rank * groups + g) * group_batches + b) * n
                synth.render_frames_gpu(det, bufs[b].data_ptr(), range(first, first + n), w, h, [4 + (first + i) % 5 for i in range(n)], 3)
            render_s += time.perf_counter() - t0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record(stream)
            for b in range(group_batches):
                det.enqueue_device(bufs[b].data_ptr(), n, w, h, w * 3, w * h * 3, 3, 5, True, 5)
            res = [det.collect(run.cap) for _ in range(group_batches)]
            e1.record(stream)
            torch.cuda.synchronize()
            tot_ms += e0.elapsed_time(e1)
            tot_frames += group_batches * n
            tot_markers += sum(int(r[1].sum()) for r in res)
            if g == 0 and rank == 0:
                sample = bufs[0][:16].cpu().numpy()
                sweep_parity = check_parity(res[0][0][:16], res[0][1][:16], sample, state, fs, os.cpu_count() or 1)
        del bufs
        tot_ms = max_over_ranks(tot_ms)
        sweep = {"workload": "BASELINE config 5 as written: 4096 distinct 3840x2160 BGR frames (4-8 markers each, seeds 100000+), rendered on the "
                             "GPU by ctag_render_frames, sharded over the ranks, 4 batches of 64 in flight per group", "frames": tot_frames * world,
                 "value": tot_frames * world / (tot_ms / 1000.0), "unit": "frames/s", "detect_ms_total": tot_ms,
                 "markers_decoded_rank0": tot_markers, "render_s_per_rank_untimed": render_s,
                 "render_frames_per_s_per_gpu": tot_frames / render_s if render_s > 0 else None, "parity_check_first_16_frames": sweep_parity}

    variants = {}
    if rank == 0 and world == 1 and not a.no_variants:
        import cv2
        # the ring as GRAY frames: detect()'s own input contract (header/CylinderTag.h:21), 1.25 algorithmic bytes per pixel
        gray = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in ring])
        rg = DeviceRun(torch, det, dev, gray, 1)
        rg.warm()
        msg, _, _, _, _ = rg.pipelined(a.steps, barrier)
        st = rg.stage_times(5)
        variants["4k_gray"] = {"workload": "the same ring as 8-bit gray frames (detect()'s own contract), batch 64", "value": n * a.steps / (msg / 1000.0),
                               "unit": "frames/s", "ms_per_step": msg / a.steps, "roofline": roofline_of(st["front"], w, h, n, 1, *gray_traffic(n)),
                               "stages_ms_per_step_unoverlapped": st}
        del rg
        # config 3 as written: 256 frames 1920x1080 BGR, one marker each
        c3 = np.stack(load_ring("config3", 0, 1))
        r3 = DeviceRun(torch, det, dev, c3, 3, cap=8)
        m3, k3, _ = r3.warm()
        ms3, _, _, _, _ = r3.pipelined(a.steps, barrier)
        st = r3.stage_times(5)
        v3 = {"workload": "BASELINE config 3 as written: 256 frames 1920x1080 BGR, one rendered 2f12c marker each (seeds 1000-1255), batch 256",
              "value": len(c3) * a.steps / (ms3 / 1000.0), "unit": "frames/s", "ms_per_step": ms3 / a.steps,
              "roofline": roofline_of(st["front"], 1920, 1080, len(c3), 3), "stages_ms_per_step_unoverlapped": st,
              "frac_of_dense_ceiling": len(c3) * a.steps / (ms3 / 1000.0) * 5.5 * 1920 * 1080 / 1e9 / hbm_peak()[0]}
        try:  # parity of all 256 frames against the frozen results of the compiled reference
            gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_config3.npz"))
            ok = all(int(k3[f]) == int(gold["marker_start"][f + 1] - gold["marker_start"][f]) and all(
                int(m3[f][k]["marker_id"]) == int(gold["marker_id"][gold["marker_start"][f] + k]) and
                float(np.abs(m3[f][k]["corners"][:int(gold["n_features"][gold["marker_start"][f] + k])] -
                             gold["corners"][gold["marker_start"][f] + k][:int(gold["n_features"][gold["marker_start"][f] + k])]).max()) <= 1e-3
                for k in range(int(k3[f]))) for f in range(len(c3)))
            v3["parity_check"] = {"status": "ok" if ok else "MISMATCH", "checker": "tests/golden/ref_config3.npz (frozen from oracle/_ref)",
                                  "frames_checked": len(c3), "markers_checked": int(k3.sum())}
        except Exception as exc:  # pragma: no cover
            v3["parity_check"] = {"status": f"error: {exc}"}
        variants["1080p_config3"] = v3
        del r3

    if rank == 0:
        cores = os.cpu_count() or 1
        ref_ring = np.stack(load_ring("config5", 0, 1)) if world > 1 else ring
        parity = check_parity(markers, counts, ring, state, fs, cores)
        traffic, traffic_src = (None, "not measured (--no-traffic or multi-GPU run)")
        if not a.no_traffic and world == 1:
            traffic, traffic_src = measure_traffic("front_bgr_slide_kernel")
        if traffic is None:
            try:
                td = json.load(open(os.path.join(ROOT, "profiles", "front_traffic.json"))).get("4k")
                if td and td.get("frames_per_launch") == n:
                    traffic, traffic_src = td["dram_bytes_total"], f"constant from {td.get('source')} ({traffic_src})"
            except Exception:
                pass
        peak = hbm_peak()[0]
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
                "data": "synthetic", "config": workload_config(),
                "roofline": roofline_of(stage_acc["front"], w, h, n, 3, traffic, traffic_src),
                "frac_of_dense_ceiling": {"value": (value / world) * 5.5 * w * h / 1e9 / peak, "bytes_per_frame": 5.5 * w * h,
                                          "ceiling_frames_per_s_per_gpu": peak * 1e9 / (5.5 * w * h),
                                          "note": "whole detect path per GPU against the S1+S2 dense traffic of BASELINE.md 4 at the measured HBM peak"},
                "stages_ms_per_step_unoverlapped": stage_acc, "pipelining": f"{depth} batches in flight, one stream each",
                "quad_fit_work": quad_counters, "gpu_launches": launches, "markers_decoded_per_step": n_markers / max(steps_done, 1), "parity_check": parity,
                "clocks": clocks, "timed_region_s": ms_max / 1000.0}
        if sustained:
            line["sustained"] = sustained
        if e2e:
            line["e2e"] = e2e
        if e2e_jpeg:
            line["e2e_jpeg"] = e2e_jpeg
        if sweep:
            line["config5_4096_distinct"] = sweep
        if multi_parity:
            line["multi_gpu_parity"] = multi_parity
        if variants:
            line["variants"] = variants
        if not a.no_e2e and world == 1:
            # warm latency of the single-frame entry point: host gray frame in, markers out, wall clock around ctag_detect
            try:
                import cv2
                det1 = Detector(state=state, feature_size=fs, device=local)
                sf = {}
                for name, img in (("test.bmp 1920x1200", cv2.imread(os.path.join(wl.DATA, "test_gray.png"), cv2.IMREAD_GRAYSCALE)),
                                  ("config-4 frame 0, 3840x2160", cv2.cvtColor(ring[0], cv2.COLOR_BGR2GRAY))):
                    for _ in range(5):
                        det1.detect(img, 5, True, 5, cap=64)
                    lat = []
                    for _ in range(50):
                        t0 = time.perf_counter()
                        rec1, _st = det1.detect(img, 5, True, 5, cap=64)
                        lat.append(time.perf_counter() - t0)
                    sf[name] = {"median_ms": float(np.median(lat) * 1e3), "markers": int(len(rec1))}
                det1.close()
                line["single_frame"] = {"call": "ctag_detect(gray host frame, adaptiveThresh=5, cornerSubPix=true, dist=5), markers out", **sf}
            except Exception as exc:  # pragma: no cover
                line["single_frame"] = {"error": str(exc)}
        if not a.no_cpu and world == 1:
            try:
                cpu_reference_pass(ref_ring[:cores], state, fs, cores)  # warm-up
                v, _ = cpu_reference_pass(ref_ring, state, fs, cores)
                v1, _ = cpu_reference_pass(ref_ring[:4], state, fs, 1)  # also one thread (4 frames)
                line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": cpu_sample_text(len(ref_ring), cores),
                                        "value_1_thread": v1}
            except Exception as exc:  # pragma: no cover
                line["cpu_baseline"] = {"error": str(exc)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    det.close()


if __name__ == "__main__":
    main()
