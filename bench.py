#!/usr/bin/env python
"""Benchmark of the CylinderTag detect hot path (BASELINE.json metric: detect frames/s at 4K / 1080p).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--res 4k|1080p] [--batch B] [--impl ours|reference]

A "step" is one pass of the whole detect path (BGR -> gray -> 2x cubic decimation -> adaptive threshold -> CCL ->
quad extraction -> pairing -> edge refinement -> decode) over one batch of B synthetic frames.
  value  : frames/s with the batch already resident in HBM (CUDA events on the detector's stream, max over ranks)
  e2e    : frames/s through the public C-ABI call with HOST (pinned) frames: H2D copy + detect + D2H of the markers
  roofline: dominant kernel (fused front end, 4.25 algorithmic bytes per full-res pixel) vs the measured HBM peak
  cpu_baseline: the CPU restatement of the reference path timed on this box's host cores (bounded sample)
`--impl reference` times the reference's CPU algorithm (oracle/) on all host cores for the same metric/config.
Under torchrun each rank owns one GPU and its own frames (weak scaling, no data-path collective).
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

RES = {"4k": (3840, 2160), "1080p": (1920, 1080)}
DATA = os.path.join(ROOT, "tests", "golden", "data")


def load_dictionary():
    toks = open(os.path.join(DATA, "CTag_2f12c.marker")).read().split()
    n, cols, fs = int(toks[0]), int(toks[1]), int(toks[2])
    return np.array([int(t) for t in toks[3:3 + n * cols]], np.int32).reshape(n, cols), fs


def _render(args):
    seed, w, h, nm = args
    cache = f"/tmp/ctag_bench_{w}x{h}_{nm}_{seed}.npy"
    if os.path.exists(cache):
        try:
            return np.load(cache)
        except Exception:
            pass
    from cylindertag_b200 import synth
    state, _ = load_dictionary()
    frame, _ = synth.synthetic_frame(seed, w, h, state, nm, channels=3)
    try:
        np.save(cache + f".{os.getpid()}.tmp.npy", frame)
        os.replace(cache + f".{os.getpid()}.tmp.npy", cache)
    except Exception:
        pass
    return frame


def render_frames(seeds, w, h, nm, workers):
    import multiprocessing as mp
    jobs = [(s, w, h, nm) for s in seeds]
    if workers <= 1 or len(jobs) == 1:
        return [_render(j) for j in jobs]
    with mp.get_context("spawn").Pool(min(workers, len(jobs))) as pool:
        return pool.map(_render, jobs)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
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
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def hbm_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def oracle_frame_job(args):
    seed, w, h, nm = args
    import cv2
    cv2.setNumThreads(1)
    from oracle import ctag_oracle as o
    frame = _render((seed, w, h, nm))
    state, fs = load_dictionary()
    t = time.perf_counter()
    gray = o.bgr2gray(frame)
    d = o.detect(gray, state, fs, 5, True, 5)
    return time.perf_counter() - t, len(d.markers)


def cpu_port_available():
    try:
        from oracle.cpu_ref import api as cpu_api
        cpu_api.load()
        return cpu_api
    except Exception:
        return None


def run_cpu_baseline(w, h, nm, seeds, threads, repeat=1):
    """Times the CPU restatement of the reference path on `threads` host threads over the given frames
    (`repeat` passes over the rendered set).  Returns (frames/s, threads, sample description, markers decoded)."""
    cpu_api = cpu_port_available()
    state, fs = load_dictionary()
    if cpu_api is not None:
        distinct = np.stack(render_frames(seeds, w, h, nm, min(8, os.cpu_count() or 1)))
        frames = np.concatenate([distinct] * repeat) if repeat > 1 else distinct
        cpu_api.detect_batch_bgr(frames[:max(1, min(len(frames), threads))], state, fs, 5, True, 5, threads)  # warm-up
        t = time.perf_counter()
        n_mk = cpu_api.detect_batch_bgr(frames, state, fs, 5, True, 5, threads)
        dt = time.perf_counter() - t
        return (len(frames) / dt, threads, f"{len(frames)} frames {w}x{h} BGR ({len(seeds)} distinct), C++ restatement "
                f"oracle/cpu_ref (-O3, no -march, one frame per thread), {threads} thread(s)", n_mk)
    # Python + cv2 oracle (every OpenCV call of the reference is the real library call; the glue is Python)
    if threads <= 1:
        t = time.perf_counter()
        n_mk = sum(oracle_frame_job((s, w, h, nm))[1] for s in seeds)
        dt = time.perf_counter() - t
    else:
        import multiprocessing as mp
        render_frames(seeds, w, h, nm, threads)  # warm the frame cache outside the timed region
        with mp.get_context("spawn").Pool(threads) as pool:
            pool.map(oracle_frame_job, [(seeds[0], w, h, nm)] * threads)  # import + warm-up
            t = time.perf_counter()
            res = pool.map(oracle_frame_job, [(s, w, h, nm) for s in seeds])
            dt = time.perf_counter() - t
        n_mk = sum(r[1] for r in res)
    return len(seeds) / dt, threads, f"{len(seeds)} frames {w}x{h} BGR, Python+cv2 oracle, {threads} process(es)", n_mk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--res", default="4k", choices=list(RES))
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--markers", type=int, default=6)
    ap.add_argument("--distinct", type=int, default=0, help="distinct rendered frames in the ring (0: 16 at 4K, 32 at 1080p)")
    ap.add_argument("--channels", type=int, default=3, choices=[1, 3], help="3 = BGR frames (caller's cvtColor fused in), 1 = gray frames (detect's own contract)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--timeline", default="", help="write the stage timeline of the timed batches (ms, per batch) to this file")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    w, h = RES[a.res]
    if a.distinct <= 0:
        a.distinct = 16 if a.res == "4k" else 32
    a.distinct = min(a.distinct, a.batch)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cname = "BGR" if a.channels == 3 else "gray"
    workload = (f"{a.res} ({w}x{h}) {cname} synthetic frames, {a.markers} rendered 2f12c markers each, batch {a.batch} "
                f"(ring of {a.distinct} distinct frames, batch input {a.batch * w * h * a.channels / 1e6:.0f} MB vs 126 MB L2), "
                f"detect(adaptiveThresh=5, cornerSubPix=true, dist=5)" + (" incl. the caller's BGR2GRAY" if a.channels == 3 else ""))
    config = {"workload": workload, "resolution": a.res, "batch": a.batch, "markers_per_frame": a.markers,
              "channels": a.channels, "l2_policy": "inputs larger than L2", "frames_per_step_per_gpu": a.batch}

    if a.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        seeds = [2000 + i for i in range(a.distinct)]
        repeat = max(1, a.batch // a.distinct) if cpu_port_available() is not None else 1
        per_step = len(seeds) * repeat
        for _ in range(min(a.warmup, 1)):
            run_cpu_baseline(w, h, a.markers, seeds[:2], min(cores, 2))
        t0 = time.perf_counter()
        fps = []
        for _ in range(max(a.steps, 1)):
            v, thr, sample, _ = run_cpu_baseline(w, h, a.markers, seeds, cores, repeat)
            fps.append(v)
            if time.perf_counter() - t0 > 150:  # keep the whole arm within a few minutes whatever K is
                break
        wall = time.perf_counter() - t0
        value = float(np.mean(fps))
        kind = "port"
        line = {"impl": "reference", "metric": "detect_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": a.gpus,
                "steps": len(fps), "warmup": a.warmup, "ms_per_step": 1000.0 * per_step / value, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 dense, f32/f64 sparse", "data": "synthetic",
                "config": dict(config, frames_per_step=per_step),
                "cpu_baseline": {"value": value, "unit": "frames/s", "cores": thr, "kind": kind, "sample": sample},
                "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "wall_s": wall}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from cylindertag_b200 import Detector, _capi
    _capi.load()  # fail loudly if the CUDA extension is missing
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    state, fs = load_dictionary()
    seeds = [2000 + rank * a.distinct + i for i in range(a.distinct)]
    distinct = render_frames(seeds, w, h, a.markers, min(8, max(1, (os.cpu_count() or 8) // max(world, 1))))
    ch = a.channels
    if ch == 1:
        import cv2
        distinct = [cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in distinct]
    host = torch.empty((a.batch, h, w, 3) if ch == 3 else (a.batch, h, w), dtype=torch.uint8).pin_memory()
    for i in range(a.batch):
        host[i] = torch.from_numpy(distinct[i % a.distinct])
    frames = host.to(dev, non_blocking=False)
    pitch, fstride = w * ch, w * h * ch
    det = Detector(state=state, feature_size=fs, device=local)
    cap = 16

    def step_device():
        det.enqueue_device(frames.data_ptr(), a.batch, w, h, pitch, fstride, ch, 5, True, 5)
        return det.collect(cap)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def enqueue():
        det.enqueue_device(frames.data_ptr(), a.batch, w, h, pitch, fstride, ch, 5, True, 5)

    for _ in range(a.warmup):
        markers, counts, info = step_device()
    # the detector keeps several batches in flight (one workspace + stream each): enqueue ahead, collect in FIFO order
    depth = det.max_in_flight()
    for _ in range(depth):
        enqueue()
    for _ in range(depth):
        det.collect(cap)
    stream = torch.cuda.ExternalStream(det.stream(), device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    launches = 0
    n_markers = 0
    timeline = []
    barrier()
    sampler.start()
    ev0.record(stream)  # both detector streams are idle here, so the event is stamped immediately
    queued = 0
    while queued < min(depth - 1, a.steps):
        enqueue()
        queued += 1
    for i in range(a.steps):
        if queued < a.steps:
            enqueue()
            queued += 1
        markers, counts, info = det.collect(cap)
        if a.timeline:
            timeline.append(det.stage_timeline_ms())
        launches += det.launch_count()
        n_markers += int(counts.sum())
    ev1.record(stream)  # every collect synchronised its stream: stamped now, after the last batch finished
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    if a.timeline and rank == 0:
        with open(a.timeline, "w") as f:
            f.write("# batch: start of front, ccl, quad, feature, decode, end (ms since detector creation)\n")
            for i, row in enumerate(timeline):
                f.write(f"{i} " + " ".join(f"{v:.3f}" for v in row) + "\n")
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_frames = a.batch * a.steps * world
    value = total_frames / (ms_max / 1000.0)

    # per-stage / per-kernel times for the roofline: batches run one at a time here, so that no other kernel shares
    # the GPU with the one being timed (CUDA events recorded by the library on the launching stream)
    stage_acc = {k: 0.0 for k in _capi.STAGE_NAMES}
    seq_steps = max(3, min(a.steps, 10))
    for _ in range(seq_steps):
        step_device()
        for k, v in det.stage_times_ms().items():
            stage_acc[k] += v
    stage_acc = {k: v / seq_steps for k, v in stage_acc.items()}

    # ---- end to end through the host-buffer C-ABI call (H2D + detect + D2H inside the timed region) ----
    e2e = None
    if not a.no_e2e:
        host_np = host.numpy()
        e2e_steps = max(1, min(a.steps, 5))
        det.detect_batch(host_np, 5, True, 5, cap)  # warm-up (staging allocation)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            m2, c2, i2 = det.detect_batch(host_np, 5, True, 5, cap)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        d2h = int(c2.sum()) * _capi.MARKER_DTYPE.itemsize + a.batch * 48
        e2e = {"value": a.batch * e2e_steps * world / float(t.item()), "unit": "frames/s",
               "h2d_bytes_per_step": int(a.batch * fstride), "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "note": "pinned host frames -> ctag_detect_batch(is_device=0); wall clock around synchronous calls"}
        # for context: the same bytes through a plain pinned-host -> device copy (nothing else running): how much of the
        # end-to-end time is the PCIe transfer alone
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        frames.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        c0.record()
        for _ in range(3):
            frames.copy_(host, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        copy_ms = c0.elapsed_time(c1) / 3
        e2e["h2d_copy_alone_ms_per_step"] = copy_ms
        e2e["h2d_copy_alone_gbs"] = a.batch * fstride / copy_ms / 1e6
        e2e["ms_per_step"] = 1000.0 * float(t.item()) / e2e_steps

    if rank == 0:
        # parity spot check of one frame against the oracle (outside every timed region)
        parity = "skipped"
        try:
            from oracle import ctag_oracle as o
            dump = o.detect(o.bgr2gray(distinct[0]) if ch == 3 else distinct[0], state, fs, 5, True, 5)
            got = markers[0][:int(counts[0])]
            ok = len(dump.markers) == int(counts[0]) and all(
                int(g["marker_id"]) == m.markerID and list(g["feature_pos"][:len(m.featurePos)]) == m.featurePos and
                float(np.abs(g["corners"][:len(m.cornerLists)] - np.array(m.cornerLists)).max()) <= 1e-3
                for g, m in zip(got, dump.markers))
            parity = "ok" if ok else "MISMATCH"
        except Exception as exc:  # pragma: no cover
            parity = f"error: {exc}"
        peak, peak_src = hbm_peak()
        front_ms = stage_acc["front"]
        alg_bytes = (4.25 if ch == 3 else 1.25) * w * h * a.batch
        achieved = alg_bytes / (front_ms / 1000.0) / 1e9
        # DRAM bytes of one launch of this kernel from the committed `ncu --set full` capture (same workload shape);
        # null for shapes that were not captured
        traffic, traffic_detail = None, None
        try:
            traffic_detail = json.load(open(os.path.join(ROOT, "profiles", "front_traffic.json"))).get(a.res)
            if traffic_detail and ch == 3 and a.batch == traffic_detail.get("frames_per_launch"):
                traffic = traffic_detail["dram_bytes_total"]
            else:
                traffic_detail = None
        except Exception:
            pass
        line = {"metric": "detect_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8/int32 dense, f32/f64 sparse", "data": "synthetic", "config": config,
                "roofline": {"bound": "hbm", "kernel": ("front_bgr_slide_kernel (gray + " if ch == 3 else "front_gray_slide_kernel (") + "cubic decimation + adaptive threshold)",
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                             "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)", "traffic_detail": traffic_detail,
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                             "kernel_ms_per_launch": front_ms},
                "stages_ms_per_step_unoverlapped": stage_acc, "pipelining": f"{depth} batches in flight, one stream each",
                "gpu_launches": launches, "markers_decoded_per_step": n_markers / a.steps, "parity_check": parity,
                "clocks": clocks}
        if e2e:
            line["e2e"] = e2e
        if not a.no_e2e and world == 1:
            # informational: warm latency of the single-frame entry point on BASELINE config 0 (the reference's own
            # test image: host gray frame in, markers out, wall clock around the synchronous ctag_detect call)
            try:
                import cv2
                tb = cv2.imread(os.path.join(DATA, "test_gray.png"), cv2.IMREAD_GRAYSCALE)
                det1 = Detector(state=state, feature_size=fs, device=local)
                for _ in range(5):
                    det1.detect(tb, 5, True, 5, cap=64)
                lat = []
                for _ in range(50):
                    t0 = time.perf_counter()
                    rec1, _st = det1.detect(tb, 5, True, 5, cap=64)
                    lat.append(time.perf_counter() - t0)
                line["single_frame"] = {"workload": "test.bmp 1920x1200 gray, ctag_detect(adaptiveThresh=5, cornerSubPix=true, dist=5), host frame in, markers out",
                                        "median_ms": float(np.median(lat) * 1e3), "markers": int(len(rec1))}
            except Exception as exc:  # pragma: no cover
                line["single_frame"] = {"error": str(exc)}
        if not a.no_cpu and world == 1:
            cores = os.cpu_count() or 1
            if cpu_port_available() is not None:
                v, thr, sample, _ = run_cpu_baseline(w, h, a.markers, [2000 + i for i in range(a.distinct)], cores,
                                                     max(1, a.batch // a.distinct))
            else:
                v, thr, sample, _ = run_cpu_baseline(w, h, a.markers, [2000 + i for i in range(8)], 1)
            line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": thr, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    det.close()


if __name__ == "__main__":
    main()
