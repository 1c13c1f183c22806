#!/usr/bin/env python
"""bench.py -- stereo frames/s and us/frame of FeatureTracker::TrackImage on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (cv2 oracle) on the host cores

A "step" is one TrackImage call (one stereo frame) per stream; the workload is BASELINE configs[1]:
752x480 EuRoC-shape synthetic stereo, cam0/cam1 MEI, max_cnt 150, min_dist 30, flow_back 1, maxLevel 3, 21x21 window,
one stream per GPU (frames of one stream depend on each other; more GPUs = more independent streams: weak scaling,
no collective on the data path).
  value : frames/s, inputs already resident in HBM when the timed region starts (K distinct frames, > L2 in total)
  e2e   : frames/s through the C-ABI call with HOST (pinned) images: H2D of both images and D2H of the records
          inside every step, host-synchronous like the reference's TrackImage
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, MAX_CNT, MIN_DIST, FLOW_BACK, LEVEL = 752, 480, 150, 30, 1, 3
METRIC = "trackImage stereo frames/s (752x480, 150 feats)"
WORKLOAD = ("configs[1]: EuRoC-shape 752x480 synthetic stereo, cam0/cam1 MEI, max_cnt 150, min_dist 30, flow_back 1, "
            "maxLevel 3, win 21x21, 1 stream per GPU")
L2_BYTES = 126 * 1024 * 1024
N_DISTINCT = 224          # distinct stereo frames resident in HBM: 224 * 722 KB = 154 MiB > 126 MiB L2


def ping_pong(k: int, d: int) -> int:
    period = 2 * (d - 1)
    m = k % period
    return m if m < d else period - m


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


SAMPLE_S = float(os.environ.get("VF_BENCH_SAMPLE_MS", "2")) * 1e-3   # NVML polling period of the clock sampler


class ClockSampler:
    """SM clock and throttle reasons DURING a timed region: NVML polled every ~2 ms from a thread (the regions are
    tens of milliseconds long, too short for `nvidia-smi -lms`); falls back to nvidia-smi when NVML is unavailable."""

    def __init__(self, gpu_index: int):
        self.idx, self.sm, self.mx, self.reasons, self.stop = gpu_index, [], [], set(), False
        self.t = None

    def _nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop:
            self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)); self.mx.append(mx)
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
            time.sleep(SAMPLE_S)

    def _smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.sm.append(float(out[0])); self.mx.append(float(out[1]))
                for k, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(k)
            except Exception:
                return

    def _run(self):
        try:
            self._nvml()
        except Exception:
            self._smi()

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        time.sleep(0.01)
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def make_frames(seed: int, n: int):
    from vins_fast_b200.synth import StereoSequence
    seq = StereoSequence(seed, W, H)
    fr = np.empty((n, 2, H, W), np.uint8)
    for k in range(n):
        _, fr[k, 0], fr[k, 1] = seq.frame(k)
    return fr


def run_reference(args, rank: int, world: int):
    """The reference's own CPU implementation of the path: TrackImage glue over the real OpenCV routines (cv2),
    all host threads; plain-C restatement if cv2 is unavailable.  Rank 0 only."""
    if rank != 0:
        return
    from oracle.camera import EUROC_CAM0, EUROC_CAM1
    n = args.warmup + args.steps
    frames = make_frames(0, N_DISTINCT)
    fidx = [ping_pong(k, N_DISTINCT) for k in range(n)]
    cores = os.cpu_count() or 1
    try:
        import cv2
        from oracle.cv2_oracle import Cv2FeatureTracker
        cv2.setNumThreads(cores)
        tr = Cv2FeatureTracker(EUROC_CAM0, EUROC_CAM1, MAX_CNT, MIN_DIST, FLOW_BACK, LEVEL)
        kind, used = "port", cv2.getNumThreads()
        sample = f"cv2 {cv2.__version__} restatement of TrackImage (oracle/cv2_oracle.py), {used} OpenCV threads"
    except Exception:
        from oracle import c_oracle
        tr = c_oracle.CFeatureTracker(EUROC_CAM0, EUROC_CAM1, W, H, MAX_CNT, MIN_DIST, FLOW_BACK, LEVEL)
        kind, used, sample = "port", 1, "plain-C restatement (oracle/vf_oracle.c), 1 thread"
    step = lambda k: tr.track_image(k / 20.0, frames[fidx[k], 0], frames[fidx[k], 1])
    for k in range(args.warmup):
        step(k)
    t0 = time.perf_counter()
    for k in range(args.warmup, n):
        step(k)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "us_per_frame": dt / args.steps * 1e6,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i16/f32 (OpenCV CPU)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "host_cores": cores},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": used, "kind": kind,
                         "sample": f"{args.steps} consecutive frames of the same sequence; {sample}; host has {cores} cores"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(frames, n_frames: int):
    from oracle.camera import EUROC_CAM0, EUROC_CAM1
    cores = os.cpu_count() or 1
    out = {}
    try:
        import cv2
        from oracle.cv2_oracle import Cv2FeatureTracker
        for threads in (cores, 1):
            cv2.setNumThreads(threads)
            tr = Cv2FeatureTracker(EUROC_CAM0, EUROC_CAM1, MAX_CNT, MIN_DIST, FLOW_BACK, LEVEL)
            n = min(n_frames, len(frames)) if threads == cores else min(n_frames // 2, len(frames))
            for k in range(3):
                tr.track_image(k / 20.0, frames[k, 0], frames[k, 1])
            t0 = time.perf_counter()
            for k in range(3, n):
                tr.track_image(k / 20.0, frames[k, 0], frames[k, 1])
            out[threads] = (n - 3) / (time.perf_counter() - t0)
        cv2.setNumThreads(cores)
        return {"value": out[cores], "unit": "frames/s", "cores": cores, "kind": "port", "value_1_thread": out[1],
                "sample": f"{min(n_frames, len(frames)) - 3} consecutive frames of the bench sequence through the cv2 {cv2.__version__} "
                          f"restatement of TrackImage (same OpenCV routines the reference calls); {cores} host cores"}
    except Exception as e:  # cv2 missing: time the plain-C restatement
        from oracle import c_oracle
        tr = c_oracle.CFeatureTracker(EUROC_CAM0, EUROC_CAM1, W, H, MAX_CNT, MIN_DIST, FLOW_BACK, LEVEL)
        n = min(n_frames, len(frames))
        t0 = time.perf_counter()
        for k in range(n):
            tr.track_image(k / 20.0, frames[k, 0], frames[k, 1])
        return {"value": n / (time.perf_counter() - t0), "unit": "frames/s", "cores": 1, "kind": "port",
                "sample": f"{n} frames, plain-C restatement oracle/vf_oracle.c (cv2 unavailable: {e})"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--multi-streams", type=int, default=64, help="streams per GPU for the extra multi-stream line (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the tracker has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import vins_fast_b200 as vf
    from vins_fast_b200 import build as vf_build
    from oracle.camera import EUROC_CAM0, EUROC_CAM1
    vf_build.build()

    K, Wm = args.steps, args.warmup
    # D distinct frames (> L2 in total), visited forward then backward (0..D-1, D-2..0, ...): consecutive steps always
    # see neighbouring frames of the sequence, and every frame is re-read only after all the others went through L2.
    D = N_DISTINCT
    frames = make_frames(rank, D)                                # one independent sequence per rank
    frame_bytes = 2 * H * W
    d_frames = torch.from_numpy(frames).cuda()                   # inputs resident in HBM
    p_frames = torch.from_numpy(frames).pin_memory()             # pinned host copies for the e2e leg
    n_frames = Wm + K
    times = np.arange(n_frames + 64) / 20.0
    fidx = [ping_pong(k, D) for k in range(n_frames + 64)]
    # the library owns its streams (the frame-to-frame chain runs on a high-priority one); the timing events are recorded
    # on that stream, the one the kernels are launched on
    cfg = vf.make_config(W, H, MAX_CNT, MIN_DIST, FLOW_BACK, EUROC_CAM0.__dict__, EUROC_CAM1.__dict__, lk_max_level=LEVEL,
                         device=local)
    batch = vf.Batch(cfg, 1)
    stream = torch.cuda.ExternalStream(batch.cuda_stream, device=torch.device("cuda", local))
    base = d_frames.data_ptr()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- value: device-resident inputs ----------------
    def run_device(first: int, count: int):
        for k in range(first, first + count):
            o = base + fidx[k] * frame_bytes
            batch.enqueue_device(times[k], o, o + H * W, W, frame_bytes)
        return batch.fetch_counts()

    with torch.cuda.stream(stream):
        run_device(0, Wm)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            h0 = batch.host_times_us()
            e0.record(stream)
            n_last = int(run_device(Wm, K)[0])
            e1.record(stream)
            h1 = batch.host_times_us()
            barrier()
        ms_value = max_over_ranks(e0.elapsed_time(e1))
        host_issue_us = max_over_ranks((h1[1] - h0[1]) / K)      # host time spent issuing one frame (slowest rank)
        launches = batch.launches_per_frame * K

        # ---------------- e2e: pinned host images in, records out, host-synchronous per frame ----------------
        batch.reset()
        out_n = None
        hp = p_frames.data_ptr()
        ptr_l = [(C.c_void_p * 1)(hp + fidx[k] * frame_bytes) for k in range(n_frames)]
        ptr_r = [(C.c_void_p * 1)(hp + fidx[k] * frame_bytes + H * W) for k in range(n_frames)]
        for k in range(Wm):
            batch.track_ptrs(times[k], ptr_l[k], W, ptr_r[k], W)
        barrier()
        with ClockSampler(local) as clk2:
            e0.record(stream)
            t_wall = time.perf_counter()
            for k in range(Wm, Wm + K):
                out_n = batch.track_ptrs(times[k], ptr_l[k], W, ptr_r[k], W)
            wall_ms = (time.perf_counter() - t_wall) * 1e3
            e1.record(stream)
            barrier()
        ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
        assert int(out_n[0]) == n_last, "device-resident and host paths disagree on the feature count"

        # ---------------- e2e, pipelined: enqueue frame k from pinned host images, collect frame k-lag meanwhile ----------------
        def run_pipelined(lag):
            batch.reset()
            for k in range(Wm):
                batch.enqueue_ptrs(times[k], ptr_l[k], W, ptr_r[k], W)
                if k >= lag:
                    batch.fetch_counts_lagged(lag)
            barrier()
            t0 = time.perf_counter()
            for k in range(Wm, Wm + K):
                batch.enqueue_ptrs(times[k], ptr_l[k], W, ptr_r[k], W)
                batch.fetch_counts_lagged(lag)            # every frame's records reach the host inside the timed region
            for back in range(lag - 1, -1, -1):
                n = batch.fetch_counts_lagged(back)
            torch.cuda.synchronize()
            ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
            assert int(n[0]) == n_last, "pipelined and synchronous host paths disagree on the feature count"
            return ms
        ms_e2e_pipe = run_pipelined(1)
        ms_e2e_pipe2 = run_pipelined(2)

        # ---------------- per-kernel times (CUDA events on the launching streams) ----------------
        batch.reset()
        batch.set_profiling(True)
        acc, n_prof, iters_t, iters_s, feats, fast_it = {}, 0, 0, 0, 0, 0
        from vins_fast_b200 import _lib as L
        for k in range(min(n_frames, Wm + 60)):
            batch.track_ptrs(times[k], ptr_l[k], W, ptr_r[k], W)
            if k >= Wm:
                for name, ms in batch.stage_times_ms().items():
                    acc[name] = acc.get(name, 0.0) + ms
                n_prof += 1
                it_t = batch.debug(0, L.DBG_LK_ITERS, 0, np.int32)
                it_s = batch.debug(0, L.DBG_LK_ITERS, 1, np.int32)
                iters_t += int((it_t & 0xffff).sum()); iters_s += int((it_s & 0xffff).sum()); feats += len(it_s)
                fast_it += int((it_t >> 16).sum()) + int((it_s >> 16).sum())
        batch.set_profiling(False)
        stages_us = {k: v / max(n_prof, 1) * 1e3 for k, v in acc.items()}

        # ---------------- kernel timeline of a synchronous frame on the GPU's own nanosecond timer ----------------
        # (every kernel stamps its first start / last end when tracing is on: no event overhead, true launch gaps)
        timeline = None
        try:
            os.environ["VF_TRACE"] = "1"
            tb = vf.Batch(vf.make_config(W, H, MAX_CNT, MIN_DIST, FLOW_BACK, EUROC_CAM0.__dict__, EUROC_CAM1.__dict__,
                                         lk_max_level=LEVEL, device=local), 1)
            del os.environ["VF_TRACE"]
            tnames = ("clear_frame pyrdown_left pad_left lk_temporal select_tracked response_nms bucket_scan bucket_scatter "
                      "mask_and_max gftt_select pyrdown_right pad_right lk_stereo undistort_left").split()
            tbuf = (C.c_ulonglong * 256)()
            tacc = []
            for k in range(min(n_frames, Wm + 40)):
                tb.track_ptrs(times[k], ptr_l[k], W, ptr_r[k], W)
                vf.load().vf_batch_trace_get(tb._h, tbuf, 256)
                if k >= Wm:
                    a = np.array(list(tbuf), np.uint64)[:192].reshape(6, 16, 2)[k % 6, :14]
                    tacc.append((a - a[:, 0].min()).astype(np.int64))
            tm = np.mean(tacc, 0) / 1e3
            timeline = {"kernels": {n: {"start_us": round(float(tm[i, 0]), 2), "end_us": round(float(tm[i, 1]), 2),
                                        "us": round(float(tm[i, 1] - tm[i, 0]), 2)} for i, n in enumerate(tnames)},
                        "frame_span_us": round(float(tm[:, 1].max()), 2),
                        "note": "synchronous frames, mean of 40; t = 0 at the first kernel start; the uploads precede it"}
            tb.close()
        except Exception as e:  # tracing is a measurement aid; the bench line does not depend on it
            os.environ.pop("VF_TRACE", None)
            timeline = {"error": str(e)}

        # ---------------- extra: many independent streams on one GPU (BASELINE config 4 shape) ----------------
        multi = None
        if args.multi_streams > 0:
            S = args.multi_streams
            mb = vf.Batch(vf.make_config(W, H, MAX_CNT, MIN_DIST, FLOW_BACK, EUROC_CAM0.__dict__, EUROC_CAM1.__dict__,
                                         lk_max_level=LEVEL, device=local), S)
            mstream = torch.cuda.ExternalStream(mb.cuda_stream, device=torch.device("cuda", local))
            mt = lambda k: times[k] + np.zeros(S)
            steps_m, warm_m = 30, 5
            for k in range(warm_m):       # stream s sees frames k+s, k+s+1, ... : consecutive frames of the same sequence
                mb.enqueue_device(mt(k), base + k * frame_bytes, base + k * frame_bytes + H * W, W, frame_bytes)
            mb.fetch_counts()
            barrier()
            e0.record(mstream)
            for k in range(warm_m, warm_m + steps_m):
                mb.enqueue_device(mt(k), base + k * frame_bytes, base + k * frame_bytes + H * W, W, frame_bytes)
            mb.fetch_counts()
            e1.record(mstream)
            barrier()
            ms_m = max_over_ranks(e0.elapsed_time(e1))
            multi = {"streams_per_gpu": S, "steps": steps_m, "value": world * S * steps_m / (ms_m * 1e-3), "unit": "frames/s",
                     "us_per_frame_per_stream": ms_m * 1e3 / (steps_m * S), "ms_per_step": ms_m / steps_m,
                     "launches_per_step": mb.launches_per_frame}
            mb.close()

    value = world * K / (ms_value * 1e-3)
    e2e = world * K / (ms_e2e * 1e-3)
    clocks = clk.summary()
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(clocks["reasons"])

    # ---------------- roofline of the dominant kernel ----------------
    kernels = {k: v for k, v in stages_us.items() if not k.startswith(("h2d", "d2h"))}
    dom = max(kernels, key=kernels.get) if kernels else "lk_stereo"
    peak, peak_src = peaks()
    passes_levels = (LEVEL + 1) * 2                      # stereo: two LK calls over levels 0..3
    if dom == "lk_temporal":
        passes_levels = (LEVEL + 1) + 2                  # forward L3 + reverse L1
        iters = iters_t
    else:
        iters = iters_s
    # algorithmic bytes of an LK launch: per point and level one 24x24 uint8 tile of I, per iteration one
    # 22x22 uint8 window of J; the derivative planes are never materialised.
    lk_bytes = (feats / max(n_prof, 1)) * passes_levels * 24 * 24 + (iters / max(n_prof, 1)) * 22 * 22
    npx = W * H
    alg_bytes = {
        "lk_stereo": lk_bytes, "lk_temporal": lk_bytes,
        # fused pyramid (read level 0 once, write levels 1..3) + border kernel (read level 0 from staging, write the padded
        # level 0 and the 28-pixel borders of levels 1..3)
        "pyrdown_left": npx * 1 + npx * (0.25 + 0.0625 + 0.015625) + npx * 1 + (H + 56) * (W + 64) + 0.35 * npx,
        "pyrdown_right": npx * 1 + npx * (0.25 + 0.0625 + 0.015625) + npx * 1 + (H + 56) * (W + 64) + 0.35 * npx,
        "response_nms": npx * 1 + npx * 4,                                            # read u8 image, write f32 response
        "mask_and_max": npx * 4 + npx * 1,                                            # read f32 response, write u8 mask
    }
    dur_us = kernels.get(dom, 0.0)
    ach = (alg_bytes.get(dom, 0.0) / (dur_us * 1e-6) / 1e9) if dur_us > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dom)
        except Exception:
            traffic = None
    roofline = {
        "kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None,
        "traffic": traffic, "peak_source": peak_src, "kernel_us": dur_us,
        "note": "the frame's working set (~2 MB) is L2-resident and the dominant kernel is a dependent chain of float "
                "additions (bit-exact OpenCV summation order): latency-bound, not HBM-bound; the HBM fraction is reported "
                "as asked, per-kernel device times are in `stages_us`",
        "streaming_kernels": {k: {"achieved_GBps": alg_bytes[k] / (kernels[k] * 1e-6) / 1e9, "frac": alg_bytes[k] / (kernels[k] * 1e-6) / 1e9 / peak}
                              for k in ("pyrdown_left", "response_nms", "mask_and_max") if kernels.get(k, 0) > 0},
    }

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_value / K, "us_per_frame": ms_value / K * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/i32 stencils, f32 LK + response, f64 undistort", "data": "synthetic",
        "config": {"workload": WORKLOAD, "streams_per_gpu": 1, "features_last_frame": n_last,
                   "l2": f"inputs larger than L2: {D} distinct stereo frames = {D * frame_bytes / 2**20:.0f} MiB resident in HBM "
                         f"(L2 = 126 MiB), visited forward/backward so a frame is re-read only after all others; the tracker's "
                         f"own state (previous pyramid, 0.5 MB) is L2-resident by design",
                   "cuda_graph": bool(cfg.use_cuda_graph),
                   "value_mode": "frames enqueued asynchronously (vf_batch_enqueue_device), records of every frame downloaded; the "
                                 "stereo pass of frame k overlaps the temporal chain of frame k+1; K steps timed with CUDA events "
                                 "on the stream the chain is launched on"},
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                   "samples": clocks["samples"], "rejected": bool(bad)},
        "e2e": {"value": e2e, "unit": "frames/s", "us_per_frame": ms_e2e / K * 1e3, "h2d_bytes_per_step": frame_bytes + 8,
                "d2h_bytes_per_step": MAX_CNT * 60 + 64, "clocks": clk2.summary(),
                "mode": "host-synchronous vf_batch_track per frame (what the reference's TrackImage does)",
                "pipelined": {"value": world * K / (ms_e2e_pipe * 1e-3), "unit": "frames/s", "us_per_frame": ms_e2e_pipe / K * 1e3,
                              "mode": "vf_batch_enqueue(k) + vf_batch_fetch_lagged(1): frame k-1 is collected while frame k runs; "
                                      "same H2D / D2H bytes per frame, host wall clock"},
                "pipelined_lag2": {"value": world * K / (ms_e2e_pipe2 * 1e-3), "unit": "frames/s", "us_per_frame": ms_e2e_pipe2 / K * 1e3,
                                   "mode": "same with vf_batch_fetch_lagged(2): the host never waits for a running frame"}},
        "gpu_launches": launches, "launches_per_frame": batch.launches_per_frame, "host_issue_us_per_frame": host_issue_us,
        "stages_us": stages_us, "timeline": timeline, "roofline": roofline, "multi_stream": multi,
        "lk": {"iterations_per_frame": (iters_t + iters_s) / max(n_prof, 1),
               "exact_integer_path_fraction": fast_it / max(iters_t + iters_s, 1)},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(frames, 160)
    batch.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
