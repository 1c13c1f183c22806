#!/usr/bin/env python
"""bench.py -- stereo frames/s and us/frame of FeatureTracker::TrackImage on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (cv2 restatement) on the host cores

A "step" is one TrackImage call (one stereo frame) per stream; the workload is BASELINE configs[1]:
752x480 EuRoC-shape synthetic stereo, cam0/cam1 MEI, max_cnt 150, min_dist 30, flow_back 1, maxLevel 3, 21x21 window,
one stream per GPU (frames of one stream depend on each other; more GPUs = more independent streams: weak scaling,
no collective on the data path).
  value : frames/s, inputs already resident in HBM when the timed region starts (224 distinct frames, > L2 in total)
  e2e   : frames/s through the C-ABI call with HOST (pinned) images: H2D of both images and D2H of the records
          inside every step, host-synchronous like the reference's TrackImage
Every timed region is EXACTLY K steps between barrier + synchronize; it is repeated `--repeats` times on consecutive
frames of the sequence and the MEDIAN region is reported (all of them are listed), so that a 20-step region is not a
single 2 ms sample.  Also in the line: `parity` (the bench sequence against the CPU oracle, outside the timed regions),
`workloads` (BASELINE configs 3, 4, 5), `roofline`, `cpu_baseline`.  Prints ONE JSON line on rank 0.
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
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from vins_fast_b200.calib import WORKLOADS, workload_cameras  # noqa: E402  (data only)

W, H, MAX_CNT, MIN_DIST, FLOW_BACK, LEVEL = 752, 480, 150, 30, 1, 3
METRIC = "trackImage stereo frames/s (752x480, 150 feats)"
WORKLOAD = ("configs[1]: EuRoC-shape 752x480 synthetic stereo, cam0/cam1 MEI, max_cnt 150, min_dist 30, flow_back 1, "
            "maxLevel 3, win 21x21, 1 stream per GPU")
L2_BYTES = 126 * 1024 * 1024
N_DISTINCT = 224          # distinct stereo frames resident in HBM: 224 * 722 KB = 154 MiB > 126 MiB L2
N_SM, N_SCHED = 148, 4    # B200: 148 SMs x 4 warp schedulers, one warp instruction per scheduler and clock
ALL_LEGS = "value,e2e,pipelined,stages,timeline,multi,c3,c4,c5,win,parity,cpu"


def workload_config(n_streams_per_gpu: int = 1) -> dict:
    """The `config` object: the SAME keys and values in both arms (our kernels / the reference's CPU path)."""
    return {"workload": WORKLOAD, "width": W, "height": H, "max_cnt": MAX_CNT, "min_dist": MIN_DIST, "flow_back": FLOW_BACK,
            "lk_max_level": LEVEL, "lk_win": 21, "cameras": "EuRoC cam0/cam1 MEI (vins/config/euroc/cam{0,1}_mei.yaml)",
            "streams_per_gpu": n_streams_per_gpu, "distinct_frames": N_DISTINCT,
            "sequence": "vins_fast_b200.synth.StereoSequence(seed 0) on every rank -- equal work per GPU; --sequence-seeds rank: one seed per rank -- frames visited forward / backward",
            "l2": f"inputs larger than L2: {N_DISTINCT} distinct stereo frames = {N_DISTINCT * 2 * H * W / 2**20:.0f} MiB "
                  f"(L2 = 126 MiB); a frame is re-read only after all the others went through"}


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
        self.t, self.nv = None, None

    def _nvml_open(self):
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        self.nv = (nv, h, mx, names)

    def sample(self):
        """One sample now (also called by the timed loops themselves, from inside a region: a 1.4 ms region is shorter than
        the polling thread's period)."""
        if self.nv is None:
            return
        nv, h, mx, names = self.nv
        try:
            self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)); self.mx.append(mx)
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def _nvml(self):
        while not self.stop:
            self.sample()
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
        if self.nv is not None:
            self._nvml()
        else:
            self._smi()

    def __enter__(self):
        try:                      # NVML is opened HERE, synchronously: importing / initialising it takes longer than a short region
            self._nvml_open()
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def make_frames(seed: int, n: int, w: int = W, h: int = H, first: int = 0):
    """[n][2][h][w] uint8: frames first..first+n-1 of the synthetic sequence `seed` (host threads: numpy drops the GIL)."""
    from vins_fast_b200.synth import StereoSequence
    seq = StereoSequence(seed, w, h)
    fr = np.empty((n, 2, h, w), np.uint8)

    def one(k):
        _, fr[k, 0], fr[k, 1] = seq.frame(first + k)
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        list(ex.map(one, range(n)))
    return fr


def make_stream_frames(seeds, n: int, w: int = W, h: int = H):
    """[n][2][len(seeds)][h][w] uint8: frames 0..n-1 of one synthetic sequence per seed (canvases and frames on host threads)."""
    from vins_fast_b200.synth import StereoSequence
    out = np.empty((n, 2, len(seeds), h, w), np.uint8)
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        seqs = list(ex.map(lambda s: StereoSequence(s, w, h), seeds))

        def one(jk):
            j, k = jk
            _, out[k, 0, j], out[k, 1, j] = seqs[j].frame(k)
        list(ex.map(one, [(j, k) for j in range(len(seeds)) for k in range(n)]))
    return out


def med(xs):
    return float(statistics.median(xs))


# ======================================================================================================================
#  reference arm: the reference's own CPU implementation of the path
# ======================================================================================================================
def _ref_stream_worker(seed, threads, n_warm, n_steps, q_ready, q_go, q_out):
    """One stream of the reference arm in its own process (N streams side by side when --gpus N)."""
    import cv2
    from oracle.camera import EUROC_CAM0, EUROC_CAM1
    from oracle.cv2_oracle import Cv2FeatureTracker
    cv2.setNumThreads(threads)
    frames = make_frames(seed, min(N_DISTINCT, n_warm + n_steps))
    fidx = [ping_pong(k, len(frames)) for k in range(n_warm + n_steps)]
    tr = Cv2FeatureTracker(EUROC_CAM0, EUROC_CAM1, MAX_CNT, MIN_DIST, FLOW_BACK, LEVEL, keep_debug=False)
    for k in range(n_warm):
        tr.track_image(k / 20.0, frames[fidx[k], 0], frames[fidx[k], 1])
    q_ready.put(seed)
    q_go.get()
    tr.cv2_seconds = 0.0
    t0 = time.perf_counter()
    for k in range(n_warm, n_warm + n_steps):
        f = tr.track_image(k / 20.0, frames[fidx[k], 0], frames[fidx[k], 1])
    t1 = time.perf_counter()
    q_out.put((seed, t0, t1, tr.cv2_seconds, len(f["ids"])))


def c_oracle_rate(frames, n: int):
    """Frames/s of the plain-C restatement (oracle/vf_oracle.c), one thread, frames 0..n-1 (3 untimed)."""
    from oracle import c_oracle
    from oracle.camera import EUROC_CAM0, EUROC_CAM1
    tr = c_oracle.CFeatureTracker(EUROC_CAM0, EUROC_CAM1, W, H, MAX_CNT, MIN_DIST, FLOW_BACK, LEVEL)
    n = min(n, len(frames))
    for k in range(3):
        tr.track_image(k / 20.0, frames[k, 0], frames[k, 1])
    t0 = time.perf_counter()
    for k in range(3, n):
        tr.track_image(k / 20.0, frames[k, 0], frames[k, 1])
    return (n - 3) / (time.perf_counter() - t0)


def run_reference(args, rank: int, world: int):
    """TrackImage glue over the real OpenCV routines (cv2 wheel), the glue vectorised so that the interpreter is not what
    is timed; all host cores; with --gpus N, N independent streams side by side (cores / N OpenCV threads each).  Rank 0 only."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_streams = max(1, args.gpus)
    threads = max(1, cores // n_streams)
    try:
        import cv2
        ver = cv2.__version__
    except Exception as e:
        ver = None
        err = str(e)
    cfg = workload_config()
    if ver is None:   # no cv2 on this host: time the plain-C restatement instead (single thread)
        frames = make_frames(0, min(N_DISTINCT, args.warmup + args.steps))
        fps = c_oracle_rate(frames, args.warmup + args.steps)
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 / fps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8/i16/f32 (plain C)", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 1, "kind": "port", "n_streams": 1,
                                 "sample": f"plain-C restatement oracle/vf_oracle.c, 1 thread (cv2 unavailable: {err})"},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    ctx = mp.get_context("fork")
    q_ready, q_go, q_out = ctx.Queue(), ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_ref_stream_worker, args=(s, threads, args.warmup, args.steps, q_ready, q_go, q_out))
             for s in range(n_streams)]
    for p in procs:
        p.start()
    for _ in procs:
        q_ready.get()
    for _ in procs:
        q_go.put(1)
    res = [q_out.get() for _ in procs]
    for p in procs:
        p.join()
    dt = max(r[2] for r in res) - min(r[1] for r in res)            # all streams: first start .. last end
    fps = n_streams * args.steps / dt
    cv2_only = sum(args.steps / r[3] for r in res if r[3] > 0)       # frames/s if only the time inside OpenCV counted
    c_fps = c_oracle_rate(make_frames(0, 24), 24)
    sample = (f"{args.steps} consecutive frames per stream, {n_streams} stream(s) side by side, cv2 {ver} restatement of TrackImage "
              f"(oracle/cv2_oracle.py: the OpenCV routines the reference calls, numpy-vectorised glue), {threads} OpenCV threads per "
              f"stream; host has {cores} cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "us_per_frame": dt / args.steps / n_streams * 1e6,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i16/f32 (OpenCV CPU)", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads * n_streams, "kind": "port", "n_streams": n_streams,
                         "cv2_calls_only": cv2_only, "c_oracle_1_thread": c_fps, "sample": sample,
                         "note": "value = wall clock of the whole TrackImage restatement; cv2_calls_only = the same frames counting only "
                                 "the time inside the OpenCV routines (what a C++ caller with free glue would see); c_oracle_1_thread "
                                 "= oracle/vf_oracle.c, scalar C, one thread, 21 frames"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(frames, n_frames: int):
    """Rank 0, N=1: the CPU path on a bounded sample of the bench sequence (about 5 s of CPU work in total)."""
    from oracle.camera import EUROC_CAM0, EUROC_CAM1
    cores = os.cpu_count() or 1
    out, cv2_only = {}, {}
    c_fps = c_oracle_rate(frames, 40)
    try:
        import cv2
        from oracle.cv2_oracle import Cv2FeatureTracker
        for threads in (cores, 1):
            cv2.setNumThreads(threads)
            tr = Cv2FeatureTracker(EUROC_CAM0, EUROC_CAM1, MAX_CNT, MIN_DIST, FLOW_BACK, LEVEL, keep_debug=False)
            n = min(n_frames, len(frames)) if threads == cores else min(n_frames // 2, len(frames))
            for k in range(3):
                tr.track_image(k / 20.0, frames[k, 0], frames[k, 1])
            tr.cv2_seconds = 0.0
            t0 = time.perf_counter()
            for k in range(3, n):
                tr.track_image(k / 20.0, frames[k, 0], frames[k, 1])
            out[threads] = (n - 3) / (time.perf_counter() - t0)
            cv2_only[threads] = (n - 3) / tr.cv2_seconds
        cv2.setNumThreads(cores)
        return {"value": out[cores], "unit": "frames/s", "cores": cores, "kind": "port", "n_streams": 1,
                "cv2_calls_only": cv2_only[cores], "value_1_thread": out[1], "cv2_calls_only_1_thread": cv2_only[1],
                "c_oracle_1_thread": c_fps,
                "sample": f"{min(n_frames, len(frames)) - 3} consecutive frames of the bench sequence through the cv2 {cv2.__version__} "
                          f"restatement of TrackImage (same OpenCV routines the reference calls, numpy-vectorised glue); {cores} host "
                          f"cores; cv2_calls_only counts only the time inside the OpenCV routines; c_oracle_1_thread = "
                          f"oracle/vf_oracle.c on 37 frames"}
    except Exception as e:  # cv2 missing: the plain-C restatement is the baseline
        return {"value": c_fps, "unit": "frames/s", "cores": 1, "kind": "port", "n_streams": 1,
                "sample": f"37 frames, plain-C restatement oracle/vf_oracle.c (cv2 unavailable: {e})"}


# ======================================================================================================================
#  parity of the bench's own sequence (outside every timed region)
# ======================================================================================================================
def parity_block(vf, cfg, frames, n_check: int):
    """Frames 0..n_check-1 of the bench sequence through the synchronous C-ABI call and through the CPU oracle
    (oracle/vf_oracle.c, itself pinned bit for bit to the cv2 restatement of the reference): the gates of BASELINE.md 3."""
    from oracle import c_oracle
    from oracle.camera import EUROC_CAM0, EUROC_CAM1
    n_check = min(n_check, len(frames))
    gpu = vf.FeatureTracker(cfg)
    ref = c_oracle.CFeatureTracker(EUROC_CAM0, EUROC_CAM1, W, H, MAX_CNT, MIN_DIST, FLOW_BACK, LEVEL)
    keys = ("ids", "track_cnt", "uv", "un", "vel", "ids_right", "uv_right", "un_right", "vel_right")
    first_div, max_px, max_un, agree, total, feats = None, 0.0, 0.0, 0, 0, 0
    for k in range(n_check):
        g = c_oracle.records_to_frame(gpu.track(k / 20.0, frames[k, 0], frames[k, 1]))
        o = ref.track_image(k / 20.0, frames[k, 0], frames[k, 1])
        same = all(np.asarray(g[key]).shape == np.asarray(o[key]).shape and np.asarray(g[key]).tobytes() == np.asarray(o[key]).tobytes()
                   for key in keys)
        if not same and first_div is None:
            first_div = k
        # tolerance view (north star): features matched by id; status = kept / has-right decisions
        gi = {int(i): n for n, i in enumerate(g["ids"])}
        oi = {int(i): n for n, i in enumerate(o["ids"])}
        common = sorted(set(gi) & set(oi))
        total += len(set(gi) | set(oi))
        gr, orr = set(int(i) for i in g["ids_right"]), set(int(i) for i in o["ids_right"])
        agree += sum(1 for i in common if (i in gr) == (i in orr))
        feats += len(oi)
        if common:
            a, b = np.array([gi[i] for i in common]), np.array([oi[i] for i in common])
            max_px = max(max_px, float(np.abs(g["uv"][a].astype(np.float64) - o["uv"][b]).max()))
            max_un = max(max_un, float(np.abs(g["un"][a].astype(np.float64) - o["un"][b]).max()))
    del gpu
    return {"frames_checked": n_check, "features_checked": feats, "bit_exact": first_div is None, "first_divergence_frame": first_div,
            "max_px_err": max_px, "max_undistorted_err": max_un, "status_agreement": agree / max(total, 1),
            "tolerances": {"px": 0.01, "undistorted": 1e-6, "status_agreement": 0.995},
            "against": "oracle/vf_oracle.c (plain-C restatement, bit-identical to the cv2 4.13 restatement of the reference on "
                       "tests/golden/*); frames 0.. of this bench's sequence, synchronous vf_tracker_track, outside the timed regions"}


# ======================================================================================================================
#  our arm
# ======================================================================================================================
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--repeats", type=int, default=5, help="timed regions of exactly --steps steps each; the median is reported")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--multi-streams", type=int, default=64, help="streams per GPU for the extra multi-stream line (0 = skip)")
    ap.add_argument("--legs", default=ALL_LEGS, help="comma list of the legs to run (profiling runs pick one)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-affinity", default="auto", choices=["auto", "off"],
                    help="auto: run this rank on the CPUs NVML names as local to its GPU (vins_fast_b200/affinity.py)")
    ap.add_argument("--sequence-seeds", default="same", choices=["same", "rank"],
                    help="same: every rank runs seed 0 (equal work per GPU); rank: one synthetic sequence per rank")
    ap.add_argument("--dist-backend", default="nccl", choices=["nccl", "gloo"], help="barrier / max-over-ranks plumbing (A/B runs)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    args.repeats = max(1, args.repeats)
    legs = set(args.legs.split(","))
    if args.no_cpu_baseline:
        legs.discard("cpu")
    if args.multi_streams <= 0:
        legs.discard("multi")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    affinity = {"bound": False, "why": "--cpu-affinity off"}
    if args.cpu_affinity == "auto":          # before CUDA comes up, so that the driver's own threads inherit it
        from vins_fast_b200.affinity import bind_to_gpu
        affinity = bind_to_gpu(local)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the tracker has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if args.dist_backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")

    import vins_fast_b200 as vf
    from vins_fast_b200 import build as vf_build
    from vins_fast_b200.shard import streams_for_rank
    vf_build.build()
    dev = torch.device("cuda", local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda" if args.dist_backend == "nccl" else "cpu", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def regions(stream, n_regions, body):
        """n_regions timed regions; body(r) issues exactly the steps of region r and returns when their results are on the
        host.  Device time (CUDA events on the stream the chain is launched on) and host wall clock, max over ranks."""
        dev_ms, wall_ms = [], []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for r in range(n_regions):
            barrier()
            e0.record(stream)
            t0 = time.perf_counter()
            body(r)
            wall = (time.perf_counter() - t0) * 1e3
            e1.record(stream)
            barrier()
            dev_ms.append(max_over_ranks(e0.elapsed_time(e1)))
            wall_ms.append(max_over_ranks(wall))
        return dev_ms, wall_ms

    def isolated_stage_times(cfg_, n_streams, feed, n_frames_, skip):
        """Per-kernel device times with every kernel ALONE on the GPU: a batch created with VF_SERIALIZE=1 puts all stages of a
        frame on one stream, profiling mode brackets every kernel with CUDA events on that stream (plain launches, no graph).
        What the frame plans overlap (pyramid beside response, stereo beside undistortion) would otherwise stretch the
        event-to-event spans; these are the launch durations the roofline fractions are computed from."""
        os.environ["VF_SERIALIZE"] = "1"
        try:
            sb = vf.Batch(cfg_, n_streams)
        finally:
            del os.environ["VF_SERIALIZE"]
        sb.set_profiling(True)
        acc_, cnt_ = {}, 0
        for k in range(n_frames_):
            feed(sb, k)
            if k >= skip:
                for nm, ms in sb.stage_times_ms().items():
                    acc_[nm] = acc_.get(nm, 0.0) + ms * 1e3
                cnt_ += 1
        sb.set_profiling(False)
        sb.close()
        return {k: v / max(cnt_, 1) for k, v in acc_.items()}

    K, Wm, R = args.steps, args.warmup, args.repeats
    cam0, cam1 = workload_cameras("c1")
    # D distinct frames (> L2 in total), visited forward then backward (0..D-1, D-2..0, ...): consecutive steps always
    # see neighbouring frames of the sequence, and every frame is re-read only after all the others went through L2.
    D = N_DISTINCT
    # Every rank tracks its own stream; by default all of them are fed the SAME synthetic sequence (seed 0), so that the work per
    # GPU is equal at every N -- what weak scaling is about.  The cost of a frame depends on the sequence (seeds 0..7 of this
    # workload: 67.5 .. 73.7 us per overlapped frame on one GPU, per-seed runs in DESIGN.md 7), and with one seed per rank the
    # MAX over ranks reports that spread as a scaling loss (round-2 lines before this change: 0.89 at 8 GPUs).
    # --sequence-seeds rank restores one seed per rank; config 4 (c4_sharded) always runs 64 distinct sequences.
    seq_seed = int(os.environ["VF_BENCH_SEED"]) if "VF_BENCH_SEED" in os.environ else (rank if args.sequence_seeds == "rank" else 0)
    frames = make_frames(seq_seed, D)
    frame_bytes = 2 * H * W
    d_frames = torch.from_numpy(frames).cuda()                   # inputs resident in HBM
    p_frames = torch.from_numpy(frames).pin_memory()             # pinned host copies for the e2e legs
    n_frames = Wm + R * K
    times = np.arange(n_frames + 64) / 20.0
    fidx = [ping_pong(k, D) for k in range(n_frames + 64)]
    cfg = vf.make_config(W, H, MAX_CNT, MIN_DIST, FLOW_BACK, cam0, cam1, lk_max_level=LEVEL, device=local)
    t_create = time.perf_counter()
    batch = vf.Batch(cfg, 1)
    create_ms = (time.perf_counter() - t_create) * 1e3           # includes capture + instantiation of every graph and the blank warm frames
    # the library owns its streams (the frame-to-frame chain runs on a high-priority one); the timing events are recorded
    # on that stream, the one the kernels are launched on
    stream = torch.cuda.ExternalStream(batch.cuda_stream, device=dev)
    base = d_frames.data_ptr()
    hp = p_frames.data_ptr()
    ptr_l = [(C.c_void_p * 1)(hp + fidx[k] * frame_bytes) for k in range(n_frames)]
    ptr_r = [(C.c_void_p * 1)(hp + fidx[k] * frame_bytes + H * W) for k in range(n_frames)]
    line = {}
    n_last = None

    with torch.cuda.stream(stream):
        # ---------------- value: device-resident inputs, frames enqueued asynchronously ----------------
        def run_device(first: int, count: int):
            for k in range(first, first + count):
                o = base + fidx[k] * frame_bytes
                batch.enqueue_device(times[k], o, o + H * W, W, frame_bytes)
            return batch.fetch_counts()

        first_frames_us = []                                     # the very first frames of a fresh handle: no warm-up cliff
        for k in range(Wm):
            t0 = time.perf_counter()
            run_device(k, 1)
            first_frames_us.append((time.perf_counter() - t0) * 1e6)
        host_issue = {}
        with ClockSampler(local) as clk:
            h0 = batch.host_times_us()
            v_dev, v_wall = regions(stream, R, lambda r: run_device(Wm + r * K, K))
            h1 = batch.host_times_us()
        n_last = int(batch.fetch_counts()[0])
        host_issue["value"] = max_over_ranks((h1[1] - h0[1]) / (R * K))
        ms_value = med(v_dev)
        launches = batch.launches_per_frame * K

        # ---------------- e2e: pinned host images in, records out, host-synchronous per frame ----------------
        e_dev = e_wall = None
        clk2 = None
        if "e2e" in legs:
            batch.reset()
            for k in range(Wm):
                batch.track_ptrs(times[k], ptr_l[k], W, ptr_r[k], W)

            def run_sync(r):
                for k in range(Wm + r * K, Wm + (r + 1) * K):
                    batch.track_ptrs(times[k], ptr_l[k], W, ptr_r[k], W)
            with ClockSampler(local) as clk2:
                h0 = batch.host_times_us()
                e_dev, e_wall = regions(stream, R, run_sync)
                h1 = batch.host_times_us()
            host_issue["e2e"] = max_over_ranks((h1[1] - h0[1]) / (R * K))
            n_sync = int(batch.fetch_counts()[0])
            assert n_sync == n_last, "device-resident and host paths disagree on the feature count"
        e_ms = [max(a, b) for a, b in zip(e_dev, e_wall)] if e_dev else None

        # ---------------- e2e, pipelined: enqueue frame k from pinned host images, collect frame k-lag meanwhile ----------------
        pipe = {}
        if "pipelined" in legs:
            for lag in (1, 2):
                batch.reset()
                for k in range(Wm):
                    batch.enqueue_ptrs(times[k], ptr_l[k], W, ptr_r[k], W)
                    if k >= lag:
                        batch.fetch_counts_lagged(lag)

                def run_pipe(r, lag=lag):
                    for k in range(Wm + r * K, Wm + (r + 1) * K):
                        batch.enqueue_ptrs(times[k], ptr_l[k], W, ptr_r[k], W)
                        batch.fetch_counts_lagged(lag)            # every frame's records reach the host inside the timed region
                h0 = batch.host_times_us()
                _, p_wall = regions(stream, R, run_pipe)
                h1 = batch.host_times_us()
                for back in range(lag - 1, -1, -1):
                    n = batch.fetch_counts_lagged(back)
                assert int(n[0]) == n_last, "pipelined and synchronous host paths disagree on the feature count"
                pipe[lag] = (med(p_wall), p_wall)
                host_issue[f"pipelined_lag{lag}"] = max_over_ranks((h1[1] - h0[1]) / (R * K))

        # ---------------- per-kernel times (CUDA events on the launching streams) ----------------
        stages_us, n_prof, iters_t, iters_s, feats, fast_it = {}, 0, 0, 0, 0, 0
        stages_in_frame_us = {}
        if "stages" in legs:
            from vins_fast_b200 import _lib as L
            batch.reset()
            batch.set_profiling(True)
            acc = {}
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
            stages_in_frame_us = {k: v / max(n_prof, 1) * 1e3 for k, v in acc.items()}   # event-to-event, stages overlapping as in a frame
            stages_us = isolated_stage_times(cfg, 1, lambda sb, k: sb.track_ptrs(times[k], ptr_l[k], W, ptr_r[k], W), min(n_frames, Wm + 40), Wm)

        # ---------------- kernel timeline of a synchronous frame on the GPU's own nanosecond timer ----------------
        # (every kernel stamps its first start / last end when tracing is on: no event overhead, true launch gaps)
        timeline = None
        if "timeline" in legs:
            try:
                os.environ["VF_TRACE"] = "1"
                tb = vf.Batch(vf.make_config(W, H, MAX_CNT, MIN_DIST, FLOW_BACK, cam0, cam1, lk_max_level=LEVEL, device=local), 1)
                del os.environ["VF_TRACE"]
                tnames = ("clear_frame pyrdown_left pad_left lk_temporal select_tracked response_nms bucket_scan bucket_scatter "
                          "mask_and_max gftt_select pyrdown_right pad_right lk_stereo undistort_left").split()
                tbuf = (C.c_ulonglong * 256)()
                tacc = []
                vf.load().vf_batch_trace_get(tb._h, tbuf, 256)        # drop the stamps of the blank warm-up frames
                for k in range(min(n_frames, Wm + 40)):
                    tb.track_ptrs(times[k], ptr_l[k], W, ptr_r[k], W)
                    vf.load().vf_batch_trace_get(tb._h, tbuf, 256)
                    if k >= Wm:
                        a = np.array(list(tbuf), np.uint64)[:192].reshape(6, 16, 2)[k % 6, :14]
                        unused = a[:, 1] == 0                                # kernels this plan does not launch
                        a = a.astype(np.int64) - int(a[~unused, 0].min())
                        a[unused] = 0
                        tacc.append(a)
                tm = np.mean(tacc, 0) / 1e3
                timeline = {"kernels": {n: {"start_us": round(float(tm[i, 0]), 2), "end_us": round(float(tm[i, 1]), 2),
                                            "us": round(float(tm[i, 1] - tm[i, 0]), 2)} for i, n in enumerate(tnames)},
                            "frame_span_us": round(float(tm[:, 1].max()), 2),
                            "note": f"synchronous frames, mean of {len(tacc)}; t = 0 at the first kernel start; the uploads precede it"}
                tb.close()
            except Exception as e:  # tracing is a measurement aid; the bench line does not depend on it
                os.environ.pop("VF_TRACE", None)
                timeline = {"error": str(e)}

        # ---------------- extra: many independent streams on one GPU, all reading neighbouring frames of the rank's sequence ----------------
        multi = None
        if "multi" in legs:
            S = args.multi_streams
            mb = vf.Batch(vf.make_config(W, H, MAX_CNT, MIN_DIST, FLOW_BACK, cam0, cam1, lk_max_level=LEVEL, device=local), S)
            mstream = torch.cuda.ExternalStream(mb.cuda_stream, device=dev)
            mt = lambda k: times[k] + np.zeros(S)
            steps_m, warm_m = 30, 5

            def run_multi(first, count):      # stream s sees frames k+s, k+s+1, ... : consecutive frames of the same sequence
                for k in range(first, first + count):
                    mb.enqueue_device(mt(k), base + k * frame_bytes, base + k * frame_bytes + H * W, W, frame_bytes)
                mb.fetch_counts()
            run_multi(0, warm_m)
            m_dev, _ = regions(mstream, 3, lambda r: run_multi(warm_m + r * steps_m, steps_m))
            ms_m = med(m_dev)
            multi = {"streams_per_gpu": S, "steps": steps_m, "value": world * S * steps_m / (ms_m * 1e-3), "unit": "frames/s",
                     "us_per_frame_per_stream": ms_m * 1e3 / (steps_m * S), "ms_per_step": ms_m / steps_m,
                     "regions_ms": m_dev, "launches_per_step": mb.launches_per_frame}
            # per-kernel device times of the batched step (CUDA events around every kernel, plain launches, synchronous frames)
            mb.close()

            def feed_multi(sb, k):
                sb.enqueue_device(mt(k), base + k * frame_bytes, base + k * frame_bytes + H * W, W, frame_bytes)
                sb.fetch_counts()
            multi["stages_us"] = isolated_stage_times(vf.make_config(W, H, MAX_CNT, MIN_DIST, FLOW_BACK, cam0, cam1, lk_max_level=LEVEL, device=local),
                                                      S, feed_multi, 12, 4)

        # ---------------- BASELINE configs 3, 4 and 5 as measured workloads ----------------
        workloads = {}
        peak, peak_src = peaks()

        def other_workload(name, n_distinct, steps, seed):
            wl = WORKLOADS[name]
            w, h, lvl = wl["width"], wl["height"], wl["lk_max_level"]
            c0, c1 = workload_cameras(name)
            fr = make_frames(seed, n_distinct, w, h)
            dfr = torch.from_numpy(fr).cuda()
            pfr = torch.from_numpy(fr).pin_memory()
            fb = 2 * h * w
            wcfg = vf.make_config(w, h, wl["max_cnt"], wl["min_dist"], wl["flow_back"], c0, c1, lk_max_level=lvl, device=local)
            wb = vf.Batch(wcfg, 1)
            wst = torch.cuda.ExternalStream(wb.cuda_stream, device=dev)
            warm = 5
            n_tot = warm + 3 * steps
            idx = [ping_pong(k, n_distinct) for k in range(n_tot)]
            with torch.cuda.stream(wst):
                def dev_steps(first, count):
                    for k in range(first, first + count):
                        o = dfr.data_ptr() + idx[k] * fb
                        wb.enqueue_device(k / 20.0, o, o + h * w, w, fb)
                    return wb.fetch_counts()
                dev_steps(0, warm)
                d_ms, _ = regions(wst, 3, lambda r: dev_steps(warm + r * steps, steps))
                nf = int(wb.fetch_counts()[0])
                wb.reset()
                pl = [(C.c_void_p * 1)(pfr.data_ptr() + idx[k] * fb) for k in range(n_tot)]
                pr = [(C.c_void_p * 1)(pfr.data_ptr() + idx[k] * fb + h * w) for k in range(n_tot)]
                for k in range(warm):
                    wb.track_ptrs(k / 20.0, pl[k], w, pr[k], w)

                def sync_steps(r):
                    for k in range(warm + r * steps, warm + (r + 1) * steps):
                        wb.track_ptrs(k / 20.0, pl[k], w, pr[k], w)
                s_dev, s_wall = regions(wst, 3, sync_steps)
                s_ms = [max(a, b) for a, b in zip(s_dev, s_wall)]
            st_us = isolated_stage_times(wcfg, 1, lambda sb, k: sb.track_ptrs(k / 20.0, pl[k], w, pr[k], w), min(n_tot, warm + 8), warm)
            npx = w * h
            s_l = sum(4.0 ** -l for l in range(lvl + 1))
            # SURVEY 8(d) compulsory bytes: pyramid read N + write N (S_L - 1); response read N + write 4 N response + N flags;
            # mask stage read 4 N response + N flags, write N mask
            alg = {"pyrdown_left": npx * s_l, "pyrdown_right": npx * s_l, "response_nms": 6 * npx, "mask_and_max": 6 * npx}
            streaming = {k: {"us": st_us[k], "alg_bytes": int(v), "GBps": v / (st_us[k] * 1e-6) / 1e9,
                             "frac_of_hbm_peak": v / (st_us[k] * 1e-6) / 1e9 / peak} for k, v in alg.items() if st_us.get(k, 0) > 0}
            wb.close()
            del dfr, pfr
            dm, sm = med(d_ms), med(s_ms)
            return {"workload": f"{w}x{h} stereo, max_cnt {wl['max_cnt']}, min_dist {wl['min_dist']}, maxLevel {lvl}, pinhole radtan, 1 stream per GPU",
                    "steps": steps, "regions": 3, "distinct_frames": n_distinct,
                    "inputs_vs_l2": f"{n_distinct * fb / 2**20:.0f} MiB of distinct frames vs 126 MiB L2",
                    "value": world * steps / (dm * 1e-3), "unit": "frames/s", "us_per_frame": dm / steps * 1e3,
                    "e2e": {"value": world * steps / (sm * 1e-3), "us_per_frame": sm / steps * 1e3, "h2d_bytes_per_step": fb + 8,
                            "d2h_bytes_per_step": wl["max_cnt"] * 60 + 4},
                    "features_last_frame": nf, "stages_us": st_us, "streaming_kernels": streaming}

        def window_workload(win, steps):
            """The headline configuration with another LK window width (`lk_win`; csrc/vf_lk_generic.cu): same frames, same
            two legs (device-resident overlapped frames; host-synchronous call from pinned images)."""
            wcfg = vf.make_config(W, H, MAX_CNT, MIN_DIST, FLOW_BACK, cam0, cam1, lk_max_level=LEVEL, device=local)
            wcfg.lk_win = win
            wb = vf.Batch(wcfg, 1)
            wst = torch.cuda.ExternalStream(wb.cuda_stream, device=dev)
            warm = 5
            with torch.cuda.stream(wst):
                def dev_steps(first, count):
                    for k in range(first, first + count):
                        o = base + fidx[k] * frame_bytes
                        wb.enqueue_device(times[k], o, o + H * W, W, frame_bytes)
                    return wb.fetch_counts()
                dev_steps(0, warm)
                d_ms, _ = regions(wst, 3, lambda r: dev_steps(warm + r * steps, steps))
                nf = int(wb.fetch_counts()[0])
                wb.reset()
                for k in range(warm):
                    wb.track_ptrs(times[k], ptr_l[k], W, ptr_r[k], W)

                def sync_steps(r):
                    for k in range(warm + r * steps, warm + (r + 1) * steps):
                        wb.track_ptrs(times[k], ptr_l[k], W, ptr_r[k], W)
                s_dev, s_wall = regions(wst, 3, sync_steps)
            wb.close()
            st_us = isolated_stage_times(wcfg, 1, lambda sb, k: sb.track_ptrs(times[k], ptr_l[k], W, ptr_r[k], W), warm + 8, warm)
            dm, sm = med(d_ms), med([max(a, b) for a, b in zip(s_dev, s_wall)])
            return {"lk_win": win, "lk_kernels_us": {k: st_us.get(k) for k in ("lk_temporal", "lk_stereo")}, "steps": steps, "regions": 3, "value": world * steps / (dm * 1e-3), "unit": "frames/s",
                    "us_per_frame": dm / steps * 1e3, "e2e": {"value": world * steps / (sm * 1e-3), "us_per_frame": sm / steps * 1e3},
                    "features_last_frame": nf}

        if "win" in legs:
            steps_w = max(1, min(K, (n_frames - 5) // 3, 100))
            workloads["lk_windows"] = {"note": "headline configuration with lk_win 15 / 31 instead of the reference's 21 (run-time-width LK kernels)",
                                       "w15": window_workload(15, steps_w), "w31": window_workload(31, steps_w)}
        if "c3" in legs:
            workloads["c3"] = other_workload("c3", 72, min(K, 100), 1000 + (rank if args.sequence_seeds == "rank" else 0))
        if "c5" in legs:
            workloads["c5"] = other_workload("c5", 8, min(K, 40), 2000 + (rank if args.sequence_seeds == "rank" else 0))
        if "c4" in legs:
            # BASELINE config 4 as specified: 64 independent sequences (seed = global stream index) partitioned over the ranks
            # (stream s -> rank s mod world), the rank's share batched in ONE vf_batch; aggregate frames/s over all ranks.
            S_TOTAL, DC = 64, 12
            mine = streams_for_rank(S_TOTAL, rank, world)
            Sl = len(mine)
            imgs = make_stream_frames(mine, DC)
            d4 = torch.from_numpy(imgs).cuda()
            p4 = torch.from_numpy(imgs).pin_memory()
            cb = vf.Batch(vf.make_config(W, H, MAX_CNT, MIN_DIST, FLOW_BACK, cam0, cam1, lk_max_level=LEVEL, device=local), Sl)
            cst = torch.cuda.ExternalStream(cb.cuda_stream, device=dev)
            steps4, warm4 = min(K, 30), 5
            idx4 = [ping_pong(k, DC) for k in range(warm4 + 3 * steps4)]
            step_bytes = 2 * Sl * H * W
            with torch.cuda.stream(cst):
                def dev4(first, count):
                    for k in range(first, first + count):
                        o = d4.data_ptr() + idx4[k] * step_bytes
                        cb.enqueue_device(k / 20.0, o, o + Sl * H * W, W, H * W)
                    return cb.fetch_counts()
                dev4(0, warm4)
                c_dev, _ = regions(cst, 3, lambda r: dev4(warm4 + r * steps4, steps4))
                cb.reset()
                pl4 = [(C.c_void_p * Sl)(*[p4.data_ptr() + idx4[k] * step_bytes + j * H * W for j in range(Sl)]) for k in range(len(idx4))]
                pr4 = [(C.c_void_p * Sl)(*[p4.data_ptr() + idx4[k] * step_bytes + (Sl + j) * H * W for j in range(Sl)]) for k in range(len(idx4))]
                for k in range(warm4):
                    cb.track_ptrs(k / 20.0, pl4[k], W, pr4[k], W)

                def sync4(r):
                    for k in range(warm4 + r * steps4, warm4 + (r + 1) * steps4):
                        cb.track_ptrs(k / 20.0, pl4[k], W, pr4[k], W)
                s4_dev, s4_wall = regions(cst, 3, sync4)
            cm, sm4 = med(c_dev), med([max(a, b) for a, b in zip(s4_dev, s4_wall)])
            workloads["c4_sharded"] = {
                "workload": "64 independent 752x480 EuRoC-shape sequences (seed = stream index) partitioned over the GPUs "
                            "(stream s -> rank s mod N, vins_fast_b200.shard.streams_for_rank), each rank's share batched in one vf_batch",
                "scaling": "strong", "streams_total": S_TOTAL, "streams_per_gpu": Sl, "steps": steps4, "regions": 3,
                "inputs_vs_l2": f"{DC} distinct steps of {step_bytes / 2**20:.0f} MiB per GPU",
                "value": S_TOTAL * steps4 / (cm * 1e-3), "unit": "frames/s", "ms_per_step": cm / steps4,
                "us_per_frame_per_stream": cm * 1e3 / (steps4 * Sl),
                "e2e": {"value": S_TOTAL * steps4 / (sm4 * 1e-3), "ms_per_step": sm4 / steps4, "h2d_bytes_per_step": step_bytes + 8 * Sl,
                        "d2h_bytes_per_step": Sl * (MAX_CNT * 60 + 4), "mode": "host-synchronous vf_batch_track from pinned host images"},
                "launches_per_step": cb.launches_per_frame}
            cb.close()
            del d4, p4

    value = world * K / (ms_value * 1e-3)
    clocks = clk.summary()
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(clocks["reasons"])
    sm_hz = (clocks["sm_mhz"] or 1965.0) * 1e6

    # ---------------- roofline of the dominant kernel ----------------
    kernels = {k: v for k, v in stages_us.items() if not k.startswith(("h2d", "d2h"))}
    dom = max(kernels, key=kernels.get) if kernels else "lk_stereo"
    passes_levels = (LEVEL + 1) * 2                      # stereo: two LK calls over levels 0..3
    iters = iters_s
    if dom == "lk_temporal":
        passes_levels, iters = (LEVEL + 1) + 2, iters_t  # forward L3 + reverse L1
    # algorithmic bytes of an LK launch: per point and level one 24x24 uint8 tile of I, per iteration one
    # 22x22 uint8 window of J; the derivative planes are never materialised.
    lk_bytes = (feats / max(n_prof, 1)) * passes_levels * 24 * 24 + (iters / max(n_prof, 1)) * 22 * 22
    npx = W * H
    alg_bytes = {"lk_stereo": lk_bytes, "lk_temporal": lk_bytes,
                 "pyrdown_left": npx * 1.328125, "pyrdown_right": npx * 1.328125,      # read N_px, write N_px (S_3 - 1)
                 "response_nms": npx * 6, "mask_and_max": npx * 6}
    dur_us = kernels.get(dom, 0.0)
    ach = (alg_bytes.get(dom, 0.0) / (dur_us * 1e-6) / 1e9) if dur_us > 0 else 0.0
    ncu = {}
    for fn in ("r2_lk_issue.json", "roofline_traffic.json"):
        p = os.path.join(ROOT, "profiles", fn)
        if os.path.exists(p):
            try:
                ncu.update(json.load(open(p)))
            except Exception:
                pass
    issue_peak = N_SM * N_SCHED * sm_hz / 1e9            # G warp-instructions / s
    rec = ncu.get(dom + "_S1", {}) if isinstance(ncu.get(dom + "_S1"), dict) else {}
    roofline = {"kernel": dom, "kernel_us": dur_us, "peak_source": peak_src}
    if dom.startswith("lk") and rec.get("warp_instructions"):
        # LK is a dependent chain (levels x Gauss-Newton iterations), not a streaming kernel: its roofline is the warp
        # instruction issue rate.  warp_instructions per launch comes from the committed ncu capture of this command
        # (profiles/r2_lk_issue.json), the launch duration is measured live above.
        a_issue = rec["warp_instructions"] / (dur_us * 1e-6) / 1e9 if dur_us > 0 else 0.0
        roofline.update({"bound": "issue", "achieved": a_issue, "peak": issue_peak, "unit": "Gwarp-inst/s", "frac": a_issue / issue_peak,
                         "traffic": rec.get("dram_bytes"), "issue_active_pct": rec.get("issue_active_pct"),
                         "warps_active_pct": rec.get("warps_active_pct"), "top_stalls": rec.get("top_stalls"),
                         "warp_instructions_per_launch": rec["warp_instructions"],
                         "hbm": {"achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None}})
    else:
        roofline.update({"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None,
                         "traffic": (ncu.get(dom) if not isinstance(ncu.get(dom), dict) else ncu[dom].get("dram_bytes"))})
    roofline["note"] = ("single stream: 150 features = 150 four-warp CTAs on 148 SMs, the frame's working set (~2 MB) is L2-resident; "
                        "the kernel is latency-bound (one dependent chain per feature), so neither fraction is near 1 by design -- "
                        "the batched launch (64 streams, multi-wave) is the one that fills the machine, see `batched`")
    roofline["streaming_kernels"] = {k: {"achieved_GBps": alg_bytes[k] / (kernels[k] * 1e-6) / 1e9,
                                         "frac": alg_bytes[k] / (kernels[k] * 1e-6) / 1e9 / peak}
                                     for k in ("pyrdown_left", "response_nms", "mask_and_max") if kernels.get(k, 0) > 0}
    if isinstance(ncu.get("lk_stereo_S64"), dict) and multi and multi.get("stages_us", {}).get("lk_stereo", 0) > 0 and multi["streams_per_gpu"] == 64:
        # the launch that fills the machine: 64 streams x 150 features, one warp per feature (vf_lk_warp.cu).  Warp instructions
        # per launch from the committed ncu capture of exactly this launch, duration measured live (events, profiling mode).
        rb = ncu["lk_stereo_S64"]
        us = multi["stages_us"]["lk_stereo"]
        a_b = rb["warp_instructions"] / (us * 1e-6) / 1e9
        roofline["batched"] = {"kernel": "lk_stereo_warp_kernel (64 streams)", "bound": "issue", "kernel_us": us, "achieved": a_b, "peak": issue_peak,
                               "unit": "Gwarp-inst/s", "frac": a_b / issue_peak, "traffic": rb.get("dram_bytes"),
                               "warp_instructions_per_launch": rb["warp_instructions"], "issue_active_pct_ncu": rb.get("issue_active_pct"),
                               "warps_active_pct_ncu": rb.get("warps_active_pct"), "top_stalls_ncu": rb.get("top_stalls"),
                               "registers_per_thread": rb.get("registers_per_thread"),
                               "note": "profiles/r2_lk_issue.json + profiles/r2_ncu_lk_batch_summary.txt"}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_value / K, "us_per_frame": ms_value / K * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/i32 stencils, f32 LK + response, f64 undistort", "data": "synthetic",
        "config": workload_config(),
        "run": {"repeats": R, "value_regions_ms": v_dev, "value_min_us_per_frame": min(v_dev) / K * 1e3,
                "value_max_us_per_frame": max(v_dev) / K * 1e3, "features_last_frame": n_last, "cuda_graph": bool(cfg.use_cuda_graph),
                "cpu_affinity": affinity, "dist_backend": args.dist_backend if world > 1 else None,
                "create_ms": create_ms, "first_frames_us": [round(v, 1) for v in first_frames_us[:8]],
                "value_mode": "frames enqueued asynchronously (vf_batch_enqueue_device), records of every frame downloaded; the "
                              "stereo pass of frame k overlaps the temporal chain of frame k+1; each region = exactly K steps timed "
                              "with CUDA events on the stream the chain is launched on, median of the regions reported"},
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                   "samples": clocks["samples"], "rejected": bool(bad)},
        "gpu_launches": launches, "launches_per_frame": batch.launches_per_frame, "host_issue_us_per_frame": host_issue,
        "stages_us": stages_us, "stages_note": "per-kernel launch durations, every kernel alone on the GPU (VF_SERIALIZE batch, CUDA events around each plain launch: includes "
                                               "~5 us of event / launch latency per stage, the timeline has the bare kernel spans); "
                                               "stages_in_frame_us = the same events with the frame's stages overlapping as they do in a synchronous frame",
        "stages_in_frame_us": stages_in_frame_us, "timeline": timeline, "roofline": roofline, "multi_stream": multi, "workloads": workloads,
        "lk": {"iterations_per_frame": (iters_t + iters_s) / max(n_prof, 1),
               "exact_integer_path_fraction": fast_it / max(iters_t + iters_s, 1)},
    }
    if e_ms:
        ms_e2e = med(e_ms)
        line["e2e"] = {"value": world * K / (ms_e2e * 1e-3), "unit": "frames/s", "us_per_frame": ms_e2e / K * 1e3,
                       "h2d_bytes_per_step": frame_bytes + 8, "d2h_bytes_per_step": MAX_CNT * 60 + 4, "clocks": clk2.summary(),
                       "regions_ms": e_ms, "min_us_per_frame": min(e_ms) / K * 1e3,
                       "mode": "host-synchronous vf_batch_track per frame (what the reference's TrackImage does); region time = "
                               "max(CUDA events, host wall clock)"}
        for lag, key in ((1, "pipelined"), (2, "pipelined_lag2")):
            if lag in pipe:
                m, allp = pipe[lag]
                line["e2e"][key] = {"value": world * K / (m * 1e-3), "unit": "frames/s", "us_per_frame": m / K * 1e3, "regions_ms": allp,
                                    "mode": f"vf_batch_enqueue(k) + vf_batch_fetch_lagged({lag}): frame k-{lag} is collected while frame k "
                                            "runs; same H2D / D2H bytes per frame, host wall clock"}
    if rank == 0 and "parity" in legs:
        line["parity"] = parity_block(vf, cfg, frames, 120)
    if rank == 0 and world == 1 and "cpu" in legs:
        line["cpu_baseline"] = cpu_baseline(frames, 160)
    batch.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
