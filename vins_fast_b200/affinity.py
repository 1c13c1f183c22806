"""Bind the calling process to the CPUs next to its GPU.

A frame is a chain of short latency-bound kernels: every launch is a doorbell write from the host thread to the GPU and a read
of the command buffer in host memory by the GPU.  On a two-socket box a rank that the scheduler happens to run on the socket
the GPU is NOT attached to pays the inter-socket hop on both, for every launch of every frame.  One process per GPU (bench.py
under torchrun, a deployment with one tracker process per camera rig) should therefore run on the GPU's own NUMA node: NVML
knows the ideal CPU set of a device (nvmlDeviceGetCpuAffinity).
"""
from __future__ import annotations

import os


def _nvml_index(local_index: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        parts = vis.split(",")
        if local_index < len(parts) and parts[local_index].strip().isdigit():
            return int(parts[local_index])
    return local_index


def gpu_cpu_set(local_index: int):
    """CPUs NVML names as local to the GPU (empty set if NVML or the query is unavailable)."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(_nvml_index(local_index))
        n_cpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        return {c for c in cpus if c < n_cpu}
    except Exception:
        return set()


def bind_to_gpu(local_index: int) -> dict:
    """Restrict the process to (its current CPU set) & (the GPU's local CPUs).  Returns what was done, for the bench line."""
    try:
        allowed = set(os.sched_getaffinity(0))
    except AttributeError:
        return {"bound": False, "why": "sched_getaffinity unavailable"}
    local = gpu_cpu_set(local_index)
    target = allowed & local
    if not local:
        return {"bound": False, "why": "NVML gave no CPU set", "cpus_allowed": len(allowed)}
    if not target:
        return {"bound": False, "why": "the GPU's CPUs are outside this process's CPU set", "cpus_allowed": len(allowed)}
    if target == allowed:
        return {"bound": False, "why": "already on the GPU's CPUs", "cpus_allowed": len(allowed), "cpus_gpu_local": len(local)}
    os.sched_setaffinity(0, target)
    return {"bound": True, "cpus_allowed": len(allowed), "cpus_gpu_local": len(local), "cpus_now": len(target)}
