#!/usr/bin/env python3
"""bench.py -- Mrays/s of the compressed-BVH traversal hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--leaf leaf|grid|box] [--configs all|none|1,3,4,5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[1]): build/bomberman.obj as one subdivision mesh, S=6 C=3 compressed leaves,
camera of build/bomberman.ecs, 1920x1080 primary rays + one cosine-diffuse bounce ray per primary hit.  A step = one pass
of the hot path over that batch: trace the primary rays, trace the bounce rays.  Both arms (ours / --impl reference) trace
the SAME two ray arrays: the bounce rays are derived from per-pixel hit distances and shading normals computed once at
build() time by the compiled reference (scenes/_built/bench_c2_primary_hits.npz).

  value      device-timed: rays resident in HBM, CUDA events around the K steps on the launching stream.
             --gpus N > 1: ONE frame sharded over the N ranks by interleaved 64x64 tiles ("scaling": "strong"), scene replicated.
  value_with_gather   the same with the hit records of every rank landing in rank 0's frame buffer inside the timed region:
             the traversal kernels store them straight into rank 0's arrays (IPC-mapped peer memory, NVLink stores) -- traversal
             and gather are one kernel, 4 bytes per ray + 16 per hit cross the link while the kernel runs.
  e2e        the same step through the reference-facing call (ecbIntersect1M == rtcIntersect1M) with pinned HOST ray
             buffers, host<->device copies inside the timed region; at N > 1 ONE call in ONE process fans out over N GPUs.
  roofline   algorithmic bytes of one launch (instrumented kernel variant, SURVEY 8d formula with this repo's node / header
             sizes) / average launch duration against MEASURED_PEAKS.json's HBM copy bandwidth -- plus what ncu says
             actually bounds the kernel (copied from the committed captures under profiles/, marked as such).
  configs    device-timed Mrays/s, work per ray, algorithmic bytes and roofline of BASELINE configs 1, 3, 4 and 5, each
             next to the compiled reference on a bounded sample of the same rays (N = 1), sharded with gather for 4 and 5.
  cpu_baseline  the compiled reference (oracle/_ref, rtcIntersect1 per ray on all host threads) on the same rays.

Scenes are images of the REFERENCE's committed BVH (built by its CPU builder at build() time, scenes/_built/) converted by
the product's importer; nothing under oracle/ is touched outside the cpu_baseline / --impl reference legs.
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

WIDTH, HEIGHT = 1920, 1080
METRIC = "Mrays/s primary+diffuse (device-timed)"
UNIT = "Mrays/s"
TILE = 64
C4_TILE = 16   # BASELINE config 4 (8K instanced frame)
SCENES = os.path.join(ROOT, "scenes", "_built")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


SUBDIV_LEVEL, COMP_LEVEL = 6, 3   # BASELINE.json configs[1] (build/bomberman.ecs); --levels S,C selects another tessellation


def scene_path(leaf):
    return os.path.join(SCENES, "bomberman_s%dc%d_%s.rimg" % (SUBDIV_LEVEL, COMP_LEVEL, leaf))


def workload_config(leaf, n_primary, n_bounce):
    """Identical in both arms (ours / --impl reference) and at every N: the driver compares it."""
    return {"workload": "bomberman.obj subdiv S=%d C=%d compress.%s, %dx%d primary + 1 cosine-diffuse bounce"
                        % (SUBDIV_LEVEL, COMP_LEVEL, leaf, WIDTH, HEIGHT),
            "rays_primary": int(n_primary), "rays_bounce_active": int(n_bounce),
            "l2": "inputs larger than L2: every step traces a fresh copy of both ray buffers (2 x 166 MB per frame); the S6/C3 "
                  "scene itself (18.8 MB) is L2-resident by construction of the workload"}


class ClockSampler:
    """SM clock and throttle reasons sampled WHILE the timed region runs (NVML polled from a thread every ~2 ms; the
    region is tens of milliseconds, too short for `nvidia-smi -lms`), nvidia-smi as fallback."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False
        self.thread = None
        self.how = None

    def _poll_nvml(self, pynvml, handle):
        while not self.stop_flag:
            try:
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM))
                bits = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                for name, bit in self.REASONS:
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                try:
                    phys = int(vis.split(",")[self.index])
                except Exception:
                    phys = self.index
            handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)
            self.how = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, args=(pynvml, handle), daemon=True)
            self.thread.start()
        except Exception:
            self.how = "nvidia-smi"

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1)
        if not self.samples:
            try:
                q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                    "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                self.samples = [float(out[0])]
                self.max_mhz = float(out[1])
                for (name, _), v in zip(self.REASONS, out[2:6]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
                self.how = "nvidia-smi (after the region)"
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz) if self.max_mhz else None,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "how": self.how}


def measured_peaks():
    """(HBM GB/s, source) from the driver-written MEASURED_PEAKS.json, else the profiling guide's fallback; plus the L2
    bandwidth this repo measured itself (profiles/measured_l2.json, tools/bw_probe.py)."""
    hbm, src = 6650.0, "fallback (B200_PROFILING.md)"
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            hbm, src = float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    l2 = None
    p = os.path.join(ROOT, "profiles", "measured_l2.json")
    if os.path.exists(p):
        try:
            l2 = json.load(open(p))
        except Exception:
            l2 = None
    return hbm, src, l2


def captured_counters(key):
    """ncu counters of the committed capture (profiles/roofline_captured.json): STATIC, copied from a profiler run of the
    same kernel on the same workload -- never measured by this process."""
    p = os.path.join(ROOT, "profiles", "roofline_captured.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------------------------
# rays of the headline workload: the same arrays in both arms
# ------------------------------------------------------------------------------------------------------------------
def config2_rays(rg, leaf, reference_scene=None):
    """(primary RTCRayHit[N], bounce RTCRayHit[N], scene diagonal).  The bounce rays leave the primary hit points of the
    compiled reference (hit distance + rtcInterpolate shading normal per pixel, stored at build() time) -- when that file
    did not travel they are recomputed here by the reference itself (reference arm) or the run stops (our arm)."""
    prim = rg.primary_rays(WIDTH, HEIGHT, **rg.BOMBERMAN_CAMERA)
    path = os.path.join(SCENES, "bench_c2_primary_hits.npz")
    if os.path.exists(path) and (SUBDIV_LEVEL, COMP_LEVEL) == (6, 3):
        z = np.load(path)
        t, normals, bounds = z["t"], z["normal_i16"].astype(np.float32) / 32767.0, z["bounds"]
    elif reference_scene is not None:
        first = prim.copy()
        reference_scene.intersect(first)
        hit = first.view(np.uint32)[:, 18] != 0xFFFFFFFF
        t = np.where(hit, first[:, 8], np.float32(np.inf)).astype(np.float32)
        normals = reference_scene.hit_normals(first)
        bounds = reference_scene.bounds()
    else:
        return prim, None, None
    first = prim.copy()
    hit = np.isfinite(t)
    first[:, 8] = t
    first.view(np.uint32)[hit, 18] = 0
    diag = float(np.linalg.norm(np.asarray(bounds[3:], np.float64) - np.asarray(bounds[:3], np.float64)))
    xs, ys = np.meshgrid(np.arange(WIDTH, dtype=np.uint32), np.arange(HEIGHT, dtype=np.uint32))
    bounce = rg.diffuse_rays(first, normals, diag, sample=0, pixel_xy=(xs.ravel(), ys.ravel()))
    return prim, bounce, diag


def bounce_from_device_hits(scene, prim, rg, torch, dev):
    """Fallback of our arm when the build-time hit file is missing (other tessellation levels): block-frame normals."""
    d = torch.from_numpy(prim).to(dev)
    blocks = torch.empty(len(prim), dtype=torch.int32, device=dev)
    scene.intersect_device(d.data_ptr(), len(prim), 80, torch.cuda.current_stream().cuda_stream, blocks.data_ptr())
    torch.cuda.synchronize()
    first = d.cpu().numpy()
    blk = blocks.cpu().numpy().view(np.uint32)
    hit = first.view(np.uint32)[:, 18] != 0xFFFFFFFF
    normals = np.zeros((len(prim), 3), np.float32)
    normals[hit] = scene.block_normals()[blk[hit]]
    b = scene.info()["bounds"]
    diag = float(np.linalg.norm(np.array(b[3:]) - np.array(b[:3])))
    xs, ys = np.meshgrid(np.arange(WIDTH, dtype=np.uint32), np.arange(HEIGHT, dtype=np.uint32))
    return rg.diffuse_rays(first, normals, diag, sample=0, pixel_xy=(xs.ravel(), ys.ravel()))


# ------------------------------------------------------------------------------------------------------------------
# multi-GPU: one frame sharded over the ranks, hit records gathered into rank 0's frame buffer by peer copies
# ------------------------------------------------------------------------------------------------------------------
class FrameGather:
    """Rank 0 owns the frame's compact result arrays (geomID[n], {tfar,u,v,primID}[n], optionally instID[n]); the other ranks
    map them through CUDA IPC and copy their shard's records straight into place over NVLink ('peer copies only
    for that step').  Falls back to NCCL point-to-point when IPC mapping is not available."""

    def __init__(self, ecb, torch, dist, rank, world, n_total, with_inst):
        self.ecb, self.torch, self.dist, self.rank, self.world = ecb, torch, dist, rank, world
        self.n_total, self.with_inst = n_total, with_inst
        self.sizes = [4 * n_total, 16 * n_total] + ([4 * n_total] if with_inst else [])
        self.ptrs = None       # device pointers of the frame arrays as seen from THIS rank
        self.owned = []
        self.transport = "fused: the traversal kernels store their hit records straight into rank 0's frame arrays (CUDA IPC mapping, peer stores over NVLink)"
        self.fallback = False
        handles = [None]
        if rank == 0:
            self.owned = [ecb.device_alloc(s) for s in self.sizes]
            self.ptrs = list(self.owned)
            try:
                handles = [[ecb.ipc_export(p) for p in self.owned]]
            except Exception as e:  # pragma: no cover - depends on the box
                handles = [repr(e)]
        if world > 1:
            dist.broadcast_object_list(handles, src=0)
            ok = isinstance(handles[0], list)
            if rank != 0 and ok:
                try:
                    self.ptrs = [ecb.ipc_import(h) for h in handles[0]]
                except Exception as e:  # pragma: no cover
                    log("rank %d: IPC import failed (%r)" % (rank, e))
                    ok = False
            flag = torch.tensor([1 if ok else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                self.transport = "NCCL send/recv after each pass (CUDA IPC mapping unavailable on this box)"
                self.fallback = True
                if rank != 0:
                    self.ptrs = None

    def frame_tensor(self, which):
        """Rank 0: torch view of frame array `which` (0 geomID int32, 1 hit float32 x4, 2 instID int32) for checks."""
        assert self.rank == 0
        import ctypes as C
        torch = self.torch
        n = self.n_total * (4 if which == 1 else 1)
        dt = torch.float32 if which == 1 else torch.int32
        buf = torch.empty(n, dtype=dt, device="cuda")
        self.ecb.peer_copy_async(buf.data_ptr(), self.owned[which], 4 * n, torch.cuda.current_stream().cuda_stream)
        torch.cuda.current_stream().synchronize()
        return buf

    def close(self):
        if self.rank != 0 and self.ptrs:
            for p in self.ptrs:
                self.ecb.ipc_close(p)
        for p in self.owned:
            self.ecb.device_free(p)
        self.ptrs, self.owned = None, []


class ShardedPass:
    """One traversal pass over this rank's shard of a ray set, compact results.  trace only: results into local arrays.
    With gather: the traversal kernels store their hit records STRAIGHT INTO rank 0's frame arrays (IPC-mapped peer memory:
    the stores travel over NVLink while the kernel runs -- traversal and gather are one kernel, nothing is copied afterwards).
    Without IPC mapping: NCCL send / recv of the local arrays after the pass."""

    def __init__(self, ecb, torch, scene, rays, offset, gather, occluded=False, chunk=None, counts=None):
        self.ecb, self.torch, self.scene, self.rays, self.offset, self.gather = ecb, torch, scene, rays, offset, gather
        self.counts = counts   # records per rank (fallback transport only)
        self.n = rays.shape[0]
        self.stride = rays.shape[1] * 4
        self.occluded = occluded
        self.chunk = chunk or self.n
        dev = rays.device
        self.inst = gather.with_inst if gather is not None else False
        self.local = (torch.empty(self.n, dtype=torch.int32, device=dev), torch.empty((self.n, 4), dtype=torch.float32, device=dev),
                      torch.empty(self.n, dtype=torch.int32, device=dev) if self.inst else None)

    def out_ptrs(self, to_frame, first):
        if to_frame:
            p = self.gather.ptrs
            o = self.offset + first
            return p[0] + 4 * o, p[1] + 16 * o, (p[2] + 4 * o) if self.inst else 0
        g, h, i = self.local
        return g.data_ptr() + 4 * first, h.data_ptr() + 16 * first, (i.data_ptr() + 4 * first) if i is not None else 0

    def launch(self, stream, gather=False):
        """Enqueue the pass on `stream`.  Returns the number of kernel launches."""
        to_frame = gather and self.gather is not None and self.gather.ptrs is not None and not self.gather.fallback
        launches = 0
        for first in range(0, self.n, self.chunk):
            m = min(self.chunk, self.n - first)
            gp, hp, ip = self.out_ptrs(to_frame, first)
            rp = self.rays.data_ptr() + first * self.stride
            if self.occluded:
                self.scene.occluded_device_compact(rp, m, gp, self.stride, stream.cuda_stream)
            else:
                self.scene.intersect_device_compact(rp, m, gp, hp, self.stride, stream.cuda_stream, ip)
            launches += 1
        if gather and self.gather is not None and self.gather.fallback:
            self._fallback_exchange(stream)
        return launches

    def _fallback_exchange(self, stream):
        """No IPC mapping: every rank sends its records to rank 0 with NCCL point-to-point on the launching stream."""
        g, dist, torch = self.gather, self.gather.dist, self.torch
        with torch.cuda.stream(stream):
            if g.rank == 0:
                if not hasattr(self, "_recv"):
                    self._recv = [(torch.empty(c, dtype=torch.int32, device=self.rays.device), torch.empty((c, 4), dtype=torch.float32, device=self.rays.device))
                                  for c in self.counts]
                for r in range(1, g.world):
                    dist.recv(self._recv[r][0], src=r)
                    if not self.occluded:
                        dist.recv(self._recv[r][1], src=r)
            else:
                lg, lh, _ = self.local
                dist.send(lg, dst=0)
                if not self.occluded:
                    dist.send(lh, dst=0)


class FramePipeline:
    """The two passes of a frame on two streams: bounce(k) is ordered after primary(k) by an event (the bounce rays of a
    renderer are made from the primary hits), primary(k+1) is not ordered after bounce(k) (a double-buffered renderer has the
    next frame's camera rays ready).  What it buys: the END of a persistent launch is a handful of warps finishing their
    longest rays while the other SMs idle (`tools/overlap_probe.py`, `profiles/r2z_overlap.txt`: 0.675 -> 0.553 ms per
    frame); with the next pass already queued on the other stream its CTAs take the SMs as they fall free."""

    def __init__(self, torch, first_stream):
        self.torch = torch
        self.streams = [first_stream, torch.cuda.Stream()]

    def frame(self, primary, bounce):
        a, b = self.streams
        primary(a)
        e = self.torch.cuda.Event()
        e.record(a)
        b.wait_event(e)
        bounce(b)


def interleaved_tiles(width, height, rank, world, sh, tile=None):
    """Pixel indices of this rank's tiles (tile t belongs to rank t mod world: sky and geometry tiles spread evenly),
    tile after tile, row-major inside a tile; and the offset of this rank's records in the rank-major frame arrays."""
    return sh.interleaved_tile_indices(width, height, rank, world, tile or TILE)


def timed_steps(torch, dist, dev, K, W, enqueue_step, streams):
    """W warm-up steps, then exactly K steps bracketed by barrier + synchronize; returns the max over ranks of the elapsed
    milliseconds between the first launch and the end of the LAST stream to finish."""
    for i in range(W):
        enqueue_step(i)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    start = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in streams]
    start.record(streams[0])
    for s in streams[1:]:
        s.wait_event(start)
    for i in range(W, W + K):
        enqueue_step(i)
    for s, e in zip(streams, ends):
        e.record(s)
    torch.cuda.synchronize()
    ms = max(start.elapsed_time(e) for e in ends)
    if dist:
        dist.barrier()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------------------------
# roofline helpers
# ------------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(c, n_records, leaf, in_bytes=80.0, hit_bytes=48.0):
    """ray in + hit out + 112 B per top-level node + 144 B per block header + 4 B per quantized node + leaf
    payload per cell (2 B heights / 4 x 12 B vertices / none); every record is read, hits are written.
    c = counters of the instrumented kernel: rays, top nodes, blocks, quantized nodes, cells, hits."""
    leaf_bytes = {"leaf": 2.0, "grid": 48.0, "box": 0.0}[leaf]
    return in_bytes * n_records + hit_bytes * c[5] + 112.0 * c[1] + 144.0 * c[2] + 4.0 * c[3] + leaf_bytes * c[4]


def work_per_ray(c):
    r = max(1.0, c[0])
    return {"top_nodes": round(c[1] / r, 3), "blocks": round(c[2] / r, 3), "cnodes": round(c[3] / r, 3), "cells": round(c[4] / r, 3),
            "hit_fraction": round(c[5] / r, 4)}


def time_kernel(torch, fn, src, reps=5):
    """Average CUDA-event duration of fn(tmp) over reps-1 launches, each on a fresh copy of `src` (first one discarded)."""
    ts = []
    for _ in range(reps):
        tmp = src.clone()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(tmp)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts[1:])), tmp


def counters_of(torch, scene, src, stream, dev):
    cnt = torch.zeros(6, dtype=torch.int64, device=dev)
    tmp = src.clone()
    scene.intersect_device_stats(tmp.data_ptr(), tmp.shape[0], cnt.data_ptr(), tmp.shape[1] * 4, stream)
    torch.cuda.synchronize()
    return cnt.cpu().numpy().astype(np.float64)


def roofline_entry(kernel, ms, bytes_alg, hbm_peak, peak_src, l2, captured):
    achieved = bytes_alg / (ms * 1e-3) / 1e9
    out = {"bound": (captured or {}).get("bound", "l2/issue (scene L2-resident; see captured counters)"),
           "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4),
           "traffic": (captured or {}).get("dram_bytes_per_launch"), "peak_source": peak_src, "kernel": kernel,
           "what": "achieved = ALGORITHMIC bytes per launch (live: instrumented-kernel counters x SURVEY 8d byte sizes) / live CUDA-event "
                   "duration; frac = achieved / HBM peak.  `traffic` and `captured` are STATIC copies of the committed ncu capture"}
    if captured:
        cap = dict(captured)
        if l2 and cap.get("lts_bytes_per_launch") and cap.get("ncu_duration_us"):
            l2_gbs = cap["lts_bytes_per_launch"] / (cap["ncu_duration_us"] * 1e-6) / 1e9
            cap["l2_gbs"] = round(l2_gbs, 1)
            cap["l2_frac_of_measured_peak"] = round(l2_gbs / l2["l2_read_gbs"], 4)
            cap["l2_peak_gbs_measured"] = l2["l2_read_gbs"]
        if cap.get("dram_bytes_per_launch") and cap.get("ncu_duration_us"):
            cap["dram_frac_of_hbm_peak"] = round(cap["dram_bytes_per_launch"] / (cap["ncu_duration_us"] * 1e-6) / 1e9 / hbm_peak, 4)
        if cap.get("issue_active_pct") and cap.get("threads_per_instruction"):
            cap["issue_slot_utilisation"] = round(cap["issue_active_pct"] / 100.0 * cap["threads_per_instruction"] / 32.0, 4)
        out["captured"] = cap
    return out


# ------------------------------------------------------------------------------------------------------------------
# BASELINE configs 1, 3, 4, 5 (device-timed, with the compiled reference beside each at N = 1)
# ------------------------------------------------------------------------------------------------------------------
def reference_rate(make_ref, jobs, threads=None):
    """Mrays/s of the compiled reference on bounded samples (rtcIntersect1 / rtcOccluded1 per ray, all host threads).
    jobs: list of (rays numpy, occluded flag); the scene is committed once.  Returns one dict per job (None without oracle/_ref)."""
    from oracle import pyoracle as po
    if not po.have_reference():
        return [None] * len(jobs)
    threads = threads or po.hardware_threads()
    ref, free = make_ref(po)
    outs = []
    for rays_np, occluded in jobs:
        active = int((rays_np[:, 3] <= rays_np[:, 8]).sum())
        best = None
        for _ in range(2):
            r = rays_np.copy()
            t0 = time.perf_counter()
            (ref.occluded if occluded else ref.intersect)(r, threads)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        outs.append({"value": round(active / best / 1e6, 2), "unit": UNIT, "cores": threads, "kind": "reference",
                     "sample": "%d rays (%d active), best of 2 passes, %s per ray on %d std::threads, AVX2 build"
                               % (len(rays_np), active, "rtcOccluded1" if occluded else "rtcIntersect1", threads)})
    free()
    return outs


def bench_configs(args, which, ecb, rg, rgd, sh, torch, dist, dev, rank, world, hbm, peak_src, l2, sm_count, sm_mhz):
    import __graft_entry__ as ge
    stream = torch.cuda.current_stream()
    out = {}
    cpu = world == 1 and not args.no_cpu_baseline

    def finish(entry, n_active, ms, kernel_name, c=None, leaf=None, n_records=None, in_bytes=80.0, hit_bytes=48.0, cap_key=None):
        entry["rays"] = int(n_active)
        entry["kernel_ms"] = round(ms, 4)
        entry["value"] = round(n_active / ms / 1e3, 1)
        entry["unit"] = UNIT
        entry["rays_per_sm_cycle"] = round(n_active / (ms * 1e-3 * sm_mhz * 1e6 * sm_count), 5)
        if c is not None:
            b = algorithmic_bytes(c, n_records, leaf, in_bytes, hit_bytes)
            entry["work_per_ray"] = work_per_ray(c)
            entry["algorithmic_bytes_per_ray"] = round(b / max(1, n_active), 1)
            entry["roofline"] = roofline_entry(kernel_name, ms, b, hbm, peak_src, l2, captured_counters(cap_key) if cap_key else None)
        return entry

    # ---- config 1: displaced cube S6/C4 + ground plane (the tutorial's scene), 1024^2 primary -------------------------
    if 1 in which and world == 1:
        path = os.path.join(SCENES, "cubeplane_s6c4_leaf.rimg")
        if os.path.exists(path):
            sc = ecb.Scene.from_image_file(path).upload(dev.index)
            rays = rgd.primary_rays(1024, 1024, device=dev, **rg.CUBE_CAMERA)
            ms, res = time_kernel(torch, lambda t: sc.intersect_device(t.data_ptr(), t.shape[0], 80, stream.cuda_stream), rays)
            c = counters_of(torch, sc, rays, stream.cuda_stream, dev)
            e = {"workload": "tutorials/displacement_geometry: displaced cube S=6 C=4 compress.leaf + ground plane (triangle accel first, "
                             "AccelN order), 1024x1024 primary; counters cover the compressed-accel pass"}
            finish(e, rays.shape[0], ms, "ecb_triangle_kernel + ecb_intersect_kernel<leaf>", c, "leaf", rays.shape[0], cap_key="config1_primary")
            if cpu:
                def mk(po):
                    r = po.RefScene("leaf", 6, 4)
                    r.add_cube(1, 1.0)
                    r.add_ground_plane()
                    r.commit()
                    return r, r.free
                e["cpu_reference"] = reference_rate(mk, [(rays.cpu().numpy(), False)])[0]
            out["1"] = e
            sc.release()
        else:
            out["1"] = {"unavailable": "scenes/_built/cubeplane_s6c4_leaf.rimg missing (run __graft_entry__.build() where the reference is mounted)"}

    # ---- config 3: high-res displacement (cube S8/C4), 4K primary + shadow rays through rtcOccluded1M ----------------
    if 3 in which and world == 1:
        path = os.path.join(SCENES, "cube_s8c4_leaf.rimg")
        if os.path.exists(path):
            sc = ecb.Scene.from_image_file(path).upload(dev.index)
            rays = rgd.primary_rays(3840, 2160, device=dev, **rg.CUBE_CAMERA)
            ms, res = time_kernel(torch, lambda t: sc.intersect_device(t.data_ptr(), t.shape[0], 80, stream.cuda_stream), rays, reps=4)
            c = counters_of(torch, sc, rays, stream.cuda_stream, dev)
            prim = finish({"workload": "displaced cube S=8 C=4 compress.leaf (786 432 micro-triangles), 3840x2160 primary"}, rays.shape[0], ms,
                          "ecb_intersect_kernel<leaf>", c, "leaf", rays.shape[0], cap_key="config3_primary")
            shadow = rgd.shadow_rays(res, light_center=(3.0, 5.0, -2.0), light_radius=1.0)
            n_shadow = int((shadow[:, 3] <= shadow[:, 8]).sum().item())
            ms_s, _ = time_kernel(torch, lambda t: sc.occluded_device(t.data_ptr(), t.shape[0], 48, stream.cuda_stream), shadow, reps=4)
            sh_e = {"workload": "one shadow ray per primary hit toward a sphere light, rtcOccluded1M semantics (block-box any-hit, compressed.h:754-756)",
                    "rays": n_shadow, "kernel_ms": round(ms_s, 4), "value": round(n_shadow / ms_s / 1e3, 1), "unit": UNIT,
                    "algorithmic_bytes_per_ray": "48 in + 4 out + 112 per top-level node tested (no instrumented any-hit variant)"}
            ms_x, _ = time_kernel(torch, lambda t: ecb._check(ecb._lib().ecbOccludedExactDevice(sc._h, 0xFFFFFFFF, t.data_ptr(), t.shape[0], 48, stream.cuda_stream), "exact"), shadow, reps=4)
            sh_e["exact_any_hit_opt_in"] = {"kernel_ms": round(ms_x, 4), "value": round(n_shadow / ms_x / 1e3, 1)}
            e = {"workload": "high-res displacement single object, 4K primary + shadow rays via rtcOccluded1M", "primary": prim, "shadow": sh_e,
                 "value": round((rays.shape[0] + n_shadow) / (ms + ms_s) / 1e3, 1), "unit": UNIT}
            if cpu:
                def mk(po):
                    r = po.RefScene("leaf", 8, 4)
                    r.add_cube(1, 1.0)
                    r.commit()
                    return r, r.free
                e["primary"]["cpu_reference"], e["shadow"]["cpu_reference"] = reference_rate(
                    mk, [(rays[::4].cpu().numpy(), False), (shadow[::4].cpu().numpy(), True)])
            out["3"] = e
            sc.release()
            del rays, res, shadow
        else:
            out["3"] = {"unavailable": "scenes/_built/cube_s8c4_leaf.rimg missing"}

    # ---- config 4: instanced scene (25 instances, compressed grids), 8K frame sharded by interleaved tiles ------------
    if 4 in which:
        path = os.path.join(SCENES, ge.CONFIG4_BLOB)
        if os.path.exists(path):
            sc = ecb.Scene.load(path).upload(dev.index)
            W8, H8 = 7680, 4320
            # the frame is submitted tile after tile at every N (what a tile renderer does, viewer_device.cpp:448-451; the one-thread-
            # per-ray instancing kernel gains 10 % from it over row-major order, profiles/r2y_ray_order.txt)
            idx, offset, counts = interleaved_tiles(W8, H8, rank, world, sh, tile=C4_TILE)
            rays = rgd.primary_rays(W8, H8, device=dev, pixel_index=torch.from_numpy(idx).to(dev), **ge.CONFIG4_CAMERA)
            gather = FrameGather(ecb, torch, dist, rank, world, W8 * H8, with_inst=True)
            p = ShardedPass(ecb, torch, sc, rays, offset, gather, chunk=1 << 22, counts=counts)
            ms = timed_steps(torch, dist if world > 1 else None, dev, 3, 3, lambda i: p.launch(stream), [stream])
            e = {"workload": "25 instances (13 x bomberman S6/C3, 12 x displaced cube S6/C3, compress.grid) with distinct affine transforms, "
                             "7680x4320 primary in %dx%d tiles" % (C4_TILE, C4_TILE) + (", sharded over %d ranks (tile t -> rank t mod N)" % world if world > 1 else ""),
                 "rays": W8 * H8, "ms_per_frame": round(ms / 3, 4), "value": round(W8 * H8 * 3 / ms / 1e3, 1), "unit": UNIT,
                 "results": "compact hit records (geomID, {tfar,u,v,primID}, instID = 24 bytes per ray)"}
            if world > 1:
                ms_g = timed_steps(torch, dist, dev, 3, 3, lambda i: p.launch(stream, gather=True), [stream])
                e["value_with_gather"] = round(W8 * H8 * 3 / ms_g / 1e3, 1)
                e["gather"] = {"transport": gather.transport, "bytes_into_rank0_per_frame_upper_bound": int(24 * (W8 * H8 - counts[0])), "bytes_per_ray": "4 (geomID, every ray) + 20 per hit",
                               "ms_per_frame_with_gather": round(ms_g / 3, 4)}
            if world == 1:
                sample = rays[::8].contiguous()   # every 8th pixel of the frame
                c = counters_of(torch, sc, sample, stream.cuda_stream, dev)
                if c[0] > 0:   # counters[0] = object-scene traversals (instance visits); the others are summed over all visits
                    n_s = float(sample.shape[0])
                    e["work_per_ray"] = {"instance_visits": round(c[0] / n_s, 3), "top_nodes": round(c[1] / n_s, 3), "blocks": round(c[2] / n_s, 3),
                                         "cnodes": round(c[3] / n_s, 3), "cells": round(c[4] / n_s, 3), "hit_fraction": round(c[5] / n_s, 4),
                                         "sample": "every 8th pixel; object-scene nodes only (the instance-level BVH4 has 10 nodes)"}
                    bts = 80.0 * n_s + 24.0 * c[5] + 64.0 * c[0] + 112.0 * c[1] + 144.0 * c[2] + 4.0 * c[3] + 48.0 * c[4]
                    e["algorithmic_bytes_per_ray"] = round(bts / n_s, 1)
                    e["roofline"] = roofline_entry("ecb_instanced_kernel (compact results)", ms / 3, bts * (W8 * H8 / n_s), hbm, peak_src, l2,
                                                   captured_counters("config4"))
                del sample
                if cpu:
                    def mk(po):
                        sys.path.insert(0, os.path.join(ROOT, "tests"))
                        from instancing_util import build_instanced
                        top, objs = build_instanced(po, ge.CONFIG4_SPEC["leaf"], ge.CONFIG4_SPEC["objects"], grid=ge.CONFIG4_SPEC["grid"],
                                                    spacing=ge.CONFIG4_SPEC["spacing"])

                        def free():
                            top.free()
                            for o in objs:
                                o.free()
                        return top, free
                    e["cpu_reference"] = reference_rate(mk, [(rays[::16].cpu().numpy(), False)])[0]
            out["4"] = e
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            gather.close()
            sc.release()
            del rays, p
        else:
            out["4"] = {"unavailable": "scenes/_built/%s missing" % ge.CONFIG4_BLOB}

    # ---- config 5: 64 Mi incoherent cosine-diffuse rays over bomberman, sharded by index range ------------------------
    if 5 in which:
        path, hits = scene_path("leaf"), os.path.join(SCENES, ge.CONFIG2_HITS)
        if os.path.exists(path) and os.path.exists(hits):
            sc = ecb.Scene.from_image_file(path).upload(dev.index)
            z = np.load(hits)
            t_np = z["t"]
            hit_pix = np.nonzero(np.isfinite(t_np))[0]
            n_hit, M5 = len(hit_pix), 1 << 26
            b, e_ = sh.index_range(M5, rank, world)
            prim_all = rg.primary_rays(WIDTH, HEIGHT, **rg.BOMBERMAN_CAMERA)
            P_hit = torch.from_numpy((prim_all[hit_pix, 0:3].astype(np.float64) + t_np[hit_pix, None].astype(np.float64) * prim_all[hit_pix, 4:7]).astype(np.float32)).to(dev)
            N_hit = torch.from_numpy(z["normal_i16"][hit_pix].astype(np.float32) / 32767.0).to(dev)
            D_hit = torch.from_numpy(prim_all[hit_pix, 4:7]).to(dev)
            pix_t = torch.from_numpy(hit_pix).to(dev)
            diag = float(np.linalg.norm(z["bounds"][3:].astype(np.float64) - z["bounds"][:3].astype(np.float64)))
            rays = torch.empty((e_ - b, 20), dtype=torch.float32, device=dev)
            step = 1 << 22
            for first in range(b, e_, step):   # ray i: hit pixel i mod n_hit, sample index i div n_hit
                i = torch.arange(first, min(first + step, e_), dtype=torch.int64, device=dev)
                k, s_idx = i % n_hit, i // n_hit
                pix = pix_t[k]
                rays[first - b: first - b + i.numel()] = rgd.diffuse_rays(P_hit[k], N_hit[k], D_hit[k], pix % WIDTH, pix // WIDTH, s_idx, 1e-3 * diag)
            gather = FrameGather(ecb, torch, dist, rank, world, M5, with_inst=False)
            p = ShardedPass(ecb, torch, sc, rays, b, gather, chunk=1 << 22,
                            counts=[sh.index_range(M5, r, world)[1] - sh.index_range(M5, r, world)[0] for r in range(world)])
            ms = timed_steps(torch, dist if world > 1 else None, dev, 3, 3, lambda i: p.launch(stream), [stream])
            e = {"workload": "64 Mi (67 108 864) cosine-diffuse rays leaving the primary hit points of the 1920x1080 frame (%d samples per hit, "
                             "tutorial sampler) over bomberman S6/C3 compress.leaf%s" % ((M5 + n_hit - 1) // n_hit, ", sharded over %d ranks by index range" % world if world > 1 else ""),
                 "rays": M5, "ms_per_batch": round(ms / 3, 4), "value": round(M5 * 3 / ms / 1e3, 1), "unit": UNIT,
                 "results": "compact hit records (geomID, {tfar,u,v,primID} = 20 bytes per ray); the 5.4 GB ray batch is read-only and larger than L2"}
            if world > 1:
                ms_g = timed_steps(torch, dist, dev, 3, 3, lambda i: p.launch(stream, gather=True), [stream])
                e["value_with_gather"] = round(M5 * 3 / ms_g / 1e3, 1)
                e["gather"] = {"transport": gather.transport, "bytes_into_rank0_per_batch_upper_bound": int(20 * (M5 - (e_ - b if rank == 0 else M5 // world))), "bytes_per_ray": "4 (geomID, every ray) + 16 per hit",
                               "ms_per_batch_with_gather": round(ms_g / 3, 4)}
            if world == 1:
                sample = rays[: 1 << 23]
                c = counters_of(torch, sc, sample, stream.cuda_stream, dev)
                bts = algorithmic_bytes(c, sample.shape[0], "leaf", 80.0, 20.0)
                e["work_per_ray"] = work_per_ray(c)
                e["algorithmic_bytes_per_ray"] = round(bts / sample.shape[0], 1)
                e["roofline"] = roofline_entry("ecb_intersect_kernel<leaf> (compact results)", ms / 3, bts * (M5 / sample.shape[0]), hbm, peak_src, l2,
                                               captured_counters("config5"))
                inc = rgd.incoherent_rays(1 << 23, z["bounds"], dev, seed=1)
                ms_i, _ = time_kernel(torch, lambda t: sc.intersect_device(t.data_ptr(), t.shape[0], 80, stream.cuda_stream), inc, reps=4)
                ci = counters_of(torch, sc, inc, stream.cuda_stream, dev)
                e["authors_incoherent"] = {"workload": "the authors' generator (viewer_device.cpp:367-392): 8 Mi rays, origin and target uniform in the scene AABB, seeded",
                                           "rays": 1 << 23, "kernel_ms": round(ms_i, 4), "value": round((1 << 23) / ms_i / 1e3, 1), "unit": UNIT,
                                           "work_per_ray": work_per_ray(ci)}
                if cpu:
                    def mk(po):
                        r = po.RefScene("leaf", 6, 3)
                        r.add_obj()
                        r.commit()
                        return r, r.free
                    e["cpu_reference"], e["authors_incoherent"]["cpu_reference"] = reference_rate(
                        mk, [(rays[::8].cpu().numpy(), False), (inc.cpu().numpy(), False)])
                del inc
            out["5"] = e
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            gather.close()
            sc.release()
            del rays, p
        else:
            out["5"] = {"unavailable": "scene image or %s missing" % ge.CONFIG2_HITS}
    return out


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world):
    import torch
    import embree_compressed_b200 as ecb
    from embree_compressed_b200 import raygen as rg
    from embree_compressed_b200 import raygen_device as rgd
    from embree_compressed_b200 import sharding as sh

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the traversal path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = host_group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")   # host-side barriers: a waiting NCCL kernel would sit on the GPU
    path = scene_path(args.leaf)
    if not os.path.exists(path):
        raise SystemExit("bench.py: %s missing -- run __graft_entry__.build() where /root/reference is mounted" % path)
    scene = ecb.Scene.from_image_file(path).upload(local)
    info = scene.info()
    prim, bounce, _ = config2_rays(rg, args.leaf)
    bounce_source = "hit distances + rtcInterpolate shading normals of the compiled reference, computed at build() time (the reference arm traces the same arrays)"
    if bounce is None:
        bounce = bounce_from_device_hits(scene, prim, rg, torch, dev)
        bounce_source = "block-frame normals of this run's own primary hits (scenes/_built/bench_c2_primary_hits.npz missing or other tessellation level)"
    n_prim = len(prim)
    n_bounce = int((bounce[:, 3] <= bounce[:, 8]).sum())
    rays_per_step = n_prim + n_bounce
    K, W = args.steps, args.warmup
    stream_t = torch.cuda.current_stream()
    stream = stream_t.cuda_stream
    hbm, peak_src, l2 = measured_peaks()
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count

    # ---- this rank's shard of the frame ----
    if world > 1:
        idx, offset, counts = interleaved_tiles(WIDTH, HEIGHT, rank, world, sh)
    else:
        idx, offset, counts = np.arange(n_prim), 0, [n_prim]
    d_prim0 = torch.from_numpy(prim[idx]).to(dev)
    d_bounce0 = torch.from_numpy(bounce[idx]).to(dev)
    n_local = len(idx)

    sampler = ClockSampler(local)
    pipe = None if args.one_stream else FramePipeline(torch, stream_t)
    timed_streams = [stream_t] if pipe is None else pipe.streams
    ms_one = None
    gather_info = None
    value_with_gather = None
    if world == 1:
        # ---- device-timed value: a fresh copy of the ray buffers for every step, results written in place (rtcIntersect1M semantics) ----
        bufs = [(d_prim0.clone(), d_bounce0.clone()) for _ in range(K + W)]

        def step_one_stream(i):
            p, b = bufs[i]
            scene.intersect_device(p.data_ptr(), n_local, 80, stream)
            scene.intersect_device(b.data_ptr(), n_local, 80, stream)

        def step(i):   # the frame pipeline (FramePipeline): bounce(k) after primary(k), primary(k+1) may start meanwhile
            p, b = bufs[i]
            pipe.frame(lambda st: scene.intersect_device(p.data_ptr(), n_local, 80, st.cuda_stream),
                       lambda st: scene.intersect_device(b.data_ptr(), n_local, 80, st.cuda_stream))

        if pipe is None:
            step = step_one_stream
        else:   # the same frames back to back on one stream first (reported next to the value; also leaves the buffers traced once)
            ms_one = timed_steps(torch, None, dev, K, W, step_one_stream, [stream_t])
            for p, b in bufs:   # fresh rays again, in the same buffers
                p.copy_(d_prim0)
                b.copy_(d_bounce0)
            torch.cuda.synchronize()
        sampler.start()
        launches0 = ecb.launch_count()
        ms_max = timed_steps(torch, None, dev, K, W, step, timed_streams)
        launches = ecb.launch_count() - launches0 - 2 * W
        clocks = sampler.stop()
        result_mode = "in place (the RTCRayHit records on the device are updated, a fresh copy of both buffers every step)"
    else:
        # ---- one frame sharded over the ranks: compact results; trace only, then trace + gather into rank 0's frame ----
        g_prim = FrameGather(ecb, torch, dist, rank, world, n_prim, with_inst=False)
        g_bounce = FrameGather(ecb, torch, dist, rank, world, n_prim, with_inst=False)
        p_prim = ShardedPass(ecb, torch, scene, d_prim0, offset, g_prim, counts=counts)
        p_bounce = ShardedPass(ecb, torch, scene, d_bounce0, offset, g_bounce, counts=counts)

        def step_one_stream(i):
            p_prim.launch(stream_t)
            p_bounce.launch(stream_t)

        def step(i):
            pipe.frame(lambda st: p_prim.launch(st), lambda st: p_bounce.launch(st))

        def step_gather(i):
            if pipe is None:
                p_prim.launch(stream_t, gather=True)
                p_bounce.launch(stream_t, gather=True)
            else:
                pipe.frame(lambda st: p_prim.launch(st, gather=True), lambda st: p_bounce.launch(st, gather=True))

        if pipe is None:
            step = step_one_stream
        else:
            ms_one = timed_steps(torch, dist, dev, K, W, step_one_stream, [stream_t])
        if rank == 0:
            sampler.start()
        launches0 = ecb.launch_count()
        ms_max = timed_steps(torch, dist, dev, K, W, step, timed_streams)
        launches = ecb.launch_count() - launches0 - 2 * W
        clocks = sampler.stop() if rank == 0 else None
        ms_gather = timed_steps(torch, dist, dev, K, W, step_gather, timed_streams)
        value_with_gather = rays_per_step * K / (ms_gather * 1e-3) / 1e6
        # the gathered frame must hold exactly what a single device leaves (rank 0 re-traces the whole frame to check)
        if rank == 0 and not g_prim.fallback:
            full = torch.from_numpy(prim).to(dev)
            scene.intersect_device(full.data_ptr(), n_prim, 80, stream)
            torch.cuda.synchronize()
            order = np.concatenate([sh.tile_pixel_indices(WIDTH, sh.tile_grid(WIDTH, HEIGHT, TILE)[r::world]) for r in range(world)])
            want = full[torch.from_numpy(order).to(dev)]
            got_geom = g_prim.frame_tensor(0)
            got_hit = g_prim.frame_tensor(1).view(-1, 4)
            hitm = want.view(torch.int32)[:, 18] != -1
            assert bool((got_geom == want.view(torch.int32)[:, 18]).all()), "gathered geomIDs differ from the single-device frame"
            assert bool((got_hit[hitm].view(torch.int32) == want[hitm][:, [8, 15, 16, 17]].view(torch.int32)).all()), "gathered hit records differ"
            del full, want
        gather_info = {"transport": g_prim.transport, "bytes_per_ray": "4 (geomID, every ray) + 16 per hit", "bytes_into_rank0_per_step_upper_bound": int(2 * 20 * (n_prim - counts[0])),
                       "ms_per_step_with_gather": round(ms_gather / K, 4), "ms_per_step_trace_only": round(ms_max / K, 4),
                       "checked": "rank 0's gathered frame == a single-device trace of the whole frame, bit for bit"}
        dist.barrier()
        g_prim.close()
        g_bounce.close()
        result_mode = "compact hit records (geomID + {tfar,u,v,primID} = 20 bytes per ray), rays read-only"
    value = rays_per_step * K / (ms_max * 1e-3) / 1e6

    out = None
    e2e = None
    configs = {}
    # ---- per-kernel timing of both launches + algorithmic bytes from the instrumented variant (this rank's shard) ----
    fn = lambda t: scene.intersect_device(t.data_ptr(), t.shape[0], 80, stream)
    kernel_ms = {"primary": time_kernel(torch, fn, d_prim0)[0], "bounce": time_kernel(torch, fn, d_bounce0)[0]}
    c_prim = counters_of(torch, scene, d_prim0, stream, dev)
    c_bounce = counters_of(torch, scene, d_bounce0, stream, dev)
    nb_local = int(c_bounce[0])
    bytes_prim = algorithmic_bytes(c_prim, n_local, args.leaf)
    bytes_bounce = algorithmic_bytes(c_bounce, n_local, args.leaf)
    dom = "primary" if kernel_ms["primary"] >= kernel_ms["bounce"] else "bounce"
    cap_key = "%s_%s" % (args.leaf, dom) if (SUBDIV_LEVEL, COMP_LEVEL) == (6, 3) else "s%dc%d_%s_%s" % (SUBDIV_LEVEL, COMP_LEVEL, args.leaf, dom)
    roofline = roofline_entry("ecb_intersect_kernel<%s> (%s pass%s)" % (args.leaf, dom, ", rank 0's shard" if world > 1 else ""), kernel_ms[dom],
                              bytes_prim if dom == "primary" else bytes_bounce, hbm, peak_src, l2, captured_counters(cap_key) if world == 1 else None)
    roofline["kernel_ms"] = {k: round(v, 4) for k, v in kernel_ms.items()}
    roofline["algorithmic_bytes_per_ray"] = {"primary": round(bytes_prim / n_local, 1), "bounce": round(bytes_bounce / max(1, nb_local), 1)}
    roofline["work_per_ray"] = {"primary": work_per_ray(c_prim), "bounce": work_per_ray(c_bounce)}
    roofline["mrays_per_s"] = {"primary": round(n_local / kernel_ms["primary"] / 1e3, 1), "bounce": round(nb_local / kernel_ms["bounce"] / 1e3, 1)}
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    roofline["rays_per_sm_cycle_live"] = {k: round((n_local if k == "primary" else nb_local) / (kernel_ms[k] * 1e-3 * sm_mhz * 1e6 * sm_count), 5)
                                          for k in kernel_ms}
    roofline["step_share"] = {k: round(v / sum(kernel_ms.values()), 3) for k, v in kernel_ms.items()}

    # ---- e2e: the reference-facing call with pinned host buffers, copies inside the timed region.  At N > 1 rank 0 makes
    # ---- ONE call per pass on a scene resident on all N devices (ecbSceneUploadMulti); the other ranks stay off their GPUs.
    if rank == 0:
        if world > 1:
            # this leg is ONE process driving all N devices: give its host pool the whole host (the library's default divides
            # the cores by LOCAL_WORLD_SIZE for one-process-per-GPU use; the other ranks are parked behind a gloo barrier)
            os.environ.setdefault("ECB200_HOST_THREADS", str(max(2, min(24, (os.cpu_count() or 4) * 3 // 4))))
            multi = ecb.Scene.from_image_file(path).upload_multi(list(range(world)))
        else:
            multi = scene
        if args.e2e_memory == "pinned":
            hp, hp_ptr = ecb.host_alloc(prim.nbytes)
            hb, hb_ptr = ecb.host_alloc(bounce.nbytes)
            hp = hp.view(np.float32).reshape(prim.shape)
            hb = hb.view(np.float32).reshape(bounce.shape)
        else:  # what an unmodified embree application hands over: plain malloc'ed ray arrays
            hp, hb, hp_ptr, hb_ptr = np.empty_like(prim), np.empty_like(bounce), None, None
        e2e_steps = max(3, min(K, 10))
        e2e_t = []
        for i in range(2 + e2e_steps):
            hp[:] = prim
            hb[:] = bounce
            t0 = time.perf_counter()
            multi.intersect1M(hp)
            multi.intersect1M(hb)
            e2e_t.append(time.perf_counter() - t0)
        e2e_s = float(np.mean(e2e_t[2:]))
        e2e_value = rays_per_step / e2e_s / 1e6
        # the host-buffer call must leave exactly the bits the device-buffer call leaves
        chk_p, chk_b = torch.from_numpy(prim).to(dev), torch.from_numpy(bounce).to(dev)
        scene.intersect_device(chk_p.data_ptr(), n_prim, 80, stream)
        scene.intersect_device(chk_b.data_ptr(), n_prim, 80, stream)
        torch.cuda.synchronize()
        assert (hp.view(np.uint32) == chk_p.cpu().numpy().view(np.uint32)).all(), "e2e primary results differ from the device path"
        assert (hb.view(np.uint32) == chk_b.cpu().numpy().view(np.uint32)).all(), "e2e bounce results differ from the device path"
        del chk_p, chk_b
        host_path = os.environ.get("ECB200_HOST_PATH") or (ecb.host_path_for(args.e2e_memory == "pinned") if world == 1 else
                                                           ("compact" if args.e2e_memory == "pinned" else "packed"))
        n_rec = len(prim) + len(bounce)
        h2d = 32 * n_rec if host_path == "packed" else int(prim.nbytes + bounce.nbytes)
        d2h = int(prim.nbytes + bounce.nbytes) if host_path == "copy" else 20 * n_rec
        how = {"packed": "the library's host pool packs the 32 bytes per ray the path reads into pinned staging, 32-byte records go up, "
                         "20-byte compact hit records come back and are scattered into the caller's records by the same pool",
               "compact": "full records go up by DMA, 20-byte compact hit records come back and are scattered into the caller's records by the library's host pool",
               "copy": "full records copied both ways"}[host_path]
        e2e = {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "ms_per_step": round(e2e_s * 1e3, 3), "host_path": host_path, "caller_memory": args.e2e_memory, "devices": multi.resident_devices(),
               "call": "ecbIntersect1M (rtcIntersect1M drop-in), %s host RTCRayHit arrays (%d bytes), results written in place%s; %s"
                       % (args.e2e_memory, int(prim.nbytes + bounce.nbytes),
                          "; ONE call per pass in ONE process, the library cuts the ray range into one part per device (no gather step)" if world > 1 else "", how),
               "host_pool_threads": os.environ.get("ECB200_HOST_THREADS", "library default"),
               "checked": "bit-identical to the device-buffer path"}
        if hp_ptr is not None:
            ecb.host_free(hp_ptr)
            ecb.host_free(hb_ptr)
        if multi is not scene:
            multi.release()
    if dist:
        dist.barrier(group=host_group)

    # ---- the other BASELINE configs ----
    which = parse_configs(args.configs, world)
    if which:
        configs = bench_configs(args, which, ecb, rg, rgd, sh, torch, dist, dev, rank, world, hbm, peak_src, l2, sm_count, sm_mhz)

    if rank == 0:
        out = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": round(ms_max / K, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": workload_config(args.leaf, n_prim, n_bounce),
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
               "bytes_per_micro_triangle": {"device_format": round(info["bytes_per_micro_triangle"], 3),
                                            "reference_layout": round(info["reference_bytes_per_micro_triangle"], 3)},
               "results": result_mode, "bounce_rays": bounce_source,
               "parallelism": ("ONE %dx%d frame sharded over %d ranks by interleaved %dx%d tiles (tile t -> rank t mod N), scene replicated per GPU, "
                               "no data-path collective; with gather the kernels store their hit records straight into rank 0's frame buffer over NVLink" % (WIDTH, HEIGHT, world, TILE, TILE))
               if world > 1 else "single GPU (at --gpus N the frame is sharded by tiles: strong scaling)"}
        out["roofline"] = roofline
        if ms_one is not None:
            out["pipeline"] = {"what": "the two passes of a frame run on two streams: bounce(k) ordered after primary(k) by an event, primary(k+1) free to start "
                                       "while bounce(k) drains (the tail of a persistent launch = a few warps on their longest rays); --one-stream times the old order",
                               "one_stream": {"value": round(rays_per_step * K / (ms_one * 1e-3) / 1e6, 1), "ms_per_step": round(ms_one / K, 4)}}
        if value_with_gather is not None:
            out["value_with_gather"] = round(value_with_gather, 1)
            out["gather"] = gather_info
        if configs:
            out["configs"] = {"2": {"workload": out["config"]["workload"], "value": out["value"], "unit": UNIT, "see": "top-level keys"}, **configs}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, prim, bounce, n_prim, n_bounce)
    if dist:
        dist.barrier(group=host_group)
        dist.barrier()
        dist.destroy_process_group()
    return out


def parse_configs(spec, world):
    if spec == "none":
        return set()
    if spec == "all":
        return {1, 3, 4, 5} if world == 1 else {4, 5}
    return {int(v) for v in spec.split(",") if v}


# ------------------------------------------------------------------------------------------------------------------
def reference_pass(ref, prim, bounce, threads):
    a, b = prim.copy(), bounce.copy()  # fresh rays (the trace mutates them); the copy is not timed
    t0 = time.perf_counter()
    ref.intersect(a, threads)
    ref.intersect(b, threads)
    return time.perf_counter() - t0


def build_reference_scene(leaf, lib_path=None):
    from oracle import pyoracle as po
    if not po.have_reference():
        return None, po
    ref = po.RefScene(leaf, SUBDIV_LEVEL, COMP_LEVEL, lib_path=lib_path)
    ref.add_obj()
    ref.commit()
    return ref, po


def cpu_baseline(args, prim, bounce, n_prim, n_bounce):
    """The compiled reference (rtcIntersect1 per ray, all host threads) on the same rays: bounded sample."""
    ref, po = build_reference_scene(args.leaf)
    threads = po.hardware_threads()
    if ref is None:
        port = po.PortScene.from_image(np.fromfile(scene_path(args.leaf), np.uint8))
        ts = []
        for _ in range(3):
            a, b = prim.copy(), bounce.copy()
            t0 = time.perf_counter()
            port.intersect(a, threads)
            port.intersect(b, threads)
            ts.append(time.perf_counter() - t0)
        kind = "port"
    else:
        ts = [reference_pass(ref, prim, bounce, threads) for _ in range(4)][1:]
        kind = "reference"
    best = min(ts)
    out = {"value": round((n_prim + n_bounce) / best / 1e6, 2), "unit": UNIT, "cores": threads, "kind": kind,
           "sample": "the full step (%d primary + %d active bounce rays), best of %d passes, rtcIntersect1 per ray on %d std::threads, AVX2 build"
                     % (n_prim, n_bounce, len(ts), threads)}
    if ref is not None:
        # also one core, and the AVX512SKX build of the reference where the host can run it
        k = max(1, len(prim) // 16)
        a, b = prim[::16].copy(), bounce[::16].copy()
        t0 = time.perf_counter()
        ref.intersect(a, 1)
        ref.intersect(b, 1)
        dt = time.perf_counter() - t0
        act = int(k + (bounce[::16, 3] <= bounce[::16, 8]).sum())
        out["single_core"] = {"value": round(act / dt / 1e6, 2), "cores": 1, "sample": "every 16th ray of the step, AVX2 build"}
        ref.free()
        if po.have_avx512_reference():
            ref512, _ = build_reference_scene(args.leaf, po.REF_SO_AVX512)
            t512 = min(reference_pass(ref512, prim, bounce, threads) for _ in range(3))
            out["avx512"] = {"value": round((n_prim + n_bounce) / t512 / 1e6, 2), "cores": threads,
                             "sample": "the full step, best of 3 passes, AVX512SKX build of the reference (its own BVH4 commit)"}
            ref512.free()
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return None
    from embree_compressed_b200 import raygen as rg
    ref, po = build_reference_scene(args.leaf)
    threads = po.hardware_threads()
    if ref is None:
        # oracle/_ref did not travel: time the scalar C restatement instead
        image = np.fromfile(scene_path(args.leaf), np.uint8)
        tracer = po.PortScene.from_image(image)
        kind = "port"

        class _Port:
            def intersect(self, r, t=None):
                return tracer.intersect(r, t)
        ref = _Port()
        prim, bounce, _ = config2_rays(rg, args.leaf)
        if bounce is None:
            first = prim.copy()
            tracer.intersect(first, threads)
            h = po.rimg_header(image)
            diag = float(np.linalg.norm(h["bounds"][3:] - h["bounds"][:3]))
            xs, ys = np.meshgrid(np.arange(WIDTH, dtype=np.uint32), np.arange(HEIGHT, dtype=np.uint32))
            bounce = rg.diffuse_rays(first, -first[:, 4:7], diag, sample=0, pixel_xy=(xs.ravel(), ys.ravel()))
    else:
        kind = "reference"
        prim, bounce, _ = config2_rays(rg, args.leaf, reference_scene=ref)
    n_prim, n_bounce = len(prim), int((bounce[:, 3] <= bounce[:, 8]).sum())
    for _ in range(args.warmup):
        reference_pass(ref, prim, bounce, threads)
    dt = 0.0
    for _ in range(args.steps):
        dt += reference_pass(ref, prim, bounce, threads)
    value = (n_prim + n_bounce) * args.steps / dt / 1e6
    return {"impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.leaf, n_prim, n_bounce),
            "cpu_baseline": {"value": round(value, 2), "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": "every step traces the full workload (%d + %d rays) with rtcIntersect1 per ray on %d std::threads (AVX2 build, internal tasking)"
                                       % (n_prim, n_bounce, threads)},
            "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--leaf", default="leaf", choices=["leaf", "grid", "box"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--one-stream", action="store_true", help="time the two passes of every frame back to back on one stream (no frame pipeline)")
    ap.add_argument("--configs", default="all", help="which of the other BASELINE configs to measure into the `configs` object: all | none | "
                                                     "comma list of 1,3,4,5 (N > 1 measures the sharded ones, 4 and 5)")
    ap.add_argument("--e2e-memory", default="pinned", choices=["pinned", "pageable"],
                    help="caller memory of the e2e leg (the contract asks for pinned; pageable is what an embree application has)")
    ap.add_argument("--levels", default="6,3", help="S,C of the scene image (default: BASELINE configs[1]); other images come "
                                                    "from `python __graft_entry__.py --scene bomberman_s8c3_leaf`")
    args = ap.parse_args()
    global SUBDIV_LEVEL, COMP_LEVEL
    SUBDIV_LEVEL, COMP_LEVEL = (int(v) for v in args.levels.split(","))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    # stdout carries exactly ONE line, the JSON: libraries that print there (NCCL's version banner under NCCL_DEBUG=VERSION,
    # a compiler, ...) are sent to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            out = run_reference(args, rank, world)
        else:
            out = run_ours(args, rank, world)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if rank == 0 and out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
