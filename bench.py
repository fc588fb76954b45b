#!/usr/bin/env python3
"""bench.py -- Mrays/s of the compressed-BVH traversal hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--leaf leaf|grid|box]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): build/bomberman.obj as one subdivision mesh, S=6 C=3 compressed leaves,
camera of build/bomberman.ecs, 1920x1080 primary rays + one cosine-diffuse bounce ray per primary hit.
A step = one pass of the hot path over that batch: trace the primary rays, trace the bounce rays.

  value      device-timed: rays resident in HBM, CUDA events around the K steps on the launching stream.
  e2e        the same step through the reference-facing call (ecbIntersect1M == rtcIntersect1M) with pinned HOST
             ray buffers; host<->device copies inside the timed region.
  roofline   algorithmic bytes of one launch (measured by the instrumented kernel variant, SURVEY 8d formula with
             this repo's node/header sizes) / average launch duration, against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline  the compiled reference (oracle/_ref, rtcIntersect1 per ray on all host threads) on the same rays.

Scenes are images of the REFERENCE's committed BVH (built by its CPU builder at build() time, scenes/_built/*.rimg)
converted by the product's importer; nothing under oracle/ is touched outside the cpu_baseline / --impl reference legs.
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


def log(*a):
    print(*a, file=sys.stderr, flush=True)


SUBDIV_LEVEL, COMP_LEVEL = 6, 3   # BASELINE.json configs[1] (build/bomberman.ecs); --levels S,C selects another tessellation


def scene_path(leaf):
    return os.path.join(ROOT, "scenes", "_built", "bomberman_s%dc%d_%s.rimg" % (SUBDIV_LEVEL, COMP_LEVEL, leaf))


def workload_config(leaf, n_primary, n_bounce, scene_bytes=None):
    if scene_bytes is None or scene_bytes < 100e6:
        l2 = "ray buffers (2 x 166 MB per step, a fresh copy every step) exceed the 126 MB L2; the %s scene is L2-resident by " \
             "construction of the workload" % ("18.8 MB" if scene_bytes is None else "%.1f MB" % (scene_bytes / 1e6))
    else:
        l2 = "ray buffers (2 x 166 MB per step, a fresh copy every step) and the %.0f MB scene both exceed the 126 MB L2" % (scene_bytes / 1e6)
    return {"workload": "bomberman.obj subdiv S=%d C=%d compress.%s, %dx%d primary + 1 cosine-diffuse bounce"
                        % (SUBDIV_LEVEL, COMP_LEVEL, leaf, WIDTH, HEIGHT),
            "rays_primary": int(n_primary), "rays_bounce_active": int(n_bounce), "l2": l2}


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


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(leaf, key=None):
    """dram bytes per launch of the dominant kernel (or another recorded field) from the committed ncu summary."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key or leaf)
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------------------------
def make_rays(scene, ecb, rg, torch, dev):
    """Primary rays of the ecs camera and their bounce rays (block-frame normals; generated once, untimed)."""
    prim = rg.primary_rays(WIDTH, HEIGHT, **rg.BOMBERMAN_CAMERA)
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
    bounce = rg.diffuse_rays(first, normals, diag, sample=0, pixel_xy=(xs.ravel(), ys.ravel()))
    return prim, bounce, int(hit.sum())


def run_ours(args, rank, world):
    import torch
    import embree_compressed_b200 as ecb
    from embree_compressed_b200 import raygen as rg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the traversal path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    path = scene_path(args.leaf)
    if not os.path.exists(path):
        raise SystemExit("bench.py: %s missing -- run __graft_entry__.build() where /root/reference is mounted" % path)
    scene = ecb.Scene.from_image_file(path).upload(local)
    info = scene.info()
    prim, bounce, n_hits = make_rays(scene, ecb, rg, torch, dev)
    n_prim = len(prim)
    n_bounce = int((bounce[:, 3] <= bounce[:, 8]).sum())
    rays_per_step = n_prim + n_bounce
    K, W = args.steps, args.warmup
    stream = torch.cuda.current_stream().cuda_stream

    # ---- device-timed value: a fresh copy of the ray buffers for every step (nothing is cached between steps) ----
    d_prim0 = torch.from_numpy(prim).to(dev)
    d_bounce0 = torch.from_numpy(bounce).to(dev)
    bufs = [(d_prim0.clone(), d_bounce0.clone()) for _ in range(K + W)]

    def step(i):
        p, b = bufs[i]
        scene.intersect_device(p.data_ptr(), n_prim, 80, stream)
        scene.intersect_device(b.data_ptr(), len(bounce), 80, stream)

    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ecb.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for i in range(W, W + K):
        step(i)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = ecb.launch_count() - launches0
    if dist:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * rays_per_step * K / (ms_max * 1e-3) / 1e6

    # ---- per-kernel timing of the dominant launch (primary pass) + algorithmic bytes from the instrumented variant ----
    cnt = torch.zeros(6, dtype=torch.int64, device=dev)
    tmp = d_prim0.clone()
    scene.intersect_device_stats(tmp.data_ptr(), n_prim, cnt.data_ptr(), 80, stream)
    torch.cuda.synchronize()
    c_prim = cnt.cpu().numpy().astype(np.float64)
    cnt.zero_()
    tmp = d_bounce0.clone()
    scene.intersect_device_stats(tmp.data_ptr(), len(bounce), cnt.data_ptr(), 80, stream)
    torch.cuda.synchronize()
    c_bounce = cnt.cpu().numpy().astype(np.float64)
    leaf_bytes = {"leaf": 2.0, "grid": 48.0, "box": 0.0}[args.leaf]

    def algorithmic_bytes(c, n_records):
        # ray in + hit out + 112 B per top-level node + 144 B per block header + 4 B per quantized node
        # + leaf payload per cell (2 B heights / 4 x 12 B vertices / none); every record is read, hits are written
        return 80.0 * n_records + 48.0 * c[5] + 112.0 * c[1] + 144.0 * c[2] + 4.0 * c[3] + leaf_bytes * c[4]

    kernel_ms = {}
    for name, src, n in (("primary", d_prim0, n_prim), ("bounce", d_bounce0, len(bounce))):
        reps = []
        for _ in range(5):
            tmp = src.clone()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            scene.intersect_device(tmp.data_ptr(), n, 80, stream)
            b.record()
            torch.cuda.synchronize()
            reps.append(a.elapsed_time(b))
        kernel_ms[name] = float(np.mean(reps[1:]))
    bytes_prim = algorithmic_bytes(c_prim, n_prim)
    bytes_bounce = algorithmic_bytes(c_bounce, len(bounce))
    dom = "primary" if kernel_ms["primary"] >= kernel_ms["bounce"] else "bounce"
    dom_bytes = bytes_prim if dom == "primary" else bytes_bounce
    peak, peak_src = measured_peak()
    achieved = dom_bytes / (kernel_ms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": recorded_traffic(args.leaf) if (SUBDIV_LEVEL, COMP_LEVEL) == (6, 3) else None, "peak_source": peak_src, "kernel": "ecb_intersect_kernel<%s> (%s pass)" % (args.leaf, dom),
                "kernel_ms": {k: round(v, 4) for k, v in kernel_ms.items()},
                "algorithmic_bytes_per_ray": {"primary": round(bytes_prim / n_prim, 1), "bounce": round(bytes_bounce / max(1, n_bounce), 1)},
                "work_per_ray": {"primary": {"top_nodes": round(c_prim[1] / c_prim[0], 3), "blocks": round(c_prim[2] / c_prim[0], 3),
                                             "cnodes": round(c_prim[3] / c_prim[0], 3), "cells": round(c_prim[4] / c_prim[0], 3)},
                                 "bounce": {"top_nodes": round(c_bounce[1] / max(1, c_bounce[0]), 3), "blocks": round(c_bounce[2] / max(1, c_bounce[0]), 3),
                                            "cnodes": round(c_bounce[3] / max(1, c_bounce[0]), 3), "cells": round(c_bounce[4] / max(1, c_bounce[0]), 3)}},
                "mrays_per_s": {"primary": round(n_prim / kernel_ms["primary"] / 1e3, 1), "bounce": round(n_bounce / kernel_ms["bounce"] / 1e3, 1)}}
    # the memory roofline is not what limits this path (the scene is L2 resident); the informative figures are
    # rays per SM-cycle (live, at the SM clock sampled during the timed region) and the issue / L2 counters of the ncu capture
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    roofline["rays_per_sm_cycle_live"] = {k: round((n_prim if k == "primary" else n_bounce) / (kernel_ms[k] * 1e-3 * sm_mhz * 1e6 * sm_count), 5)
                                          for k in kernel_ms}
    if args.leaf == "leaf" and (SUBDIV_LEVEL, COMP_LEVEL) == (6, 3):
        roofline["ncu"] = recorded_traffic(args.leaf, "ncu_primary_pass")

    # ---- e2e: the reference-facing call with pinned host buffers, copies inside the timed region ----
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
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        scene.intersect1M(hp)
        scene.intersect1M(hb)
        e2e_t.append(time.perf_counter() - t0)
    e2e_s = float(np.mean(e2e_t[2:]))
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * rays_per_step / float(t.item()) / 1e6
    e2e_hits = int((hp.view(np.uint32)[:, 18] != 0xFFFFFFFF).sum())
    assert e2e_hits == n_hits, "e2e path found %d hits, device path %d" % (e2e_hits, n_hits)
    # the host-buffer call must leave exactly the bits the device-buffer call leaves
    p_dev, b_dev = bufs[W]
    assert (hp.view(np.uint32) == p_dev.cpu().numpy().view(np.uint32)).all(), "e2e primary results differ from the device path"
    assert (hb.view(np.uint32) == b_dev.cpu().numpy().view(np.uint32)).all(), "e2e bounce results differ from the device path"
    host_path = os.environ.get("ECB200_HOST_PATH") or ecb.host_path_for(args.e2e_memory == "pinned")
    n_rec = len(prim) + len(bounce)
    h2d = 32 * n_rec if host_path == "packed" else int(prim.nbytes + bounce.nbytes)
    d2h = int(prim.nbytes + bounce.nbytes) if host_path == "copy" else 20 * n_rec
    how = {"packed": "the library's host pool packs the 32 bytes per ray the path reads into pinned staging, 32-byte records go up, "
                     "20-byte compact hit records come back and are scattered into the caller's records by the same pool",
           "compact": "full records go up, 20-byte compact hit records come back and are scattered into the caller's records by the library's host pool",
           "copy": "full records copied both ways"}[host_path]
    e2e = {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
           "ms_per_step": round(float(t.item()) * 1e3, 3), "host_path": host_path, "caller_memory": args.e2e_memory,
           "call": "ecbIntersect1M (rtcIntersect1M drop-in), %s host RTCRayHit arrays (%d bytes), results written in place; %s"
                   % (args.e2e_memory, int(prim.nbytes + bounce.nbytes), how)}
    if hp_ptr is not None:
        ecb.host_free(hp_ptr)
        ecb.host_free(hb_ptr)

    # ---- multi-GPU: the one exchange step of the path, hit records gathered to rank 0 over NCCL (timed apart) ----
    gather = None
    if dist:
        p, b = bufs[W]
        compact = torch.cat([p[:, 8:9], p[:, 15:20]], 1).contiguous()  # tfar,u,v,primID,geomID,instID = 24 B per ray
        outs = [torch.empty_like(compact) for _ in range(world)] if rank == 0 else None
        for _ in range(2):
            dist.gather(compact, outs, dst=0)
        torch.cuda.synchronize()
        dist.barrier()
        a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            dist.gather(compact, outs, dst=0)
        bb.record()
        torch.cuda.synchronize()
        tg = torch.tensor([a.elapsed_time(bb) / 5], device=dev, dtype=torch.float64)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather = {"ms_per_pass": round(float(tg.item()), 3), "bytes_per_rank": int(compact.numel() * 4),
                  "what": "NCCL gather of 24-byte hit records of the primary pass to rank 0 (not inside `value`)"}

    out = None
    if rank == 0:
        out = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": round(ms_max / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": workload_config(args.leaf, n_prim, n_bounce, info["device_bytes"]),
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
               "bytes_per_micro_triangle": {"device_format": round(info["bytes_per_micro_triangle"], 3),
                                            "reference_layout": round(info["reference_bytes_per_micro_triangle"], 3)},
               "parallelism": "rays sharded per rank (each rank traces its own full frame), scene replicated, no data-path collective"}
        if gather:
            out["gather"] = gather
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, prim, bounce, n_prim, n_bounce)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return out


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
    prim = rg.primary_rays(WIDTH, HEIGHT, **rg.BOMBERMAN_CAMERA)
    xs, ys = np.meshgrid(np.arange(WIDTH, dtype=np.uint32), np.arange(HEIGHT, dtype=np.uint32))
    if ref is None:
        # oracle/_ref did not travel: time the scalar C restatement instead (bounce normals = reversed incoming ray)
        image = np.fromfile(scene_path(args.leaf), np.uint8)
        tracer = po.PortScene.from_image(image)
        kind = "port"
        first = prim.copy()
        tracer.intersect(first, threads)
        h = po.rimg_header(image)
        diag = float(np.linalg.norm(h["bounds"][3:] - h["bounds"][:3]))
        normals = -first[:, 4:7]

        class _Port:
            def intersect(self, r, t):
                return tracer.intersect(r, t)
        ref = _Port()
    else:
        kind = "reference"
        first = prim.copy()
        ref.intersect(first, threads)
        b = ref.bounds()
        diag = float(np.linalg.norm(b[3:] - b[:3]))
        normals = ref.hit_normals(first)
    bounce = rg.diffuse_rays(first, normals, diag, sample=0, pixel_xy=(xs.ravel(), ys.ravel()))
    n_prim, n_bounce = len(prim), int((bounce[:, 3] <= bounce[:, 8]).sum())
    for _ in range(args.warmup):
        reference_pass(ref, prim, bounce, threads)
    dt = 0.0
    for _ in range(args.steps):
        dt += reference_pass(ref, prim, bounce, threads)
    value = (n_prim + n_bounce) * args.steps / dt / 1e6
    return {"impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
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
