"""Developer probe: how much of the BASELINE config 2 step is the TAIL of the two persistent launches (the last warps finishing
their long rays while the rest of the device idles)?  Times K frames of (primary pass, bounce pass)
  (a) on one stream, back to back (what bench.py times), and
  (b) pipelined over two streams: bounce(k) is ordered after primary(k) by an event, primary(k+1) may start while bounce(k)
      drains -- what a double-buffered renderer would do.

    python tools/overlap_probe.py [steps]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg
import bench

K = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda", 0)
scene = ecb.Scene.from_image_file(os.path.join(ROOT, "scenes", "_built", "bomberman_s6c3_leaf.rimg")).upload(0)
prim, bounce, _ = bench.config2_rays(rg, "leaf")
n = len(prim)
n_active = n + int((bounce[:, 3] <= bounce[:, 8]).sum())
dp, db = torch.from_numpy(prim).to(dev), torch.from_numpy(bounce).to(dev)
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()


def one_stream(bufs):
    for p, b in bufs:
        scene.intersect_device(p.data_ptr(), n, 80, s0.cuda_stream)
        scene.intersect_device(b.data_ptr(), n, 80, s0.cuda_stream)


def two_streams(bufs):
    for p, b in bufs:
        scene.intersect_device(p.data_ptr(), n, 80, s0.cuda_stream)
        e = torch.cuda.Event()
        e.record(s0)
        s1.wait_event(e)
        scene.intersect_device(b.data_ptr(), n, 80, s1.cuda_stream)


def slice_chains(n_slices, pair):
    """The frame cut into n_slices contiguous ray ranges, slice j on its own stream(s): primary(k, j) -> bounce(k, j) ->
    primary(k+1, j) in order (pair=False), or the two passes of a slice on two streams as in two_streams (pair=True)."""
    streams = [(torch.cuda.Stream(), torch.cuda.Stream()) for _ in range(n_slices)]
    cuts = [n * j // n_slices for j in range(n_slices + 1)]

    def run(bufs):
        for p, b in bufs:
            for j, (sa, sb) in enumerate(streams):
                lo, m = cuts[j], cuts[j + 1] - cuts[j]
                scene.intersect_device(p.data_ptr() + 80 * lo, m, 80, sa.cuda_stream)
                if pair:
                    e = torch.cuda.Event()
                    e.record(sa)
                    sb.wait_event(e)
                    scene.intersect_device(b.data_ptr() + 80 * lo, m, 80, sb.cuda_stream)
                else:
                    scene.intersect_device(b.data_ptr() + 80 * lo, m, 80, sa.cuda_stream)
    run.streams = [s for pr in streams for s in pr]
    return run


one_stream.streams = [s0]
two_streams.streams = [s0, s1]
modes = [("one stream", one_stream), ("two streams", two_streams), ("2 slice chains", slice_chains(2, False)), ("4 slice chains", slice_chains(4, False)),
         ("8 slice chains", slice_chains(8, False)), ("2 slices x 2 streams", slice_chains(2, True)), ("4 slices x 2 streams", slice_chains(4, True))]
for rep in range(2):
    for name, fn in modes:
        bufs = [(dp.clone(), db.clone()) for _ in range(K + 3)]
        fn(bufs[:3])
        torch.cuda.synchronize()
        a, e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s0)
        for st in fn.streams:
            if st is not s0:
                st.wait_event(a)
        fn(bufs[3:])
        ends = []
        for st in fn.streams:
            e = torch.cuda.Event(enable_timing=True)
            e.record(st)
            ends.append(e)
        torch.cuda.synchronize()
        ms = max(a.elapsed_time(e) for e in ends) / K
        ok = all(bool((p.view(torch.int32) == bufs[0][0].view(torch.int32)).all()) and bool((b.view(torch.int32) == bufs[0][1].view(torch.int32)).all()) for p, b in bufs[3::13])
        print("%-22s %.4f ms per frame  %6.0f Mrays/s  %s" % (name, ms, n_active / ms / 1e3, "same bits" if ok else "DIFFERENT"), flush=True)
        del bufs
