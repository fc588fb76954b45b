"""Developer tool: throughput of the block-construction kernel (ecbBuildBlocks) on the benchmark asset.

    python tools/build_bench.py [grid|leaf|box]

Inputs = the displaced vertex grids and frames of every block of scenes/_built/bomberman_s6c3_grid.rimg (46 528
blocks); frames of the requested kind come from its own image.  Prints the end-to-end call time (host buffers, copies
included) and, for comparison, the scalar oracle on all host threads.  Run under `ncu --metrics gpu__time_duration.sum
-k regex:ecb_build` for the kernel time alone."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import embree_compressed_b200 as ecb
import build_util as bu

leaf = sys.argv[1] if len(sys.argv) > 1 else "leaf"
kind = {"grid": 0, "leaf": 1, "box": 2}[leaf]
img = lambda l: np.fromfile(os.path.join(ROOT, "scenes", "_built", "bomberman_s6c3_%s.rimg" % l), np.uint8)
t0 = time.perf_counter()
inp, blocks = bu.inputs_from_images(img("grid"), img(leaf))
print("parsed %d blocks in %.1f s" % (len(blocks), time.perf_counter() - t0), flush=True)
args = (leaf, 3, inp["grids"], inp["space"], inp["world"], inp["uv"], inp["geomIDs"], inp["primIDs"])
ecb.build_blocks(*args)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); out, bounds, ok = ecb.build_blocks(*args); ts.append(time.perf_counter() - t0)
n = len(blocks)
print("ecbBuildBlocks %s: %d blocks, %.2f ms per call (host buffers, %.1f MB in, %.1f MB out) = %.1f M blocks/s, %.0f M micro-triangles/s"
      % (leaf, n, 1e3 * min(ts), inp["grids"].nbytes / 1e6, out.nbytes / 1e6, n / min(ts) / 1e6, 128 * n / min(ts) / 1e6), flush=True)
t0 = time.perf_counter()
bu.port_build(inp["grids"], inp["space"], inp["world"], kind, 3, threads=os.cpu_count())
print("scalar oracle on %d threads: %.1f ms (includes the ctypes marshalling of the inputs)" % (os.cpu_count(), 1e3 * (time.perf_counter() - t0)))
