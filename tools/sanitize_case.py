"""Small end-to-end case that launches every kernel of the library once or more: closest hit x 3 leaf kinds x both traversal
kernels (one lane per ray / four lanes per ray), any-hit, exact any-hit, the triangle pass of a mixed scene, the instancing
kernel and the block-construction phases -- through host-buffer and device-buffer entry points.

Run it against a -DECB_BOUNDS_CHECK build (ECB200_LIB=embree_compressed_b200/libecb200_check.so; tests/test_bounds_check_gpu.py
does) to have every index the kernels derive from scene data checked, or under compute-sanitizer where that tool is open.
Prints `bounds check: enabled=<0|1> violations=[...]` and exits non-zero on any violation or result mismatch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg

G = os.path.join(ROOT, "tests", "golden")
cam = dict(vp=(1.6, 1.45, -1.35), vi=(0.05, -0.02, 0.03), vu=(0.0, 1.0, 0.0), fov=90.0)
failures = 0
for kernel in ("lane", "quad"):
    ecb.set_traversal_kernel(kernel)
    for name in ("cube_s4c2_leaf", "cube_s4c2_grid", "cube_s4c2_box", "cube_s5c3_leaf", "cube_s4c1_grid", "cubeplane_s4c2_leaf", "cubeplane_s4c2_grid"):
        s = ecb.Scene.from_image(np.fromfile(os.path.join(G, name + ".rimg"), np.uint8))
        g = np.load(os.path.join(G, name + ".npz"))
        rays = rg.primary_rays(int(g["size"][0]), int(g["size"][1]), **cam)
        r = rays.copy()
        s.intersect1M(r)
        o = rg.rays_only(rays)
        s.occluded1M(o)
        x = rg.rays_only(rays)
        s.occluded_exact1M(x)
        ok = (r[:, [8, 15, 16, 17, 18, 19]].view(np.uint32) == g["port_primary"].view(np.uint32)).all()
        failures += 0 if ok else 1
        print(kernel, name, "hits", int((r.view(np.uint32)[:, 18] != 0xFFFFFFFF).sum()), "bit-exact vs golden port:", bool(ok), flush=True)
        s.release()
ecb.set_traversal_kernel(None)
# the instancing kernel on the benchmark's instanced scene (a small frame), when its blob travelled
blob = os.path.join(ROOT, "scenes", "_built", "config4_instanced_grid.ecb")
if os.path.exists(blob):
    inst = ecb.Scene.load(blob)
    rays = rg.primary_rays(320, 180, vp=(90.0, 70.0, -100.0), vi=(0.0, 0.0, 0.0), vu=(0.0, 1.0, 0.0), fov=60.0)
    r = rays.copy()
    inst.intersect1M(r)
    o = rg.rays_only(rays)
    inst.occluded1M(o)
    print("instanced: hits", int((r.view(np.uint32)[:, 18] != 0xFFFFFFFF).sum()), "occluded", int(np.isneginf(o[:, 8]).sum()), flush=True)
    inst.release()
# block construction: rebuild the blocks of a golden scene from decoded corner data is covered by tests/test_build_gpu.py; here
# one launch of every build phase on synthetic grids
C_ = 2
W = (1 << C_) + 1
n = 64
rng = np.random.default_rng(5)
gx, gy = np.meshgrid(np.linspace(0, 1, W), np.linspace(0, 1, W))
grids = np.stack([np.stack([gx + k, gy, 0.05 * rng.random((W, W))], -1) for k in range(n)]).astype(np.float32).reshape(n, W * W, 3)
eye = np.tile(np.eye(3, dtype=np.float32).ravel(), (n, 1))
uv = np.tile(np.array([0, 0, 1, 1], np.float32), (n, 1))
for kind in ("leaf", "grid", "box"):
    blocks, bounds, ok = ecb.build_blocks(kind, C_, grids, eye, eye, uv, np.zeros(n, np.uint32), np.arange(n, dtype=np.uint32))
    print("build", kind, blocks.shape, "finite bounds:", bool(np.isfinite(bounds).all()), flush=True)
enabled, counts = ecb.bounds_check_failures()
print("bounds check: enabled=%d violations=%s" % (int(enabled), counts))
print("launches", ecb.launch_count())
sys.exit(1 if (failures or any(counts)) else 0)
