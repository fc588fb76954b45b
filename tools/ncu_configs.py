"""One launch of the dominant kernel of every BASELINE config, for `ncu --set full` (profiles/r2_*): warm-up launches first,
then, in this order, the launches to capture:
  0 c2 primary   1 c2 bounce   2 c1 triangle pass   3 c1 compressed pass   4 c3 primary   5 c3 shadow (any-hit)
  6 c4 instanced (1920x1080 crop of the 8K frame: every 4th pixel in x and y)   7 c5 diffuse batch (8 Mi rays)
Run:  python tools/ncu_configs.py            (prints CUDA-event times; the same command line goes under ncu with -k regex:ecb_ -s <warm-ups> -c 8)
      --s8 adds bomberman S8/C3 (needs scenes/_built/bomberman_s8c3_leaf.rimg): launches 8 (primary) and 9 (bounce)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg, raygen_device as rgd
import bench, __graft_entry__ as ge

dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
S = os.path.join(ROOT, "scenes", "_built")
jobs = []   # (name, fn(tensor), rays tensor)

sc2 = ecb.Scene.from_image_file(os.path.join(S, "bomberman_s6c3_leaf.rimg")).upload(0)
prim, bounce, _ = bench.config2_rays(rg, "leaf")
it = lambda sc: (lambda t: sc.intersect_device(t.data_ptr(), t.shape[0], 80, stream))
jobs.append(("c2 primary", it(sc2), torch.from_numpy(prim).to(dev)))
jobs.append(("c2 bounce", it(sc2), torch.from_numpy(bounce).to(dev)))
sc1 = ecb.Scene.from_image_file(os.path.join(S, "cubeplane_s6c4_leaf.rimg")).upload(0)
jobs.append(("c1 primary (triangle pass + compressed pass)", it(sc1), rgd.primary_rays(1024, 1024, device=dev, **rg.CUBE_CAMERA)))
sc3 = ecb.Scene.from_image_file(os.path.join(S, "cube_s8c4_leaf.rimg")).upload(0)
r3 = rgd.primary_rays(3840, 2160, device=dev, **rg.CUBE_CAMERA)
jobs.append(("c3 primary", it(sc3), r3))
tmp = r3.clone(); sc3.intersect_device(tmp.data_ptr(), tmp.shape[0], 80, stream); torch.cuda.synchronize()
shadow = rgd.shadow_rays(tmp, light_center=(3.0, 5.0, -2.0), light_radius=1.0)
jobs.append(("c3 shadow", lambda t: sc3.occluded_device(t.data_ptr(), t.shape[0], 48, stream), shadow))
sc4 = ecb.Scene.load(os.path.join(S, ge.CONFIG4_BLOB)).upload(0)
ys, xs = torch.meshgrid(torch.arange(0, 4320, 4, device=dev), torch.arange(0, 7680, 4, device=dev), indexing="ij")
r4 = rgd.primary_rays(7680, 4320, device=dev, pixel_index=(ys * 7680 + xs).reshape(-1).to(torch.int64), **ge.CONFIG4_CAMERA)
jobs.append(("c4 instanced (every 4th pixel of the 8K frame)", it(sc4), r4))
z = np.load(os.path.join(S, ge.CONFIG2_HITS))
hit_pix = np.nonzero(np.isfinite(z["t"]))[0]
pa = prim
P = torch.from_numpy((pa[hit_pix, 0:3].astype(np.float64) + z["t"][hit_pix, None].astype(np.float64) * pa[hit_pix, 4:7]).astype(np.float32)).to(dev)
N = torch.from_numpy(z["normal_i16"][hit_pix].astype(np.float32) / 32767.0).to(dev)
D = torch.from_numpy(pa[hit_pix, 4:7]).to(dev)
pix = torch.from_numpy(hit_pix).to(dev)
i = torch.arange(0, 1 << 23, dtype=torch.int64, device=dev)
k = i % len(hit_pix)
diag = float(np.linalg.norm(z["bounds"][3:].astype(np.float64) - z["bounds"][:3].astype(np.float64)))
r5 = rgd.diffuse_rays(P[k], N[k], D[k], pix[k] % 1920, pix[k] // 1920, i // len(hit_pix), 1e-3 * diag)
jobs.append(("c5 diffuse (8 Mi of the 64 Mi batch)", it(sc2), r5))
if "--s8" in sys.argv:
    sc8 = ecb.Scene.from_image_file(os.path.join(S, "bomberman_s8c3_leaf.rimg")).upload(0)
    jobs.append(("s8 primary", it(sc8), torch.from_numpy(prim).to(dev)))
    jobs.append(("s8 bounce", it(sc8), torch.from_numpy(bounce).to(dev)))

copies = [(n, f, r, [r.clone() for _ in range(3)]) for n, f, r in jobs]
torch.cuda.synchronize()
for rep in range(3):          # reps 0 and 1 warm up, rep 2 is the one to capture
    for n, f, r, cs in copies:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(cs[rep]); b.record(); torch.cuda.synchronize()
        if rep == 2:
            act = int((r[:, 3] <= r[:, 8]).sum().item())
            print("%-48s %9d active rays  %8.4f ms  %8.1f Mrays/s" % (n, act, a.elapsed_time(b), act / a.elapsed_time(b) / 1e3), flush=True)
print("launches", ecb.launch_count())
