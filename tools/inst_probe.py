"""Developer probe: device time of the instancing kernel on every 4th pixel of the 8K frame of BASELINE config 4."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, zlib
import embree_compressed_b200 as ecb
from embree_compressed_b200 import raygen as rg, raygen_device as rgd
import __graft_entry__ as ge
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream().cuda_stream
sc = ecb.Scene.load(os.path.join(ROOT, "scenes", "_built", ge.CONFIG4_BLOB)).upload(0)
ys, xs = torch.meshgrid(torch.arange(0, 4320, 2, device=dev), torch.arange(0, 7680, 2, device=dev), indexing="ij")
rays = rgd.primary_rays(7680, 4320, device=dev, pixel_index=(ys * 7680 + xs).reshape(-1).to(torch.int64), **ge.CONFIG4_CAMERA)
ts = []
for _ in range(6):
    tmp = rays.clone(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); sc.intersect_device(tmp.data_ptr(), tmp.shape[0], 80, stream); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = float(np.median(ts[1:]))
print("%s: %d rays %.3f ms %.0f Mrays/s crc=%08x" % (os.path.basename(ecb.LIB_PATH), rays.shape[0], ms, rays.shape[0] / ms / 1e3, zlib.crc32(tmp.cpu().numpy().tobytes())), flush=True)
