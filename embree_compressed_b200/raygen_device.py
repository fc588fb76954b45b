"""Synthetic ray sets generated ON the device with torch (benchmark plumbing for the large configurations: an 8K frame
is 33 M rays, BASELINE config 5 is 64 Mi -- too many to build with numpy on the host inside a benchmark run).

Same definitions as raygen.py (the numpy generators the parity tests use): tutorial pinhole camera, cosine-weighted
hemisphere sampling driven by the tutorial's MurmurHash3 + LCG sampler, the authors' AABB point-to-point generator.
All functions return float32 tensors shaped (N, 20) = RTCRayHit[N] (or (N, 12) = RTCRay[N]) on the given device.
"""
import math

import numpy as np
import torch

from . import raygen as rg

M32 = 0xFFFFFFFF
INVALID_F32 = float(np.array([0xFFFFFFFF], np.uint32).view(np.float32)[0])  # NaN bit pattern of RTC_INVALID_GEOMETRY_ID


def new_rayhits(n, device):
    r = torch.zeros((n, 20), dtype=torch.float32, device=device)
    r[:, rg.TFAR] = float("inf")
    ri = r.view(torch.int32)
    for c in (rg.MASK, rg.PRIMID, rg.GEOMID, rg.INSTID):
        ri[:, c] = -1
    return r


def primary_rays(width, height, vp, vi, vu, fov, device, pixel_index=None):
    """One ray per pixel (row-major), or for the flat pixel indices `pixel_index` (int64 tensor: y * width + x)."""
    vx, vy, vz, p = rg.camera_frame(width, height, vp, vi, vu, fov)
    f64 = dict(dtype=torch.float64, device=device)
    if pixel_index is None:
        pixel_index = torch.arange(width * height, dtype=torch.int64, device=device)
    xs = (pixel_index % width).to(torch.float64)[:, None]
    ys = (pixel_index // width).to(torch.float64)[:, None]
    d = xs * torch.tensor(vx, **f64) + ys * torch.tensor(vy, **f64) + torch.tensor(vz, **f64)
    d = d / torch.linalg.norm(d, dim=1, keepdim=True)
    r = new_rayhits(pixel_index.numel(), device)
    r[:, 0:3] = torch.tensor(p, **f64).to(torch.float32)
    r[:, 4:7] = d.to(torch.float32)
    return r


# ---- tutorial sampler: MurmurHash3 mix + LCG (tutorials/common/math/random_sampler.h:28-100), 32-bit arithmetic in int64 ----
def _rotl(x, r):
    return ((x << r) | (x >> (32 - r))) & M32


def _mix(h, k):
    k = (k * 0xCC9E2D51) & M32
    k = _rotl(k, 15)
    k = (k * 0x1B873593) & M32
    h = h ^ k
    return (_rotl(h, 13) * 5 + 0xE6546B64) & M32


def _final(h):
    h = h ^ (h >> 16)
    h = (h * 0x85EBCA6B) & M32
    h = h ^ (h >> 13)
    h = (h * 0xC2B2AE35) & M32
    return h ^ (h >> 16)


def sampler_init(px, py, sample):
    pid = (px | (py << 16)) & M32
    h = _mix(torch.zeros_like(pid), pid)
    h = _mix(h, sample & M32)
    return _final(h)


def sampler_float(state):
    state = (state * 1664525 + 1013904223) & M32
    return state, (state >> 1).to(torch.float32) * 4.656612873077392578125e-10


def _normalize(v):
    return v / torch.linalg.norm(v, dim=1, keepdim=True)


def diffuse_rays(P, n, d_in, px, py, sample, tnear):
    """Cosine-weighted bounce rays (tutorials/common/math/sampling.h:48-54) leaving the points P (N,3) about the normals n
    (N,3; face-forwarded against the incoming directions d_in here, viewer_device.cpp:295-296).  px, py, sample: int64
    tensors (N,) feeding RandomSampler_init(x, y, sample).  All rays are active."""
    dev = P.device
    n = _normalize(n.to(torch.float32))
    flip = (n * d_in).sum(1) > 0
    n = torch.where(flip[:, None], -n, n)
    ex = torch.tensor([1.0, 0.0, 0.0], device=dev).expand_as(n)
    ey = torch.tensor([0.0, 1.0, 0.0], device=dev).expand_as(n)
    dx0 = torch.linalg.cross(ex, n)
    dx1 = torch.linalg.cross(ey, n)
    pick = ((dx0 * dx0).sum(1) > (dx1 * dx1).sum(1))[:, None]
    dx = _normalize(torch.where(pick, dx0, dx1))
    dy = _normalize(torch.linalg.cross(n, dx))
    st = sampler_init(px, py, sample)
    st, s1 = sampler_float(st)
    st, s2 = sampler_float(st)
    phi = (2.0 * math.pi) * s1
    cos_t = torch.sqrt(s2)
    sin_t = torch.sqrt(1.0 - s2)
    d = (torch.cos(phi) * sin_t)[:, None] * dx + (torch.sin(phi) * sin_t)[:, None] * dy + cos_t[:, None] * n
    d = _normalize(d)
    r = new_rayhits(P.shape[0], dev)
    r[:, 0:3] = P
    r[:, 4:7] = d
    r[:, rg.TNEAR] = float(tnear)
    return r


def incoherent_rays(n, bounds, device, seed=1):
    """The authors' incoherent benchmark rays (tutorials/viewer/viewer_device.cpp:367-392): uniform point -> uniform
    point inside the scene AABB, seeded."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lo = torch.tensor(bounds[:3], dtype=torch.float32, device=device)
    hi = torch.tensor(bounds[3:], dtype=torch.float32, device=device)
    p1 = lo + torch.rand((n, 3), generator=g, device=device) * (hi - lo)
    p2 = lo + torch.rand((n, 3), generator=g, device=device) * (hi - lo)
    r = new_rayhits(n, device)
    r[:, 0:3] = p1
    r[:, 4:7] = _normalize(p2 - p1)
    return r


def shadow_rays(rayhits, light_center, light_radius, seed=3):
    """One shadow ray per hit of `rayhits` toward a uniformly sampled point on a sphere light (BASELINE config 3): RTCRay
    records (N, 12), tnear = 1e-3, tfar = distance to the sample; misses become inactive rays (tnear > tfar)."""
    dev = rayhits.device
    n = rayhits.shape[0]
    hit = rayhits.view(torch.int32)[:, rg.GEOMID] != -1
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    z = 2.0 * torch.rand(n, generator=g, device=dev) - 1.0
    phi = (2.0 * math.pi) * torch.rand(n, generator=g, device=dev)
    s = torch.sqrt(torch.clamp(1.0 - z * z, min=0.0))
    target = torch.tensor(light_center, dtype=torch.float32, device=dev) + light_radius * torch.stack([s * torch.cos(phi), s * torch.sin(phi), z], 1)
    t = torch.where(hit, rayhits[:, rg.TFAR], torch.zeros_like(rayhits[:, rg.TFAR]))
    P = torch.where(hit[:, None], rayhits[:, 0:3] + t[:, None] * rayhits[:, 4:7], torch.zeros_like(rayhits[:, 0:3]))
    d = target - P
    dist = torch.linalg.norm(d, dim=1)
    r = torch.zeros((n, 12), dtype=torch.float32, device=dev)
    r[:, 0:3] = P
    r[:, 4:7] = d / torch.clamp(dist, min=1e-30)[:, None]
    r[:, rg.TNEAR] = torch.where(hit, torch.full_like(dist, 1e-3), torch.ones_like(dist))
    r[:, rg.TFAR] = torch.where(hit, dist, torch.zeros_like(dist))
    r.view(torch.int32)[:, rg.MASK] = -1
    return r
