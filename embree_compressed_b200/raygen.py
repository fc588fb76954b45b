"""Deterministic synthetic ray sets for the parity tests and the benchmark (numpy, host side).

The generators restate the ray sets the reference's own viewer traces:
  * primary rays of the tutorial pinhole camera
    (tutorials/common/tutorial/camera.h:75-88, tutorials/viewer/viewer_device.cpp:250-262);
  * cosine-weighted diffuse bounce rays (tutorials/common/math/sampling.h:48-54) driven by the
    tutorial's per-pixel sampler = MurmurHash3 mix + LCG (tutorials/common/math/random_sampler.h:28-100);
  * the authors' "incoherent" generator: origin and target uniform in the scene AABB
    (tutorials/viewer/viewer_device.cpp:367-392), seeded here instead of drand48.
Rays are float32 arrays (N, 20) laid out as RTCRayHit (include/embree3/rtcore_ray.h:26-64).
"""
import numpy as np

INVALID = 0xFFFFFFFF
ORG, TNEAR, DIR, TIME, TFAR, MASK, ID, FLAGS, NG, U, V, PRIMID, GEOMID, INSTID = (
    slice(0, 3), 3, slice(4, 7), 7, 8, 9, 10, 11, slice(12, 15), 15, 16, 17, 18, 19)

# camera of build/bomberman.ecs:2-6 and of tutorials/displacement_geometry/displacement_geometry.cpp:28-29
BOMBERMAN_CAMERA = dict(vp=(18.21240425, 20.05745888, 15.46878433), vi=(0.0, 0.0, 0.0), vu=(0.0, 1.0, 0.0), fov=90.0)
CUBE_CAMERA = dict(vp=(1.5, 1.5, -1.5), vi=(0.0, 0.0, 0.0), vu=(0.0, 1.0, 0.0), fov=90.0)


def _normalize(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def new_rayhits(n):
    r = np.zeros((n, 20), np.float32)
    ri = r.view(np.uint32)
    r[:, TFAR] = np.inf
    ri[:, MASK] = INVALID
    ri[:, PRIMID] = INVALID
    ri[:, GEOMID] = INVALID
    ri[:, INSTID] = INVALID
    return r


def camera_frame(width, height, vp, vi, vu, fov):
    vp, vi, vu = (np.asarray(a, np.float64) for a in (vp, vi, vu))
    Z = _normalize(vi - vp)
    U = _normalize(np.cross(vu, Z))
    V = _normalize(np.cross(Z, U))
    fov_scale = 1.0 / np.tan(np.deg2rad(0.5 * fov))
    vx = U
    vy = -V
    vz = -0.5 * width * U + 0.5 * height * V + 0.5 * height * fov_scale * Z
    return vx, vy, vz, vp


def primary_rays(width, height, vp, vi, vu, fov, x0=0, y0=0, w=None, h=None):
    """One ray per pixel of the window [x0,x0+w) x [y0,y0+h) of a width x height frame, row-major."""
    w = width if w is None else w
    h = height if h is None else h
    vx, vy, vz, p = camera_frame(width, height, vp, vi, vu, fov)
    xs, ys = np.meshgrid(np.arange(x0, x0 + w, dtype=np.float64), np.arange(y0, y0 + h, dtype=np.float64))
    d = xs[..., None] * vx + ys[..., None] * vy + vz
    d = _normalize(d).reshape(-1, 3)
    r = new_rayhits(w * h)
    r[:, ORG] = p.astype(np.float32)
    r[:, DIR] = d.astype(np.float32)
    return r


# ---- tutorial sampler ------------------------------------------------------------------------------------
def _rotl(x, r):
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))).astype(np.uint32)


def _murmur_mix(h, k):
    with np.errstate(over="ignore"):
        k = (k * np.uint32(0xCC9E2D51)).astype(np.uint32)
        k = _rotl(k, 15)
        k = (k * np.uint32(0x1B873593)).astype(np.uint32)
        h = (h ^ k).astype(np.uint32)
        h = (_rotl(h, 13) * np.uint32(5) + np.uint32(0xE6546B64)).astype(np.uint32)
    return h


def _murmur_final(h):
    with np.errstate(over="ignore"):
        h = h ^ (h >> np.uint32(16))
        h = (h * np.uint32(0x85EBCA6B)).astype(np.uint32)
        h = h ^ (h >> np.uint32(13))
        h = (h * np.uint32(0xC2B2AE35)).astype(np.uint32)
        h = h ^ (h >> np.uint32(16))
    return h.astype(np.uint32)


def sampler_init(x, y, sample):
    pid = (np.asarray(x, np.uint32) | (np.asarray(y, np.uint32) << np.uint32(16))).astype(np.uint32)
    h = _murmur_mix(np.zeros_like(pid), pid)
    h = _murmur_mix(h, np.full_like(pid, sample, dtype=np.uint32))
    return _murmur_final(h)


def sampler_float(state):
    with np.errstate(over="ignore"):
        state = (state * np.uint32(1664525) + np.uint32(1013904223)).astype(np.uint32)
    return state, ((state >> np.uint32(1)).astype(np.float32) * np.float32(4.656612873077392578125e-10))


def _frame(n):
    dx0 = np.cross(np.array([1.0, 0.0, 0.0]), n)
    dx1 = np.cross(np.array([0.0, 1.0, 0.0]), n)
    pick = (np.sum(dx0 * dx0, -1) > np.sum(dx1 * dx1, -1))[:, None]
    dx = _normalize(np.where(pick, dx0, dx1))
    dy = _normalize(np.cross(n, dx))
    return dx, dy


def diffuse_rays(rayhits, normals, scene_diag, sample=0, pixel_xy=None, samples_per_hit=1):
    """Cosine-weighted bounce rays from every hit of `rayhits` (misses produce inactive rays tnear>tfar).

    normals: (N,3) shading normals (any length); face-forwarded against the incoming direction here, as
    tutorials/viewer/viewer_device.cpp:295-296 does.  Sample index k uses sampler_init(x, y, sample+k).
    """
    n_in = rayhits.shape[0]
    ri = rayhits.view(np.uint32)
    hit = ri[:, GEOMID] != INVALID
    nlen = np.linalg.norm(normals, axis=-1)
    hit &= nlen > 0
    if pixel_xy is None:
        idx = np.arange(n_in, dtype=np.uint32)
        px, py = idx & np.uint32(0xFFFF), (idx >> np.uint32(16)) & np.uint32(0xFFFF)
    else:
        px, py = pixel_xy
    out = []
    d_in = rayhits[:, DIR].astype(np.float64)
    P = rayhits[:, ORG].astype(np.float64) + rayhits[:, TFAR:TFAR + 1].astype(np.float64) * d_in
    n = np.where(hit[:, None], normals.astype(np.float64), np.array([0.0, 0.0, 1.0]))
    n = _normalize(n)
    flip = np.sum(n * d_in, -1) > 0
    n = np.where(flip[:, None], -n, n)
    dx, dy = _frame(n)
    for k in range(samples_per_hit):
        st = sampler_init(px, py, sample + k)
        st, s1 = sampler_float(st)
        st, s2 = sampler_float(st)
        phi = 2.0 * np.pi * s1.astype(np.float64)
        cos_t = np.sqrt(s2.astype(np.float64))
        sin_t = np.sqrt(1.0 - s2.astype(np.float64))
        local = np.stack([np.cos(phi) * sin_t, np.sin(phi) * sin_t, cos_t], -1)
        d = local[:, 0:1] * dx + local[:, 1:2] * dy + local[:, 2:3] * n
        d = _normalize(d)
        r = new_rayhits(n_in)
        r[:, ORG] = P.astype(np.float32)
        r[:, DIR] = d.astype(np.float32)
        r[:, TNEAR] = np.float32(1e-3 * scene_diag)
        r[~hit, TNEAR] = np.float32(1.0)
        r[~hit, TFAR] = np.float32(0.0)  # inactive (tnear > tfar)
        r[~hit, ORG] = 0
        r[~hit, DIR] = np.array([0, 0, 1], np.float32)
        out.append(r)
    return np.concatenate(out, 0) if len(out) > 1 else out[0]


def incoherent_rays(n, bounds, seed=1):
    """The authors' incoherent benchmark rays: uniform point -> uniform point inside the scene AABB."""
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(bounds[:3], np.float64), np.asarray(bounds[3:], np.float64)
    p1 = lo + rng.random((n, 3)) * (hi - lo)
    p2 = lo + rng.random((n, 3)) * (hi - lo)
    r = new_rayhits(n)
    r[:, ORG] = p1.astype(np.float32)
    r[:, DIR] = _normalize(p2 - p1).astype(np.float32)
    return r


def rays_only(rayhits):
    """RTCRay (48-byte) view copy of an RTCRayHit array, for the occluded entry points."""
    return np.ascontiguousarray(rayhits[:, :12])


def shadow_rays(rayhits, light_center, light_radius, seed=3):
    """One shadow ray per hit toward a uniformly sampled point on a sphere light (BASELINE config 3): RTCRay records
    (N,12) with tnear = 1e-3, tfar = distance to the sample; misses become inactive rays."""
    n = rayhits.shape[0]
    ri = rayhits.view(np.uint32)
    hit = ri[:, GEOMID] != INVALID
    rng = np.random.default_rng(seed)
    z = 2.0 * rng.random(n) - 1.0
    phi = 2.0 * np.pi * rng.random(n)
    s = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    target = np.asarray(light_center, np.float64) + light_radius * np.stack([s * np.cos(phi), s * np.sin(phi), z], -1)
    P = rayhits[:, ORG].astype(np.float64) + rayhits[:, TFAR:TFAR + 1].astype(np.float64) * rayhits[:, DIR].astype(np.float64)
    P = np.where(hit[:, None], P, 0.0)
    d = target - P
    dist = np.linalg.norm(d, axis=-1)
    r = np.zeros((n, 12), np.float32)
    r[:, ORG] = P.astype(np.float32)
    r[:, DIR] = (d / np.maximum(dist, 1e-30)[:, None]).astype(np.float32)
    r[:, TNEAR] = np.float32(1e-3)
    r[:, TFAR] = dist.astype(np.float32)
    r.view(np.uint32)[:, MASK] = INVALID
    r[~hit, TNEAR] = 1.0
    r[~hit, TFAR] = 0.0
    return r
