"""embree_compressed_b200 -- B200-native traversal of Embree's compressed subdivision/displacement BVH.

Python face of the C ABI in include/ecb200.h (libecb200.so: hand-written sm_100a CUDA kernels + a thin C++ host
layer).  Python is plumbing only: it moves pointers.  There is no CPU or PyTorch fallback -- importing works without
a GPU (so the library and its symbols can be checked), every traversal call raises EcbError without one.

    scene = Scene.from_image_file("scenes/_built/bomberman_s6c3_leaf.rimg")   # or Scene.import_bvh4(root=..., ...)
    scene.intersect1M(rayhits)            # numpy (N,20) float32 = RTCRayHit[N], mutated in place (rtcIntersect1M)
    scene.intersect_device(ptr, n)        # device-resident rays (torch tensor .data_ptr())
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# ECB200_LIB: developer override used to A/B kernel variants (tools/kbench.py); the shipped library is in-tree
LIB_PATH = os.environ.get("ECB200_LIB") or os.path.join(HERE, "libecb200.so")

INVALID_GEOMETRY_ID = 0xFFFFFFFF
LEAF_GRID, LEAF_LEAF, LEAF_BOX = 0, 1, 2
LEAF_KINDS = {"grid": LEAF_GRID, "leaf": LEAF_LEAF, "box": LEAF_BOX}
ERROR_NO_DEVICE = 100


class EcbError(RuntimeError):
    def __init__(self, code, what, msg):
        super().__init__("%s failed: %s (%d)%s" % (what, _lib().ecbErrorString(code).decode(), code,
                                                   (": " + msg) if msg else ""))
        self.code = code


class ImportDesc(C.Structure):
    _fields_ = [("base", C.c_void_p), ("root", C.c_uint64), ("leaf_kind", C.c_uint32), ("proj_col_major", C.c_uint32),
                ("subdiv_level", C.c_uint32), ("comp_level", C.c_uint32), ("bounds", C.c_float * 6),
                ("image_bytes", C.c_uint64)]


class InstanceDesc(C.Structure):
    _fields_ = [("base", C.c_void_p), ("root", C.c_uint64), ("n_instances", C.c_uint32), ("reserved", C.c_uint32),
                ("geomIDs", C.c_void_p), ("world2local", C.c_void_p), ("objects", C.c_void_p), ("bounds", C.c_float * 6),
                ("image_bytes", C.c_uint64)]


class TriangleDesc(C.Structure):
    _fields_ = [("base", C.c_void_p), ("root", C.c_uint64), ("image_bytes", C.c_uint64)]


class BuildDesc(C.Structure):
    _fields_ = [("leaf_kind", C.c_uint32), ("comp_level", C.c_uint32), ("n_blocks", C.c_uint32), ("reserved", C.c_uint32),
                ("grids", C.c_void_p), ("space", C.c_void_p), ("world", C.c_void_p), ("uv", C.c_void_p),
                ("geomIDs", C.c_void_p), ("primIDs", C.c_void_p)]


class DeviceHits(C.Structure):
    """ecbDeviceHits: compact result arrays of a device-buffer call (device pointers, one entry per ray)."""
    _fields_ = [("geomID", C.c_void_p), ("hit", C.c_void_p), ("instID", C.c_void_p), ("Ng", C.c_void_p)]


class RayNp(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in "org_x org_y org_z tnear dir_x dir_y dir_z time tfar mask id flags".split()]


class HitNp(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in "Ng_x Ng_y Ng_z u v primID geomID instID".split()]


class RayHitNp(C.Structure):
    _fields_ = [("ray", RayNp), ("hit", HitNp)]


class SceneInfo(C.Structure):
    _fields_ = [("leaf_kind", C.c_uint32), ("subdiv_level", C.c_uint32), ("comp_level", C.c_uint32),
                ("n_top_nodes", C.c_uint32), ("n_blocks", C.c_uint32), ("top_depth", C.c_uint32),
                ("block_stride", C.c_uint32), ("top_node_bytes", C.c_uint32), ("device_bytes", C.c_uint64),
                ("reference_bytes", C.c_uint64), ("micro_triangles", C.c_uint64), ("bounds", C.c_float * 6),
                ("device", C.c_int32)]


# every entry point include/ecb200.h declares: (restype, argtypes)
_vp, _u32, _u64, _sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t
SYMBOLS = {
    "ecbSceneImportBVH4": (C.c_int, [C.POINTER(ImportDesc), C.POINTER(_vp)]),
    "ecbSceneImportInstances": (C.c_int, [C.POINTER(InstanceDesc), C.POINTER(_vp)]),
    "ecbSceneAttachTriangles4v": (C.c_int, [_vp, C.POINTER(TriangleDesc)]),
    "ecbBlockStride": (_u32, [_u32, _u32]),
    "ecbBuildBlocks": (C.c_int, [C.POINTER(BuildDesc), _vp, _vp, _vp]),
    "ecbSceneCopyBlocks": (C.c_int, [_vp, _vp, _sz]),
    "ecbSceneSetBlocks": (C.c_int, [_vp, _vp, _sz]),
    "ecbSceneImportImage": (C.c_int, [_vp, _sz, C.POINTER(_vp)]),
    "ecbSceneImportImageFile": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "ecbSceneSave": (C.c_int, [_vp, C.c_char_p]),
    "ecbSceneLoad": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "ecbSceneUpload": (C.c_int, [_vp, C.c_int]),
    "ecbSceneUploadMulti": (C.c_int, [_vp, C.POINTER(C.c_int), _u32]),
    "ecbSceneResidentDevices": (_u32, [_vp, C.POINTER(C.c_int), _u32]),
    "ecbSceneGetInfo": (C.c_int, [_vp, C.POINTER(SceneInfo)]),
    "ecbSceneRelease": (None, [_vp]),
    "ecbIntersect1": (C.c_int, [_vp, _u32, _vp]),
    "ecbIntersect1M": (C.c_int, [_vp, _u32, _vp, _u32, _sz]),
    "ecbIntersect16": (C.c_int, [_vp, _vp, _u32, _vp]),
    "ecbOccluded1": (C.c_int, [_vp, _u32, _vp]),
    "ecbOccluded1M": (C.c_int, [_vp, _u32, _vp, _u32, _sz]),
    "ecbIntersectNp": (C.c_int, [_vp, _u32, C.POINTER(RayHitNp), _u64]),
    "ecbOccludedNp": (C.c_int, [_vp, _u32, C.POINTER(RayNp), _u64]),
    "ecbOccludedExact1M": (C.c_int, [_vp, _u32, _vp, _u32, _sz]),
    "ecbOccludedExactDevice": (C.c_int, [_vp, _u32, _vp, _u64, _sz, _vp]),
    "ecbIntersectDevice": (C.c_int, [_vp, _u32, _vp, _u64, _sz, _vp]),
    "ecbIntersectDeviceEx": (C.c_int, [_vp, _u32, _vp, _u64, _sz, _vp, _vp]),
    "ecbIntersectDeviceStats": (C.c_int, [_vp, _u32, _vp, _u64, _sz, _vp, _vp]),
    "ecbOccludedDevice": (C.c_int, [_vp, _u32, _vp, _u64, _sz, _vp]),
    "ecbIntersectDeviceCompact": (C.c_int, [_vp, _u32, _vp, _u64, _sz, C.POINTER(DeviceHits), _vp]),
    "ecbOccludedDeviceCompact": (C.c_int, [_vp, _u32, _vp, _u64, _sz, C.POINTER(DeviceHits), _vp]),
    "ecbIntersectDeviceSoA": (C.c_int, [_vp, _u32, C.POINTER(_vp), _u64, C.POINTER(DeviceHits), _vp]),
    "ecbDeviceAlloc": (C.c_int, [C.POINTER(_vp), _sz]),
    "ecbDeviceFree": (None, [_vp]),
    "ecbIpcExport": (C.c_int, [_vp, _vp]),
    "ecbIpcImport": (C.c_int, [_vp, C.POINTER(_vp)]),
    "ecbIpcClose": (None, [_vp]),
    "ecbPeerCopyAsync": (C.c_int, [_vp, _vp, _sz, _vp]),
    "ecbSceneBlockNormals": (C.c_int, [_vp, _vp, _sz]),
    "ecbHostAlloc": (C.c_int, [C.POINTER(_vp), _sz]),
    "ecbHostFree": (None, [_vp]),
    "ecbHostPathName": (C.c_char_p, [C.c_int]),
    "ecbBoundsCheckFailures": (C.c_int, [C.POINTER(_u64)]),
    "ecbSetTraversalKernel": (None, [C.c_int]),
    "ecbGetTraversalKernel": (C.c_int, []),
    "ecbProbeReadBandwidth": (C.c_int, [_sz, C.c_int, C.POINTER(C.c_double)]),
    "ecbLaunchCount": (_u64, []),
    "ecbErrorString": (C.c_char_p, [C.c_int]),
    "ecbLastErrorMessage": (C.c_char_p, []),
}

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no fallback implementation)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def _check(code, what):
    if code != 0:
        raise EcbError(code, what, _lib().ecbLastErrorMessage().decode())


def launch_count():
    return int(_lib().ecbLaunchCount())


def host_path_for(pinned):
    """Name of the transfer strategy the host-buffer entry points pick for this kind of caller memory."""
    return _lib().ecbHostPathName(1 if pinned else 0).decode()


def bounds_check_failures():
    """(enabled, [8 violation counters]) of a -DECB_BOUNDS_CHECK build; enabled is False for the shipping library."""
    c = (_u64 * 8)()
    r = _lib().ecbBoundsCheckFailures(c)
    if r < 0:
        raise RuntimeError("ecbBoundsCheckFailures: CUDA error")
    return bool(r), [int(v) for v in c]


KERNEL_LANE, KERNEL_QUAD = 0, 1


def set_traversal_kernel(mode):
    """ECB_KERNEL_LANE / ECB_KERNEL_QUAD ("lane" / "quad"), or -1 / None for the default."""
    mode = {"lane": 0, "quad": 1, None: -1}.get(mode, mode)
    _lib().ecbSetTraversalKernel(int(mode))


def traversal_kernel():
    return "quad" if _lib().ecbGetTraversalKernel() == 1 else "lane"


def probe_read_bandwidth(working_set_bytes, passes):
    """GB/s of 128-bit L1-bypassing reads over a device buffer (32 MB x 64 passes: L2; 4 GB x 1: HBM)."""
    g = C.c_double()
    _check(_lib().ecbProbeReadBandwidth(working_set_bytes, passes, C.byref(g)), "ecbProbeReadBandwidth")
    return g.value


def host_alloc(n_bytes):
    """Pinned host memory: returns (numpy uint8 view, pointer); free the pointer with host_free."""
    p = _vp()
    _check(_lib().ecbHostAlloc(C.byref(p), n_bytes), "ecbHostAlloc")
    buf = (C.c_uint8 * n_bytes).from_address(p.value)
    return np.frombuffer(buf, np.uint8), p.value


def host_free(ptr):
    _lib().ecbHostFree(ptr)


def build_blocks(leaf_kind, comp_level, grids, space, world, uv, geom_ids, prim_ids):
    """ecbBuildBlocks: compressed blocks from displaced vertex grids and frames, built on the GPU.
    Returns (blocks uint8 (n, stride), bounds float32 (n, 6), patch_ok uint32 (n,))."""
    kind = LEAF_KINDS.get(leaf_kind, leaf_kind)
    arrs = [np.ascontiguousarray(a, np.float32) for a in (grids, space, world, uv)]
    ids = [np.ascontiguousarray(a, np.uint32) for a in (geom_ids, prim_ids)]
    n = len(ids[0])
    stride = int(_lib().ecbBlockStride(kind, comp_level))
    d = BuildDesc()
    d.leaf_kind, d.comp_level, d.n_blocks = kind, comp_level, n
    d.grids, d.space, d.world, d.uv = (a.ctypes.data for a in arrs)
    d.geomIDs, d.primIDs = (a.ctypes.data for a in ids)
    blocks = np.zeros((n, max(stride, 1)), np.uint8)
    bounds = np.zeros((n, 6), np.float32)
    ok = np.zeros(n, np.uint32)
    _check(_lib().ecbBuildBlocks(C.byref(d), blocks.ctypes.data, bounds.ctypes.data, ok.ctypes.data), "ecbBuildBlocks")
    return blocks, bounds, ok


def _addr(a):
    return int(a) if isinstance(a, int) else a.ctypes.data


def soa_from_rayhits(rayhits, alloc=None):
    """RTCRayHitNp view of an (N, 20) RTCRayHit array: one contiguous array per field (copies).  alloc(nbytes) -> (uint8 view,
    ptr) allocates each array (e.g. host_alloc for pinned memory); default numpy."""
    names = Scene.RAY_FIELDS + Scene.HIT_FIELDS
    out = {}
    for k, f in enumerate(names):
        col = np.ascontiguousarray(rayhits[:, k]) if k < rayhits.shape[1] else np.zeros(len(rayhits), np.float32)
        if alloc is not None:
            buf, _ = alloc(col.nbytes)
            arr = buf.view(np.float32)
            arr[:] = col
            col = arr
        out[f] = col.view(np.uint32) if f in ("mask", "id", "flags", "primID", "geomID", "instID") else col
    return out


def rayhits_from_soa(soa, n):
    """Inverse of soa_from_rayhits: (n, 20) float32."""
    names = Scene.RAY_FIELDS + Scene.HIT_FIELDS
    r = np.zeros((n, 20), np.float32)
    for k, f in enumerate(names):
        r[:, k] = soa[f][:n].view(np.float32)
    return r


class Scene:
    """A committed compressed-BVH scene in the device format (host copy + lazily uploaded device copy)."""

    def __init__(self, handle, keepalive=None):
        self._h = handle
        self._keepalive = keepalive

    # -- construction -------------------------------------------------------------------------------------------
    @classmethod
    def import_bvh4(cls, root, leaf_kind, comp_level, subdiv_level=0, proj_col_major=1, base=0, bounds=None):
        """ecbSceneImportBVH4: base = 0 walks live reference memory in this process."""
        d = ImportDesc()
        d.base = base
        d.root = root
        d.leaf_kind = LEAF_KINDS.get(leaf_kind, leaf_kind)
        d.proj_col_major = proj_col_major
        d.subdiv_level = subdiv_level
        d.comp_level = comp_level
        for i in range(6):
            d.bounds[i] = float(bounds[i]) if bounds is not None else 0.0
        h = _vp()
        _check(_lib().ecbSceneImportBVH4(C.byref(d), C.byref(h)), "ecbSceneImportBVH4")
        return cls(h.value)

    @classmethod
    def import_instances(cls, root, geom_ids, world2local, objects, base=0, bounds=None):
        """ecbSceneImportInstances: objects[i] is the imported Scene instance i refers to (entries may repeat)."""
        ids = np.ascontiguousarray(geom_ids, np.uint32)
        xf = np.ascontiguousarray(world2local, np.float32).reshape(-1, 12)
        handles = (C.c_void_p * len(objects))(*[o._h for o in objects])
        d = InstanceDesc()
        d.base = base
        d.root = root
        d.n_instances = len(ids)
        d.geomIDs = ids.ctypes.data
        d.world2local = xf.ctypes.data
        d.objects = C.cast(handles, C.c_void_p).value
        for i in range(6):
            d.bounds[i] = float(bounds[i]) if bounds is not None else 0.0
        h = _vp()
        _check(_lib().ecbSceneImportInstances(C.byref(d), C.byref(h)), "ecbSceneImportInstances")
        return cls(h.value, keepalive=list(objects))

    def attach_triangles4v(self, root, base=0):
        """ecbSceneAttachTriangles4v: the scene's BVH4 over Triangle4v blocks (mixed scenes; before upload)."""
        d = TriangleDesc()
        d.base = base
        d.root = root
        _check(_lib().ecbSceneAttachTriangles4v(self._h, C.byref(d)), "ecbSceneAttachTriangles4v")
        return self

    @classmethod
    def from_image(cls, image):
        image = np.ascontiguousarray(image, np.uint8)
        h = _vp()
        _check(_lib().ecbSceneImportImage(image.ctypes.data, image.nbytes, C.byref(h)), "ecbSceneImportImage")
        return cls(h.value)

    @classmethod
    def from_image_file(cls, path):
        h = _vp()
        _check(_lib().ecbSceneImportImageFile(os.fsencode(path), C.byref(h)), "ecbSceneImportImageFile")
        return cls(h.value)

    @classmethod
    def load(cls, path):
        h = _vp()
        _check(_lib().ecbSceneLoad(os.fsencode(path), C.byref(h)), "ecbSceneLoad")
        return cls(h.value)

    def save(self, path):
        _check(_lib().ecbSceneSave(self._h, os.fsencode(path)), "ecbSceneSave")

    def upload(self, device=-1):
        _check(_lib().ecbSceneUpload(self._h, device), "ecbSceneUpload")
        return self

    def upload_multi(self, devices):
        """ecbSceneUploadMulti: replicate on several CUDA devices; long host-buffer calls then fan out over them."""
        arr = (C.c_int * len(devices))(*devices)
        _check(_lib().ecbSceneUploadMulti(self._h, arr, len(devices)), "ecbSceneUploadMulti")
        return self

    def resident_devices(self):
        arr = (C.c_int * 64)()
        n = _lib().ecbSceneResidentDevices(self._h, arr, 64)
        return [arr[i] for i in range(min(n, 64))]

    def release(self):
        if self._h:
            _lib().ecbSceneRelease(self._h)
            self._h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    # -- queries ------------------------------------------------------------------------------------------------
    def info(self):
        i = SceneInfo()
        _check(_lib().ecbSceneGetInfo(self._h, C.byref(i)), "ecbSceneGetInfo")
        d = {n: getattr(i, n) for n, _ in SceneInfo._fields_ if n != "bounds"}
        d["bounds"] = [i.bounds[k] for k in range(6)]
        d["bytes_per_micro_triangle"] = d["device_bytes"] / max(1, d["micro_triangles"])
        d["reference_bytes_per_micro_triangle"] = d["reference_bytes"] / max(1, d["micro_triangles"])
        return d

    def copy_blocks(self):
        """Host copy of the converted blocks: uint8 (n_blocks, block_stride)."""
        i = self.info()
        out = np.zeros((i["n_blocks"], i["block_stride"]), np.uint8)
        _check(_lib().ecbSceneCopyBlocks(self._h, out.ctypes.data, out.nbytes), "ecbSceneCopyBlocks")
        return out

    def set_blocks(self, blocks):
        b = np.ascontiguousarray(blocks, np.uint8)
        _check(_lib().ecbSceneSetBlocks(self._h, b.ctypes.data, b.nbytes), "ecbSceneSetBlocks")
        return self

    def block_normals(self):
        n = self.info()["n_blocks"]
        out = np.zeros((n, 3), np.float32)
        _check(_lib().ecbSceneBlockNormals(self._h, out.ctypes.data, out.size), "ecbSceneBlockNormals")
        return out

    # -- traversal, host buffers (rtcIntersect1M / rtcOccluded1M / rtcIntersect16 semantics) -----------------------
    @staticmethod
    def _host_args(a, width):
        if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.ndim == 2 and a.shape[1] >= width):
            raise ValueError("rays must be a float32 array of shape (N, >=%d)" % width)
        if a.shape[0] == 0:
            return None, 0, 4 * a.shape[1]
        if a.strides[1] != 4:
            raise ValueError("ray fields must be contiguous")
        return a.ctypes.data, a.shape[0], a.strides[0]

    def intersect1M(self, rayhits, inst_id=INVALID_GEOMETRY_ID):
        p, n, stride = self._host_args(rayhits, 20)
        _check(_lib().ecbIntersect1M(self._h, inst_id, p, n, stride), "ecbIntersect1M")

    def occluded1M(self, rays, inst_id=INVALID_GEOMETRY_ID):
        p, n, stride = self._host_args(rays, 12)
        _check(_lib().ecbOccluded1M(self._h, inst_id, p, n, stride), "ecbOccluded1M")

    def occluded_exact1M(self, rays, inst_id=INVALID_GEOMETRY_ID):
        """Opt-in exact shadow query (not the reference's conservative any-hit): enters the blocks."""
        p, n, stride = self._host_args(rays, 12)
        _check(_lib().ecbOccludedExact1M(self._h, inst_id, p, n, stride), "ecbOccludedExact1M")

    RAY_FIELDS = "org_x org_y org_z tnear dir_x dir_y dir_z time tfar mask id flags".split()
    HIT_FIELDS = "Ng_x Ng_y Ng_z u v primID geomID instID".split()

    def intersectNp(self, soa, n, inst_id=INVALID_GEOMETRY_ID):
        """ecbIntersectNp (rtcIntersectNp): soa maps field name -> 1-D float32 / uint32 array of at least n entries, or
        -> address (int); the arrays are mutated in place."""
        s = RayHitNp()
        for f in self.RAY_FIELDS:
            setattr(s.ray, f, _addr(soa[f]))
        for f in self.HIT_FIELDS:
            setattr(s.hit, f, _addr(soa[f]))
        _check(_lib().ecbIntersectNp(self._h, inst_id, C.byref(s), n), "ecbIntersectNp")

    def occludedNp(self, soa, n, inst_id=INVALID_GEOMETRY_ID):
        s = RayNp()
        for f in self.RAY_FIELDS:
            setattr(s, f, _addr(soa[f]))
        _check(_lib().ecbOccludedNp(self._h, inst_id, C.byref(s), n), "ecbOccludedNp")

    def intersect1(self, rayhit, inst_id=INVALID_GEOMETRY_ID):
        _check(_lib().ecbIntersect1(self._h, inst_id, rayhit.ctypes.data), "ecbIntersect1")

    def occluded1(self, ray, inst_id=INVALID_GEOMETRY_ID):
        _check(_lib().ecbOccluded1(self._h, inst_id, ray.ctypes.data), "ecbOccluded1")

    def intersect16(self, valid, rayhit16, inst_id=INVALID_GEOMETRY_ID):
        """valid: int32[16] (-1 active / 0 inactive); rayhit16: float32 (20,16) SoA = RTCRayHit16."""
        valid = np.ascontiguousarray(valid, np.int32)
        assert rayhit16.dtype == np.float32 and rayhit16.shape == (20, 16) and rayhit16.flags.c_contiguous
        _check(_lib().ecbIntersect16(valid.ctypes.data, self._h, inst_id, rayhit16.ctypes.data), "ecbIntersect16")

    # -- traversal, device buffers ----------------------------------------------------------------------------------
    def intersect_device(self, ptr, n, stride=80, stream=0, block_out_ptr=0, inst_id=INVALID_GEOMETRY_ID):
        if block_out_ptr:
            _check(_lib().ecbIntersectDeviceEx(self._h, inst_id, ptr, n, stride, block_out_ptr, stream),
                   "ecbIntersectDeviceEx")
        else:
            _check(_lib().ecbIntersectDevice(self._h, inst_id, ptr, n, stride, stream), "ecbIntersectDevice")

    def intersect_device_stats(self, ptr, n, counters_ptr, stride=80, stream=0, inst_id=INVALID_GEOMETRY_ID):
        """counters_ptr: device uint64[6] = rays, top nodes, blocks, quantized nodes, cells, hits (accumulated)."""
        _check(_lib().ecbIntersectDeviceStats(self._h, inst_id, ptr, n, stride, counters_ptr, stream),
               "ecbIntersectDeviceStats")

    def intersect_device_compact(self, ptr, n, geom_ptr, hit_ptr, stride=80, stream=0, inst_ptr=0, inst_id=INVALID_GEOMETRY_ID):
        """ecbIntersectDeviceCompact: rays are only read; geomID[n] (uint32) and {tfar,u,v,primID}[n] (4 floats, hits only)
        are written to the given device arrays (+ instID[n] on instancing scenes)."""
        out = DeviceHits(geom_ptr, hit_ptr, inst_ptr or None, None)
        _check(_lib().ecbIntersectDeviceCompact(self._h, inst_id, ptr, n, stride, C.byref(out), stream), "ecbIntersectDeviceCompact")

    def intersect_device_soa(self, soa_ptrs, n, geom_ptr, hit_ptr, stream=0, inst_id=INVALID_GEOMETRY_ID):
        """ecbIntersectDeviceSoA: soa_ptrs = 8 device pointers (org_x, org_y, org_z, tnear, dir_x, dir_y, dir_z, tfar)."""
        arr = (_vp * 8)(*soa_ptrs)
        out = DeviceHits(geom_ptr, hit_ptr, None, None)
        _check(_lib().ecbIntersectDeviceSoA(self._h, inst_id, arr, n, C.byref(out), stream), "ecbIntersectDeviceSoA")

    def occluded_device_compact(self, ptr, n, geom_ptr, stride=48, stream=0, inst_id=INVALID_GEOMETRY_ID):
        out = DeviceHits(geom_ptr, None, None, None)
        _check(_lib().ecbOccludedDeviceCompact(self._h, inst_id, ptr, n, stride, C.byref(out), stream), "ecbOccludedDeviceCompact")

    def occluded_device(self, ptr, n, stride=48, stream=0, inst_id=INVALID_GEOMETRY_ID):
        _check(_lib().ecbOccludedDevice(self._h, inst_id, ptr, n, stride, stream), "ecbOccludedDevice")


# ---- multi-GPU gather plumbing (one process per GPU): device buffers another process can map over NVLink ----
def device_alloc(n_bytes):
    p = _vp()
    _check(_lib().ecbDeviceAlloc(C.byref(p), n_bytes), "ecbDeviceAlloc")
    return p.value


def device_free(ptr):
    _lib().ecbDeviceFree(ptr)


def ipc_export(ptr):
    """64-byte CUDA IPC handle of a buffer from device_alloc (bytes)."""
    h = (C.c_uint8 * 64)()
    _check(_lib().ecbIpcExport(ptr, h), "ecbIpcExport")
    return bytes(h)


def ipc_import(handle):
    h = (C.c_uint8 * 64).from_buffer_copy(handle)
    p = _vp()
    _check(_lib().ecbIpcImport(h, C.byref(p)), "ecbIpcImport")
    return p.value


def ipc_close(ptr):
    _lib().ecbIpcClose(ptr)


def peer_copy_async(dst, src, n_bytes, stream=0):
    _check(_lib().ecbPeerCopyAsync(dst, src, n_bytes, stream), "ecbPeerCopyAsync")
