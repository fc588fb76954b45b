"""ctypes bindings of the parity oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  It wraps
  * oracle/_ref/libembree3_ref.so  -- the UNMODIFIED reference compiled by oracle/build_ref.py plus the
    harness oracle/refhost.cpp (scene construction through rtc*, rtcIntersect1/rtcOccluded1 loops);
  * oracle/_ref/libcbvh_port.so    -- the scalar C restatement oracle/cbvh_port.c.
Rays are numpy float32 arrays of shape (N, 20) = RTCRayHit (80 bytes) or (N, 12) = RTCRay (48 bytes); integer
fields are read through a uint32 view (columns: see RAYHIT_COLUMNS).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_SO = os.path.join(REF_DIR, "libembree3_ref.so")
REF_SO_AVX512 = os.path.join(REF_DIR, "libembree3_ref_avx512.so")   # timing baseline only (EMBREE_MAX_ISA=AVX512SKX)
PORT_SO = os.path.join(REF_DIR, "libcbvh_port.so")
BOMBERMAN_OBJ = os.path.join(REF_DIR, "assets", "bomberman.obj")

# RTCRayHit columns (include/embree3/rtcore_ray.h:26-64)
ORG, TNEAR, DIR, TIME, TFAR, MASK, ID, FLAGS, NG, U, V, PRIMID, GEOMID, INSTID = (
    slice(0, 3), 3, slice(4, 7), 7, 8, 9, 10, 11, slice(12, 15), 15, 16, 17, 18, 19)
INVALID = 0xFFFFFFFF
LEAF_KINDS = {"grid": 0, "leaf": 1, "box": 2}


def build_port(force=False):
    """gcc the C restatement (no contraction: see cbvh_port.h)."""
    srcs = [os.path.join(HERE, "cbvh_port.c"), os.path.join(HERE, "cbvh_build_port.c")]
    os.makedirs(REF_DIR, exist_ok=True)
    deps = srcs + [os.path.join(HERE, "cbvh_port.h")]
    if not force and os.path.exists(PORT_SO) and all(os.path.getmtime(PORT_SO) >= os.path.getmtime(d) for d in deps):
        return PORT_SO
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-Wall", "-o", PORT_SO] + srcs +
                   ["-lm"], check=True)
    return PORT_SO


def have_reference():
    return os.path.exists(REF_SO)


class _Layout(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "sizeof_block off_geomID off_primID off_uv off_rcp_edges off_nodes off_leaves off_grid off_extent "
        "off_elems off_grid_width off_space off_proj off_box proj_col_major sizeof_node sizeof_leaf sizeof_vec3f "
        "sizeof_aligned_node sizeof_rayhit off_an_children off_an_lower_x").split()]


_ref = {}


def have_avx512_reference():
    """The AVX512SKX build exists and this CPU can run it."""
    try:
        flags = open("/proc/cpuinfo").read()
    except Exception:
        return False
    return os.path.exists(REF_SO_AVX512) and all(f in flags for f in ("avx512f", "avx512dq", "avx512cd", "avx512bw", "avx512vl"))


def ref_lib(path=None):
    path = path or REF_SO
    if path not in _ref:
        L = C.CDLL(path)
        vp = C.c_void_p
        L.ref_layout.argtypes = [C.POINTER(_Layout)]
        L.ref_scene_new.restype = vp
        L.ref_scene_new.argtypes = [C.c_char_p, C.c_uint, C.c_uint, C.c_int]
        L.ref_scene_free.argtypes = [vp]
        L.ref_scene_add_subdiv.restype = C.c_uint
        L.ref_scene_add_subdiv.argtypes = [vp, vp, C.c_uint, vp, C.c_uint, vp, C.c_uint, C.c_int, C.c_float]
        L.ref_scene_add_cube.restype = C.c_uint
        L.ref_scene_add_cube.argtypes = [vp, C.c_int, C.c_float]
        L.ref_scene_add_obj.restype = C.c_int
        L.ref_scene_add_obj.argtypes = [vp, C.c_char_p]
        L.ref_scene_add_ground_plane.restype = C.c_uint
        L.ref_scene_add_ground_plane.argtypes = [vp]
        L.ref_scene_add_instance.restype = C.c_uint
        L.ref_scene_add_instance.argtypes = [vp, vp, vp]
        L.ref_scene_commit.argtypes = [vp]
        L.ref_scene_error.restype = C.c_int
        L.ref_scene_error.argtypes = [vp]
        L.ref_scene_bounds.argtypes = [vp, vp]
        L.ref_scene_handle.restype = vp
        L.ref_scene_handle.argtypes = [vp]
        L.ref_scene_compressed_root.restype = C.c_uint64
        L.ref_scene_compressed_root.argtypes = [vp, C.POINTER(C.c_int)]
        L.ref_intersect.restype = C.c_double
        L.ref_intersect.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int]
        L.ref_occluded.restype = C.c_double
        L.ref_occluded.argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int]
        L.ref_hit_normals.argtypes = [vp, vp, C.c_size_t, C.c_size_t, vp]
        L.ref_scene_dump_image.restype = C.c_size_t
        L.ref_scene_dump_image.argtypes = [vp, vp, C.c_size_t]
        L.ref_hardware_threads.restype = C.c_int
        L.ref_scene_triangle_root.restype = C.c_uint64
        L.ref_scene_triangle_root.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint)]
        L.ref_scene_instance_info.restype = C.c_int
        L.ref_scene_instance_info.argtypes = [vp, C.POINTER(C.c_uint64), C.c_uint, vp, vp, vp]
        _ref[path] = L
    return _ref[path]


def ref_layout():
    lay = _Layout()
    ref_lib().ref_layout(C.byref(lay))
    return {n: getattr(lay, n) for n, _ in _Layout._fields_}


def hardware_threads():
    return os.cpu_count() or 1


class RefScene:
    """A scene built and traced by the compiled reference."""

    def __init__(self, leaf, S, C_, build_threads=0, lib_path=None):
        self.L = ref_lib(lib_path)
        self.leaf, self.S, self.C = leaf, S, C_
        self.h = self.L.ref_scene_new(leaf.encode(), S, C_, build_threads)
        self._children = []

    def add_cube(self, displacement=1, scale=1.0):
        return self.L.ref_scene_add_cube(self.h, displacement, scale)

    def add_obj(self, path=BOMBERMAN_OBJ):
        r = self.L.ref_scene_add_obj(self.h, path.encode())
        if r < 0:
            raise RuntimeError("cannot read %s (%d)" % (path, r))
        return r

    def add_subdiv(self, verts, indices, faces, displacement=0, scale=0.0):
        v = np.ascontiguousarray(verts, np.float32)
        i = np.ascontiguousarray(indices, np.uint32)
        f = np.ascontiguousarray(faces, np.uint32)
        return self.L.ref_scene_add_subdiv(self.h, v.ctypes.data, len(v), i.ctypes.data, len(i), f.ctypes.data, len(f),
                                           displacement, scale)

    def add_ground_plane(self):
        return self.L.ref_scene_add_ground_plane(self.h)

    def add_instance(self, obj, xfm12):
        x = np.ascontiguousarray(xfm12, np.float32)
        self._children.append(obj)
        return self.L.ref_scene_add_instance(self.h, obj.h, x.ctypes.data)

    def commit(self):
        self.L.ref_scene_commit(self.h)
        err = self.L.ref_scene_error(self.h)
        if err:
            raise RuntimeError("reference reported RTCError %d on commit" % err)
        return self

    def bounds(self):
        b = np.zeros(6, np.float32)
        self.L.ref_scene_bounds(self.h, b.ctypes.data)
        return b

    def root(self):
        n = C.c_int()
        r = self.L.ref_scene_compressed_root(self.h, C.byref(n))
        return r, n.value

    def handle(self):
        return self.L.ref_scene_handle(self.h)

    def triangle_root(self):
        """(root of the BVH4 over Triangle4v blocks or 0, its index and the compressed accel's index in the AccelN list,
        sizeof(Triangle4v))."""
        ti, si, sz = C.c_int(), C.c_int(), C.c_uint()
        r = self.L.ref_scene_triangle_root(self.h, C.byref(ti), C.byref(si), C.byref(sz))
        return r, ti.value, si.value, sz.value

    def intersect(self, rayhits, threads=None):
        assert rayhits.dtype == np.float32 and rayhits.flags.c_contiguous
        return self.L.ref_intersect(self.h, rayhits.ctypes.data, rayhits.shape[0], rayhits.strides[0],
                                    threads or hardware_threads())

    def occluded(self, rays, threads=None):
        assert rays.dtype == np.float32 and rays.flags.c_contiguous
        return self.L.ref_occluded(self.h, rays.ctypes.data, rays.shape[0], rays.strides[0],
                                   threads or hardware_threads())

    def hit_normals(self, rayhits):
        n = np.zeros((rayhits.shape[0], 3), np.float32)
        self.L.ref_hit_normals(self.h, rayhits.ctypes.data, rayhits.shape[0], rayhits.strides[0], n.ctypes.data)
        return n

    def instance_info(self, cap=4096):
        """(root of the user-geometry BVH4, geomIDs[n], world2local[n,12], instanced RTCScene handles[n])."""
        root = C.c_uint64()
        ids = np.zeros(cap, np.uint32)
        xf = np.zeros((cap, 12), np.float32)
        objs = (C.c_void_p * cap)()
        n = self.L.ref_scene_instance_info(self.h, C.byref(root), cap, ids.ctypes.data, xf.ctypes.data, objs)
        if n < 0:
            raise RuntimeError("scene has no instance accel")
        assert n <= cap
        return root.value, ids[:n].copy(), xf[:n].copy(), [objs[i] for i in range(n)]

    def dump_image(self):
        size = self.L.ref_scene_dump_image(self.h, None, 0)
        if size == 0:
            raise RuntimeError("scene has no compressed accel")
        buf = np.zeros(size, np.uint8)
        got = self.L.ref_scene_dump_image(self.h, buf.ctypes.data, size)
        assert got == size
        return buf

    def free(self):
        if self.h:
            self.L.ref_scene_free(self.h)
            self.h = None


# ---------------------------------------------------------------------------------------------- port
class _PortScene(C.Structure):
    _fields_ = [("base", C.c_void_p), ("root", C.c_uint64), ("leaf_kind", C.c_int32), ("proj_col_major", C.c_int32),
                ("tri_root", C.c_uint64)]


class _Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in "rays top_nodes blocks cnodes cells hits".split()]


_port = None


def port_lib():
    global _port
    if _port is None:
        build_port()
        L = C.CDLL(PORT_SO)
        L.cbvh_intersectN.restype = C.c_double
        L.cbvh_intersectN.argtypes = [C.POINTER(_PortScene), C.c_void_p, C.c_size_t, C.c_size_t, C.c_int,
                                      C.POINTER(_Stats)]
        L.cbvh_occludedN.restype = C.c_double
        L.cbvh_occludedN.argtypes = L.cbvh_intersectN.argtypes
        L.cbvh_candidatesN.argtypes = [C.POINTER(_PortScene), C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
        _port = L
    return _port


RIMG_HEADER = np.dtype([("magic", "S8"), ("version", "<u4"), ("leaf_kind", "<u4"), ("subdiv_level", "<u4"),
                        ("comp_level", "<u4"), ("proj_col_major", "<u4"), ("n_nodes", "<u4"), ("n_blocks", "<u4"),
                        ("reserved0", "<u4"), ("root", "<u8"), ("total_bytes", "<u8"), ("bounds", "<f4", (6,)),
                        ("tri_root", "<u8"), ("reserved1", "<u4", (2,))])
assert RIMG_HEADER.itemsize == 96


def rimg_header(image):
    return np.frombuffer(image[:96].tobytes(), RIMG_HEADER)[0]


class PortScene:
    """The C restatement bound either to a live reference scene (base=0) or to an RIMG buffer."""

    def __init__(self, base, root, leaf_kind, proj_col_major, keepalive=None, tri_root=0):
        self.L = port_lib()
        self.s = _PortScene(base, root, leaf_kind, proj_col_major, tri_root)
        self.keepalive = keepalive
        self.last_stats = None

    @classmethod
    def from_ref(cls, ref_scene):
        root, _ = ref_scene.root()
        if not root:
            raise RuntimeError("no compressed accel")
        return cls(0, root, LEAF_KINDS[ref_scene.leaf.split(",")[0]], ref_layout()["proj_col_major"], keepalive=ref_scene,
                   tri_root=ref_scene.triangle_root()[0])

    @classmethod
    def from_image(cls, image):
        h = rimg_header(image)
        assert h["magic"] == b"ECBRIMG1"
        return cls(image.ctypes.data, int(h["root"]), int(h["leaf_kind"]), int(h["proj_col_major"]), keepalive=image,
                   tri_root=int(h["tri_root"]))

    def _stats(self, st):
        self.last_stats = {n: getattr(st, n) for n, _ in _Stats._fields_}

    def intersect(self, rayhits, threads=None):
        st = _Stats()
        dt = self.L.cbvh_intersectN(C.byref(self.s), rayhits.ctypes.data, rayhits.shape[0], rayhits.strides[0],
                                    threads or hardware_threads(), C.byref(st))
        self._stats(st)
        return dt

    def occluded(self, rays, threads=None):
        st = _Stats()
        dt = self.L.cbvh_occludedN(C.byref(self.s), rays.ctypes.data, rays.shape[0], rays.strides[0],
                                   threads or hardware_threads(), C.byref(st))
        self._stats(st)
        return dt

    def candidates(self, rayhits, threads=None):
        out = np.zeros((rayhits.shape[0], 4), np.float32)
        self.L.cbvh_candidatesN(C.byref(self.s), rayhits.ctypes.data, rayhits.shape[0], rayhits.strides[0],
                                threads or hardware_threads(), out.ctypes.data)
        return out  # columns: t_best, t_second, prim_best(bits), prim_second(bits)


# ---------------------------------------------------------------------------------------------- instanced port
class _PortInstance(C.Structure):
    _fields_ = [("world2local", C.c_float * 12), ("geomID", C.c_uint32), ("object", C.c_uint32)]


class _PortInstScene(C.Structure):
    _fields_ = [("base", C.c_void_p), ("root", C.c_uint64), ("instances", C.POINTER(_PortInstance)), ("n_instances", C.c_uint32),
                ("reserved", C.c_uint32), ("geom_to_instance", C.POINTER(C.c_uint32)), ("n_geoms", C.c_uint32),
                ("reserved2", C.c_uint32), ("objects", C.POINTER(_PortScene))]


class PortInstancedScene:
    """The C restatement of the instancing path bound to a live reference scene of instances."""

    def __init__(self, ref_top, object_scenes):
        """ref_top: committed RefScene whose geometries are instances of the RefScenes in object_scenes."""
        self.L = port_lib()
        self.L.cbvh_inst_intersectN.restype = C.c_double
        self.L.cbvh_inst_intersectN.argtypes = [C.POINTER(_PortInstScene), C.c_void_p, C.c_size_t, C.c_size_t, C.c_int,
                                                C.POINTER(_Stats)]
        self.L.cbvh_inst_occludedN.restype = C.c_double
        self.L.cbvh_inst_occludedN.argtypes = self.L.cbvh_inst_intersectN.argtypes
        root, ids, xf, handles = ref_top.instance_info()
        by_handle = {o.handle(): i for i, o in enumerate(object_scenes)}
        self.objects = (_PortScene * len(object_scenes))()
        lay = ref_layout()
        for i, o in enumerate(object_scenes):
            r, _ = o.root()
            self.objects[i] = _PortScene(0, r, LEAF_KINDS[o.leaf.split(",")[0]], lay["proj_col_major"], o.triangle_root()[0])
        self.instances = (_PortInstance * len(ids))()
        n_geoms = int(ids.max()) + 1 if len(ids) else 0
        self.g2i = (C.c_uint32 * max(1, n_geoms))(*([INVALID] * max(1, n_geoms)))
        for k in range(len(ids)):
            self.instances[k].geomID = int(ids[k])
            self.instances[k].object = by_handle[handles[k]]
            for j in range(12):
                self.instances[k].world2local[j] = float(xf[k, j])
            self.g2i[int(ids[k])] = k
        self.s = _PortInstScene(0, root, self.instances, len(ids), 0, self.g2i, n_geoms, 0, self.objects)
        self.keepalive = (ref_top, object_scenes)
        self.info = (root, ids, xf, [by_handle[h] for h in handles])

    def intersect(self, rayhits, threads=None):
        st = _Stats()
        return self.L.cbvh_inst_intersectN(C.byref(self.s), rayhits.ctypes.data, rayhits.shape[0], rayhits.strides[0],
                                           threads or hardware_threads(), C.byref(st))

    def candidates(self, rayhits, threads=None):
        self.L.cbvh_inst_candidatesN.argtypes = [C.POINTER(_PortInstScene), C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
        out = np.zeros((rayhits.shape[0], 4), np.float32)
        self.L.cbvh_inst_candidatesN(C.byref(self.s), rayhits.ctypes.data, rayhits.shape[0], rayhits.strides[0],
                                     threads or hardware_threads(), out.ctypes.data)
        return out

    def occluded(self, rays, threads=None):
        st = _Stats()
        return self.L.cbvh_inst_occludedN(C.byref(self.s), rays.ctypes.data, rays.shape[0], rays.strides[0],
                                          threads or hardware_threads(), C.byref(st))
