#!/usr/bin/env python3
"""Compile the UNMODIFIED reference (Embree 3.0.0 + compressed-BVH patch) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  The product never links or loads anything produced here; the
result is the parity oracle and the `cpu_baseline` / `--impl reference` arm.

What this does (and does not do):
  * g++ is run directly on the reference's own source files where they lie under
    /root/reference (read-only) -- the reference's CMake build system is NOT run.
  * The source lists and per-ISA flags restate kernels/CMakeLists.txt:31-176,
    common/*/CMakeLists.txt and common/cmake/gcc.cmake:17-46 for the reference's default
    configuration (EMBREE_MAX_ISA=AVX2 => SSE2 (lowest) + SSE4.2 + AVX + AVX2; Release flags),
    with EMBREE_TASKING_SYSTEM=INTERNAL because TBB is not installed.
  * The two headers CMake would generate (kernels/config.h, kernels/hash.h) are supplied from
    oracle/gen/ through an include path; Eigen3 (absent in the image) is supplied by the
    stand-in in oracle/eigen_shim/ (see its header for what it implements).
  * Outputs: oracle/_ref/obj/*.o, oracle/_ref/libembree3_ref.so (reference + oracle/refhost.cpp
    harness), oracle/_ref/assets/bomberman.obj (the reference's demo asset, copied so that it
    travels to the GPU box), oracle/_ref/BUILD_INFO.json.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ECB_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")

FLAGS_ISA = {
    "sse2": "-msse2",
    "sse42": "-msse4.2",
    "avx": "-mavx",
    "avx2": "-mf16c -mavx2 -mfma -mlzcnt -mbmi -mbmi2",
    # common/cmake/gcc.cmake:22 (EMBREE_MAX_ISA=AVX512SKX; only the second library below is built with it)
    "avx512skx": "-mavx512f -mavx512dq -mavx512cd -mavx512bw -mavx512vl -mf16c -mavx2 -mfma -mlzcnt -mbmi -mbmi2",
}
# common/cmake/gcc.cmake:26-40 + Release (:52-54); -Wall dropped (noise only), -fPIE dropped (we link a .so)
COMMON = ("-fPIC -std=c++11 -fvisibility=hidden -fvisibility-inlines-hidden -fno-strict-aliasing "
          "-fno-tree-vectorize -DNDEBUG -O3 -w")
DEFS = "-DEMBREE_TARGET_SSE2 -DEMBREE_TARGET_SSE42 -DEMBREE_TARGET_AVX -DEMBREE_TARGET_AVX2 -DTASKING_INTERNAL"

# Two libraries are built from the same sources:
#   libembree3_ref.so         EMBREE_MAX_ISA=AVX2       the parity oracle and the default CPU baseline (pinned ISA set, so
#                                                       that the reference's numerics do not depend on the box's CPU)
#   libembree3_ref_avx512.so  EMBREE_MAX_ISA=AVX512SKX  timing baseline only ("AVX2 and AVX512SKX builds")
VARIANTS = {
    "avx2": dict(so="libembree3_ref.so", obj="obj", defs=DEFS, isas=("sse42", "avx", "avx2")),
    "avx512": dict(so="libembree3_ref_avx512.so", obj="obj512", defs=DEFS + " -DEMBREE_TARGET_AVX512SKX",
                   isas=("sse42", "avx", "avx2", "avx512skx")),
}

COMMON_LIBS = [
    "common/sys/sysinfo.cpp", "common/sys/alloc.cpp", "common/sys/filename.cpp", "common/sys/library.cpp",
    "common/sys/thread.cpp", "common/sys/string.cpp", "common/sys/regression.cpp", "common/sys/mutex.cpp",
    "common/sys/condition.cpp", "common/sys/barrier.cpp",
    "common/math/constants.cpp", "common/simd/sse.cpp",
    "common/lexers/stringstream.cpp", "common/lexers/tokenstream.cpp",
    "common/tasking/taskschedulerinternal.cpp",
    "common/algorithms/parallel_for.cpp", "common/algorithms/parallel_reduce.cpp",
    "common/algorithms/parallel_prefix_sum.cpp", "common/algorithms/parallel_for_for.cpp",
    "common/algorithms/parallel_for_for_prefix_sum.cpp", "common/algorithms/parallel_partition.cpp",
    "common/algorithms/parallel_sort.cpp", "common/algorithms/parallel_set.cpp",
    "common/algorithms/parallel_map.cpp", "common/algorithms/parallel_filter.cpp",
]

# kernels/CMakeLists.txt:31-96 (EMBREE_LIBRARY_FILES, lowest ISA, -DEMBREE_LOWEST_ISA)
BASE = [
    "common/device.cpp", "common/stat.cpp", "common/acceln.cpp", "common/accelset.cpp", "common/state.cpp",
    "common/rtcore.cpp", "common/rtcore_builder.cpp", "common/scene.cpp", "common/alloc.cpp",
    "common/geometry.cpp", "common/scene_user_geometry.cpp", "common/scene_instance.cpp",
    "common/scene_geometry_instance.cpp", "common/scene_triangle_mesh.cpp", "common/scene_quad_mesh.cpp",
    "common/scene_bezier_curves.cpp", "common/scene_line_segments.cpp",
    "subdiv/bezier_curve.cpp", "subdiv/bspline_curve.cpp",
    "geometry/primitive.cpp", "geometry/instance_intersector.cpp", "builders/primrefgen.cpp",
    "bvh/bvh.cpp", "bvh/bvh_statistics.cpp", "bvh/bvh4_factory.cpp", "bvh/bvh8_factory.cpp",
    "bvh/bvh_rotate.cpp", "bvh/bvh_refit.cpp", "bvh/bvh_builder.cpp", "bvh/bvh_builder_hair.cpp",
    "bvh/bvh_builder_hair_mb.cpp", "bvh/bvh_builder_morton.cpp", "bvh/bvh_builder_sah.cpp",
    "bvh/bvh_builder_twolevel.cpp", "bvh/bvh_builder_instancing.cpp", "bvh/bvh_intersector1_bvh4.cpp",
    # EMBREE_GEOMETRY_SUBDIVISION
    "common/scene_subdiv_mesh.cpp", "subdiv/tessellation_cache.cpp", "subdiv/subdivpatch1base.cpp",
    "subdiv/catmullclark_coefficients.cpp", "geometry/grid_soa.cpp", "geometry/compressed_node.cpp",
    "subdiv/subdivpatch1base_eval.cpp", "bvh/bvh_builder_subdiv.cpp",
    # EMBREE_RAY_PACKETS
    "bvh/bvh_intersector_hybrid4_bvh4.cpp", "bvh/bvh_intersector_stream_bvh4.cpp",
    "bvh/bvh_intersector_stream_filters.cpp",
]


def isa_files(isa):
    """kernels/CMakeLists.txt:100-172 (macro embree_files) for ISA in {sse42, avx, avx2, avx512skx}."""
    order = {"sse2": 0, "sse42": 1, "avx": 2, "avx2": 3, "avx512skx": 5}[isa]
    f = ["geometry/instance_intersector.cpp", "bvh/bvh_intersector1_bvh4.cpp"]
    if isa in ("sse2", "avx"):
        f += ["common/scene_triangle_mesh.cpp", "common/scene_quad_mesh.cpp", "common/scene_bezier_curves.cpp",
              "common/scene_line_segments.cpp", "bvh/bvh_refit.cpp", "bvh/bvh_builder.cpp",
              "bvh/bvh_builder_hair.cpp", "bvh/bvh_builder_hair_mb.cpp", "bvh/bvh_builder_sah.cpp",
              "bvh/bvh_builder_twolevel.cpp", "bvh/bvh_builder_instancing.cpp", "bvh/bvh_builder_subdiv.cpp"]
    if isa in ("sse2", "avx", "avx2"):
        f += ["bvh/bvh_builder_morton.cpp", "bvh/bvh_rotate.cpp", "builders/primrefgen.cpp"]
    if order > 1:
        f += ["bvh/bvh_intersector1_bvh8.cpp"]
    if isa == "avx":
        f += ["bvh/bvh.cpp", "bvh/bvh_statistics.cpp"]
    f += ["common/scene_subdiv_mesh.cpp", "geometry/grid_soa.cpp", "geometry/compressed_node.cpp",
          "subdiv/subdivpatch1base_eval.cpp"]
    f += ["bvh/bvh_intersector_hybrid4_bvh4.cpp", "bvh/bvh_intersector_stream_bvh4.cpp",
          "bvh/bvh_intersector_stream_filters.cpp"]
    if order > 1:
        f += ["bvh/bvh_intersector_hybrid8_bvh4.cpp", "bvh/bvh_intersector_hybrid4_bvh8.cpp",
              "bvh/bvh_intersector_hybrid8_bvh8.cpp", "bvh/bvh_intersector_stream_bvh8.cpp"]
    if order > 3:
        f += ["bvh/bvh_intersector_hybrid16_bvh8.cpp", "bvh/bvh_intersector_hybrid16_bvh4.cpp"]
    return f


def include_flags():
    # "../config.h" (kernels/common/default.h:51) and "../hash.h" (kernels/common/device.cpp:18) are
    # resolved through -I oracle/gen/kernels/common  =>  oracle/gen/kernels/{config,hash}.h
    return " ".join([
        "-I" + os.path.join(HERE, "gen", "kernels", "common"),
        "-I" + os.path.join(HERE, "eigen_shim"),
        "-I" + os.path.join(REF, "include"),
        "-I" + REF,
    ])


def jobs(variant="avx2"):
    js = []
    for s in COMMON_LIBS:
        js.append((os.path.join(REF, s), "c_" + s.replace("/", "_")[:-4] + ".o", FLAGS_ISA["sse2"]))
    for s in BASE:
        js.append((os.path.join(REF, "kernels", s), "k_sse2_" + s.replace("/", "_")[:-4] + ".o",
                   FLAGS_ISA["sse2"] + " -DEMBREE_LOWEST_ISA"))
    for isa in VARIANTS[variant]["isas"]:
        for s in isa_files(isa):
            js.append((os.path.join(REF, "kernels", s), "k_%s_" % isa + s.replace("/", "_")[:-4] + ".o",
                       FLAGS_ISA[isa]))
    return js


def run(cmd):
    r = subprocess.run(cmd, shell=True, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write("FAILED: %s\n%s\n" % (cmd, r.stderr[-4000:]))
        raise SystemExit(1)
    return r


def compile_one(job, variant="avx2"):
    src, obj, flags = job
    v = VARIANTS[variant]
    out = os.path.join(OUT, v["obj"], obj)
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    run("g++ %s %s %s %s -c %s -o %s" % (COMMON, v["defs"], flags, include_flags(), src, out))
    return out


def source_digest():
    h = hashlib.sha256()
    for p in [os.path.join(HERE, "refhost.cpp"), os.path.join(HERE, "refhost.h"),
              os.path.join(HERE, "eigen_shim", "Eigen", "Dense"), __file__]:
        if os.path.exists(p):
            h.update(open(p, "rb").read())
    return h.hexdigest()


def cpu_has_avx512():
    try:
        flags = open("/proc/cpuinfo").read()
        return all(f in flags for f in ("avx512f", "avx512dq", "avx512cd", "avx512bw", "avx512vl"))
    except Exception:
        return False


def build(force=False, variants=None):
    """Builds the AVX2 library (always) and the AVX512SKX timing library (ECB_REF_AVX512=0 skips it)."""
    if variants is None:
        variants = ["avx2"] + (["avx512"] if os.environ.get("ECB_REF_AVX512", "1") != "0" else [])
    ok = True
    for v in variants:
        ok = build_variant(v, force) and ok if v == "avx2" else (build_variant(v, force) or True) and ok
    return ok


def build_variant(variant, force=False):
    v = VARIANTS[variant]
    so = os.path.join(OUT, v["so"])
    if not os.path.isdir(REF):
        if os.path.exists(so):
            print("[build_ref] %s absent; using prebuilt %s" % (REF, v["so"]))
            return True
        print("[build_ref] %s absent and no prebuilt %s" % (REF, v["so"]))
        return False
    info_path = os.path.join(OUT, "BUILD_INFO.json" if variant == "avx2" else "BUILD_INFO_%s.json" % variant)
    digest = source_digest()
    if not force and os.path.exists(so) and os.path.exists(info_path):
        try:
            if json.load(open(info_path)).get("digest") == digest:
                print("[build_ref] %s up to date" % v["so"])
                return True
        except Exception:
            pass
    obj_dir = os.path.join(OUT, v["obj"])
    os.makedirs(obj_dir, exist_ok=True)
    # the include path -I gen/kernels/common must exist for "../config.h" to resolve (git keeps no empty directories)
    os.makedirs(os.path.join(HERE, "gen", "kernels", "common"), exist_ok=True)
    hash_h = os.path.join(HERE, "gen", "kernels", "hash.h")
    if not os.path.exists(hash_h):
        open(hash_h, "w").write('#define RTC_HASH "oracle-build"\n')
    t0 = time.time()
    js = jobs(variant)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(lambda j: compile_one(j, variant), js))
    print("[build_ref] %s: compiled %d reference translation units in %.0f s" % (variant, len(objs), time.time() - t0))
    # the harness is compiled against the reference's internal headers (it only *reads* the committed BVH and drives rtc* calls)
    host_o = os.path.join(obj_dir, "refhost.o")
    run("g++ %s %s %s %s -I%s -c %s -o %s" % (COMMON.replace("-fvisibility=hidden", ""), v["defs"],
                                                FLAGS_ISA["sse2"] + " -DEMBREE_LOWEST_ISA", include_flags(), HERE,
                                                os.path.join(HERE, "refhost.cpp"), host_o))
    # -Bsymbolic: the library's own references to rtc* bind locally, so that it can be dlopen'ed by a process that
    # also exports rtc* symbols (the drop-in shim libembree3_b200.so) without being interposed
    run("g++ -shared -Wl,-Bsymbolic -o %s %s %s -lpthread -ldl -Wl,-z,noexecstack" % (so, " ".join(objs), host_o))
    os.makedirs(os.path.join(OUT, "assets"), exist_ok=True)
    shutil.copyfile(os.path.join(REF, "build", "bomberman.obj"), os.path.join(OUT, "assets", "bomberman.obj"))
    json.dump({"digest": digest, "units": len(objs), "isas": ["sse2"] + list(v["isas"]),
               "tasking": "internal", "eigen": "oracle/eigen_shim stand-in",
               "compiler": subprocess.run("g++ --version", shell=True, capture_output=True, text=True).stdout.splitlines()[0],
               "seconds": round(time.time() - t0, 1)}, open(info_path, "w"), indent=1)
    print("[build_ref] wrote %s" % so)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    sys.exit(0 if ok else 1)
