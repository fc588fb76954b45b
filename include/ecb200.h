/* ecb200.h -- C ABI of the B200-native compressed-BVH traversal plugin (libecb200.so).
 *
 * This is the drop-in boundary for ONE hot path of lispbub/embree-compressed: ray traversal and
 * intersection over the compressed subdivision/displacement BVH ("cBVH").  The scene is still
 * built by the reference's CPU builder; the host application (the reference's rtcCommitScene, or
 * the embree3-compatible shim in embree_compressed_b200/csrc/rtcore_shim.cpp) hands the committed
 * BVH4 to ecbSceneImportBVH4() once and then routes its ray queries to the ecbIntersect* /
 * ecbOccluded* entry points below.  Plain pointers and sizes only; no C++ or torch types.
 *
 * All traversal runs in hand-written sm_100a CUDA kernels.  There is no CPU fallback: every
 * traversal entry point returns ECB_ERROR_NO_DEVICE when no CUDA device is usable.
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   ecbSceneImportBVH4   <- hand-off after Scene::commit_task, kernels/common/scene.cpp:632-697;
 *                           input = BVH4::root of the accel created by BVH4Factory::BVH4SubdivPatch1c{Grid,Leaf,Box}
 *                           (kernels/bvh/bvh4_factory.cpp:1653-1675), blocks = CompressedBVH (kernels/geometry/compressed.h:408-433)
 *   ecbIntersect1        <- rtcIntersect1   kernels/common/rtcore.cpp:287-304
 *   ecbIntersect16       <- rtcIntersect16  kernels/common/rtcore.cpp:370-401   (segfaults in the reference on this accel)
 *   ecbIntersect1M       <- rtcIntersect1M  kernels/common/rtcore.cpp:403-432   (same)
 *   ecbOccluded1         <- rtcOccluded1    kernels/common/rtcore.cpp:541-555
 *   ecbOccluded1M        <- rtcOccluded1M   kernels/common/rtcore.cpp:656-683
 *   per-ray semantics    <- BVHNIntersector1<4,BVH_AN1,true,SubdivPatch1c*Intersector1>::intersect/occluded
 *                           kernels/bvh/bvh_intersector1.cpp:40-209 + CompressedBVHIntersector1 kernels/geometry/compressed.h:454-756
 */
#ifndef ECB200_H
#define ECB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECB_API __attribute__((visibility("default")))

/* error codes: values follow enum RTCError (include/embree3/rtcore_device.h) where one exists */
typedef enum ecbError {
  ECB_SUCCESS = 0,
  ECB_ERROR_UNKNOWN = 1,
  ECB_ERROR_INVALID_ARGUMENT = 2,
  ECB_ERROR_INVALID_OPERATION = 3,
  ECB_ERROR_OUT_OF_MEMORY = 4,
  ECB_ERROR_UNSUPPORTED = 5, /* RTC_ERROR_UNSUPPORTED_CPU slot: scene uses a feature outside the hot path */
  ECB_ERROR_NO_DEVICE = 100, /* no usable CUDA device / kernel image: the path fails loudly, it never falls back */
  ECB_ERROR_CUDA = 101
} ecbError;

/* leaf flavour of the committed accel: which reference factory built it */
typedef enum ecbLeafKind {
  ECB_LEAF_GRID = 0, /* bvh4.compressed.grid : full-precision vertex grid, exact micro-triangles */
  ECB_LEAF_LEAF = 1, /* bvh4.compressed.leaf : 2-byte "pizza box" bilinear slabs             */
  ECB_LEAF_BOX = 2   /* bvh4.compressed.box  : voxel leaves                                   */
} ecbLeafKind;

typedef struct ecbScene ecbScene; /* opaque */

/* What the host hands over at commit time. */
typedef struct ecbImportDesc {
  const void* base;        /* added to every reference pointer: NULL for live reference memory in this process,
                              start of the buffer for a relocated image (oracle/rimg.h) */
  uint64_t root;           /* BVH4::root NodeRef (tagged pointer, kernels/bvh/bvh.h:150-400,1492) */
  uint32_t leaf_kind;      /* ecbLeafKind */
  uint32_t proj_col_major; /* storage order of Eigen::Matrix3f CompressedBVH::proj in the host build (1 = Eigen default) */
  uint32_t subdiv_level;   /* S of rtcSetSceneLevels (informational: bytes per micro-triangle) */
  uint32_t comp_level;     /* C of rtcSetSceneLevels; verified against every block's `elems` */
  float bounds[6];         /* scene bounds (rtcGetSceneBounds), lower xyz then upper xyz; informational */
  uint64_t image_bytes;    /* size of the relocated image when base != NULL: every offset read from it is range-checked
                              (a truncated or corrupt image is rejected, never dereferenced); 0 for live memory */
} ecbImportDesc;

typedef struct ecbSceneInfo {
  uint32_t leaf_kind, subdiv_level, comp_level;
  uint32_t n_top_nodes, n_blocks, top_depth;
  uint32_t block_stride;       /* bytes per block in the device format */
  uint32_t top_node_bytes;     /* bytes per top-level node in the device format */
  uint64_t device_bytes;       /* total bytes of the device scene */
  uint64_t reference_bytes;    /* bytes the same accel occupies in the reference layout (nodes+blocks, no allocator slack) */
  uint64_t micro_triangles;    /* 2 * 4^C * n_blocks */
  float bounds[6];
  int32_t device;              /* CUDA device ordinal the scene is resident on, -1 if host only */
} ecbSceneInfo;

/* An instancing scene: the host's user-geometry BVH4 over RTC_GEOMETRY_TYPE_INSTANCE geometries
 * (device config object_accel=bvh4.object) plus, per instance, what kernels/common/scene_instance.h holds. */
typedef struct ecbInstanceDesc {
  const void* base;          /* as in ecbImportDesc */
  uint64_t root;             /* BVH4::root of the accel created by BVH4Factory::BVH4UserGeometry (bvh4_factory.cpp:1678) */
  uint32_t n_instances;
  uint32_t reserved;
  const uint32_t* geomIDs;   /* geomID of each instance geometry inside the instancing scene (becomes hit.instID) */
  const float* world2local;  /* 12 floats per instance: l.vx, l.vy, l.vz, p of Instance::world2local0 */
  ecbScene* const* objects;  /* imported scene each instance refers to (entries may repeat; retained by the new scene) */
  float bounds[6];
  uint64_t image_bytes;      /* as in ecbImportDesc */
} ecbInstanceDesc;

/* Triangle meshes in the same scene (SURVEY 8f-1, config 1's ground plane): what the host hands over for the accel
 * that AccelN::intersect visits BEFORE the compressed accel (kernels/common/acceln.cpp:51-56, creation order
 * kernels/common/scene.cpp:650-654). */
typedef struct ecbTriangleDesc {
  const void* base;  /* as in ecbImportDesc */
  uint64_t root;     /* BVH4::root of the accel created by BVH4Factory::BVH4Triangle4v (kernels/bvh/bvh4_factory.cpp:1368-1394):
                        AlignedNodes, leaves = arrays of TriangleMv<4> blocks (kernels/geometry/trianglev.h, 176 bytes) */
  uint64_t image_bytes; /* as in ecbImportDesc */
} ecbTriangleDesc;

/* ---- scene ---------------------------------------------------------------------------------------- */
/* Converts the reference's committed compressed BVH4 into the device scene format (host work only). */
ECB_API ecbError ecbSceneImportBVH4(const ecbImportDesc* desc, ecbScene** scene_out);
/* Builds the device structure of an instancing scene; ray queries on it restate FastInstanceIntersectorN::intersect1 /
 * occluded1 (kernels/geometry/instance_intersector.cpp:45-83) over the non-robust BVH4 traversal of the instance level. */
ECB_API ecbError ecbSceneImportInstances(const ecbInstanceDesc* desc, ecbScene** scene_out);
/* Attaches the scene's triangle accel to an imported (not yet uploaded, non-instancing) scene.  Ray queries then
 * restate AccelN::intersect / occluded: BVHNIntersector1<4,BVH_AN1,true,ArrayIntersector1<TriangleMvIntersector1Pluecker<4,4,true>>>
 * (kernels/bvh/bvh_intersector1_bvh4.cpp:44, kernels/geometry/triangle_intersector_pluecker.h:77-127) first, the
 * compressed accel second, on the same ray.  Triangle hits carry the real geometry normal in Ng. */
ECB_API ecbError ecbSceneAttachTriangles4v(ecbScene* scene, const ecbTriangleDesc* desc);
/* Convenience: import from an RIMG file / buffer (header carries root, leaf kind, S, C, bounds). */
ECB_API ecbError ecbSceneImportImage(const void* image, size_t bytes, ecbScene** scene_out);
ECB_API ecbError ecbSceneImportImageFile(const char* path, ecbScene** scene_out);
/* Native serialisation of the converted scene (the reference has none). */
ECB_API ecbError ecbSceneSave(const ecbScene* scene, const char* path);
ECB_API ecbError ecbSceneLoad(const char* path, ecbScene** scene_out);
/* Makes the scene resident on CUDA device `device` (cudaMalloc + copy; -1 = the calling thread's current device).
 * Traversal calls do this lazily.  A scene may be resident on several devices: every call adds one copy. */
ECB_API ecbError ecbSceneUpload(ecbScene* scene, int device);
/* Multi-GPU in ONE process (the reference's analogue is the tile-parallel parallel_for over screen tiles,
 * tutorials/viewer/viewer_device.cpp:548-554): replicates the scene on every listed device.  From then on a host-buffer
 * call (ecbIntersect1M / ecbOccluded1M / ...) of at least 256 Ki rays is cut into one contiguous ray range per device;
 * every range runs its own upload -> traversal -> download pipeline over its own PCIe link and the results land in the
 * caller's array -- no gather step.  Device-buffer calls run on the device that owns the ray buffer. */
ECB_API ecbError ecbSceneUploadMulti(ecbScene* scene, const int* devices, uint32_t n_devices);
/* Devices the scene is resident on (first = primary); returns their number. */
ECB_API uint32_t ecbSceneResidentDevices(const ecbScene* scene, int* devices_out, uint32_t capacity);
ECB_API ecbError ecbSceneGetInfo(const ecbScene* scene, ecbSceneInfo* info);
ECB_API void ecbSceneRelease(ecbScene* scene);

/* ---- block construction on the GPU (the step in front of the path) -------------------------------
 * Builds n compressed blocks from their displaced vertex grids and frames, exactly as CompressedBVH::CompressedBVH does
 * (kernels/geometry/compressed.h:128-337: homography, box hierarchy, Node::setAABB quantization against re-expanded
 * parents, frustum box, conservative world bounds, pizza-box leaves), in the device format.  What stays on the host is
 * what needs the application: subdivision + displacement callbacks (evalGrid) and the 3x3 block frame (compressed.h:91-126).
 * All pointers are HOST pointers; the work runs on the GPU (no CPU fallback). */
typedef struct ecbBuildDesc {
  uint32_t leaf_kind;        /* ecbLeafKind */
  uint32_t comp_level;       /* C: blocks of 2^C x 2^C cells, (2^C+1)^2 vertices */
  uint32_t n_blocks;
  uint32_t reserved;
  const float* grids;        /* n x (2^C+1)^2 x 3: displaced vertices of each block, row-major y*width+x */
  const float* space;        /* n x 9: LinearSpace3f world->local of each block (vx,vy,vz columns), compressed.h:125-126 */
  const float* world;        /* n x 9: LinearSpace3f(vx,vy,vz) it was inverted from */
  const float* uv;           /* n x 4: patch uv of vertex (0,0) and uv extent of the block (compressed.h:84-86) */
  const uint32_t* geomIDs;   /* n */
  const uint32_t* primIDs;   /* n */
} ecbBuildDesc;
/* bytes per block of the device format for a leaf kind / compression level (what blocks_out must hold per block) */
ECB_API uint32_t ecbBlockStride(uint32_t leaf_kind, uint32_t comp_level);
/* blocks_out: n x ecbBlockStride bytes; bounds_out: n x 6 floats (lower xyz, upper xyz: the PrimRef bounds the host's
 * top-level BVH4 builder needs, compressed.h:252-264); patch_ok_out: optional, n flags (homography applicable). */
ECB_API ecbError ecbBuildBlocks(const ecbBuildDesc* desc, void* blocks_out, float* bounds_out, uint32_t* patch_ok_out);
/* Host copy of a scene's converted blocks (n_blocks x block_stride bytes, ecbSceneGetInfo), and the reverse: install
 * blocks built by ecbBuildBlocks into a scene that is not yet resident (same count, kind and level). */
ECB_API ecbError ecbSceneCopyBlocks(const ecbScene* scene, void* blocks_out, size_t capacity_bytes);
ECB_API ecbError ecbSceneSetBlocks(ecbScene* scene, const void* blocks, size_t bytes);

/* ---- traversal, host buffers (drop-in semantics of the rtc* calls named above) ---------------------
 * rayhit / ray point to RTCRayHit (80 B) / RTCRay (48 B) structs, include/embree3/rtcore_ray.h:26-64, mutated in
 * place exactly like the reference does: on a hit tfar,Ng,u,v,primID,geomID,instID are overwritten, on a miss
 * nothing is written; occluded sets tfar=-inf.  Stream rays with tnear > tfar are inactive.  instID is what
 * RTCIntersectContext::instID[0] holds (RTC_INVALID_GEOMETRY_ID at top level).
 * The arrays may live in any host memory (malloc'ed, stack, pinned), 4-byte aligned, byteStride a multiple of 4 that is
 * at least the struct size.  Only the fields this path reads (org, tnear, dir, tfar: 32 bytes per ray) travel to the
 * device and only result records travel back (20-32 bytes per ray); time, mask, id and flags are never read
 * (the reference's intersectors for this accel do not read them either; EMBREE_RAY_MASK is off). */
ECB_API ecbError ecbIntersect1(ecbScene* scene, uint32_t instID, void* rayhit);
ECB_API ecbError ecbIntersect1M(ecbScene* scene, uint32_t instID, void* rayhit, uint32_t M, size_t byteStride);
ECB_API ecbError ecbIntersect16(const int* valid, ecbScene* scene, uint32_t instID, void* rayhit16);
ECB_API ecbError ecbOccluded1(ecbScene* scene, uint32_t instID, void* ray);
ECB_API ecbError ecbOccluded1M(ecbScene* scene, uint32_t instID, void* ray, uint32_t M, size_t byteStride);

/* SoA streams: ecbIntersectNp <- rtcIntersectNp, ecbOccludedNp <- rtcOccludedNp (include/embree3/rtcore_scene.h:118,142;
 * kernels/common/rtcore.cpp:507-539,744-772).  The structs repeat RTCRayNp / RTCHitNp / RTCRayHitNp
 * (include/embree3/rtcore_ray.h:189-228) field for field, so an embree3 caller's struct can be passed as it is.  The eight
 * ray arrays this path reads (org, tnear, dir, tfar) are packed by the library's host pool from eight sequential streams (32
 * bytes read per ray instead of the 80-byte records of the AoS call); hits are written into the hit arrays and tfar, a miss
 * writes nothing; time, mask, id and flags are never read. */
typedef struct ecbRayNp {
  float *org_x, *org_y, *org_z, *tnear;
  float *dir_x, *dir_y, *dir_z, *time;
  float* tfar;
  uint32_t *mask, *id, *flags;
} ecbRayNp;
typedef struct ecbHitNp {
  float *Ng_x, *Ng_y, *Ng_z;
  float *u, *v;
  uint32_t *primID, *geomID, *instID;
} ecbHitNp;
typedef struct ecbRayHitNp {
  ecbRayNp ray;
  ecbHitNp hit;
} ecbRayHitNp;
ECB_API ecbError ecbIntersectNp(ecbScene* scene, uint32_t instID, const ecbRayHitNp* rayhit, uint64_t N);
ECB_API ecbError ecbOccludedNp(ecbScene* scene, uint32_t instID, const ecbRayNp* ray, uint64_t N);

/* Opt-in EXACT shadow query (NOT the reference's semantics): the ray is occluded iff the closest-hit
 * traversal would accept a leaf hit inside [tnear, tfar] -- the blocks are entered instead of stopping at their world
 * boxes.  Same RTCRay records, tfar = -inf when occluded.  Not available on instancing scenes. */
ECB_API ecbError ecbOccludedExact1M(ecbScene* scene, uint32_t instID, void* ray, uint32_t M, size_t byteStride);
ECB_API ecbError ecbOccludedExactDevice(ecbScene* scene, uint32_t instID, void* d_ray, uint64_t M, size_t byteStride, void* stream);

/* ---- traversal, device-resident buffers (extension; what the benchmark's `value` times) ------------
 * d_rayhit / d_ray are device pointers on the scene's device; stream is a cudaStream_t (NULL = default stream).
 * Asynchronous: returns after the launch. */
ECB_API ecbError ecbIntersectDevice(ecbScene* scene, uint32_t instID, void* d_rayhit, uint64_t M, size_t byteStride, void* stream);
ECB_API ecbError ecbOccludedDevice(ecbScene* scene, uint32_t instID, void* d_ray, uint64_t M, size_t byteStride, void* stream);
/* Like ecbIntersectDevice, and additionally writes, per ray, the index of the block that produced the final hit
 * (0xFFFFFFFF for a miss / unchanged) to d_block_out[M] when it is not NULL.  Used by the benchmark's bounce-ray
 * generator (the reference returns a dummy Ng, compressed.h:574, so normals must come from the block frame). */
ECB_API ecbError ecbIntersectDeviceEx(ecbScene* scene, uint32_t instID, void* d_rayhit, uint64_t M, size_t byteStride,
                                      uint32_t* d_block_out, void* stream);
/* Instrumented variant (same results, slower): adds this launch's work to d_counters[6] = {rays traced, top-level
 * nodes tested, blocks entered, quantized nodes decoded, leaf cells tested, hits} -- the per-ray counts the roofline's
 * algorithmic-bytes figure is built from.  d_counters is a device pointer the caller zeroes. */
ECB_API ecbError ecbIntersectDeviceStats(ecbScene* scene, uint32_t instID, void* d_rayhit, uint64_t M, size_t byteStride,
                                         uint64_t* d_counters, void* stream);
/* Compact result records of a device-buffer call (device pointers on the scene's device, one entry per ray): what a
 * multi-GPU frame gather moves instead of the 80-byte RTCRayHit records.  The ray records are only read. */
typedef struct ecbDeviceHits {
  uint32_t* geomID; /* geomID of the hit, 0xFFFFFFFF for a miss / inactive ray; any-hit queries: 0 = occluded */
  float* hit;       /* 4 floats per ray: tfar, u, v, primID (bit pattern); written for hits only */
  uint32_t* instID; /* instancing scenes: instID[0] of the hit (geomID of the instance); may be NULL otherwise */
  float* Ng;        /* reserved for mixed scenes (3 floats per ray); may be NULL otherwise */
} ecbDeviceHits;
ECB_API ecbError ecbIntersectDeviceCompact(ecbScene* scene, uint32_t instID, const void* d_rayhit, uint64_t M, size_t byteStride,
                                           const ecbDeviceHits* out, void* stream);
ECB_API ecbError ecbOccludedDeviceCompact(ecbScene* scene, uint32_t instID, const void* d_ray, uint64_t M, size_t byteStride,
                                          const ecbDeviceHits* out, void* stream);

/* SoA ray input on the device (north_star: "SoA ray data coalesced"): d_soa = org_x, org_y, org_z, tnear, dir_x, dir_y, dir_z,
 * tfar -- eight device arrays of M floats; consecutive lanes load consecutive words.  Plain compressed-accel scenes. */
ECB_API ecbError ecbIntersectDeviceSoA(ecbScene* scene, uint32_t instID, const float* const d_soa[8], uint64_t M,
                                       const ecbDeviceHits* out, void* stream);

/* ---- multi-GPU, one process per GPU: the hit-record gather over NVLink -------------------------------
 * The gathering rank allocates the frame's result arrays with ecbDeviceAlloc and publishes their 64-byte handles
 * (ecbIpcExport, any transport: torch.distributed broadcast, a file, ...); every other rank maps them once (ecbIpcImport:
 * CUDA IPC, peer access over NVLink/NVSwitch) and, after its traversal launches, copies its shard's compact records
 * straight into place with ecbPeerCopyAsync on its own stream -- no collective, no staging on the gathering GPU. */
ECB_API ecbError ecbDeviceAlloc(void** d_ptr, size_t bytes);
ECB_API void ecbDeviceFree(void* d_ptr);
ECB_API ecbError ecbIpcExport(void* d_ptr, void* handle64_out);
ECB_API ecbError ecbIpcImport(const void* handle64, void** d_ptr_out);
ECB_API void ecbIpcClose(void* d_ptr);
ECB_API ecbError ecbPeerCopyAsync(void* d_dst, const void* d_src, size_t bytes, void* stream);

/* World-space unit normal of a block's local frame (third row of CompressedBVH::space), 3 floats per block, host copy. */
ECB_API ecbError ecbSceneBlockNormals(const ecbScene* scene, float* normals_out, size_t capacity_floats);

/* Pinned host memory.  The host-buffer entry points do not need it (see above); it matters for the "compact" and
 * "copy" transfer strategies, which DMA the caller's records directly. */
ECB_API ecbError ecbHostAlloc(void** ptr, size_t bytes);
ECB_API void ecbHostFree(void* ptr);

/* Transfer strategy the host-buffer entry points use for a large call on a compressed-accel scene, given the kind of
 * caller memory: "packed" (32 B per ray up, 20 B back), "compact" (full records up by DMA, 20 B back) or "copy" (full
 * records both ways).  Chosen from ECB200_HOST_PATH, else from the host cores this process can use (csrc/ecb_api.cpp).
 * Diagnostic only (bench.py reports it with its byte counts); the returned string is static. */
ECB_API const char* ecbHostPathName(int caller_memory_is_pinned);

/* The closest-hit traversal exists as two kernels that return the same bits: ECB_KERNEL_LANE (one lane per ray, best for
 * coherent rays on shallow blocks) and ECB_KERNEL_QUAD (four lanes per ray, one child box per lane: best for incoherent
 * rays and displaced geometry).  Process-wide; -1 restores the default (or $ECB200_KERNEL = lane | quad). */
#define ECB_KERNEL_LANE 0
#define ECB_KERNEL_QUAD 1
ECB_API void ecbSetTraversalKernel(int mode);
ECB_API int ecbGetTraversalKernel(void);

/* Libraries built with -DECB_BOUNDS_CHECK (embree_compressed_b200.build.build_variant("check", ["ECB_BOUNDS_CHECK=1"]))
 * check every index the kernels derive from scene data against the array it indexes and count violations in eight classes
 * (top-level stack, top-level node, block, block stack, in-block offset, instance tables, triangle blocks, ray index).
 * Returns 1 and the counters for such a build, 0 (and zeros) for the shipping build, -1 on a CUDA error.  compute-sanitizer
 * is not available on the B200 pool; tests/test_bounds_check_gpu.py runs the golden cases on the checking build instead. */
ECB_API int ecbBoundsCheckFailures(uint64_t counts_out[8]);

/* Measurement support (not on the traversal path): read bandwidth in GB/s of a device buffer of `working_set_bytes`
 * streamed `passes` times with 128-bit L1-bypassing loads from every SM.  32 MB x 64 passes measures the L2 the scene
 * traffic of the traversal kernels is served from; 4 GB x 1 pass measures HBM.  The roofline's L2 denominator
 * (profiles/measured_l2.json) comes from this call. */
ECB_API ecbError ecbProbeReadBandwidth(size_t working_set_bytes, int passes, double* gbs_out);

/* Number of kernels this library has launched since load (bench.py's gpu_launches claim). */
ECB_API uint64_t ecbLaunchCount(void);
ECB_API const char* ecbErrorString(ecbError e);
ECB_API const char* ecbLastErrorMessage(void);

#ifdef __cplusplus
}
#endif
#endif /* ECB200_H */
