/* rtcore_subset.h -- the part of the embree3 C API (Embree 3.0.0) that the compressed-BVH hot path sits behind,
 * restated for callers that build without the reference's include/embree3/ tree (tests, demos, the shim itself).
 *
 * An application written against the reference's own headers links against libembree3_b200.so unchanged: every
 * name, enum value and struct layout below is an ABI fact of the reference
 * (include/embree3/rtcore_{common,device,buffer,ray,geometry,scene}.h) and is verified against the compiled
 * reference by tests/test_shim_gpu.py.  Nothing here adds API except the two *_b200 entries at the bottom.
 */
#ifndef EMBREE3_B200_RTCORE_SUBSET_H
#define EMBREE3_B200_RTCORE_SUBSET_H

#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

#if !defined(RTC_API)
#define RTC_API __attribute__((visibility("default")))
#endif
#define RTC_INVALID_GEOMETRY_ID ((unsigned int)-1)
#define RTC_MAX_INSTANCE_LEVEL_COUNT 1

typedef struct RTCDeviceTy* RTCDevice;
typedef struct RTCSceneTy* RTCScene;
typedef struct RTCGeometryTy* RTCGeometry;

enum RTCError { RTC_ERROR_NONE = 0, RTC_ERROR_UNKNOWN = 1, RTC_ERROR_INVALID_ARGUMENT = 2, RTC_ERROR_INVALID_OPERATION = 3,
                RTC_ERROR_OUT_OF_MEMORY = 4, RTC_ERROR_UNSUPPORTED_CPU = 5, RTC_ERROR_CANCELLED = 6 };
enum RTCGeometryType { RTC_GEOMETRY_TYPE_TRIANGLE = 0, RTC_GEOMETRY_TYPE_QUAD = 1, RTC_GEOMETRY_TYPE_SUBDIVISION = 8,
                       RTC_GEOMETRY_TYPE_USER = 120, RTC_GEOMETRY_TYPE_INSTANCE = 121 };
enum RTCBufferType { RTC_BUFFER_TYPE_INDEX = 0, RTC_BUFFER_TYPE_VERTEX = 1, RTC_BUFFER_TYPE_VERTEX_ATTRIBUTE = 2,
                     RTC_BUFFER_TYPE_FACE = 16, RTC_BUFFER_TYPE_LEVEL = 17, RTC_BUFFER_TYPE_EDGE_CREASE_INDEX = 18,
                     RTC_BUFFER_TYPE_EDGE_CREASE_WEIGHT = 19, RTC_BUFFER_TYPE_VERTEX_CREASE_INDEX = 20,
                     RTC_BUFFER_TYPE_VERTEX_CREASE_WEIGHT = 21, RTC_BUFFER_TYPE_HOLE = 22 };
enum RTCFormat { RTC_FORMAT_UINT = 0x5001, RTC_FORMAT_UINT2 = 0x5002, RTC_FORMAT_UINT3 = 0x5003, RTC_FORMAT_UINT4 = 0x5004,
                 RTC_FORMAT_FLOAT = 0x9001, RTC_FORMAT_FLOAT2 = 0x9002, RTC_FORMAT_FLOAT3 = 0x9003, RTC_FORMAT_FLOAT4 = 0x9004,
                 RTC_FORMAT_FLOAT3X4_ROW_MAJOR = 0x9134, RTC_FORMAT_FLOAT3X4_COLUMN_MAJOR = 0x9234 };
enum RTCBuildQuality { RTC_BUILD_QUALITY_LOW = 0, RTC_BUILD_QUALITY_MEDIUM = 1, RTC_BUILD_QUALITY_HIGH = 2, RTC_BUILD_QUALITY_REFIT = 3 };
enum RTCSceneFlags { RTC_SCENE_FLAG_NONE = 0, RTC_SCENE_FLAG_DYNAMIC = 1, RTC_SCENE_FLAG_COMPACT = 2, RTC_SCENE_FLAG_ROBUST = 4,
                     RTC_SCENE_FLAG_CONTEXT_FILTER_FUNCTION = 8 };
enum RTCIntersectContextFlags { RTC_INTERSECT_CONTEXT_FLAG_NONE = 0, RTC_INTERSECT_CONTEXT_FLAG_INCOHERENT = 0,
                                RTC_INTERSECT_CONTEXT_FLAG_COHERENT = 1 };
enum RTCSubdivisionMode { RTC_SUBDIVISION_MODE_NO_BOUNDARY = 0, RTC_SUBDIVISION_MODE_SMOOTH_BOUNDARY = 1,
                          RTC_SUBDIVISION_MODE_PIN_CORNERS = 2, RTC_SUBDIVISION_MODE_PIN_BOUNDARY = 3, RTC_SUBDIVISION_MODE_PIN_ALL = 4 };

/* rtcore_ray.h:26-64 */
struct __attribute__((aligned(16))) RTCRay { float org_x, org_y, org_z, tnear, dir_x, dir_y, dir_z, time, tfar; unsigned int mask, id, flags; };
struct __attribute__((aligned(16))) RTCHit { float Ng_x, Ng_y, Ng_z, u, v; unsigned int primID, geomID, instID[RTC_MAX_INSTANCE_LEVEL_COUNT]; };
struct RTCRayHit { struct RTCRay ray; struct RTCHit hit; };
#define RTC_B200_PACKET(N)                                                                                                     \
  struct __attribute__((aligned(4 * N))) RTCRay##N { float org_x[N], org_y[N], org_z[N], tnear[N], dir_x[N], dir_y[N], dir_z[N], time[N], tfar[N]; \
                                                     unsigned int mask[N], id[N], flags[N]; };                                 \
  struct __attribute__((aligned(4 * N))) RTCHit##N { float Ng_x[N], Ng_y[N], Ng_z[N], u[N], v[N]; unsigned int primID[N], geomID[N],   \
                                                     instID[RTC_MAX_INSTANCE_LEVEL_COUNT][N]; };                              \
  struct RTCRayHit##N { struct RTCRay##N ray; struct RTCHit##N hit; };
RTC_B200_PACKET(4)
RTC_B200_PACKET(8)
RTC_B200_PACKET(16)

/* streams of N-wide SoA packets (opaque: 12 / 20 arrays of N words, rtcore_ray.h:189-191) and pointer-SoA streams (:192-228) */
struct RTCRayN;
struct RTCRayHitN;
struct RTCRayNp { float *org_x, *org_y, *org_z, *tnear, *dir_x, *dir_y, *dir_z, *time, *tfar; unsigned int *mask, *id, *flags; };
struct RTCHitNp { float *Ng_x, *Ng_y, *Ng_z, *u, *v; unsigned int *primID, *geomID, *instID[RTC_MAX_INSTANCE_LEVEL_COUNT]; };
struct RTCRayHitNp { struct RTCRayNp ray; struct RTCHitNp hit; };

struct __attribute__((aligned(16))) RTCBounds { float lower_x, lower_y, lower_z, align0, upper_x, upper_y, upper_z, align1; };
struct RTCFilterFunctionNArguments;
typedef void (*RTCFilterFunctionN)(const struct RTCFilterFunctionNArguments* args);
struct RTCIntersectContext { enum RTCIntersectContextFlags flags; RTCFilterFunctionN filter; unsigned int instID[RTC_MAX_INSTANCE_LEVEL_COUNT]; };
static inline void rtcInitIntersectContext(struct RTCIntersectContext* c) {
  c->flags = RTC_INTERSECT_CONTEXT_FLAG_INCOHERENT; c->filter = NULL; c->instID[0] = (unsigned int)-1;
}
struct RTCDisplacementFunctionNArguments { void* geometryUserPtr; RTCGeometry geometry; unsigned int primID; unsigned int timeStep;
                                           const float *u, *v, *Ng_x, *Ng_y, *Ng_z; float *P_x, *P_y, *P_z; unsigned int N; };
typedef void (*RTCDisplacementFunctionN)(const struct RTCDisplacementFunctionNArguments* args);
struct RTCInterpolateArguments { RTCGeometry geometry; unsigned int primID; float u, v; enum RTCBufferType bufferType; unsigned int bufferSlot;
                                 float *P, *dPdu, *dPdv, *ddPdudu, *ddPdvdv, *ddPdudv; unsigned int valueCount; };
typedef void (*RTCErrorFunction)(void* userPtr, enum RTCError code, const char* str);

/* device (rtcore_device.h) */
RTC_API RTCDevice rtcNewDevice(const char* config);
RTC_API void rtcRetainDevice(RTCDevice device);
RTC_API void rtcReleaseDevice(RTCDevice device);
RTC_API enum RTCError rtcGetDeviceError(RTCDevice device);
RTC_API void rtcSetDeviceErrorFunction(RTCDevice device, RTCErrorFunction error, void* userPtr);
/* geometry (rtcore_geometry.h) */
RTC_API RTCGeometry rtcNewGeometry(RTCDevice device, enum RTCGeometryType type);
RTC_API void rtcRetainGeometry(RTCGeometry geometry);
RTC_API void rtcReleaseGeometry(RTCGeometry geometry);
RTC_API void rtcCommitGeometry(RTCGeometry geometry);
RTC_API void rtcEnableGeometry(RTCGeometry geometry);
RTC_API void rtcDisableGeometry(RTCGeometry geometry);
RTC_API void rtcSetGeometryTimeStepCount(RTCGeometry geometry, unsigned int timeStepCount);
RTC_API void rtcSetGeometryVertexAttributeCount(RTCGeometry geometry, unsigned int vertexAttributeCount);
RTC_API void rtcSetGeometryBuildQuality(RTCGeometry geometry, enum RTCBuildQuality quality);
RTC_API void rtcSetSharedGeometryBuffer(RTCGeometry geometry, enum RTCBufferType type, unsigned int slot, enum RTCFormat format,
                                        const void* ptr, size_t byteOffset, size_t byteStride, size_t itemCount);
RTC_API void* rtcSetNewGeometryBuffer(RTCGeometry geometry, enum RTCBufferType type, unsigned int slot, enum RTCFormat format,
                                      size_t byteStride, size_t itemCount);
RTC_API void* rtcGetGeometryBufferData(RTCGeometry geometry, enum RTCBufferType type, unsigned int slot);
RTC_API void rtcUpdateGeometryBuffer(RTCGeometry geometry, enum RTCBufferType type, unsigned int slot);
RTC_API void rtcSetGeometryUserData(RTCGeometry geometry, void* ptr);
RTC_API void* rtcGetGeometryUserData(RTCGeometry geometry);
RTC_API void rtcSetGeometryTessellationRate(RTCGeometry geometry, float tessellationRate);
RTC_API void rtcSetGeometryTopologyCount(RTCGeometry geometry, unsigned int topologyCount);
RTC_API void rtcSetGeometrySubdivisionMode(RTCGeometry geometry, unsigned int topologyID, enum RTCSubdivisionMode mode);
RTC_API void rtcSetGeometryVertexAttributeTopology(RTCGeometry geometry, unsigned int vertexAttributeID, unsigned int topologyID);
RTC_API void rtcSetGeometryDisplacementFunction(RTCGeometry geometry, RTCDisplacementFunctionN displacement);
RTC_API void rtcInterpolate(const struct RTCInterpolateArguments* args);
RTC_API void rtcSetGeometryInstancedScene(RTCGeometry geometry, RTCScene scene);
RTC_API void rtcSetGeometryTransform(RTCGeometry geometry, unsigned int timeStep, enum RTCFormat format, const void* xfm);
/* scene (rtcore_scene.h) */
RTC_API RTCScene rtcNewScene(RTCDevice device);
RTC_API void rtcRetainScene(RTCScene scene);
RTC_API void rtcReleaseScene(RTCScene scene);
RTC_API unsigned int rtcAttachGeometry(RTCScene scene, RTCGeometry geometry);
RTC_API void rtcAttachGeometryByID(RTCScene scene, RTCGeometry geometry, unsigned int geomID);
RTC_API void rtcDetachGeometry(RTCScene scene, unsigned int geomID);
RTC_API RTCGeometry rtcGetGeometry(RTCScene scene, unsigned int geomID);
RTC_API void rtcSetSceneLevels(RTCScene scene, unsigned subdivisionLevel, unsigned compressionLevel); /* rtcore_scene.h:65 */
RTC_API void rtcCommitScene(RTCScene scene);
RTC_API void rtcJoinCommitScene(RTCScene scene);
RTC_API void rtcSetSceneBuildQuality(RTCScene scene, enum RTCBuildQuality quality);
RTC_API void rtcSetSceneFlags(RTCScene scene, enum RTCSceneFlags flags);
RTC_API enum RTCSceneFlags rtcGetSceneFlags(RTCScene scene);
RTC_API void rtcGetSceneBounds(RTCScene scene, struct RTCBounds* bounds_o);
RTC_API void rtcIntersect1(RTCScene scene, struct RTCIntersectContext* context, struct RTCRayHit* rayhit);
RTC_API void rtcIntersect4(const int* valid, RTCScene scene, struct RTCIntersectContext* context, struct RTCRayHit4* rayhit);
RTC_API void rtcIntersect8(const int* valid, RTCScene scene, struct RTCIntersectContext* context, struct RTCRayHit8* rayhit);
RTC_API void rtcIntersect16(const int* valid, RTCScene scene, struct RTCIntersectContext* context, struct RTCRayHit16* rayhit);
RTC_API void rtcIntersect1M(RTCScene scene, struct RTCIntersectContext* context, struct RTCRayHit* rayhit, unsigned int M, size_t byteStride);
RTC_API void rtcIntersect1Mp(RTCScene scene, struct RTCIntersectContext* context, struct RTCRayHit** rayhit, unsigned int M);
RTC_API void rtcIntersectNM(RTCScene scene, struct RTCIntersectContext* context, struct RTCRayHitN* rayhit, unsigned int N, unsigned int M, size_t byteStride);
RTC_API void rtcIntersectNp(RTCScene scene, struct RTCIntersectContext* context, const struct RTCRayHitNp* rayhit, unsigned int N);
RTC_API void rtcOccluded1(RTCScene scene, struct RTCIntersectContext* context, struct RTCRay* ray);
RTC_API void rtcOccluded4(const int* valid, RTCScene scene, struct RTCIntersectContext* context, struct RTCRay4* ray);
RTC_API void rtcOccluded8(const int* valid, RTCScene scene, struct RTCIntersectContext* context, struct RTCRay8* ray);
RTC_API void rtcOccluded16(const int* valid, RTCScene scene, struct RTCIntersectContext* context, struct RTCRay16* ray);
RTC_API void rtcOccluded1M(RTCScene scene, struct RTCIntersectContext* context, struct RTCRay* ray, unsigned int M, size_t byteStride);
RTC_API void rtcOccluded1Mp(RTCScene scene, struct RTCIntersectContext* context, struct RTCRay** ray, unsigned int M);
RTC_API void rtcOccludedNM(RTCScene scene, struct RTCIntersectContext* context, struct RTCRayN* ray, unsigned int N, unsigned int M, size_t byteStride);
RTC_API void rtcOccludedNp(RTCScene scene, struct RTCIntersectContext* context, const struct RTCRayNp* ray, unsigned int N);

/* ---- the rest of the embree3 surface (include/embree3/rtcore_{device,buffer,geometry,scene,builder}.h): object plumbing
 * that never touches a ray.  Forwarded to the host Embree unchanged, except that (a) rtcGetDeviceProperty answers for the
 * device path where it differs (packets and streams are served, filter callbacks and user geometries are not) and
 * (b) installing a filter callback is refused with RTC_ERROR_INVALID_OPERATION: host callbacks cannot run inside the
 * CUDA traversal, and silently ignoring them would change results. ---- */
typedef struct RTCBufferTy* RTCBuffer;
typedef struct RTCBVHTy* RTCBVH;
typedef struct RTCThreadLocalAllocatorTy* RTCThreadLocalAllocator;
enum RTCDeviceProperty { RTC_DEVICE_PROPERTY_VERSION = 0, RTC_DEVICE_PROPERTY_VERSION_MAJOR = 1, RTC_DEVICE_PROPERTY_VERSION_MINOR = 2,
                         RTC_DEVICE_PROPERTY_VERSION_PATCH = 3, RTC_DEVICE_PROPERTY_NATIVE_RAY4_SUPPORTED = 32,
                         RTC_DEVICE_PROPERTY_NATIVE_RAY8_SUPPORTED = 33, RTC_DEVICE_PROPERTY_NATIVE_RAY16_SUPPORTED = 34,
                         RTC_DEVICE_PROPERTY_RAY_STREAM_SUPPORTED = 35, RTC_DEVICE_PROPERTY_RAY_MASK_SUPPORTED = 64,
                         RTC_DEVICE_PROPERTY_BACKFACE_CULLING_ENABLED = 65, RTC_DEVICE_PROPERTY_FILTER_FUNCTION_SUPPORTED = 66,
                         RTC_DEVICE_PROPERTY_IGNORE_INVALID_RAYS_ENABLED = 67, RTC_DEVICE_PROPERTY_TRIANGLE_GEOMETRY_SUPPORTED = 96,
                         RTC_DEVICE_PROPERTY_QUAD_GEOMETRY_SUPPORTED = 97, RTC_DEVICE_PROPERTY_SUBDIVISION_GEOMETRY_SUPPORTED = 98,
                         RTC_DEVICE_PROPERTY_CURVE_GEOMETRY_SUPPORTED = 99, RTC_DEVICE_PROPERTY_USER_GEOMETRY_SUPPORTED = 100,
                         RTC_DEVICE_PROPERTY_TASKING_SYSTEM = 128, RTC_DEVICE_PROPERTY_JOIN_COMMIT_SUPPORTED = 129 };
typedef bool (*RTCMemoryMonitorFunction)(void* ptr, ssize_t bytes, bool post);
typedef bool (*RTCProgressMonitorFunction)(void* ptr, double n);
struct RTCBoundsFunctionArguments;
struct RTCIntersectFunctionNArguments;
struct RTCOccludedFunctionNArguments;
struct RTCInterpolateNArguments;
struct RTCLinearBounds;
struct RTCBuildArguments;
typedef void (*RTCBoundsFunction)(const struct RTCBoundsFunctionArguments* args);
typedef void (*RTCIntersectFunctionN)(const struct RTCIntersectFunctionNArguments* args);
typedef void (*RTCOccludedFunctionN)(const struct RTCOccludedFunctionNArguments* args);
RTC_API ssize_t rtcGetDeviceProperty(RTCDevice device, enum RTCDeviceProperty prop);                         /* rtcore_device.h:67 */
RTC_API void rtcSetDeviceMemoryMonitorFunction(RTCDevice device, RTCMemoryMonitorFunction memoryMonitor, void* userPtr);
RTC_API RTCBuffer rtcNewBuffer(RTCDevice device, size_t byteSize);                                            /* rtcore_buffer.h:47-59 */
RTC_API RTCBuffer rtcNewSharedBuffer(RTCDevice device, void* ptr, size_t byteSize);
RTC_API void* rtcGetBufferData(RTCBuffer buffer);
RTC_API void rtcRetainBuffer(RTCBuffer buffer);
RTC_API void rtcReleaseBuffer(RTCBuffer buffer);
RTC_API void rtcSetGeometryBuffer(RTCGeometry geometry, enum RTCBufferType type, unsigned int slot, enum RTCFormat format, RTCBuffer buffer,
                                  size_t byteOffset, size_t byteStride, size_t itemCount);
RTC_API void rtcSetGeometryMask(RTCGeometry geometry, unsigned int mask);
RTC_API void rtcGetGeometryTransform(RTCGeometry geometry, float time, enum RTCFormat format, void* xfm);
RTC_API void rtcInterpolateN(const struct RTCInterpolateNArguments* args);
RTC_API void rtcSetGeometryIntersectFilterFunction(RTCGeometry geometry, RTCFilterFunctionN filter);          /* refused unless NULL */
RTC_API void rtcSetGeometryOccludedFilterFunction(RTCGeometry geometry, RTCFilterFunctionN filter);           /* refused unless NULL */
RTC_API void rtcSetGeometryUserPrimitiveCount(RTCGeometry geometry, unsigned int userPrimitiveCount);         /* user geometries build on the */
RTC_API void rtcSetGeometryBoundsFunction(RTCGeometry geometry, RTCBoundsFunction bounds, void* userPtr);     /* host; a scene holding one is */
RTC_API void rtcSetGeometryIntersectFunction(RTCGeometry geometry, RTCIntersectFunctionN intersect);          /* rejected by rtcCommitScene   */
RTC_API void rtcSetGeometryOccludedFunction(RTCGeometry geometry, RTCOccludedFunctionN occluded);
RTC_API void rtcFilterIntersection(const struct RTCIntersectFunctionNArguments* args, const struct RTCFilterFunctionNArguments* filterArgs);
RTC_API void rtcFilterOcclusion(const struct RTCOccludedFunctionNArguments* args, const struct RTCFilterFunctionNArguments* filterArgs);
RTC_API void rtcSetSceneProgressMonitorFunction(RTCScene scene, RTCProgressMonitorFunction progress, void* ptr);
RTC_API void rtcGetSceneLinearBounds(RTCScene scene, struct RTCLinearBounds* bounds_o);
RTC_API RTCBVH rtcNewBVH(RTCDevice device);                                                                   /* rtcore_builder.h:105-117 */
RTC_API void* rtcBuildBVH(const struct RTCBuildArguments* args);
RTC_API void* rtcThreadLocalAlloc(RTCThreadLocalAllocator allocator, size_t bytes, size_t align);
RTC_API void rtcRetainBVH(RTCBVH bvh);
RTC_API void rtcReleaseBVH(RTCBVH bvh);

/* ---- additions of this repo ----------------------------------------------------------------------------------- */
/* What the host Embree must export so that rtcCommitScene can hand the committed compressed accel to the plugin
 * (INTEGRATION.md shows the accessor a maintainer adds to kernels/common/rtcore.cpp; oracle/refhost.cpp supplies it
 * for the test build of the reference). */
struct RTCCompressedBVHInfo_b200 { uint64_t root; uint32_t leaf_kind, subdiv_level, comp_level, proj_col_major, n_accels; float bounds[6]; };
typedef int (*rtcGetCompressedBVH_b200_t)(RTCScene host_scene, struct RTCCompressedBVHInfo_b200* info);
struct RTCInstanceBVHInfo_b200 { uint64_t root; uint32_t n_instances; uint32_t reserved; float bounds[6]; };
typedef int (*rtcGetInstanceBVH_b200_t)(RTCScene host_scene, struct RTCInstanceBVHInfo_b200* info, unsigned cap, unsigned* geomIDs,
                                        float* world2local12, RTCScene* objects);
/* optional third accessor: root of the BVH4<Triangle4v> that AccelN visits before the compressed accel (0 if none) */
typedef int (*rtcGetTriangleBVH_b200_t)(RTCScene host_scene, uint64_t* root);
/* Pinned host memory for ray arrays.  Optional: the stream calls run at full speed on plain malloc'ed arrays (the host
 * pool packs the rays out of them); pinned arrays only matter for the "compact" / "copy" transfer strategies. */
RTC_API void* rtcAllocRays_b200(size_t bytes);
RTC_API void rtcFreeRays_b200(void* ptr);

#ifdef __cplusplus
}
#endif
#endif
