// rtcore_shim.cpp -- libembree3_b200.so: the embree3 API surface of the compressed-BVH path with the traversal on B200.
//
// A caller of the reference's API (rtcNewDevice ... rtcCommitScene, rtcIntersect1/4/8/16/1M, rtcOccluded1/4/8/16/1M,
// same struct layouts; include/embree3/rtcore_scene.h:43-133) links against this library instead of libembree3.so:
//   * scene CONSTRUCTION is forwarded to a host Embree (the reference: its CPU builder stays the builder), which is
//     dlopen'ed from the path given by the device config key `host_embree=` or by $ECB200_HOST_EMBREE;
//   * rtcCommitScene commits on the host, asks the host for the committed compressed BVH4
//     (rtcGetCompressedBVH_b200, the accessor INTEGRATION.md adds to kernels/common/rtcore.cpp) and converts it with
//     ecbSceneImportBVH4 into the device scene;
//   * every ray query runs in the CUDA kernels through the C ABI of include/ecb200.h.  There is NO fallback to the
//     host's intersectors: if the device path cannot serve a scene or there is no GPU, the call raises an RTCError
//     through the reference's sticky-error convention (rtcGetDeviceError / error callback, kernels/common/rtcore.h:40-69)
//     and leaves the rays untouched.
// rtcIntersect4/8/16/1M crash in the reference on this accel (null intersector pointers); here they
// are defined as "rtcIntersect1 applied to every active ray", which is Embree's contract for packets and streams.
#include <dlfcn.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ecb200.h"
#include "../../include/embree3_b200/rtcore_subset.h"

namespace {

struct HostApi {
  void* lib = nullptr;
#define HOST_FN(ret, name, args) ret(*name) args = nullptr;
  HOST_FN(RTCDevice, rtcNewDevice, (const char*))
  HOST_FN(void, rtcReleaseDevice, (RTCDevice))
  HOST_FN(RTCError, rtcGetDeviceError, (RTCDevice))
  HOST_FN(RTCGeometry, rtcNewGeometry, (RTCDevice, RTCGeometryType))
  HOST_FN(void, rtcRetainGeometry, (RTCGeometry))
  HOST_FN(void, rtcReleaseGeometry, (RTCGeometry))
  HOST_FN(void, rtcCommitGeometry, (RTCGeometry))
  HOST_FN(void, rtcEnableGeometry, (RTCGeometry))
  HOST_FN(void, rtcDisableGeometry, (RTCGeometry))
  HOST_FN(void, rtcSetGeometryTimeStepCount, (RTCGeometry, unsigned))
  HOST_FN(void, rtcSetGeometryVertexAttributeCount, (RTCGeometry, unsigned))
  HOST_FN(void, rtcSetGeometryBuildQuality, (RTCGeometry, RTCBuildQuality))
  HOST_FN(void, rtcSetSharedGeometryBuffer, (RTCGeometry, RTCBufferType, unsigned, RTCFormat, const void*, size_t, size_t, size_t))
  HOST_FN(void*, rtcSetNewGeometryBuffer, (RTCGeometry, RTCBufferType, unsigned, RTCFormat, size_t, size_t))
  HOST_FN(void*, rtcGetGeometryBufferData, (RTCGeometry, RTCBufferType, unsigned))
  HOST_FN(void, rtcUpdateGeometryBuffer, (RTCGeometry, RTCBufferType, unsigned))
  HOST_FN(void, rtcSetGeometryUserData, (RTCGeometry, void*))
  HOST_FN(void*, rtcGetGeometryUserData, (RTCGeometry))
  HOST_FN(void, rtcSetGeometryTessellationRate, (RTCGeometry, float))
  HOST_FN(void, rtcSetGeometryTopologyCount, (RTCGeometry, unsigned))
  HOST_FN(void, rtcSetGeometrySubdivisionMode, (RTCGeometry, unsigned, RTCSubdivisionMode))
  HOST_FN(void, rtcSetGeometryVertexAttributeTopology, (RTCGeometry, unsigned, unsigned))
  HOST_FN(void, rtcSetGeometryDisplacementFunction, (RTCGeometry, RTCDisplacementFunctionN))
  HOST_FN(void, rtcInterpolate, (const RTCInterpolateArguments*))
  HOST_FN(RTCScene, rtcNewScene, (RTCDevice))
  HOST_FN(void, rtcReleaseScene, (RTCScene))
  HOST_FN(unsigned, rtcAttachGeometry, (RTCScene, RTCGeometry))
  HOST_FN(void, rtcAttachGeometryByID, (RTCScene, RTCGeometry, unsigned))
  HOST_FN(void, rtcDetachGeometry, (RTCScene, unsigned))
  HOST_FN(RTCGeometry, rtcGetGeometry, (RTCScene, unsigned))
  HOST_FN(void, rtcSetSceneLevels, (RTCScene, unsigned, unsigned))
  HOST_FN(void, rtcCommitScene, (RTCScene))
  HOST_FN(void, rtcSetSceneBuildQuality, (RTCScene, RTCBuildQuality))
  HOST_FN(void, rtcSetSceneFlags, (RTCScene, RTCSceneFlags))
  HOST_FN(RTCSceneFlags, rtcGetSceneFlags, (RTCScene))
  HOST_FN(void, rtcGetSceneBounds, (RTCScene, RTCBounds*))
  HOST_FN(int, rtcGetCompressedBVH_b200, (RTCScene, RTCCompressedBVHInfo_b200*))
  HOST_FN(void, rtcSetGeometryInstancedScene, (RTCGeometry, RTCScene))
  HOST_FN(void, rtcSetGeometryTransform, (RTCGeometry, unsigned, RTCFormat, const void*))
  HOST_FN(int, rtcGetInstanceBVH_b200, (RTCScene, RTCInstanceBVHInfo_b200*, unsigned, unsigned*, float*, RTCScene*))
  HOST_FN(int, rtcGetTriangleBVH_b200, (RTCScene, uint64_t*))
  // object plumbing that never touches a ray (resolved lazily: a host without one of them only loses that call)
  HOST_FN(ssize_t, rtcGetDeviceProperty, (RTCDevice, RTCDeviceProperty))
  HOST_FN(void, rtcSetDeviceMemoryMonitorFunction, (RTCDevice, RTCMemoryMonitorFunction, void*))
  HOST_FN(RTCBuffer, rtcNewBuffer, (RTCDevice, size_t))
  HOST_FN(RTCBuffer, rtcNewSharedBuffer, (RTCDevice, void*, size_t))
  HOST_FN(void*, rtcGetBufferData, (RTCBuffer))
  HOST_FN(void, rtcRetainBuffer, (RTCBuffer))
  HOST_FN(void, rtcReleaseBuffer, (RTCBuffer))
  HOST_FN(void, rtcSetGeometryBuffer, (RTCGeometry, RTCBufferType, unsigned, RTCFormat, RTCBuffer, size_t, size_t, size_t))
  HOST_FN(void, rtcSetGeometryMask, (RTCGeometry, unsigned))
  HOST_FN(void, rtcGetGeometryTransform, (RTCGeometry, float, RTCFormat, void*))
  HOST_FN(void, rtcInterpolateN, (const RTCInterpolateNArguments*))
  HOST_FN(void, rtcSetGeometryIntersectFilterFunction, (RTCGeometry, RTCFilterFunctionN))
  HOST_FN(void, rtcSetGeometryOccludedFilterFunction, (RTCGeometry, RTCFilterFunctionN))
  HOST_FN(void, rtcSetGeometryUserPrimitiveCount, (RTCGeometry, unsigned))
  HOST_FN(void, rtcSetGeometryBoundsFunction, (RTCGeometry, RTCBoundsFunction, void*))
  HOST_FN(void, rtcSetGeometryIntersectFunction, (RTCGeometry, RTCIntersectFunctionN))
  HOST_FN(void, rtcSetGeometryOccludedFunction, (RTCGeometry, RTCOccludedFunctionN))
  HOST_FN(void, rtcSetSceneProgressMonitorFunction, (RTCScene, RTCProgressMonitorFunction, void*))
  HOST_FN(void, rtcGetSceneLinearBounds, (RTCScene, RTCLinearBounds*))
  HOST_FN(RTCBVH, rtcNewBVH, (RTCDevice))
  HOST_FN(void*, rtcBuildBVH, (const RTCBuildArguments*))
  HOST_FN(void*, rtcThreadLocalAlloc, (RTCThreadLocalAllocator, size_t, size_t))
  HOST_FN(void, rtcRetainBVH, (RTCBVH))
  HOST_FN(void, rtcReleaseBVH, (RTCBVH))
#undef HOST_FN
  const HostApi* registry = nullptr;  // the never-freed instance this copy was made from
};

struct ShimDevice {
  std::atomic<int> refs{1};
  HostApi host;
  RTCDevice hdev = nullptr;
  std::mutex mutex;
  RTCError error = RTC_ERROR_NONE;
  RTCErrorFunction callback = nullptr;
  void* callback_ptr = nullptr;
  int verbose = 0;
  int gpus = 1;  // device config key gpus=N

  void raise(RTCError code, const std::string& msg) {
    RTCErrorFunction cb;
    void* p;
    {
      std::lock_guard<std::mutex> lock(mutex);
      if (error == RTC_ERROR_NONE) error = code;  // first error sticks (kernels/common/device.cpp process_error)
      cb = callback;
      p = callback_ptr;
    }
    if (cb) cb(p, code, msg.c_str());
    if (verbose) fprintf(stderr, "embree3_b200: error %d: %s\n", (int)code, msg.c_str());
  }
};

struct ShimScene {
  std::atomic<int> refs{1};
  ShimDevice* dev = nullptr;
  RTCScene hscene = nullptr;
  ecbScene* gpu = nullptr;
  std::mutex mutex;
};

// handles of one geometry are the host's own handles; scenes and devices are wrapped
inline ShimDevice* D(RTCDevice d) { return reinterpret_cast<ShimDevice*>(d); }
inline ShimScene* S(RTCScene s) { return reinterpret_cast<ShimScene*>(s); }

// Geometry and buffer handles are the host library's own handles and may outlive the device object that created them
// (the reference's Geometry ref-counts its device, kernels/common/geometry.cpp:29).  Pass-through calls therefore never
// go through a ShimDevice: every host library is resolved once into a HostApi that is never unloaded or freed, a geometry
// remembers the library that created it, and anything else falls back to the first library this process loaded.
std::mutex g_host_mutex;
std::map<std::string, HostApi*> g_host_libs;         // path -> resolved entry points (kept for the life of the process)
std::map<RTCGeometry, const HostApi*> g_geom_host;   // geometry -> the library that created it
std::atomic<const HostApi*> g_first_host{nullptr};
std::atomic<ShimDevice*> g_any_device{nullptr};      // only used to REPORT errors of calls that carry no device

const HostApi* host_of(RTCGeometry g) {
  if (g) {
    std::lock_guard<std::mutex> lock(g_host_mutex);
    auto it = g_geom_host.find(g);
    if (it != g_geom_host.end()) return it->second;
  }
  const HostApi* h = g_first_host.load();
  return h;
}
// host scene handle -> wrapper, so that instances can find the device scene of the scene they refer to
std::mutex g_scene_map_mutex;
std::map<RTCScene, ShimScene*> g_scene_map;
thread_local RTCError g_thread_error = RTC_ERROR_NONE;  // errors without a device (rtcore.cpp:79-98)
const HostApi* host_or_error(RTCGeometry g) {
  const HostApi* h = host_of(g);
  if (!h) g_thread_error = RTC_ERROR_INVALID_OPERATION;  // no host library loaded yet: nothing to forward to
  return h;
}

RTCError map_error(ecbError e) {
  switch (e) {
    case ECB_SUCCESS: return RTC_ERROR_NONE;
    case ECB_ERROR_INVALID_ARGUMENT: return RTC_ERROR_INVALID_ARGUMENT;
    case ECB_ERROR_INVALID_OPERATION: return RTC_ERROR_INVALID_OPERATION;
    case ECB_ERROR_OUT_OF_MEMORY: return RTC_ERROR_OUT_OF_MEMORY;
    case ECB_ERROR_UNSUPPORTED: return RTC_ERROR_INVALID_OPERATION;
    default: return RTC_ERROR_UNKNOWN;  // CUDA errors and "no device": SURVEY 5 (RTC_ERROR_UNKNOWN, sticky)
  }
}

bool resolve_host(HostApi& h, const std::string& path, std::string& err);
bool load_host(HostApi& out, const std::string& path, std::string& err) {
  std::lock_guard<std::mutex> lock(g_host_mutex);
  auto it = g_host_libs.find(path);
  if (it == g_host_libs.end()) {
    HostApi* h = new HostApi();
    if (!resolve_host(*h, path, err)) {
      delete h;
      return false;
    }
    it = g_host_libs.emplace(path, h).first;
    const HostApi* expect = nullptr;
    g_first_host.compare_exchange_strong(expect, h);
  }
  out = *it->second;  // a copy of the entry points; the library itself stays loaded for the life of the process
  out.registry = it->second;
  return true;
}
bool resolve_host(HostApi& h, const std::string& path, std::string& err) {
  h.lib = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h.lib) {
    err = std::string("cannot load host embree '") + path + "': " + dlerror();
    return false;
  }
#define RESOLVE(name)                                                   \
  h.name = reinterpret_cast<decltype(h.name)>(dlsym(h.lib, #name));     \
  if (!h.name) { err = std::string("host embree lacks ") + #name; return false; }
  RESOLVE(rtcNewDevice) RESOLVE(rtcReleaseDevice) RESOLVE(rtcGetDeviceError) RESOLVE(rtcNewGeometry) RESOLVE(rtcRetainGeometry)
  RESOLVE(rtcReleaseGeometry) RESOLVE(rtcCommitGeometry) RESOLVE(rtcEnableGeometry) RESOLVE(rtcDisableGeometry)
  RESOLVE(rtcSetGeometryTimeStepCount) RESOLVE(rtcSetGeometryVertexAttributeCount) RESOLVE(rtcSetGeometryBuildQuality)
  RESOLVE(rtcSetSharedGeometryBuffer) RESOLVE(rtcSetNewGeometryBuffer) RESOLVE(rtcGetGeometryBufferData) RESOLVE(rtcUpdateGeometryBuffer)
  RESOLVE(rtcSetGeometryUserData) RESOLVE(rtcGetGeometryUserData) RESOLVE(rtcSetGeometryTessellationRate)
  RESOLVE(rtcSetGeometryTopologyCount) RESOLVE(rtcSetGeometrySubdivisionMode) RESOLVE(rtcSetGeometryVertexAttributeTopology)
  RESOLVE(rtcSetGeometryDisplacementFunction) RESOLVE(rtcInterpolate) RESOLVE(rtcNewScene) RESOLVE(rtcReleaseScene)
  RESOLVE(rtcAttachGeometry) RESOLVE(rtcAttachGeometryByID) RESOLVE(rtcDetachGeometry) RESOLVE(rtcGetGeometry) RESOLVE(rtcSetSceneLevels)
  RESOLVE(rtcSetGeometryInstancedScene) RESOLVE(rtcSetGeometryTransform)
  RESOLVE(rtcCommitScene) RESOLVE(rtcSetSceneBuildQuality) RESOLVE(rtcSetSceneFlags) RESOLVE(rtcGetSceneFlags) RESOLVE(rtcGetSceneBounds)
#undef RESOLVE
  h.rtcGetCompressedBVH_b200 = reinterpret_cast<decltype(h.rtcGetCompressedBVH_b200)>(dlsym(h.lib, "rtcGetCompressedBVH_b200"));
  h.rtcGetInstanceBVH_b200 = reinterpret_cast<decltype(h.rtcGetInstanceBVH_b200)>(dlsym(h.lib, "rtcGetInstanceBVH_b200"));  // optional
  h.rtcGetTriangleBVH_b200 = reinterpret_cast<decltype(h.rtcGetTriangleBVH_b200)>(dlsym(h.lib, "rtcGetTriangleBVH_b200"));  // optional
#define OPTIONAL(name) h.name = reinterpret_cast<decltype(h.name)>(dlsym(h.lib, #name));
  OPTIONAL(rtcGetDeviceProperty) OPTIONAL(rtcSetDeviceMemoryMonitorFunction) OPTIONAL(rtcNewBuffer) OPTIONAL(rtcNewSharedBuffer)
  OPTIONAL(rtcGetBufferData) OPTIONAL(rtcRetainBuffer) OPTIONAL(rtcReleaseBuffer) OPTIONAL(rtcSetGeometryBuffer) OPTIONAL(rtcSetGeometryMask)
  OPTIONAL(rtcGetGeometryTransform) OPTIONAL(rtcInterpolateN) OPTIONAL(rtcSetGeometryIntersectFilterFunction)
  OPTIONAL(rtcSetGeometryOccludedFilterFunction) OPTIONAL(rtcSetGeometryUserPrimitiveCount) OPTIONAL(rtcSetGeometryBoundsFunction)
  OPTIONAL(rtcSetGeometryIntersectFunction) OPTIONAL(rtcSetGeometryOccludedFunction) OPTIONAL(rtcSetSceneProgressMonitorFunction)
  OPTIONAL(rtcGetSceneLinearBounds) OPTIONAL(rtcNewBVH) OPTIONAL(rtcBuildBVH) OPTIONAL(rtcThreadLocalAlloc) OPTIONAL(rtcRetainBVH)
  OPTIONAL(rtcReleaseBVH)
#undef OPTIONAL
  if (!h.rtcGetCompressedBVH_b200) {
    err = "host embree does not export rtcGetCompressedBVH_b200 (apply the accessor of INTEGRATION.md)";
    return false;
  }
  return true;
}

// splits "key=value,key=value" ; returns the config for the host with our keys removed
std::string parse_config(const char* cfg, std::string& host_path, int& verbose, int& gpus) {
  std::string rest;
  std::string s = cfg ? cfg : "";
  size_t i = 0;
  while (i <= s.size()) {
    size_t j = s.find(',', i);
    if (j == std::string::npos) j = s.size();
    std::string item = s.substr(i, j - i);
    i = j + 1;
    if (item.empty()) continue;
    if (item.compare(0, 12, "host_embree=") == 0) { host_path = item.substr(12); continue; }
    // gpus=N: committed scenes are replicated on CUDA devices 0..N-1 and long stream calls fan out over them
    if (item.compare(0, 5, "gpus=") == 0) { gpus = atoi(item.c_str() + 5); continue; }
    if (item.compare(0, 8, "verbose=") == 0) verbose = atoi(item.c_str() + 8);
    if (!rest.empty()) rest += ",";
    rest += item;
  }
  return rest;
}

// makes a freshly imported scene resident: on the current device, or on devices 0..gpus-1 (device config gpus=N)
ecbError upload_scene(ecbScene* g, int gpus) {
  if (gpus <= 1) return ecbSceneUpload(g, -1);
  std::vector<int> devs;
  for (int i = 0; i < gpus; i++) devs.push_back(i);
  return ecbSceneUploadMulti(g, devs.data(), (uint32_t)devs.size());
}

uint32_t ctx_inst(const RTCIntersectContext* c) { return c ? c->instID[0] : RTC_INVALID_GEOMETRY_ID; }

bool ready(ShimScene* s, const char* what, const RTCIntersectContext* ctx) {
  if (!s) { g_thread_error = RTC_ERROR_INVALID_ARGUMENT; return false; }
  if (ctx && ctx->filter) {
    // a per-query filter callback (rtcore_common.h:214-219) is host code the reference calls for every hit; it cannot run
    // inside the CUDA traversal, and tracing without it would return different hits: refuse, rays untouched
    s->dev->raise(RTC_ERROR_INVALID_OPERATION, std::string(what) + ": RTCIntersectContext::filter is not supported on the device path");
    return false;
  }
  if (!s->gpu) {
    s->dev->raise(RTC_ERROR_INVALID_OPERATION, std::string(what) + ": scene not committed to the B200 path");
    return false;
  }
  return true;
}

void check(ShimScene* s, ecbError e, const char* what) {
  if (e != ECB_SUCCESS) s->dev->raise(map_error(e), std::string(what) + ": " + ecbErrorString(e) + " " + ecbLastErrorMessage());
}

// SoA packets of N rays <-> N AoS RTCRayHit records.  The SoA structs are 20 (12 for RTCRayN) arrays of N words.
template <int N>
void packet_intersect(const int* valid, ShimScene* s, RTCIntersectContext* ctx, void* packet, bool occluded) {
  const int fields = occluded ? 12 : 20;
  alignas(16) uint32_t aos[N][20];
  uint32_t* soa = static_cast<uint32_t*>(packet);
  const float one = 1.f, zero = 0.f;
  for (int l = 0; l < N; l++) {
    for (int f = 0; f < fields; f++) aos[l][f] = soa[f * N + l];
    if (valid[l] != -1) {  // inactive lane -> inactive stream ray
      memcpy(&aos[l][3], &one, 4);
      memcpy(&aos[l][8], &zero, 4);
    }
  }
  ecbError e;
  if (occluded) {
    alignas(16) uint32_t rays[N][12];
    for (int l = 0; l < N; l++) memcpy(rays[l], aos[l], 48);
    e = ecbOccluded1M(s->gpu, ctx_inst(ctx), rays, N, 48);
    for (int l = 0; l < N; l++) aos[l][8] = rays[l][8];
  } else {
    e = ecbIntersect1M(s->gpu, ctx_inst(ctx), aos, N, 80);
  }
  check(s, e, "packet query");
  if (e != ECB_SUCCESS) return;
  for (int l = 0; l < N; l++) {
    if (valid[l] != -1) continue;
    soa[8 * N + l] = aos[l][8];
    if (!occluded)
      for (int f = 12; f < 20; f++) soa[f * N + l] = aos[l][f];
  }
}

// The remaining stream layouts of the API (rtcore_scene.h:112-142) are gathered into one AoS batch, traced with the
// stream entry point and scattered back: pointer streams (1Mp), streams of N-wide SoA packets (NM) and pointer-SoA (Np).
// `word(i, f)` returns the address of field f (0..19, RTCRayHit word order) of ray i.
template <typename WordAt>
void stream_any_layout(ShimScene* s, RTCIntersectContext* ctx, size_t n, bool occluded, WordAt word, const char* what) {
  if (n == 0) return;
  const int fields = occluded ? 12 : 20;
  void* pinned = nullptr;
  const size_t bytes = n * (size_t)fields * 4;
  std::vector<uint32_t> pageable;
  uint32_t* aos = nullptr;
  if (bytes >= (1u << 20) && ecbHostAlloc(&pinned, bytes) == ECB_SUCCESS) aos = static_cast<uint32_t*>(pinned);
  else { pageable.resize(n * (size_t)fields); aos = pageable.data(); }
  for (size_t i = 0; i < n; i++)
    for (int f = 0; f < fields; f++) aos[i * fields + f] = *word(i, f);
  // gathered in chunks: the C ABI counts rays in 32 bits
  ecbError e = ECB_SUCCESS;
  for (size_t first = 0; first < n && e == ECB_SUCCESS; first += (1u << 30)) {
    const uint32_t m = (uint32_t)((n - first) < (1u << 30) ? (n - first) : (1u << 30));
    e = occluded ? ecbOccluded1M(s->gpu, ctx_inst(ctx), aos + first * fields, m, 48)
                 : ecbIntersect1M(s->gpu, ctx_inst(ctx), aos + first * fields, m, 80);
  }
  check(s, e, what);
  if (e == ECB_SUCCESS)
    for (size_t i = 0; i < n; i++) {
      *word(i, 8) = aos[i * fields + 8];  // tfar
      if (!occluded)
        for (int f = 12; f < 20; f++) *word(i, f) = aos[i * fields + f];
    }
  if (pinned) ecbHostFree(pinned);
}

}  // namespace

extern "C" {

// ---- device ----------------------------------------------------------------------------------------------------------
RTC_API RTCDevice rtcNewDevice(const char* config) {
  ShimDevice* d = new ShimDevice;
  std::string host_path = getenv("ECB200_HOST_EMBREE") ? getenv("ECB200_HOST_EMBREE") : "";
  const std::string host_cfg = parse_config(config, host_path, d->verbose, d->gpus);
  std::string err;
  if (host_path.empty()) err = "no host embree given (device config host_embree=<libembree3.so> or $ECB200_HOST_EMBREE)";
  if (!err.empty() || !load_host(d->host, host_path, err)) {
    g_thread_error = RTC_ERROR_INVALID_OPERATION;
    fprintf(stderr, "embree3_b200: %s\n", err.c_str());
    delete d;
    return nullptr;
  }
  d->hdev = d->host.rtcNewDevice(host_cfg.c_str());
  if (!d->hdev) {
    g_thread_error = RTC_ERROR_UNKNOWN;
    delete d;
    return nullptr;
  }
  g_any_device = d;
  if (d->verbose) fprintf(stderr, "embree3_b200: host embree %s, traversal on CUDA (libecb200)\n", host_path.c_str());
  return reinterpret_cast<RTCDevice>(d);
}
RTC_API void rtcRetainDevice(RTCDevice device) { if (device) D(device)->refs++; }
RTC_API void rtcReleaseDevice(RTCDevice device) {
  if (!device) return;
  ShimDevice* d = D(device);
  if (--d->refs == 0) {
    d->host.rtcReleaseDevice(d->hdev);
    ShimDevice* expect = d;
    g_any_device.compare_exchange_strong(expect, nullptr);
    delete d;  // the host library stays loaded: geometry handles may outlive the device object
  }
}
RTC_API RTCError rtcGetDeviceError(RTCDevice device) {
  if (!device) { RTCError e = g_thread_error; g_thread_error = RTC_ERROR_NONE; return e; }
  ShimDevice* d = D(device);
  RTCError host = d->host.rtcGetDeviceError(d->hdev);
  std::lock_guard<std::mutex> lock(d->mutex);
  RTCError e = d->error != RTC_ERROR_NONE ? d->error : host;
  d->error = RTC_ERROR_NONE;
  return e;
}
RTC_API void rtcSetDeviceErrorFunction(RTCDevice device, RTCErrorFunction error, void* userPtr) {
  if (!device) return;
  std::lock_guard<std::mutex> lock(D(device)->mutex);
  D(device)->callback = error;
  D(device)->callback_ptr = userPtr;
}

// ---- geometry: forwarded to the host builder --------------------------------------------------------------------------
RTC_API RTCGeometry rtcNewGeometry(RTCDevice device, RTCGeometryType type) {
  if (!device) { g_thread_error = RTC_ERROR_INVALID_ARGUMENT; return nullptr; }
  RTCGeometry g = D(device)->host.rtcNewGeometry(D(device)->hdev, type);
  if (g) {
    std::lock_guard<std::mutex> lock(g_host_mutex);
    g_geom_host[g] = D(device)->host.registry;  // a recycled handle value simply overwrites its stale entry
  }
  return g;
}
RTC_API void rtcRetainGeometry(RTCGeometry g) { if (const HostApi* h = host_or_error(g)) h->rtcRetainGeometry(g); }
RTC_API void rtcReleaseGeometry(RTCGeometry g) { if (const HostApi* h = host_or_error(g)) h->rtcReleaseGeometry(g); }
RTC_API void rtcCommitGeometry(RTCGeometry g) { if (const HostApi* h = host_or_error(g)) h->rtcCommitGeometry(g); }
RTC_API void rtcEnableGeometry(RTCGeometry g) { if (const HostApi* h = host_or_error(g)) h->rtcEnableGeometry(g); }
RTC_API void rtcDisableGeometry(RTCGeometry g) { if (const HostApi* h = host_or_error(g)) h->rtcDisableGeometry(g); }
RTC_API void rtcSetGeometryTimeStepCount(RTCGeometry g, unsigned n) { if (const HostApi* h = host_or_error(g)) h->rtcSetGeometryTimeStepCount(g, n); }
RTC_API void rtcSetGeometryVertexAttributeCount(RTCGeometry g, unsigned n) { if (const HostApi* h = host_or_error(g)) h->rtcSetGeometryVertexAttributeCount(g, n); }
RTC_API void rtcSetGeometryBuildQuality(RTCGeometry g, RTCBuildQuality q) { if (const HostApi* h = host_or_error(g)) h->rtcSetGeometryBuildQuality(g, q); }
RTC_API void rtcSetSharedGeometryBuffer(RTCGeometry g, RTCBufferType t, unsigned slot, RTCFormat f, const void* p, size_t off, size_t stride,
                                        size_t n) {
  if (const HostApi* h = host_or_error(g)) h->rtcSetSharedGeometryBuffer(g, t, slot, f, p, off, stride, n);
}
RTC_API void* rtcSetNewGeometryBuffer(RTCGeometry g, RTCBufferType t, unsigned slot, RTCFormat f, size_t stride, size_t n) {
  const HostApi* h = host_or_error(g);
  return h ? h->rtcSetNewGeometryBuffer(g, t, slot, f, stride, n) : nullptr;
}
RTC_API void* rtcGetGeometryBufferData(RTCGeometry g, RTCBufferType t, unsigned slot) { const HostApi* h = host_or_error(g); return h ? h->rtcGetGeometryBufferData(g, t, slot) : nullptr; }
RTC_API void rtcUpdateGeometryBuffer(RTCGeometry g, RTCBufferType t, unsigned slot) { if (const HostApi* h = host_or_error(g)) h->rtcUpdateGeometryBuffer(g, t, slot); }
RTC_API void rtcSetGeometryUserData(RTCGeometry g, void* p) { if (const HostApi* h = host_or_error(g)) h->rtcSetGeometryUserData(g, p); }
RTC_API void* rtcGetGeometryUserData(RTCGeometry g) { const HostApi* h = host_or_error(g); return h ? h->rtcGetGeometryUserData(g) : nullptr; }
RTC_API void rtcSetGeometryTessellationRate(RTCGeometry g, float r) { if (const HostApi* h = host_or_error(g)) h->rtcSetGeometryTessellationRate(g, r); }
RTC_API void rtcSetGeometryTopologyCount(RTCGeometry g, unsigned n) { if (const HostApi* h = host_or_error(g)) h->rtcSetGeometryTopologyCount(g, n); }
RTC_API void rtcSetGeometrySubdivisionMode(RTCGeometry g, unsigned t, RTCSubdivisionMode m) { if (const HostApi* h = host_or_error(g)) h->rtcSetGeometrySubdivisionMode(g, t, m); }
RTC_API void rtcSetGeometryVertexAttributeTopology(RTCGeometry g, unsigned a, unsigned t) { if (const HostApi* h = host_or_error(g)) h->rtcSetGeometryVertexAttributeTopology(g, a, t); }
RTC_API void rtcSetGeometryDisplacementFunction(RTCGeometry g, RTCDisplacementFunctionN f) { if (const HostApi* h = host_or_error(g)) h->rtcSetGeometryDisplacementFunction(g, f); }
RTC_API void rtcInterpolate(const RTCInterpolateArguments* args) { if (const HostApi* h = host_or_error(args ? args->geometry : nullptr)) h->rtcInterpolate(args); }
RTC_API void rtcSetGeometryInstancedScene(RTCGeometry g, RTCScene scene) { if (const HostApi* h = host_or_error(g)) h->rtcSetGeometryInstancedScene(g, scene ? reinterpret_cast<ShimScene*>(scene)->hscene : nullptr); }
RTC_API void rtcSetGeometryTransform(RTCGeometry g, unsigned t, RTCFormat f, const void* xfm) { if (const HostApi* h = host_or_error(g)) h->rtcSetGeometryTransform(g, t, f, xfm); }


// ---- the rest of the embree3 surface: plumbing that never touches a ray, forwarded to the host -------------------------
// (a call the host library does not export, or made before any device exists, raises RTC_ERROR_INVALID_OPERATION)
static ShimDevice* any_device(const char* what) {
  ShimDevice* d = g_any_device.load();
  if (!d) g_thread_error = RTC_ERROR_INVALID_OPERATION;
  (void)what;
  return d;
}
#define FORWARD(name, what, fail, ...)                                                                    \
  const HostApi* h__ = host_or_error(nullptr);                                                            \
  if (!h__) return fail;                                                                                  \
  if (!h__->name) {                                                                                       \
    if (ShimDevice* d__ = any_device(what))                                                               \
      d__->raise(RTC_ERROR_INVALID_OPERATION, std::string(what) + ": the host embree does not export it"); \
    return fail;                                                                                          \
  }                                                                                                       \
  return h__->name(__VA_ARGS__)

RTC_API ssize_t rtcGetDeviceProperty(RTCDevice device, RTCDeviceProperty prop) {
  if (!device) { g_thread_error = RTC_ERROR_INVALID_ARGUMENT; return 0; }
  switch (prop) {
    // the device path serves every packet width and the stream calls (the reference's compressed accel serves none of them)
    case RTC_DEVICE_PROPERTY_NATIVE_RAY4_SUPPORTED: case RTC_DEVICE_PROPERTY_NATIVE_RAY8_SUPPORTED:
    case RTC_DEVICE_PROPERTY_NATIVE_RAY16_SUPPORTED: case RTC_DEVICE_PROPERTY_RAY_STREAM_SUPPORTED: return 1;
    // host callbacks cannot run inside the CUDA traversal; ray masks are not evaluated (EMBREE_RAY_MASK is off in the reference build too)
    case RTC_DEVICE_PROPERTY_FILTER_FUNCTION_SUPPORTED: case RTC_DEVICE_PROPERTY_USER_GEOMETRY_SUPPORTED:
    case RTC_DEVICE_PROPERTY_RAY_MASK_SUPPORTED: case RTC_DEVICE_PROPERTY_QUAD_GEOMETRY_SUPPORTED:
    case RTC_DEVICE_PROPERTY_CURVE_GEOMETRY_SUPPORTED: return 0;
    default: break;
  }
  ShimDevice* d = D(device);
  if (!d->host.rtcGetDeviceProperty) { d->raise(RTC_ERROR_INVALID_OPERATION, "rtcGetDeviceProperty: the host embree does not export it"); return 0; }
  return d->host.rtcGetDeviceProperty(d->hdev, prop);
}
RTC_API void rtcSetDeviceMemoryMonitorFunction(RTCDevice device, RTCMemoryMonitorFunction fn, void* ptr) {
  if (!device) { g_thread_error = RTC_ERROR_INVALID_ARGUMENT; return; }
  if (D(device)->host.rtcSetDeviceMemoryMonitorFunction) D(device)->host.rtcSetDeviceMemoryMonitorFunction(D(device)->hdev, fn, ptr);
}
RTC_API RTCBuffer rtcNewBuffer(RTCDevice device, size_t bytes) {
  if (!device || !D(device)->host.rtcNewBuffer) { g_thread_error = RTC_ERROR_INVALID_ARGUMENT; return nullptr; }
  return D(device)->host.rtcNewBuffer(D(device)->hdev, bytes);
}
RTC_API RTCBuffer rtcNewSharedBuffer(RTCDevice device, void* ptr, size_t bytes) {
  if (!device || !D(device)->host.rtcNewSharedBuffer) { g_thread_error = RTC_ERROR_INVALID_ARGUMENT; return nullptr; }
  return D(device)->host.rtcNewSharedBuffer(D(device)->hdev, ptr, bytes);
}
RTC_API void* rtcGetBufferData(RTCBuffer b) { FORWARD(rtcGetBufferData, "rtcGetBufferData", nullptr, b); }
RTC_API void rtcRetainBuffer(RTCBuffer b) { FORWARD(rtcRetainBuffer, "rtcRetainBuffer", , b); }
RTC_API void rtcReleaseBuffer(RTCBuffer b) { FORWARD(rtcReleaseBuffer, "rtcReleaseBuffer", , b); }
RTC_API void rtcSetGeometryBuffer(RTCGeometry g, RTCBufferType t, unsigned slot, RTCFormat f, RTCBuffer b, size_t off, size_t stride, size_t n) {
  FORWARD(rtcSetGeometryBuffer, "rtcSetGeometryBuffer", , g, t, slot, f, b, off, stride, n);
}
RTC_API void rtcSetGeometryMask(RTCGeometry g, unsigned mask) { FORWARD(rtcSetGeometryMask, "rtcSetGeometryMask", , g, mask); }
RTC_API void rtcGetGeometryTransform(RTCGeometry g, float time, RTCFormat f, void* xfm) {
  FORWARD(rtcGetGeometryTransform, "rtcGetGeometryTransform", , g, time, f, xfm);
}
RTC_API void rtcInterpolateN(const RTCInterpolateNArguments* args) { FORWARD(rtcInterpolateN, "rtcInterpolateN", , args); }
RTC_API void rtcSetGeometryIntersectFilterFunction(RTCGeometry g, RTCFilterFunctionN filter) {
  if (filter) {
    if (ShimDevice* d = any_device("rtcSetGeometryIntersectFilterFunction"))
      d->raise(RTC_ERROR_INVALID_OPERATION, "rtcSetGeometryIntersectFilterFunction: host callbacks cannot run inside the CUDA traversal");
    return;
  }
  FORWARD(rtcSetGeometryIntersectFilterFunction, "rtcSetGeometryIntersectFilterFunction", , g, filter);
}
RTC_API void rtcSetGeometryOccludedFilterFunction(RTCGeometry g, RTCFilterFunctionN filter) {
  if (filter) {
    if (ShimDevice* d = any_device("rtcSetGeometryOccludedFilterFunction"))
      d->raise(RTC_ERROR_INVALID_OPERATION, "rtcSetGeometryOccludedFilterFunction: host callbacks cannot run inside the CUDA traversal");
    return;
  }
  FORWARD(rtcSetGeometryOccludedFilterFunction, "rtcSetGeometryOccludedFilterFunction", , g, filter);
}
// user geometries are built by the host like any other geometry; rtcCommitScene refuses a scene that holds one
RTC_API void rtcSetGeometryUserPrimitiveCount(RTCGeometry g, unsigned n) { FORWARD(rtcSetGeometryUserPrimitiveCount, "rtcSetGeometryUserPrimitiveCount", , g, n); }
RTC_API void rtcSetGeometryBoundsFunction(RTCGeometry g, RTCBoundsFunction fn, void* ptr) { FORWARD(rtcSetGeometryBoundsFunction, "rtcSetGeometryBoundsFunction", , g, fn, ptr); }
RTC_API void rtcSetGeometryIntersectFunction(RTCGeometry g, RTCIntersectFunctionN fn) { FORWARD(rtcSetGeometryIntersectFunction, "rtcSetGeometryIntersectFunction", , g, fn); }
RTC_API void rtcSetGeometryOccludedFunction(RTCGeometry g, RTCOccludedFunctionN fn) { FORWARD(rtcSetGeometryOccludedFunction, "rtcSetGeometryOccludedFunction", , g, fn); }
// only callable from inside a user-geometry callback, which never runs on the device path
RTC_API void rtcFilterIntersection(const RTCIntersectFunctionNArguments*, const RTCFilterFunctionNArguments*) {
  if (ShimDevice* d = any_device("rtcFilterIntersection")) d->raise(RTC_ERROR_INVALID_OPERATION, "rtcFilterIntersection: no user-geometry callbacks on the device path");
}
RTC_API void rtcFilterOcclusion(const RTCOccludedFunctionNArguments*, const RTCFilterFunctionNArguments*) {
  if (ShimDevice* d = any_device("rtcFilterOcclusion")) d->raise(RTC_ERROR_INVALID_OPERATION, "rtcFilterOcclusion: no user-geometry callbacks on the device path");
}
RTC_API RTCBVH rtcNewBVH(RTCDevice device) {
  if (!device || !D(device)->host.rtcNewBVH) { g_thread_error = RTC_ERROR_INVALID_ARGUMENT; return nullptr; }
  return D(device)->host.rtcNewBVH(D(device)->hdev);
}
RTC_API void* rtcBuildBVH(const RTCBuildArguments* args) { FORWARD(rtcBuildBVH, "rtcBuildBVH", nullptr, args); }
RTC_API void* rtcThreadLocalAlloc(RTCThreadLocalAllocator a, size_t bytes, size_t align) { FORWARD(rtcThreadLocalAlloc, "rtcThreadLocalAlloc", nullptr, a, bytes, align); }
RTC_API void rtcRetainBVH(RTCBVH b) { FORWARD(rtcRetainBVH, "rtcRetainBVH", , b); }
RTC_API void rtcReleaseBVH(RTCBVH b) { FORWARD(rtcReleaseBVH, "rtcReleaseBVH", , b); }
#undef FORWARD

// ---- scene ------------------------------------------------------------------------------------------------------------
RTC_API RTCScene rtcNewScene(RTCDevice device) {
  if (!device) { g_thread_error = RTC_ERROR_INVALID_ARGUMENT; return nullptr; }
  ShimScene* s = new ShimScene;
  s->dev = D(device);
  s->dev->refs++;
  s->hscene = s->dev->host.rtcNewScene(s->dev->hdev);
  {
    std::lock_guard<std::mutex> lock(g_scene_map_mutex);
    g_scene_map[s->hscene] = s;
  }
  return reinterpret_cast<RTCScene>(s);
}
RTC_API void rtcRetainScene(RTCScene scene) { if (scene) S(scene)->refs++; }
RTC_API void rtcReleaseScene(RTCScene scene) {
  if (!scene) return;
  ShimScene* s = S(scene);
  if (--s->refs == 0) {
    if (s->gpu) ecbSceneRelease(s->gpu);
    {
      std::lock_guard<std::mutex> lock(g_scene_map_mutex);
      g_scene_map.erase(s->hscene);
    }
    s->dev->host.rtcReleaseScene(s->hscene);
    rtcReleaseDevice(reinterpret_cast<RTCDevice>(s->dev));
    delete s;
  }
}
RTC_API unsigned rtcAttachGeometry(RTCScene scene, RTCGeometry g) { return S(scene)->dev->host.rtcAttachGeometry(S(scene)->hscene, g); }
RTC_API void rtcAttachGeometryByID(RTCScene scene, RTCGeometry g, unsigned id) { S(scene)->dev->host.rtcAttachGeometryByID(S(scene)->hscene, g, id); }
RTC_API void rtcDetachGeometry(RTCScene scene, unsigned id) { S(scene)->dev->host.rtcDetachGeometry(S(scene)->hscene, id); }
RTC_API RTCGeometry rtcGetGeometry(RTCScene scene, unsigned id) { return S(scene)->dev->host.rtcGetGeometry(S(scene)->hscene, id); }
RTC_API void rtcSetSceneLevels(RTCScene scene, unsigned S_, unsigned C_) { S(scene)->dev->host.rtcSetSceneLevels(S(scene)->hscene, S_, C_); }
RTC_API void rtcSetSceneBuildQuality(RTCScene scene, RTCBuildQuality q) { S(scene)->dev->host.rtcSetSceneBuildQuality(S(scene)->hscene, q); }
RTC_API void rtcSetSceneFlags(RTCScene scene, RTCSceneFlags f) { S(scene)->dev->host.rtcSetSceneFlags(S(scene)->hscene, f); }
RTC_API RTCSceneFlags rtcGetSceneFlags(RTCScene scene) { return S(scene)->dev->host.rtcGetSceneFlags(S(scene)->hscene); }
RTC_API void rtcGetSceneBounds(RTCScene scene, RTCBounds* b) { S(scene)->dev->host.rtcGetSceneBounds(S(scene)->hscene, b); }
RTC_API void rtcGetSceneLinearBounds(RTCScene scene, RTCLinearBounds* b) {
  if (!scene) { g_thread_error = RTC_ERROR_INVALID_ARGUMENT; return; }
  if (S(scene)->dev->host.rtcGetSceneLinearBounds) S(scene)->dev->host.rtcGetSceneLinearBounds(S(scene)->hscene, b);
  else S(scene)->dev->raise(RTC_ERROR_INVALID_OPERATION, "rtcGetSceneLinearBounds: the host embree does not export it");
}
RTC_API void rtcSetSceneProgressMonitorFunction(RTCScene scene, RTCProgressMonitorFunction fn, void* ptr) {
  if (!scene) { g_thread_error = RTC_ERROR_INVALID_ARGUMENT; return; }
  if (S(scene)->dev->host.rtcSetSceneProgressMonitorFunction) S(scene)->dev->host.rtcSetSceneProgressMonitorFunction(S(scene)->hscene, fn, ptr);
}

// Commit: host build (kernels/common/scene.cpp:632-697), then hand-off of the committed compressed BVH4 to the device path.
RTC_API void rtcCommitScene(RTCScene scene) {
  if (!scene) { g_thread_error = RTC_ERROR_INVALID_ARGUMENT; return; }
  ShimScene* s = S(scene);
  std::lock_guard<std::mutex> lock(s->mutex);
  s->dev->host.rtcCommitScene(s->hscene);
  if (s->gpu) { ecbSceneRelease(s->gpu); s->gpu = nullptr; }
  RTCCompressedBVHInfo_b200 info;
  const int rc = s->dev->host.rtcGetCompressedBVH_b200(s->hscene, &info);
  if (rc == -2 && s->dev->host.rtcGetInstanceBVH_b200) {
    // not a compressed accel: a scene of instances of (already committed) compressed scenes?  
    RTCInstanceBVHInfo_b200 ii;
    const unsigned cap = 1u << 16;
    std::vector<unsigned> ids(cap);
    std::vector<float> xf(12 * (size_t)cap);
    std::vector<RTCScene> hobj(cap);
    if (s->dev->host.rtcGetInstanceBVH_b200(s->hscene, &ii, cap, ids.data(), xf.data(), hobj.data()) == 0 && ii.n_instances <= cap) {
      std::vector<ecbScene*> objs(ii.n_instances);
      {
        std::lock_guard<std::mutex> map_lock(g_scene_map_mutex);
        for (unsigned i = 0; i < ii.n_instances; i++) {
          auto it = g_scene_map.find(hobj[i]);
          objs[i] = (it != g_scene_map.end()) ? it->second->gpu : nullptr;
        }
      }
      for (unsigned i = 0; i < ii.n_instances; i++)
        if (!objs[i]) {
          s->dev->raise(RTC_ERROR_INVALID_OPERATION, "instanced scene must be committed (on the B200 path) before the scene that instances it");
          return;
        }
      ecbInstanceDesc d;
      memset(&d, 0, sizeof(d));
      d.root = ii.root;
      d.n_instances = ii.n_instances;
      d.geomIDs = ids.data();
      d.world2local = xf.data();
      d.objects = objs.data();
      memcpy(d.bounds, ii.bounds, sizeof(d.bounds));
      ecbScene* g = nullptr;
      ecbError e = ecbSceneImportInstances(&d, &g);
      if (e == ECB_SUCCESS) e = upload_scene(g, s->dev->gpus);
      if (e != ECB_SUCCESS) {
        if (g) ecbSceneRelease(g);
        s->dev->raise(map_error(e), std::string("rtcCommitScene (instances): ") + ecbErrorString(e) + " " + ecbLastErrorMessage());
        return;
      }
      s->gpu = g;
      return;
    }
  }
  // mixed scenes: a triangle accel (BVH4<Triangle4v>, tri_accel=bvh4.triangle4v) next to the compressed one
  uint64_t tri_root = 0;
  if (rc == 0 && info.n_accels == 2 && s->dev->host.rtcGetTriangleBVH_b200 &&
      s->dev->host.rtcGetTriangleBVH_b200(s->hscene, &tri_root) != 0)
    tri_root = 0;
  if (rc != 0 || !(info.n_accels == 1 || (info.n_accels == 2 && tri_root))) {
    s->dev->raise(RTC_ERROR_INVALID_OPERATION,
                  rc == -1 ? "device config must select subdiv_accel=bvh4.compressed.{grid,leaf,box}"
                           : "scene is not a compressed subdivision accel, optionally with triangle meshes in a BVH4<Triangle4v> "
                             "(tri_accel=bvh4.triangle4v); other geometry types are outside the B200 hot path");
    return;
  }
  ecbImportDesc d;
  memset(&d, 0, sizeof(d));
  d.base = nullptr;  // live host memory of this process
  d.root = info.root;
  d.leaf_kind = info.leaf_kind;
  d.proj_col_major = info.proj_col_major;
  d.subdiv_level = info.subdiv_level;
  d.comp_level = info.comp_level;
  memcpy(d.bounds, info.bounds, sizeof(d.bounds));
  ecbScene* g = nullptr;
  ecbError e = ecbSceneImportBVH4(&d, &g);
  if (e == ECB_SUCCESS && tri_root) {
    ecbTriangleDesc t;
    memset(&t, 0, sizeof(t));
    t.root = tri_root;
    e = ecbSceneAttachTriangles4v(g, &t);
  }
  if (e == ECB_SUCCESS) e = upload_scene(g, s->dev->gpus);
  if (e != ECB_SUCCESS) {
    if (g) ecbSceneRelease(g);
    s->dev->raise(map_error(e), std::string("rtcCommitScene: ") + ecbErrorString(e) + " " + ecbLastErrorMessage());
    return;
  }
  s->gpu = g;
}
RTC_API void rtcJoinCommitScene(RTCScene scene) { rtcCommitScene(scene); }

// ---- ray queries: all on the device -----------------------------------------------------------------------------------
RTC_API void rtcIntersect1(RTCScene scene, RTCIntersectContext* ctx, RTCRayHit* rayhit) {
  ShimScene* s = S(scene);
  if (ready(s, "rtcIntersect1", ctx)) check(s, ecbIntersect1(s->gpu, ctx_inst(ctx), rayhit), "rtcIntersect1");
}
RTC_API void rtcIntersect1M(RTCScene scene, RTCIntersectContext* ctx, RTCRayHit* rayhit, unsigned M, size_t byteStride) {
  ShimScene* s = S(scene);
  if (ready(s, "rtcIntersect1M", ctx)) check(s, ecbIntersect1M(s->gpu, ctx_inst(ctx), rayhit, M, byteStride), "rtcIntersect1M");
}
RTC_API void rtcIntersect4(const int* valid, RTCScene scene, RTCIntersectContext* ctx, RTCRayHit4* rh) {
  if (ready(S(scene), "rtcIntersect4", ctx)) packet_intersect<4>(valid, S(scene), ctx, rh, false);
}
RTC_API void rtcIntersect8(const int* valid, RTCScene scene, RTCIntersectContext* ctx, RTCRayHit8* rh) {
  if (ready(S(scene), "rtcIntersect8", ctx)) packet_intersect<8>(valid, S(scene), ctx, rh, false);
}
RTC_API void rtcIntersect16(const int* valid, RTCScene scene, RTCIntersectContext* ctx, RTCRayHit16* rh) {
  if (ready(S(scene), "rtcIntersect16", ctx)) packet_intersect<16>(valid, S(scene), ctx, rh, false);
}
RTC_API void rtcOccluded1(RTCScene scene, RTCIntersectContext* ctx, RTCRay* ray) {
  ShimScene* s = S(scene);
  if (ready(s, "rtcOccluded1", ctx)) check(s, ecbOccluded1(s->gpu, ctx_inst(ctx), ray), "rtcOccluded1");
}
RTC_API void rtcOccluded1M(RTCScene scene, RTCIntersectContext* ctx, RTCRay* ray, unsigned M, size_t byteStride) {
  ShimScene* s = S(scene);
  if (ready(s, "rtcOccluded1M", ctx)) check(s, ecbOccluded1M(s->gpu, ctx_inst(ctx), ray, M, byteStride), "rtcOccluded1M");
}
RTC_API void rtcOccluded4(const int* valid, RTCScene scene, RTCIntersectContext* ctx, RTCRay4* r) {
  if (ready(S(scene), "rtcOccluded4", ctx)) packet_intersect<4>(valid, S(scene), ctx, r, true);
}
RTC_API void rtcOccluded8(const int* valid, RTCScene scene, RTCIntersectContext* ctx, RTCRay8* r) {
  if (ready(S(scene), "rtcOccluded8", ctx)) packet_intersect<8>(valid, S(scene), ctx, r, true);
}
RTC_API void rtcOccluded16(const int* valid, RTCScene scene, RTCIntersectContext* ctx, RTCRay16* r) {
  if (ready(S(scene), "rtcOccluded16", ctx)) packet_intersect<16>(valid, S(scene), ctx, r, true);
}

RTC_API void rtcIntersect1Mp(RTCScene scene, RTCIntersectContext* ctx, RTCRayHit** rayhit, unsigned M) {
  ShimScene* s = S(scene);
  if (!ready(s, "rtcIntersect1Mp", ctx)) return;
  stream_any_layout(s, ctx, M, false, [&](size_t i, int f) { return reinterpret_cast<uint32_t*>(rayhit[i]) + f; }, "rtcIntersect1Mp");
}
RTC_API void rtcOccluded1Mp(RTCScene scene, RTCIntersectContext* ctx, RTCRay** ray, unsigned M) {
  ShimScene* s = S(scene);
  if (!ready(s, "rtcOccluded1Mp", ctx)) return;
  stream_any_layout(s, ctx, M, true, [&](size_t i, int f) { return reinterpret_cast<uint32_t*>(ray[i]) + f; }, "rtcOccluded1Mp");
}
// M packets of N rays in SoA layout, byteStride between packets (rtcore_scene.h:115,139)
RTC_API void rtcIntersectNM(RTCScene scene, RTCIntersectContext* ctx, RTCRayHitN* rayhit, unsigned N, unsigned M, size_t byteStride) {
  ShimScene* s = S(scene);
  if (!ready(s, "rtcIntersectNM", ctx)) return;
  uint8_t* base = reinterpret_cast<uint8_t*>(rayhit);
  stream_any_layout(s, ctx, (size_t)N * M, false,
                    [&](size_t i, int f) { return reinterpret_cast<uint32_t*>(base + (i / N) * byteStride) + (size_t)f * N + (i % N); }, "rtcIntersectNM");
}
RTC_API void rtcOccludedNM(RTCScene scene, RTCIntersectContext* ctx, RTCRayN* ray, unsigned N, unsigned M, size_t byteStride) {
  ShimScene* s = S(scene);
  if (!ready(s, "rtcOccludedNM", ctx)) return;
  uint8_t* base = reinterpret_cast<uint8_t*>(ray);
  stream_any_layout(s, ctx, (size_t)N * M, true,
                    [&](size_t i, int f) { return reinterpret_cast<uint32_t*>(base + (i / N) * byteStride) + (size_t)f * N + (i % N); }, "rtcOccludedNM");
}
RTC_API void rtcIntersectNp(RTCScene scene, RTCIntersectContext* ctx, const RTCRayHitNp* rayhit, unsigned N) {
  ShimScene* s = S(scene);
  if (!ready(s, "rtcIntersectNp", ctx)) return;
  // RTCRayHitNp and ecbRayHitNp have the same layout (rtcore_ray.h:189-228): the arrays go to the device as they are
  static_assert(sizeof(RTCRayHitNp) == sizeof(ecbRayHitNp), "RTCRayHitNp layout");
  check(s, ecbIntersectNp(s->gpu, ctx_inst(ctx), reinterpret_cast<const ecbRayHitNp*>(rayhit), N), "rtcIntersectNp");
}
RTC_API void rtcOccludedNp(RTCScene scene, RTCIntersectContext* ctx, const RTCRayNp* ray, unsigned N) {
  ShimScene* s = S(scene);
  if (!ready(s, "rtcOccludedNp", ctx)) return;
  static_assert(sizeof(RTCRayNp) == sizeof(ecbRayNp), "RTCRayNp layout");
  check(s, ecbOccludedNp(s->gpu, ctx_inst(ctx), reinterpret_cast<const ecbRayNp*>(ray), N), "rtcOccludedNp");
}

RTC_API void* rtcAllocRays_b200(size_t bytes) {
  void* p = nullptr;
  return ecbHostAlloc(&p, bytes) == ECB_SUCCESS ? p : nullptr;
}
RTC_API void rtcFreeRays_b200(void* ptr) { ecbHostFree(ptr); }

}  // extern "C"
