// ecb_api.cpp -- C-ABI host layer of libecb200.so (include/ecb200.h): scene residency, host-buffer staging
// pipeline, device-buffer launches, serialisation.  No traversal arithmetic lives here and there is no CPU
// fallback: without a usable CUDA device every traversal entry point fails with ECB_ERROR_NO_DEVICE.
#include <cuda_runtime.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <new>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "ecb_kernels.h"
#include "ecb_scene.h"

namespace {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

constexpr size_t kRayHitBytes = 80, kRayBytes = 48;
constexpr int kStages = 8;
constexpr int kSmallStages = 2;              // stages of a small staging context (calls of up to kInlineRaysMax rays)
constexpr size_t kPackedBytes = 32;         // packed input record of the host-buffer path: {org.xyz,tnear},{dir.xyz,tfar}
constexpr int kPackPartsDefault = 8;        // pack jobs per chunk (ECB200_PACK_PARTS overrides for experiments)
constexpr int kDirectPercentDefault = 0;     // packed path, pinned caller memory: share of chunks uploaded as full records by DMA instead of packed (ECB200_DIRECT_PERCENT)
constexpr int kPackAheadDefault = 3;        // chunks packed ahead of the one being issued to the GPU (ECB200_PACK_AHEAD)
constexpr size_t kChunkRaysMax = 1u << 17;  // capacity of a staging buffer: 128 Ki rays
constexpr int kScatterParts = 4;            // scatter jobs per chunk
constexpr uint32_t kQueueRing = 256;
constexpr uint64_t kMultiDeviceRaysMin = 1u << 18;  // host-buffer calls of at least this many rays fan out over all devices the scene is resident on
constexpr uint64_t kInlineRaysMax = 16384;  // calls up to this size pack and scatter on the caller's thread (measured: 8 Ki rays 59 us inline, 101 us through the pool)

ecbError cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorNoKernelImageForDevice ||
      e == cudaErrorInvalidDevice)
    return ECB_ERROR_NO_DEVICE;
  if (e == cudaErrorMemoryAllocation) return ECB_ERROR_OUT_OF_MEMORY;
  return ECB_ERROR_CUDA;
}
#define CU(call)                                    \
  do {                                              \
    cudaError_t e__ = (call);                       \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

}  // namespace

void ecb_set_error(const std::string& msg) { g_last_error = msg; }

// Host-buffer pipeline context: `n_stages` chunks in flight, each H2D -> kernel(s) -> D2H on its own stream over fixed
// device buffers of `cap` rays.  Contexts belong to a Residency (one device copy of a scene) and are handed out to one
// call at a time: concurrent callers get their own context, so rtcIntersect1 / rtcIntersect1M stay re-entrant like the
// reference's (kernels/common/rtcore.cpp:287-304: no lock on the query path).  Two sizes: small (calls of up to
// kInlineRaysMax rays: packets, rtcIntersect1, short streams) and large (the 8-stage pipeline of long streams).
struct Staging {
  int device = -1;
  int n_stages = 0;
  size_t cap = 0;  // rays per stage
  cudaStream_t stream[kStages] = {};
  cudaEvent_t done[kStages] = {};
  void* d_buf[kStages] = {};
  void* h_pin[kStages] = {};  // pinned bounce buffers for pageable caller memory
  void* d_cmp[kStages] = {};  // compact results of a chunk (device) ...
  void* h_cmp[kStages] = {};  // ... and their pinned landing zone, scattered into the caller's array by the host pool
  void* h_pk[kStages] = {};   // packed input records of a chunk (pinned), written by the host pool, read by the H2D copy
  int pack_pending[kStages] = {};
  uint64_t gpu_issued = 0;      // packed path: chunks whose device work and `done` event have been issued ...
  uint64_t scat_submitted = 0;  // ... and chunks whose scatter jobs the completion thread has handed to the pool
  bool stop = false;
  // scatter jobs of the chunk that currently owns a stage
  std::mutex mutex;
  std::condition_variable cv;
  int pending[kStages] = {};
  cudaError_t worker_error = cudaSuccess;
  ~Staging() {
    if (device >= 0) cudaSetDevice(device);
    for (int i = 0; i < kStages; i++) {
      if (d_buf[i]) cudaFree(d_buf[i]);
      if (h_pin[i]) cudaFreeHost(h_pin[i]);
      if (d_cmp[i]) cudaFree(d_cmp[i]);
      if (h_cmp[i]) cudaFreeHost(h_cmp[i]);
      if (h_pk[i]) cudaFreeHost(h_pk[i]);
      if (done[i]) cudaEventDestroy(done[i]);
      if (stream[i]) cudaStreamDestroy(stream[i]);
    }
  }
};

// One device copy of a scene: device arrays, the kernel-argument structs that point at them, the ring of dynamic
// ray-fetch cursors and the staging contexts of the host-buffer entry points.  A scene has one Residency per device it
// was uploaded to; whatever was allocated is freed by the destructor (also when an upload fails half way).
struct Residency {
  int device = -1;
  void* d_top = nullptr;
  void* d_blocks = nullptr;
  EcbDeviceScene dev{};
  void* d_tri_top = nullptr;
  void* d_tri_blocks = nullptr;
  EcbDeviceTriScene tri_dev{nullptr, nullptr, ECB_REF_EMPTY, 0u, 0u};
  void* d_items = nullptr;
  void* d_instances = nullptr;
  void* d_objects = nullptr;
  void* d_pre = nullptr;  // experiment: bfloat16 pre-test boxes (ecb_format.h)
  EcbDeviceInstScene inst_dev{};
  unsigned long long* d_queues = nullptr;  // ring of device words used as dynamic ray-fetch cursors, one per launch in flight
  std::atomic<uint32_t> next_queue{0};
  // staging contexts: [0] small, [1] large
  std::mutex pool_mutex;
  std::condition_variable pool_cv;
  std::vector<Staging*> idle[2];
  int created[2] = {0, 0};
  ~Residency() {
    if (device >= 0) cudaSetDevice(device);
    for (int k = 0; k < 2; k++)
      for (Staging* st : idle[k]) delete st;
    void* ptrs[] = {d_top, d_blocks, d_tri_top, d_tri_blocks, d_items, d_instances, d_objects, d_queues, d_pre};
    for (void* p : ptrs)
      if (p) cudaFree(p);
  }
};

namespace {

Residency* find_residency(const ecbScene* s, int device) {  // caller holds s->mutex
  for (Residency* r : s->res)
    if (r->device == device) return r;
  return nullptr;
}

ecbError ensure_resident(ecbScene* s, int device, Residency** out);

cudaError_t upload_array(void** dst, const void* src, size_t bytes) {
  if (!bytes) return cudaSuccess;
  cudaError_t e = cudaMalloc(dst, bytes);
  return e != cudaSuccess ? e : cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
}

// instancing scene: object scenes first, then instance-level nodes / items / instance records / object table
ecbError upload_instanced(ecbScene* s, Residency* r) {
  std::vector<EcbDeviceScene> objs;
  for (ecbScene* o : s->objects) {
    Residency* ro = nullptr;
    ecbError e = ensure_resident(o, r->device, &ro);
    if (e != ECB_SUCCESS) return e;
    objs.push_back(ro->dev);
  }
  CU(cudaSetDevice(r->device));
  CU(upload_array(&r->d_top, s->top.data(), s->top.size() * sizeof(EcbTopNode)));
  CU(upload_array(&r->d_items, s->items.data(), s->items.size() * 4));
  CU(upload_array(&r->d_instances, s->instances.data(), s->instances.size() * sizeof(EcbInstance)));
  CU(upload_array(&r->d_objects, objs.data(), objs.size() * sizeof(EcbDeviceScene)));
  EcbDeviceInstScene& d = r->inst_dev;
  d.nodes = static_cast<const EcbTopNode*>(r->d_top);
  d.items = static_cast<const uint32_t*>(r->d_items);
  d.instances = static_cast<const EcbInstance*>(r->d_instances);
  d.objects = static_cast<const EcbDeviceScene*>(r->d_objects);
  d.root = s->root;
  d.n_nodes = (uint32_t)s->top.size();
  d.n_items = (uint32_t)s->items.size();
  d.n_instances = (uint32_t)s->instances.size();
  d.n_objects = (uint32_t)objs.size();
  d.uniform_leaf_kind = objs.empty() ? -1 : (int32_t)objs[0].leaf_kind;
  for (const EcbDeviceScene& o : objs)
    if ((int32_t)o.leaf_kind != d.uniform_leaf_kind) d.uniform_leaf_kind = -1;
  return ECB_SUCCESS;
}

ecbError upload_plain(ecbScene* s, Residency* r) {
  CU(upload_array(&r->d_top, s->top.data(), s->top.size() * sizeof(EcbTopNode)));
  CU(upload_array(&r->d_blocks, s->blocks.data(), s->blocks.size()));
  EcbDeviceScene& d = r->dev;
  d.top = static_cast<const EcbTopNode*>(r->d_top);
  d.blocks = static_cast<const uint8_t*>(r->d_blocks);
  d.root = s->root;
  d.n_top = (uint32_t)s->top.size();
  d.n_blocks = s->info.n_blocks;
  d.block_stride = s->block_stride;
  d.leaf_offset = s->leaf_offset;
  d.leaf_kind = s->info.leaf_kind;
  d.comp_level = s->info.comp_level;
  d.elems = s->elems;
  d.grid_width = s->grid_width;
  d.rcp_edges = s->rcp_edges;
  d.top_depth = s->info.top_depth;
  d.pre = nullptr;
  if (const char* e = getenv("ECB200_PRETEST"); e && atoi(e) == 1 && !s->top.empty()) {  // experiment, see ecb_format.h
    std::vector<uint16_t> pre(24 * s->top.size());
    auto down = [](float f) {  // largest bfloat16 <= f
      uint32_t u; memcpy(&u, &f, 4);
      uint32_t h = u >> 16;
      if ((u & 0xFFFFu) && (u >> 31)) h += 1;  // negative with dropped bits: one step away from zero
      return (uint16_t)h;
    };
    auto up = [](float f) {  // smallest bfloat16 >= f
      uint32_t u; memcpy(&u, &f, 4);
      uint32_t h = u >> 16;
      if ((u & 0xFFFFu) && !(u >> 31)) h += 1;
      return (uint16_t)h;
    };
    for (size_t n = 0; n < s->top.size(); n++)
      for (int p = 0; p < 6; p++)
        for (int c = 0; c < 4; c++) pre[24 * n + 4 * p + c] = (p & 1) ? up(s->top[n].bounds[p][c]) : down(s->top[n].bounds[p][c]);
    CU(upload_array(&r->d_pre, pre.data(), pre.size() * 2));
    d.pre = r->d_pre;
  }
  if (s->has_triangles()) {  // mixed scene: triangle accel next to the compressed one
    CU(upload_array(&r->d_tri_top, s->tri_top.data(), s->tri_top.size() * sizeof(EcbTopNode)));
    CU(upload_array(&r->d_tri_blocks, s->tri_blocks.data(), s->tri_blocks.size() * sizeof(EcbTriBlock)));
    r->tri_dev.nodes = static_cast<const EcbTopNode*>(r->d_tri_top);
    r->tri_dev.blocks = static_cast<const EcbTriBlock*>(r->d_tri_blocks);
    r->tri_dev.root = s->tri_root;
    r->tri_dev.n_nodes = (uint32_t)s->tri_top.size();
    r->tri_dev.n_blocks = (uint32_t)s->tri_blocks.size();
  }
  return ECB_SUCCESS;
}

// Makes the scene resident on `device` (-1: the primary residency, created on the calling thread's current device when
// there is none yet) and returns that residency.  Leaves the calling thread's current device set to it.
ecbError ensure_resident(ecbScene* s, int device, Residency** out) {
  if (device < 0) {
    if (Residency* p = s->primary.load(std::memory_order_acquire)) {
      *out = p;
      return ECB_SUCCESS;
    }
  }
  std::lock_guard<std::mutex> lock(s->mutex);
  if (device < 0 && !s->res.empty()) {
    *out = s->res[0];
    return ECB_SUCCESS;
  }
  if (device >= 0)
    if (Residency* r = find_residency(s, device)) {
      *out = r;
      return ECB_SUCCESS;
    }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_last_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return ECB_ERROR_NO_DEVICE;
  }
  if (device < 0) CU(cudaGetDevice(&device));
  CU(cudaSetDevice(device));
  Residency* r = new (std::nothrow) Residency;
  if (!r) return ECB_ERROR_OUT_OF_MEMORY;
  r->device = device;
  ecbError err = s->instanced ? upload_instanced(s, r) : upload_plain(s, r);
  if (err == ECB_SUCCESS) {
    cudaError_t ce = cudaMalloc(reinterpret_cast<void**>(&r->d_queues), kQueueRing * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMemset(r->d_queues, 0, kQueueRing * sizeof(unsigned long long));
    if (ce != cudaSuccess) err = cuda_fail(ce, "ray-fetch cursor allocation");
  }
  if (err != ECB_SUCCESS) {
    delete r;  // frees whatever the failed upload had allocated
    return err;
  }
  s->res.push_back(r);
  if (s->res.size() == 1) {
    s->info.device = device;
    s->primary.store(r, std::memory_order_release);
  }
  *out = r;
  return ECB_SUCCESS;
}

unsigned long long* next_queue(Residency* r) { return r->d_queues + (r->next_queue.fetch_add(1) % kQueueRing); }

int env_int(const char* name, int dflt, int lo, int hi) {
  const char* e = getenv(name);
  int v = e ? atoi(e) : dflt;
  return v < lo ? lo : (v > hi ? hi : v);
}

ecbError create_staging(int device, bool large, Staging** out) {
  Staging* st = new (std::nothrow) Staging;
  if (!st) return ECB_ERROR_OUT_OF_MEMORY;
  st->device = device;
  st->n_stages = large ? kStages : kSmallStages;
  st->cap = large ? kChunkRaysMax : (size_t)kInlineRaysMax;
  const size_t cmp_bytes = st->cap * 32 + 64;  // per chunk: geomID[n] (padded to 16 B) + {tfar,u,v,primID}[n] (+ instID[n]: instancing; + Ng[3n]: mixed scenes)
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < st->n_stages && e == cudaSuccess; i++) {
    e = cudaStreamCreateWithFlags(&st->stream[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&st->done[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&st->d_buf[i], st->cap * kRayHitBytes);
    if (e == cudaSuccess) e = cudaMalloc(&st->d_cmp[i], cmp_bytes);
    if (e == cudaSuccess) e = cudaMallocHost(&st->h_cmp[i], cmp_bytes);
    if (e == cudaSuccess) e = cudaMallocHost(&st->h_pk[i], st->cap * kPackedBytes);
  }
  if (e != cudaSuccess) {
    delete st;
    return cuda_fail(e, "staging allocation");
  }
  *out = st;
  return ECB_SUCCESS;
}

// Hands a staging context of the residency to the calling thread for the duration of one call.  Small contexts are
// created on demand (one per concurrent caller, kept for reuse); large ones up to ECB200_HOST_CONTEXTS (default 2: two
// long streams overlap, a third waits for one of them).
struct StagingLease {
  Residency* r = nullptr;
  Staging* st = nullptr;
  int kind = 0;
  ~StagingLease() {
    if (!st) return;
    {
      std::lock_guard<std::mutex> l(r->pool_mutex);
      r->idle[kind].push_back(st);
    }
    r->pool_cv.notify_one();
  }
};

ecbError acquire_staging(Residency* r, bool large, StagingLease* lease) {
  static const int kLargeMax = env_int("ECB200_HOST_CONTEXTS", 2, 1, 16);
  const int kind = large ? 1 : 0, limit = large ? kLargeMax : 256;
  lease->r = r;
  lease->kind = kind;
  std::unique_lock<std::mutex> l(r->pool_mutex);
  for (;;) {
    if (!r->idle[kind].empty()) {
      lease->st = r->idle[kind].back();
      r->idle[kind].pop_back();
      return ECB_SUCCESS;
    }
    if (r->created[kind] < limit) {
      r->created[kind]++;
      l.unlock();
      Staging* st = nullptr;
      const ecbError ce = create_staging(r->device, large, &st);
      if (ce != ECB_SUCCESS) {
        l.lock();
        r->created[kind]--;
        return ce;
      }
      lease->st = st;
      return ECB_SUCCESS;
    }
    r->pool_cv.wait(l);
  }
}

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// ---- host worker pool: scatters compact results into the caller's records while the next chunks are in flight ----
// and, with the packed strategy, packs the rays of the chunks ahead.  (ECB200_HOST_THREADS overrides the size; default
// min(24, 3/4 of the hardware threads / LOCAL_WORLD_SIZE): measured 8 / 10 / 12 / 14 / 16 threads on the 16-core host = 530 / 660 / 680 / 690 /
// 550 Mrays/s.)  Process-wide, never torn down.
class HostPool {
 public:
  static HostPool& get() {
    static HostPool* pool = new HostPool;  // leaked on purpose: no destruction-order issues with the CUDA runtime at exit
    return *pool;
  }
  void submit(std::function<void()> job) {
    {
      std::lock_guard<std::mutex> lock(mutex_);
      jobs_.push_back(std::move(job));
      pending_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_one();
  }
  size_t size() const { return threads_.size(); }

 private:
  HostPool() {
    unsigned n = std::thread::hardware_concurrency() * 3 / 4;
    // one process per GPU on a shared host (torchrun exports LOCAL_WORLD_SIZE): the processes split the host's cores
    if (const char* w = getenv("LOCAL_WORLD_SIZE")) {
      const int lw = atoi(w);
      if (lw > 1) n /= (unsigned)lw;
    }
    if (n < 2) n = 2;
    if (n > 24) n = 24;  // 3/4 of the cores: 12 of 16 on the single-GPU host (measured optimum), 24 of 32 on the 8-GPU host (984 vs 881 Mrays/s)
    if (const char* e = getenv("ECB200_HOST_THREADS")) n = (unsigned)atoi(e);
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    for (unsigned i = 0; i < n; i++) {
      threads_.emplace_back([this] { run(); });
      threads_.back().detach();
    }
  }
  void run() {
    for (;;) {
      std::function<void()> job;
      {
        std::unique_lock<std::mutex> lock(mutex_);
        if (jobs_.empty()) {
          // the jobs of one call arrive in bursts a few tens of microseconds apart (pack, then scatter as the chunks come
          // back): poll that long before going to sleep, a futex wake-up costs more than the wait it would save
          lock.unlock();
          const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(kPoolSpinUs);
          bool found = false;
          while (!found && std::chrono::steady_clock::now() < until) {
            for (int i = 0; i < 64; i++) cpu_relax();
            found = pending_.load(std::memory_order_acquire) != 0;
          }
          lock.lock();
          cv_.wait(lock, [this] { return !jobs_.empty(); });
        }
        job = std::move(jobs_.front());
        jobs_.pop_front();
        pending_.fetch_sub(1, std::memory_order_relaxed);
      }
      job();
    }
  }
  static void cpu_relax() {
#if defined(__SSE2__)
    _mm_pause();
#endif
  }
  static constexpr int kPoolSpinUs = 200;
  std::atomic<int> pending_{0};
  std::mutex mutex_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> jobs_;
  std::vector<std::thread> threads_;
};

// rays per pipeline stage of the host-buffer entry points (ECB200_HOST_CHUNK overrides for experiments)
size_t pipeline_chunk() {
  static const size_t n = [] {
    const char* e = getenv("ECB200_HOST_CHUNK");
    size_t v = e ? (size_t)strtoull(e, nullptr, 10) : (size_t)(1u << 16);
    if (v < 1024) v = 1024;
    if (v > kChunkRaysMax) v = kChunkRaysMax;
    return v;
  }();
  return n;
}

// Transfer strategy of the host-buffer entry points (measured on B200 / PCIe Gen5 x16: 55 GB/s one way, 35+35 GB/s both
// ways at once; the host is a 16-core VM whose cores stream 5-9 GB/s each -- bytes over the link AND bytes through the host
// cores decide the end-to-end rate; bench step = 3.8 M rays, 332 MB of records):
//   packed           : the host pool packs the 32 bytes per ray this path reads ({org,tnear},{dir,tfar}) into pinned
//                      staging buffers (prefetch + streaming stores) three chunks ahead of the upload, 32 bytes per ray
//                      go up, 20 bytes per ray ({tfar,u,v,primID} + geomID) come back into a pinned landing zone and
//                      the pool writes them into the caller's records while the next chunks are in flight (a miss
//                      writes nothing, as the API demands).  Pinned and pageable caller memory alike
//                                                                        -> 650-710 Mrays/s (12 pool threads; 8: 530)
//   compact          : the full records go up (DMA straight from pinned caller memory; pageable memory is copied into
//                      bounce buffers by the caller's thread), compact results come back as above
//                      (ECB200_HOST_PATH=compact)                          -> 504-521 Mrays/s pinned, 96 pageable
//   copy             : the full records travel both ways (ECB200_HOST_PATH=copy)   -> 452-462 Mrays/s pinned, 53 pageable
// Instancing scenes (hits carry the instance's geomID as instID[0]: one more word per ray in the result records) and mixed
// scenes (triangle hits carry a real normal: three more floats) take the same strategies.
// Default (no ECB200_HOST_PATH): packed for pageable caller memory and whenever the process has a full pool of 12 threads,
// compact for pinned caller memory on a host whose cores are shared between ranks (see run_host).
// Tried and dropped: results written by the kernel straight into the caller's mapped pinned array 363 Mrays/s (small PCIe
// writes), the same with a 48-byte 2-D H2D copy 35 Mrays/s (row-sized DMAs); packing with plain stores and 2 x 4 pack
// jobs in flight 210-310 Mrays/s (one core's streaming rate bounds a pack job: the look-ahead has to cover it).
enum HostPath { kPathAuto, kPathPacked, kPathCompact, kPathCopy };
HostPath host_path() {
  static const HostPath path = [] {
    const char* e = getenv("ECB200_HOST_PATH");
    if (e && strcmp(e, "copy") == 0) return kPathCopy;
    if (e && strcmp(e, "compact") == 0) return kPathCompact;
    if (e && strcmp(e, "packed") == 0) return kPathPacked;
    return kPathAuto;
  }();
  return path;
}

// copies the fields this path reads of rays [b, e) of a chunk into 32-byte packed records: {org.xyz, tnear}, {dir.xyz, tfar}
// (RTCRay layout, rtcore_ray.h:26-44: org 0, tnear 12, dir 16, time 28, tfar 32; time/mask/id/flags are not read by the
// compressed-accel intersectors)
void pack_rays(const uint8_t* rays, size_t stride, uint8_t* dst, size_t b, size_t e) {
#if defined(__SSE2__)
  // the source is read once and the staging buffer is read next by the DMA engine: prefetch ahead, streaming stores
  static const bool nt = !(getenv("ECB200_PACK_NT") && atoi(getenv("ECB200_PACK_NT")) == 0);
  if (nt) {
    for (size_t i = b; i < e; i++) {
      const uint8_t* r = rays + i * stride;
      __builtin_prefetch(r + 1024);
      const __m128 o = _mm_loadu_ps(reinterpret_cast<const float*>(r));
      const __m128 d = _mm_loadu_ps(reinterpret_cast<const float*>(r + 16));
      const __m128 t = _mm_load_ss(reinterpret_cast<const float*>(r + 32));
      const __m128 zt = _mm_shuffle_ps(d, t, _MM_SHUFFLE(0, 0, 2, 2));                        // dir.z, dir.z, tfar, tfar
      _mm_stream_ps(reinterpret_cast<float*>(dst + i * kPackedBytes), o);
      _mm_stream_ps(reinterpret_cast<float*>(dst + i * kPackedBytes + 16), _mm_shuffle_ps(d, zt, _MM_SHUFFLE(2, 0, 1, 0)));  // dir.xyz, tfar
    }
    _mm_sfence();
    return;
  }
  for (size_t i = b; i < e; i++) {
    const uint8_t* r = rays + i * stride;
    __builtin_prefetch(r + 1024);
    const __m128 o = _mm_loadu_ps(reinterpret_cast<const float*>(r));
    const __m128 d = _mm_loadu_ps(reinterpret_cast<const float*>(r + 16));
    const __m128 t = _mm_load_ss(reinterpret_cast<const float*>(r + 32));
    const __m128 zt = _mm_shuffle_ps(d, t, _MM_SHUFFLE(0, 0, 2, 2));
    _mm_store_ps(reinterpret_cast<float*>(dst + i * kPackedBytes), o);
    _mm_store_ps(reinterpret_cast<float*>(dst + i * kPackedBytes + 16), _mm_shuffle_ps(d, zt, _MM_SHUFFLE(2, 0, 1, 0)));
  }
#else
  for (size_t i = b; i < e; i++) {
    const uint8_t* r = rays + i * stride;
    uint32_t w[8];
    memcpy(w, r, 28);
    memcpy(w + 7, r + 32, 4);
    memcpy(dst + i * kPackedBytes, w, 32);
  }
#endif
}

enum { kExtraNone = 0, kExtraInst = 1, kExtraNg = 2 };  // what follows {tfar,u,v,primID}[n] in a chunk's compact results

// writes the compact results of rays [b, e) of a chunk into the caller's records (hit semantics of rtcIntersect1 /
// rtcOccluded1: compressed.h:570-591, bvh_intersector1.cpp:199)
void scatter_results(uint8_t* rays, size_t stride, const uint8_t* cmp, size_t n, size_t b, size_t e, bool any_hit, uint32_t instID,
                     int extra = kExtraNone) {
  const uint32_t* geom = reinterpret_cast<const uint32_t*>(cmp);
  const float* rec = reinterpret_cast<const float*>(cmp + ((4 * n + 15) & ~size_t(15)));
  // instancing scenes: instID[0] of a hit is the geomID of the instance (instance_intersector.cpp:45-83), one word per ray;
  // mixed scenes: triangle hits carry their geometric normal (intersector_epilog.h:227-306), three floats per ray
  const uint8_t* tail_arrays = cmp + ((4 * n + 15) & ~size_t(15)) + 16 * n;
  const uint32_t* inst = extra == kExtraInst ? reinterpret_cast<const uint32_t*>(tail_arrays) : nullptr;
  const float* ng = extra == kExtraNg ? reinterpret_cast<const float*>(tail_arrays) : nullptr;
  const float ninf = -__builtin_inff();
  constexpr size_t kAhead = 16;  // the caller's records are not in cache: fetch the lines of the hits ahead (measured on the B200 host: -25 %)
  for (size_t i = b; i < e; i++) {
    if (i + kAhead < e && geom[i + kAhead] != 0xFFFFFFFFu) {
      __builtin_prefetch(rays + (i + kAhead) * stride + 32, 1);
      __builtin_prefetch(rays + (i + kAhead) * stride + 79, 1);
    }
    const uint32_t g = geom[i];
    if (g == 0xFFFFFFFFu) continue;
    uint8_t* r = rays + i * stride;
    if (any_hit) {
      memcpy(r + 32, &ninf, 4);
      continue;
    }
    const float* q = rec + 4 * i;
    const float ngu[4] = {ng ? ng[3 * i] : 1.f, ng ? ng[3 * i + 1] : 0.f, ng ? ng[3 * i + 2] : 0.f, q[1]};  // Ng (dummy (1,0,0) for compressed hits, compressed.h:574), u
    uint32_t tail[4];
    memcpy(&tail[0], &q[2], 4);  // v
    memcpy(&tail[1], &q[3], 4);  // primID
    tail[2] = g;
    tail[3] = inst ? inst[i] : instID;
    memcpy(r + 32, &q[0], 4);
    memcpy(r + 48, ngu, 16);
    memcpy(r + 64, tail, 16);
  }
}

// ---- SoA streams (rtcIntersectNp / rtcOccludedNp): the eight arrays the path reads are contiguous, so the pack jobs of the
// packed pipeline read eight sequential streams (32 bytes per ray instead of the 80-byte records of the AoS call) and the
// scatter jobs write the hits into the caller's nine result arrays.  (A direct variant -- eight DMAs per chunk straight
// from the caller's pinned arrays into SoA device buffers that the kernels read coalesced through EcbLaunch::soa -- needs
// no host work on the way up but ran at half the rate: 336 vs 646 Mrays/s on the bench step, eight 256 KB copies per
// 64 Ki-ray chunk keep the copy engine at a fraction of the link rate.  The kernels keep the SoA input mode for
// device-resident arrays: ecbIntersectDeviceSoA.)
struct NpView {
  float* in[8];  // org_x, org_y, org_z, tnear, dir_x, dir_y, dir_z, tfar
  float *Ng_x, *Ng_y, *Ng_z, *u, *v;
  uint32_t *primID, *geomID, *instID;
};

void scatter_results_np(const NpView& a, size_t first, const uint8_t* cmp, size_t n, size_t b, size_t e, bool any_hit, uint32_t instID, int extra) {
  const uint32_t* geom = reinterpret_cast<const uint32_t*>(cmp);
  const float* rec = reinterpret_cast<const float*>(cmp + ((4 * n + 15) & ~size_t(15)));
  const uint8_t* tail_arrays = cmp + ((4 * n + 15) & ~size_t(15)) + 16 * n;
  const uint32_t* inst = extra == kExtraInst ? reinterpret_cast<const uint32_t*>(tail_arrays) : nullptr;
  const float* ng = extra == kExtraNg ? reinterpret_cast<const float*>(tail_arrays) : nullptr;
  const float ninf = -__builtin_inff();
  for (size_t i = b; i < e; i++) {
    const uint32_t g = geom[i];
    if (g == 0xFFFFFFFFu) continue;  // a miss writes nothing
    const size_t j = first + i;
    if (any_hit) {
      a.in[7][j] = ninf;
      continue;
    }
    const float* q = rec + 4 * i;
    a.in[7][j] = q[0];
    a.Ng_x[j] = ng ? ng[3 * i] : 1.f;  // dummy (1,0,0) for compressed hits (compressed.h:574)
    a.Ng_y[j] = ng ? ng[3 * i + 1] : 0.f;
    a.Ng_z[j] = ng ? ng[3 * i + 2] : 0.f;
    a.u[j] = q[1];
    a.v[j] = q[2];
    memcpy(&a.primID[j], &q[3], 4);
    a.geomID[j] = g;
    a.instID[j] = inst ? inst[i] : instID;
  }
}

void pack_rays_np(const NpView& a, size_t first, uint8_t* dst, size_t b, size_t e) {
  const float* ox = a.in[0] + first; const float* oy = a.in[1] + first; const float* oz = a.in[2] + first; const float* tn = a.in[3] + first;
  const float* dx = a.in[4] + first; const float* dy = a.in[5] + first; const float* dz = a.in[6] + first; const float* tf = a.in[7] + first;
  for (size_t i = b; i < e; i++) {
    float w[8] = {ox[i], oy[i], oz[i], tn[i], dx[i], dy[i], dz[i], tf[i]};
#if defined(__SSE2__)
    _mm_stream_ps(reinterpret_cast<float*>(dst + i * kPackedBytes), _mm_loadu_ps(w));
    _mm_stream_ps(reinterpret_cast<float*>(dst + i * kPackedBytes + 16), _mm_loadu_ps(w + 4));
#else
    memcpy(dst + i * kPackedBytes, w, 32);
#endif
  }
#if defined(__SSE2__)
  _mm_sfence();
#endif
}

// Host-buffer entry points on a compressed-accel scene, packed strategy (see host_path()).  Per chunk of rays:
//   pool: pack -> caller's thread: H2D (32 B/ray), kernel, D2H (20 B/ray, 4 for any-hit), event -> completion thread: waits
//   for the event, hands the scatter jobs to the pool -> pool: scatter into the caller's records.
// Pool threads never wait on the GPU; the caller's thread stays kPackAhead chunks ahead with the pack jobs.
ecbError run_host_packed(ecbScene* s, Residency* r, Staging* st, bool occluded, bool exact, uint32_t instID, uint8_t* base, uint64_t M, size_t stride,
                         const NpView* np = nullptr, size_t np_first = 0, bool direct_ok = false) {
  static const int kPackParts = env_int("ECB200_PACK_PARTS", kPackPartsDefault, 1, 64);
  // Hybrid upload (pinned, densely packed caller records only): a fraction of the chunks is not packed by the pool but DMA'd
  // straight from the caller's array as full records -- more bytes over the link (80 instead of 32 per ray), no host-core
  // and less host-memory work for those chunks.  The pool and the link then both run at capacity instead of the link idling
  // behind the pack jobs.  ECB200_DIRECT_PERCENT = share of direct chunks in percent (0 = always pack).
  static const int kDirectPercent = env_int("ECB200_DIRECT_PERCENT", kDirectPercentDefault, 0, 100);
  const bool hybrid = direct_ok && !np && kDirectPercent > 0 && M > kInlineRaysMax;
  auto is_direct = [&](uint64_t c) { return hybrid && ((c + 1) * (uint64_t)kDirectPercent) / 100 > (c * (uint64_t)kDirectPercent) / 100; };
  const size_t elem = occluded ? kRayBytes : kRayHitBytes;
  static const int kPackAheadCfg = env_int("ECB200_PACK_AHEAD", kPackAheadDefault, 1, kStages - 2);
  const int nst = st->n_stages;
  const int kPackAhead = kPackAheadCfg < nst - 1 ? kPackAheadCfg : (nst > 1 ? nst - 1 : 1);
  // rays per chunk: the configured size for long calls; a mid-sized call (a tile, a scanline block) is cut into about four
  // chunks (each chunk costs the caller's thread ~45 us of job hand-offs and CUDA calls, so not more) so that its pack, upload, traversal, download and scatter overlap instead of running one after the other
  size_t chunk = pipeline_chunk() < st->cap ? pipeline_chunk() : st->cap;
  if (M > kInlineRaysMax && M < 4 * chunk) {
    size_t c4 = (size_t)((M + 3) / 4);
    c4 = (c4 + 1023) & ~size_t(1023);
    if (c4 < 16384) c4 = 16384;
    if (c4 < chunk) chunk = c4;
  }
  const uint64_t n_chunks = (M + chunk - 1) / chunk;
  const bool any_hit = occluded;
  auto chunk_size = [&](uint64_t c) { return (size_t)((M - c * chunk) < chunk ? (M - c * chunk) : chunk); };
  const bool inst = s->instanced;  // hits of an instancing scene carry the instance's geomID as instID[0]: one more word per ray
  const bool mixed = s->has_triangles();  // triangle hits carry a real normal: three more floats per ray
  const int extra = inst ? kExtraInst : mixed ? kExtraNg : kExtraNone;
  auto cmp_bytes = [&](size_t n) { return ((4 * n + 15) & ~size_t(15)) + (any_hit ? 0 : 16 * n + (inst ? 4 * n : 0) + (mixed ? 12 * n : 0)); };
  auto issue = [&](uint64_t c) -> ecbError {  // H2D of the packed records, kernel, D2H of the compact results
    const int k = (int)(c % nst);
    const size_t n = chunk_size(c);
    cudaStream_t q = st->stream[k];
    const bool direct = is_direct(c);
    if (direct) CU(cudaMemcpyAsync(st->d_buf[k], base + c * chunk * stride, n * elem, cudaMemcpyHostToDevice, q));
    else CU(cudaMemcpyAsync(st->d_buf[k], st->h_pk[k], n * kPackedBytes, cudaMemcpyHostToDevice, q));
    EcbLaunch L;
    L.in = static_cast<const uint8_t*>(st->d_buf[k]);
    L.in_stride = direct ? elem : kPackedBytes;
    L.tfar_offset = direct ? 32 : 28;
    L.out = nullptr;
    L.out_stride = 0;
    L.M = n;
    L.instID = instID;
    L.any_hit_exact = exact ? 1u : 0u;
    L.hit_geom = static_cast<uint32_t*>(st->d_cmp[k]);
    L.hit_rec = reinterpret_cast<float4*>(static_cast<uint8_t*>(st->d_cmp[k]) + ((4 * n + 15) & ~size_t(15)));
    L.hit_inst = inst ? reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(L.hit_rec) + 16 * n) : nullptr;
    L.hit_ng = mixed ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(L.hit_rec) + 16 * n) : nullptr;
    L.block_out = nullptr;
    L.counters = nullptr;
    L.queue = next_queue(r);
    g_launches.fetch_add(1);
    if (mixed) {  // AccelN order: triangle accel first, then the compressed accel on the same staging records
      cudaError_t te = ecb_launch_triangles(r->tri_dev, r->device, L, occluded, q);
      if (te != cudaSuccess) return cuda_fail(te, "kernel launch");
      L.keep_results = 1;
      g_launches.fetch_add(1);
    }
    cudaError_t le = inst                  ? ecb_launch_instanced(r->inst_dev, r->device, L, occluded, q)
                     : (occluded && !exact) ? ecb_launch_occluded(r->dev, r->device, L, q)
                                            : ecb_launch_intersect(r->dev, r->device, L, q);
    if (le != cudaSuccess) return cuda_fail(le, "kernel launch");
    CU(cudaMemcpyAsync(st->h_cmp[k], st->d_cmp[k], cmp_bytes(n), cudaMemcpyDeviceToHost, q));
    return ECB_SUCCESS;
  };

  if (M <= kInlineRaysMax) {  // small calls (rtcIntersect1, packets, short streams): everything inline on the caller's thread
    for (uint64_t c0 = 0; c0 < n_chunks; c0 += nst) {
      const uint64_t c1 = c0 + nst < n_chunks ? c0 + nst : n_chunks;
      for (uint64_t c = c0; c < c1; c++) {
        if (np) pack_rays_np(*np, np_first + (size_t)(c * chunk), static_cast<uint8_t*>(st->h_pk[c % nst]), 0, chunk_size(c));
        else pack_rays(base + c * chunk * stride, stride, static_cast<uint8_t*>(st->h_pk[c % nst]), 0, chunk_size(c));
        ecbError e = issue(c);
        if (e != ECB_SUCCESS) return e;
      }
      for (uint64_t c = c0; c < c1; c++) {
        CU(cudaStreamSynchronize(st->stream[c % nst]));
        const size_t n = chunk_size(c);
        if (np) scatter_results_np(*np, np_first + (size_t)(c * chunk), static_cast<const uint8_t*>(st->h_cmp[c % nst]), n, 0, n, any_hit, instID, extra);
        else scatter_results(base + c * chunk * stride, stride, static_cast<const uint8_t*>(st->h_cmp[c % nst]), n, 0, n, any_hit, instID, extra);
      }
    }
    return ECB_SUCCESS;
  }

  const NpView np_copy = np ? *np : NpView{};  // the completion thread and the pool jobs outlive nothing, but capture by value anyway
  const bool is_np = np != nullptr;
  {
    std::lock_guard<std::mutex> wl(st->mutex);
    st->worker_error = cudaSuccess;
    st->gpu_issued = 0;
    st->scat_submitted = 0;
    st->stop = false;
  }
  const int device = r->device;
  // completion thread: chunk by chunk, waits for the chunk's results and hands the scatter jobs to the pool
  auto completion_body = [=] {
    cudaError_t se = cudaSetDevice(device);
    for (uint64_t c = 0; c < n_chunks; c++) {
      const int k = (int)(c % nst);
      {
        std::unique_lock<std::mutex> wl(st->mutex);
        st->cv.wait(wl, [&] { return st->gpu_issued > c || st->stop; });
        if (st->gpu_issued <= c) break;  // the call ended early
      }
      cudaError_t ce = se != cudaSuccess ? se : cudaEventSynchronize(st->done[k]);
      const size_t n = (size_t)((M - c * chunk) < chunk ? (M - c * chunk) : chunk);
      uint8_t* src = base + c * chunk * stride;
      const uint8_t* cmp = static_cast<const uint8_t*>(st->h_cmp[k]);
      {
        std::lock_guard<std::mutex> wl(st->mutex);
        if (ce != cudaSuccess && st->worker_error == cudaSuccess) st->worker_error = ce;
        st->pending[k] = ce == cudaSuccess ? kScatterParts : 0;
      }
      if (ce == cudaSuccess) {
        for (int part = 0; part < kScatterParts; part++) {
          const size_t b = n * part / kScatterParts, e = n * (part + 1) / kScatterParts;
          const size_t np_base = np_first + (size_t)(c * chunk);
          HostPool::get().submit([=] {
            if (is_np) scatter_results_np(np_copy, np_base, cmp, n, b, e, any_hit, instID, extra);
            else scatter_results(src, stride, cmp, n, b, e, any_hit, instID, extra);
            std::lock_guard<std::mutex> wl(st->mutex);
            if (--st->pending[k] == 0) st->cv.notify_all();
          });
        }
      }
      {
        std::lock_guard<std::mutex> wl(st->mutex);
        st->scat_submitted = c + 1;
      }
      st->cv.notify_all();
    }
  };
  std::thread completion;
  try {
    completion = std::thread(completion_body);
  } catch (const std::exception& ex) {  // never let an exception cross the C ABI
    g_last_error = std::string("cannot start the completion thread: ") + ex.what();
    return ECB_ERROR_OUT_OF_MEMORY;
  }
  // whatever way this call ends (CUDA errors included), no pool job and no completion thread may outlive it: they read and
  // write the caller's array
  struct Drain {
    Staging* st;
    std::thread* completion;
    ~Drain() {
      {
        std::lock_guard<std::mutex> wl(st->mutex);
        st->stop = true;
      }
      st->cv.notify_all();
      completion->join();
      std::unique_lock<std::mutex> wl(st->mutex);
      st->cv.wait(wl, [&] {
        for (int k = 0; k < st->n_stages; k++)
          if (st->pending[k] || st->pack_pending[k]) return false;
        return true;
      });
    }
  } drain{st, &completion};

  for (uint64_t c = 0; c < n_chunks + kPackAhead; c++) {
    if (c < n_chunks) {
      const int k = (int)(c % nst);
      const size_t n = chunk_size(c);
      const int parts = is_direct(c) ? 0 : (int)(n / 4096 < 1 ? 1 : (n / 4096 > (size_t)kPackParts ? (size_t)kPackParts : n / 4096));  // >= 4 Ki rays per pack job; none for a direct chunk
      {
        std::unique_lock<std::mutex> wl(st->mutex);
        if (c >= (uint64_t)nst)  // the stage's previous chunk is done down to its last scatter job
          st->cv.wait(wl, [&] { return st->scat_submitted > c - nst && st->pending[k] == 0; });
        st->pack_pending[k] = parts;
      }
      const uint8_t* src = base + c * chunk * stride;
      uint8_t* dst = static_cast<uint8_t*>(st->h_pk[k]);
      for (int part = 0; part < parts; part++) {
        const size_t b = n * part / parts, e = n * (part + 1) / parts;
        const size_t np_base = np_first + (size_t)(c * chunk);
        HostPool::get().submit([=] {
          if (is_np) pack_rays_np(np_copy, np_base, dst, b, e);
          else pack_rays(src, stride, dst, b, e);
          std::lock_guard<std::mutex> wl(st->mutex);
          if (--st->pack_pending[k] == 0) st->cv.notify_all();
        });
      }
    }
    if (c >= (uint64_t)kPackAhead) {
      const uint64_t d = c - kPackAhead;
      const int k = (int)(d % nst);
      {
        std::unique_lock<std::mutex> wl(st->mutex);
        st->cv.wait(wl, [&] { return st->pack_pending[k] == 0; });
      }
      ecbError e = issue(d);
      if (e != ECB_SUCCESS) return e;
      CU(cudaEventRecord(st->done[k], st->stream[k]));
      {
        std::lock_guard<std::mutex> wl(st->mutex);
        st->gpu_issued = d + 1;
      }
      st->cv.notify_all();
    }
  }
  {
    std::unique_lock<std::mutex> wl(st->mutex);
    st->cv.wait(wl, [&] {
      if (st->scat_submitted < n_chunks) return false;
      for (int k = 0; k < st->n_stages; k++)
        if (st->pending[k]) return false;
      return true;
    });
    if (st->worker_error != cudaSuccess) return cuda_fail(st->worker_error, "result scatter");
  }
  return ECB_SUCCESS;
}

// Transfer strategy of one host-buffer call (see host_path()).  auto: pageable caller memory has to go through host threads
// anyway -> packed.  Pinned caller memory can be DMA'd as it is, which costs no host cycles: packed wins when the process
// has a full pool (12 threads: 650 vs 510 Mrays/s on the 16-core single-GPU host), compact when the host's cores are shared
// -- between ranks (2 ranks on 24 cores, 9 threads each: 743 packed vs 855 compact) or between the devices of one scene
// (ecbSceneUploadMulti: the links add up, the pool does not).
HostPath choose_path(bool pinned, size_t n_devices) {
  const HostPath p = host_path();
  if (p != kPathAuto) return p;
  if (!pinned) return kPathPacked;
  if (n_devices > 1) return kPathCompact;
  return HostPool::get().size() >= 12 ? kPathPacked : kPathCompact;
}

// One host-buffer call (or one device's share of it) on one residency.  Shared by intersect (elem = 80) and occluded (48).
ecbError run_host_on(ecbScene* s, Residency* r, bool occluded, uint32_t instID, void* rays, uint64_t M, size_t stride, bool exact,
                     HostPath path) {
  const size_t elem = occluded ? kRayBytes : kRayHitBytes;
  CU(cudaSetDevice(r->device));
  StagingLease lease;
  ecbError err = acquire_staging(r, M > kInlineRaysMax, &lease);
  if (err != ECB_SUCCESS) return err;
  Staging* st = lease.st;
  const int nst = st->n_stages;
  uint8_t* base = static_cast<uint8_t*>(rays);
  const size_t chunk = pipeline_chunk() < st->cap ? pipeline_chunk() : st->cap;
  const uint64_t n_chunks = (M + chunk - 1) / chunk;
  const bool pinned = is_pinned(rays);
  const bool compact = path != kPathCopy;
  const int extra = s->instanced ? kExtraInst : s->has_triangles() ? kExtraNg : kExtraNone;
  if (path == kPathPacked) return run_host_packed(s, r, st, occluded, exact, instID, base, M, stride, nullptr, 0, pinned && stride == elem);
  const bool use_pool = compact && M > 8192;  // small calls (rtcIntersect1, packets) scatter inline

  auto launch = [&](int k, size_t n) -> cudaError_t {
    EcbLaunch L;
    L.in = static_cast<const uint8_t*>(st->d_buf[k]);
    L.in_stride = elem;
    L.out = compact ? nullptr : static_cast<uint8_t*>(st->d_buf[k]);
    L.out_stride = elem;
    L.M = n;
    L.instID = instID;
    L.any_hit_exact = exact ? 1u : 0u;
    L.hit_geom = compact ? static_cast<uint32_t*>(st->d_cmp[k]) : nullptr;
    L.hit_rec = compact ? reinterpret_cast<float4*>(static_cast<uint8_t*>(st->d_cmp[k]) + ((4 * n + 15) & ~size_t(15))) : nullptr;
    L.hit_inst = (compact && s->instanced) ? reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(L.hit_rec) + 16 * n) : nullptr;
    L.hit_ng = (compact && s->has_triangles()) ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(L.hit_rec) + 16 * n) : nullptr;
    L.block_out = nullptr;
    L.counters = nullptr;
    L.queue = next_queue(r);
    g_launches.fetch_add(1);
    if (s->instanced) return ecb_launch_instanced(r->inst_dev, r->device, L, occluded, st->stream[k]);
    if (s->has_triangles()) {  // AccelN order: triangle accel first, in place, then the compressed accel on the same rays
      cudaError_t te = ecb_launch_triangles(r->tri_dev, r->device, L, occluded, st->stream[k]);
      if (te != cudaSuccess) return te;
      L.keep_results = compact ? 1u : 0u;
      g_launches.fetch_add(1);
    }
    if (exact) return ecb_launch_intersect(r->dev, r->device, L, st->stream[k]);
    return occluded ? ecb_launch_occluded(r->dev, r->device, L, st->stream[k]) : ecb_launch_intersect(r->dev, r->device, L, st->stream[k]);
  };

  if (!pinned) {
    for (int i = 0; i < st->n_stages; i++)
      if (!st->h_pin[i]) CU(cudaMallocHost(&st->h_pin[i], st->cap * kRayHitBytes));
  }
  auto chunk_size = [&](uint64_t c) { return (size_t)((M - c * chunk) < chunk ? (M - c * chunk) : chunk); };
  auto unpack = [&](uint64_t c, int k) {  // full-record mode, pageable caller memory: bounce buffer -> caller's array
    const size_t n = chunk_size(c);
    uint8_t* dst = base + c * chunk * stride;
    const uint8_t* hb = static_cast<const uint8_t*>(st->h_pin[k]);
    if (stride == elem) memcpy(dst, hb, n * elem);
    else for (size_t i = 0; i < n; i++) memcpy(dst + i * stride, hb + i * elem, elem);
  };
  auto wait_stage = [&](int k) {  // the scatter jobs of the chunk that last used stage k are done
    std::unique_lock<std::mutex> wl(st->mutex);
    st->cv.wait(wl, [&] { return st->pending[k] == 0; });
  };
  st->worker_error = cudaSuccess;
  const bool any_hit = occluded;  // exact any-hit runs through the occluded entry points as well
  // whatever way this call ends (CUDA errors included), no scatter job may outlive it: they write into the caller's array
  struct Drain {
    Staging* st;
    ~Drain() {
      std::unique_lock<std::mutex> wl(st->mutex);
      st->cv.wait(wl, [&] {
        for (int k = 0; k < st->n_stages; k++)
          if (st->pending[k]) return false;
        return true;
      });
    }
  } drain{st};

  for (uint64_t c = 0; c < n_chunks; c++) {
    const int k = (int)(c % nst);
    const size_t n = chunk_size(c);
    uint8_t* src = base + c * chunk * stride;
    cudaStream_t q = st->stream[k];
    if (use_pool) wait_stage(k);
    if (pinned) {
      if (stride == elem) CU(cudaMemcpyAsync(st->d_buf[k], src, n * elem, cudaMemcpyHostToDevice, q));
      else CU(cudaMemcpy2DAsync(st->d_buf[k], elem, src, stride, elem, n, cudaMemcpyHostToDevice, q));
    } else {
      // pageable caller memory: bounce through pinned buffers; first drain this stage's previous round trip
      CU(cudaStreamSynchronize(q));
      if (!compact && c >= (uint64_t)nst) unpack(c - nst, k);
      uint8_t* hb = static_cast<uint8_t*>(st->h_pin[k]);
      if (stride == elem) memcpy(hb, src, n * elem);
      else for (size_t i = 0; i < n; i++) memcpy(hb + i * elem, src + i * stride, elem);
      CU(cudaMemcpyAsync(st->d_buf[k], hb, n * elem, cudaMemcpyHostToDevice, q));
    }
    cudaError_t le = launch(k, n);
    if (le != cudaSuccess) return cuda_fail(le, "kernel launch");
    if (compact) {
      const size_t bytes = ((4 * n + 15) & ~size_t(15)) + (any_hit ? 0 : 16 * n + (s->instanced ? 4 * n : 0) + (s->has_triangles() ? 12 * n : 0));
      CU(cudaMemcpyAsync(st->h_cmp[k], st->d_cmp[k], bytes, cudaMemcpyDeviceToHost, q));
      if (use_pool) {
        CU(cudaEventRecord(st->done[k], q));
        {
          std::lock_guard<std::mutex> wl(st->mutex);
          st->pending[k] = kScatterParts;
        }
        for (int part = 0; part < kScatterParts; part++) {
          const size_t b = n * part / kScatterParts, e = n * (part + 1) / kScatterParts;
          const uint8_t* cmp = static_cast<const uint8_t*>(st->h_cmp[k]);
          const int device = r->device;
          cudaEvent_t ev = st->done[k];
          HostPool::get().submit([=] {
            cudaError_t ce = cudaSetDevice(device);
            if (ce == cudaSuccess) ce = cudaEventSynchronize(ev);
            if (ce == cudaSuccess) scatter_results(src, stride, cmp, n, b, e, any_hit, instID, extra);
            std::lock_guard<std::mutex> wl(st->mutex);
            if (ce != cudaSuccess && st->worker_error == cudaSuccess) st->worker_error = ce;
            if (--st->pending[k] == 0) st->cv.notify_all();
          });
        }
      }
    } else if (pinned) {
      if (stride == elem) CU(cudaMemcpyAsync(src, st->d_buf[k], n * elem, cudaMemcpyDeviceToHost, q));
      else CU(cudaMemcpy2DAsync(src, stride, st->d_buf[k], elem, elem, n, cudaMemcpyDeviceToHost, q));
    } else {
      CU(cudaMemcpyAsync(st->h_pin[k], st->d_buf[k], n * elem, cudaMemcpyDeviceToHost, q));
    }
  }
  if (use_pool) {
    for (int k = 0; k < st->n_stages; k++) wait_stage(k);
    if (st->worker_error != cudaSuccess) return cuda_fail(st->worker_error, "result scatter");
    return ECB_SUCCESS;
  }
  for (int k = 0; k < st->n_stages; k++) CU(cudaStreamSynchronize(st->stream[k]));
  if (compact) {  // small call: scatter inline
    for (uint64_t c = 0; c < n_chunks; c++) {
      const size_t n = chunk_size(c);
      scatter_results(base + c * chunk * stride, stride, static_cast<const uint8_t*>(st->h_cmp[c % nst]), n, 0, n, any_hit, instID, extra);
    }
  } else if (!pinned) {
    const uint64_t tail0 = n_chunks > (uint64_t)nst ? n_chunks - nst : 0;
    for (uint64_t c = tail0; c < n_chunks; c++) unpack(c, (int)(c % nst));
  }
  return ECB_SUCCESS;
}

ecbError run_host_np_on(ecbScene* s, Residency* r, bool occluded, uint32_t instID, const NpView& a, size_t first, uint64_t M) {
  CU(cudaSetDevice(r->device));
  StagingLease lease;
  ecbError err = acquire_staging(r, M > kInlineRaysMax, &lease);
  if (err != ECB_SUCCESS) return err;
  return run_host_packed(s, r, lease.st, occluded, false, instID, nullptr, M, 0, &a, first);
}

ecbError run_host_np(ecbScene* s, bool occluded, uint32_t instID, const NpView& a, uint64_t M) {
  if (!s) return ECB_ERROR_INVALID_ARGUMENT;
  if (M == 0) return ECB_SUCCESS;
  for (int k = 0; k < 8; k++)
    if (!a.in[k]) return ECB_ERROR_INVALID_ARGUMENT;
  if (!occluded && (!a.Ng_x || !a.Ng_y || !a.Ng_z || !a.u || !a.v || !a.primID || !a.geomID || !a.instID)) return ECB_ERROR_INVALID_ARGUMENT;
  Residency* r0 = nullptr;
  ecbError err = ensure_resident(s, -1, &r0);
  if (err != ECB_SUCCESS) return err;
  std::vector<Residency*> devs;
  if (M >= kMultiDeviceRaysMin) {
    std::lock_guard<std::mutex> lock(s->mutex);
    devs = s->res;
  }
  if (devs.size() <= 1) return run_host_np_on(s, r0, occluded, instID, a, 0, M);
  const size_t n = devs.size();
  std::vector<ecbError> errs(n, ECB_SUCCESS);
  std::vector<std::string> msgs(n);
  std::vector<std::thread> threads;
  auto part = [&](size_t i) {
    const uint64_t per = ((M + n - 1) / n + 1023) & ~uint64_t(1023);
    const uint64_t b = per * i < M ? per * i : M, e = per * (i + 1) < M ? per * (i + 1) : M;
    if (e <= b) return;
    errs[i] = run_host_np_on(s, devs[i], occluded, instID, a, (size_t)b, e - b);
    if (errs[i] != ECB_SUCCESS) msgs[i] = g_last_error;
  };
  try {
    for (size_t i = 1; i < n; i++) threads.emplace_back(part, i);
  } catch (const std::exception& ex) {
    for (auto& t : threads) t.join();
    g_last_error = std::string("cannot start a device thread: ") + ex.what();
    return ECB_ERROR_OUT_OF_MEMORY;
  }
  part(0);
  for (auto& t : threads) t.join();
  cudaSetDevice(r0->device);
  for (size_t i = 0; i < n; i++)
    if (errs[i] != ECB_SUCCESS) {
      g_last_error = "device " + std::to_string(devs[i]->device) + ": " + msgs[i];
      return errs[i];
    }
  return ECB_SUCCESS;
}

// Entry of every host-buffer call: argument checks, lazy upload, and -- for a scene resident on several devices
// (ecbSceneUploadMulti) and a call long enough to pay for it -- the fan-out: the ray range is cut into one contiguous part
// per device, each part runs the single-device pipeline on its own thread (its own staging context, stream set and PCIe
// link), all of them share the host pool, and the results land in the caller's array -- there is no gather step.
ecbError run_host(ecbScene* s, bool occluded, uint32_t instID, void* rays, uint64_t M, size_t stride, bool exact = false) {
  if (!s || (!rays && M)) return ECB_ERROR_INVALID_ARGUMENT;
  if (exact && s->instanced) {
    g_last_error = "exact any-hit is not available on instancing scenes";
    return ECB_ERROR_INVALID_OPERATION;
  }
  const size_t elem = occluded ? kRayBytes : kRayHitBytes;
  if (stride < elem || (stride & 3)) {
    g_last_error = "byteStride must be a multiple of 4 and at least the struct size";
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  if (M == 0) return ECB_SUCCESS;
  Residency* r0 = nullptr;
  ecbError err = ensure_resident(s, -1, &r0);
  if (err != ECB_SUCCESS) return err;
  std::vector<Residency*> devs;
  if (M >= kMultiDeviceRaysMin) {
    std::lock_guard<std::mutex> lock(s->mutex);
    devs = s->res;
  }
  const bool pinned = is_pinned(rays);
  if (devs.size() <= 1) return run_host_on(s, r0, occluded, instID, rays, M, stride, exact, choose_path(pinned, 1));
  const HostPath path = choose_path(pinned, devs.size());
  const size_t n = devs.size();
  std::vector<ecbError> errs(n, ECB_SUCCESS);
  std::vector<std::string> msgs(n);
  std::vector<std::thread> threads;
  auto part = [&](size_t i) {
    // contiguous parts, multiples of 1024 rays
    const uint64_t per = ((M + n - 1) / n + 1023) & ~uint64_t(1023);
    const uint64_t b = per * i < M ? per * i : M, e = per * (i + 1) < M ? per * (i + 1) : M;
    if (e <= b) return;
    errs[i] = run_host_on(s, devs[i], occluded, instID, static_cast<uint8_t*>(rays) + b * stride, e - b, stride, exact, path);
    if (errs[i] != ECB_SUCCESS) msgs[i] = g_last_error;
  };
  try {
    for (size_t i = 1; i < n; i++) threads.emplace_back(part, i);
  } catch (const std::exception& ex) {  // never let an exception cross the C ABI
    for (auto& t : threads) t.join();
    g_last_error = std::string("cannot start a device thread: ") + ex.what();
    return ECB_ERROR_OUT_OF_MEMORY;
  }
  part(0);
  for (auto& t : threads) t.join();
  cudaSetDevice(r0->device);
  for (size_t i = 0; i < n; i++)
    if (errs[i] != ECB_SUCCESS) {
      g_last_error = "device " + std::to_string(devs[i]->device) + ": " + msgs[i];
      return errs[i];
    }
  return ECB_SUCCESS;
}

// residency a device pointer belongs to: the only one, or the one on the device that owns the allocation
ecbError residency_for_pointer(ecbScene* s, const void* d_ptr, Residency** out) {
  Residency* r0 = nullptr;
  ecbError err = ensure_resident(s, -1, &r0);
  if (err != ECB_SUCCESS) return err;
  *out = r0;
  std::lock_guard<std::mutex> lock(s->mutex);
  if (s->res.size() <= 1 || !d_ptr) return ECB_SUCCESS;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, d_ptr) != cudaSuccess) {
    cudaGetLastError();
    return ECB_SUCCESS;
  }
  if (a.type == cudaMemoryTypeDevice)
    if (Residency* r = find_residency(s, a.device)) *out = r;
  return ECB_SUCCESS;
}

ecbError run_device(ecbScene* s, bool occluded, uint32_t instID, void* d_rays, uint64_t M, size_t stride, uint32_t* d_block_out,
                    void* stream, unsigned long long* d_counters = nullptr, bool exact = false, const ecbDeviceHits* compact_out = nullptr) {
  if (!s || (!d_rays && M)) return ECB_ERROR_INVALID_ARGUMENT;
  const size_t elem = occluded ? kRayBytes : kRayHitBytes;
  if (stride < elem || (stride & 15) || (reinterpret_cast<uintptr_t>(d_rays) & 15)) {
    g_last_error = "device rays must be 16-byte aligned with a byteStride that is a multiple of 16";
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  Residency* r = nullptr;
  ecbError err = residency_for_pointer(s, d_rays, &r);
  if (err != ECB_SUCCESS) return err;
  if (M == 0) return ECB_SUCCESS;
  cudaStream_t q = static_cast<cudaStream_t>(stream);
  EcbLaunch L;
  L.in = static_cast<const uint8_t*>(d_rays);
  L.in_stride = stride;
  L.out = static_cast<uint8_t*>(d_rays);
  L.out_stride = stride;
  L.M = M;
  L.instID = instID;
  L.any_hit_exact = exact ? 1u : 0u;
  L.hit_rec = nullptr;
  L.hit_geom = nullptr;
  L.block_out = d_block_out;
  L.counters = d_counters;
  L.queue = next_queue(r);
  if (s->instanced && d_block_out) {
    g_last_error = "block output is not available on instancing scenes";
    return ECB_ERROR_INVALID_OPERATION;
  }
  if (exact && s->instanced) {
    g_last_error = "exact any-hit is not available on instancing scenes";
    return ECB_ERROR_INVALID_OPERATION;
  }
  if (compact_out) {  // compact result records instead of in-place write-back (what a multi-GPU frame gather moves)
    if (!compact_out->geomID || (!occluded && !compact_out->hit) || (s->instanced && !occluded && !compact_out->instID) ||
        (s->has_triangles() && !occluded && !compact_out->Ng)) {
      g_last_error = "compact output needs geomID[], hit[] (closest hit), instID[] (instancing scenes) and Ng[] (mixed scenes)";
      return ECB_ERROR_INVALID_ARGUMENT;
    }
    L.out = nullptr;
    L.hit_geom = compact_out->geomID;
    L.hit_rec = reinterpret_cast<float4*>(compact_out->hit);
    L.hit_inst = s->instanced ? compact_out->instID : nullptr;
    L.hit_ng = s->has_triangles() ? compact_out->Ng : nullptr;
  }
  if (!s->instanced && s->has_triangles()) {
    if (compact_out) {
      g_last_error = "compact output of a mixed scene needs writable input records (the triangle pass hands tfar to the second pass through them): use the in-place call";
      return ECB_ERROR_INVALID_OPERATION;
    }
    cudaError_t te = ecb_launch_triangles(r->tri_dev, r->device, L, occluded, q);
    if (te != cudaSuccess) return cuda_fail(te, "kernel launch");
    g_launches.fetch_add(1);
  }
  cudaError_t le = s->instanced ? ecb_launch_instanced(r->inst_dev, r->device, L, occluded, q)
                   : exact      ? ecb_launch_intersect(r->dev, r->device, L, q)
                   : occluded   ? ecb_launch_occluded(r->dev, r->device, L, q)
                                : ecb_launch_intersect(r->dev, r->device, L, q);
  if (le != cudaSuccess) return cuda_fail(le, "kernel launch");
  g_launches.fetch_add(1);
  return ECB_SUCCESS;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
extern "C" {

ecbError ecbSceneImportBVH4(const ecbImportDesc* desc, ecbScene** scene_out) { return ecb_import_bvh4(desc, scene_out); }
ecbError ecbSceneImportInstances(const ecbInstanceDesc* desc, ecbScene** scene_out) { return ecb_import_instances(desc, scene_out); }

ecbError ecbSceneAttachTriangles4v(ecbScene* scene, const ecbTriangleDesc* desc) { return ecb_import_triangles(desc, scene); }

ecbError ecbSceneImportImage(const void* image, size_t bytes, ecbScene** scene_out) {
  // header of oracle/rimg.h restated as offsets (plain data: magic[8], 8 x u32, root u64, total u64, bounds[6])
  if (!image || bytes < 96 || memcmp(image, "ECBRIMG1", 8) != 0) {
    g_last_error = "not an RIMG image";
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  const uint8_t* p = static_cast<const uint8_t*>(image);
  uint32_t w[8];
  memcpy(w, p + 8, sizeof(w));
  uint64_t root, total;
  memcpy(&root, p + 40, 8);
  memcpy(&total, p + 48, 8);
  if (w[0] != 1 || total > bytes) {
    g_last_error = "RIMG version/size mismatch";
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  ecbImportDesc d;
  memset(&d, 0, sizeof(d));
  d.base = image;
  d.image_bytes = total;
  d.root = root;
  d.leaf_kind = w[1];
  d.subdiv_level = w[2];
  d.comp_level = w[3];
  d.proj_col_major = w[4];
  memcpy(d.bounds, p + 56, 24);
  ecbError e = ecb_import_bvh4(&d, scene_out);
  uint64_t tri_root;
  memcpy(&tri_root, p + 80, 8);  // mixed scene: the triangle accel travels in the same image
  if (e == ECB_SUCCESS && tri_root) {
    ecbTriangleDesc t;
    t.base = image;
    t.image_bytes = total;
    t.root = tri_root;
    e = ecb_import_triangles(&t, *scene_out);
    if (e != ECB_SUCCESS) {
      ecbSceneRelease(*scene_out);
      *scene_out = nullptr;
    }
  }
  return e;
}

}  // extern "C"

static bool read_file(const char* path, std::vector<uint8_t>& buf) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) return false;
  const std::streamoff end = f.tellg();
  if (end < 0) return false;
  buf.resize((size_t)end);
  f.seekg(0);
  if (end > 0 && !f.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)end)) return false;
  return true;
}

extern "C" ecbError ecbSceneImportImageFile(const char* path, ecbScene** scene_out) {
  if (!path || !scene_out) return ECB_ERROR_INVALID_ARGUMENT;
  std::vector<uint8_t> buf;
  try {
    if (!read_file(path, buf)) {
      g_last_error = std::string("cannot read ") + path;
      return ECB_ERROR_INVALID_ARGUMENT;
    }
  } catch (const std::exception&) {
    return ECB_ERROR_OUT_OF_MEMORY;
  }
  return ecbSceneImportImage(buf.data(), buf.size(), scene_out);
}

// ---- native blob (ecbSceneSave / ecbSceneLoad) -----------------------------------------------------------------------
// plain scene:      EcbBlobHeader | top nodes | blocks | [EcbBlobTriTrailer | triangle nodes | triangle blocks]
// instancing scene: EcbBlobInstHeader | n_objects x { u64 bytes | plain-scene blob } | instance-level nodes | items | instances
namespace {

struct EcbBlobInstHeader {
  char magic[8];  // "ECB200I1"
  uint32_t version, n_objects, n_instances, n_items, n_nodes, root, depth, reserved;
  float bounds[6];
};

void append(std::vector<uint8_t>& out, const void* p, size_t n) {
  const uint8_t* b = static_cast<const uint8_t*>(p);
  out.insert(out.end(), b, b + n);
}

void save_plain(const ecbScene* s, std::vector<uint8_t>& out) {
  EcbBlobHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "ECB200S1", 8);
  h.version = 1;
  h.leaf_kind = s->info.leaf_kind;
  h.subdiv_level = s->info.subdiv_level;
  h.comp_level = s->info.comp_level;
  h.n_top = (uint32_t)s->top.size();
  h.n_blocks = s->info.n_blocks;
  h.block_stride = s->block_stride;
  h.leaf_offset = s->leaf_offset;
  h.root = s->root;
  h.top_depth = s->info.top_depth;
  h.elems = s->elems;
  h.grid_width = s->grid_width;
  h.rcp_edges = s->rcp_edges;
  memcpy(h.bounds, s->info.bounds, sizeof(h.bounds));
  h.reference_bytes = s->info.reference_bytes;
  h.top_bytes = s->top.size() * sizeof(EcbTopNode);
  h.block_bytes = s->blocks.size();
  append(out, &h, sizeof(h));
  append(out, s->top.data(), h.top_bytes);
  append(out, s->blocks.data(), h.block_bytes);
  if (s->has_triangles()) {
    EcbBlobTriTrailer t;
    memset(&t, 0, sizeof(t));
    memcpy(t.magic, "ECB200T1", 8);
    t.n_nodes = (uint32_t)s->tri_top.size();
    t.n_blocks = (uint32_t)s->tri_blocks.size();
    t.root = s->tri_root;
    t.depth = s->tri_depth;
    append(out, &t, sizeof(t));
    append(out, s->tri_top.data(), s->tri_top.size() * sizeof(EcbTopNode));
    append(out, s->tri_blocks.data(), s->tri_blocks.size() * sizeof(EcbTriBlock));
  }
}

// The kernels trust the arrays, so a blob is checked before it becomes a scene.  `leaf_ok(ref)` validates a leaf reference.
// Nodes are emitted in pre-order, so every child node index must be larger than its parent's: that rules out cycles (and
// shared subtrees are harmless).  Returns the height of the tree (the importer's max leaf depth), or -1.
template <typename LeafOk>
int validate_tree(const std::vector<EcbTopNode>& nodes, uint32_t root, LeafOk leaf_ok) {
  const uint32_t n = (uint32_t)nodes.size();
  if (root == ECB_REF_EMPTY) return 0;
  if (root & ECB_REF_LEAF) return leaf_ok(root) ? 0 : -1;
  if (root >= n) return -1;
  std::vector<uint16_t> height(n, 0);
  for (uint32_t i = n; i-- > 0;) {
    uint32_t h = 0;
    for (int k = 0; k < 4; k++) {
      const uint32_t c = nodes[i].child[k];
      if (c == ECB_REF_EMPTY) continue;
      if (c & ECB_REF_LEAF) {
        if (!leaf_ok(c)) return -1;
        h = h > 1 ? h : 1;
      } else {
        if (c <= i || c >= n) return -1;
        h = h > (uint32_t)height[c] + 1 ? h : (uint32_t)height[c] + 1;
      }
    }
    if (h > 1000) return -1;
    height[i] = (uint16_t)h;
  }
  return (int)height[root];
}

struct Cursor {
  const uint8_t* p;
  size_t left;
  bool take(void* dst, size_t n) {
    if (n > left) return false;
    memcpy(dst, p, n);
    p += n;
    left -= n;
    return true;
  }
};

ecbError load_plain(Cursor& c, ecbScene** out) {
  EcbBlobHeader h;
  if (!c.take(&h, sizeof(h)) || memcmp(h.magic, "ECB200S1", 8) != 0 || h.version != 1) {
    g_last_error = "not an ECB200 scene blob";
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  // every derived constant is recomputed from (leaf kind, C) and must match; the array sizes must fit the data that is there
  bool ok = h.comp_level >= 1 && h.comp_level <= ECB_MAX_C && h.leaf_kind <= ECB_LEAF_BOX;
  uint32_t leaf_offset = 0, stride = 0;
  if (ok) {
    ecb_block_layout(h.comp_level, h.leaf_kind, &leaf_offset, &stride);
    const uint32_t cells = 1u << (2 * h.comp_level);
    ok = h.block_stride == stride && h.leaf_offset == leaf_offset && h.elems == (cells - 1) / 3 && h.grid_width == (1u << h.comp_level) + 1 &&
         h.rcp_edges == 1.0f / (float)(1u << h.comp_level) && h.top_bytes == (uint64_t)h.n_top * sizeof(EcbTopNode) &&
         h.block_bytes == (uint64_t)h.n_blocks * stride && h.top_bytes <= c.left && h.block_bytes <= c.left - h.top_bytes &&
         h.n_blocks < ECB_REF_LEAF;
  }
  if (!ok) {
    g_last_error = "corrupt scene blob header";
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  ecbScene* s = new (std::nothrow) ecbScene;
  if (!s) return ECB_ERROR_OUT_OF_MEMORY;
  s->top.resize(h.n_top);
  s->blocks.resize(h.block_bytes);
  c.take(s->top.data(), h.top_bytes);
  c.take(s->blocks.data(), h.block_bytes);
  const int depth = validate_tree(s->top, h.root, [&](uint32_t ref) { return (ref & ~ECB_REF_LEAF) < h.n_blocks; });
  if (depth < 0 || 3 * depth + 4 > (int)ECB_MAX_TOP_STACK) {
    delete s;
    g_last_error = "scene blob has an out-of-range / backward child reference or a tree too deep for the traversal stack";
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  s->root = h.root;
  s->block_stride = stride;
  s->leaf_offset = leaf_offset;
  s->elems = h.elems;
  s->grid_width = h.grid_width;
  s->rcp_edges = h.rcp_edges;
  ecbSceneInfo& in = s->info;
  in.leaf_kind = h.leaf_kind;
  in.subdiv_level = h.subdiv_level;
  in.comp_level = h.comp_level;
  in.n_top_nodes = h.n_top;
  in.n_blocks = h.n_blocks;
  in.top_depth = (uint32_t)depth;
  in.block_stride = stride;
  in.top_node_bytes = sizeof(EcbTopNode);
  in.device_bytes = h.top_bytes + h.block_bytes;
  in.reference_bytes = h.reference_bytes;
  in.micro_triangles = 2ull * (1ull << (2 * h.comp_level)) * h.n_blocks;
  memcpy(in.bounds, h.bounds, sizeof(in.bounds));
  in.device = -1;
  EcbBlobTriTrailer t;
  if (c.left >= sizeof(t) && memcmp(c.p, "ECB200T1", 8) == 0) {  // optional: triangle accel of a mixed scene
    c.take(&t, sizeof(t));
    ok = t.n_blocks < (1u << 27) && (uint64_t)t.n_nodes * sizeof(EcbTopNode) <= c.left &&
         (uint64_t)t.n_blocks * sizeof(EcbTriBlock) <= c.left - (uint64_t)t.n_nodes * sizeof(EcbTopNode) && t.root != ECB_REF_EMPTY;
    int tdepth = -1;
    if (ok) {
      s->tri_top.resize(t.n_nodes);
      s->tri_blocks.resize(t.n_blocks);
      c.take(s->tri_top.data(), t.n_nodes * sizeof(EcbTopNode));
      c.take(s->tri_blocks.data(), (size_t)t.n_blocks * sizeof(EcbTriBlock));
      tdepth = validate_tree(s->tri_top, t.root, [&](uint32_t ref) {
        const uint32_t first = (ref & ~ECB_REF_LEAF) >> 3, num = ref & 7u;
        return (uint64_t)first + num <= t.n_blocks;
      });
    }
    if (tdepth < 0 || 3 * tdepth + 4 > (int)ECB_MAX_TOP_STACK) {
      delete s;
      g_last_error = "corrupt triangle section in scene blob";
      return ECB_ERROR_INVALID_ARGUMENT;
    }
    s->tri_root = t.root;
    s->tri_depth = (uint32_t)tdepth;
    in.device_bytes += t.n_nodes * sizeof(EcbTopNode) + (uint64_t)t.n_blocks * sizeof(EcbTriBlock);
  }
  *out = s;
  return ECB_SUCCESS;
}

ecbError load_instanced(Cursor& c, ecbScene** out) {
  EcbBlobInstHeader h;
  if (!c.take(&h, sizeof(h)) || h.version != 1 || h.n_objects > (1u << 20) || h.n_items >= (1u << 27)) {
    g_last_error = "corrupt instancing-scene blob header";
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  ecbScene* s = new (std::nothrow) ecbScene;
  if (!s) return ECB_ERROR_OUT_OF_MEMORY;
  s->instanced = true;
  auto fail = [&](const char* msg, ecbError e) {
    for (ecbScene* o : s->objects) ecbSceneRelease(o);
    s->objects.clear();
    delete s;
    if (msg) g_last_error = msg;
    return e;
  };
  for (uint32_t i = 0; i < h.n_objects; i++) {
    uint64_t bytes = 0;
    if (!c.take(&bytes, 8) || bytes > c.left) return fail("truncated instancing-scene blob", ECB_ERROR_INVALID_ARGUMENT);
    Cursor sub{c.p, (size_t)bytes};
    ecbScene* o = nullptr;
    const ecbError e = load_plain(sub, &o);
    if (e != ECB_SUCCESS) return fail(nullptr, e);
    if (o->has_triangles()) {
      ecbSceneRelease(o);
      return fail("instanced object scenes cannot carry a triangle accel", ECB_ERROR_INVALID_ARGUMENT);
    }
    s->objects.push_back(o);  // the instancing scene owns the reference the load created
    c.p += bytes;
    c.left -= (size_t)bytes;
  }
  const uint64_t need = (uint64_t)h.n_nodes * sizeof(EcbTopNode) + (uint64_t)h.n_items * 4 + (uint64_t)h.n_instances * sizeof(EcbInstance);
  if (need > c.left) return fail("truncated instancing-scene blob", ECB_ERROR_INVALID_ARGUMENT);
  s->top.resize(h.n_nodes);
  s->items.resize(h.n_items);
  s->instances.resize(h.n_instances);
  c.take(s->top.data(), h.n_nodes * sizeof(EcbTopNode));
  c.take(s->items.data(), (size_t)h.n_items * 4);
  c.take(s->instances.data(), (size_t)h.n_instances * sizeof(EcbInstance));
  bool ok = true;
  for (uint32_t it : s->items) ok = ok && it < h.n_instances;
  for (const EcbInstance& in : s->instances) ok = ok && in.object < h.n_objects;
  const int depth = ok ? validate_tree(s->top, h.root, [&](uint32_t ref) {
    const uint32_t first = (ref & ~ECB_REF_LEAF) >> 3, num = ref & 7u;
    return (uint64_t)first + num <= h.n_items;
  }) : -1;
  if (depth < 0 || 3 * depth + 4 > (int)ECB_MAX_TOP_STACK) return fail("instancing-scene blob has an out-of-range reference", ECB_ERROR_INVALID_ARGUMENT);
  s->root = h.root;
  ecbSceneInfo& in = s->info;
  memset(&in, 0, sizeof(in));
  in.n_top_nodes = h.n_nodes;
  in.top_depth = (uint32_t)depth;
  in.top_node_bytes = sizeof(EcbTopNode);
  in.device_bytes = s->top.size() * sizeof(EcbTopNode) + s->items.size() * 4 + s->instances.size() * sizeof(EcbInstance);
  for (const EcbInstance& i : s->instances) {
    in.micro_triangles += s->objects[i.object]->info.micro_triangles;
    in.n_blocks += s->objects[i.object]->info.n_blocks;
  }
  for (ecbScene* o : s->objects) in.device_bytes += o->info.device_bytes;
  in.leaf_kind = s->objects.empty() ? 0 : s->objects[0]->info.leaf_kind;
  memcpy(in.bounds, h.bounds, sizeof(in.bounds));
  in.device = -1;
  *out = s;
  return ECB_SUCCESS;
}

}  // namespace

extern "C" {

ecbError ecbSceneSave(const ecbScene* s, const char* path) {
  if (!s || !path) return ECB_ERROR_INVALID_ARGUMENT;
  try {
    std::vector<uint8_t> out;
    if (s->instanced) {
      EcbBlobInstHeader h;
      memset(&h, 0, sizeof(h));
      memcpy(h.magic, "ECB200I1", 8);
      h.version = 1;
      h.n_objects = (uint32_t)s->objects.size();
      h.n_instances = (uint32_t)s->instances.size();
      h.n_items = (uint32_t)s->items.size();
      h.n_nodes = (uint32_t)s->top.size();
      h.root = s->root;
      h.depth = s->info.top_depth;
      memcpy(h.bounds, s->info.bounds, sizeof(h.bounds));
      append(out, &h, sizeof(h));
      for (const ecbScene* o : s->objects) {
        std::vector<uint8_t> sub;
        save_plain(o, sub);
        const uint64_t bytes = sub.size();
        append(out, &bytes, 8);
        append(out, sub.data(), sub.size());
      }
      append(out, s->top.data(), s->top.size() * sizeof(EcbTopNode));
      append(out, s->items.data(), s->items.size() * 4);
      append(out, s->instances.data(), s->instances.size() * sizeof(EcbInstance));
    } else {
      save_plain(s, out);
    }
    std::ofstream f(path, std::ios::binary);
    if (!f) {
      g_last_error = std::string("cannot write ") + path;
      return ECB_ERROR_INVALID_ARGUMENT;
    }
    f.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)out.size());
    return f ? ECB_SUCCESS : ECB_ERROR_UNKNOWN;
  } catch (const std::exception&) {  // never let an exception cross the C ABI
    return ECB_ERROR_OUT_OF_MEMORY;
  }
}

ecbError ecbSceneLoad(const char* path, ecbScene** out) {
  if (!path || !out) return ECB_ERROR_INVALID_ARGUMENT;
  try {
    std::vector<uint8_t> buf;
    if (!read_file(path, buf) || buf.size() < 8) {
      g_last_error = std::string("not an ECB200 scene blob: ") + path;
      return ECB_ERROR_INVALID_ARGUMENT;
    }
    Cursor c{buf.data(), buf.size()};
    if (memcmp(buf.data(), "ECB200I1", 8) == 0) return load_instanced(c, out);
    return load_plain(c, out);
  } catch (const std::exception&) {  // std::bad_alloc / length_error from a crafted size: never across the C ABI
    g_last_error = "out of memory while loading the scene blob";
    return ECB_ERROR_OUT_OF_MEMORY;
  }
}

ecbError ecbSceneUpload(ecbScene* s, int device) {
  if (!s) return ECB_ERROR_INVALID_ARGUMENT;
  Residency* r = nullptr;
  return ensure_resident(s, device, &r);
}

ecbError ecbSceneUploadMulti(ecbScene* s, const int* devices, uint32_t n_devices) {
  if (!s || (n_devices && !devices)) return ECB_ERROR_INVALID_ARGUMENT;
  int current = -1;
  cudaGetDevice(&current);
  ecbError err = ECB_SUCCESS;
  for (uint32_t i = 0; i < n_devices && err == ECB_SUCCESS; i++) {
    if (devices[i] < 0) {
      g_last_error = "device ordinals must be >= 0";
      err = ECB_ERROR_INVALID_ARGUMENT;
      break;
    }
    Residency* r = nullptr;
    err = ensure_resident(s, devices[i], &r);
  }
  if (current >= 0) cudaSetDevice(current);  // uploads switch the calling thread's device: put it back
  return err;
}

uint32_t ecbSceneResidentDevices(const ecbScene* s, int* devices_out, uint32_t capacity) {
  if (!s) return 0;
  std::lock_guard<std::mutex> lock(s->mutex);
  for (uint32_t i = 0; i < s->res.size() && i < capacity && devices_out; i++) devices_out[i] = s->res[i]->device;
  return (uint32_t)s->res.size();
}

ecbError ecbSceneGetInfo(const ecbScene* s, ecbSceneInfo* info) {
  if (!s || !info) return ECB_ERROR_INVALID_ARGUMENT;
  *info = s->info;
  return ECB_SUCCESS;
}

void ecbSceneRelease(ecbScene* s) {
  if (!s) return;
  if (--s->refs > 0) return;  // still referenced by an instancing scene
  for (ecbScene* o : s->objects) ecbSceneRelease(o);
  int current = -1;
  if (!s->res.empty()) cudaGetDevice(&current);
  for (Residency* r : s->res) delete r;  // frees the device arrays and the staging contexts on the residency's device
  if (current >= 0) cudaSetDevice(current);
  delete s;
}

ecbError ecbIntersect1(ecbScene* s, uint32_t instID, void* rayhit) { return run_host(s, false, instID, rayhit, 1, kRayHitBytes); }
ecbError ecbIntersect1M(ecbScene* s, uint32_t instID, void* rayhit, uint32_t M, size_t byteStride) {
  return run_host(s, false, instID, rayhit, M, byteStride);
}
ecbError ecbOccluded1(ecbScene* s, uint32_t instID, void* ray) { return run_host(s, true, instID, ray, 1, kRayBytes); }
ecbError ecbOccluded1M(ecbScene* s, uint32_t instID, void* ray, uint32_t M, size_t byteStride) {
  return run_host(s, true, instID, ray, M, byteStride);
}

static NpView np_view(const ecbRayNp* ray, const ecbHitNp* hit) {
  NpView a;
  memset(&a, 0, sizeof(a));
  float* const in[8] = {ray->org_x, ray->org_y, ray->org_z, ray->tnear, ray->dir_x, ray->dir_y, ray->dir_z, ray->tfar};
  for (int k = 0; k < 8; k++) a.in[k] = in[k];
  if (hit) {
    a.Ng_x = hit->Ng_x; a.Ng_y = hit->Ng_y; a.Ng_z = hit->Ng_z;
    a.u = hit->u; a.v = hit->v;
    a.primID = hit->primID; a.geomID = hit->geomID; a.instID = hit->instID;
  }
  return a;
}
ecbError ecbIntersectNp(ecbScene* s, uint32_t instID, const ecbRayHitNp* rayhit, uint64_t N) {
  if (!rayhit) return ECB_ERROR_INVALID_ARGUMENT;
  return run_host_np(s, false, instID, np_view(&rayhit->ray, &rayhit->hit), N);
}
ecbError ecbOccludedNp(ecbScene* s, uint32_t instID, const ecbRayNp* ray, uint64_t N) {
  if (!ray) return ECB_ERROR_INVALID_ARGUMENT;
  return run_host_np(s, true, instID, np_view(ray, nullptr), N);
}

ecbError ecbOccludedExact1M(ecbScene* s, uint32_t instID, void* ray, uint32_t M, size_t byteStride) {
  return run_host(s, true, instID, ray, M, byteStride, true);
}
ecbError ecbOccludedExactDevice(ecbScene* s, uint32_t instID, void* d_ray, uint64_t M, size_t byteStride, void* stream) {
  return run_device(s, true, instID, d_ray, M, byteStride, nullptr, stream, nullptr, true);
}

// RTCRayHit16 (include/embree3/rtcore_ray.h:149-187) is the SoA transpose of 16 RTCRayHit: 20 arrays of 16 words.
ecbError ecbIntersect16(const int* valid, ecbScene* s, uint32_t instID, void* rayhit16) {
  if (!valid || !rayhit16) return ECB_ERROR_INVALID_ARGUMENT;
  alignas(16) uint32_t aos[16][20];
  uint32_t* soa = static_cast<uint32_t*>(rayhit16);
  for (int l = 0; l < 16; l++) {
    for (int f = 0; f < 20; f++) aos[l][f] = soa[f * 16 + l];
    if (valid[l] != -1) {  // inactive lane -> inactive stream ray (tnear > tfar)
      const float one = 1.f, zero = 0.f;
      memcpy(&aos[l][3], &one, 4);
      memcpy(&aos[l][8], &zero, 4);
    }
  }
  ecbError e = run_host(s, false, instID, aos, 16, kRayHitBytes);
  if (e != ECB_SUCCESS) return e;
  for (int l = 0; l < 16; l++) {
    if (valid[l] != -1) continue;
    soa[8 * 16 + l] = aos[l][8];
    for (int f = 12; f < 20; f++) soa[f * 16 + l] = aos[l][f];
  }
  return ECB_SUCCESS;
}

ecbError ecbIntersectDevice(ecbScene* s, uint32_t instID, void* d_rayhit, uint64_t M, size_t byteStride, void* stream) {
  return run_device(s, false, instID, d_rayhit, M, byteStride, nullptr, stream);
}
ecbError ecbIntersectDeviceEx(ecbScene* s, uint32_t instID, void* d_rayhit, uint64_t M, size_t byteStride, uint32_t* d_block_out,
                              void* stream) {
  return run_device(s, false, instID, d_rayhit, M, byteStride, d_block_out, stream);
}
ecbError ecbIntersectDeviceStats(ecbScene* s, uint32_t instID, void* d_rayhit, uint64_t M, size_t byteStride, uint64_t* d_counters,
                                 void* stream) {
  if (!d_counters) return ECB_ERROR_INVALID_ARGUMENT;
  return run_device(s, false, instID, d_rayhit, M, byteStride, nullptr, stream, reinterpret_cast<unsigned long long*>(d_counters));
}
ecbError ecbOccludedDevice(ecbScene* s, uint32_t instID, void* d_ray, uint64_t M, size_t byteStride, void* stream) {
  return run_device(s, true, instID, d_ray, M, byteStride, nullptr, stream);
}

ecbError ecbIntersectDeviceCompact(ecbScene* s, uint32_t instID, const void* d_rayhit, uint64_t M, size_t byteStride,
                                   const ecbDeviceHits* out, void* stream) {
  if (!out) return ECB_ERROR_INVALID_ARGUMENT;
  return run_device(s, false, instID, const_cast<void*>(d_rayhit), M, byteStride, nullptr, stream, nullptr, false, out);
}
ecbError ecbOccludedDeviceCompact(ecbScene* s, uint32_t instID, const void* d_ray, uint64_t M, size_t byteStride, const ecbDeviceHits* out,
                                  void* stream) {
  if (!out) return ECB_ERROR_INVALID_ARGUMENT;
  return run_device(s, true, instID, const_cast<void*>(d_ray), M, byteStride, nullptr, stream, nullptr, false, out);
}

ecbError ecbIntersectDeviceSoA(ecbScene* s, uint32_t instID, const float* const d_soa[8], uint64_t M, const ecbDeviceHits* out, void* stream) {
  if (!s || !d_soa || !out || !out->geomID || !out->hit) return ECB_ERROR_INVALID_ARGUMENT;
  for (int k = 0; k < 8; k++)
    if (!d_soa[k] && M) return ECB_ERROR_INVALID_ARGUMENT;
  if (s->instanced || s->has_triangles()) {
    g_last_error = "SoA device input is available on plain compressed-accel scenes";
    return ECB_ERROR_INVALID_OPERATION;
  }
  Residency* r = nullptr;
  ecbError err = residency_for_pointer(s, d_soa[0], &r);
  if (err != ECB_SUCCESS) return err;
  if (M == 0) return ECB_SUCCESS;
  EcbLaunch L;
  L.in = nullptr;
  L.in_stride = 0;
  for (int k = 0; k < 8; k++) L.soa[k] = d_soa[k];
  L.out = nullptr;
  L.out_stride = 0;
  L.M = M;
  L.instID = instID;
  L.any_hit_exact = 0;
  L.hit_geom = out->geomID;
  L.hit_rec = reinterpret_cast<float4*>(out->hit);
  L.block_out = nullptr;
  L.counters = nullptr;
  L.queue = next_queue(r);
  cudaError_t le = ecb_launch_intersect(r->dev, r->device, L, static_cast<cudaStream_t>(stream));
  if (le != cudaSuccess) return cuda_fail(le, "kernel launch");
  g_launches.fetch_add(1);
  return ECB_SUCCESS;
}

ecbError ecbDeviceAlloc(void** d_ptr, size_t bytes) {
  if (!d_ptr) return ECB_ERROR_INVALID_ARGUMENT;
  CU(cudaMalloc(d_ptr, bytes ? bytes : 1));
  return ECB_SUCCESS;
}
void ecbDeviceFree(void* d_ptr) {
  if (d_ptr) cudaFree(d_ptr);
}
ecbError ecbIpcExport(void* d_ptr, void* handle64_out) {
  if (!d_ptr || !handle64_out) return ECB_ERROR_INVALID_ARGUMENT;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, d_ptr));
  memcpy(handle64_out, &h, sizeof(h));
  return ECB_SUCCESS;
}
ecbError ecbIpcImport(const void* handle64, void** d_ptr_out) {
  if (!handle64 || !d_ptr_out) return ECB_ERROR_INVALID_ARGUMENT;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  CU(cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return ECB_SUCCESS;
}
void ecbIpcClose(void* d_ptr) {
  if (d_ptr) cudaIpcCloseMemHandle(d_ptr);
}
ecbError ecbPeerCopyAsync(void* d_dst, const void* d_src, size_t bytes, void* stream) {
  if ((!d_dst || !d_src) && bytes) return ECB_ERROR_INVALID_ARGUMENT;
  if (!bytes) return ECB_SUCCESS;
  CU(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
  return ECB_SUCCESS;
}

int ecbBoundsCheckFailures(uint64_t counts_out[8]) {
  unsigned long long c[8] = {0};
  int enabled = 0;
  if (ecb_bounds_check_read(c, &enabled) != cudaSuccess) return -1;
  if (counts_out)
    for (int i = 0; i < 8; i++) counts_out[i] = c[i];
  return enabled;
}

void ecbSetTraversalKernel(int mode) { ecb_set_kernel_mode(mode); }
int ecbGetTraversalKernel(void) { return ecb_kernel_mode(); }

ecbError ecbProbeReadBandwidth(size_t working_set_bytes, int passes, double* gbs_out) {
  if (!gbs_out || working_set_bytes < (1u << 20) || passes < 1) return ECB_ERROR_INVALID_ARGUMENT;
  int count = 0, device = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    g_last_error = "no CUDA device";
    return ECB_ERROR_NO_DEVICE;
  }
  CU(cudaGetDevice(&device));
  CU(ecb_probe_read_bandwidth(working_set_bytes, passes, device, gbs_out));
  g_launches.fetch_add(6);
  return ECB_SUCCESS;
}

ecbError ecbSceneBlockNormals(const ecbScene* s, float* out, size_t cap) {
  if (!s || !out || s->instanced) return ECB_ERROR_INVALID_ARGUMENT;
  if (cap < 3 * (size_t)s->info.n_blocks) return ECB_ERROR_INVALID_ARGUMENT;
  for (uint32_t b = 0; b < s->info.n_blocks; b++) {
    const EcbBlockHeader* h = reinterpret_cast<const EcbBlockHeader*>(s->blocks.data() + (size_t)b * s->block_stride);
    // space = inverse([vx vy vz]) stored by columns; its third ROW is vz (unit, orthogonal to vx and vy; compressed.h:121-126)
    out[3 * b + 0] = h->space[2];
    out[3 * b + 1] = h->space[5];
    out[3 * b + 2] = h->space[8];
  }
  return ECB_SUCCESS;
}

uint32_t ecbBlockStride(uint32_t leaf_kind, uint32_t comp_level) {
  if (leaf_kind > ECB_LEAF_BOX || comp_level < 1 || comp_level > ECB_MAX_C) return 0;
  uint32_t leaf_offset, stride;
  ecb_block_layout(comp_level, leaf_kind, &leaf_offset, &stride);
  return stride;
}

ecbError ecbBuildBlocks(const ecbBuildDesc* d, void* blocks_out, float* bounds_out, uint32_t* patch_ok_out) {
  if (!d || !blocks_out || !bounds_out) return ECB_ERROR_INVALID_ARGUMENT;
  if (d->leaf_kind > ECB_LEAF_BOX || d->comp_level < 1 || d->comp_level > ECB_MAX_C) {
    g_last_error = "leaf kind / compression level out of range";
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  if (d->n_blocks == 0) return ECB_SUCCESS;
  if (!d->grids || !d->space || !d->world || !d->uv || !d->geomIDs || !d->primIDs) return ECB_ERROR_INVALID_ARGUMENT;
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0) {
    g_last_error = std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0");
    return ECB_ERROR_NO_DEVICE;
  }
  int device = 0;
  CU(cudaGetDevice(&device));
  const uint32_t n = d->n_blocks, W = (1u << d->comp_level) + 1;
  EcbBuildLaunch b{};
  ecb_block_layout(d->comp_level, d->leaf_kind, &b.leaf_offset, &b.stride);
  b.n = n;
  b.comp_level = d->comp_level;
  b.leaf_kind = d->leaf_kind;
  const size_t grid_bytes = (size_t)n * W * W * 12, out_bytes = (size_t)n * b.stride;
  // one device allocation, carved into the ten arrays (256-byte aligned)
  const size_t sizes[10] = {grid_bytes, (size_t)n * 36, (size_t)n * 36, (size_t)n * 16, (size_t)n * 4, (size_t)n * 4,
                            out_bytes, (size_t)n * 24, (size_t)n * 4, ecb_build_tmp_bytes(n)};
  size_t offs[10], total = 0;
  for (int i = 0; i < 10; i++) { offs[i] = total; total += (sizes[i] + 255) & ~size_t(255); }
  uint8_t* base = nullptr;
  ce = cudaMalloc(reinterpret_cast<void**>(&base), total);
  const void* srcs[6] = {d->grids, d->space, d->world, d->uv, d->geomIDs, d->primIDs};
  for (int i = 0; i < 6 && ce == cudaSuccess; i++) ce = cudaMemcpyAsync(base + offs[i], srcs[i], sizes[i], cudaMemcpyHostToDevice, nullptr);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(base + offs[6], 0, out_bytes, nullptr);
  if (ce == cudaSuccess) {
    b.grids = reinterpret_cast<const float*>(base + offs[0]);
    b.space = reinterpret_cast<const float*>(base + offs[1]);
    b.world = reinterpret_cast<const float*>(base + offs[2]);
    b.uv = reinterpret_cast<const float*>(base + offs[3]);
    b.geomIDs = reinterpret_cast<const uint32_t*>(base + offs[4]);
    b.primIDs = reinterpret_cast<const uint32_t*>(base + offs[5]);
    b.blocks = base + offs[6];
    b.bounds = reinterpret_cast<float*>(base + offs[7]);
    b.patch_ok = reinterpret_cast<uint32_t*>(base + offs[8]);
    b.tmp = base + offs[9];
    ce = ecb_launch_build(b, device, nullptr);
    g_launches.fetch_add(d->leaf_kind == ECB_LEAF_GRID ? 3 : 5);
  }
  if (ce == cudaSuccess) ce = cudaMemcpy(blocks_out, base + offs[6], out_bytes, cudaMemcpyDeviceToHost);
  if (ce == cudaSuccess) ce = cudaMemcpy(bounds_out, base + offs[7], (size_t)n * 24, cudaMemcpyDeviceToHost);
  if (ce == cudaSuccess && patch_ok_out) ce = cudaMemcpy(patch_ok_out, base + offs[8], (size_t)n * 4, cudaMemcpyDeviceToHost);
  if (base) cudaFree(base);
  if (ce != cudaSuccess) return cuda_fail(ce, "ecbBuildBlocks");
  return ECB_SUCCESS;
}

ecbError ecbSceneCopyBlocks(const ecbScene* s, void* out, size_t capacity) {
  if (!s || !out || s->instanced) return ECB_ERROR_INVALID_ARGUMENT;
  if (capacity < s->blocks.size()) {
    g_last_error = "buffer smaller than n_blocks x block_stride";
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  memcpy(out, s->blocks.data(), s->blocks.size());
  return ECB_SUCCESS;
}

ecbError ecbSceneSetBlocks(ecbScene* s, const void* blocks, size_t bytes) {
  if (!s || !blocks || s->instanced) return ECB_ERROR_INVALID_ARGUMENT;
  if (s->resident() || bytes != s->blocks.size()) {
    g_last_error = "blocks can be replaced before the scene is resident, by the same number of bytes";
    return ECB_ERROR_INVALID_OPERATION;
  }
  memcpy(s->blocks.data(), blocks, bytes);
  return ECB_SUCCESS;
}

ecbError ecbHostAlloc(void** ptr, size_t bytes) {
  if (!ptr) return ECB_ERROR_INVALID_ARGUMENT;
  CU(cudaMallocHost(ptr, bytes));
  return ECB_SUCCESS;
}
void ecbHostFree(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

const char* ecbHostPathName(int caller_memory_is_pinned) {
  switch (host_path()) {
    case kPathCopy: return "copy";
    case kPathCompact: return "compact";
    case kPathPacked: return "packed";
    default: return (!caller_memory_is_pinned || HostPool::get().size() >= 12) ? "packed" : "compact";
  }
}

uint64_t ecbLaunchCount(void) { return g_launches.load(); }

const char* ecbErrorString(ecbError e) {
  switch (e) {
    case ECB_SUCCESS: return "success";
    case ECB_ERROR_UNKNOWN: return "unknown error";
    case ECB_ERROR_INVALID_ARGUMENT: return "invalid argument";
    case ECB_ERROR_INVALID_OPERATION: return "invalid operation";
    case ECB_ERROR_OUT_OF_MEMORY: return "out of memory";
    case ECB_ERROR_UNSUPPORTED: return "scene uses a feature outside the compressed-BVH hot path";
    case ECB_ERROR_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
    case ECB_ERROR_CUDA: return "CUDA error";
  }
  return "?";
}
const char* ecbLastErrorMessage(void) { return g_last_error.c_str(); }

}  // extern "C"
