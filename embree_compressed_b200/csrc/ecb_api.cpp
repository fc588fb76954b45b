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
constexpr size_t kPackedBytes = 32;         // packed input record of the host-buffer path: {org.xyz,tnear},{dir.xyz,tfar}
constexpr int kPackPartsDefault = 8;        // pack jobs per chunk (ECB200_PACK_PARTS overrides for experiments)
constexpr int kPackAheadDefault = 3;        // chunks packed ahead of the one being issued to the GPU (ECB200_PACK_AHEAD)
constexpr size_t kChunkRaysMax = 1u << 17;  // capacity of a staging buffer: 128 Ki rays
constexpr size_t kCompactBytesMax = kChunkRaysMax * 32 + 64;  // per chunk: geomID[n] (padded to 16 B) + {tfar,u,v,primID}[n] (+ instID[n]: instancing; + Ng[3n]: mixed scenes)
constexpr int kScatterParts = 4;            // scatter jobs per chunk
constexpr uint32_t kQueueRing = 256;
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

// Host-buffer pipeline: kStages chunks in flight, each H2D -> kernel(s) -> D2H on its own stream over fixed device
// buffers; created lazily per scene.
struct Staging {
  int device = -1;
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

namespace {

ecbError ensure_resident(ecbScene* s, int device);

// instancing scene: object scenes first, then instance-level nodes / items / instance records / object table
ecbError upload_instanced(ecbScene* s, int device) {
  std::vector<EcbDeviceScene> objs;
  for (ecbScene* o : s->objects) {
    ecbError e = ensure_resident(o, device);
    if (e != ECB_SUCCESS) return e;
    objs.push_back(o->dev);
  }
  auto up = [](void** dst, const void* src, size_t bytes) -> cudaError_t {
    if (!bytes) return cudaSuccess;
    cudaError_t e = cudaMalloc(dst, bytes);
    return e != cudaSuccess ? e : cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  CU(up(&s->d_top, s->top.data(), s->top.size() * sizeof(EcbTopNode)));
  CU(up(&s->d_items, s->items.data(), s->items.size() * 4));
  CU(up(&s->d_instances, s->instances.data(), s->instances.size() * sizeof(EcbInstance)));
  CU(up(&s->d_objects, objs.data(), objs.size() * sizeof(EcbDeviceScene)));
  EcbDeviceInstScene& d = s->inst_dev;
  d.nodes = static_cast<const EcbTopNode*>(s->d_top);
  d.items = static_cast<const uint32_t*>(s->d_items);
  d.instances = static_cast<const EcbInstance*>(s->d_instances);
  d.objects = static_cast<const EcbDeviceScene*>(s->d_objects);
  d.root = s->root;
  d.n_nodes = (uint32_t)s->top.size();
  d.n_items = (uint32_t)s->items.size();
  d.n_instances = (uint32_t)s->instances.size();
  d.n_objects = (uint32_t)objs.size();
  CU(cudaMalloc(reinterpret_cast<void**>(&s->d_queues), kQueueRing * sizeof(unsigned long long)));
  CU(cudaMemset(s->d_queues, 0, kQueueRing * sizeof(unsigned long long)));
  s->device = device;
  s->info.device = device;
  return ECB_SUCCESS;
}

ecbError ensure_resident(ecbScene* s, int device) {
  std::lock_guard<std::mutex> lock(s->mutex);
  if (s->device >= 0) {
    if (device >= 0 && device != s->device) {
      g_last_error = "scene is resident on another device";
      return ECB_ERROR_INVALID_OPERATION;
    }
    return ECB_SUCCESS;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_last_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return ECB_ERROR_NO_DEVICE;
  }
  if (device < 0) {
    CU(cudaGetDevice(&device));
  }
  CU(cudaSetDevice(device));
  if (s->instanced) return upload_instanced(s, device);
  const size_t top_bytes = s->top.size() * sizeof(EcbTopNode);
  if (top_bytes) {
    CU(cudaMalloc(&s->d_top, top_bytes));
    CU(cudaMemcpy(s->d_top, s->top.data(), top_bytes, cudaMemcpyHostToDevice));
  }
  if (!s->blocks.empty()) {
    CU(cudaMalloc(&s->d_blocks, s->blocks.size()));
    CU(cudaMemcpy(s->d_blocks, s->blocks.data(), s->blocks.size(), cudaMemcpyHostToDevice));
  }
  EcbDeviceScene& d = s->dev;
  d.top = static_cast<const EcbTopNode*>(s->d_top);
  d.blocks = static_cast<const uint8_t*>(s->d_blocks);
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
  s->tri_dev = EcbDeviceTriScene{nullptr, nullptr, ECB_REF_EMPTY, 0u, 0u};
  if (s->has_triangles()) {  // mixed scene: triangle accel next to the compressed one
    if (!s->tri_top.empty()) {
      CU(cudaMalloc(&s->d_tri_top, s->tri_top.size() * sizeof(EcbTopNode)));
      CU(cudaMemcpy(s->d_tri_top, s->tri_top.data(), s->tri_top.size() * sizeof(EcbTopNode), cudaMemcpyHostToDevice));
    }
    if (!s->tri_blocks.empty()) {
      CU(cudaMalloc(&s->d_tri_blocks, s->tri_blocks.size() * sizeof(EcbTriBlock)));
      CU(cudaMemcpy(s->d_tri_blocks, s->tri_blocks.data(), s->tri_blocks.size() * sizeof(EcbTriBlock), cudaMemcpyHostToDevice));
    }
    s->tri_dev.nodes = static_cast<const EcbTopNode*>(s->d_tri_top);
    s->tri_dev.blocks = static_cast<const EcbTriBlock*>(s->d_tri_blocks);
    s->tri_dev.root = s->tri_root;
    s->tri_dev.n_nodes = (uint32_t)s->tri_top.size();
    s->tri_dev.n_blocks = (uint32_t)s->tri_blocks.size();
  }
  CU(cudaMalloc(reinterpret_cast<void**>(&s->d_queues), kQueueRing * sizeof(unsigned long long)));
  CU(cudaMemset(s->d_queues, 0, kQueueRing * sizeof(unsigned long long)));
  s->device = device;
  s->info.device = device;
  return ECB_SUCCESS;
}

unsigned long long* next_queue(ecbScene* s) { return s->d_queues + (s->next_queue.fetch_add(1) % kQueueRing); }

ecbError ensure_staging(ecbScene* s) {
  if (s->staging) return ECB_SUCCESS;
  Staging* st = new (std::nothrow) Staging;
  if (!st) return ECB_ERROR_OUT_OF_MEMORY;
  st->device = s->device;
  for (int i = 0; i < kStages; i++) {
    cudaError_t e = cudaStreamCreateWithFlags(&st->stream[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&st->done[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&st->d_buf[i], kChunkRaysMax * kRayHitBytes);
    if (e == cudaSuccess) e = cudaMalloc(&st->d_cmp[i], kCompactBytesMax);
    if (e == cudaSuccess) e = cudaMallocHost(&st->h_cmp[i], kCompactBytesMax);
    if (e == cudaSuccess) e = cudaMallocHost(&st->h_pk[i], kChunkRaysMax * kPackedBytes);
    if (e != cudaSuccess) {
      delete st;
      return cuda_fail(e, "staging allocation");
    }
  }
  s->staging = st;
  return ECB_SUCCESS;
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
// min(12, 3/4 of the hardware threads / LOCAL_WORLD_SIZE): measured 8 / 10 / 12 / 14 / 16 threads on the 16-core host = 530 / 660 / 680 / 690 /
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
    if (n > 12) n = 12;
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
bool use_compact_results() { return host_path() != kPathCopy; }

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

// Host-buffer entry points on a compressed-accel scene, packed strategy (see host_path()).  Per chunk of rays:
//   pool: pack -> caller's thread: H2D (32 B/ray), kernel, D2H (20 B/ray, 4 for any-hit), event -> completion thread: waits
//   for the event, hands the scatter jobs to the pool -> pool: scatter into the caller's records.
// Pool threads never wait on the GPU; the caller's thread stays kPackAhead chunks ahead with the pack jobs.
int env_int(const char* name, int dflt, int lo, int hi) {
  const char* e = getenv(name);
  int v = e ? atoi(e) : dflt;
  return v < lo ? lo : (v > hi ? hi : v);
}

ecbError run_host_packed(ecbScene* s, Staging* st, bool occluded, bool exact, uint32_t instID, uint8_t* base, uint64_t M, size_t stride) {
  static const int kPackParts = env_int("ECB200_PACK_PARTS", kPackPartsDefault, 1, 64);
  static const int kPackAhead = env_int("ECB200_PACK_AHEAD", kPackAheadDefault, 1, kStages - 2);
  // rays per chunk: the configured size for long calls; a mid-sized call (a tile, a scanline block) is cut into about four
  // chunks (each chunk costs the caller's thread ~45 us of job hand-offs and CUDA calls, so not more) so that its pack, upload, traversal, download and scatter overlap instead of running one after the other
  size_t chunk = pipeline_chunk();
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
    const int k = (int)(c % kStages);
    const size_t n = chunk_size(c);
    cudaStream_t q = st->stream[k];
    CU(cudaMemcpyAsync(st->d_buf[k], st->h_pk[k], n * kPackedBytes, cudaMemcpyHostToDevice, q));
    EcbLaunch L;
    L.in = static_cast<const uint8_t*>(st->d_buf[k]);
    L.in_stride = kPackedBytes;
    L.tfar_offset = 28;
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
    L.queue = next_queue(s);
    g_launches.fetch_add(1);
    if (mixed) {  // AccelN order: triangle accel first, then the compressed accel on the same staging records
      cudaError_t te = ecb_launch_triangles(s->tri_dev, s->device, L, occluded, q);
      if (te != cudaSuccess) return cuda_fail(te, "kernel launch");
      L.keep_results = 1;
      g_launches.fetch_add(1);
    }
    cudaError_t le = inst                  ? ecb_launch_instanced(s->inst_dev, s->device, L, occluded, q)
                     : (occluded && !exact) ? ecb_launch_occluded(s->dev, s->device, L, q)
                                            : ecb_launch_intersect(s->dev, s->device, L, q);
    if (le != cudaSuccess) return cuda_fail(le, "kernel launch");
    CU(cudaMemcpyAsync(st->h_cmp[k], st->d_cmp[k], cmp_bytes(n), cudaMemcpyDeviceToHost, q));
    return ECB_SUCCESS;
  };

  if (M <= kInlineRaysMax) {  // small calls (rtcIntersect1, packets, short streams): everything inline on the caller's thread
    for (uint64_t c0 = 0; c0 < n_chunks; c0 += kStages) {
      const uint64_t c1 = c0 + kStages < n_chunks ? c0 + kStages : n_chunks;
      for (uint64_t c = c0; c < c1; c++) {
        pack_rays(base + c * chunk * stride, stride, static_cast<uint8_t*>(st->h_pk[c % kStages]), 0, chunk_size(c));
        ecbError e = issue(c);
        if (e != ECB_SUCCESS) return e;
      }
      for (uint64_t c = c0; c < c1; c++) {
        CU(cudaStreamSynchronize(st->stream[c % kStages]));
        const size_t n = chunk_size(c);
        scatter_results(base + c * chunk * stride, stride, static_cast<const uint8_t*>(st->h_cmp[c % kStages]), n, 0, n, any_hit, instID, extra);
      }
    }
    return ECB_SUCCESS;
  }

  {
    std::lock_guard<std::mutex> wl(st->mutex);
    st->worker_error = cudaSuccess;
    st->gpu_issued = 0;
    st->scat_submitted = 0;
    st->stop = false;
  }
  const int device = s->device;
  // completion thread: chunk by chunk, waits for the chunk's results and hands the scatter jobs to the pool
  auto completion_body = [=] {
    cudaError_t se = cudaSetDevice(device);
    for (uint64_t c = 0; c < n_chunks; c++) {
      const int k = (int)(c % kStages);
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
          HostPool::get().submit([=] {
            scatter_results(src, stride, cmp, n, b, e, any_hit, instID, extra);
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
        for (int k = 0; k < kStages; k++)
          if (st->pending[k] || st->pack_pending[k]) return false;
        return true;
      });
    }
  } drain{st, &completion};

  for (uint64_t c = 0; c < n_chunks + kPackAhead; c++) {
    if (c < n_chunks) {
      const int k = (int)(c % kStages);
      const size_t n = chunk_size(c);
      const int parts = (int)(n / 4096 < 1 ? 1 : (n / 4096 > (size_t)kPackParts ? (size_t)kPackParts : n / 4096));  // >= 4 Ki rays per pack job
      {
        std::unique_lock<std::mutex> wl(st->mutex);
        if (c >= (uint64_t)kStages)  // the stage's previous chunk is done down to its last scatter job
          st->cv.wait(wl, [&] { return st->scat_submitted > c - kStages && st->pending[k] == 0; });
        st->pack_pending[k] = parts;
      }
      const uint8_t* src = base + c * chunk * stride;
      uint8_t* dst = static_cast<uint8_t*>(st->h_pk[k]);
      for (int part = 0; part < parts; part++) {
        const size_t b = n * part / parts, e = n * (part + 1) / parts;
        HostPool::get().submit([=] {
          pack_rays(src, stride, dst, b, e);
          std::lock_guard<std::mutex> wl(st->mutex);
          if (--st->pack_pending[k] == 0) st->cv.notify_all();
        });
      }
    }
    if (c >= (uint64_t)kPackAhead) {
      const uint64_t d = c - kPackAhead;
      const int k = (int)(d % kStages);
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
      for (int k = 0; k < kStages; k++)
        if (st->pending[k]) return false;
      return true;
    });
    if (st->worker_error != cudaSuccess) return cuda_fail(st->worker_error, "result scatter");
  }
  return ECB_SUCCESS;
}

// Shared by intersect (elem = 80) and occluded (elem = 48).
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
  ecbError err = ensure_resident(s, -1);
  if (err != ECB_SUCCESS) return err;
  std::lock_guard<std::mutex> lock(s->mutex);
  CU(cudaSetDevice(s->device));
  err = ensure_staging(s);
  if (err != ECB_SUCCESS) return err;
  Staging* st = s->staging;
  uint8_t* base = static_cast<uint8_t*>(rays);
  const size_t chunk = pipeline_chunk();
  const uint64_t n_chunks = (M + chunk - 1) / chunk;
  const bool pinned = is_pinned(rays);
  const bool compact = use_compact_results();
  const int extra = s->instanced ? kExtraInst : s->has_triangles() ? kExtraNg : kExtraNone;
  // auto: pageable caller memory has to go through host threads anyway -> packed.  Pinned caller memory can be DMA'd as it
  // is, which costs no host cycles: packed wins when the process has a full pool (12 threads: 650 vs 510 Mrays/s on the
  // 16-core single-GPU host), compact when the host's cores are shared between ranks (2 ranks on 24 cores, 9 threads each:
  // 743 packed vs 855 compact)
  if (compact && (host_path() == kPathPacked || (host_path() == kPathAuto && (!pinned || HostPool::get().size() >= 12)))) return run_host_packed(s, st, occluded, exact, instID, base, M, stride);
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
    L.queue = next_queue(s);
    g_launches.fetch_add(1);
    if (s->instanced) return ecb_launch_instanced(s->inst_dev, s->device, L, occluded, st->stream[k]);
    if (s->has_triangles()) {  // AccelN order: triangle accel first, in place, then the compressed accel on the same rays
      cudaError_t te = ecb_launch_triangles(s->tri_dev, s->device, L, occluded, st->stream[k]);
      if (te != cudaSuccess) return te;
      L.keep_results = compact ? 1u : 0u;
      g_launches.fetch_add(1);
    }
    if (exact) return ecb_launch_intersect(s->dev, s->device, L, st->stream[k]);
    return occluded ? ecb_launch_occluded(s->dev, s->device, L, st->stream[k]) : ecb_launch_intersect(s->dev, s->device, L, st->stream[k]);
  };

  if (!pinned) {
    for (int i = 0; i < kStages; i++)
      if (!st->h_pin[i]) CU(cudaMallocHost(&st->h_pin[i], kChunkRaysMax * kRayHitBytes));
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
        for (int k = 0; k < kStages; k++)
          if (st->pending[k]) return false;
        return true;
      });
    }
  } drain{st};

  for (uint64_t c = 0; c < n_chunks; c++) {
    const int k = (int)(c % kStages);
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
      if (!compact && c >= (uint64_t)kStages) unpack(c - kStages, k);
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
          const int device = s->device;
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
    for (int k = 0; k < kStages; k++) wait_stage(k);
    if (st->worker_error != cudaSuccess) return cuda_fail(st->worker_error, "result scatter");
    return ECB_SUCCESS;
  }
  for (int k = 0; k < kStages; k++) CU(cudaStreamSynchronize(st->stream[k]));
  if (compact) {  // small call: scatter inline
    for (uint64_t c = 0; c < n_chunks; c++) {
      const size_t n = chunk_size(c);
      scatter_results(base + c * chunk * stride, stride, static_cast<const uint8_t*>(st->h_cmp[c % kStages]), n, 0, n, any_hit, instID, extra);
    }
  } else if (!pinned) {
    const uint64_t tail0 = n_chunks > (uint64_t)kStages ? n_chunks - kStages : 0;
    for (uint64_t c = tail0; c < n_chunks; c++) unpack(c, (int)(c % kStages));
  }
  return ECB_SUCCESS;
}

ecbError run_device(ecbScene* s, bool occluded, uint32_t instID, void* d_rays, uint64_t M, size_t stride, uint32_t* d_block_out,
                    void* stream, unsigned long long* d_counters = nullptr, bool exact = false) {
  if (!s || (!d_rays && M)) return ECB_ERROR_INVALID_ARGUMENT;
  const size_t elem = occluded ? kRayBytes : kRayHitBytes;
  if (stride < elem || (stride & 15) || (reinterpret_cast<uintptr_t>(d_rays) & 15)) {
    g_last_error = "device rays must be 16-byte aligned with a byteStride that is a multiple of 16";
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  ecbError err = ensure_resident(s, -1);
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
  L.queue = next_queue(s);
  if (s->instanced && (d_block_out || d_counters)) {
    g_last_error = "block output / work counters are not available on instancing scenes";
    return ECB_ERROR_INVALID_OPERATION;
  }
  if (exact && s->instanced) {
    g_last_error = "exact any-hit is not available on instancing scenes";
    return ECB_ERROR_INVALID_OPERATION;
  }
  if (!s->instanced && s->has_triangles()) {
    cudaError_t te = ecb_launch_triangles(s->tri_dev, s->device, L, occluded, q);
    if (te != cudaSuccess) return cuda_fail(te, "kernel launch");
    g_launches.fetch_add(1);
  }
  cudaError_t le = s->instanced ? ecb_launch_instanced(s->inst_dev, s->device, L, occluded, q)
                   : exact      ? ecb_launch_intersect(s->dev, s->device, L, q)
                   : occluded   ? ecb_launch_occluded(s->dev, s->device, L, q)
                                : ecb_launch_intersect(s->dev, s->device, L, q);
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
    t.root = tri_root;
    e = ecb_import_triangles(&t, *scene_out);
    if (e != ECB_SUCCESS) {
      ecbSceneRelease(*scene_out);
      *scene_out = nullptr;
    }
  }
  return e;
}

ecbError ecbSceneImportImageFile(const char* path, ecbScene** scene_out) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) {
    g_last_error = std::string("cannot open ") + (path ? path : "(null)");
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  const size_t n = (size_t)f.tellg();
  std::vector<char> buf(n);
  f.seekg(0);
  f.read(buf.data(), (std::streamsize)n);
  return ecbSceneImportImage(buf.data(), n, scene_out);
}

ecbError ecbSceneSave(const ecbScene* s, const char* path) {
  if (!s || !path) return ECB_ERROR_INVALID_ARGUMENT;
  if (s->instanced) {
    g_last_error = "instancing scenes are not serialisable (save the object scenes)";
    return ECB_ERROR_INVALID_OPERATION;
  }
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
  std::ofstream f(path, std::ios::binary);
  if (!f) {
    g_last_error = std::string("cannot write ") + path;
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  f.write(reinterpret_cast<const char*>(&h), sizeof(h));
  f.write(reinterpret_cast<const char*>(s->top.data()), (std::streamsize)h.top_bytes);
  f.write(reinterpret_cast<const char*>(s->blocks.data()), (std::streamsize)h.block_bytes);
  if (s->has_triangles()) {
    EcbBlobTriTrailer t;
    memset(&t, 0, sizeof(t));
    memcpy(t.magic, "ECB200T1", 8);
    t.n_nodes = (uint32_t)s->tri_top.size();
    t.n_blocks = (uint32_t)s->tri_blocks.size();
    t.root = s->tri_root;
    t.depth = s->tri_depth;
    f.write(reinterpret_cast<const char*>(&t), sizeof(t));
    f.write(reinterpret_cast<const char*>(s->tri_top.data()), (std::streamsize)(s->tri_top.size() * sizeof(EcbTopNode)));
    f.write(reinterpret_cast<const char*>(s->tri_blocks.data()), (std::streamsize)(s->tri_blocks.size() * sizeof(EcbTriBlock)));
  }
  return f ? ECB_SUCCESS : ECB_ERROR_UNKNOWN;
}

ecbError ecbSceneLoad(const char* path, ecbScene** out) {
  if (!path || !out) return ECB_ERROR_INVALID_ARGUMENT;
  std::ifstream f(path, std::ios::binary);
  EcbBlobHeader h;
  if (!f || !f.read(reinterpret_cast<char*>(&h), sizeof(h)) || memcmp(h.magic, "ECB200S1", 8) != 0 || h.version != 1) {
    g_last_error = std::string("not an ECB200 scene blob: ") + path;
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  if (h.top_bytes != (uint64_t)h.n_top * sizeof(EcbTopNode) || h.block_bytes != (uint64_t)h.n_blocks * h.block_stride ||
      h.comp_level < 1 || h.comp_level > ECB_MAX_C || h.leaf_kind > ECB_LEAF_BOX || 3 * h.top_depth + 4 > ECB_MAX_TOP_STACK) {
    g_last_error = "corrupt scene blob header";
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  ecbScene* s = new (std::nothrow) ecbScene;
  if (!s) return ECB_ERROR_OUT_OF_MEMORY;
  s->top.resize(h.n_top);
  s->blocks.resize(h.block_bytes);
  f.read(reinterpret_cast<char*>(s->top.data()), (std::streamsize)h.top_bytes);
  f.read(reinterpret_cast<char*>(s->blocks.data()), (std::streamsize)h.block_bytes);
  if (!f) {
    delete s;
    g_last_error = "truncated scene blob";
    return ECB_ERROR_INVALID_ARGUMENT;
  }
  // every child reference must stay inside the arrays: the kernels trust the blob
  for (const EcbTopNode& n : s->top)
    for (int i = 0; i < 4; i++) {
      const uint32_t c = n.child[i];
      const bool ok = c == ECB_REF_EMPTY || ((c & ECB_REF_LEAF) ? (c & ~ECB_REF_LEAF) < h.n_blocks : c < h.n_top);
      if (!ok) {
        delete s;
        g_last_error = "scene blob has an out-of-range child reference";
        return ECB_ERROR_INVALID_ARGUMENT;
      }
    }
  s->root = h.root;
  s->block_stride = h.block_stride;
  s->leaf_offset = h.leaf_offset;
  s->elems = h.elems;
  s->grid_width = h.grid_width;
  s->rcp_edges = h.rcp_edges;
  ecbSceneInfo& in = s->info;
  in.leaf_kind = h.leaf_kind;
  in.subdiv_level = h.subdiv_level;
  in.comp_level = h.comp_level;
  in.n_top_nodes = h.n_top;
  in.n_blocks = h.n_blocks;
  in.top_depth = h.top_depth;
  in.block_stride = h.block_stride;
  in.top_node_bytes = sizeof(EcbTopNode);
  in.device_bytes = h.top_bytes + h.block_bytes;
  in.reference_bytes = h.reference_bytes;
  in.micro_triangles = 2ull * (1ull << (2 * h.comp_level)) * h.n_blocks;
  memcpy(in.bounds, h.bounds, sizeof(in.bounds));
  in.device = -1;
  EcbBlobTriTrailer t;
  if (f.read(reinterpret_cast<char*>(&t), sizeof(t))) {  // optional: triangle accel of a mixed scene
    bool ok = memcmp(t.magic, "ECB200T1", 8) == 0 && 3 * t.depth + 4 <= ECB_MAX_TOP_STACK && t.n_blocks < (1u << 27);
    if (ok) {
      s->tri_top.resize(t.n_nodes);
      s->tri_blocks.resize(t.n_blocks);
      f.read(reinterpret_cast<char*>(s->tri_top.data()), (std::streamsize)(t.n_nodes * sizeof(EcbTopNode)));
      f.read(reinterpret_cast<char*>(s->tri_blocks.data()), (std::streamsize)(t.n_blocks * sizeof(EcbTriBlock)));
      ok = (bool)f;
    }
    auto ref_ok = [&](uint32_t c) {
      if (c == ECB_REF_EMPTY) return true;
      if (!(c & ECB_REF_LEAF)) return c < t.n_nodes;
      const uint32_t first = (c & ~ECB_REF_LEAF) >> 3, num = c & 7u;
      return (uint64_t)first + num <= t.n_blocks;
    };
    if (ok) {
      ok = ref_ok(t.root) && t.root != ECB_REF_EMPTY;
      for (const EcbTopNode& n : s->tri_top)
        for (int i = 0; i < 4; i++) ok = ok && ref_ok(n.child[i]);
    }
    if (!ok) {
      delete s;
      g_last_error = "corrupt triangle section in scene blob";
      return ECB_ERROR_INVALID_ARGUMENT;
    }
    s->tri_root = t.root;
    s->tri_depth = t.depth;
    in.device_bytes += t.n_nodes * sizeof(EcbTopNode) + (uint64_t)t.n_blocks * sizeof(EcbTriBlock);
  }
  *out = s;
  return ECB_SUCCESS;
}

ecbError ecbSceneUpload(ecbScene* s, int device) {
  if (!s) return ECB_ERROR_INVALID_ARGUMENT;
  return ensure_resident(s, device);
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
  if (s->device >= 0) {
    cudaSetDevice(s->device);
    delete s->staging;
    if (s->d_top) cudaFree(s->d_top);
    if (s->d_blocks) cudaFree(s->d_blocks);
    if (s->d_queues) cudaFree(s->d_queues);
    if (s->d_items) cudaFree(s->d_items);
    if (s->d_instances) cudaFree(s->d_instances);
    if (s->d_objects) cudaFree(s->d_objects);
    if (s->d_tri_top) cudaFree(s->d_tri_top);
    if (s->d_tri_blocks) cudaFree(s->d_tri_blocks);
  }
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
  if (s->device >= 0 || bytes != s->blocks.size()) {
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
