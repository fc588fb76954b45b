// tests/mock_cuda/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.  A host-only stand-in for the CUDA runtime calls that
// embree_compressed_b200/csrc/ecb_api.cpp makes, so that the HOST pipeline of the product (chunking, packing, stage reuse,
// the thread pool, the completion thread, scatter) can be compiled with g++ -fsanitize=thread and run on a machine without
// a GPU (tests/test_host_pipeline_cpu.py).  It is found instead of the real header only through the -I of that test's
// build line; nothing in the product includes or links it.
//
// Model: "device memory" is host memory; every stream is a worker thread that executes the operations enqueued on it in
// order (copies, memsets, the fake kernels of mock_kernels.cpp, event records), so asynchronous copies really are
// asynchronous and ThreadSanitizer sees a staging buffer that is rewritten before its upload has run.
#pragma once
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

enum cudaError_t {
  cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInsufficientDriver = 35, cudaErrorNoDevice = 100,
  cudaErrorInvalidDevice = 101, cudaErrorNoKernelImageForDevice = 209, cudaErrorUnknown = 999
};
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; int device; };
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventBlockingSync = 1 };
struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

struct MockStream {
  std::mutex m;
  std::condition_variable cv;
  std::deque<std::function<void()>> q;
  bool running = false, quit = false;
  std::thread worker;
  MockStream() {
    worker = std::thread([this] {
      for (;;) {
        std::function<void()> op;
        {
          std::unique_lock<std::mutex> l(m);
          cv.wait(l, [this] { return quit || !q.empty(); });
          if (q.empty()) return;
          op = std::move(q.front());
          q.pop_front();
          running = true;
        }
        op();
        {
          std::lock_guard<std::mutex> l(m);
          running = false;
        }
        cv.notify_all();
      }
    });
  }
  ~MockStream() {
    {
      std::lock_guard<std::mutex> l(m);
      quit = true;
    }
    cv.notify_all();
    worker.join();
  }
  void enqueue(std::function<void()> op) {
    {
      std::lock_guard<std::mutex> l(m);
      q.push_back(std::move(op));
    }
    cv.notify_all();
  }
  void sync() {
    std::unique_lock<std::mutex> l(m);
    cv.wait(l, [this] { return q.empty() && !running; });
  }
};
typedef MockStream* cudaStream_t;

// the state of an event is shared with the record operations still sitting in a stream, so that destroying the handle
// while one of them completes is as harmless as it is with the real runtime
struct MockEventState {
  std::mutex m;
  std::condition_variable cv;
  uint64_t recorded = 0, completed = 0;
};
struct MockEvent {
  std::shared_ptr<MockEventState> state = std::make_shared<MockEventState>();
};
typedef MockEvent* cudaEvent_t;

// operations on the null stream run at once on the calling thread
inline void mock_enqueue(cudaStream_t s, std::function<void()> op) {
  if (s) s->enqueue(std::move(op));
  else op();
}

struct MockPinned {
  std::mutex m;
  std::map<uintptr_t, size_t> ranges;
};
inline MockPinned& mock_pinned() {
  static MockPinned* p = new MockPinned;
  return *p;
}

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "mock CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
// MOCK_CUDA_DEVICES (default 1) fake devices: "current device" is per thread, like the real runtime's
inline int mock_device_count() {
  static const int n = [] { const char* e = getenv("MOCK_CUDA_DEVICES"); const int v = e ? atoi(e) : 1; return v < 1 ? 1 : v; }();
  return n;
}
inline int& mock_current_device() { static thread_local int d = 0; return d; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = mock_device_count(); return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = mock_current_device(); return cudaSuccess; }
inline cudaError_t cudaSetDevice(int d) {
  if (d < 0 || d >= mock_device_count()) return cudaErrorInvalidDevice;
  mock_current_device() = d;
  return cudaSuccess;
}
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof(*h)); memcpy(h, &p, sizeof(p)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, &h, sizeof(*p)); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) {
  *p = malloc(n ? n : 1);
  if (!*p) return cudaErrorMemoryAllocation;
  memset(*p, 0xCD, n);  // nothing may rely on fresh device memory being zero
  return cudaSuccess;
}
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void** p, size_t n) {
  *p = aligned_alloc(64, (n + 63) & ~size_t(63));
  if (!*p) return cudaErrorMemoryAllocation;
  std::lock_guard<std::mutex> l(mock_pinned().m);
  mock_pinned().ranges[reinterpret_cast<uintptr_t>(*p)] = n;
  return cudaSuccess;
}
inline cudaError_t cudaFreeHost(void* p) {
  {
    std::lock_guard<std::mutex> l(mock_pinned().m);
    mock_pinned().ranges.erase(reinterpret_cast<uintptr_t>(p));
  }
  free(p);
  return cudaSuccess;
}
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  const uintptr_t v = reinterpret_cast<uintptr_t>(p);
  std::lock_guard<std::mutex> l(mock_pinned().m);
  auto it = mock_pinned().ranges.upper_bound(v);
  a->type = cudaMemoryTypeUnregistered;
  a->device = 0;
  if (it != mock_pinned().ranges.begin()) {
    --it;
    if (v < it->first + it->second) a->type = cudaMemoryTypeHost;
  }
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t q) {
  mock_enqueue(q, [=] { memcpy(d, s, n); });
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind,
                                     cudaStream_t q) {
  mock_enqueue(q, [=] {
    for (size_t r = 0; r < height; r++) memcpy(static_cast<uint8_t*>(d) + r * dpitch, static_cast<const uint8_t*>(s) + r * spitch, width);
  });
  return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t q) {
  mock_enqueue(q, [=] { memset(d, v, n); });
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new MockStream; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t s) { if (s) s->sync(); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new MockEvent; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t q) {
  std::shared_ptr<MockEventState> st = e->state;
  uint64_t ticket;
  {
    std::lock_guard<std::mutex> l(st->m);
    ticket = ++st->recorded;
  }
  mock_enqueue(q, [st, ticket] {
    std::lock_guard<std::mutex> l(st->m);
    if (st->completed < ticket) st->completed = ticket;
    st->cv.notify_all();
  });
  return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t e) {
  std::shared_ptr<MockEventState> st = e->state;
  std::unique_lock<std::mutex> l(st->m);
  const uint64_t want = st->recorded;
  st->cv.wait(l, [&] { return st->completed >= want; });
  return cudaSuccess;
}
