// tests/cpp/host_pipeline_check.cpp -- TEST INFRASTRUCTURE ONLY (built and run by tests/test_host_pipeline_cpu.py).
// The product's host layer (csrc/ecb_api.cpp + csrc/ecb_import.cpp, unmodified) compiled against tests/mock_cuda: streams
// are worker threads, kernels are tests/mock_cuda/mock_kernels.cpp, results are fake_trace.h's functions of the ray.
// What is checked, for the transfer strategy named by ECB200_HOST_PATH: every record of every call holds exactly what the
// API contract says (hits written, misses / inactive rays / padding / input fields untouched, any-hit tfar = -inf), for
// plain and mixed scenes, pinned and pageable caller memory, dense and padded strides, ray counts from 1 to several
// pipeline rounds, and concurrent callers -- under ThreadSanitizer, which turns a staging buffer reused too early, a
// scatter job outliving its call or an unprotected counter into a report.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ecb200.h"
#include "../mock_cuda/fake_trace.h"

static std::atomic<int> g_failures{0};
#define CHECK(cond, ...)                                  \
  do {                                                    \
    if (!(cond)) {                                        \
      if (g_failures.fetch_add(1) < 20) { fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } \
    }                                                     \
  } while (0)

static uint32_t rng(uint64_t& s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
static float frand(uint64_t& s, float lo, float hi) { return lo + (hi - lo) * (float)(rng(s) & 0xFFFFFF) / 16777216.0f; }

// one record of `width` floats: RTCRayHit (20 floats) or RTCRay (12 floats) + padding
static void make_record(float* r, size_t width, bool rayhit, uint64_t& s) {
  for (int k = 0; k < 3; k++) r[k] = frand(s, -5.f, 5.f);
  r[3] = 0.f;
  for (int k = 0; k < 3; k++) r[4 + k] = frand(s, -1.f, 1.f);
  r[7] = frand(s, 0.f, 1.f);  // time: never read, never written
  const uint32_t kind = rng(s) % 10u;
  r[8] = kind < 6 ? __builtin_inff() : kind == 6 ? 0.8f : kind == 7 ? 0.6f : kind == 8 ? -1.0f : 1.0f;
  if (kind == 9) r[3] = 2.0f;  // tnear > tfar: inactive stream ray
  uint32_t ids[3] = {rng(s), rng(s), rng(s)};  // mask, id, flags: never read, never written
  memcpy(r + 9, ids, 12);
  if (rayhit) {
    for (int k = 12; k < 17; k++) r[k] = frand(s, 10.f, 20.f);  // stale Ng, u, v
    const uint32_t tail[3] = {0xABCD0000u + (rng(s) & 0xFFFFu), 0xFFFFFFFFu, 0x55555555u};  // primID, geomID = invalid, instID
    memcpy(r + 17, tail, 12);
  }
  for (size_t k = rayhit ? 20 : 12; k < width; k++) r[k] = 7.0f;
}

static void expect_closest(float* r, bool mixed, uint32_t instID) {
  float tfar = r[8];
  FakeHit h, best;
  bool found = false;
  if (mixed && fake_closest(1, r, r[3], r + 4, tfar, h)) { best = h; tfar = h.t; found = true; }
  if (fake_closest(0, r, r[3], r + 4, tfar, h)) { best = h; tfar = h.t; found = true; }
  if (!found) return;
  r[8] = best.t;
  memcpy(r + 12, best.ng, 12);
  r[15] = best.u;
  r[16] = best.v;
  const uint32_t tail[3] = {best.primID, best.geomID, instID};
  memcpy(r + 17, tail, 12);
}
static void expect_any(float* r, bool mixed) {
  if ((mixed && fake_any(1, r, r[3], r + 4, r[8])) || fake_any(0, r, r[3], r + 4, r[8])) r[8] = -__builtin_inff();
}

struct Buffer {
  float* p = nullptr;
  bool pinned = false;
  Buffer(size_t floats, bool pin) : pinned(pin) {
    if (pin) { void* v = nullptr; if (ecbHostAlloc(&v, floats * 4) != ECB_SUCCESS) abort(); p = (float*)v; }
    else p = (float*)malloc(floats * 4 + 4);
  }
  ~Buffer() { if (pinned) ecbHostFree(p); else free(p); }
};

static void run_case(ecbScene* scene, bool mixed, bool pinned, size_t width, uint32_t M, int query, uint64_t seed) {
  // query: 0 closest hit, 1 any hit, 2 exact any hit
  const bool rayhit = query == 0;
  Buffer got(width * (size_t)M + 1, pinned);
  std::vector<float> want(width * (size_t)M);
  uint64_t s = seed;
  for (uint32_t i = 0; i < M; i++) make_record(want.data() + i * width, width, rayhit, s);
  memcpy(got.p, want.data(), want.size() * 4);
  const uint32_t instID = 0x1234u + (uint32_t)seed;
  for (uint32_t i = 0; i < M; i++) {
    if (query == 0) expect_closest(want.data() + i * width, mixed, instID);
    else expect_any(want.data() + i * width, mixed);
  }
  ecbError e = query == 0   ? ecbIntersect1M(scene, instID, got.p, M, width * 4)
               : query == 1 ? ecbOccluded1M(scene, instID, got.p, M, width * 4)
                            : ecbOccludedExact1M(scene, instID, got.p, M, width * 4);
  CHECK(e == ECB_SUCCESS, "call failed: %s %s", ecbErrorString(e), ecbLastErrorMessage());
  size_t bad = 0, first = 0;
  for (size_t k = 0; k < want.size(); k++)
    if (memcmp(&want[k], &got.p[k], 4) != 0) { if (!bad) first = k; bad++; }
  CHECK(bad == 0, "mixed=%d pinned=%d width=%zu M=%u query=%d: %zu words differ, first at record %zu word %zu", (int)mixed, (int)pinned,
        width, M, query, bad, first / width, first % width);
}

// SoA streams (ecbIntersectNp / ecbOccludedNp): the same records handed over as 20 (12) separate arrays must end up with
// exactly the words the AoS call leaves
static void run_case_np(ecbScene* scene, bool mixed, bool pinned, uint32_t M, bool occluded, uint64_t seed) {
  const size_t width = occluded ? 12 : 20;
  std::vector<float> want(width * (size_t)M);
  uint64_t s = seed;
  for (uint32_t i = 0; i < M; i++) make_record(want.data() + i * width, width, !occluded, s);
  std::vector<Buffer*> arrays;
  for (size_t f = 0; f < width; f++) {
    arrays.push_back(new Buffer((size_t)M + 1, pinned));
    for (uint32_t i = 0; i < M; i++) arrays[f]->p[i] = want[i * width + f];
  }
  const uint32_t instID = 0x777u + (uint32_t)seed;
  for (uint32_t i = 0; i < M; i++) {
    if (!occluded) expect_closest(want.data() + i * width, mixed, instID);
    else expect_any(want.data() + i * width, mixed);
  }
  ecbRayHitNp np;
  float** rf = &np.ray.org_x;
  for (int f = 0; f < 9; f++) rf[f] = arrays[f]->p;
  np.ray.mask = (uint32_t*)arrays[9]->p; np.ray.id = (uint32_t*)arrays[10]->p; np.ray.flags = (uint32_t*)arrays[11]->p;
  if (!occluded) {
    np.hit.Ng_x = arrays[12]->p; np.hit.Ng_y = arrays[13]->p; np.hit.Ng_z = arrays[14]->p; np.hit.u = arrays[15]->p; np.hit.v = arrays[16]->p;
    np.hit.primID = (uint32_t*)arrays[17]->p; np.hit.geomID = (uint32_t*)arrays[18]->p; np.hit.instID = (uint32_t*)arrays[19]->p;
  }
  ecbError e = occluded ? ecbOccludedNp(scene, instID, &np.ray, M) : ecbIntersectNp(scene, instID, &np, M);
  CHECK(e == ECB_SUCCESS, "Np call failed: %s %s", ecbErrorString(e), ecbLastErrorMessage());
  size_t bad = 0;
  for (size_t f = 0; f < width; f++)
    for (uint32_t i = 0; i < M; i++)
      if (memcmp(&arrays[f]->p[i], &want[i * width + f], 4) != 0) bad++;
  CHECK(bad == 0, "Np mixed=%d pinned=%d M=%u occluded=%d: %zu words differ", (int)mixed, (int)pinned, M, (int)occluded, bad);
  for (Buffer* b : arrays) delete b;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: host_pipeline_check plain.rimg mixed.rimg\n"); return 2; }
  ecbScene* scenes[2] = {nullptr, nullptr};
  for (int k = 0; k < 2; k++)
    if (ecbSceneImportImageFile(argv[1 + k], &scenes[k]) != ECB_SUCCESS) { fprintf(stderr, "import %s: %s\n", argv[1 + k], ecbLastErrorMessage()); return 3; }
  const uint32_t sizes[] = {1u, 100u, 8192u, 16384u, 16385u, 50000u, 150001u};
  uint64_t seed = 1;
  for (int mixed = 0; mixed < 2; mixed++)
    for (int pinned = 0; pinned < 2; pinned++)
      for (uint32_t M : sizes) {
        run_case(scenes[mixed], mixed, pinned, 20, M, 0, seed++);
        run_case(scenes[mixed], mixed, pinned, 24, M, 0, seed++);
        run_case(scenes[mixed], mixed, pinned, 12, M, 1, seed++);
        run_case(scenes[mixed], mixed, pinned, 16, M, 1, seed++);
        if (M <= 50000u) run_case(scenes[mixed], mixed, pinned, 12, M, 2, seed++);
      }
  for (int mixed = 0; mixed < 2; mixed++)
    for (int pinned = 0; pinned < 2; pinned++)
      for (uint32_t M : {1u, 5000u, 16385u, 70001u}) {
        run_case_np(scenes[mixed], mixed, pinned, M, false, seed++);
        run_case_np(scenes[mixed], mixed, pinned, M, true, seed++);
      }
  // M = 0 and argument checks never reach the pipeline
  CHECK(ecbIntersect1M(scenes[0], 0, nullptr, 0, 80) == ECB_SUCCESS, "M = 0");
  float dummy[20] = {0};
  CHECK(ecbIntersect1M(scenes[0], 0, dummy, 1, 78) == ECB_ERROR_INVALID_ARGUMENT, "bad stride accepted");
  // concurrent callers: two threads on one scene (each with its own staging context), two more on the other scene (sharing the pool)
  {
    std::vector<std::thread> th;
    for (int t = 0; t < 4; t++)
      th.emplace_back([&, t] {
        for (int rep = 0; rep < 3; rep++) run_case(scenes[t & 1], (t & 1) != 0, (t & 2) != 0, 20, 40000u + 1000u * t, 0, 1000 + 10 * t + rep);
      });
    for (auto& x : th) x.join();
  }
  // a scene resident on several (mock) devices: long calls fan out, one contiguous ray range per device, every range
  // through its own staging context; short calls stay on the primary device.  MOCK_CUDA_DEVICES > 1 enables this part.
  if (getenv("MOCK_CUDA_DEVICES") && atoi(getenv("MOCK_CUDA_DEVICES")) > 1) {
    const int n_dev = atoi(getenv("MOCK_CUDA_DEVICES"));
    std::vector<int> devs;
    for (int i = 0; i < n_dev; i++) devs.push_back(i);
    for (int k = 0; k < 2; k++) {
      CHECK(ecbSceneUploadMulti(scenes[k], devs.data(), (uint32_t)devs.size()) == ECB_SUCCESS, "upload multi: %s", ecbLastErrorMessage());
      int got[8] = {0};
      CHECK(ecbSceneResidentDevices(scenes[k], got, 8) == (uint32_t)n_dev, "resident device count");
    }
    for (int pinned = 0; pinned < 2; pinned++)
      for (uint32_t M : {1000u, 262144u, 300001u, 700003u}) {
        run_case(scenes[0], false, pinned, 20, M, 0, seed++);
        run_case(scenes[1], true, pinned, 24, M, 0, seed++);
        run_case(scenes[0], false, pinned, 12, M, 1, seed++);
        if (M >= 262144u) run_case_np(scenes[M & 1], (M & 1) != 0, pinned, M, false, seed++);
      }
    std::vector<std::thread> th;
    for (int t = 0; t < 3; t++)
      th.emplace_back([&, t] { run_case(scenes[0], false, (t & 1) != 0, 20, 400000u + 1000u * t, 0, 5000 + t); });
    for (auto& x : th) x.join();
    const int bad_dev[1] = {n_dev};
    CHECK(ecbSceneUploadMulti(scenes[0], bad_dev, 1) != ECB_SUCCESS, "upload to a device that does not exist accepted");
  }
  for (int k = 0; k < 2; k++) ecbSceneRelease(scenes[k]);
  const int f = g_failures.load();
  printf("host pipeline check: %s (path=%s)\n", f ? "FAILED" : "ok", getenv("ECB200_HOST_PATH") ? getenv("ECB200_HOST_PATH") : "default");
  return f ? 1 : 0;
}
