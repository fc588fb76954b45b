// Developer tool: aggregate throughput of T threads calling ecbIntersect1M concurrently on ONE scene (the reference's usage:
// every tile task calls rtcIntersect1 / rtcIntersect1M on its own, with no lock -- kernels/common/rtcore.cpp:287-304).
//   g++ -O2 -std=c++17 -Iinclude tools/cpp/concurrent_callers.cpp -Lembree_compressed_b200 -lecb200 -Wl,-rpath,$PWD/embree_compressed_b200 -lpthread -o /tmp/concurrent_callers
//   /tmp/concurrent_callers scenes/_built/bomberman_s6c3_leaf.rimg [rays per call] [calls per thread]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "ecb200.h"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const uint32_t m = argc > 2 ? (uint32_t)atoi(argv[2]) : 4096;
  const int calls = argc > 3 ? atoi(argv[3]) : 500;
  ecbScene* scene = nullptr;
  if (ecbSceneImportImageFile(argv[1], &scene) != ECB_SUCCESS || ecbSceneUpload(scene, 0) != ECB_SUCCESS) {
    fprintf(stderr, "scene: %s\n", ecbLastErrorMessage());
    return 3;
  }
  // camera of build/bomberman.ecs, a 2048-pixel-wide strip of rays starting at a different row per thread
  auto make = [&](int t, std::vector<float>& r) {
    r.assign((size_t)m * 20, 0.f);
    for (uint32_t i = 0; i < m; i++) {
      float* q = r.data() + (size_t)i * 20;
      const float x = (float)((i + 37 * t) % 1920), y = (float)(300 + 60 * t + (i / 1920));
      const float d[3] = {x - 960.f - 0.4f * (y - 540.f), -(y - 540.f) * 0.9f - 300.f, -540.f + 0.3f * (x - 960.f)};
      const float n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      q[0] = 18.2124f; q[1] = 20.0575f; q[2] = 15.4688f; q[3] = 0.f;
      q[4] = d[0] / n; q[5] = d[1] / n; q[6] = d[2] / n;
      q[8] = INFINITY;
      uint32_t inv = 0xFFFFFFFFu;
      memcpy(q + 9, &inv, 4); memcpy(q + 17, &inv, 4); memcpy(q + 18, &inv, 4); memcpy(q + 19, &inv, 4);
    }
  };
  double base = 0;
  for (int T : {1, 2, 4, 8, 16}) {
    std::vector<std::vector<float>> src(T), buf(T);
    for (int t = 0; t < T; t++) { make(t, src[t]); buf[t] = src[t]; }
    std::vector<long> hits(T, 0);
    auto body = [&](int t) {
      for (int i = 0; i < calls; i++) {
        // tfar / hit fields back to "no hit yet" (8 words per ray, not the whole record)
        for (uint32_t k = 0; k < m; k++) { buf[t][(size_t)k * 20 + 8] = INFINITY; memcpy(&buf[t][(size_t)k * 20 + 18], &src[t][(size_t)k * 20 + 18], 4); }
        if (ecbIntersect1M(scene, 0xFFFFFFFFu, buf[t].data(), m, 80) != ECB_SUCCESS) { fprintf(stderr, "call failed: %s\n", ecbLastErrorMessage()); exit(4); }
      }
      for (uint32_t k = 0; k < m; k++) { uint32_t g; memcpy(&g, &buf[t][(size_t)k * 20 + 18], 4); hits[t] += g != 0xFFFFFFFFu; }
    };
    body(0);  // warm-up: staging contexts
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back(body, t);
    for (auto& x : th) x.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const double rate = (double)T * calls * m / dt / 1e6;
    if (T == 1) base = rate;
    printf("%2d caller thread(s) x %d calls of %u rays: %8.1f Mrays/s aggregate (x%.2f), %.1f us per call per thread, hits of thread 0: %ld\n", T, calls, m, rate,
           rate / base, dt / calls * 1e6, hits[0]);
    fflush(stdout);
  }
  ecbSceneRelease(scene);
  return 0;
}
