#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cstdint>
int main(int argc, char** argv) {
  const size_t M = 4150000;
  int nt = argc > 1 ? atoi(argv[1]) : 1;
  std::vector<uint8_t> rays(M * 80, 1);
  std::vector<float> rec(M * 4, 0.5f);
  std::vector<uint32_t> geom(M);
  for (size_t i = 0; i < M; i++) geom[i] = (i * 2654435761u >> 8) % 100 < 46 ? 0 : 0xFFFFFFFFu;
  for (int rep = 0; rep < 4; rep++) {
    // evict
    std::vector<uint8_t> junk(256 << 20, rep);
    volatile uint8_t sink = 0; for (size_t i = 0; i < junk.size(); i += 64) sink += junk[i];
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&, t, argc, argv]() {
      size_t b = M * t / nt, e = M * (t + 1) / nt;
      const int D = argc > 2 ? atoi(argv[2]) : 0;
      for (size_t i = b; i < e; i++) {
        if (D && i + D < e && geom[i + D] != 0xFFFFFFFFu) { __builtin_prefetch(rays.data() + (i + D) * 80 + 32, 1); __builtin_prefetch(rays.data() + (i + D) * 80 + 79, 1); }
        if (geom[i] == 0xFFFFFFFFu) continue;
        uint8_t* r = rays.data() + i * 80;
        const float* q = rec.data() + 4 * i;
        memcpy(r + 32, q, 4);
        float a[4] = {1.f, 0.f, 0.f, q[1]};
        memcpy(r + 48, a, 16);
        uint32_t c[4]; memcpy(c, q + 2, 8); c[2] = geom[i]; c[3] = 0xFFFFFFFFu;
        memcpy(r + 64, c, 16);
      }
    });
    for (auto& x : th) x.join();
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("threads %d prefetch %d: %.2f ms\n", nt, argc > 2 ? atoi(argv[2]) : 0, ms);
  }
}
