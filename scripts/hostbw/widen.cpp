// build: g++ -O3 -march=native -pthread scripts/hostbw/widen.cpp -o scripts/hostbw/widen
// host-side experiment: how fast can T threads widen 444 M int32 row indices to int64 (+1: 0-based -> 1-based)?
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
int main(int argc, char **argv) {
  const size_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 444194947ull;
  int32_t *src = (int32_t *)aligned_alloc(64, n * 4);
  int64_t *dst = (int64_t *)aligned_alloc(64, n * 8);
  for (size_t i = 0; i < n; i++) src[i] = (int32_t)i;
  for (size_t i = 0; i < n; i += 512) dst[i] = 0;   // touch pages
  printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
  for (int T : {1, 2, 4, 8, 16, 32}) {
    for (int rep = 0; rep < 2; rep++) {
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for (int t = 0; t < T; t++)
        th.emplace_back([=] {
          const size_t b = n * t / T, e = n * (t + 1) / T;
          for (size_t i = b; i < e; i++) dst[i] = (int64_t)src[i] + 1;
        });
      for (auto &x : th) x.join();
      double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (rep) printf("threads %2d: %.1f ms (%.1f GB/s of traffic)\n", T, ms, n * 12.0 / ms / 1e6);
    }
  }
  return (int)(dst[n / 2] & 1);
}
