// Host-side worker threads shared by the front-end and the engine's layout pass.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <thread>
#include <vector>

// SGB_FRONTEND_THREADS, default: the hardware's, at most 16.  Calls are independent -- each owns its RNG -- so
// anything per call can run in parallel; only the order in which results enter the batch is kept serial.
inline int sgb_host_threads() {
  static const int n = [] {
    const char *e = getenv("SGB_FRONTEND_THREADS");
    int v = e ? atoi(e) : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    return std::max(1, v);
  }();
  return n;
}
inline std::atomic<int> &sgb_host_threads_override() { static std::atomic<int> v(0); return v; }
template <typename F>
inline void parallel_for(int n, F f, int grain = 8) {
  const int ov = sgb_host_threads_override().load();
  const int T = std::min(ov > 0 ? ov : sgb_host_threads(), (n + grain - 1) / grain);
  if (T <= 1) { for (int i = 0; i < n; i++) f(i); return; }
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  auto work = [&] { for (int i = next.fetch_add(grain); i < n; i = next.fetch_add(grain)) for (int j = i; j < std::min(n, i + grain); j++) f(j); };
  for (int t = 1; t < T; t++) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
}
