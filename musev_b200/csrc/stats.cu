#include "stats.cuh"

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/musev_b200.h"

namespace mvb {

static std::atomic<long long> g_launches[KC_COUNT];
static std::atomic<bool> g_profiling{false};
struct EvPair { cudaEvent_t a, b; int cat; };
static std::vector<EvPair> g_events;
static std::mutex g_mu;

void stats_note_launch(int category, int n) { g_launches[category].fetch_add(n, std::memory_order_relaxed); }
bool stats_profiling() { return g_profiling.load(std::memory_order_relaxed); }

ProfScope::ProfScope(cudaStream_t stream, int category, int launches) : s(stream), cat(category) {
  stats_note_launch(category, launches);
  if (stats_profiling()) {
    cudaEventCreate(&e0);
    cudaEventRecord(e0, s);
  }
}
ProfScope::~ProfScope() {
  if (e0) {
    cudaEvent_t e1;
    cudaEventCreate(&e1);
    cudaEventRecord(e1, s);
    std::lock_guard<std::mutex> lk(g_mu);
    g_events.push_back({e0, e1, cat});
  }
}

}  // namespace mvb

extern "C" {

long long mvb_launch_count(int category) {
  if (category >= 0 && category < mvb::KC_COUNT) return mvb::g_launches[category].load();
  long long t = 0;
  for (int i = 0; i < mvb::KC_COUNT; ++i) t += mvb::g_launches[i].load();
  return t;
}

void mvb_profile_enable(int on) { mvb::g_profiling.store(on != 0); }

int mvb_profile_collect(double* ms_per_category, long long* scopes_per_category) {
  cudaError_t e = cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(mvb::g_mu);
  for (int i = 0; i < mvb::KC_COUNT; ++i) { ms_per_category[i] = 0.0; scopes_per_category[i] = 0; }
  for (auto& p : mvb::g_events) {
    float ms = 0.f;
    if (e == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      ms_per_category[p.cat] += ms;
      scopes_per_category[p.cat] += 1;
    }
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  mvb::g_events.clear();
  return e == cudaSuccess ? MVB_OK : MVB_ERR_CUDA;
}

}  // extern "C"
