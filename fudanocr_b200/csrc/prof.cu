// Tiny in-library profiler: named CUDA-event scopes around kernel families, recorded on the launching
// stream (so bench.py can report per-kernel device time "live", without an external profiler), plus a
// global launch counter (bench.py's `gpu_launches`).  Disabled by default; when disabled a scope costs
// one branch.
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace {
struct Rec {
  const char* name;
  cudaEvent_t a, b;
};
std::mutex g_mu;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
int g_mode = 0;            // 0 off, 1 all scopes, 2 only scopes whose name contains g_focus
std::string g_focus;
long long g_launches = 0;
int g_depth = 0;           // open recorded scopes: a scope opened inside another one (the GEMM inside conv_wgrad_tc) is not recorded,
                           // so the per-family times add up to the step instead of counting the inner launches twice
void* const kNested = reinterpret_cast<void*>(~static_cast<uintptr_t>(0));

cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void focr_count_launch(int n) { g_launches += n; }

bool prof_begin(const char* name, cudaStream_t s, void** tok) {
  *tok = nullptr;
  if (g_mode == 0) return false;
  if (g_mode == 2 && std::string(name).find(g_focus) == std::string::npos) return false;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_depth > 0) {
    ++g_depth;
    *tok = kNested;
    return true;
  }
  ++g_depth;
  Rec r{name, get_event(), get_event()};
  cudaEventRecord(r.a, s);
  g_recs.push_back(r);
  *tok = reinterpret_cast<void*>(g_recs.size());
  return true;
}
void prof_end(void* tok, cudaStream_t s) {
  if (!tok) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_depth > 0) --g_depth;
  if (tok == kNested) return;
  const size_t i = reinterpret_cast<size_t>(tok) - 1;
  if (i < g_recs.size()) cudaEventRecord(g_recs[i].b, s);
}

extern "C" {
// mode 0 off, 1 all, 2 focus (substring match)
int focr_prof_enable(int mode, const char* focus) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_mode = mode;
  g_depth = 0;
  g_focus = focus ? focus : "";
  return 0;
}
// Synchronises, then writes up to `cap` lines "name count total_ms" into buf; clears the records.
int focr_prof_collect(char* buf, int cap) {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_mu);
  std::map<std::string, std::pair<int, double>> agg;
  for (auto& r : g_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.name];
      e.first += 1;
      e.second += ms;
    }
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  g_recs.clear();
  std::string out;
  for (auto& kv : agg) {
    char line[256];
    snprintf(line, sizeof(line), "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (cap > 0) {
    const size_t n = out.size() < (size_t)cap - 1 ? out.size() : (size_t)cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int)agg.size();
}
long long focr_launch_count(void) { return g_launches; }
}
