#include "common.cuh"
#include <cstring>
#include <mutex>
#include <vector>

namespace dsb {
thread_local char g_err[512] = {0};
std::atomic<uint64_t> g_launches{0};
Tune g_tune;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ---- stage timers ----
struct Span { cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::mutex g_prof_mu;
static std::vector<Span> g_spans[ST_COUNT];
static std::vector<Span> g_pool;
static Span g_open[ST_COUNT];
static bool g_is_open[ST_COUNT] = {false};

void prof_begin(int stage, cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  Span s;
  if (!g_pool.empty()) {
    s = g_pool.back();
    g_pool.pop_back();
  } else {
    cudaEventCreate(&s.a);
    cudaEventCreate(&s.b);
  }
  cudaEventRecord(s.a, st);
  g_open[stage] = s;
  g_is_open[stage] = true;
}
void prof_end(int stage, cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_is_open[stage]) return;
  cudaEventRecord(g_open[stage].b, st);
  g_spans[stage].push_back(g_open[stage]);
  g_is_open[stage] = false;
}
}  // namespace dsb

extern "C" void dsb_profile_enable(int on) { dsb::g_prof_on = on != 0; }
extern "C" void dsb_profile_reset(void) {
  std::lock_guard<std::mutex> lk(dsb::g_prof_mu);
  for (int i = 0; i < dsb::ST_COUNT; ++i) {
    for (auto& s : dsb::g_spans[i]) dsb::g_pool.push_back(s);
    dsb::g_spans[i].clear();
  }
}
extern "C" int dsb_profile_read(int stage, double* total_ms, int* spans) {
  if (stage < 0 || stage >= dsb::ST_COUNT || !total_ms || !spans)
    return dsb::set_error(DSB_ERR_INVALID, "dsb_profile_read: bad argument");
  std::lock_guard<std::mutex> lk(dsb::g_prof_mu);
  double t = 0.0;
  for (auto& s : dsb::g_spans[stage]) {
    cudaError_t e = cudaEventSynchronize(s.b);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, s.a, s.b);
    if (e != cudaSuccess) return dsb::set_error(DSB_ERR_CUDA, "dsb_profile_read: %s", cudaGetErrorString(e));
    t += ms;
  }
  *total_ms = t;
  *spans = (int)dsb::g_spans[stage].size();
  return 0;
}

namespace {
struct Knob { const char* name; std::atomic<int>* v; int lo, hi; };
const Knob* knobs(int* n) {
  static const Knob k[] = {{"rnn_in_flight", &dsb::g_tune.rnn_in_flight, 1, 3},
                           {"rnn_max_slots", &dsb::g_tune.rnn_max_slots, 0, 1 << 20},
                           {"rnn_ksplit", &dsb::g_tune.rnn_ksplit, 0, 1},
                           {"rnn_ring_gsz", &dsb::g_tune.rnn_ring_gsz, 0, 4},
                           {"rnn_producers", &dsb::g_tune.rnn_producers, 1, 2},
                           {"rnn_pair", &dsb::g_tune.rnn_pair, 0, 1},
                           {"rnn_batch_minor", &dsb::g_tune.rnn_batch_minor, 0, 1},
                           {"rnn_pair_min_rows", &dsb::g_tune.rnn_pair_min_rows, 1, 1 << 20},
                           {"rnn_pair_in_flight", &dsb::g_tune.rnn_pair_in_flight, 1, 3}};
  *n = (int)(sizeof(k) / sizeof(k[0]));
  return k;
}
}  // namespace
extern "C" int dsb_tune_set(const char* key, int value) {
  int n = 0;
  const Knob* k = knobs(&n);
  for (int i = 0; key && i < n; ++i)
    if (strcmp(key, k[i].name) == 0) {
      if (value < k[i].lo || value > k[i].hi)
        return dsb::set_error(DSB_ERR_INVALID, "dsb_tune_set: %s = %d outside [%d, %d]", key, value, k[i].lo, k[i].hi);
      k[i].v->store(value);
      return 0;
    }
  return dsb::set_error(DSB_ERR_INVALID, "dsb_tune_set: unknown key '%s'", key ? key : "(null)");
}
extern "C" int dsb_tune_get(const char* key) {
  int n = 0;
  const Knob* k = knobs(&n);
  for (int i = 0; key && i < n; ++i)
    if (strcmp(key, k[i].name) == 0) return k[i].v->load();
  return -1;
}

extern "C" const char* dsb_last_error(void) { return dsb::g_err; }
extern "C" int dsb_abi_version(void) { return 2; }
extern "C" uint64_t dsb_kernel_launch_count(void) { return dsb::g_launches.load(); }
