// Launch accounting: per-class launch counters (always on) and optional CUDA-event timing of every launch.
#include <mutex>
#include <vector>

#include "../../include/pvcr_b200.h"
#include "common.cuh"
#include "persist.cuh"
#include "kernels.cuh"

namespace pvcr {

struct EventPair { cudaEvent_t a, b; int cls; double work; };
static std::mutex g_mu;
static unsigned long long g_launches[KC_COUNT];
static double g_work[KC_COUNT];
static int g_timing = 0;        // 1: time eager launches; 2: also launches captured into a CUDA graph (external event nodes)
static std::vector<EventPair> g_pending;
static std::vector<EventPair> g_pool;
static const char* const g_names[KC_COUNT] = {"gemm_tcgen05", "operand_staging", "rnn_gates", "attention",
                                              "loss", "gru_persistent_fwd", "gru_persistent_bwd",
                                              "decoder_persistent_fwd", "decoder_persistent_bwd", "misc"};

LaunchScope::LaunchScope(int c, cudaStream_t s, double work) : cls(c), st(s), rec(nullptr), ext(false) {
  std::lock_guard<std::mutex> g(g_mu);
  g_launches[c]++;
  g_work[c] += work;
  if (!g_timing) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cs) != cudaSuccess) return;
  const bool capturing = cs != cudaStreamCaptureStatusNone;
  if (capturing && g_timing < 2) return;
  EventPair ep;
  if (!g_pool.empty()) { ep = g_pool.back(); g_pool.pop_back(); }
  else if (cudaEventCreate(&ep.a) != cudaSuccess || cudaEventCreate(&ep.b) != cudaSuccess) return;
  ep.cls = c;
  ep.work = work;
  cudaEventRecordWithFlags(ep.a, s, capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
  ext = capturing;
  g_pending.push_back(ep);
  rec = reinterpret_cast<void*>(g_pending.size());     // 1-based index
}
LaunchScope::~LaunchScope() {
  if (!rec) return;
  std::lock_guard<std::mutex> g(g_mu);
  const size_t i = reinterpret_cast<size_t>(rec) - 1;
  if (i < g_pending.size()) cudaEventRecordWithFlags(g_pending[i].b, st, ext ? cudaEventRecordExternal : cudaEventRecordDefault);
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

static const unsigned long long* g_seed_step = nullptr;
const unsigned long long* seed_step_ptr() { return g_seed_step; }
static long long* g_phase_buf = nullptr;
static bool g_phase_on = false;
constexpr int PHASE_STEPS = 256;
long long* debug_phase_buffer() { return g_phase_on ? g_phase_buf : nullptr; }

}  // namespace pvcr

using namespace pvcr;

extern "C" {

int pvcr_prof_num_classes(void) { return KC_COUNT; }
const char* pvcr_prof_class_name(int cls) { return (cls >= 0 && cls < KC_COUNT) ? g_names[cls] : ""; }
void pvcr_prof_enable(int on) {
  std::lock_guard<std::mutex> g(g_mu);
  g_timing = on < 0 ? 0 : (on > 2 ? 2 : on);
}
void pvcr_prof_reset(void) {
  std::lock_guard<std::mutex> g(g_mu);
  for (int i = 0; i < KC_COUNT; ++i) { g_launches[i] = 0; g_work[i] = 0.0; }
  for (auto& e : g_pending) g_pool.push_back(e);
  g_pending.clear();
}
// launches[cls], ms[cls] (sum of event-timed durations since the last reset; 0 when timing is off), work[cls]
int pvcr_prof_read(uint64_t* launches, double* ms, double* work) {
  std::lock_guard<std::mutex> g(g_mu);
  for (int i = 0; i < KC_COUNT; ++i) { launches[i] = g_launches[i]; ms[i] = 0.0; work[i] = g_work[i]; }
  for (auto& e : g_pending) {
    if (cudaEventSynchronize(e.b) != cudaSuccess) return PVCR_ERR_CUDA;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, e.a, e.b) != cudaSuccess) return PVCR_ERR_CUDA;
    ms[e.cls] += t;
  }
  return PVCR_OK;
}

// Timeline of the event-timed launches since the last reset: for launch i (in host enqueue order) cls[i] and its
// start / end in milliseconds relative to the first launch's start.  Returns the number of launches written.
int pvcr_prof_timeline(int* cls, float* t0, float* t1, int cap) {
  std::lock_guard<std::mutex> g(g_mu);
  int n = 0;
  for (auto& e : g_pending) {
    if (n >= cap) break;
    if (cudaEventSynchronize(e.b) != cudaSuccess) return PVCR_ERR_CUDA;
    float a = 0.f, b = 0.f;
    if (cudaEventElapsedTime(&a, g_pending[0].a, e.a) != cudaSuccess) return PVCR_ERR_CUDA;
    if (cudaEventElapsedTime(&b, g_pending[0].a, e.b) != cudaSuccess) return PVCR_ERR_CUDA;
    cls[n] = e.cls; t0[n] = a; t1[n] = b;
    ++n;
  }
  return n;
}

// Every event-timed launch since the last reset, in host enqueue order: class, duration in ms, work (executed
// tensor-core FLOPs for GEMM launches, else 0).  Returns the number of entries written (synchronises).
int pvcr_prof_launch_list(int* cls, float* ms, double* work, int cap) {
  std::lock_guard<std::mutex> g(g_mu);
  int n = 0;
  for (auto& e : g_pending) {
    if (n >= cap) break;
    if (cudaEventSynchronize(e.b) != cudaSuccess) return PVCR_ERR_CUDA;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, e.a, e.b) != cudaSuccess) return PVCR_ERR_CUDA;
    cls[n] = e.cls; ms[n] = t; work[n] = e.work;
    ++n;
  }
  return n;
}

// Device counter mixed into every dropout / Gumbel seed at kernel run time (NULL = off).  A CUDA graph captures the
// pointer, not the value: incrementing the counter between (or inside) replays gives every replay fresh masks.
void pvcr_set_seed_step(const uint64_t* device_counter) {
  g_seed_step = reinterpret_cast<const unsigned long long*>(device_counter);
}

// Tuning aid: in-kernel phase timestamps of the persistent kernels (CTA 0).  enable allocates a small device
// buffer; read copies [steps][8] clock64 stamps of the LAST persistent launch to the host (synchronises).
int pvcr_debug_phase_timing(int on) {
  if (on && !g_phase_buf) {
    if (cudaMalloc(&g_phase_buf, sizeof(long long) * PHASE_STEPS * PHASE_SLOTS) != cudaSuccess) return PVCR_ERR_CUDA;
  }
  g_phase_on = on != 0;
  return PVCR_OK;
}
int pvcr_debug_phase_read(long long* out, int steps) {
  if (!g_phase_buf || steps > PHASE_STEPS) return PVCR_ERR_ARG;
  if (cudaDeviceSynchronize() != cudaSuccess) return PVCR_ERR_CUDA;
  if (cudaMemcpy(out, g_phase_buf, sizeof(long long) * steps * PHASE_SLOTS, cudaMemcpyDeviceToHost) != cudaSuccess)
    return PVCR_ERR_CUDA;
  return PVCR_OK;
}

}  // extern "C"
