// Side lane: a second (lower-priority) stream owned by the library.  The persistent recurrent kernels occupy 128 of
// the 148 SMs for hundreds of microseconds and several hoisted gradient GEMMs have far fewer output tiles than SMs;
// work that is off the step's critical path (weight gradients, bias column sums, embedding scatter, operand casts
// that only feed those) is forked onto this lane and joined back either at the end of the C-ABI call (mode 1) or at
// an explicit pvcr_side_join() (mode 2), so it runs in the shadow of the sweeps instead of between them.
// Fork / join are event record + stream wait pairs: capturable into a CUDA graph (they become graph edges).
#include <cstdlib>
#include <mutex>

#include "../../include/pvcr_b200.h"
#include "host.h"

namespace pvcr {

namespace {
struct Lane {
  cudaStream_t s = nullptr;
  cudaEvent_t ev[16];
  int next = 0;
  bool pending = false;        // work enqueued on the lane since the last join
  int device = -1;
};
std::mutex g_mu;
Lane g_lane;
int g_mode = getenv("PVCR_SIDE_MODE") ? atoi(getenv("PVCR_SIDE_MODE")) : 1;     // tuning override of the default
thread_local int g_cta_cap = 0;

int ensure_lane() {
  int dev = 0;
  PVCR_CUDA_CHECK(cudaGetDevice(&dev));
  if (g_lane.s && g_lane.device == dev) return PVCR_OK;
  PVCR_REQUIRE(g_lane.s == nullptr, "side lane: one device per process (lane lives on device %d, current %d)",
               g_lane.device, dev);
  int lo = 0, hi = 0;
  PVCR_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));     // lo = least priority (numerically greatest)
  PVCR_CUDA_CHECK(cudaStreamCreateWithPriority(&g_lane.s, cudaStreamNonBlocking, lo));
  for (auto& e : g_lane.ev) PVCR_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  g_lane.device = dev;
  return PVCR_OK;
}
cudaEvent_t next_event() {
  cudaEvent_t e = g_lane.ev[g_lane.next];
  g_lane.next = (g_lane.next + 1) % 16;
  return e;
}
}  // namespace

int side_mode() { return g_mode; }
bool side_site(int bit) {
  static const int mask = getenv("PVCR_SIDE_MASK") ? atoi(getenv("PVCR_SIDE_MASK")) : 0xff;    // tuning aid
  return g_mode != 0 && ((mask >> bit) & 1);
}

// The lane waits for everything enqueued on `main` so far.  *out = lane stream (== main when the lane is off).
int side_fork(cudaStream_t main, cudaStream_t* out) {
  *out = main;
  if (g_mode == 0) return PVCR_OK;
  std::lock_guard<std::mutex> g(g_mu);
  PVCR_TRY(ensure_lane());
  cudaEvent_t e = next_event();
  PVCR_CUDA_CHECK(cudaEventRecord(e, main));
  PVCR_CUDA_CHECK(cudaStreamWaitEvent(g_lane.s, e, 0));
  g_lane.pending = true;
  *out = g_lane.s;
  return PVCR_OK;
}

// `main` waits for everything enqueued on the lane so far.
int side_join(cudaStream_t main) {
  std::lock_guard<std::mutex> g(g_mu);
  if (!g_lane.s || !g_lane.pending) return PVCR_OK;
  cudaEvent_t e = next_event();
  PVCR_CUDA_CHECK(cudaEventRecord(e, g_lane.s));
  PVCR_CUDA_CHECK(cudaStreamWaitEvent(main, e, 0));
  g_lane.pending = false;
  return PVCR_OK;
}

// End of a C-ABI call: mode 1 joins here, mode 2 leaves the lane running until pvcr_side_join().
int side_call_end(cudaStream_t main) { return g_mode == 2 ? PVCR_OK : side_join(main); }

int gemm_cta_cap() { return g_cta_cap; }
CtaCap::CtaCap(int cap) : prev(g_cta_cap) { g_cta_cap = cap; }
CtaCap::~CtaCap() { g_cta_cap = prev; }

}  // namespace pvcr

using namespace pvcr;

extern "C" {

// 0 = off (everything on the caller's stream), 1 = fork / join inside each call (default),
// 2 = joins deferred: the caller must call pvcr_side_join(stream) before it consumes gradients, ends a stream
//     capture or reuses the workspaces.  Returns the previous mode.
int pvcr_side_mode(int mode) {
  std::lock_guard<std::mutex> g(g_mu);
  const int prev = g_mode;
  if (mode >= 0 && mode <= 2) g_mode = mode;
  return prev;
}

int pvcr_side_join(void* stream) { return side_join(static_cast<cudaStream_t>(stream)); }

}  // extern "C"
