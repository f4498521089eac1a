// Side lane: a second (lower-priority) stream owned by the library.  The persistent recurrent kernels occupy 128 of
// the 148 SMs for hundreds of microseconds and several hoisted gradient GEMMs have far fewer output tiles than SMs;
// work that is off the step's critical path (weight gradients, bias column sums, embedding scatter, operand casts
// that only feed those) is forked onto this lane and joined back either at the end of the C-ABI call (mode 1) or at
// an explicit pvcr_side_join() (mode 2), so it runs in the shadow of the sweeps instead of between them.
// Fork / join are event record + stream wait pairs: capturable into a CUDA graph (they become graph edges).
#include <cstdlib>
#include <mutex>
#include <vector>

#include "../../include/pvcr_b200.h"
#include "host.h"

namespace pvcr {

namespace {
constexpr int NLANES = 3;
struct Lane {
  cudaStream_t s = nullptr;
  bool pending = false;        // work enqueued on the lane since the last join
};
std::mutex g_mu;
Lane g_lane[NLANES];
cudaEvent_t g_ev[32];
int g_next = 0, g_device = -1;
int g_mode = getenv("PVCR_SIDE_MODE") ? atoi(getenv("PVCR_SIDE_MODE")) : 1;     // tuning override of the default
int g_nlanes = getenv("PVCR_SIDE_LANES") ? atoi(getenv("PVCR_SIDE_LANES")) : NLANES;   // tuning aid: fold lanes together
thread_local int g_cta_cap = 0;

int ensure_lanes() {
  int dev = 0;
  PVCR_CUDA_CHECK(cudaGetDevice(&dev));
  if (g_lane[0].s && g_device == dev) return PVCR_OK;
  PVCR_REQUIRE(g_lane[0].s == nullptr, "side lane: one device per process (lanes live on device %d, current %d)",
               g_device, dev);
  int lo = 0, hi = 0;
  PVCR_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));     // lo = least priority (numerically greatest)
  for (auto& l : g_lane) PVCR_CUDA_CHECK(cudaStreamCreateWithPriority(&l.s, cudaStreamNonBlocking, lo));
  for (auto& e : g_ev) PVCR_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  g_device = dev;
  if (g_nlanes < 1) g_nlanes = 1;
  if (g_nlanes > NLANES) g_nlanes = NLANES;
  return PVCR_OK;
}
cudaEvent_t next_event() {
  cudaEvent_t e = g_ev[g_next];
  g_next = (g_next + 1) % 32;
  return e;
}
}  // namespace

int side_mode() { return g_mode; }
bool side_site(int bit) {
  static const int mask = getenv("PVCR_SIDE_MASK") ? atoi(getenv("PVCR_SIDE_MASK")) : 0xff;    // tuning aid
  return g_mode != 0 && ((mask >> bit) & 1);
}

// Lane `id` waits for everything enqueued on `main` so far.  *out = lane stream (== main when the lanes are off).
// Work on different lanes is unordered; work on one lane runs in enqueue order.
int side_fork(cudaStream_t main, cudaStream_t* out, int id) {
  *out = main;
  if (g_mode == 0) return PVCR_OK;
  std::lock_guard<std::mutex> g(g_mu);
  PVCR_TRY(ensure_lanes());
  Lane& l = g_lane[(id < 0 ? 0 : id) % g_nlanes];
  cudaEvent_t e = next_event();
  PVCR_CUDA_CHECK(cudaEventRecord(e, main));
  PVCR_CUDA_CHECK(cudaStreamWaitEvent(l.s, e, 0));
  l.pending = true;
  *out = l.s;
  return PVCR_OK;
}

// `main` waits for everything enqueued on lane `id` so far.
int side_join_lane(cudaStream_t main, int id) {
  std::lock_guard<std::mutex> g(g_mu);
  Lane& l = g_lane[(id < 0 ? 0 : id) % g_nlanes];
  if (!l.s || !l.pending) return PVCR_OK;
  cudaEvent_t e = next_event();
  PVCR_CUDA_CHECK(cudaEventRecord(e, l.s));
  PVCR_CUDA_CHECK(cudaStreamWaitEvent(main, e, 0));
  // the lane stays "pending": later work may be enqueued on it and a full join is still due
  return PVCR_OK;
}

// `main` waits for everything enqueued on every lane so far.
int side_join(cudaStream_t main) {
  std::lock_guard<std::mutex> g(g_mu);
  for (auto& l : g_lane) {
    if (!l.s || !l.pending) continue;
    cudaEvent_t e = next_event();
    PVCR_CUDA_CHECK(cudaEventRecord(e, l.s));
    PVCR_CUDA_CHECK(cudaStreamWaitEvent(main, e, 0));
    l.pending = false;
  }
  return PVCR_OK;
}

// Milestones: an event recorded in the MIDDLE of a lane's work by the C code that enqueues it (e.g. "the embedding
// gradient is final" while more weight-gradient GEMMs follow on the same lane); pvcr_side_wait_milestone makes a caller's
// stream wait for exactly that point instead of the lane's tail.
namespace {
cudaEvent_t g_milestone[4] = {nullptr, nullptr, nullptr, nullptr};
bool g_milestone_set[4] = {false, false, false, false};
}
int side_milestone(int id, cudaStream_t lane) {
  std::lock_guard<std::mutex> g(g_mu);
  if (id < 0 || id >= 4) return PVCR_OK;
  if (!g_milestone[id]) PVCR_CUDA_CHECK(cudaEventCreateWithFlags(&g_milestone[id], cudaEventDisableTiming));
  PVCR_CUDA_CHECK(cudaEventRecord(g_milestone[id], lane));
  g_milestone_set[id] = true;
  return PVCR_OK;
}

// One-shot notes between the C-ABI calls of one step, keyed by a workspace pointer: a call that already produced
// something a later call would otherwise compute (e.g. the forward pass staging the transposed weights of the backward
// sweep on a lane) leaves a note; the later call takes it.  A call that could have left a note but did not clears it.
namespace {
struct Note { const void* key; int tag; const void* what; };
std::vector<Note> g_notes;
}
// `what` names the data the note is about (e.g. the fp32 weight matrix that was staged): a note only matches a taker
// that asks about the same data, so a workspace address recycled by the caller's allocator cannot make a later call
// trust planes that were staged from other weights (or never finished: see pvcr_side_join, which drops them).
void side_note_put(const void* key, int tag, const void* what) {
  std::lock_guard<std::mutex> g(g_mu);
  for (size_t i = 0; i < g_notes.size();)        // at most one note per (workspace, tag): the newest
    if (g_notes[i].key == key && g_notes[i].tag == tag) g_notes.erase(g_notes.begin() + i); else ++i;
  if (g_notes.size() > 256) g_notes.clear();
  g_notes.push_back(Note{key, tag, what});
}
bool side_note_take(const void* key, int tag, const void* what) {
  std::lock_guard<std::mutex> g(g_mu);
  bool hit = false;
  for (size_t i = 0; i < g_notes.size();)
    if (g_notes[i].key == key && g_notes[i].tag == tag) {
      hit = hit || g_notes[i].what == what;      // a note about other data is stale: dropped, not honoured
      g_notes.erase(g_notes.begin() + i);
    } else {
      ++i;
    }
  return hit;
}

// End of a C-ABI call: mode 1 joins here, mode 2 leaves the lane running until pvcr_side_join().
int side_call_end(cudaStream_t main) { return g_mode == 2 ? PVCR_OK : side_join(main); }

static thread_local bool g_pdl = false;
bool pdl_enabled() { return g_pdl; }
PdlScope::PdlScope(bool on) : prev(g_pdl) { g_pdl = on; }
PdlScope::~PdlScope() { g_pdl = prev; }
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

// `stream` waits for what has been enqueued on ONE lane so far (the other lanes keep running un-joined): lets a
// data-parallel caller start the all-reduce of a gradient bucket as soon as the lane that produces it is done.
int pvcr_side_join_lane(void* stream, int lane) {
  if (lane < 0 || lane >= NLANES) { set_last_error("pvcr_side_join_lane: lane %d not in 0..%d", lane, NLANES - 1); return PVCR_ERR_ARG; }
  return side_join_lane(static_cast<cudaStream_t>(stream), lane);
}

// `stream` waits for milestone `id` of the latest call that recorded it (0: the embedding gradient of
// pvcr_s2vtatt_bwd_part(part = 1) is final).  No-op if that milestone was never recorded (lanes off).
int pvcr_side_wait_milestone(void* stream, int id) {
  std::lock_guard<std::mutex> g(g_mu);
  if (id < 0 || id >= 4) { set_last_error("pvcr_side_wait_milestone: id %d not in 0..3", id); return PVCR_ERR_ARG; }
  if (!g_milestone_set[id]) return PVCR_OK;
  PVCR_CUDA_CHECK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), g_milestone[id], 0));
  return PVCR_OK;
}

int pvcr_side_join(void* stream) {
  {   // the explicit join ends a step: notes about work staged for "the coming call" of that step do not outlive it
    std::lock_guard<std::mutex> g(g_mu);
    for (size_t i = 0; i < g_notes.size();)
      if (g_notes[i].tag == NOTE_VOCAB_WV) g_notes.erase(g_notes.begin() + i); else ++i;
  }
  return side_join(static_cast<cudaStream_t>(stream));
}

}  // extern "C"
