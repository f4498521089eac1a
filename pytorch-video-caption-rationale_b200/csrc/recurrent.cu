// Recurrent layers as sequences of (tcgen05 GEMM  h_{t-1} W_hh^T) + fused gate kernels, and the shared
// gradient GEMM helpers.  Reference: torch.nn.GRU call sites model/S2VTAttModel.py:88-93,142 and
// model/S2VTModel.py:84,107,122,129; torch.nn.LSTM model/RationaleNet.py:43.
#include <cstdlib>

#include "host.h"

namespace pvcr {

int grad_w(Arena& a, const float* dy, long long lddy, int R, int N, const float* x, long long ldx, int K,
           const long long* x_row_ids, const float* x_row_scale, float* dw, long long lddw, int accumulate,
           int nsplit, cudaStream_t st, Dropout x_drop) {
  static const bool mn_off = getenv("PVCR_NO_MN_WGRAD") != nullptr;
  if (nsplit == 1 && !mn_off) {
    // bf16 mode: MN-major tensor-core operands -- dy and x are only cast (gathered / scaled) row-major, the
    // contraction runs over their rows; no transposed copies
    // operands already staged in this call (or by the forward pass) are reused; new ones are staged once and kept
    StageCache* sc = a.measuring() ? nullptr : a.cache;
    const bool x_plain = !x_row_scale && x_drop.p <= 0.f;
    const void* xkey = x_row_ids ? (const void*)x_row_ids : (const void*)x;
    const long long xkey_ld = x_row_ids ? -1 : ldx;
    const Planes* dy_hit = sc ? sc->find(dy, lddy, R, N) : nullptr;
    const Planes* x_hit = sc ? sc->find(xkey, xkey_ld, R, K) : nullptr;
    Planes dyp, xp;
    int rc = PVCR_OK;
    if (sc && !dy_hit) {                       // persistent (cached) allocation: outside the mark / release window
      dyp = alloc_planes(a, R, N, 1);
      if (a.failed) { set_last_error("grad_w: workspace too small"); return PVCR_ERR_WORKSPACE; }
      PVCR_TRY(stage(dy, lddy, R, N, dyp, 0, nullptr, NO_DROPOUT, st));
      sc->put(dy, lddy, R, N, dyp);
      dy_hit = sc->find(dy, lddy, R, N);
    }
    if (sc && !x_hit && x_plain && !x_row_ids) {
      xp = alloc_planes(a, R, K, 1);
      if (a.failed) { set_last_error("grad_w: workspace too small"); return PVCR_ERR_WORKSPACE; }
      PVCR_TRY(cast_split(x, ldx, R, K, xp.ptr, xp.ld, xp.Kp, 1, 0, nullptr, NO_DROPOUT, st));
      sc->put(xkey, xkey_ld, R, K, xp);
      x_hit = sc->find(xkey, xkey_ld, R, K);
    }
    const size_t m = a.mark();
    if (!dy_hit) dyp = alloc_planes(a, R, N, 1);
    if (!x_hit) xp = alloc_planes(a, R, K, 1);
    if (!a.measuring()) {
      if (a.failed) { set_last_error("grad_w: workspace too small"); return PVCR_ERR_WORKSPACE; }
      if (!dy_hit) rc = stage(dy, lddy, R, N, dyp, 0, nullptr, NO_DROPOUT, st);
      if (rc == PVCR_OK && !x_hit) {
        if (x_row_ids) rc = gather_split(x, K, x_row_ids, R, xp.ptr, xp.ld, xp.Kp, 1, x_drop, st);
        else rc = cast_split(x, ldx, R, K, xp.ptr, xp.ld, xp.Kp, 1, 0, x_row_scale, x_drop, st);
      }
      if (rc == PVCR_OK)
        rc = gemm_mn_store((dy_hit ? *dy_hit : dyp).view_rows(0, R), (x_hit ? *x_hit : xp).view_rows(0, R), N, K, R, dw,
                           lddw, accumulate, st);
    }
    a.release(m);
    return rc;
  }
  const size_t m = a.mark();
  Planes dyT = alloc_planes(a, N, R, nsplit);
  Planes xT = alloc_planes(a, K, R, nsplit);
  int rc = PVCR_OK;
  if (!a.measuring()) {
    if (a.failed) { set_last_error("grad_w: workspace too small"); return PVCR_ERR_WORKSPACE; }
    rc = transpose_split(dy, lddy, R, N, dyT.ptr, dyT.ld, dyT.Kp, 0, 1, nsplit, 0, nullptr, nullptr, st, NO_DROPOUT);
    if (rc == PVCR_OK) rc = transpose_split(x, ldx, R, K, xT.ptr, xT.ld, xT.Kp, 0, 1, nsplit, 1, x_row_ids, x_row_scale, st, x_drop);
    if (rc == PVCR_OK) rc = gemm_planes(dyT.view(), xT.view(), N, K, (int)dyT.ld, dw, lddw, nullptr, accumulate, st);
  }
  a.release(m);
  return rc;
}

int grad_x(Arena& a, const float* dy, long long lddy, int R, int N, const Planes& wT, float* dx, long long lddx,
           int accumulate, cudaStream_t st) {
  StageCache* sc = (a.measuring() || wT.nsplit != 1) ? nullptr : a.cache;
  const Planes* hit = sc ? sc->find(dy, lddy, R, N) : nullptr;
  if (sc && !hit) {
    Planes keep = alloc_planes(a, R, N, 1);
    if (a.failed) { set_last_error("grad_x: workspace too small"); return PVCR_ERR_WORKSPACE; }
    PVCR_TRY(stage(dy, lddy, R, N, keep, 0, nullptr, NO_DROPOUT, st));
    sc->put(dy, lddy, R, N, keep);
    hit = sc->find(dy, lddy, R, N);
  }
  const size_t m = a.mark();
  Planes dya;
  if (!hit) dya = alloc_planes(a, R, N, wT.nsplit);
  int rc = PVCR_OK;
  if (!a.measuring()) {
    if (a.failed) { set_last_error("grad_x: workspace too small"); return PVCR_ERR_WORKSPACE; }
    if (!hit) rc = stage(dy, lddy, R, N, dya, 0, nullptr, NO_DROPOUT, st);
    const Planes& A = hit ? *hit : dya;
    if (rc == PVCR_OK) rc = gemm_planes(A.view(), wT.view(), R, wT.rows, (int)A.ld, dx, lddx, nullptr, accumulate, st);
  }
  a.release(m);
  return rc;
}

// bf16 mode: dx[R,Kout] (+)= dy[R,N] W[N,Kout] with W = the forward weight planes (MN-major B operand): no W^T copy.
int grad_x_fwdw(Arena& a, const float* dy, long long lddy, int R, int N, const Planes& w, int Kout, float* dx,
                long long lddx, int accumulate, cudaStream_t st) {
  StageCache* sc = a.measuring() ? nullptr : a.cache;
  const Planes* hit = sc ? sc->find(dy, lddy, R, N) : nullptr;
  if (sc && !hit) {
    Planes keep = alloc_planes(a, R, N, 1);
    if (a.failed) { set_last_error("grad_x: workspace too small"); return PVCR_ERR_WORKSPACE; }
    PVCR_TRY(stage(dy, lddy, R, N, keep, 0, nullptr, NO_DROPOUT, st));
    sc->put(dy, lddy, R, N, keep);
    hit = sc->find(dy, lddy, R, N);
  }
  const size_t m = a.mark();
  Planes dya;
  if (!hit) dya = alloc_planes(a, R, N, 1);
  int rc = PVCR_OK;
  if (!a.measuring()) {
    if (a.failed) { set_last_error("grad_x: workspace too small"); return PVCR_ERR_WORKSPACE; }
    if (!hit) rc = stage(dy, lddy, R, N, dya, 0, nullptr, NO_DROPOUT, st);
    const Planes& A = hit ? *hit : dya;
    if (rc == PVCR_OK) rc = gemm_kn_store(A.view(), w.view_rows(0, N), R, Kout, N, dx, lddx, accumulate, st);
  }
  a.release(m);
  return rc;
}

int gru_seq_fwd(const GruSeq& s, cudaStream_t st) {
  if (gru_persist_eligible(s) && gru_cluster_eligible(s)) return gru_cluster_fwd(s, st);
  if (gru_persist_eligible(s)) return gru_persist_fwd(s, st);
  const int H3 = 3 * s.H;
  for (int t = 0; t < s.T; ++t) {
    const bool has_prev = (t > 0) || (s.h0 != nullptr);
    if (has_prev) {
      OperandView a = (t > 0) ? OperandView{s.hp + (t - 1) * s.hp_ts, s.hp_ld, 0, s.B, 1}
                              : OperandView{s.h0_planes, s.h0_planes_ld, 0, s.B, 1};
      PVCR_TRY(gemm_planes(a, s.whh.view(), s.B, H3, (int)s.whh.ld, s.gh, H3, nullptr, 0, st));
    }
    GruFwdArgs g{};
    g.B = s.B; g.H = s.H;
    g.gi_a = s.gi_a ? s.gi_a + t * s.gi_a_ts : nullptr; g.gi_a_ld = s.gi_a_ld;
    if (s.gi_b && t >= s.gi_b_from) { g.gi_b = s.gi_b + (t - s.gi_b_from) * s.gi_b_ts; g.gi_b_ld = s.gi_b_ld; }
    g.gi_bias = s.gi_bias;
    g.gh = has_prev ? s.gh : nullptr; g.gh_ld = H3;
    g.b_hh = s.b_hh;
    if (t > 0) { g.h_prev = s.h + (t - 1) * s.h_ts; g.h_prev_ld = s.h_ld; }
    else { g.h_prev = s.h0; g.h_prev_ld = s.h0_ld; }
    g.h_out = s.h + t * s.h_ts; g.h_out_ld = s.h_ld;
    g.h_planes = s.hp ? s.hp + t * s.hp_ts : nullptr; g.h_planes_ld = s.hp_ld; g.Hp = s.Hp; g.nsplit = s.nsplit;
    const long long o = (long long)t * s.B * s.H;
    g.r = s.r + o; g.z = s.z + o; g.n = s.n + o; g.ghn = s.ghn + o;
    PVCR_TRY(gru_gate_fwd(g, st));
  }
  return PVCR_OK;
}

int gru_seq_bwd(const GruSeq& s, const GruSeqGrad& g, cudaStream_t st) {
  if (g.xch && gru_persist_eligible(s)) return gru_persist_bwd(s, g, st);
  for (int t = s.T - 1; t >= 0; --t) {
    const bool has_prev = (t > 0) || (s.h0 != nullptr);
    GruBwdArgs b{};
    b.B = s.B; b.H = s.H;
    b.dh_a = g.dh_carry; b.dh_a_ld = s.H;
    if (g.dh_ext) { b.dh_b = g.dh_ext + t * g.dh_ext_ts; b.dh_b_ld = g.dh_ext_ld; }
    const long long o = (long long)t * s.B * s.H;
    b.r = s.r + o; b.z = s.z + o; b.n = s.n + o; b.ghn = s.ghn + o;
    if (t > 0) { b.h_prev = s.h + (t - 1) * s.h_ts; b.h_prev_ld = s.h_ld; }
    else { b.h_prev = s.h0; b.h_prev_ld = s.h0_ld; }
    b.dgi = g.dgi + t * g.dgi_ts; b.dgi_ld = g.dgi_ld;
    b.dgh = g.dgh + t * g.dgh_ts; b.dgh_ld = g.dgh_ld;
    b.dgh_planes = has_prev ? g.dgh_a.ptr : nullptr; b.dgh_planes_ld = g.dgh_a.ld; b.dgh_Kp = g.dgh_a.Kp; b.dgh_col0 = 0;
    b.nsplit = s.nsplit;
    b.dh_direct = g.dh_carry; b.dh_direct_ld = s.H;
    PVCR_TRY(gru_gate_bwd(b, st));
    if (has_prev)
      PVCR_TRY(gemm_planes(g.dgh_a.view(), g.whhT.view(), s.B, s.H, (int)g.dgh_a.ld, g.dh_carry, s.H, nullptr, 1, st));
  }
  return PVCR_OK;
}

// ---- one GRU step as a self-contained op (the reference's `encode_step`, model/S2VTAttModel.py:63-78 and
// model/S2VTModel.py:57-72: `self.rnn(vid_feat.unsqueeze(0), rnn_state)`), differentiable in x, h_prev and the
// parameters.  This is the compatibility path for callers that drive the encoder frame by frame from Python
// (SpatialNet.py:120-138); the sequence entry points hoist the input projection and run a persistent sweep instead.
size_t linear_fwd_workspace(int M, int N, int K, int nsplit);
size_t linear_bwd_workspace(int M, int N, int K, int nsplit);
int linear_fwd(const float* x, long long ldx, const float* w, long long ldw, const float* bias, float* y, long long ldy,
               int M, int N, int K, int nsplit, void* ws, size_t ws_bytes, cudaStream_t st);
int linear_bwd(const float* dy, long long lddy, const float* x, long long ldx, const float* w, long long ldw, float* dx,
               long long lddx, float* dw, long long lddw, float* db, int M, int N, int K, int nsplit, int accumulate,
               void* ws, size_t ws_bytes, cudaStream_t st);

static size_t gru_step_scratch(int B, int V, int H, int nsplit) {
  size_t m = linear_fwd_workspace(B, 3 * H, V, nsplit);
  const size_t c[3] = {linear_fwd_workspace(B, 3 * H, H, nsplit), linear_bwd_workspace(B, 3 * H, V, nsplit),
                       linear_bwd_workspace(B, 3 * H, H, nsplit)};
  for (size_t v : c) if (v > m) m = v;
  return m;
}
size_t gru_step_workspace(int B, int V, int H, int nsplit) {
  Arena a(nullptr, 0);
  a.alloc<float>((size_t)4 * B * 3 * H);        // gi, gh (forward) / dgi, dgh (backward)
  a.alloc<float>((size_t)B * H);                // W_hh^T dgh before it is added to dh * z
  return a.off + gru_step_scratch(B, V, H, nsplit) + 1024;
}

// saved: [4][B,H] = r, z, n, W_hn h + b_hn (kept for the backward).  h_prev null = zeros.
int gru_step_fwd(const float* x, const float* h_prev, const float* w_ih, const float* w_hh, const float* b_ih,
                 const float* b_hh, int B, int V, int H, int nsplit, float* h_out, float* saved, void* ws,
                 size_t ws_bytes, cudaStream_t st) {
  PVCR_REQUIRE(B > 0 && V > 0 && H > 0 && nsplit >= 1 && nsplit <= 3, "gru_step_fwd: B=%d V=%d H=%d nsplit=%d", B, V, H, nsplit);
  Arena a(ws, ws_bytes);
  float* gi = a.alloc<float>((size_t)B * 3 * H);
  float* gh = a.alloc<float>((size_t)B * 3 * H);
  a.alloc<float>((size_t)2 * B * 3 * H);
  a.alloc<float>((size_t)B * H);
  const size_t need = gru_step_scratch(B, V, H, nsplit);
  char* scratch = a.alloc<char>(need);
  if (a.failed) { set_last_error("gru_step_fwd: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  PVCR_TRY(linear_fwd(x, V, w_ih, V, b_ih, gi, 3 * H, B, 3 * H, V, nsplit, scratch, need, st));
  if (h_prev) PVCR_TRY(linear_fwd(h_prev, H, w_hh, H, nullptr, gh, 3 * H, B, 3 * H, H, nsplit, scratch, need, st));
  GruFwdArgs g{};
  g.B = B; g.H = H;
  g.gi_a = gi; g.gi_a_ld = 3 * H;
  g.gh = h_prev ? gh : nullptr; g.gh_ld = 3 * H; g.b_hh = b_hh;
  g.h_prev = h_prev; g.h_prev_ld = H;
  g.h_out = h_out; g.h_out_ld = H;
  g.nsplit = nsplit;
  const size_t o = (size_t)B * H;
  g.r = saved; g.z = saved + o; g.n = saved + 2 * o; g.ghn = saved + 3 * o;
  return gru_gate_fwd(g, st);
}

// d_x [B,V], d_h_prev [B,H] overwritten (nullable); d_w_* / d_b_* overwritten, or accumulated into when accumulate != 0.
int gru_step_bwd(const float* d_h, const float* x, const float* h_prev, const float* w_ih, const float* w_hh,
                 const float* saved, int B, int V, int H, int nsplit, float* d_x, float* d_h_prev, float* d_w_ih,
                 float* d_w_hh, float* d_b_ih, float* d_b_hh, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  PVCR_REQUIRE(B > 0 && V > 0 && H > 0 && nsplit >= 1 && nsplit <= 3, "gru_step_bwd: B=%d V=%d H=%d nsplit=%d", B, V, H, nsplit);
  Arena a(ws, ws_bytes);
  a.alloc<float>((size_t)2 * B * 3 * H);
  float* dgi = a.alloc<float>((size_t)B * 3 * H);
  float* dgh = a.alloc<float>((size_t)B * 3 * H);
  float* dhz = a.alloc<float>((size_t)B * H);
  const size_t need = gru_step_scratch(B, V, H, nsplit);
  char* scratch = a.alloc<char>(need);
  if (a.failed) { set_last_error("gru_step_bwd: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  GruBwdArgs b{};
  b.B = B; b.H = H;
  b.dh_a = d_h; b.dh_a_ld = H;
  const size_t o = (size_t)B * H;
  b.r = saved; b.z = saved + o; b.n = saved + 2 * o; b.ghn = saved + 3 * o;
  b.h_prev = h_prev; b.h_prev_ld = H;
  b.dgi = dgi; b.dgi_ld = 3 * H; b.dgh = dgh; b.dgh_ld = 3 * H;
  b.nsplit = nsplit;
  b.dh_direct = dhz; b.dh_direct_ld = H;          // dh * z
  PVCR_TRY(gru_gate_bwd(b, st));
  PVCR_TRY(linear_bwd(dgi, 3 * H, x, V, w_ih, V, d_x, V, d_w_ih, V, d_b_ih, B, 3 * H, V, nsplit, accumulate, scratch, need, st));
  if (h_prev) {
    PVCR_TRY(linear_bwd(dgh, 3 * H, h_prev, H, w_hh, H, d_h_prev, H, d_w_hh, H, d_b_hh, B, 3 * H, H, nsplit, accumulate,
                        scratch, need, st));
    if (d_h_prev) PVCR_TRY(add_inplace(d_h_prev, dhz, (long long)B * H, st));
  } else {
    if (d_b_hh) PVCR_TRY(colsum(dgh, 3 * H, B, 3 * H, d_b_hh, accumulate, st));
    if (d_w_hh && !accumulate) PVCR_TRY(fill_zero(d_w_hh, sizeof(float) * (size_t)3 * H * H, st));
    if (d_h_prev) PVCR_CUDA_CHECK(cudaMemcpyAsync(d_h_prev, dhz, sizeof(float) * o, cudaMemcpyDeviceToDevice, st));
  }
  return PVCR_OK;
}

}  // namespace pvcr
