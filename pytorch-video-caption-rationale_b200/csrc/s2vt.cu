// S2VTModel (two stacked GRUs, Venugopalan et al. 2015), teacher-forced forward and hand-written backward.
// Reference: model/S2VTModel.py:74-86 (encode), :88-145 (decode, training branch), :179-202 (forward).
//
// Restructuring relative to the reference loop (same arithmetic):
//   * rnn1 encoding stage: input projection hoisted into one GEMM over all B*N frames; its decoding stage sees
//     an all-zero input, so its input projection is just b_ih (the reference multiplies W_ih by zeros);
//   * rnn2 encoding stage: the word half of its input is zero padding, so only W_ih[:, :H] out1 is computed,
//     hoisted over all frames; decoding stage (teacher forced): W_ih[:, :H] h1 and W_ih[:, H:] Emb[w] are both
//     hoisted over all B*L tokens because rnn1's decoding states do not depend on rnn2;
//   * four GRU sequences (rnn1 enc/dec, rnn2 enc/dec) run on the persistent recurrent kernels when eligible;
//   * every weight gradient is one GEMM over all steps.
#include "../../include/pvcr_b200.h"
#include "host.h"

namespace pvcr {

struct SeqBuf {                 // one GRU sequence of T steps over B videos, rows ordered (b, t)
  int T;
  float *h, *r, *z, *n, *ghn;   // h [B*T, H] ; saved gates [T][B,H]
  Planes hp;                    // bf16 planes of h, rows b*T + t
  float *dgi, *dgh, *hprev;     // backward: [B*T, 3H] x2, [B*T, H]
};

struct S2vtWs {
  Planes w1i, w1h, w2o, w2e, w2h, x_a, emb_a;
  SeqBuf e1, d1, e2, d2;        // rnn1 enc / dec, rnn2 enc / dec   (d2.h is the caller's hs buffer)
  float *gi1, *gi2e, *gi2d, *gh;
  Planes w1hT, w2hT, w2oT, w2eT, w1iT, dgh_a;
  float *dh1, *dh2, *d_h1d, *d_out1, *demb_rows, *dxsel;
  Planes h0_a;                  // decode(): bf16 rows / fp32 copy of a caller-given rnn1 state
  float* h0_f;
  unsigned* sync;
  bf16* xch;
};

static void carve_seq(Arena& a, int B, int T, int H, int ns, SeqBuf& s, bool own_h) {
  const size_t R = (size_t)B * T;
  s.T = T;
  s.h = own_h ? a.alloc<float>(R * H) : nullptr;
  s.r = a.alloc<float>(R * H); s.z = a.alloc<float>(R * H); s.n = a.alloc<float>(R * H); s.ghn = a.alloc<float>(R * H);
  s.hp = alloc_planes(a, (int)R, H, ns);
  s.dgi = a.alloc<float>(R * 3 * H); s.dgh = a.alloc<float>(R * 3 * H); s.hprev = a.alloc<float>(R * H);
}

static void carve_s2vt(Arena& a, const PvcrDims& d, int need_frame_grad, S2vtWs& w) {
  const int B = d.B, N = d.N, V = d.V, H = d.H, E = d.E, L = d.L, ns = d.nsplit;
  const size_t BN = (size_t)B * N, BL = (size_t)B * L;
  w.w1i = alloc_planes(a, 3 * H, V, ns); w.w1h = alloc_planes(a, 3 * H, H, ns);
  w.w2o = alloc_planes(a, 3 * H, H, ns); w.w2e = alloc_planes(a, 3 * H, E, ns); w.w2h = alloc_planes(a, 3 * H, H, ns);
  w.x_a = alloc_planes(a, (int)BN, V, ns);
  w.emb_a = alloc_planes(a, (int)BL, E, ns);
  carve_seq(a, B, N, H, ns, w.e1, true); carve_seq(a, B, L, H, ns, w.d1, true);
  carve_seq(a, B, N, H, ns, w.e2, true); carve_seq(a, B, L, H, ns, w.d2, false);
  w.gi1 = a.alloc<float>(BN * 3 * H); w.gi2e = a.alloc<float>(BN * 3 * H); w.gi2d = a.alloc<float>(BL * 3 * H);
  w.gh = a.alloc<float>((size_t)B * 3 * H);
  w.w1hT = alloc_planes(a, H, 3 * H, ns); w.w2hT = alloc_planes(a, H, 3 * H, ns);
  w.w2oT = alloc_planes(a, H, 3 * H, ns); w.w2eT = alloc_planes(a, E, 3 * H, ns);
  if (need_frame_grad) w.w1iT = alloc_planes(a, V, 3 * H, ns); else w.w1iT = Planes{};
  w.dgh_a = alloc_planes(a, B, 3 * H, ns);
  w.dh1 = a.alloc<float>((size_t)B * H); w.dh2 = a.alloc<float>((size_t)B * H);
  w.d_h1d = a.alloc<float>(BL * H); w.d_out1 = a.alloc<float>(BN * H);
  w.demb_rows = a.alloc<float>(BL * E);
  w.dxsel = need_frame_grad ? a.alloc<float>(BN * V) : nullptr;
  w.sync = a.alloc<unsigned>(32 * 160);
  w.xch = a.alloc<bf16>((size_t)2 * B * 4 * H);
  w.h0_a = alloc_planes(a, B, H, ns);
  w.h0_f = a.alloc<float>((size_t)B * H);
}

static size_t s2vt_scratch(const PvcrDims& d, int need_frame_grad) {
  Arena a(nullptr, 0);
  size_t peak = 0;
  auto gw = [&](int R, int N, int K) {
    const size_t need = a.mark() + grad_w_scratch(R, N, K, d.nsplit);
    if (need > peak) peak = need;
  };
  auto gx = [&](int R, int N) {
    size_t m = a.mark();
    alloc_planes(a, R, N, d.nsplit);
    if (a.off > peak) peak = a.off;
    a.release(m);
  };
  const int BL = d.B * d.L, BN = d.B * d.N, H = d.H;
  gw(BL, 3 * H, H); gw(BL, 3 * H, d.E); gx(BL, 3 * H); gw(BN, 3 * H, H); gx(BN, 3 * H); gw(BN, 3 * H, d.V);
  return peak + 4096;
}

size_t s2vt_workspace(const PvcrDims& d, int need_frame_grad) {
  Arena a(nullptr, 0);
  S2vtWs w;
  carve_s2vt(a, d, need_frame_grad, w);
  return a.off + s2vt_scratch(d, need_frame_grad) + 1024;
}

// GRU sequence descriptor over a SeqBuf; gi rows (b, t) with 3H columns (nullable), initial state = step
// `prev_T - 1` of `prev` (nullable: zeros).
static GruSeq make_seq(const PvcrDims& d, const SeqBuf& s, const float* gi, const float* gi_bias, const float* b_hh,
                       const Planes& whh, const SeqBuf* prev, float* gh, unsigned* sync) {
  const int H = d.H, T = s.T;
  GruSeq q{};
  q.T = T; q.B = d.B; q.H = H; q.nsplit = d.nsplit;
  q.gi_a = gi; q.gi_a_ts = 3 * H; q.gi_a_ld = (long long)T * 3 * H;
  q.gi_bias = gi_bias; q.b_hh = b_hh; q.whh = whh;
  if (prev) {
    q.h0 = prev->h + (long long)(prev->T - 1) * H; q.h0_ld = (long long)prev->T * H;
    q.h0_planes = prev->hp.ptr + (long long)(prev->T - 1) * prev->hp.ld; q.h0_planes_ld = (long long)prev->T * prev->hp.ld;
  }
  q.h = s.h; q.h_ts = H; q.h_ld = (long long)T * H;
  q.hp = s.hp.ptr; q.hp_ts = s.hp.ld; q.hp_ld = (long long)T * s.hp.ld; q.Hp = s.hp.Kp;
  q.gh = gh; q.r = s.r; q.z = s.z; q.n = s.n; q.ghn = s.ghn; q.sync = sync;
  return q;
}

struct S2vtSeqs { GruSeq e1, d1, e2, d2; };
// given: decode() mode (model/S2VTModel.py:88, called by SpatialNet.py:140) -- rnn1's outputs over the frames and its
// state are caller inputs; the decoding stage of rnn1 starts from the staged copy of that state (h0_f / h0_a).
static S2vtSeqs make_seqs(const PvcrDims& d, const PvcrS2vtParams& p, S2vtWs& w, float* hs, bool given = false) {
  w.d2.h = hs;
  S2vtSeqs q;
  q.e1 = make_seq(d, w.e1, w.gi1, nullptr, p.rnn1_b_hh, w.w1h, nullptr, w.gh, w.sync);
  q.d1 = make_seq(d, w.d1, nullptr, p.rnn1_b_ih, p.rnn1_b_hh, w.w1h, &w.e1, w.gh, w.sync);
  if (given) { q.d1.h0 = w.h0_f; q.d1.h0_ld = d.H; q.d1.h0_planes = w.h0_a.ptr; q.d1.h0_planes_ld = w.h0_a.ld; }
  q.e2 = make_seq(d, w.e2, w.gi2e, nullptr, p.rnn2_b_hh, w.w2h, nullptr, w.gh, w.sync);
  q.d2 = make_seq(d, w.d2, w.gi2d, nullptr, p.rnn2_b_hh, w.w2h, &w.e2, w.gh, w.sync);
  return q;
}

static int check_s2vt_dims(const PvcrDims& d) {
  PVCR_REQUIRE(d.B > 0 && d.N > 0 && d.V > 0 && d.H > 0 && d.E > 0 && d.L > 0 && d.Vc > 0,
               "dims must be positive: B=%d N=%d V=%d H=%d E=%d L=%d Vc=%d", d.B, d.N, d.V, d.H, d.E, d.L, d.Vc);
  PVCR_REQUIRE(d.nsplit >= 1 && d.nsplit <= 3, "nsplit=%d not in 1..3", d.nsplit);
  return PVCR_OK;
}

static Dropout emb_dropout(const PvcrDims& d) { return make_dropout(d.dropout_p, d.seed, 0x3000000000ull); }

// weights -> planes, rnn1 over the frames, rnn2 encoding stage
static int s2vt_encode(const PvcrDims& d, const PvcrS2vtParams& p, S2vtWs& w, const S2vtSeqs& q, const float* vid,
                       const float* frame_scale, cudaStream_t st, const float* out1_given = nullptr,
                       const float* state1_given = nullptr) {
  const int B = d.B, N = d.N, V = d.V, H = d.H, E = d.E;
  const int BN = B * N, H3 = 3 * H;
  if (!out1_given) PVCR_TRY(prep_weight(p.rnn1_w_ih, V, H3, V, w.w1i, st));
  PVCR_TRY(prep_weight(p.rnn1_w_hh, H, H3, H, w.w1h, st));
  PVCR_TRY(prep_weight(p.rnn2_w_ih, H + E, H3, H, w.w2o, st));
  PVCR_TRY(prep_weight(p.rnn2_w_ih + H, H + E, H3, E, w.w2e, st));
  PVCR_TRY(prep_weight(p.rnn2_w_hh, H, H3, H, w.w2h, st));
  if (w.e1.hp.Kp != H) {
    for (SeqBuf* s : {&w.e1, &w.d1, &w.e2, &w.d2})
      PVCR_TRY(fill_zero(s->hp.ptr, sizeof(bf16) * (size_t)s->hp.rows * s->hp.ld, st));
  }
  if (out1_given) {
    PVCR_CUDA_CHECK(cudaMemcpyAsync(w.e1.h, out1_given, sizeof(float) * (size_t)BN * H, cudaMemcpyDeviceToDevice, st));
    PVCR_CUDA_CHECK(cudaMemcpyAsync(w.h0_f, state1_given, sizeof(float) * (size_t)B * H, cudaMemcpyDeviceToDevice, st));
    if (w.h0_a.Kp != H) PVCR_TRY(fill_zero(w.h0_a.ptr, sizeof(bf16) * (size_t)B * w.h0_a.ld, st));
    PVCR_TRY(stage(out1_given, H, BN, H, w.e1.hp, 0, nullptr, NO_DROPOUT, st));
    PVCR_TRY(stage(state1_given, H, B, H, w.h0_a, 0, nullptr, NO_DROPOUT, st));
  } else {
  PVCR_TRY(stage(vid, V, BN, V, w.x_a, 0, frame_scale, NO_DROPOUT, st));
  PVCR_TRY(gemm_planes(w.x_a.view(), w.w1i.view(), BN, H3, (int)w.x_a.ld, w.gi1, H3, p.rnn1_b_ih, 0, st));
  PVCR_TRY(gru_seq_fwd(q.e1, st));
  }
  // rnn2 encoding stage: [out1 ; 0] -> only the out1 half of W_ih contributes
  PVCR_TRY(gemm_planes(w.e1.hp.view(), w.w2o.view(), BN, H3, (int)w.e1.hp.ld, w.gi2e, H3, p.rnn2_b_ih, 0, st));
  PVCR_TRY(gru_seq_fwd(q.e2, st));
  return PVCR_OK;
}

static int s2vt_fwd_impl(const PvcrDims& d, const PvcrS2vtParams& p, const float* vid, const float* frame_scale,
                         const float* out1_given, const float* state1_given, const long long* s_in, float* hs, void* ws,
                         size_t ws_bytes, cudaStream_t st) {
  PVCR_TRY(check_s2vt_dims(d));
  const int B = d.B, H = d.H, E = d.E, L = d.L;
  const int BL = B * L, H3 = 3 * H;
  Arena a(ws, ws_bytes);
  S2vtWs w;
  carve_s2vt(a, d, 0, w);
  if (a.failed) { set_last_error("s2vt_fwd: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  S2vtSeqs q = make_seqs(d, p, w, hs, out1_given != nullptr);
  PVCR_TRY(s2vt_encode(d, p, w, q, vid, frame_scale, st, out1_given, state1_given));
  PVCR_TRY(gru_seq_fwd(q.d1, st));
  // rnn2 decoding stage: [h1_dec ; Dropout(Emb[w])]
  PVCR_TRY(gemm_planes(w.d1.hp.view(), w.w2o.view(), BL, H3, (int)w.d1.hp.ld, w.gi2d, H3, p.rnn2_b_ih, 0, st));
  PVCR_TRY(gather_split(p.emb, E, s_in, BL, w.emb_a.ptr, w.emb_a.ld, w.emb_a.Kp, d.nsplit, emb_dropout(d), st));
  PVCR_TRY(gemm_planes(w.emb_a.view(), w.w2e.view(), BL, H3, (int)w.emb_a.ld, w.gi2d, H3, nullptr, 1, st));
  PVCR_TRY(gru_seq_fwd(q.d2, st));
  return PVCR_OK;
}

int s2vt_fwd(const PvcrDims& d, const PvcrS2vtParams& p, const float* vid, const float* frame_scale,
             const long long* s_in, float* hs, void* ws, size_t ws_bytes, cudaStream_t st) {
  return s2vt_fwd_impl(d, p, vid, frame_scale, nullptr, nullptr, s_in, hs, ws, ws_bytes, st);
}
int s2vt_decode_fwd(const PvcrDims& d, const PvcrS2vtParams& p, const float* out1, const float* state1,
                    const long long* s_in, float* hs, void* ws, size_t ws_bytes, cudaStream_t st) {
  PVCR_REQUIRE(out1 && state1, "s2vt_decode_fwd: null rnn1 outputs / state");
  return s2vt_fwd_impl(d, p, nullptr, nullptr, out1, state1, s_in, hs, ws, ws_bytes, st);
}

// rows (b, t) of h_{t-1}: step 0 takes `first` (row stride first_ld; null = zeros), steps >= 1 the sequence itself
static int build_hprev(const SeqBuf& s, int B, int H, const float* first, long long first_ld, cudaStream_t st) {
  const int T = s.T;
  if (first)
    PVCR_CUDA_CHECK(cudaMemcpy2DAsync(s.hprev, sizeof(float) * (size_t)T * H, first, sizeof(float) * (size_t)first_ld,
                                      sizeof(float) * H, B, cudaMemcpyDeviceToDevice, st));
  else
    PVCR_CUDA_CHECK(cudaMemset2DAsync(s.hprev, sizeof(float) * (size_t)T * H, 0, sizeof(float) * H, B, st));
  if (T > 1)
    PVCR_CUDA_CHECK(cudaMemcpy2DAsync(s.hprev + H, sizeof(float) * (size_t)T * H, s.h, sizeof(float) * (size_t)T * H,
                                      sizeof(float) * (size_t)(T - 1) * H, B, cudaMemcpyDeviceToDevice, st));
  return PVCR_OK;
}

static GruSeqGrad make_grad(const SeqBuf& s, int H, const float* dh_ext, float* dh_carry, const S2vtWs& w,
                            const Planes& whhT) {
  GruSeqGrad g{};
  g.dh_ext = dh_ext; g.dh_ext_ts = H; g.dh_ext_ld = (long long)s.T * H;
  g.dh_carry = dh_carry;
  g.dgi = s.dgi; g.dgi_ts = 3 * H; g.dgi_ld = (long long)s.T * 3 * H;
  g.dgh = s.dgh; g.dgh_ts = 3 * H; g.dgh_ld = (long long)s.T * 3 * H;
  g.dgh_a = w.dgh_a; g.whhT = whhT; g.xch = w.xch;
  return g;
}

static int s2vt_bwd_impl(const PvcrDims& d, const PvcrS2vtParams& p, const float* vid, const float* frame_scale,
             const long long* s_in, const float* d_hs, float* hs, PvcrS2vtGrads& g, float* d_frame_scale,
             float* d_out1, float* d_state1, void* ws, size_t ws_bytes, cudaStream_t st) {
  PVCR_TRY(check_s2vt_dims(d));
  const bool given = d_out1 != nullptr;
  const int B = d.B, N = d.N, V = d.V, H = d.H, E = d.E, L = d.L, ns = d.nsplit;
  const int BN = B * N, BL = B * L, H3 = 3 * H;
  const int need_frame_grad = d_frame_scale != nullptr;
  Arena a(ws, ws_bytes);
  S2vtWs w;
  carve_s2vt(a, d, need_frame_grad, w);
  if (a.failed || a.off + s2vt_scratch(d, need_frame_grad) > ws_bytes) {
    set_last_error("s2vt_bwd: workspace too small (%zu bytes)", ws_bytes);
    return PVCR_ERR_WORKSPACE;
  }
  S2vtSeqs q = make_seqs(d, p, w, hs, given);
  PVCR_TRY(prep_weight_T(p.rnn1_w_hh, H, H3, H, w.w1hT, 0, 1, st));
  PVCR_TRY(prep_weight_T(p.rnn2_w_hh, H, H3, H, w.w2hT, 0, 1, st));
  PVCR_TRY(prep_weight_T(p.rnn2_w_ih, H + E, H3, H, w.w2oT, 0, 1, st));
  PVCR_TRY(prep_weight_T(p.rnn2_w_ih + H, H + E, H3, E, w.w2eT, 0, 1, st));
  if (need_frame_grad) PVCR_TRY(prep_weight_T(p.rnn1_w_ih, V, H3, V, w.w1iT, 0, 1, st));
  if (w.dgh_a.Kp != H3) PVCR_TRY(fill_zero(w.dgh_a.ptr, sizeof(bf16) * (size_t)B * w.dgh_a.ld, st));

  // ---- rnn2, reverse time: decoding stage (gradient d_hs on every state), then encoding stage ----
  PVCR_TRY(fill_zero(w.dh2, sizeof(float) * (size_t)B * H, st));
  PVCR_TRY(gru_seq_bwd(q.d2, make_grad(w.d2, H, d_hs, w.dh2, w, w.w2hT), st));
  PVCR_TRY(gru_seq_bwd(q.e2, make_grad(w.e2, H, nullptr, w.dh2, w, w.w2hT), st));
  PVCR_TRY(build_hprev(w.d2, B, H, w.e2.h + (long long)(N - 1) * H, (long long)N * H, st));
  PVCR_TRY(build_hprev(w.e2, B, H, nullptr, 0, st));
  PVCR_TRY(grad_w(a, w.d2.dgh, H3, BL, H3, w.d2.hprev, H, H, nullptr, nullptr, g.rnn2_w_hh, H, 0, ns, st));
  PVCR_TRY(grad_w(a, w.e2.dgh, H3, BN, H3, w.e2.hprev, H, H, nullptr, nullptr, g.rnn2_w_hh, H, 1, ns, st));
  PVCR_TRY(colsum(w.d2.dgh, H3, BL, H3, g.rnn2_b_hh, 0, st));
  PVCR_TRY(colsum(w.e2.dgh, H3, BN, H3, g.rnn2_b_hh, 1, st));
  PVCR_TRY(colsum(w.d2.dgi, H3, BL, H3, g.rnn2_b_ih, 0, st));
  PVCR_TRY(colsum(w.e2.dgi, H3, BN, H3, g.rnn2_b_ih, 1, st));
  // W_ih of rnn2 = [W_o | W_e]: W_o sees h1 (decoding) and out1 (encoding), W_e the embedded words
  PVCR_TRY(grad_w(a, w.d2.dgi, H3, BL, H3, w.d1.h, H, H, nullptr, nullptr, g.rnn2_w_ih, H + E, 0, ns, st));
  PVCR_TRY(grad_w(a, w.e2.dgi, H3, BN, H3, w.e1.h, H, H, nullptr, nullptr, g.rnn2_w_ih, H + E, 1, ns, st));
  // (dropout mask of the embedded words re-applied while gathering)
  PVCR_TRY(grad_w(a, w.d2.dgi, H3, BL, H3, p.emb, E, E, s_in, nullptr, g.rnn2_w_ih + H, H + E, 0, ns, st, emb_dropout(d)));
  PVCR_TRY(grad_x(a, w.d2.dgi, H3, BL, H3, w.w2eT, w.demb_rows, E, 0, st));
  PVCR_TRY(fill_zero(g.emb, sizeof(float) * (size_t)d.Vc * E, st));
  PVCR_TRY(scatter_add_rows(w.demb_rows, E, s_in, BL, E, g.emb, emb_dropout(d), st));
  PVCR_TRY(grad_x(a, w.d2.dgi, H3, BL, H3, w.w2oT, w.d_h1d, H, 0, st));
  PVCR_TRY(grad_x(a, w.e2.dgi, H3, BN, H3, w.w2oT, w.d_out1, H, 0, st));

  // ---- rnn1, reverse time ----
  PVCR_TRY(fill_zero(w.dh1, sizeof(float) * (size_t)B * H, st));
  PVCR_TRY(gru_seq_bwd(q.d1, make_grad(w.d1, H, w.d_h1d, w.dh1, w, w.w1hT), st));
  if (given) {      // decode(): rnn1's frame outputs and state were inputs -- return their gradients, no frame sweep
    PVCR_CUDA_CHECK(cudaMemcpyAsync(d_out1, w.d_out1, sizeof(float) * (size_t)BN * H, cudaMemcpyDeviceToDevice, st));
    PVCR_CUDA_CHECK(cudaMemcpyAsync(d_state1, w.dh1, sizeof(float) * (size_t)B * H, cudaMemcpyDeviceToDevice, st));
    PVCR_TRY(build_hprev(w.d1, B, H, w.h0_f, H, st));
    PVCR_TRY(grad_w(a, w.d1.dgh, H3, BL, H3, w.d1.hprev, H, H, nullptr, nullptr, g.rnn1_w_hh, H, 0, ns, st));
    PVCR_TRY(colsum(w.d1.dgh, H3, BL, H3, g.rnn1_b_hh, 0, st));
    PVCR_TRY(colsum(w.d1.dgi, H3, BL, H3, g.rnn1_b_ih, 0, st));
    return PVCR_OK;
  }
  PVCR_TRY(gru_seq_bwd(q.e1, make_grad(w.e1, H, w.d_out1, w.dh1, w, w.w1hT), st));
  PVCR_TRY(build_hprev(w.d1, B, H, w.e1.h + (long long)(N - 1) * H, (long long)N * H, st));
  PVCR_TRY(build_hprev(w.e1, B, H, nullptr, 0, st));
  PVCR_TRY(grad_w(a, w.d1.dgh, H3, BL, H3, w.d1.hprev, H, H, nullptr, nullptr, g.rnn1_w_hh, H, 0, ns, st));
  PVCR_TRY(grad_w(a, w.e1.dgh, H3, BN, H3, w.e1.hprev, H, H, nullptr, nullptr, g.rnn1_w_hh, H, 1, ns, st));
  PVCR_TRY(colsum(w.d1.dgh, H3, BL, H3, g.rnn1_b_hh, 0, st));
  PVCR_TRY(colsum(w.e1.dgh, H3, BN, H3, g.rnn1_b_hh, 1, st));
  PVCR_TRY(colsum(w.d1.dgi, H3, BL, H3, g.rnn1_b_ih, 0, st));
  PVCR_TRY(colsum(w.e1.dgi, H3, BN, H3, g.rnn1_b_ih, 1, st));
  PVCR_TRY(grad_w(a, w.e1.dgi, H3, BN, H3, vid, V, V, nullptr, frame_scale, g.rnn1_w_ih, V, 0, ns, st));
  if (need_frame_grad) {
    PVCR_TRY(grad_x(a, w.e1.dgi, H3, BN, H3, w.w1iT, w.dxsel, V, 0, st));
    PVCR_TRY(rowdot(vid, w.dxsel, BN, V, d_frame_scale, st));
  }
  return PVCR_OK;
}

int s2vt_bwd(const PvcrDims& d, const PvcrS2vtParams& p, const float* vid, const float* frame_scale,
             const long long* s_in, const float* d_hs, float* hs, PvcrS2vtGrads& g, float* d_frame_scale, void* ws,
             size_t ws_bytes, cudaStream_t st) {
  return s2vt_bwd_impl(d, p, vid, frame_scale, s_in, d_hs, hs, g, d_frame_scale, nullptr, nullptr, ws, ws_bytes, st);
}
// Backward of s2vt_decode_fwd: all gradients of `g` except rnn1_w_ih (not on this path), d_out1 [B,N,H], d_state1 [B,H].
int s2vt_decode_bwd(const PvcrDims& d, const PvcrS2vtParams& p, const long long* s_in, const float* d_hs, float* hs,
                    PvcrS2vtGrads& g, float* d_out1, float* d_state1, void* ws, size_t ws_bytes, cudaStream_t st) {
  PVCR_REQUIRE(d_out1 && d_state1, "s2vt_decode_bwd: null gradient outputs");
  return s2vt_bwd_impl(d, p, nullptr, nullptr, s_in, d_hs, hs, g, nullptr, d_out1, d_state1, ws, ws_bytes, st);
}

// ---- step-wise decoding with word feedback -----------------------------------------------------------------------
// Eval branch (model/S2VTModel.py:147-177: arg-max always fed back) and the scheduled-sampling training branch
// (:121-141: per step, the teacher word or the arg-max according to a host-drawn coin).
struct S2vtStepWs {
  S2vtWs w;
  Planes wv, emb_step, hdrop;
  float *logits_step, *hs, *g2;
  long long* words;
};
static void carve_steps(Arena& a, const PvcrDims& d, S2vtStepWs& g) {
  carve_s2vt(a, d, 0, g.w);
  g.wv = alloc_planes(a, d.Vc, d.H, d.nsplit);
  g.emb_step = alloc_planes(a, d.B, d.E, d.nsplit);
  g.hdrop = alloc_planes(a, d.B, d.H, d.nsplit);
  g.logits_step = a.alloc<float>((size_t)d.B * round_up(d.Vc, 4));
  g.hs = a.alloc<float>((size_t)d.B * d.L * d.H);
  g.g2 = a.alloc<float>((size_t)d.B * 3 * d.H);
  g.words = a.alloc<long long>(d.B);
}
size_t s2vt_decode_steps_workspace(const PvcrDims& d) {
  Arena a(nullptr, 0);
  S2vtStepWs g;
  carve_steps(a, d, g);
  return a.off + 4096;
}

static int gru_single_step(const PvcrDims& d, const SeqBuf& s, int i, const SeqBuf& prev_stage, const Planes& whh,
                           const float* gi, const float* gi_bias, const float* b_hh, float* gh, cudaStream_t st) {
  const int B = d.B, H = d.H, T = s.T, H3 = 3 * H;
  OperandView hprev_a = (i == 0)
      ? OperandView{prev_stage.hp.ptr + (long long)(prev_stage.T - 1) * prev_stage.hp.ld,
                    (long long)prev_stage.T * prev_stage.hp.ld, 0, B, 1}
      : OperandView{s.hp.ptr + (long long)(i - 1) * s.hp.ld, (long long)T * s.hp.ld, 0, B, 1};
  PVCR_TRY(gemm_planes(hprev_a, whh.view(), B, H3, (int)whh.ld, gh, H3, nullptr, 0, st));
  GruFwdArgs g{};
  g.B = B; g.H = H;
  g.gi_a = gi; g.gi_a_ld = H3; g.gi_bias = gi_bias;
  g.gh = gh; g.gh_ld = H3; g.b_hh = b_hh;
  if (i == 0) { g.h_prev = prev_stage.h + (long long)(prev_stage.T - 1) * H; g.h_prev_ld = (long long)prev_stage.T * H; }
  else { g.h_prev = s.h + (long long)(i - 1) * H; g.h_prev_ld = (long long)T * H; }
  g.h_out = s.h + (long long)i * H; g.h_out_ld = (long long)T * H;
  g.h_planes = s.hp.ptr + (long long)i * s.hp.ld; g.h_planes_ld = (long long)T * s.hp.ld;
  g.Hp = s.hp.Kp; g.nsplit = d.nsplit;
  return gru_gate_fwd(g, st);
}

// teacher_mask: HOST array of L ints (coin of step i decides the word fed to step i+1); null = always feed back.
// out1_given / state1_given: decode() mode (see s2vt_decode_fwd) -- vid / frame_scale unused.
int s2vt_decode_steps_impl(const PvcrDims& d, const PvcrS2vtParams& p, const float* vid, const float* frame_scale,
                      const float* out1_given, const float* state1_given,
                      long long sos_id, const long long* teacher_words, const int* teacher_mask, float out_dropout_p,
                      long long* ids, long long* fed, float* logits, void* ws, size_t ws_bytes, cudaStream_t st) {
  PVCR_TRY(check_s2vt_dims(d));
  PVCR_REQUIRE((out1_given != nullptr) == (state1_given != nullptr), "s2vt decode: rnn1 outputs and state come together");
  const int B = d.B, H = d.H, E = d.E, L = d.L, Vc = d.Vc, H3 = 3 * H;
  Arena a(ws, ws_bytes);
  S2vtStepWs gw;
  carve_steps(a, d, gw);
  if (a.failed) { set_last_error("s2vt_decode_steps: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  S2vtWs& w = gw.w;
  S2vtSeqs q = make_seqs(d, p, w, gw.hs, out1_given != nullptr);
  PVCR_TRY(s2vt_encode(d, p, w, q, vid, frame_scale, st, out1_given, state1_given));
  SeqBuf first1 = w.e1;          // the stage whose last state starts rnn1's decoding steps
  if (out1_given) { first1 = SeqBuf{}; first1.T = 1; first1.h = w.h0_f; first1.hp = w.h0_a; }
  PVCR_TRY(prep_weight(p.out_w, H, Vc, H, gw.wv, st));
  if (gw.hdrop.Kp != H) PVCR_TRY(fill_zero(gw.hdrop.ptr, sizeof(bf16) * (size_t)B * gw.hdrop.ld, st));
  PVCR_TRY(fill_i64(gw.words, sos_id, B, st));
  for (int i = 0; i < L; ++i) {
    if (fed) PVCR_CUDA_CHECK(cudaMemcpy2DAsync(fed + i, sizeof(long long) * L, gw.words, sizeof(long long),
                                               sizeof(long long), B, cudaMemcpyDeviceToDevice, st));
    PVCR_TRY(gru_single_step(d, w.d1, i, first1, w.w1h, nullptr, p.rnn1_b_ih, p.rnn1_b_hh, w.gh, st));
    OperandView h1_a{w.d1.hp.ptr + (long long)i * w.d1.hp.ld, (long long)L * w.d1.hp.ld, 0, B, 1};
    PVCR_TRY(gemm_planes(h1_a, w.w2o.view(), B, H3, (int)w.w2o.ld, gw.g2, H3, p.rnn2_b_ih, 0, st));
    Dropout ed = emb_dropout(d);
    ed.row_mul = L; ed.row_add = i;
    PVCR_TRY(gather_split(p.emb, E, gw.words, B, gw.emb_step.ptr, gw.emb_step.ld, gw.emb_step.Kp, d.nsplit, ed, st));
    PVCR_TRY(gemm_planes(gw.emb_step.view(), w.w2e.view(), B, H3, (int)gw.emb_step.ld, gw.g2, H3, nullptr, 1, st));
    PVCR_TRY(gru_single_step(d, w.d2, i, w.e2, w.w2h, gw.g2, nullptr, p.rnn2_b_hh, w.gh, st));
    Dropout od = make_dropout(out_dropout_p, d.seed, 0x5000000000ull);
    od.row_mul = L; od.row_add = i;
    PVCR_TRY(cast_split(gw.hs + (long long)i * H, (long long)L * H, B, H, gw.hdrop.ptr, gw.hdrop.ld, gw.hdrop.Kp,
                        d.nsplit, 0, nullptr, od, st));
    float* lg = logits ? logits + (long long)i * Vc : gw.logits_step;
    const long long ldl = logits ? (long long)L * Vc : round_up(Vc, 4);
    PVCR_TRY(gemm_planes(gw.hdrop.view(), gw.wv.view(), B, Vc, (int)gw.wv.ld, lg, ldl, p.out_b, 0, st));
    const int use_teacher = (teacher_mask && teacher_words && i + 1 < L) ? teacher_mask[i] : 0;
    PVCR_TRY(argmax_rows(lg, ldl, B, Vc, ids + i, L, gw.words, teacher_words ? teacher_words + i + 1 : nullptr, L,
                         use_teacher, st));
  }
  return PVCR_OK;
}
int s2vt_decode_steps(const PvcrDims& d, const PvcrS2vtParams& p, const float* vid, const float* frame_scale,
                      long long sos_id, const long long* teacher_words, const int* teacher_mask, float out_dropout_p,
                      long long* ids, long long* fed, float* logits, void* ws, size_t ws_bytes, cudaStream_t st) {
  return s2vt_decode_steps_impl(d, p, vid, frame_scale, nullptr, nullptr, sos_id, teacher_words, teacher_mask,
                                out_dropout_p, ids, fed, logits, ws, ws_bytes, st);
}

}  // namespace pvcr
