// RationaleNet frame-selection generator (model/RationaleNet.py:14-54): bidirectional LSTM over the frames,
// Dropout, Linear(2H -> 2), 2-class Gumbel-softmax (soft in training, straight-through one-hot in eval), and the
// brevity / continuity penalties of train_utils.py:73-95.  The selected features sel = vid_feats * p1 are never
// materialised: p1 [B,N] is handed to the caption network as `frame_scale` and applied while its encoder input
// is staged; the matching gradient d p1 = sum_v vid * d sel comes back from the caption network's backward.
//
// Restructuring: both directions' input projections are one GEMM [B*N, V] x [V, 8H]; the recurrences are
// per-step (tcgen05 GEMM h W_hh^T + fused LSTM cell kernel); every weight gradient is one GEMM over all steps.
#include "../../include/pvcr_b200.h"
#include "host.h"

namespace pvcr {

struct DirBuf {                   // one LSTM direction, rows ordered (b, t)
  Planes whh, whhT, hp;
  float *h, *i, *f, *g, *o, *c, *hprev;
};
struct GenWs {
  Planes wih, x_a, da_a;
  DirBuf dir[2];
  float *bias_cat, *gi, *gh, *y, *dgi, *dh_carry, *dc, *dhf, *dhb, *dlogit;
  unsigned* sync;
  bf16* xch;
};

static LstmSeqArgs dir_seq(const PvcrDims& d, const PvcrGenParams& p, const GenWs& w, int k) {
  const int N = d.N, H = d.H;
  const DirBuf& r = w.dir[k];
  LstmSeqArgs s{};
  s.T = N; s.B = d.B; s.H = H; s.rev = k;
  s.whh = r.whh; s.b_hh = k ? p.b_hh_r : p.b_hh;
  s.gi = w.gi + (long long)k * 4 * H; s.gi_ts = 8 * H; s.gi_ld = (long long)N * 8 * H;
  s.h = r.h; s.h_ts = H; s.h_ld = (long long)N * H;
  s.hp = r.hp.ptr; s.hp_ts = r.hp.ld; s.hp_ld = (long long)N * r.hp.ld;
  s.si = r.i; s.sf = r.f; s.sg = r.g; s.so = r.o; s.sc = r.c;
  s.sync = w.sync + (size_t)k * 32 * 160;       // each direction its own group counters (the two may run side by side)
  return s;
}

static void carve_gen(Arena& a, const PvcrDims& d, GenWs& w) {
  const int B = d.B, N = d.N, V = d.V, H = d.H, ns = d.nsplit;
  const size_t BN = (size_t)B * N;
  w.wih = alloc_planes(a, 8 * H, V, ns);
  w.x_a = alloc_planes(a, (int)BN, V, ns);
  w.da_a = alloc_planes(a, B, 4 * H, ns);
  for (int k = 0; k < 2; ++k) {
    DirBuf& r = w.dir[k];
    r.whh = alloc_planes(a, 4 * H, H, ns);
    r.whhT = alloc_planes(a, H, 4 * H, ns);
    r.hp = alloc_planes(a, (int)BN, H, ns);
    r.h = a.alloc<float>(BN * H);
    r.i = a.alloc<float>(BN * H); r.f = a.alloc<float>(BN * H); r.g = a.alloc<float>(BN * H);
    r.o = a.alloc<float>(BN * H); r.c = a.alloc<float>(BN * H);
    r.hprev = a.alloc<float>(BN * H);
  }
  w.bias_cat = a.alloc<float>((size_t)8 * H);
  w.gi = a.alloc<float>(BN * 8 * H);
  w.gh = a.alloc<float>((size_t)B * 4 * H);
  w.y = a.alloc<float>(BN * 2);
  w.dgi = a.alloc<float>(BN * 8 * H);
  w.dh_carry = a.alloc<float>((size_t)B * H);
  w.dc = a.alloc<float>((size_t)B * H);
  w.dhf = a.alloc<float>(BN * H); w.dhb = a.alloc<float>(BN * H);
  w.dlogit = a.alloc<float>(BN * 2);
  w.sync = a.alloc<unsigned>(2 * 32 * 160);
  w.xch = a.alloc<bf16>((size_t)2 * 2 * B * 4 * H);      // [direction][2][B][4H]
}

static size_t gen_scratch(const PvcrDims& d) {
  Arena a(nullptr, 0);
  size_t peak = 0;
  auto gw = [&](int R, int N, int K) {
    const size_t need = a.mark() + grad_w_scratch(R, N, K, d.nsplit);
    if (need > peak) peak = need;
  };
  gw(d.B * d.N, 8 * d.H, d.V); gw(d.B * d.N, 4 * d.H, d.H);
  return peak + 4096;
}

size_t generator_workspace(const PvcrDims& d) {
  Arena a(nullptr, 0);
  GenWs w;
  carve_gen(a, d, w);
  return a.off + gen_scratch(d) + 1024;
}

static Dropout gen_dropout(const PvcrDims& d) { return make_dropout(d.dropout_p, d.seed, 0x7000000000ull); }

// noise: [B*N, 2] Exp(1) draws (row b*N + n) or NULL (drawn in-kernel from dims.seed).
int generator_fwd(const PvcrDims& d, const PvcrGenParams& p, const float* vid, const float* noise, float tau, int hard,
                  float* probs, float* p1, float* pen, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int B = d.B, N = d.N, V = d.V, H = d.H, BN = B * N, H4 = 4 * H, H8 = 8 * H;
  PVCR_REQUIRE(B > 0 && N > 0 && V > 0 && H > 0 && d.nsplit >= 1 && d.nsplit <= 3, "generator_fwd: bad dims");
  PVCR_REQUIRE(tau > 0.f, "generator_fwd: tau=%f must be positive", tau);
  Arena a(ws, ws_bytes);
  GenWs w;
  carve_gen(a, d, w);
  if (a.failed) { set_last_error("generator_fwd: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  PVCR_TRY(prep_weight(p.w_ih, V, H4, V, w.wih, st, 0));
  PVCR_TRY(prep_weight(p.w_ih_r, V, H4, V, w.wih, st, H4));
  PVCR_TRY(prep_weight(p.w_hh, H, H4, H, w.dir[0].whh, st));
  PVCR_TRY(prep_weight(p.w_hh_r, H, H4, H, w.dir[1].whh, st));
  PVCR_CUDA_CHECK(cudaMemcpyAsync(w.bias_cat, p.b_ih, sizeof(float) * H4, cudaMemcpyDeviceToDevice, st));
  PVCR_CUDA_CHECK(cudaMemcpyAsync(w.bias_cat + H4, p.b_ih_r, sizeof(float) * H4, cudaMemcpyDeviceToDevice, st));
  PVCR_TRY(stage(vid, V, BN, V, w.x_a, 0, nullptr, NO_DROPOUT, st));
  PVCR_TRY(gemm_planes(w.x_a.view(), w.wih.view(), BN, H8, (int)w.x_a.ld, w.gi, H8, w.bias_cat, 0, st));
  // The two directions are independent: when both fit the SMs with 32-video groups (2 x 64 CTAs at B = 128, H = 512)
  // the reverse sweep runs on a side lane next to the forward one instead of behind it.
  bool paired = false;
  if (lstm_persist_pair_ok(B, H, d.nsplit, w.dir[0].hp.Kp) && side_site(4)) {
    cudaStream_t lane = st;
    PVCR_TRY(side_fork(st, &lane, 0));
    if (lane != st) {
      PVCR_TRY(lstm_persist_fwd(dir_seq(d, p, w, 0), st, true));
      PVCR_TRY(lstm_persist_fwd(dir_seq(d, p, w, 1), lane, true));
      PVCR_TRY(side_join_lane(st, 0));
      paired = true;
    }
  }
  for (int k = 0; k < 2 && !paired; ++k) {
    DirBuf& r = w.dir[k];
    if (r.hp.Kp != H) PVCR_TRY(fill_zero(r.hp.ptr, sizeof(bf16) * (size_t)BN * r.hp.ld, st));
    const float* b_hh = k ? p.b_hh_r : p.b_hh;
    if (lstm_persist_eligible(B, H, d.nsplit, r.hp.Kp)) {
      PVCR_TRY(lstm_persist_fwd(dir_seq(d, p, w, k), st));
      continue;
    }
    for (int s = 0; s < N; ++s) {
      const int t = k ? N - 1 - s : s, tp = k ? t + 1 : t - 1;
      if (s > 0) {
        OperandView hp_a{r.hp.ptr + (long long)tp * r.hp.ld, (long long)N * r.hp.ld, 0, B, 1};
        PVCR_TRY(gemm_planes(hp_a, r.whh.view(), B, H4, (int)r.whh.ld, w.gh, H4, nullptr, 0, st));
      }
      LstmFwdArgs g{};
      g.B = B; g.H = H;
      g.gi = w.gi + (long long)t * H8 + (long long)k * H4; g.gi_ld = (long long)N * H8;
      g.gh = s > 0 ? w.gh : nullptr; g.gh_ld = H4;
      g.b_hh = b_hh;
      g.c_prev = s > 0 ? r.c + (long long)tp * B * H : nullptr;
      g.h_out = r.h + (long long)t * H; g.h_out_ld = (long long)N * H;
      g.h_planes = r.hp.ptr + (long long)t * r.hp.ld; g.h_planes_ld = (long long)N * r.hp.ld;
      g.Hp = r.hp.Kp; g.nsplit = d.nsplit;
      const long long o = (long long)t * B * H;
      g.i = r.i + o; g.f = r.f + o; g.g = r.g + o; g.o = r.o + o; g.c = r.c + o;
      PVCR_TRY(lstm_gate_fwd(g, st));
    }
  }
  GumbelArgs ga{};
  ga.B = B; ga.N = N; ga.H = H;
  ga.hf = w.dir[0].h; ga.hb = w.dir[1].h; ga.h_ts = H; ga.h_bs = (long long)N * H;
  ga.w = p.lin_w; ga.bias = p.lin_b; ga.noise = noise; ga.seed = d.seed; ga.seed_step = seed_step_ptr(); ga.tau = tau; ga.hard = hard;
  ga.drop = gen_dropout(d);
  ga.probs = probs; ga.p1 = p1; ga.y = w.y; ga.pen = pen;
  return gumbel_select_fwd(ga, st);
}

// d_p1: gradient on the frame scale from the caption network (nullable); d_probs: external gradient on probs
// [B,N,2] (nullable); g_pen: device [2] = d loss / d (brevity, continuity) (nullable).
int generator_bwd(const PvcrDims& d, const PvcrGenParams& p, const float* vid, float tau, const float* d_p1,
                  const float* d_probs, const float* g_pen, PvcrGenGrads& g, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
  const int B = d.B, N = d.N, V = d.V, H = d.H, BN = B * N, H4 = 4 * H, H8 = 8 * H, ns = d.nsplit;
  Arena a(ws, ws_bytes);
  GenWs w;
  carve_gen(a, d, w);
  if (a.failed || a.off + gen_scratch(d) > ws_bytes) {
    set_last_error("generator_bwd: workspace too small (%zu bytes)", ws_bytes);
    return PVCR_ERR_WORKSPACE;
  }
  GumbelBwdArgs gb{};
  gb.B = B; gb.N = N; gb.H = H;
  gb.hf = w.dir[0].h; gb.hb = w.dir[1].h; gb.h_ts = H; gb.h_bs = (long long)N * H;
  gb.w = p.lin_w; gb.y = w.y; gb.tau = tau; gb.drop = gen_dropout(d);
  gb.dp1_sel = d_p1; gb.dprobs = d_probs; gb.g_pen = g_pen;
  gb.dhf = w.dhf; gb.dhb = w.dhb; gb.dw = g.lin_w; gb.dbias = g.lin_b; gb.scratch = w.dlogit;
  PVCR_TRY(gumbel_select_bwd(gb, st));
  if (w.da_a.Kp != H4) PVCR_TRY(fill_zero(w.da_a.ptr, sizeof(bf16) * (size_t)B * w.da_a.ld, st));
  // Pairing the two BACKWARD sweeps the same way is measured slower (cfg3: 655 us for the pair vs 2 x 314 us one after the
  // other): a 32-video group's gate gradients (128 KB) do not fit beside the W_hh^T slice, so they come back in two
  // serialised 64 KB passes per step.  Off unless asked for.
  static const bool pair_bwd = getenv("PVCR_LSTM_PAIR_BWD") != nullptr;
  bool paired = false;
  if (pair_bwd && lstm_persist_pair_ok(B, H, ns, w.dir[0].hp.Kp) && side_site(4)) {
    for (int k = 0; k < 2; ++k) PVCR_TRY(prep_weight_T(k ? p.w_hh_r : p.w_hh, H, H4, H, w.dir[k].whhT, 0, 1, st));
    cudaStream_t lane = st;
    PVCR_TRY(side_fork(st, &lane, 0));
    if (lane != st) {
      for (int k = 0; k < 2; ++k)
        PVCR_TRY(lstm_persist_bwd(dir_seq(d, p, w, k), w.dir[k].whhT, k ? w.dhb : w.dhf, H, (long long)N * H,
                                  w.dgi + (long long)k * H4, H8, (long long)N * H8, w.xch + (size_t)k * 2 * B * H4,
                                  k ? lane : st, true));
      PVCR_TRY(side_join_lane(st, 0));
      paired = true;
    }
  }
  for (int k = 0; k < 2; ++k) {
    DirBuf& r = w.dir[k];
    if (!paired) PVCR_TRY(prep_weight_T(k ? p.w_hh_r : p.w_hh, H, H4, H, r.whhT, 0, 1, st));
    PVCR_TRY(fill_zero(w.dh_carry, sizeof(float) * (size_t)B * H, st));
    PVCR_TRY(fill_zero(w.dc, sizeof(float) * (size_t)B * H, st));
    const float* dh_ext = k ? w.dhb : w.dhf;
    const bool persist = paired || lstm_persist_eligible(B, H, ns, r.hp.Kp);
    if (persist && !paired)
      PVCR_TRY(lstm_persist_bwd(dir_seq(d, p, w, k), r.whhT, dh_ext, H, (long long)N * H, w.dgi + (long long)k * H4, H8,
                                (long long)N * H8, w.xch, st));
    for (int s = N - 1; s >= 0 && !persist; --s) {
      const int t = k ? N - 1 - s : s, tp = k ? t + 1 : t - 1;
      LstmBwdArgs b{};
      b.B = B; b.H = H;
      b.dh_a = w.dh_carry; b.dh_a_ld = H;
      b.dh_b = dh_ext + (long long)t * H; b.dh_b_ld = (long long)N * H;
      b.dc = w.dc;
      const long long o = (long long)t * B * H;
      b.i = r.i + o; b.f = r.f + o; b.g = r.g + o; b.o = r.o + o; b.c = r.c + o;
      b.c_prev = s > 0 ? r.c + (long long)tp * B * H : nullptr;
      b.da = w.dgi + (long long)t * H8 + (long long)k * H4; b.da_ld = (long long)N * H8;
      b.da_planes = s > 0 ? w.da_a.ptr : nullptr; b.da_planes_ld = w.da_a.ld; b.Kp = w.da_a.Kp; b.nsplit = ns;
      PVCR_TRY(lstm_gate_bwd(b, st));
      if (s > 0)
        PVCR_TRY(gemm_planes(w.da_a.view(), r.whhT.view(), B, H, (int)w.da_a.ld, w.dh_carry, H, nullptr, 0, st));
    }
    // h_{prev} rows in (b, t) order: forward direction h[t-1] (zero at t = 0), reverse direction h[t+1] (zero at N-1)
    PVCR_TRY(fill_zero(r.hprev, sizeof(float) * (size_t)BN * H, st));
    if (N > 1) {
      const float* src = k ? r.h + H : r.h;
      float* dst = k ? r.hprev : r.hprev + H;
      PVCR_CUDA_CHECK(cudaMemcpy2DAsync(dst, sizeof(float) * (size_t)N * H, src, sizeof(float) * (size_t)N * H,
                                        sizeof(float) * (size_t)(N - 1) * H, B, cudaMemcpyDeviceToDevice, st));
    }
    float* dwhh = k ? g.w_hh_r : g.w_hh;
    PVCR_TRY(grad_w(a, w.dgi + (long long)k * H4, H8, BN, H4, r.hprev, H, H, nullptr, nullptr, dwhh, H, 0, ns, st));
    float* dbih = k ? g.b_ih_r : g.b_ih;
    float* dbhh = k ? g.b_hh_r : g.b_hh;
    PVCR_TRY(colsum(w.dgi + (long long)k * H4, H8, BN, H4, dbih, 0, st));
    PVCR_CUDA_CHECK(cudaMemcpyAsync(dbhh, dbih, sizeof(float) * H4, cudaMemcpyDeviceToDevice, st));
  }
  if (g.w_ih_r == g.w_ih + (size_t)H4 * V) {     // contiguous [8H, V] gradient buffer: one GEMM for both directions
    PVCR_TRY(grad_w(a, w.dgi, H8, BN, H8, vid, V, V, nullptr, nullptr, g.w_ih, V, 0, ns, st));
  } else {
    PVCR_TRY(grad_w(a, w.dgi, H8, BN, H4, vid, V, V, nullptr, nullptr, g.w_ih, V, 0, ns, st));
    PVCR_TRY(grad_w(a, w.dgi + H4, H8, BN, H4, vid, V, V, nullptr, nullptr, g.w_ih_r, V, 0, ns, st));
  }
  return PVCR_OK;
}

}  // namespace pvcr
