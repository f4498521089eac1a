// S2VTAttModel encoder + attention decoder, forward and hand-written backward.
// Reference: model/S2VTAttModel.py:80-96 (Encoder.forward), :125-148 (Decoder.forward_step),
// :150-196 (Decoder.forward), :25-48 (Attention.forward).
//
// Restructuring relative to the reference loop (same arithmetic):
//   * encoder input projection hoisted into one GEMM over all B*N frames (K1);
//   * W_ih of the decoder split into [Wc | We]; the embedding half is hoisted into one GEMM over all B*L
//     tokens (K8/K9), the context half stays in the step;
//   * W_q and W_hh of the decoder are concatenated so that q = W_q h and gh = W_hh h are one GEMM per step;
//   * every weight gradient is hoisted out of the time loop into one GEMM over all steps.
#include <cstdlib>

#include "../../include/pvcr_b200.h"
#include "host.h"

namespace pvcr {

struct AttWs {
  // prepared weights (forward)
  Planes wih_enc, whh_enc, wk, wcat, wc, we;
  // forward activations
  Planes x_a, enc_a, emb_a, ctx_a, hs_a, h0_a;
  float* h0_f;
  float *gi_enc, *enc, *er, *ez, *en, *eghn, *gh, *pk, *ep, *g1_all, *alpha_all, *ctx_all, *g2, *dr, *dz, *dn, *dghn;
  // backward
  Planes wcT, wcatT, wkT, weT, whh_encT, wih_encT, dgi_a, d1_a, dgh_a;
  float *dh_carry, *dgi_all, *d1_all, *dctx, *dpk, *denc, *dv_part, *dgi_enc, *dgh_enc, *hprev_dec, *hprev_enc,
      *demb_rows, *dxsel;
  unsigned* sync;
  bf16* xch;
  bf16* ctx_x;
  float *dctx_all, *ds_all;
  bf16* xg;
};

static size_t scratch_need(const PvcrDims& d, int need_frame_grad) {
  // peak of the transient operand planes used by grad_w / grad_x in the hoisted gradient GEMMs
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
  gw(BL, 4 * H, H); gw(BL, 3 * H, H); gw(BL, 3 * H, d.E); gx(BL, 3 * H);
  gw(BN, H, H); gx(BN, H); gw(BN, 3 * H, H); gw(BN, 3 * H, d.V);
  if (need_frame_grad) gx(BN, 3 * H);
  return peak + 4096;
}

static void carve(Arena& a, const PvcrDims& d, int need_frame_grad, AttWs& w) {
  const int B = d.B, N = d.N, V = d.V, H = d.H, E = d.E, L = d.L, ns = d.nsplit;
  const size_t BN = (size_t)B * N, BL = (size_t)B * L;
  w.wih_enc = alloc_planes(a, 3 * H, V, ns);
  w.whh_enc = alloc_planes(a, 3 * H, H, ns);
  w.wk = alloc_planes(a, H, H, ns);
  w.wcat = alloc_planes(a, 4 * H, H, ns);
  w.wc = alloc_planes(a, 3 * H, H, ns);
  w.we = alloc_planes(a, 3 * H, E, ns);
  w.x_a = alloc_planes(a, (int)BN, V, ns);
  w.enc_a = alloc_planes(a, (int)BN, H, ns);
  w.emb_a = alloc_planes(a, (int)BL, E, ns);
  w.ctx_a = alloc_planes(a, B, H, ns);
  w.hs_a = alloc_planes(a, (int)BL, H, ns);
  w.h0_a = alloc_planes(a, B, H, ns);                 // decode(): bf16 rows of a caller-given initial state
  w.h0_f = a.alloc<float>((size_t)B * H);
  w.gi_enc = a.alloc<float>(BN * 3 * H);
  w.enc = a.alloc<float>(BN * H);
  w.er = a.alloc<float>(BN * H); w.ez = a.alloc<float>(BN * H); w.en = a.alloc<float>(BN * H); w.eghn = a.alloc<float>(BN * H);
  w.gh = a.alloc<float>((size_t)B * 3 * H);
  w.pk = a.alloc<float>(BN * H);
  w.ep = a.alloc<float>(BL * 3 * H);
  w.g1_all = a.alloc<float>(BL * 4 * H);
  w.alpha_all = a.alloc<float>(BL * N);
  w.ctx_all = a.alloc<float>(BL * H);
  w.g2 = a.alloc<float>((size_t)B * 3 * H);
  w.dr = a.alloc<float>(BL * H); w.dz = a.alloc<float>(BL * H); w.dn = a.alloc<float>(BL * H); w.dghn = a.alloc<float>(BL * H);
  // backward
  w.wcT = alloc_planes(a, H, 3 * H, ns);
  w.wcatT = alloc_planes(a, H, 4 * H, ns);
  w.wkT = alloc_planes(a, H, H, ns);
  w.weT = alloc_planes(a, E, 3 * H, ns);
  w.whh_encT = alloc_planes(a, H, 3 * H, ns);
  if (need_frame_grad) w.wih_encT = alloc_planes(a, V, 3 * H, ns); else w.wih_encT = Planes{};
  w.dgi_a = alloc_planes(a, B, 3 * H, ns);
  w.d1_a = alloc_planes(a, B, 4 * H, ns);
  w.dgh_a = alloc_planes(a, B, 3 * H, ns);
  w.dh_carry = a.alloc<float>((size_t)B * H);
  w.dgi_all = a.alloc<float>(BL * 3 * H);
  w.d1_all = a.alloc<float>(BL * 4 * H);
  w.dctx = a.alloc<float>((size_t)B * H);
  w.dpk = a.alloc<float>(BN * H);
  w.denc = a.alloc<float>(BN * H);
  w.dv_part = a.alloc<float>((size_t)B * H);
  w.dgi_enc = a.alloc<float>(BN * 3 * H);
  w.dgh_enc = a.alloc<float>(BN * 3 * H);
  w.hprev_dec = a.alloc<float>(BL * H);
  w.hprev_enc = a.alloc<float>(BN * H);
  w.demb_rows = a.alloc<float>(BL * E);
  w.dxsel = need_frame_grad ? a.alloc<float>(BN * V) : nullptr;
  w.sync = a.alloc<unsigned>(32 * 160);
  w.dctx_all = a.alloc<float>(BL * H);
  w.ds_all = a.alloc<float>(BL * N);
  w.xg = a.alloc<bf16>((size_t)2 * B * 5 * H);
  w.xch = a.alloc<bf16>((size_t)2 * B * 4 * H);
  w.ctx_x = a.alloc<bf16>(BL * H);
}

// bf16 copies kept by the staging cache of one backward call (bf16 mode): the decoder half's planes first, the
// encoder half's after them.  A part-2 call starts its cache behind the decoder half's region: with deferred side-lane
// joins the lane may still be reading those planes while part 2 stages its own.
static size_t dec_cache_need(const PvcrDims& d) {
  Arena a(nullptr, 0);
  const int BL = d.B * d.L, BN = d.B * d.N, H = d.H;
  alloc_planes(a, BL, H, 1); alloc_planes(a, BL, 3 * H, 1); alloc_planes(a, BL, H, 1);      // dq, dgh, h_prev
  alloc_planes(a, BL, 3 * H, 1); alloc_planes(a, BL, H, 1);                                  // dgi, ctx
  alloc_planes(a, BN, H, 1);                                                                 // dpk
  return a.off + 1024;
}
static size_t stage_cache_need(const PvcrDims& d) {
  Arena a(nullptr, 0);
  const int BN = d.B * d.N, H = d.H;
  alloc_planes(a, BN, 3 * H, 1); alloc_planes(a, BN, H, 1); alloc_planes(a, BN, 3 * H, 1);   // dgh_enc, h_prev_enc, dgi_enc
  return dec_cache_need(d) + a.off + 4096;
}

size_t s2vtatt_workspace(const PvcrDims& d, int need_frame_grad) {
  Arena a(nullptr, 0);
  AttWs w;
  carve(a, d, need_frame_grad, w);
  return a.off + scratch_need(d, need_frame_grad) + stage_cache_need(d) + 1024;
}

static int check_dims(const PvcrDims& d) {
  PVCR_REQUIRE(d.B > 0 && d.N > 0 && d.V > 0 && d.H > 0 && d.E > 0 && d.L > 0 && d.Vc > 0,
               "dims must be positive: B=%d N=%d V=%d H=%d E=%d L=%d Vc=%d", d.B, d.N, d.V, d.H, d.E, d.L, d.Vc);
  PVCR_REQUIRE(d.nsplit >= 1 && d.nsplit <= 3, "nsplit=%d not in 1..3", d.nsplit);
  return PVCR_OK;
}

static GruSeq encoder_seq(const PvcrDims& d, const PvcrS2vtAttParams& p, const AttWs& w) {
  const int H = d.H, N = d.N;
  GruSeq s{};
  s.T = N; s.B = d.B; s.H = H; s.nsplit = d.nsplit;
  s.gi_a = w.gi_enc; s.gi_a_ts = 3 * H; s.gi_a_ld = (long long)N * 3 * H;       // rows b*N + t
  s.b_hh = p.enc_b_hh;
  s.whh = w.whh_enc;
  s.h = w.enc; s.h_ts = H; s.h_ld = (long long)N * H;
  s.hp = w.enc_a.ptr; s.hp_ts = w.enc_a.ld; s.hp_ld = (long long)N * w.enc_a.ld; s.Hp = w.enc_a.Kp;
  s.gh = w.gh;
  s.r = w.er; s.z = w.ez; s.n = w.en; s.ghn = w.eghn;
  s.sync = w.sync;
  return s;
}

// Initial decoder state h_{-1}.  forward(): the encoder's final state = frame N-1 of its outputs (Decoder.forward is
// handed encoder_final, model/S2VTAttModel.py:261-262).  decode(): a separate caller-given tensor
// (model/S2VTAttModel.py:231-243, called by SpatialNet.py:140), staged into h0_f / h0_a by the forward call.
struct H0 { const float* f; long long f_ld; const bf16* a; long long a_ld; };
static H0 initial_state(const PvcrDims& d, const AttWs& w, bool given) {
  if (given) return H0{w.h0_f, d.H, w.h0_a.ptr, w.h0_a.ld};
  return H0{w.enc + (long long)(d.N - 1) * d.H, (long long)d.N * d.H, w.enc_a.ptr + (long long)(d.N - 1) * w.enc_a.ld,
            (long long)d.N * w.enc_a.ld};
}

// Transposed weight planes of the backward sweeps (they depend on the parameters only).
static int att_bwd_weights(const PvcrDims& d, const PvcrS2vtAttParams& p, const AttWs& w, cudaStream_t st) {
  const int H = d.H, E = d.E, H3 = 3 * H;
  PVCR_TRY(prep_weight_T(p.dec_w_ih, H + E, H3, H, w.wcT, 0, 1, st));
  PVCR_TRY(fill_zero(w.wcatT.ptr, sizeof(bf16) * (size_t)w.wcatT.rows * w.wcatT.ld, st));
  PVCR_TRY(prep_weight_T(p.att_wq, H, H, H, w.wcatT, 0, 0, st));
  PVCR_TRY(prep_weight_T(p.dec_w_hh, H, H3, H, w.wcatT, H, 0, st));
  PVCR_TRY(prep_weight_T(p.enc_w_hh, H, H3, H, w.whh_encT, 0, 1, st));
  return PVCR_OK;
}

// enc_given / final_given (both or neither): decode() mode -- the encoder outputs [B,N,H] and the initial decoder state
// [B,H] come from the caller (SpatialNet drives its own per-frame encoder loop), vid / frame_scale are unused.
static int s2vtatt_fwd_impl(const PvcrDims& d, const PvcrS2vtAttParams& p, const float* vid, const float* frame_scale,
                            const float* enc_given, const float* final_given, const long long* s_in, float* hs,
                            float* alphas, void* ws, size_t ws_bytes, cudaStream_t st) {
  PVCR_TRY(check_dims(d));
  const bool given = enc_given != nullptr;
  PVCR_REQUIRE(given == (final_given != nullptr), "s2vtatt decode: encoder outputs and final state come together");
  const int B = d.B, N = d.N, V = d.V, H = d.H, E = d.E, L = d.L;
  const int BN = B * N, BL = B * L, H3 = 3 * H, H4 = 4 * H;
  Arena a(ws, ws_bytes);
  AttWs w;
  carve(a, d, 0, w);
  if (a.failed) { set_last_error("s2vtatt_fwd: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }

  // Critical chain: {W_ih cast | frame staging} -> input-projection GEMM -> {W_hh cast} -> encoder sweep -> proj_key.
  // Everything else that only depends on the inputs runs on side lanes next to it:
  //   lane 0: frame staging (joined before the GEMM);
  //   lane 1: the recurrent / attention weight casts (joined before the sweep);
  //   lane 2: the hoisted embedding half of the decoder input projection (+ b_ih), then the transposed weight planes of
  //           the backward sweeps (the backward call finds a note and skips them); joined before the decoder sweep.
  cudaStream_t l0 = st, l1 = st, l2 = st;
  const bool fork = side_site(0);
  if (fork) { PVCR_TRY(side_fork(st, &l0, 0)); PVCR_TRY(side_fork(st, &l1, 1)); PVCR_TRY(side_fork(st, &l2, 2)); }
  if (!given) {
  PVCR_TRY(stage(vid, V, BN, V, w.x_a, 0, frame_scale, NO_DROPOUT, l0));
  PVCR_TRY(prep_weight(p.enc_w_ih, V, H3, V, w.wih_enc, st));
  if (fork) PVCR_TRY(side_join_lane(st, 0));         // waits for the frame staging only (nothing else is on lane 0 yet)
  PVCR_TRY(gemm_planes(w.x_a.view(), w.wih_enc.view(), BN, H3, (int)w.x_a.ld, w.gi_enc, H3, p.enc_b_ih, 0, st));
  PVCR_TRY(prep_weight(p.enc_w_hh, H, H3, H, w.whh_enc, l1));
  if (fork) PVCR_TRY(side_join_lane(st, 1));         // the sweep waits for its W_hh planes only
  }

  PVCR_TRY(prep_weight(p.att_wk, H, H, H, w.wk, l0));
  PVCR_TRY(prep_weight(p.att_wq, H, H, H, w.wcat, l0, 0));
  PVCR_TRY(prep_weight(p.dec_w_hh, H, H3, H, w.wcat, l1, H));
  PVCR_TRY(prep_weight(p.dec_w_ih, H + E, H3, H, w.wc, l1));
  if (w.enc_a.Kp != H) {       // contraction padding of the planes written by the gate kernels must read as zero
    PVCR_TRY(fill_zero(w.enc_a.ptr, sizeof(bf16) * (size_t)BN * w.enc_a.ld, st));
    PVCR_TRY(fill_zero(w.hs_a.ptr, sizeof(bf16) * (size_t)BL * w.hs_a.ld, l1));
    PVCR_TRY(fill_zero(w.ctx_a.ptr, sizeof(bf16) * (size_t)B * w.ctx_a.ld, l1));
  }

  PVCR_TRY(prep_weight(p.dec_w_ih + H, H + E, H3, E, w.we, l2));
  PVCR_TRY(gather_split(p.emb, E, s_in, BL, w.emb_a.ptr, w.emb_a.ld, w.emb_a.Kp, d.nsplit, NO_DROPOUT, l2));
  PVCR_TRY(gemm_planes(w.emb_a.view(), w.we.view(), BL, H3, (int)w.emb_a.ld, w.ep, H3, p.dec_b_ih, 0, l2));
  side_note_take(ws, NOTE_ATT_BWD_WEIGHTS);          // a stale note of an earlier forward on this workspace
  if (fork) {
    PVCR_TRY(att_bwd_weights(d, p, w, l2));
    side_note_put(ws, NOTE_ATT_BWD_WEIGHTS, p.dec_w_hh);
  }
  if (given) {
    // decode(): encoder outputs and the initial state are inputs; stage their fp32 copies and bf16 operand rows
    PVCR_CUDA_CHECK(cudaMemcpyAsync(w.enc, enc_given, sizeof(float) * (size_t)BN * H, cudaMemcpyDeviceToDevice, st));
    PVCR_CUDA_CHECK(cudaMemcpyAsync(w.h0_f, final_given, sizeof(float) * (size_t)B * H, cudaMemcpyDeviceToDevice, st));
    if (w.h0_a.Kp != H) PVCR_TRY(fill_zero(w.h0_a.ptr, sizeof(bf16) * (size_t)B * w.h0_a.ld, st));
    PVCR_TRY(stage(enc_given, H, BN, H, w.enc_a, 0, nullptr, NO_DROPOUT, st));
    PVCR_TRY(stage(final_given, H, B, H, w.h0_a, 0, nullptr, NO_DROPOUT, st));
  } else {
  // encoder: N recurrent steps on gi = (vid * frame_scale) W_ih^T + b_ih
  PVCR_TRY(gru_seq_fwd(encoder_seq(d, p, w), st));
  }
  const H0 h0 = initial_state(d, w, given);

  // proj_key = enc W_k^T
  PVCR_TRY(side_join(st));
  PVCR_TRY(gemm_planes(w.enc_a.view(), w.wk.view(), BN, H, (int)w.enc_a.ld, w.pk, H, nullptr, 0, st));

  // decoder steps: one persistent cooperative kernel when the shape allows, else per-step launches
  if (dec_persist_eligible(B, N, H, d.nsplit, w.enc_a.Kp)) {
    DecPersistFwd q{};
    q.L = L; q.B = B; q.N = N; q.H = H;
    q.w1 = w.wcat.ptr; q.w1_ld = w.wcat.ld; q.w3 = w.wc.ptr; q.w3_ld = w.wc.ld;
    q.b_hh = p.dec_b_hh; q.v = p.att_v; q.pk = w.pk; q.enc_a = w.enc_a.ptr; q.enc_ld = w.enc_a.ld; q.enc = w.enc;
    q.h0 = h0.f; q.h0_ld = h0.f_ld; q.h0_a = h0.a; q.h0_a_ld = h0.a_ld;
    q.ep = w.ep; q.q_all = w.g1_all; q.q_ld = H4; q.ctx_x = w.ctx_x; q.ctx_all = w.ctx_all; q.alpha = w.alpha_all;
    q.hs = hs; q.hs_a = w.hs_a.ptr; q.hs_a_ld = w.hs_a.ld;
    q.r = w.dr; q.z = w.dz; q.n = w.dn; q.ghn = w.dghn; q.counters = w.sync;
    PVCR_TRY(dec_persist_fwd(q, st));
  } else
  for (int i = 0; i < L; ++i) {
    OperandView hprev_a = (i == 0)
        ? OperandView{const_cast<bf16*>(h0.a), h0.a_ld, 0, B, 1}
        : OperandView{w.hs_a.ptr + (long long)(i - 1) * w.hs_a.ld, (long long)L * w.hs_a.ld, 0, B, 1};
    float* g1 = w.g1_all + (long long)i * B * H4;
    PVCR_TRY(gemm_planes(hprev_a, w.wcat.view(), B, H4, (int)w.wcat.ld, g1, H4, nullptr, 0, st));
    AttnFwdArgs at{};
    at.B = B; at.N = N; at.H = H;
    at.q = g1; at.q_ld = H4; at.pk = w.pk; at.enc = w.enc; at.v = p.att_v;
    at.alpha = w.alpha_all + (long long)i * B * N;
    at.ctx = w.ctx_all + (long long)i * H; at.ctx_ld = (long long)L * H;             // rows b*L + i
    at.ctx_planes = w.ctx_a.ptr; at.ctx_planes_ld = w.ctx_a.ld; at.Hp = w.ctx_a.Kp; at.nsplit = d.nsplit;
    PVCR_TRY(attn_fwd(at, st));
    PVCR_TRY(gemm_planes(w.ctx_a.view(), w.wc.view(), B, H3, (int)w.ctx_a.ld, w.g2, H3, nullptr, 0, st));
    GruFwdArgs g{};
    g.B = B; g.H = H;
    g.gi_a = w.g2; g.gi_a_ld = H3;
    g.gi_b = w.ep + (long long)i * H3; g.gi_b_ld = (long long)L * H3;
    g.gh = g1 + H; g.gh_ld = H4; g.b_hh = p.dec_b_hh;
    if (i == 0) { g.h_prev = h0.f; g.h_prev_ld = h0.f_ld; }
    else { g.h_prev = hs + (long long)(i - 1) * H; g.h_prev_ld = (long long)L * H; }
    g.h_out = hs + (long long)i * H; g.h_out_ld = (long long)L * H;
    g.h_planes = w.hs_a.ptr + (long long)i * w.hs_a.ld; g.h_planes_ld = (long long)L * w.hs_a.ld;
    g.Hp = w.hs_a.Kp; g.nsplit = d.nsplit;
    const long long o = (long long)i * B * H;
    g.r = w.dr + o; g.z = w.dz + o; g.n = w.dn + o; g.ghn = w.dghn + o;
    PVCR_TRY(gru_gate_fwd(g, st));
  }
  if (alphas)
    PVCR_CUDA_CHECK(cudaMemcpyAsync(alphas, w.alpha_all, sizeof(float) * (size_t)L * B * N, cudaMemcpyDeviceToDevice, st));
  return PVCR_OK;
}

int s2vtatt_fwd(const PvcrDims& d, const PvcrS2vtAttParams& p, const float* vid, const float* frame_scale,
                const long long* s_in, float* hs, float* alphas, void* ws, size_t ws_bytes, cudaStream_t st) {
  return s2vtatt_fwd_impl(d, p, vid, frame_scale, nullptr, nullptr, s_in, hs, alphas, ws, ws_bytes, st);
}
int s2vtatt_decode_fwd(const PvcrDims& d, const PvcrS2vtAttParams& p, const float* enc_outs, const float* enc_final,
                       const long long* s_in, float* hs, float* alphas, void* ws, size_t ws_bytes, cudaStream_t st) {
  PVCR_REQUIRE(enc_outs && enc_final, "s2vtatt_decode_fwd: null encoder outputs / final state");
  return s2vtatt_fwd_impl(d, p, nullptr, nullptr, enc_outs, enc_final, s_in, hs, alphas, ws, ws_bytes, st);
}

// part: 0 = whole backward; 1 = decoder half only (every decoder / attention / embedding gradient is final when it
// returns); 2 = encoder half only (must follow part 1 on the same workspace).  The split lets a data-parallel caller
// all-reduce the decoder gradients while the encoder sweep runs.
// d_enc_outs / d_enc_final (both or neither): decode() mode (forward ran through s2vtatt_decode_fwd) -- the gradients on
// the caller's encoder outputs and initial state are returned instead of being swept through the encoder.
static int s2vtatt_bwd_impl(const PvcrDims& d, const PvcrS2vtAttParams& p, const float* vid, const float* frame_scale,
                const long long* s_in, const float* d_hs, const float* hs, PvcrS2vtAttGrads& g, float* d_frame_scale,
                float* d_enc_outs, float* d_enc_final, void* ws, size_t ws_bytes, cudaStream_t st, int part) {
  PVCR_TRY(check_dims(d));
  const bool given = d_enc_outs != nullptr;
  if (given) part = 1;
  const int B = d.B, N = d.N, V = d.V, H = d.H, E = d.E, L = d.L, ns = d.nsplit;
  const int BN = B * N, BL = B * L, H3 = 3 * H, H4 = 4 * H;
  const int need_frame_grad = d_frame_scale != nullptr;
  Arena a(ws, ws_bytes);
  AttWs w;
  carve(a, d, need_frame_grad, w);
  if (a.failed || a.off + scratch_need(d, need_frame_grad) + stage_cache_need(d) > ws_bytes) {
    set_last_error("s2vtatt_bwd: workspace too small (%zu bytes)", ws_bytes);
    return PVCR_ERR_WORKSPACE;
  }
  StageCache cache;
  if (part == 2) a.alloc<char>(dec_cache_need(d));
  if (ns == 1) {
    a.cache = &cache;
    // operands the forward pass already staged as bf16 planes (same values: frame scale / no dropout included)
    if (!frame_scale && !given) cache.put(vid, V, BN, V, w.x_a);
    cache.put(w.enc, H, BN, H, w.enc_a);
    cache.put(s_in, -1, BL, E, w.emb_a);
  }
  // transposed weights of the sweeps: already staged by the forward call on this workspace, or staged here
  if (part != 2 && !side_note_take(ws, NOTE_ATT_BWD_WEIGHTS, p.dec_w_hh)) PVCR_TRY(att_bwd_weights(d, p, w, st));
  if (part != 2) {
  if (ns > 1) {     // bf16 mode multiplies by the forward weight planes directly (MN-major operand)
    PVCR_TRY(prep_weight_T(p.att_wk, H, H, H, w.wkT, 0, 1, st));
    PVCR_TRY(prep_weight_T(p.dec_w_ih + H, H + E, H3, E, w.weT, 0, 1, st));
  }
  if (need_frame_grad && ns > 1) PVCR_TRY(prep_weight_T(p.enc_w_ih, V, H3, V, w.wih_encT, 0, 1, st));

  const H0 h0 = initial_state(d, w, given);
  const bool persist_dec = dec_persist_eligible(B, N, H, ns, w.enc_a.Kp);
  Planes dgi_p{}, d1_p{};
  bool sweep_planes = false, emb_zeroed = false;
  if (persist_dec) {
    DecPersistBwd q{};
    q.L = L; q.B = B; q.N = N; q.H = H;
    q.wcT = w.wcT.ptr; q.wcT_ld = w.wcT.ld; q.wcatT = w.wcatT.ptr; q.wcatT_ld = w.wcatT.ld;
    q.v = p.att_v; q.pk = w.pk; q.enc_a = w.enc_a.ptr; q.enc_ld = w.enc_a.ld; q.enc = w.enc; q.hs = hs; q.d_hs = d_hs;
    q.h0 = h0.f; q.h0_ld = h0.f_ld;
    q.q_all = w.g1_all; q.q_ld = H4; q.alpha = w.alpha_all;
    q.r = w.dr; q.z = w.dz; q.n = w.dn; q.ghn = w.dghn;
    q.dgi_all = w.dgi_all; q.d1_all = w.d1_all; q.dctx_all = w.dctx_all; q.ds_all = w.ds_all;
    q.dh_carry = w.dh_carry; q.xg = w.xg; q.counters = w.sync;
    if (ns == 1) {          // the sweep emits the bf16 operand planes of the hoisted weight-gradient GEMMs itself
      dgi_p = alloc_planes(a, BL, H3, 1); d1_p = alloc_planes(a, BL, H4, 1);
      if (a.failed) { set_last_error("s2vtatt_bwd: workspace too small (decoder gradient planes)"); return PVCR_ERR_WORKSPACE; }
      q.dgi_p = dgi_p.ptr; q.dgi_p_ld = dgi_p.ld; q.d1_p = d1_p.ptr; q.d1_p_ld = d1_p.ld;
      cache.put(w.dgi_all, H3, BL, H3, dgi_p);
      sweep_planes = true;
    }
    // the dense embedding gradient is zeroed on lane 1 in the shadow of the sweep (nothing there depends on the sweep)
    static const bool emb_early = getenv("PVCR_NO_EMB_EARLY_ZERO") == nullptr;      // A/B knob
    if (ns == 1 && side_site(2) && emb_early) {
      cudaStream_t lz;
      PVCR_TRY(side_fork(st, &lz, 1));
      PVCR_TRY(fill_zero(g.emb, sizeof(float) * (size_t)d.Vc * E, lz));
      emb_zeroed = true;
    }
    PVCR_TRY(dec_persist_bwd(q, st));
  } else {
  PVCR_TRY(fill_zero(w.dh_carry, sizeof(float) * (size_t)B * H, st));
  PVCR_TRY(fill_zero(w.dpk, sizeof(float) * (size_t)BN * H, st));
  PVCR_TRY(fill_zero(w.denc, sizeof(float) * (size_t)BN * H, st));
  PVCR_TRY(fill_zero(w.dv_part, sizeof(float) * (size_t)B * H, st));
  if (w.d1_a.Kp != H4 || w.dgi_a.Kp != H3) {     // contraction padding must read as zero
    PVCR_TRY(fill_zero(w.d1_a.ptr, sizeof(bf16) * (size_t)B * w.d1_a.ld, st));
    PVCR_TRY(fill_zero(w.dgi_a.ptr, sizeof(bf16) * (size_t)B * w.dgi_a.ld, st));
    PVCR_TRY(fill_zero(w.dgh_a.ptr, sizeof(bf16) * (size_t)B * w.dgh_a.ld, st));
  }
  // ---- decoder, reverse time ----
  for (int i = L - 1; i >= 0; --i) {
    GruBwdArgs b{};
    b.B = B; b.H = H;
    b.dh_a = w.dh_carry; b.dh_a_ld = H;
    b.dh_b = d_hs + (long long)i * H; b.dh_b_ld = (long long)L * H;
    const long long o = (long long)i * B * H;
    b.r = w.dr + o; b.z = w.dz + o; b.n = w.dn + o; b.ghn = w.dghn + o;
    if (i == 0) { b.h_prev = h0.f; b.h_prev_ld = h0.f_ld; }
    else { b.h_prev = hs + (long long)(i - 1) * H; b.h_prev_ld = (long long)L * H; }
    b.dgi = w.dgi_all + (long long)i * H3; b.dgi_ld = (long long)L * H3;           // rows b*L + i
    b.dgh = w.d1_all + (long long)i * H4 + H; b.dgh_ld = (long long)L * H4;
    b.dgi_planes = w.dgi_a.ptr; b.dgi_planes_ld = w.dgi_a.ld; b.dgi_Kp = w.dgi_a.Kp; b.dgi_col0 = 0;
    b.dgh_planes = w.d1_a.ptr; b.dgh_planes_ld = w.d1_a.ld; b.dgh_Kp = w.d1_a.Kp; b.dgh_col0 = H;
    b.nsplit = ns;
    b.dh_direct = w.dh_carry; b.dh_direct_ld = H;
    PVCR_TRY(gru_gate_bwd(b, st));
    // dctx = dgi Wc
    PVCR_TRY(gemm_planes(w.dgi_a.view(), w.wcT.view(), B, H, (int)w.dgi_a.ld, w.dctx, H, nullptr, 0, st));
    AttnBwdArgs at{};
    at.B = B; at.N = N; at.H = H;
    at.dctx = w.dctx; at.dctx_ld = H;
    at.q = w.g1_all + (long long)i * B * H4; at.q_ld = H4;
    at.pk = w.pk; at.enc = w.enc; at.v = p.att_v; at.alpha = w.alpha_all + (long long)i * B * N;
    at.dq = w.d1_all + (long long)i * H4; at.dq_ld = (long long)L * H4;
    at.dq_planes = w.d1_a.ptr; at.dq_planes_ld = w.d1_a.ld; at.dq_Kp = w.d1_a.Kp; at.nsplit = ns;
    at.dpk = w.dpk; at.denc = w.denc; at.dv_part = w.dv_part;
    PVCR_TRY(attn_bwd(at, st));
    // dh_{i-1} = dh*z + [dq | dgh] [Wq ; Whh]
    PVCR_TRY(gemm_planes(w.d1_a.view(), w.wcatT.view(), B, H, (int)w.d1_a.ld, w.dh_carry, H, nullptr, 1, st));
  }
  }

  // ---- decoder weight gradients, hoisted over all (b, i) rows ----
  // None of them feeds the encoder half: in bf16 mode they run on the side lane (every operand is cast once into the
  // staging cache, so no transient scratch is shared between the two streams) while this stream goes on with
  // d proj_key -> d enc -> encoder sweep.  Persistent GEMMs on the lane can be capped to the SMs a sweep leaves free (PVCR_SIDE_CAP; uncapped measured fastest).
  const bool fork = persist_dec && ns == 1 && side_site(2);
  static const int side_cap = getenv("PVCR_SIDE_CAP") ? atoi(getenv("PVCR_SIDE_CAP")) : 0;   // measured: 0 (no cap) fastest
  // lane A: gradients fed by [dq | dgh] and h_{i-1};  lane B: gradients fed by dgi;  lane C: d v, d W_k.  No staged
  // operand is shared between lanes (d proj_key, which this stream needs too, is staged here before lane C forks).
  cudaStream_t la = st, lb = st, lc = st;
  if (fork) { PVCR_TRY(side_fork(st, &la, 0)); PVCR_TRY(side_fork(st, &lb, 1)); }          // after the sweep
  {
  CtaCap cap_(fork && side_mode() == 2 ? side_cap : 0);
  if (sweep_planes) {
    // dW = sum_{b,i} d[b,i]^T h_{i-1}[b]:  rows i >= 1 pair with the forward's hs planes shifted by one row (the sweep
    // wrote the i = 0 rows of [dq | dgh] as zeros), rows i = 0 pair with the encoder's final state: no h_{i-1} copy
    const OperandView dq1{d1_p.ptr + d1_p.ld, d1_p.ld, 0, BL - 1, 1}, dgh1{d1_p.ptr + d1_p.ld + H, d1_p.ld, 0, BL - 1, 1};
    const OperandView hsm{w.hs_a.ptr, w.hs_a.ld, 0, BL - 1, 1};
    if (BL > 1) {
      PVCR_TRY(gemm_mn_store(dgh1, hsm, H3, H, BL - 1, g.dec_w_hh, H, 0, la));
      PVCR_TRY(gemm_mn_store(dq1, hsm, H, H, BL - 1, g.att_wq, H, 0, la));
    } else {
      PVCR_TRY(fill_zero(g.dec_w_hh, sizeof(float) * (size_t)H3 * H, la));
      PVCR_TRY(fill_zero(g.att_wq, sizeof(float) * (size_t)H * H, la));
    }
    Planes d0 = alloc_planes(a, B, H4, 1);                   // [dq | dgh] of step 0, rows b
    if (a.failed) { set_last_error("s2vtatt_bwd: workspace too small (step-0 planes)"); return PVCR_ERR_WORKSPACE; }
    PVCR_TRY(cast_split(w.d1_all, (long long)L * H4, B, H4, d0.ptr, d0.ld, d0.Kp, 1, 0, nullptr, NO_DROPOUT, la));
    const OperandView e_last{const_cast<bf16*>(h0.a), h0.a_ld, 0, B, 1};
    PVCR_TRY(gemm_mn_store(OperandView{d0.ptr + H, d0.ld, 0, B, 1}, e_last, H3, H, B, g.dec_w_hh, H, 1, la));
    PVCR_TRY(gemm_mn_store(OperandView{d0.ptr, d0.ld, 0, B, 1}, e_last, H, H, B, g.att_wq, H, 1, la));
  } else {
  // h_{i-1} rows in (b, i) order: i = 0 -> encoder final state, i >= 1 -> hs[b, i-1]
  PVCR_CUDA_CHECK(cudaMemcpy2DAsync(w.hprev_dec, sizeof(float) * (size_t)L * H, h0.f, sizeof(float) * (size_t)h0.f_ld,
                                    sizeof(float) * H, B, cudaMemcpyDeviceToDevice, la));
  if (L > 1)
    PVCR_CUDA_CHECK(cudaMemcpy2DAsync(w.hprev_dec + H, sizeof(float) * (size_t)L * H, hs, sizeof(float) * (size_t)L * H,
                                      sizeof(float) * (size_t)(L - 1) * H, B, cudaMemcpyDeviceToDevice, la));
  PVCR_TRY(grad_w(a, w.d1_all + H, H4, BL, H3, w.hprev_dec, H, H, nullptr, nullptr, g.dec_w_hh, H, 0, ns, la));
  PVCR_TRY(grad_w(a, w.d1_all, H4, BL, H, w.hprev_dec, H, H, nullptr, nullptr, g.att_wq, H, 0, ns, la));
  }
  // lane B: the embedding gradient FIRST (the largest decoder-side gradient, 27.6 MB at cfg2: a data-parallel caller
  // starts its all-reduce at the milestone below, under the encoder sweep), then the two halves of d W_ih and d b_ih
  // PVCR_EMB_FIRST=1: embedding gradient before the two halves of d W_ih on this lane.  Measured on one GPU: +16 us per
  // step (2.160 vs 2.144 ms), and no gain at 2 GPUs from the earlier all-reduce start -- off.
  static const bool emb_first = getenv("PVCR_EMB_FIRST") != nullptr;
  if (!emb_first) {
    PVCR_TRY(grad_w(a, w.dgi_all, H3, BL, H3, w.ctx_all, H, H, nullptr, nullptr, g.dec_w_ih, H + E, 0, ns, lb));
    PVCR_TRY(grad_w(a, w.dgi_all, H3, BL, H3, p.emb, E, E, s_in, nullptr, g.dec_w_ih + H, H + E, 0, ns, lb));
  }
  if (ns == 1) PVCR_TRY(grad_x_fwdw(a, w.dgi_all, H3, BL, H3, w.we, E, w.demb_rows, E, 0, lb));
  else PVCR_TRY(grad_x(a, w.dgi_all, H3, BL, H3, w.weT, w.demb_rows, E, 0, lb));
  if (!emb_zeroed) PVCR_TRY(fill_zero(g.emb, sizeof(float) * (size_t)d.Vc * E, lb));
  PVCR_TRY(scatter_add_rows(w.demb_rows, E, s_in, BL, E, g.emb, NO_DROPOUT, lb));
  if (fork) PVCR_TRY(side_milestone(0, lb));
  if (emb_first) {
    PVCR_TRY(grad_w(a, w.dgi_all, H3, BL, H3, w.ctx_all, H, H, nullptr, nullptr, g.dec_w_ih, H + E, 0, ns, lb));
    PVCR_TRY(grad_w(a, w.dgi_all, H3, BL, H3, p.emb, E, E, s_in, nullptr, g.dec_w_ih + H, H + E, 0, ns, lb));
  }
  PVCR_TRY(colsum(w.dgi_all, H3, BL, H3, g.dec_b_ih, 0, lb));
  if (persist_dec) {
    AttnGradArgs ag{};
    ag.L = L; ag.B = B; ag.N = N; ag.H = H;
    ag.alpha = w.alpha_all; ag.ds = w.ds_all; ag.dctx = w.dctx_all; ag.q = w.g1_all; ag.q_ld = H4;
    ag.pk = w.pk; ag.v = p.att_v; ag.dpk = w.dpk; ag.denc = w.denc; ag.dv_part = w.dv_part;
    if (ns == 1) {
      // d proj_key feeds two GEMMs (d W_k on a lane, d enc here): the kernel emits its bf16 operand planes itself
      Planes dpk_a = alloc_planes(a, BN, H, 1);
      if (a.failed) { set_last_error("s2vtatt_bwd: workspace too small (dpk planes)"); return PVCR_ERR_WORKSPACE; }
      ag.dpk_a = dpk_a.ptr; ag.dpk_a_ld = dpk_a.ld;
      cache.put(w.dpk, H, BN, H, dpk_a);
    }
    PVCR_TRY(attn_grad_hoisted(ag, st));
  }
  if (fork) PVCR_TRY(side_fork(st, &lc, 2));
  PVCR_TRY(colsum(w.dv_part, H, B, H, g.att_v, 0, lc));
  PVCR_TRY(colsum(w.d1_all + H, H4, BL, H3, g.dec_b_hh, 0, lc));
  // key projection: dWk = dpk^T enc ; denc += dpk Wk
  PVCR_TRY(grad_w(a, w.dpk, H, BN, H, w.enc, H, H, nullptr, nullptr, g.att_wk, H, 0, ns, lc));
  }
  if (ns == 1) PVCR_TRY(grad_x_fwdw(a, w.dpk, H, BN, H, w.wk, H, w.denc, H, 1, st));
  else PVCR_TRY(grad_x(a, w.dpk, H, BN, H, w.wkT, w.denc, H, 1, st));

  }   // decoder half
  if (given) {      // decode(): hand the gradients on the caller's encoder outputs / initial state back
    PVCR_CUDA_CHECK(cudaMemcpyAsync(d_enc_outs, w.denc, sizeof(float) * (size_t)BN * H, cudaMemcpyDeviceToDevice, st));
    PVCR_CUDA_CHECK(cudaMemcpyAsync(d_enc_final, w.dh_carry, sizeof(float) * (size_t)B * H, cudaMemcpyDeviceToDevice, st));
    return side_call_end(st);
  }
  if (part == 1) return side_call_end(st);
  // ---- encoder, reverse time (dh_carry already holds the gradient on the final state) ----
  GruSeq es = encoder_seq(d, p, w);
  GruSeqGrad eg{};
  eg.dh_ext = w.denc; eg.dh_ext_ts = H; eg.dh_ext_ld = (long long)N * H;
  eg.dh_carry = w.dh_carry;
  eg.dgi = w.dgi_enc; eg.dgi_ts = H3; eg.dgi_ld = (long long)N * H3;
  eg.dgh = w.dgh_enc; eg.dgh_ts = H3; eg.dgh_ld = (long long)N * H3;
  eg.dgh_a = w.dgh_a; eg.whhT = w.whh_encT; eg.xch = w.xch;
  const bool enc_planes = ns == 1 && gru_persist_eligible(es);
  if (enc_planes) {     // the persistent sweep emits the bf16 operand planes of the two weight-gradient GEMMs itself
    Planes dgi_p = alloc_planes(a, BN, H3, 1), dgh_p = alloc_planes(a, BN, H3, 1);
    if (a.failed) { set_last_error("s2vtatt_bwd: workspace too small (encoder gradient planes)"); return PVCR_ERR_WORKSPACE; }
    eg.dgi_p = dgi_p.ptr; eg.dgi_p_ts = dgi_p.ld; eg.dgi_p_ld = (long long)N * dgi_p.ld;
    eg.dgh_p = dgh_p.ptr; eg.dgh_p_ts = dgh_p.ld; eg.dgh_p_ld = (long long)N * dgh_p.ld;
    cache.put(w.dgi_enc, H3, BN, H3, dgi_p);
  }
  PVCR_TRY(gru_seq_bwd(es, eg, st));
  // the two encoder weight gradients are independent: W_hh and the bias column sums on side lanes, W_ih here
  const bool fork2 = ns == 1 && side_site(3);
  cudaStream_t ln2 = st, ln3 = st, ln4 = st;
  if (fork2) { PVCR_TRY(side_fork(st, &ln2, 2)); PVCR_TRY(side_fork(st, &ln3, 1)); PVCR_TRY(side_fork(st, &ln4, 0)); }
  if (enc_planes) {
    // dW_hh = sum_{b,t} dgh[b,t]^T h_{t-1}[b]: the sweep wrote the t = 0 rows of the dgh planes as zeros (h_{-1} = 0), so
    // the product runs on the forward's enc planes shifted by one row; no h_{t-1} copy
    if (BN > 1)
      PVCR_TRY(gemm_mn_store(OperandView{eg.dgh_p + eg.dgh_p_ts, eg.dgh_p_ts, 0, BN - 1, 1},
                             OperandView{w.enc_a.ptr, w.enc_a.ld, 0, BN - 1, 1}, H3, H, BN - 1, g.enc_w_hh, H, 0, ln2));
    else PVCR_TRY(fill_zero(g.enc_w_hh, sizeof(float) * (size_t)H3 * H, ln2));
  } else {
  // h_{t-1} rows in (b, t) order: zero for t = 0
  PVCR_TRY(fill_zero(w.hprev_enc, sizeof(float) * (size_t)BN * H, ln2));
  if (N > 1)
    PVCR_CUDA_CHECK(cudaMemcpy2DAsync(w.hprev_enc + H, sizeof(float) * (size_t)N * H, w.enc, sizeof(float) * (size_t)N * H,
                                      sizeof(float) * (size_t)(N - 1) * H, B, cudaMemcpyDeviceToDevice, ln2));
  PVCR_TRY(grad_w(a, w.dgh_enc, H3, BN, H3, w.hprev_enc, H, H, nullptr, nullptr, g.enc_w_hh, H, 0, ns, ln2));
  }
  PVCR_TRY(colsum(w.dgh_enc, H3, BN, H3, g.enc_b_hh, 0, ln4));
  PVCR_TRY(colsum(w.dgi_enc, H3, BN, H3, g.enc_b_ih, 0, ln3));
  if (ns == 1 && frame_scale) cache.put(vid, V, BN, V, w.x_a);      // x_a = vid * frame_scale, exactly this operand
  PVCR_TRY(grad_w(a, w.dgi_enc, H3, BN, H3, vid, V, V, nullptr, frame_scale, g.enc_w_ih, V, 0, ns, st));
  if (need_frame_grad) {
    // d(sel) = dgi W_ih ; d frame_scale[b,n] = sum_v vid[b,n,v] * dsel[b,n,v]   (model/RationaleNet.py:52)
    if (ns == 1) PVCR_TRY(grad_x_fwdw(a, w.dgi_enc, H3, BN, H3, w.wih_enc, V, w.dxsel, V, 0, st));
    else PVCR_TRY(grad_x(a, w.dgi_enc, H3, BN, H3, w.wih_encT, w.dxsel, V, 0, st));
    PVCR_TRY(rowdot(vid, w.dxsel, BN, V, d_frame_scale, st));
  }
  return side_call_end(st);
}

int s2vtatt_bwd(const PvcrDims& d, const PvcrS2vtAttParams& p, const float* vid, const float* frame_scale,
                const long long* s_in, const float* d_hs, const float* hs, PvcrS2vtAttGrads& g, float* d_frame_scale,
                void* ws, size_t ws_bytes, cudaStream_t st, int part) {
  return s2vtatt_bwd_impl(d, p, vid, frame_scale, s_in, d_hs, hs, g, d_frame_scale, nullptr, nullptr, ws, ws_bytes, st, part);
}
// Backward of s2vtatt_decode_fwd: every decoder / attention / embedding gradient of `g` (its encoder entries are not
// touched), d_enc_outs [B,N,H] and d_enc_final [B,H].
int s2vtatt_decode_bwd(const PvcrDims& d, const PvcrS2vtAttParams& p, const long long* s_in, const float* d_hs,
                       const float* hs, PvcrS2vtAttGrads& g, float* d_enc_outs, float* d_enc_final, void* ws,
                       size_t ws_bytes, cudaStream_t st) {
  PVCR_REQUIRE(d_enc_outs && d_enc_final, "s2vtatt_decode_bwd: null gradient outputs");
  return s2vtatt_bwd_impl(d, p, nullptr, nullptr, s_in, d_hs, hs, g, nullptr, d_enc_outs, d_enc_final, ws, ws_bytes, st, 1);
}

// ---- fixed-length greedy decoding (eval branch, model/S2VTAttModel.py:172-191) ---------------------------------
struct GreedyWs {
  AttWs w;
  Planes wv, emb_all, wcat_c, wc_c;      // weights of the per-step products: each split term stored once (compact planes)
  bf16* wv_blocked;                      // W_v planes again, K-blocked (block_planes): what the per-step projection streams
  float *logits_step, *hs, *emb_table, *pc;
  long long* words;
  void* argmax_scratch;
};
static void carve_greedy(Arena& a, const PvcrDims& d, GreedyWs& g) {
  carve(a, d, 0, g.w);
  g.wv = alloc_planes_compact(a, d.Vc, d.H, d.nsplit);           // each term of W_v once: the planes stay L2-resident between steps
  g.wv_blocked = a.alloc<bf16>((size_t)d.Vc * g.wv.ld);
  g.wcat_c = alloc_planes_compact(a, 4 * d.H, d.H, d.nsplit);
  g.wc_c = alloc_planes_compact(a, 3 * d.H, d.H, d.nsplit);
  g.emb_all = alloc_planes(a, d.Vc, d.E, d.nsplit);              // the whole embedding table as an A-role operand (prepare only)
  g.emb_table = a.alloc<float>((size_t)d.Vc * 3 * d.H);           // W_e Emb[w] + b_ih for EVERY word w
  g.logits_step = a.alloc<float>((size_t)d.B * round_up(d.Vc, 4));
  g.hs = a.alloc<float>((size_t)d.B * d.L * d.H);
  g.words = a.alloc<long long>(d.B);
  g.argmax_scratch = a.alloc<char>(gemm_argmax_scratch(d.B, d.Vc));
  g.pc = a.alloc<float>((size_t)d.B * d.N * 3 * d.H);            // enc W_c^T: the context projection of every frame
}
size_t s2vtatt_greedy_workspace(const PvcrDims& d) {
  Arena a(nullptr, 0);
  GreedyWs g;
  carve_greedy(a, d, g);
  return a.off + 4096;
}

// enc_given / final_given: decode() in eval mode (model/S2VTAttModel.py:231-243 with self.training == False).
int s2vtatt_greedy_impl(const PvcrDims& d, const PvcrS2vtAttParams& p, const float* vid, const float* frame_scale,
                   const float* enc_given, const float* final_given,
                   long long sos_id, long long* ids, float* logits, float* alphas, void* ws, size_t ws_bytes,
                   cudaStream_t st, int flags) {
  PVCR_TRY(check_dims(d));
  const bool given = enc_given != nullptr;
  PVCR_REQUIRE(given == (final_given != nullptr), "s2vtatt greedy decode: encoder outputs and final state come together");
  const int B = d.B, N = d.N, V = d.V, H = d.H, E = d.E, L = d.L, Vc = d.Vc;
  const int BN = B * N, BL = B * L, H3 = 3 * H, H4 = 4 * H;
  Arena a(ws, ws_bytes);
  GreedyWs gw;
  carve_greedy(a, d, gw);
  if (a.failed) { set_last_error("s2vtatt_greedy: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  AttWs& w = gw.w;
  // Parameter-only work: the weight planes and the word table  T[w] = W_e Emb[w] + b_ih  (decoder input projection of
  // EVERY vocabulary word, one [Vc,E] x [E,3H] GEMM) -- a decoding step then reads row T[argmax] instead of gathering the
  // embedding and multiplying it.  A caller that decodes batch after batch with unchanged parameters keeps the workspace
  // and passes PVCR_DECODE_REUSE_PREPARED: everything in this block is skipped.
  if (!(flags & PVCR_DECODE_REUSE_PREPARED)) {
  if (!given) {
  PVCR_TRY(prep_weight(p.enc_w_ih, V, H3, V, w.wih_enc, st));
  PVCR_TRY(prep_weight(p.enc_w_hh, H, H3, H, w.whh_enc, st));
  }
  PVCR_TRY(prep_weight(p.att_wk, H, H, H, w.wk, st));
  PVCR_TRY(prep_weight(p.att_wq, H, H, H, gw.wcat_c, st, 0));
  PVCR_TRY(prep_weight(p.dec_w_hh, H, H3, H, gw.wcat_c, st, H));
  PVCR_TRY(prep_weight(p.dec_w_ih, H + E, H3, H, gw.wc_c, st));
  PVCR_TRY(prep_weight(p.dec_w_ih + H, H + E, H3, E, w.we, st));
  PVCR_TRY(prep_weight(p.out_w, H, Vc, H, gw.wv, st));
  PVCR_TRY(block_planes(gw.wv.ptr, gw.wv.ld, Vc, gw.wv_blocked, st));
  if (gw.emb_all.Kp != E) PVCR_TRY(fill_zero(gw.emb_all.ptr, sizeof(bf16) * (size_t)Vc * gw.emb_all.ld, st));
  PVCR_TRY(stage(p.emb, E, Vc, E, gw.emb_all, 0, nullptr, NO_DROPOUT, st));
  PVCR_TRY(gemm_planes(gw.emb_all.view(), w.we.view(), Vc, H3, (int)gw.emb_all.ld, gw.emb_table, H3, p.dec_b_ih, 0, st));
  }
  if (w.enc_a.Kp != H) {
    PVCR_TRY(fill_zero(w.enc_a.ptr, sizeof(bf16) * (size_t)BN * w.enc_a.ld, st));
    PVCR_TRY(fill_zero(w.hs_a.ptr, sizeof(bf16) * (size_t)BL * w.hs_a.ld, st));
    PVCR_TRY(fill_zero(w.ctx_a.ptr, sizeof(bf16) * (size_t)B * w.ctx_a.ld, st));
  }
  if (given) {
    PVCR_CUDA_CHECK(cudaMemcpyAsync(w.enc, enc_given, sizeof(float) * (size_t)BN * H, cudaMemcpyDeviceToDevice, st));
    PVCR_CUDA_CHECK(cudaMemcpyAsync(w.h0_f, final_given, sizeof(float) * (size_t)B * H, cudaMemcpyDeviceToDevice, st));
    if (w.h0_a.Kp != H) PVCR_TRY(fill_zero(w.h0_a.ptr, sizeof(bf16) * (size_t)B * w.h0_a.ld, st));
    PVCR_TRY(stage(enc_given, H, BN, H, w.enc_a, 0, nullptr, NO_DROPOUT, st));
    PVCR_TRY(stage(final_given, H, B, H, w.h0_a, 0, nullptr, NO_DROPOUT, st));
  } else {
  PVCR_TRY(stage(vid, V, BN, V, w.x_a, 0, frame_scale, NO_DROPOUT, st));
  PVCR_TRY(gemm_planes(w.x_a.view(), w.wih_enc.view(), BN, H3, (int)w.x_a.ld, w.gi_enc, H3, p.enc_b_ih, 0, st));
  if (gru_f32_persist_eligible(B, H)) {
    // the N encoder steps in one launch, exact fp32 on the CUDA cores (gru_f32_persist.cu)
    GruF32Fwd q{};
    q.T = N; q.B = B; q.H = H;
    q.gi = w.gi_enc; q.gi_ts = H3; q.gi_ld = (long long)N * H3;
    q.w_hh = p.enc_w_hh; q.b_hh = p.enc_b_hh;
    q.h = w.enc; q.h_ts = H; q.h_ld = (long long)N * H;
    q.counters = w.sync;
    PVCR_TRY(gru_f32_persist_fwd(q, st));
    PVCR_TRY(stage(w.enc, H, BN, H, w.enc_a, 0, nullptr, NO_DROPOUT, st));      // split planes of all N outputs in one pass
  } else {
    PVCR_TRY(gru_seq_fwd(encoder_seq(d, p, w), st));
  }
  }
  const H0 h0 = initial_state(d, w, given);
  PVCR_TRY(gemm_planes(w.enc_a.view(), w.wk.view(), BN, H, (int)w.enc_a.ld, w.pk, H, nullptr, 0, st));
  PVCR_TRY(fill_i64(gw.words, sos_id, B, st));
  // W_c ctx = sum_n alpha_n (W_c enc_n): the context projection is hoisted to ONE [B N, H] x [H, 3H] GEMM per batch and the
  // attention kernel sums the projected frames -- no context vector, no per-step GEMM behind the attention
  static const bool pc_off = getenv("PVCR_NO_DECODE_PC") != nullptr;                 // A/B knob
  // (only while the projected frames, B N 3H fp32, stay L2-resident next to the keys: 31 MB at B = 128; at B = 512 they would
  // be re-read from HBM every step: 80 k captions/s instead of 81 k without, 90 k instead of 103 k at B = 1024)
  const bool use_pc = !pc_off && attn_fwd_projected_ok(N, H, H3) && (size_t)B * N * H3 * sizeof(float) <= ((size_t)40 << 20);
  if (use_pc) PVCR_TRY(gemm_planes(w.enc_a.view(), gw.wc_c.view(), BN, H3, (int)w.enc_a.ld, gw.pc, H3, nullptr, 0, st));
  float* hs = gw.hs;
  // Step i: [q | gh] = [W_q; W_hh] h_{i-1} -> attention -> W_c ctx   (needs h_{i-1} only)
  //         gates with T[word_i] -> h_i -> logits_i -> word_{i+1}    (the only place the fed-back word enters)
  // so the first half of step i+1 does not wait for the vocabulary projection and arg-max of step i: it runs on a side
  // lane next to them and the two meet again at the gates of step i+1.
  // The chain gates -> [q | gh] product -> attention is launched programmatically (PdlScope): each kernel's launch latency
  // and prologue overlap the tail of its predecessor.  A/B knob: PVCR_NO_DECODE_PDL.
  static const bool pdl_off = getenv("PVCR_NO_DECODE_PDL") != nullptr;
  // A handful of videos (B <= 4): the per-step products are matrix-vector products, taken from the fp32 parameters on the
  // CUDA cores (gemv_f32.cu: 4 bytes per weight and step instead of three bf16 terms through 128-row tiles).  Knob: PVCR_NO_DECODE_GEMV.
  static const bool gemv_off = getenv("PVCR_NO_DECODE_GEMV") != nullptr;
  const bool small = !gemv_off && gemv_f32_eligible(B, H) && h0.f_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(h0.f) & 15) == 0;
  auto recurrent_half = [&](int i, cudaStream_t s) -> int {
    PdlScope pdl(!pdl_off && i > 0 && !small);
    OperandView hprev_a = (i == 0)
        ? OperandView{const_cast<bf16*>(h0.a), h0.a_ld, 0, B, 1}
        : OperandView{w.hs_a.ptr + (long long)(i - 1) * w.hs_a.ld, (long long)L * w.hs_a.ld, 0, B, 1};
    if (small) {
      GemvF32 q{};
      q.B = B; q.K = H;
      q.x = i == 0 ? h0.f : gw.hs + (long long)(i - 1) * H; q.x_ld = i == 0 ? h0.f_ld : (long long)L * H;
      q.w0 = p.att_wq; q.w0_ld = H; q.rows0 = H;
      q.w1 = p.dec_w_hh; q.w1_ld = H; q.rows1 = H3;
      q.out = w.g1_all; q.out_ld = H4;
      PVCR_TRY(gemv_f32(q, s));
    } else {
      PVCR_TRY(gemm_planes(hprev_a, gw.wcat_c.view(), B, H4, (int)w.hs_a.ld, w.g1_all, H4, nullptr, 0, s));
    }
    AttnFwdArgs at{};
    at.B = B; at.N = N; at.H = H;
    at.q = w.g1_all; at.q_ld = H4; at.pk = w.pk; at.enc = w.enc; at.v = p.att_v;
    at.alpha = alphas ? alphas + (long long)i * B * N : w.alpha_all;
    at.ctx = w.ctx_all; at.ctx_ld = H;
    at.ctx_planes = w.ctx_a.ptr; at.ctx_planes_ld = w.ctx_a.ld; at.Hp = w.ctx_a.Kp; at.nsplit = d.nsplit;
    if (use_pc) { at.val = gw.pc; at.W = H3; at.out = w.g2; at.out_ld = H3; }
    PVCR_TRY(attn_fwd(at, s));
    if (!use_pc) PVCR_TRY(gemm_planes(w.ctx_a.view(), gw.wc_c.view(), B, H3, (int)w.ctx_a.ld, w.g2, H3, nullptr, 0, s));
    return PVCR_OK;
  };
  static const bool overlap_off = getenv("PVCR_NO_DECODE_OVERLAP") != nullptr;       // A/B knob
  // The (max, index) partials of step i's vocabulary GEMM are combined by the gate kernel of step i+1 (which is where the
  // fed-back word is needed): one launch less on the critical path of every step.  A/B knob: PVCR_NO_DECODE_FOLD_ARGMAX.
  static const bool fold_off = getenv("PVCR_NO_DECODE_FOLD_ARGMAX") != nullptr;
  const bool fold = !fold_off && H >= 32;
  ArgmaxParts parts{};
  PVCR_TRY(recurrent_half(0, st));
  for (int i = 0; i < L; ++i) {
    float* g1 = w.g1_all;
    GruFwdArgs g{};
    g.B = B; g.H = H;
    g.gi_a = w.g2; g.gi_a_ld = H3;
    g.gi_b = gw.emb_table; g.gi_b_ld = H3; g.gi_b_rows = gw.words;      // + W_e Emb[word] + b_ih
    if (fold && i > 0) {                                                // word_i = arg-max of step i-1, combined here
      g.am_pmax = parts.pmax; g.am_pidx = parts.pidx; g.am_nparts = parts.nparts;
      g.am_out = ids + (i - 1); g.am_out_stride = L;
    }
    g.gh = g1 + H; g.gh_ld = H4; g.b_hh = p.dec_b_hh;
    if (i == 0) { g.h_prev = h0.f; g.h_prev_ld = h0.f_ld; }
    else { g.h_prev = hs + (long long)(i - 1) * H; g.h_prev_ld = (long long)L * H; }
    g.h_out = hs + (long long)i * H; g.h_out_ld = (long long)L * H;
    g.h_planes = w.hs_a.ptr + (long long)i * w.hs_a.ld; g.h_planes_ld = (long long)L * w.hs_a.ld;
    g.Hp = w.hs_a.Kp; g.nsplit = d.nsplit;
    PVCR_TRY(gru_gate_fwd(g, st));
    // vocabulary projection + arg-max of step i on a side lane; the first half of step i+1 stays on the caller's stream
    // (which of the two halves gets the lane makes no measurable difference: 2.888 vs 2.890 ms per batch)
    cudaStream_t lane = st;
    if (!overlap_off && side_site(3) && i + 1 < L) PVCR_TRY(side_fork(st, &lane, 0));
    OperandView h_a{w.hs_a.ptr + (long long)i * w.hs_a.ld, (long long)L * w.hs_a.ld, 0, B, 1};
    // logits_i and word_{i+1} in one pass: the arg-max is taken in the GEMM epilogue (the logits are stored only when the
    // caller asked for them and never read back)
    if (i + 1 < L) PVCR_TRY(recurrent_half(i + 1, st));
    const bool combine_here = !fold || i + 1 == L;
    OperandView wv_v = gw.wv.view();
    wv_v.blocked = gw.wv_blocked;
    if (small) {
      GemvF32 q{};
      q.B = B; q.K = H;
      q.x = hs + (long long)i * H; q.x_ld = (long long)L * H;
      q.w0 = p.out_w; q.w0_ld = H; q.rows0 = Vc;
      q.bias = p.out_b;
      q.out = logits ? logits + (long long)i * Vc : nullptr; q.out_ld = (long long)L * Vc;
      const int np = gemv_f32_parts(Vc);
      q.pmax = reinterpret_cast<float*>(gw.argmax_scratch); q.pidx = reinterpret_cast<int*>(q.pmax + (size_t)B * np);
      q.stream = 1;
      PVCR_TRY(gemv_f32(q, lane));
      parts.pmax = q.pmax; parts.pidx = q.pidx; parts.nparts = np;
      if (combine_here) PVCR_TRY(argmax_combine(q.pmax, q.pidx, B, np, ids + i, L, gw.words, lane));
    } else
    PVCR_TRY(gemm_argmax(h_a, wv_v, B, Vc, (int)w.hs_a.ld, p.out_b, logits ? logits + (long long)i * Vc : nullptr,
                         (long long)L * Vc, combine_here ? ids + i : nullptr, L, gw.words, gw.argmax_scratch, lane, &parts));
    if (lane != st) PVCR_TRY(side_join_lane(st, 0));
  }
  PVCR_TRY(side_call_end(st));
  return PVCR_OK;
}
int s2vtatt_greedy(const PvcrDims& d, const PvcrS2vtAttParams& p, const float* vid, const float* frame_scale,
                   long long sos_id, long long* ids, float* logits, float* alphas, void* ws, size_t ws_bytes,
                   cudaStream_t st) {
  return s2vtatt_greedy_impl(d, p, vid, frame_scale, nullptr, nullptr, sos_id, ids, logits, alphas, ws, ws_bytes, st, 0);
}

// ---- fixed-length beam search over the decoder step (definition: oracle s2vtatt_beam_search) --------
// Encoder and proj_key once per video, then K hypotheses per video share them (rows b*K + k of the replicated copies).
struct BeamWs {
  AttWs w;                    // encoder side, B rows
  Planes wv, emb_step, hp, ctx_p;
  float *encR, *pkR, *hA, *hB, *g1, *g2, *ctx, *alpha, *logits, *scoreA, *scoreB;
  long long *words, *histA, *histB;
  int* parent;
  char* sel_scratch;
  long long ldl;
};
static void carve_beam(Arena& a, const PvcrDims& d, int K, BeamWs& g) {
  carve(a, d, 0, g.w);
  const size_t R = (size_t)d.B * K, H = d.H, N = d.N;
  g.wv = alloc_planes(a, d.Vc, d.H, d.nsplit);
  g.emb_step = alloc_planes(a, (int)R, d.E, d.nsplit);
  g.hp = alloc_planes(a, (int)R, d.H, d.nsplit);
  g.ctx_p = alloc_planes(a, (int)R, d.H, d.nsplit);
  g.encR = a.alloc<float>(R * N * H); g.pkR = a.alloc<float>(R * N * H);
  g.hA = a.alloc<float>(R * H); g.hB = a.alloc<float>(R * H);
  g.g1 = a.alloc<float>(R * 4 * H); g.g2 = a.alloc<float>(R * 3 * H);
  g.ctx = a.alloc<float>(R * H); g.alpha = a.alloc<float>(R * N);
  g.ldl = round_up(d.Vc, 4);
  g.logits = a.alloc<float>(R * g.ldl);
  g.scoreA = a.alloc<float>(R); g.scoreB = a.alloc<float>(R);
  g.words = a.alloc<long long>(R); g.histA = a.alloc<long long>(R * d.L); g.histB = a.alloc<long long>(R * d.L);
  g.parent = a.alloc<int>(R);
  g.sel_scratch = a.alloc<char>(beam_select_scratch(d.B, K));
}
size_t s2vtatt_beam_workspace(const PvcrDims& d, int K) {
  Arena a(nullptr, 0);
  BeamWs g;
  carve_beam(a, d, K < 1 ? 1 : K, g);
  return a.off + 4096;
}

int s2vtatt_beam(const PvcrDims& d, const PvcrS2vtAttParams& p, const float* vid, const float* frame_scale, long long sos_id,
                 int K, long long* ids, float* scores, void* ws, size_t ws_bytes, cudaStream_t st) {
  PVCR_TRY(check_dims(d));
  PVCR_REQUIRE(K >= 1 && K <= 8, "s2vtatt_beam: beam width %d not in 1..8", K);
  const int B = d.B, N = d.N, V = d.V, H = d.H, E = d.E, L = d.L, Vc = d.Vc, R = B * K;
  const int BN = B * N, H3 = 3 * H, H4 = 4 * H;
  Arena a(ws, ws_bytes);
  BeamWs g;
  carve_beam(a, d, K, g);
  if (a.failed) { set_last_error("s2vtatt_beam: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  AttWs& w = g.w;
  PVCR_TRY(prep_weight(p.enc_w_ih, V, H3, V, w.wih_enc, st));
  PVCR_TRY(prep_weight(p.enc_w_hh, H, H3, H, w.whh_enc, st));
  PVCR_TRY(prep_weight(p.att_wk, H, H, H, w.wk, st));
  PVCR_TRY(prep_weight(p.att_wq, H, H, H, w.wcat, st, 0));
  PVCR_TRY(prep_weight(p.dec_w_hh, H, H3, H, w.wcat, st, H));
  PVCR_TRY(prep_weight(p.dec_w_ih, H + E, H3, H, w.wc, st));
  PVCR_TRY(prep_weight(p.dec_w_ih + H, H + E, H3, E, w.we, st));
  PVCR_TRY(prep_weight(p.out_w, H, Vc, H, g.wv, st));
  if (w.enc_a.Kp != H) {
    PVCR_TRY(fill_zero(w.enc_a.ptr, sizeof(bf16) * (size_t)BN * w.enc_a.ld, st));
    PVCR_TRY(fill_zero(g.hp.ptr, sizeof(bf16) * (size_t)R * g.hp.ld, st));
    PVCR_TRY(fill_zero(g.ctx_p.ptr, sizeof(bf16) * (size_t)R * g.ctx_p.ld, st));
  }
  // encoder and proj_key once per video
  PVCR_TRY(stage(vid, V, BN, V, w.x_a, 0, frame_scale, NO_DROPOUT, st));
  PVCR_TRY(gemm_planes(w.x_a.view(), w.wih_enc.view(), BN, H3, (int)w.x_a.ld, w.gi_enc, H3, p.enc_b_ih, 0, st));
  PVCR_TRY(gru_seq_fwd(encoder_seq(d, p, w), st));
  PVCR_TRY(gemm_planes(w.enc_a.view(), w.wk.view(), BN, H, (int)w.enc_a.ld, w.pk, H, nullptr, 0, st));
  // K hypotheses per video: rows b*K + k
  PVCR_TRY(repeat_rows(w.enc, g.encR, B, K, (long long)N * H, st));
  PVCR_TRY(repeat_rows(w.pk, g.pkR, B, K, (long long)N * H, st));
  // initial state = encoder final state (frame N-1): gather it with a 2-D copy, then replicate
  PVCR_CUDA_CHECK(cudaMemcpy2DAsync(g.hB, sizeof(float) * H, w.enc + (long long)(N - 1) * H, sizeof(float) * (size_t)N * H,
                                    sizeof(float) * H, B, cudaMemcpyDeviceToDevice, st));
  PVCR_TRY(repeat_rows(g.hB, g.hA, B, K, H, st));
  PVCR_TRY(fill_i64(g.words, sos_id, R, st));
  PVCR_TRY(fill_zero(g.scoreA, sizeof(float) * R, st));
  PVCR_TRY(fill_zero(g.histA, sizeof(long long) * (size_t)R * L, st));
  float *h_cur = g.hA, *h_new = g.hB, *sc_cur = g.scoreA, *sc_new = g.scoreB;
  long long *hist_cur = g.histA, *hist_new = g.histB;
  for (int i = 0; i < L; ++i) {
    PVCR_TRY(stage(h_cur, H, R, H, g.hp, 0, nullptr, NO_DROPOUT, st));
    PVCR_TRY(gemm_planes(g.hp.view(), w.wcat.view(), R, H4, (int)w.wcat.ld, g.g1, H4, nullptr, 0, st));
    AttnFwdArgs at{};
    at.B = R; at.N = N; at.H = H;
    at.q = g.g1; at.q_ld = H4; at.pk = g.pkR; at.enc = g.encR; at.v = p.att_v;
    at.alpha = g.alpha; at.ctx = g.ctx; at.ctx_ld = H;
    at.ctx_planes = g.ctx_p.ptr; at.ctx_planes_ld = g.ctx_p.ld; at.Hp = g.ctx_p.Kp; at.nsplit = d.nsplit;
    PVCR_TRY(attn_fwd(at, st));
    PVCR_TRY(gemm_planes(g.ctx_p.view(), w.wc.view(), R, H3, (int)g.ctx_p.ld, g.g2, H3, nullptr, 0, st));
    PVCR_TRY(gather_split(p.emb, E, g.words, R, g.emb_step.ptr, g.emb_step.ld, g.emb_step.Kp, d.nsplit, NO_DROPOUT, st));
    PVCR_TRY(gemm_planes(g.emb_step.view(), w.we.view(), R, H3, (int)g.emb_step.ld, g.g2, H3, p.dec_b_ih, 1, st));
    GruFwdArgs gf{};
    gf.B = R; gf.H = H;
    gf.gi_a = g.g2; gf.gi_a_ld = H3;
    gf.gh = g.g1 + H; gf.gh_ld = H4; gf.b_hh = p.dec_b_hh;
    gf.h_prev = h_cur; gf.h_prev_ld = H;
    gf.h_out = h_new; gf.h_out_ld = H;
    gf.h_planes = g.hp.ptr; gf.h_planes_ld = g.hp.ld; gf.Hp = g.hp.Kp; gf.nsplit = d.nsplit;
    PVCR_TRY(gru_gate_fwd(gf, st));
    PVCR_TRY(gemm_planes(g.hp.view(), g.wv.view(), R, Vc, (int)g.wv.ld, g.logits, g.ldl, p.out_b, 0, st));
    PVCR_TRY(beam_select(g.logits, g.ldl, B, Vc, K, i == 0, sc_cur, sc_new, g.parent, g.words, g.sel_scratch, st));
    PVCR_TRY(beam_reorder(h_new, h_cur, R, H, hist_cur, hist_new, L, i, g.parent, g.words, K, st));
    { float* t = sc_cur; sc_cur = sc_new; sc_new = t; }
    { long long* t = hist_cur; hist_cur = hist_new; hist_new = t; }
  }
  PVCR_CUDA_CHECK(cudaMemcpyAsync(ids, hist_cur, sizeof(long long) * (size_t)R * L, cudaMemcpyDeviceToDevice, st));
  if (scores) PVCR_CUDA_CHECK(cudaMemcpyAsync(scores, sc_cur, sizeof(float) * R, cudaMemcpyDeviceToDevice, st));
  return PVCR_OK;
}

}  // namespace pvcr
