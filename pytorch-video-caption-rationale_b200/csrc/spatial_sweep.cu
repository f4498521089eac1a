// SpatialNet's frame loop as ONE library call per direction (reference model/SpatialNet.py:114-138):
//     for every frame t:  q = query_layer(h_{t-1});  alpha_t, ctx_t = attention over the K*K cells (keys = key_layer(conv_feats),
//                         values = the frame's input features);  h_t = GRU(ctx_t, h_{t-1})          (encode_step, :127)
// The step-wise drop-in (functional.Linear + SpatialAttnStep + GruStep per frame) staged every weight once per FRAME
// (W_ih [3H, F] alone is 12.6 MB of fp32 at cfg4, 40 x forward and 40 x transposed backward), produced forty [3H, F] weight
// gradients that autograd then summed, and issued ~1 500 launches per fwd+bwd.  Here the weights are staged once per call, [W_q; W_hh]
// is one stacked B operand (q and W_hh h from one product, as in the decoder sweeps), the gate kernels write the bf16 operand planes
// of the next product directly, and the weight gradients are three products over all frames after the loop
// (d W_ih = d gi_all^T ctx_all,  d [W_q; W_hh] = [dq | d gh]_all^T h_prev_all).  Per frame: 5 launches forward, 5 backward.
//
// Layouts: proj_key [B, N, Kc, H] and feats [B, N, Kc, F] as the front leaves them (frame t of video b at ((b N + t) Kc) rows);
// outs [N, B, H] (= torch.cat of the reference's per-frame outputs, :129-132), alphas [N, B, Kc].  Initial state zeros (:114).
#include "../../include/pvcr_b200.h"
#include "host.h"

namespace pvcr {

int spatial_attn_fwd_launch(int B, int Kc, int H, int Fv, const float* q, long long q_ld, const float* proj_key, long long pk_batch_stride,
                            const float* feats, long long feats_batch_stride, const float* v, float* alpha, float* ctx, cudaStream_t stream);
int spatial_attn_bwd_launch(int B, int Kc, int H, int Fv, const float* dctx, const float* q, long long q_ld, const float* proj_key,
                            long long pk_batch_stride, const float* feats, long long feats_batch_stride, const float* v,
                            const float* alpha, float* dq, long long dq_ld, float* dproj_key, long long dpk_batch_stride,
                            float* dv_part, cudaStream_t stream);

namespace {

struct SweepWs {
  // forward (kept for the backward)
  Planes wqhh;       // B role [4H, H]: rows [0, H) = W_q, [H, 4H) = W_hh
  Planes wih;        // B role [3H, F]
  Planes hp;         // A role [B, H]: h_{t-1}
  Planes cp;         // A role [B, F]: ctx_t
  float* qgh;        // [N][B, 4H]: q_t | W_hh h_{t-1}
  float* gi;         // [B, 3H]
  float* saved;      // [N][4][B, H]: r, z, n, W_hn h + b_hn
  float* ctx;        // [N][B, F]
  // backward
  Planes wihT;       // B role [F, 3H]:  d ctx = d gi W_ih
  Planes wqhhT;      // B role [H, 4H]:  d h  += [dq | d gh] [W_q; W_hh]
  Planes dgip;       // A role [B, 3H]
  Planes d1p;        // A role [B, 4H]
  float* dgi_all;    // [N][B, 3H]
  float* d1_all;     // [N][B, 4H]: dq_t | d gh_t
  float* dctx;       // [B, F]
  float* dh;         // [B, H] carry
  float* dvp;        // [N][B, H]
};

void carve(Arena& a, int B, int N, int H, int F, int ns, SweepWs& w) {
  w.wqhh = alloc_planes(a, 4 * H, H, ns);
  w.wih = alloc_planes(a, 3 * H, F, ns);
  w.hp = alloc_planes(a, B, H, ns);
  w.cp = alloc_planes(a, B, F, ns);
  w.qgh = a.alloc<float>((size_t)N * B * 4 * H);
  w.gi = a.alloc<float>((size_t)B * 3 * H);
  w.saved = a.alloc<float>((size_t)N * 4 * B * H);
  w.ctx = a.alloc<float>((size_t)N * B * F);
  w.wihT = alloc_planes(a, F, 3 * H, ns);
  w.wqhhT = alloc_planes(a, H, 4 * H, ns);
  w.dgip = alloc_planes(a, B, 3 * H, ns);
  w.d1p = alloc_planes(a, B, 4 * H, ns);
  w.dgi_all = a.alloc<float>((size_t)N * B * 3 * H);
  w.d1_all = a.alloc<float>((size_t)N * B * 4 * H);
  w.dctx = a.alloc<float>((size_t)B * F);
  w.dh = a.alloc<float>((size_t)B * H);
  w.dvp = a.alloc<float>((size_t)N * B * H);
}
size_t grad_scratch(int B, int N, int H, int F, int ns) {
  const size_t s1 = grad_w_scratch(N * B, 3 * H, F, ns), s2 = grad_w_scratch(N * B, 3 * H, H, ns);
  return (s1 > s2 ? s1 : s2) + 1024;
}

}  // namespace

size_t spatial_encode_workspace(int B, int N, int Kc, int H, int F, int ns) {
  (void)Kc;
  Arena a(nullptr, 0);
  SweepWs w;
  carve(a, B, N, H, F, ns, w);
  return a.off + grad_scratch(B, N, H, F, ns) + 4096;
}

int spatial_encode_fwd(int B, int N, int Kc, int H, int F, int ns, const float* proj_key, const float* feats, const float* w_q,
                       const float* v, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* outs,
                       float* alphas, void* ws, size_t ws_bytes, cudaStream_t st) {
  PVCR_REQUIRE(B > 0 && N > 0 && Kc > 0 && H > 0 && F > 0 && ns >= 1 && ns <= 3, "spatial_encode_fwd: B=%d N=%d Kc=%d H=%d F=%d nsplit=%d",
               B, N, Kc, H, F, ns);
  Arena a(ws, ws_bytes);
  SweepWs w;
  carve(a, B, N, H, F, ns, w);
  if (a.failed) { set_last_error("spatial_encode_fwd: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  const int H3 = 3 * H, H4 = 4 * H;
  PVCR_TRY(prep_weight(w_q, H, H, H, w.wqhh, st, 0));
  PVCR_TRY(prep_weight(w_hh, H, H3, H, w.wqhh, st, H));
  PVCR_TRY(prep_weight(w_ih, F, H3, F, w.wih, st, 0));
  PVCR_TRY(fill_zero(w.qgh, sizeof(float) * (size_t)B * H4, st));              // q_0 = W_q 0
  PVCR_TRY(fill_zero(w.hp.ptr, sizeof(bf16) * (size_t)B * w.hp.ld, st));       // (padding columns of the planes stay zero)
  PVCR_TRY(fill_zero(w.cp.ptr, sizeof(bf16) * (size_t)B * w.cp.ld, st));
  const long long pk_bs = (long long)N * Kc * H, f_bs = (long long)N * Kc * F;
  for (int t = 0; t < N; ++t) {
    float* qgh = w.qgh + (size_t)t * B * H4;
    float* ctx = w.ctx + (size_t)t * B * F;
    if (t > 0) PVCR_TRY(gemm_planes(w.hp.view(), w.wqhh.view(), B, H4, (int)w.hp.ld, qgh, H4, nullptr, 0, st));
    PVCR_TRY(spatial_attn_fwd_launch(B, Kc, H, F, qgh, H4, proj_key + (size_t)t * Kc * H, pk_bs, feats + (size_t)t * Kc * F, f_bs, v,
                                     alphas + (size_t)t * B * Kc, ctx, st));
    PVCR_TRY(stage(ctx, F, B, F, w.cp, 0, nullptr, NO_DROPOUT, st));
    PVCR_TRY(gemm_planes(w.cp.view(), w.wih.view(), B, H3, (int)w.cp.ld, w.gi, H3, b_ih, 0, st));
    GruFwdArgs g{};
    g.B = B; g.H = H;
    g.gi_a = w.gi; g.gi_a_ld = H3;
    g.gh = t > 0 ? qgh + H : nullptr; g.gh_ld = H4;
    g.b_hh = b_hh;
    g.h_prev = t > 0 ? outs + (size_t)(t - 1) * B * H : nullptr; g.h_prev_ld = H;
    g.h_out = outs + (size_t)t * B * H; g.h_out_ld = H;
    g.h_planes = w.hp.ptr; g.h_planes_ld = w.hp.ld; g.Hp = w.hp.Kp; g.nsplit = ns;
    float* sv = w.saved + (size_t)t * 4 * B * H;
    const size_t o = (size_t)B * H;
    g.r = sv; g.z = sv + o; g.n = sv + 2 * o; g.ghn = sv + 3 * o;
    PVCR_TRY(gru_gate_fwd(g, st));
  }
  return PVCR_OK;
}

// d_outs [N, B, H]: gradient on every h_t.  d_proj_key [B, N, Kc, H] (every element written), d_w_q [H, H], d_v [H], d_w_ih [3H, F],
// d_w_hh [3H, H], d_b_ih / d_b_hh [3H]: all overwritten.  `ws` is the forward call's workspace, `outs` / `alphas` its outputs.
int spatial_encode_bwd(int B, int N, int Kc, int H, int F, int ns, const float* proj_key, const float* feats, const float* w_q,
                       const float* v, const float* w_ih, const float* w_hh, const float* outs, const float* alphas,
                       const float* d_outs, float* d_proj_key, float* d_w_q, float* d_v, float* d_w_ih, float* d_w_hh, float* d_b_ih,
                       float* d_b_hh, void* ws, size_t ws_bytes, cudaStream_t st) {
  PVCR_REQUIRE(B > 0 && N > 0 && Kc > 0 && H > 0 && F > 0 && ns >= 1 && ns <= 3, "spatial_encode_bwd: B=%d N=%d Kc=%d H=%d F=%d nsplit=%d",
               B, N, Kc, H, F, ns);
  Arena a(ws, ws_bytes);
  SweepWs w;
  carve(a, B, N, H, F, ns, w);
  const size_t gs = grad_scratch(B, N, H, F, ns);
  char* scratch = a.alloc<char>(gs);
  if (a.failed) { set_last_error("spatial_encode_bwd: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  const int H3 = 3 * H, H4 = 4 * H;
  PVCR_TRY(prep_weight_T(w_ih, F, H3, F, w.wihT, 0, 1, st));
  PVCR_TRY(prep_weight_T(w_q, H, H, H, w.wqhhT, 0, 0, st));
  PVCR_TRY(prep_weight_T(w_hh, H, H3, H, w.wqhhT, H, 1, st));
  PVCR_TRY(fill_zero(w.dh, sizeof(float) * (size_t)B * H, st));
  PVCR_TRY(fill_zero(w.dgip.ptr, sizeof(bf16) * (size_t)B * w.dgip.ld, st));
  PVCR_TRY(fill_zero(w.d1p.ptr, sizeof(bf16) * (size_t)B * w.d1p.ld, st));
  const long long pk_bs = (long long)N * Kc * H, f_bs = (long long)N * Kc * F;
  for (int t = N - 1; t >= 0; --t) {
    float* dgi = w.dgi_all + (size_t)t * B * H3;
    float* d1 = w.d1_all + (size_t)t * B * H4;
    const float* sv = w.saved + (size_t)t * 4 * B * H;
    const size_t o = (size_t)B * H;
    GruBwdArgs b{};
    b.B = B; b.H = H;
    b.dh_a = w.dh; b.dh_a_ld = H;
    b.dh_b = d_outs + (size_t)t * B * H; b.dh_b_ld = H;
    b.r = sv; b.z = sv + o; b.n = sv + 2 * o; b.ghn = sv + 3 * o;
    b.h_prev = t > 0 ? outs + (size_t)(t - 1) * B * H : nullptr; b.h_prev_ld = H;
    b.dgi = dgi; b.dgi_ld = H3;
    b.dgh = d1 + H; b.dgh_ld = H4;
    b.dgi_planes = w.dgip.ptr; b.dgi_planes_ld = w.dgip.ld; b.dgi_Kp = w.dgip.Kp; b.dgi_col0 = 0;
    b.nsplit = ns;
    b.dh_direct = w.dh; b.dh_direct_ld = H;                  // dh * z: the direct path into h_{t-1}
    PVCR_TRY(gru_gate_bwd(b, st));
    PVCR_TRY(gemm_planes(w.dgip.view(), w.wihT.view(), B, F, (int)w.dgip.ld, w.dctx, F, nullptr, 0, st));
    PVCR_TRY(spatial_attn_bwd_launch(B, Kc, H, F, w.dctx, w.qgh + (size_t)t * B * H4, H4, proj_key + (size_t)t * Kc * H, pk_bs,
                                     feats + (size_t)t * Kc * F, f_bs, v, alphas + (size_t)t * B * Kc, d1, H4,
                                     d_proj_key + (size_t)t * Kc * H, pk_bs, w.dvp + (size_t)t * B * H, st));
    if (t > 0) {
      PVCR_TRY(stage(d1, H4, B, H4, w.d1p, 0, nullptr, NO_DROPOUT, st));
      PVCR_TRY(gemm_planes(w.d1p.view(), w.wqhhT.view(), B, H, (int)w.d1p.ld, w.dh, H, nullptr, 1, st));
    }
  }
  // parameter gradients: products over all frames
  Arena sa(scratch, gs);
  const int R = N * B;
  PVCR_TRY(grad_w(sa, w.dgi_all, H3, R, H3, w.ctx, F, F, nullptr, nullptr, d_w_ih, F, 0, ns, st));
  if (N > 1) {
    const int R1 = (N - 1) * B;              // frame t pairs with h_{t-1} = outs[t-1]; h_{-1} = 0 contributes nothing
    const float* d1 = w.d1_all + (size_t)B * H4;
    PVCR_TRY(grad_w(sa, d1, H4, R1, H, outs, H, H, nullptr, nullptr, d_w_q, H, 0, ns, st));
    PVCR_TRY(grad_w(sa, d1 + H, H4, R1, H3, outs, H, H, nullptr, nullptr, d_w_hh, H, 0, ns, st));
  } else {
    PVCR_TRY(fill_zero(d_w_q, sizeof(float) * (size_t)H * H, st));
    PVCR_TRY(fill_zero(d_w_hh, sizeof(float) * (size_t)H3 * H, st));
  }
  PVCR_TRY(colsum(w.dgi_all, H3, R, H3, d_b_ih, 0, st));
  PVCR_TRY(colsum(w.d1_all + H, H4, R, H3, d_b_hh, 0, st));
  PVCR_TRY(colsum(w.dvp, H, R, H, d_v, 0, st));
  return PVCR_OK;
}

}  // namespace pvcr

using namespace pvcr;

extern "C" {

size_t pvcr_spatial_encode_workspace(int B, int N, int Kc, int H, int F, int nsplit) {
  return spatial_encode_workspace(B, N, Kc, H, F, nsplit);
}
int pvcr_spatial_encode_fwd(int B, int N, int Kc, int H, int F, int nsplit, const float* proj_key, const float* feats, const float* w_q,
                            const float* v, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* outs,
                            float* alphas, void* workspace, size_t workspace_bytes, void* stream) {
  if (!proj_key || !feats || !w_q || !v || !w_ih || !w_hh || !b_ih || !b_hh || !outs || !alphas || !workspace) {
    set_last_error("pvcr_spatial_encode_fwd: null argument");
    return PVCR_ERR_ARG;
  }
  return spatial_encode_fwd(B, N, Kc, H, F, nsplit, proj_key, feats, w_q, v, w_ih, w_hh, b_ih, b_hh, outs, alphas, workspace,
                            workspace_bytes, (cudaStream_t)stream);
}
int pvcr_spatial_encode_bwd(int B, int N, int Kc, int H, int F, int nsplit, const float* proj_key, const float* feats, const float* w_q,
                            const float* v, const float* w_ih, const float* w_hh, const float* outs, const float* alphas,
                            const float* d_outs, float* d_proj_key, float* d_w_q, float* d_v, float* d_w_ih, float* d_w_hh,
                            float* d_b_ih, float* d_b_hh, void* workspace, size_t workspace_bytes, void* stream) {
  if (!proj_key || !feats || !w_q || !v || !w_ih || !w_hh || !outs || !alphas || !d_outs || !d_proj_key || !d_w_q || !d_v || !d_w_ih ||
      !d_w_hh || !d_b_ih || !d_b_hh || !workspace) {
    set_last_error("pvcr_spatial_encode_bwd: null argument");
    return PVCR_ERR_ARG;
  }
  return spatial_encode_bwd(B, N, Kc, H, F, nsplit, proj_key, feats, w_q, v, w_ih, w_hh, outs, alphas, d_outs, d_proj_key, d_w_q, d_v,
                            d_w_ih, d_w_hh, d_b_ih, d_b_hh, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
