// extern "C" surface declared in include/pvcr_b200.h.
#include "../../include/pvcr_b200.h"
#include "host.h"

namespace pvcr {
size_t linear_fwd_workspace(int M, int N, int K, int nsplit);
int linear_fwd(const float*, long long, const float*, long long, const float*, float*, long long, int, int, int, int,
               void*, size_t, cudaStream_t);
size_t linear_bwd_workspace(int M, int N, int K, int nsplit);
int linear_bwd(const float*, long long, const float*, long long, const float*, long long, float*, long long, float*,
               long long, float*, int, int, int, int, int, void*, size_t, cudaStream_t);
size_t wgrad_mn_workspace(int R, int N, int K);
int wgrad_mn(const float*, long long, const float*, long long, float*, long long, int, int, int, int, void*, size_t,
             cudaStream_t);
size_t s2vtatt_workspace(const PvcrDims& d, int need_frame_grad);
int s2vtatt_fwd(const PvcrDims&, const PvcrS2vtAttParams&, const float*, const float*, const long long*, float*, float*,
                void*, size_t, cudaStream_t);
int s2vtatt_bwd(const PvcrDims&, const PvcrS2vtAttParams&, const float*, const float*, const long long*, const float*,
                const float*, PvcrS2vtAttGrads&, float*, void*, size_t, cudaStream_t, int);
size_t s2vt_workspace(const PvcrDims& d, int need_frame_grad);
int s2vt_fwd(const PvcrDims&, const PvcrS2vtParams&, const float*, const float*, const long long*, float*, void*, size_t,
             cudaStream_t);
int s2vt_bwd(const PvcrDims&, const PvcrS2vtParams&, const float*, const float*, const long long*, const float*, float*,
             PvcrS2vtGrads&, float*, void*, size_t, cudaStream_t);
size_t s2vtatt_greedy_workspace(const PvcrDims& d);
int s2vtatt_greedy(const PvcrDims&, const PvcrS2vtAttParams&, const float*, const float*, long long, long long*, float*,
                   float*, void*, size_t, cudaStream_t);
size_t s2vt_decode_steps_workspace(const PvcrDims& d);
int s2vt_decode_steps(const PvcrDims&, const PvcrS2vtParams&, const float*, const float*, long long, const long long*,
                      const int*, float, long long*, long long*, float*, void*, size_t, cudaStream_t);
size_t generator_workspace(const PvcrDims& d);
int generator_fwd(const PvcrDims&, const PvcrGenParams&, const float*, const float*, float, int, float*, float*, float*,
                  void*, size_t, cudaStream_t);
int generator_bwd(const PvcrDims&, const PvcrGenParams&, const float*, float, const float*, const float*, const float*,
                  PvcrGenGrads&, void*, size_t, cudaStream_t);
size_t vocab_ce_workspace(int B, int L, int H, int Vc, int nsplit, float dropout_p);
int vocab_ce_prepare(const float*, int, int, int, int, int, void*, size_t, cudaStream_t);
size_t s2vtatt_beam_workspace(const PvcrDims& d, int K);
int s2vtatt_beam(const PvcrDims& d, const PvcrS2vtAttParams& p, const float* vid, const float* frame_scale, long long sos_id,
                 int K, long long* ids, float* scores, void* ws, size_t ws_bytes, cudaStream_t st);
int vocab_ce_fwd(const float*, const float*, const float*, const long long*, const long long*, int, int, int, int, int,
                 float, unsigned long long, float*, long long*, float*, float*, float*, long long, void*, size_t,
                 cudaStream_t);
int s2vtatt_decode_fwd(const PvcrDims&, const PvcrS2vtAttParams&, const float*, const float*, const long long*, float*,
                       float*, void*, size_t, cudaStream_t);
int s2vtatt_decode_bwd(const PvcrDims&, const PvcrS2vtAttParams&, const long long*, const float*, const float*,
                       PvcrS2vtAttGrads&, float*, float*, void*, size_t, cudaStream_t);
int s2vt_decode_fwd(const PvcrDims&, const PvcrS2vtParams&, const float*, const float*, const long long*, float*, void*,
                    size_t, cudaStream_t);
int s2vt_decode_bwd(const PvcrDims&, const PvcrS2vtParams&, const long long*, const float*, float*, PvcrS2vtGrads&, float*,
                    float*, void*, size_t, cudaStream_t);
int s2vtatt_greedy_impl(const PvcrDims&, const PvcrS2vtAttParams&, const float*, const float*, const float*, const float*,
                        long long, long long*, float*, float*, void*, size_t, cudaStream_t, int);
int s2vt_decode_steps_impl(const PvcrDims&, const PvcrS2vtParams&, const float*, const float*, const float*, const float*,
                           long long, const long long*, const int*, float, long long*, long long*, float*, void*, size_t,
                           cudaStream_t);
size_t gru_step_workspace(int B, int V, int H, int nsplit);
int gru_step_fwd(const float*, const float*, const float*, const float*, const float*, const float*, int, int, int, int,
                 float*, float*, void*, size_t, cudaStream_t);
int gru_step_bwd(const float*, const float*, const float*, const float*, const float*, const float*, int, int, int, int,
                 float*, float*, float*, float*, float*, float*, int, void*, size_t, cudaStream_t);
int out_dropout_apply(const float*, float*, long long, float, unsigned long long, cudaStream_t);
int philox_minmax(unsigned long long seed, unsigned long long idx0, unsigned long long n, float* minmax, cudaStream_t st);
int vocab_ce_bwd(const float*, const float*, const float*, const long long*, const long long*, int, int, int, int, int, float,
                 unsigned long long, const float*, float*, float*, float*, float*, long long*, void*, size_t,
                 cudaStream_t);
}  // namespace pvcr

using namespace pvcr;

extern "C" {

const char* pvcr_last_error(void) { return last_error(); }
int pvcr_version(void) { return 100; }

size_t pvcr_linear_fwd_workspace(int M, int N, int K, int nsplit) { return linear_fwd_workspace(M, N, K, nsplit); }
int pvcr_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y,
                    int64_t ldy, int M, int N, int K, int nsplit, void* workspace, size_t workspace_bytes,
                    void* stream) {
  return linear_fwd(x, ldx, w, ldw, bias, y, ldy, M, N, K, nsplit, workspace, workspace_bytes, (cudaStream_t)stream);
}
size_t pvcr_linear_bwd_workspace(int M, int N, int K, int nsplit) { return linear_bwd_workspace(M, N, K, nsplit); }
int pvcr_linear_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* w, int64_t ldw,
                    float* dx, int64_t lddx, float* dw, int64_t lddw, float* db, int M, int N, int K, int nsplit,
                    int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  return linear_bwd(dy, lddy, x, ldx, w, ldw, dx, lddx, dw, lddw, db, M, N, K, nsplit, accumulate, workspace,
                    workspace_bytes, (cudaStream_t)stream);
}

size_t pvcr_wgrad_mn_workspace(int R, int N, int K) { return wgrad_mn_workspace(R, N, K); }
int pvcr_wgrad_mn(const float* dy, int64_t lddy, const float* x, int64_t ldx, float* dw, int64_t lddw, int R, int N, int K,
                  int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  return wgrad_mn(dy, lddy, x, ldx, dw, lddw, R, N, K, accumulate, workspace, workspace_bytes, (cudaStream_t)stream);
}
size_t pvcr_s2vtatt_workspace(const PvcrDims* d, int need_frame_grad) { return s2vtatt_workspace(*d, need_frame_grad); }
int pvcr_s2vtatt_fwd(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* vid_feats, const float* frame_scale,
                     const int64_t* s_in, float* hs, float* alphas, void* workspace, size_t workspace_bytes,
                     void* stream) {
  return s2vtatt_fwd(*d, *p, vid_feats, frame_scale, (const long long*)s_in, hs, alphas, workspace, workspace_bytes,
                     (cudaStream_t)stream);
}
int pvcr_s2vtatt_bwd(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* vid_feats, const float* frame_scale,
                     const int64_t* s_in, const float* hs, const float* d_hs, PvcrS2vtAttGrads* g,
                     float* d_frame_scale, void* workspace, size_t workspace_bytes, void* stream) {
  return s2vtatt_bwd(*d, *p, vid_feats, frame_scale, (const long long*)s_in, d_hs, hs, *g, d_frame_scale, workspace,
                     workspace_bytes, (cudaStream_t)stream, 0);
}
int pvcr_s2vtatt_bwd_part(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* vid_feats, const float* frame_scale,
                          const int64_t* s_in, const float* hs, const float* d_hs, PvcrS2vtAttGrads* g,
                          float* d_frame_scale, void* workspace, size_t workspace_bytes, void* stream, int part) {
  if (part < 0 || part > 2) { set_last_error("pvcr_s2vtatt_bwd_part: part=%d not in 0..2", part); return PVCR_ERR_ARG; }
  return s2vtatt_bwd(*d, *p, vid_feats, frame_scale, (const long long*)s_in, d_hs, hs, *g, d_frame_scale, workspace,
                     workspace_bytes, (cudaStream_t)stream, part);
}
size_t pvcr_s2vt_workspace(const PvcrDims* d, int need_frame_grad) { return s2vt_workspace(*d, need_frame_grad); }
int pvcr_s2vt_fwd(const PvcrDims* d, const PvcrS2vtParams* p, const float* vid_feats, const float* frame_scale,
                  const int64_t* s_in, float* hs, void* workspace, size_t workspace_bytes, void* stream) {
  return s2vt_fwd(*d, *p, vid_feats, frame_scale, (const long long*)s_in, hs, workspace, workspace_bytes,
                  (cudaStream_t)stream);
}
int pvcr_s2vt_bwd(const PvcrDims* d, const PvcrS2vtParams* p, const float* vid_feats, const float* frame_scale,
                  const int64_t* s_in, float* hs, const float* d_hs, PvcrS2vtGrads* g, float* d_frame_scale,
                  void* workspace, size_t workspace_bytes, void* stream) {
  return s2vt_bwd(*d, *p, vid_feats, frame_scale, (const long long*)s_in, d_hs, hs, *g, d_frame_scale, workspace,
                  workspace_bytes, (cudaStream_t)stream);
}
size_t pvcr_s2vtatt_greedy_workspace(const PvcrDims* d) { return s2vtatt_greedy_workspace(*d); }
int pvcr_s2vtatt_greedy(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* vid_feats, const float* frame_scale,
                        int64_t sos_id, int64_t* ids, float* logits, float* alphas, void* workspace,
                        size_t workspace_bytes, void* stream) {
  return s2vtatt_greedy(*d, *p, vid_feats, frame_scale, sos_id, (long long*)ids, logits, alphas, workspace,
                        workspace_bytes, (cudaStream_t)stream);
}
size_t pvcr_s2vt_decode_steps_workspace(const PvcrDims* d) { return s2vt_decode_steps_workspace(*d); }
int pvcr_s2vt_decode_steps(const PvcrDims* d, const PvcrS2vtParams* p, const float* vid_feats, const float* frame_scale,
                           int64_t sos_id, const int64_t* teacher_words, const int32_t* teacher_mask,
                           float out_dropout_p, int64_t* ids, int64_t* fed, float* logits, void* workspace,
                           size_t workspace_bytes, void* stream) {
  return s2vt_decode_steps(*d, *p, vid_feats, frame_scale, sos_id, (const long long*)teacher_words, teacher_mask,
                           out_dropout_p, (long long*)ids, (long long*)fed, logits, workspace, workspace_bytes,
                           (cudaStream_t)stream);
}
size_t pvcr_generator_workspace(const PvcrDims* d) { return generator_workspace(*d); }
int pvcr_generator_fwd(const PvcrDims* d, const PvcrGenParams* p, const float* vid_feats, const float* noise, float tau,
                       int hard, float* probs, float* p1, float* pen, void* workspace, size_t workspace_bytes,
                       void* stream) {
  return generator_fwd(*d, *p, vid_feats, noise, tau, hard, probs, p1, pen, workspace, workspace_bytes,
                       (cudaStream_t)stream);
}
int pvcr_generator_bwd(const PvcrDims* d, const PvcrGenParams* p, const float* vid_feats, float tau, const float* d_p1,
                       const float* d_probs, const float* g_pen, PvcrGenGrads* g, void* workspace,
                       size_t workspace_bytes, void* stream) {
  return generator_bwd(*d, *p, vid_feats, tau, d_p1, d_probs, g_pen, *g, workspace, workspace_bytes,
                       (cudaStream_t)stream);
}
int pvcr_masked_ce(const float* logits, int64_t ld, int B, int L, int Vc, const int64_t* target, const int64_t* s_len,
                   const float* gscale, float* loss3, int64_t* pred, float* lse, float* nll, float* dlogits,
                   int64_t ld_d, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  PVCR_TRY(ce_rows(logits, ld, B, L, Vc, (const long long*)target, (const long long*)s_len, lse, nll, (long long*)pred,
                   dlogits, ld_d, gscale, st));
  if (loss3) PVCR_TRY(loss_finalize(nll, (const long long*)pred, (const long long*)target, (const long long*)s_len, B, L, loss3, st));
  return PVCR_OK;
}
int pvcr_rationale_penalties(const float* probs, int B, int N, float* pen, void* stream) {
  return penalties_fwd(probs, B, N, pen, (cudaStream_t)stream);
}
int pvcr_rationale_penalties_bwd(const float* probs, int B, int N, const float* g_pen, float* dprobs, void* stream) {
  return penalties_bwd(probs, B, N, g_pen, dprobs, (cudaStream_t)stream);
}
size_t pvcr_vocab_ce_workspace(int B, int L, int H, int Vc, int nsplit, float dropout_p) {
  return vocab_ce_workspace(B, L, H, Vc, nsplit, dropout_p);
}
size_t pvcr_s2vtatt_beam_workspace(const PvcrDims* d, int beam) { return d ? s2vtatt_beam_workspace(*d, beam) : 0; }
int pvcr_s2vtatt_beam(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* vid_feats, const float* frame_scale,
                      int64_t sos_id, int beam, int64_t* ids, float* scores, void* workspace, size_t workspace_bytes,
                      void* stream) {
  if (!d || !p) { set_last_error("pvcr_s2vtatt_beam: null dims / params"); return PVCR_ERR_ARG; }
  return s2vtatt_beam(*d, *p, vid_feats, frame_scale, (long long)sos_id, beam, (long long*)ids, scores, workspace,
                      workspace_bytes, (cudaStream_t)stream);
}
int pvcr_vocab_ce_prepare(const float* out_w, int B, int L, int H, int Vc, int nsplit, void* workspace,
                          size_t workspace_bytes, void* stream) {
  return vocab_ce_prepare(out_w, B, L, H, Vc, nsplit, workspace, workspace_bytes, (cudaStream_t)stream);
}
int pvcr_vocab_ce_fwd(const float* hs, const float* out_w, const float* out_b, const int64_t* target,
                      const int64_t* s_len, int B, int L, int H, int Vc, int nsplit, float dropout_p, uint64_t seed,
                      float* loss3, int64_t* pred, float* lse, float* token_nll, float* logits_out,
                      int64_t ld_logits_out, void* workspace, size_t workspace_bytes, void* stream) {
  return vocab_ce_fwd(hs, out_w, out_b, (const long long*)target, (const long long*)s_len, B, L, H, Vc, nsplit,
                      dropout_p, seed, loss3, (long long*)pred, lse, token_nll, logits_out, ld_logits_out, workspace,
                      workspace_bytes, (cudaStream_t)stream);
}
int pvcr_s2vtatt_decode_fwd(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* enc_outs, const float* enc_final,
                            const int64_t* s_in, float* hs, float* alphas, void* workspace, size_t workspace_bytes,
                            void* stream) {
  if (!d || !p) { set_last_error("pvcr_s2vtatt_decode_fwd: null dims / params"); return PVCR_ERR_ARG; }
  return s2vtatt_decode_fwd(*d, *p, enc_outs, enc_final, (const long long*)s_in, hs, alphas, workspace, workspace_bytes,
                            (cudaStream_t)stream);
}
int pvcr_s2vtatt_decode_bwd(const PvcrDims* d, const PvcrS2vtAttParams* p, const int64_t* s_in, const float* hs,
                            const float* d_hs, PvcrS2vtAttGrads* grads, float* d_enc_outs, float* d_enc_final,
                            void* workspace, size_t workspace_bytes, void* stream) {
  if (!d || !p || !grads) { set_last_error("pvcr_s2vtatt_decode_bwd: null dims / params / grads"); return PVCR_ERR_ARG; }
  return s2vtatt_decode_bwd(*d, *p, (const long long*)s_in, d_hs, hs, *grads, d_enc_outs, d_enc_final, workspace,
                            workspace_bytes, (cudaStream_t)stream);
}
int pvcr_s2vt_decode_fwd(const PvcrDims* d, const PvcrS2vtParams* p, const float* out1, const float* state1,
                         const int64_t* s_in, float* hs, void* workspace, size_t workspace_bytes, void* stream) {
  if (!d || !p) { set_last_error("pvcr_s2vt_decode_fwd: null dims / params"); return PVCR_ERR_ARG; }
  return s2vt_decode_fwd(*d, *p, out1, state1, (const long long*)s_in, hs, workspace, workspace_bytes, (cudaStream_t)stream);
}
int pvcr_s2vt_decode_bwd(const PvcrDims* d, const PvcrS2vtParams* p, const int64_t* s_in, float* hs, const float* d_hs,
                         PvcrS2vtGrads* grads, float* d_out1, float* d_state1, void* workspace, size_t workspace_bytes,
                         void* stream) {
  if (!d || !p || !grads) { set_last_error("pvcr_s2vt_decode_bwd: null dims / params / grads"); return PVCR_ERR_ARG; }
  return s2vt_decode_bwd(*d, *p, (const long long*)s_in, d_hs, hs, *grads, d_out1, d_state1, workspace, workspace_bytes,
                         (cudaStream_t)stream);
}
int pvcr_s2vtatt_decode_greedy(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* enc_outs, const float* enc_final,
                               int64_t sos_id, int64_t* ids, float* logits, float* alphas, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (!d || !p || !enc_outs || !enc_final) { set_last_error("pvcr_s2vtatt_decode_greedy: null argument"); return PVCR_ERR_ARG; }
  return s2vtatt_greedy_impl(*d, *p, nullptr, nullptr, enc_outs, enc_final, (long long)sos_id, (long long*)ids, logits,
                             alphas, workspace, workspace_bytes, (cudaStream_t)stream, 0);
}
int pvcr_s2vtatt_greedy_ex(const PvcrDims* d, const PvcrS2vtAttParams* p, const float* vid_feats, const float* frame_scale,
                           const float* enc_outs, const float* enc_final, int64_t sos_id, int64_t* ids, float* logits,
                           float* alphas, void* workspace, size_t workspace_bytes, int flags, void* stream) {
  if (!d || !p || !ids || (!vid_feats && !enc_outs)) { set_last_error("pvcr_s2vtatt_greedy_ex: null argument"); return PVCR_ERR_ARG; }
  if ((enc_outs != nullptr) != (enc_final != nullptr)) {
    set_last_error("pvcr_s2vtatt_greedy_ex: enc_outs and enc_final come together"); return PVCR_ERR_ARG;
  }
  return s2vtatt_greedy_impl(*d, *p, enc_outs ? nullptr : vid_feats, enc_outs ? nullptr : frame_scale, enc_outs, enc_final,
                             (long long)sos_id, (long long*)ids, logits, alphas, workspace, workspace_bytes,
                             (cudaStream_t)stream, flags);
}
int pvcr_s2vt_decode_greedy(const PvcrDims* d, const PvcrS2vtParams* p, const float* out1, const float* state1,
                            int64_t sos_id, int64_t* ids, float* logits, void* workspace, size_t workspace_bytes,
                            void* stream) {
  if (!d || !p || !out1 || !state1) { set_last_error("pvcr_s2vt_decode_greedy: null argument"); return PVCR_ERR_ARG; }
  return s2vt_decode_steps_impl(*d, *p, nullptr, nullptr, out1, state1, (long long)sos_id, nullptr, nullptr, 0.f,
                                (long long*)ids, nullptr, logits, workspace, workspace_bytes, (cudaStream_t)stream);
}
size_t pvcr_gru_step_workspace(int B, int V, int H, int nsplit) { return gru_step_workspace(B, V, H, nsplit); }
int pvcr_gru_step_fwd(const float* x, const float* h_prev, const float* w_ih, const float* w_hh, const float* b_ih,
                      const float* b_hh, int B, int V, int H, int nsplit, float* h_out, float* saved, void* workspace,
                      size_t workspace_bytes, void* stream) {
  return gru_step_fwd(x, h_prev, w_ih, w_hh, b_ih, b_hh, B, V, H, nsplit, h_out, saved, workspace, workspace_bytes,
                      (cudaStream_t)stream);
}
int pvcr_gru_step_bwd(const float* d_h, const float* x, const float* h_prev, const float* w_ih, const float* w_hh,
                      const float* saved, int B, int V, int H, int nsplit, float* d_x, float* d_h_prev, float* d_w_ih,
                      float* d_w_hh, float* d_b_ih, float* d_b_hh, int accumulate, void* workspace,
                      size_t workspace_bytes, void* stream) {
  return gru_step_bwd(d_h, x, h_prev, w_ih, w_hh, saved, B, V, H, nsplit, d_x, d_h_prev, d_w_ih, d_w_hh, d_b_ih, d_b_hh,
                      accumulate, workspace, workspace_bytes, (cudaStream_t)stream);
}
int pvcr_out_dropout_apply(const float* x, float* y, int64_t n, float dropout_p, uint64_t seed, void* stream) {
  return out_dropout_apply(x, y, (long long)n, dropout_p, seed, (cudaStream_t)stream);
}
int pvcr_debug_philox_minmax(uint64_t seed, uint64_t idx0, uint64_t n, float* minmax, void* stream) {
  return philox_minmax(seed, idx0, n, minmax, (cudaStream_t)stream);
}
int pvcr_vocab_ce_bwd(const float* hs, const float* out_w, const float* out_b, const int64_t* target,
                      const int64_t* s_len, int B, int L,
                      int H, int Vc, int nsplit, float dropout_p, uint64_t seed, const float* gscale, float* d_hs,
                      float* d_out_w, float* d_out_b, float* lse, int64_t* pred, void* workspace,
                      size_t workspace_bytes, void* stream) {
  return vocab_ce_bwd(hs, out_w, out_b, (const long long*)target, (const long long*)s_len, B, L, H, Vc, nsplit, dropout_p,
                      seed, gscale, d_hs, d_out_w, d_out_b, lse, (long long*)pred, workspace, workspace_bytes,
                      (cudaStream_t)stream);
}

}  // extern "C"
