// Launchers of the non-GEMM kernels (operand staging, recurrent gate math, attention, loss).
// All pointers are device pointers; all launches are stream-ordered; no allocation inside.
#pragma once
#include "common.cuh"

namespace pvcr {

// Counter-based dropout: keep iff philox(seed, element index) >= p; kept values are scaled by 1/(1-p).
struct Dropout {
  float p;                  // 0 = disabled
  unsigned long long seed;
  unsigned long long offset;   // added to the element index (distinct per tensor / per step)
  // element (r, c) of a [R, C] call maps to index ((r * row_mul + row_add) * C + c): lets a step-wise call on
  // B rows reproduce the mask of row b*L + i of the all-steps call (row_mul = L, row_add = i)
  int row_mul = 1, row_add = 0;
  // optional device counter mixed into the seed (see pvcr_set_seed_step): lets a captured CUDA graph draw a fresh
  // mask on every replay
  const unsigned long long* step = nullptr;
};
// device counter registered with pvcr_set_seed_step (null when unset); host side of the Dropout.step field
const unsigned long long* seed_step_ptr();
inline Dropout make_dropout(float p, unsigned long long seed, unsigned long long offset) {
  Dropout d{p, seed, offset};
  d.step = p > 0.f ? seed_step_ptr() : nullptr;
  return d;
}

// ---- operand staging: fp32 -> bf16 split planes -------------------------------------------------
// out[r][p*Cp + c] = term_{role,p}( in[r*ld_in + c] * (row_scale ? row_scale[r] : 1) * dropout ), zero for c >= C.
int cast_split(const float* in, long long ld_in, int R, int C, bf16* out, long long ld_out, int Cp, int nsplit,
               int role_b, const float* row_scale, Dropout drop, cudaStream_t st);
// out[c][p*Rp + r_off + r] = term_{role,p}( src(r)[c] * (row_scale ? row_scale[r] : 1) ) for r < R, where
// src(r) = in + (row_ids ? row_ids[r] : r) * ld_in.  With zero_pad, columns r_off+R .. Rp-1 are zero-filled.
int block_planes(const bf16* in, long long ld, int R, bf16* out, cudaStream_t st);      // [R, ld] -> [ld / 64][R][64]
int transpose_split(const float* in, long long ld_in, int R, int C, bf16* out, long long ld_out, int Rp, int r_off,
                    int zero_pad, int nsplit, int role_b, const long long* row_ids, const float* row_scale,
                    cudaStream_t st, Dropout drop);     // drop: element index r*C + c (as gather_split)
// out[i][p*Ep + e] = term_{A,p}( table[ids[i]][e] * dropout )     (embedding rows as an A operand)
int gather_split(const float* table, int E, const long long* ids, int n_ids, bf16* out, long long ld_out, int Ep,
                 int nsplit, Dropout drop, cudaStream_t st);
// table_grad[ids[i]][e] += rows[i*ld + e] * dropout-mask      (dense embedding gradient)
int scatter_add_rows(const float* rows, long long ld, const long long* ids, int n_ids, int E, float* table_grad,
                     Dropout drop, cudaStream_t st);
// out[c] (+)= sum_r in[r*ld + c]
int colsum(const float* in, long long ld, int R, int C, float* out, int accumulate, cudaStream_t st);
// y[i] = a[i] * mask_i/(1-p)  (dropout backward / forward on fp32 data), in place allowed
int dropout_apply(const float* in, float* out, long long n, Dropout drop, cudaStream_t st);
int fill_zero(void* p, size_t bytes, cudaStream_t st);
int add_inplace(float* y, const float* x, long long n, cudaStream_t st);      // y += x
// out[r * out_stride] = argmax_c in[r*ld + c] (first maximal index), and optionally
// next[r] = use_teacher ? teacher[r * teacher_stride] : argmax   (the word fed to the next decoding step)
int argmax_rows(const float* in, long long ld, int R, int C, long long* out, long long out_stride, long long* next,
                const long long* teacher, long long teacher_stride, int use_teacher, cudaStream_t st);
int fill_i64(long long* p, long long v, int n, cudaStream_t st);
// beam-search bookkeeping (beam.cu)
int repeat_rows(const float* in, float* out, int rows_in, int K, long long len, cudaStream_t st);
size_t beam_select_scratch(int B, int K);
int beam_select(const float* logits, long long ld, int B, int Vc, int K, int first, const float* score_in, float* score_out,
                int* parent, long long* word, void* scratch, cudaStream_t st);
int beam_reorder(const float* h_in, float* h_out, int R, int H, const long long* hist_in, long long* hist_out, int L, int step,
                 const int* parent, const long long* word, int K, cudaStream_t st);

// ---- GRU gate math ---------------------------------------------------------------------------------
struct GruFwdArgs {
  int B, H;
  const float* gi_a; long long gi_a_ld;     // [B,3H] (nullable: the step has no input, e.g. S2VT rnn1 while decoding)
  const float* gi_b; long long gi_b_ld;     // optional second addend
  const long long* gi_b_rows = nullptr;     // optional: video b reads row gi_b_rows[b] of gi_b (a table indexed by word id)
  // optional (greedy decoding): the row index of video b is the arg-max over am_nparts (max, index) partials that the
  // vocabulary GEMM's epilogue left for row b (gemm_argmax with out == nullptr) -- the combine pass of the previous step
  // folded into this kernel: one launch less on the fed-back-word path.  The word goes to am_out[b * am_out_stride].
  const float* am_pmax = nullptr; const int* am_pidx = nullptr; int am_nparts = 0;
  long long* am_out = nullptr; long long am_out_stride = 0;
  const float* gi_bias;                     // optional [3H] bias addend
  const float* gh; long long gh_ld;         // [B,3H] W_hh h (no bias) or null (h_prev == 0)
  const float* b_hh;                        // [3H]
  const float* h_prev; long long h_prev_ld; // null => zeros
  float* h_out; long long h_out_ld;
  bf16* h_planes; long long h_planes_ld; int Hp, nsplit;   // optional A-role planes of h_out
  float *r, *z, *n, *ghn;                   // saved [B,H] each (nullable as a group)
};
int gru_gate_fwd(const GruFwdArgs& a, cudaStream_t st);

struct GruBwdArgs {
  int B, H;
  const float* dh_a; long long dh_a_ld;     // nullable
  const float* dh_b; long long dh_b_ld;     // nullable
  const float *r, *z, *n, *ghn;             // saved
  const float* h_prev; long long h_prev_ld; // null => zeros
  float* dgi; long long dgi_ld;             // [B,3H] fp32
  float* dgh; long long dgh_ld;             // [B,3H] fp32
  bf16* dgi_planes; long long dgi_planes_ld; int dgi_Kp, dgi_col0;   // optional A-role planes (k = dgi_col0 + c)
  bf16* dgh_planes; long long dgh_planes_ld; int dgh_Kp, dgh_col0;
  int nsplit;
  float* dh_direct; long long dh_direct_ld; // dh * z
};
int gru_gate_bwd(const GruBwdArgs& a, cudaStream_t st);

// ---- LSTM gate math (gate order i,f,g,o) -------------------------------------------------------------
struct LstmFwdArgs {
  int B, H;
  const float* gi; long long gi_ld;         // [B,4H] (includes b_ih)
  const float* gh; long long gh_ld;         // [B,4H] or null (h_prev == 0)
  const float* b_hh;
  const float* c_prev;                      // [B,H] contiguous or null
  float* h_out; long long h_out_ld;
  bf16* h_planes; long long h_planes_ld; int Hp, nsplit;
  float *i, *f, *g, *o, *c;                 // saved [B,H] each
};
int lstm_gate_fwd(const LstmFwdArgs& a, cudaStream_t st);
struct LstmBwdArgs {
  int B, H;
  const float* dh_a; long long dh_a_ld;     // carry (nullable)
  const float* dh_b; long long dh_b_ld;     // external (nullable)
  float* dc;                                // [B,H] carry, in/out (zero-initialised by the caller)
  const float *i, *f, *g, *o, *c, *c_prev;  // saved; c_prev null => zeros
  float* da; long long da_ld;               // [B,4H]
  bf16* da_planes; long long da_planes_ld; int Kp, nsplit;
};
int lstm_gate_bwd(const LstmBwdArgs& a, cudaStream_t st);

// ---- additive attention step ---------------------------------------------------------------------------
struct AttnFwdArgs {
  int B, N, H;
  const float* q; long long q_ld;           // [B,H] = W_q h
  const float* pk;                          // [B,N,H]
  const float* enc;                         // [B,N,H]
  const float* v;                           // [H]
  float* alpha;                             // [B,N]
  float* ctx; long long ctx_ld;             // [B,H]
  bf16* ctx_planes; long long ctx_planes_ld; int Hp, nsplit;
  // projected-value mode (decoding): val [B,N,W] = enc W_c^T hoisted out of the step loop; the kernel then writes
  // out [B,W] = sum_n alpha_n val[b,n,:] (= W_c ctx) instead of ctx -- no context vector, no GEMM behind the attention
  const float* val = nullptr; int W = 0;
  float* out = nullptr; long long out_ld = 0;
};
int attn_fwd(const AttnFwdArgs& a, cudaStream_t st);
bool attn_fwd_projected_ok(int N, int H, int W);      // shapes the projected-value mode serves
struct AttnBwdArgs {
  int B, N, H;
  const float* dctx; long long dctx_ld;
  const float* q; long long q_ld;
  const float *pk, *enc, *v, *alpha;
  float* dq; long long dq_ld;               // [B,H] fp32
  bf16* dq_planes; long long dq_planes_ld; int dq_Kp, nsplit;   // k = c
  float* dpk;                               // [B,N,H] +=
  float* denc;                              // [B,N,H] +=
  float* dv_part;                           // [B,H] +=
};
int attn_bwd(const AttnBwdArgs& a, cudaStream_t st);

// ---- loss --------------------------------------------------------------------------------------------
// Row-wise cross entropy on materialised fp32 logits [R = B*L, Vc]; token (b,l) = row b*L + l.
// Writes lse/nll/pred per row; if dlogits != null also dlogits = (softmax - onehot) * w[row] * gscale
// where w = (l < s_len[b]) / (s_len[b] * B).
int ce_rows(const float* logits, long long ld, int B, int L, int Vc, const long long* target, const long long* s_len,
            float* lse, float* nll, long long* pred, float* dlogits, long long ld_d, const float* gscale,
            cudaStream_t st);
// out[0] = mean_b( sum_l nll*mask / s_len ), out[1] = #correct (masked), out[2] = #mask
int loss_finalize(const float* nll, const long long* pred, const long long* target, const long long* s_len, int B,
                  int L, float* out3, cudaStream_t st);

// ---- RationaleNet generator head ------------------------------------------------------------------------
// logits = [hf;hb] W^T + b (2 classes) ; y = softmax((logits - log(noise)) / tau) ; probs = hard ? onehot-st : y
struct GumbelArgs {
  int B, N, H;                               // hf/hb: forward / reverse direction LSTM outputs, element (b, n) at
  const float *hf, *hb;                      //        n*h_ts + b*h_bs
  long long h_ts, h_bs;
  const float* w;                            // [2, 2H]
  const float* bias;                         // [2]
  const float* noise;                        // [B*N, 2] Exp(1) draws (row = b*N + n), or null => philox
  unsigned long long seed;
  const unsigned long long* seed_step;       // as Dropout.step
  float tau; int hard;
  Dropout drop;                              // dropout on the LSTM outputs
  float* probs;                              // [B,N,2]
  float* p1;                                 // [B,N] = probs[:,:,1] (the frame scale; nullable)
  float* y;                                  // [B,N,2] soft sample (saved for backward)
  float* pen;                                // [2]: brevity, continuity losses (unscaled)
};
int gumbel_select_fwd(const GumbelArgs& a, cudaStream_t st);
struct GumbelBwdArgs {
  int B, N, H;
  const float *hf, *hb, *w, *y;
  long long h_ts, h_bs;
  float tau;
  Dropout drop;
  const float* dp1_sel;                      // [B,N] d(loss)/d(p1) through the feature scaling (nullable)
  const float* dprobs;                       // [B,N,2] external gradient on probs (nullable)
  const float* g_pen;                        // device [2]: d(loss)/d(brevity), d(loss)/d(continuity) (nullable = 0)
  float* dhf; float* dhb;                    // [N,B,H] each
  float* dw;                                 // [2,2H]  (overwritten)
  float* dbias;                              // [2]
  float* scratch;                            // [B*N*2] dlogits
};
int gumbel_select_bwd(const GumbelBwdArgs& a, cudaStream_t st);
// pen[2] = { mean_b sum_n p1, mean_{b, n>=1} |p1[b,n] - p1[b,n-1]| } with p1 = probs[:,:,1]; and its gradient
int penalties_fwd(const float* probs, int B, int N, float* pen, cudaStream_t st);
int penalties_bwd(const float* probs, int B, int N, const float* g_pen, float* dprobs, cudaStream_t st);
// dp1[b,n] = sum_v x[b,n,v] * dsel[b,n,v]
int rowdot(const float* x, const float* dsel, int R, int C, float* out, cudaStream_t st);

}  // namespace pvcr
