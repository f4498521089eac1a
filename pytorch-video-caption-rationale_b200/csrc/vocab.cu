// Vocabulary projection + masked cross entropy (train_utils.py:37-71 on top of
// model/S2VTAttModel.py:145 / model/S2VTModel.py:130 `Dropout + Linear`).
//
// Generic-precision path (any nsplit): logits are materialised once in the workspace (fp32), reduced
// row-wise, and overwritten in place by d(loss)/d(logits) for the gradient GEMMs.
#include "../../include/pvcr_b200.h"
#include <cstdlib>

#include "host.h"

namespace pvcr {

size_t vocab_fused_workspace(int M, int H, int Vc);
int vocab_fused_fwd(const float*, const float*, const float*, const long long*, const long long*, int, int, int, int,
                    float, unsigned long long, float*, long long*, float*, float*, void*, size_t, cudaStream_t);
int vocab_fused_bwd(const float*, const float*, const float*, const long long*, const long long*, int, int, int, int,
                    float, unsigned long long, const float*, float*, float*, float*, const float*, void*, size_t,
                    cudaStream_t);
int vocab_fused_prepare(const float*, int, int, int, int, void*, size_t, cudaStream_t);
static bool fused_off() { static const bool off = getenv("PVCR_NO_FUSED_CE") != nullptr; return off; }

struct VocabWs {
  Planes hs_a, wv;
  float *logits, *nll, *hs_drop;
  long long ldl;
};

static void carve_vocab(Arena& a, int M, int H, int Vc, int nsplit, float dropout_p, VocabWs& w) {
  w.hs_a = alloc_planes(a, M, H, nsplit);
  w.wv = alloc_planes(a, Vc, H, nsplit);
  w.ldl = round_up(Vc, 4);
  w.logits = a.alloc<float>((size_t)M * w.ldl);
  w.nll = a.alloc<float>((size_t)M);
  w.hs_drop = dropout_p > 0.f ? a.alloc<float>((size_t)M * H) : nullptr;
}

size_t vocab_ce_workspace(int B, int L, int H, int Vc, int nsplit, float dropout_p) {
  Arena a(nullptr, 0);
  VocabWs w;
  const int M = B * L;
  carve_vocab(a, M, H, Vc, nsplit, dropout_p, w);
  // backward scratch: grad_x (dlogits planes + W^T planes) and grad_w (dlogits^T, hs^T planes)
  size_t peak = 0;
  { size_t m = a.mark(); alloc_planes(a, H, Vc, nsplit); alloc_planes(a, M, Vc, nsplit); peak = a.off; a.release(m); }
  { const size_t need = a.mark() + grad_w_scratch(M, Vc, H, nsplit); if (need > peak) peak = need; }
  const size_t fused = nsplit == 1 ? vocab_fused_workspace(M, H, Vc) : 0;
  return (peak > fused ? peak : fused) + 4096;
}

static Dropout out_dropout(float p, unsigned long long seed) { return make_dropout(p, seed, 0x5000000000ull); }

int vocab_ce_prepare(const float* wv, int B, int L, int H, int Vc, int nsplit, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (nsplit != 1 || fused_off()) return PVCR_OK;
  return vocab_fused_prepare(wv, B, L, H, Vc, ws, ws_bytes, st);
}

// loss3 = {masked loss, #correct, #mask}; pred [B*L] int64; logits stay in the workspace for vocab_ce_bwd.
int vocab_ce_fwd(const float* hs, const float* wv, const float* bv, const long long* target, const long long* s_len,
                 int B, int L, int H, int Vc, int nsplit, float dropout_p, unsigned long long seed, float* loss3,
                 long long* pred, float* lse, float* token_nll, float* logits_out, long long ld_logits_out, void* ws,
                 size_t ws_bytes, cudaStream_t st) {
  PVCR_REQUIRE(nsplit >= 1 && nsplit <= 3, "vocab_ce_fwd: nsplit=%d", nsplit);
  if (nsplit == 1 && !logits_out && target && !fused_off())
    return vocab_fused_fwd(hs, wv, bv, target, s_len, B, L, H, Vc, dropout_p, seed, loss3, pred, lse, token_nll, ws,
                           ws_bytes, st);
  if (side_note_take(ws, NOTE_VOCAB_WV, wv)) PVCR_TRY(side_join(st));     // a prepare whose fused layout is not used here
  const int M = B * L;
  Arena a(ws, ws_bytes);
  VocabWs w;
  carve_vocab(a, M, H, Vc, nsplit, dropout_p, w);
  if (a.failed) { set_last_error("vocab_ce_fwd: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  const Dropout dr = out_dropout(dropout_p, seed);
  PVCR_TRY(stage(hs, H, M, H, w.hs_a, 0, nullptr, dr, st));
  if (w.hs_drop) PVCR_TRY(dropout_apply(hs, w.hs_drop, (long long)M * H, dr, st));
  PVCR_TRY(prep_weight(wv, H, Vc, H, w.wv, st));
  float* logits = logits_out ? logits_out : w.logits;
  const long long ldl = logits_out ? ld_logits_out : w.ldl;
  PVCR_TRY(gemm_planes(w.hs_a.view(), w.wv.view(), M, Vc, (int)w.hs_a.ld, logits, ldl, bv, 0, st));
  if (target) {
    PVCR_TRY(ce_rows(logits, ldl, B, L, Vc, target, s_len, lse, w.nll, pred, nullptr, 0, nullptr, st));
    PVCR_TRY(loss_finalize(w.nll, pred, target, s_len, B, L, loss3, st));
    if (token_nll) PVCR_CUDA_CHECK(cudaMemcpyAsync(token_nll, w.nll, sizeof(float) * M, cudaMemcpyDeviceToDevice, st));
  }
  return PVCR_OK;
}

// Requires the workspace of the preceding vocab_ce_fwd (logits inside).  gscale: device scalar d(total)/d(loss) or null.
int vocab_ce_bwd(const float* hs, const float* wv, const float* bv, const long long* target, const long long* s_len, int B, int L, int H,
                 int Vc, int nsplit, float dropout_p, unsigned long long seed, const float* gscale, float* d_hs,
                 float* d_wv, float* d_bv, float* lse, long long* pred, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (nsplit == 1 && !fused_off())
    return vocab_fused_bwd(hs, wv, bv, target, s_len, B, L, H, Vc, dropout_p, seed, gscale, d_hs, d_wv, d_bv, lse, ws,
                           ws_bytes, st);
  const int M = B * L;
  Arena a(ws, ws_bytes);
  VocabWs w;
  carve_vocab(a, M, H, Vc, nsplit, dropout_p, w);
  if (a.failed) { set_last_error("vocab_ce_bwd: workspace too small"); return PVCR_ERR_WORKSPACE; }
  // dlogits in place of logits
  PVCR_TRY(ce_rows(w.logits, w.ldl, B, L, Vc, target, s_len, lse, w.nll, pred, w.logits, w.ldl, gscale, st));
  {
    const size_t m = a.mark();
    Planes wvT = alloc_planes(a, H, Vc, nsplit);
    if (a.failed) { set_last_error("vocab_ce_bwd: workspace too small"); return PVCR_ERR_WORKSPACE; }
    PVCR_TRY(prep_weight_T(wv, H, Vc, H, wvT, 0, 1, st));
    PVCR_TRY(grad_x(a, w.logits, w.ldl, M, Vc, wvT, d_hs, H, 0, st));
    if (a.failed) { set_last_error("vocab_ce_bwd: workspace too small"); return PVCR_ERR_WORKSPACE; }
    a.release(m);
  }
  if (dropout_p > 0.f) PVCR_TRY(dropout_apply(d_hs, d_hs, (long long)M * H, out_dropout(dropout_p, seed), st));
  PVCR_TRY(grad_w(a, w.logits, w.ldl, M, Vc, w.hs_drop ? w.hs_drop : hs, H, H, nullptr, nullptr, d_wv, H, 0, nsplit, st));
  PVCR_TRY(colsum(w.logits, w.ldl, M, Vc, d_bv, 0, st));
  return PVCR_OK;
}

// y = x * mask / (1 - p) with the mask pvcr_vocab_ce_fwd / _bwd draw for Dropout(hs) under (p, seed): element index =
// flat index into the [B*L, H] matrix.  Used by the materialised-logits backward and by the dropout parity tests.
int out_dropout_apply(const float* x, float* y, long long n, float p, unsigned long long seed, cudaStream_t st) {
  return dropout_apply(x, y, n, out_dropout(p, seed), st);
}

}  // namespace pvcr
