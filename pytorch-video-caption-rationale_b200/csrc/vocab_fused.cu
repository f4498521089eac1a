// Vocabulary projection fused with the loss contract (bf16 mode): the logits never leave the GEMM.
//   forward : logits tile = hs W^T + b in TMEM -> epilogue keeps per-row running (max, sum-exp, arg-max) and the
//             target logit; one partial per (row, N-tile); a row-wise finalize gives lse / nll / pred.
//   backward: the GEMM is recomputed and its epilogue writes d logits = (softmax - onehot) * w_row directly as bf16,
//             both row-major [M, Vcp] (A operand of dH = dlogits W) and transposed [Vc, Mp] (A operand of
//             dW = dlogits^T hs), so no fp32 logits / dlogits tensor and no staging pass over them exists.
// Reference semantics: Dropout + Linear (model/S2VTAttModel.py:145, model/S2VTModel.py:130), calc_masked_loss,
// calc_masked_accuracy, torch.argmax (train_utils.py:37-71, train.py:38).
#include "../../include/pvcr_b200.h"
#include "host.h"

namespace pvcr {

constexpr int CE_BN = 128;

struct EpiCeFwd {
  const float* bias; const long long* target;
  float *pmax, *psum, *tgt; int* pidx;
  int M, N, ntiles;
  float m_run, s_run; int i_run; long long t;
  __device__ __forceinline__ void begin(int row, int) {
    m_run = -INFINITY; s_run = 0.f; i_run = 0x7fffffff;
    t = row < M ? target[row] : -1;
  }
  __device__ __forceinline__ void chunk(int row, int col0, int, float (&v)[32]) {
    float cm = -INFINITY; int ci = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = col0 + j;
      float x = -INFINITY;
      if (col < N) {
        x = v[j] + __ldg(bias + col);
        if (col == t) tgt[row] = x;
      }
      v[j] = x;
      if (x > cm) { cm = x; ci = col; }
    }
    if (cm > m_run) { s_run *= __expf(m_run - cm); m_run = cm; i_run = ci; }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += __expf(v[j] - m_run);       // exp(-inf) = 0 for the padding columns
    s_run += s;
  }
  __device__ __forceinline__ void end(int row, int tile, int) {
    if (row < M) {
      const long long o = (long long)row * ntiles + tile;
      pmax[o] = m_run; psum[o] = s_run; pidx[o] = i_run;
    }
  }
};

struct EpiCeBwd {
  const float* bias; const long long* target; const float *lse, *roww;
  bf16* D; long long ldD; bf16* DT; long long ldDT;
  int M, N, Mp;
  float l, w; long long t;
  __device__ __forceinline__ void begin(int row, int) {
    l = 0.f; w = 0.f; t = -1;
    if (row < M) { l = lse[row]; w = roww[row]; t = target[row]; }
  }
  __device__ __forceinline__ void chunk(int row, int col0, int, float (&v)[32]) {
    __align__(16) bf16 o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = col0 + j;
      float d = 0.f;
      if (col < N && w != 0.f) d = (__expf(v[j] + __ldg(bias + col) - l) - (col == t ? 1.f : 0.f)) * w;
      o[j] = __float2bfloat16_rn(d);
    }
    if (row < M) {
      uint4* dst = reinterpret_cast<uint4*>(D + (long long)row * ldD + col0);
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[j] = reinterpret_cast<const uint4*>(o)[j];
    }
    if (row < Mp) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < N) DT[(long long)(col0 + j) * ldDT + row] = o[j];
    }
  }
  __device__ __forceinline__ void end(int, int, int) {}
};

// one warp per row: combine the per-tile partials
__global__ void __launch_bounds__(256) ce_finalize_rows_kernel(const float* pmax, const float* psum, const int* pidx,
                                                               const float* tgt, int M, int ntiles, float* lse_out,
                                                               float* nll_out, long long* pred_out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  float m = -INFINITY; int mi = 0x7fffffff;
  for (int k = lane; k < ntiles; k += 32) {
    const float v = pmax[(long long)row * ntiles + k];
    const int i = pidx[(long long)row * ntiles + k];
    if (v > m || (v == m && i < mi)) { m = v; mi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
  }
  float s = 0.f;
  for (int k = lane; k < ntiles; k += 32)
    s += psum[(long long)row * ntiles + k] * expf(pmax[(long long)row * ntiles + k] - m);
  s = warp_sum(s);
  if (lane == 0) {
    const float lse = m + logf(s);
    lse_out[row] = lse;
    nll_out[row] = lse - tgt[row];
    pred_out[row] = mi == 0x7fffffff ? 0 : mi;
  }
}

// w[row] = (l < s_len[b]) / (s_len[b] * B) * gscale
__global__ void row_weights_kernel(const long long* s_len, int B, int L, const float* gscale, float* w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * L) return;
  const int b = i / L, l = i % L;
  const long long len = s_len[b];
  w[i] = (l < len ? 1.f / ((float)len * (float)B) : 0.f) * (gscale ? gscale[0] : 1.f);
}

// out[r] = sum_c in[r*ld + c]  (bf16 in, fp32 accumulate): one warp per row
__global__ void __launch_bounds__(256) rowsum_bf16_kernel(const bf16* in, long long ld, int R, int C, float* out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= R) return;
  const bf16* x = in + (long long)row * ld;
  float s = 0.f;
  const int C8 = C & ~7;
  for (int c = lane * 8; c < C8; c += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + c);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); s += f.x + f.y; }
  }
  for (int c = C8 + lane; c < C; c += 32) s += __bfloat162float(x[c]);
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}

struct FusedWs {
  Planes hs_a, wv, wvT, hsT;
  float *pmax, *psum, *tgt, *nll, *roww;
  int* pidx;
  bf16 *D, *DT;
  long long ldD, ldDT;
  int ntiles;
};
static void carve_fused(Arena& a, int M, int H, int Vc, FusedWs& w) {
  w.ntiles = cdiv(Vc, CE_BN);
  w.hs_a = alloc_planes(a, M, H, 1);
  w.wv = alloc_planes(a, Vc, H, 1);
  w.pmax = a.alloc<float>((size_t)M * w.ntiles); w.psum = a.alloc<float>((size_t)M * w.ntiles);
  w.pidx = a.alloc<int>((size_t)M * w.ntiles);
  w.tgt = a.alloc<float>(M); w.nll = a.alloc<float>(M); w.roww = a.alloc<float>(M);
  w.wvT = alloc_planes(a, H, Vc, 1);
  w.hsT = alloc_planes(a, H, M, 1);
  w.ldD = (long long)w.ntiles * CE_BN;
  w.ldDT = w.hsT.Kp;
  w.D = a.alloc<bf16>((size_t)M * w.ldD);
  w.DT = a.alloc<bf16>((size_t)Vc * w.ldDT);
}
size_t vocab_fused_workspace(int M, int H, int Vc) {
  Arena a(nullptr, 0);
  FusedWs w;
  carve_fused(a, M, H, Vc, w);
  return a.off + 4096;
}

static Dropout fused_out_dropout(float p, unsigned long long seed) { return Dropout{p, seed, 0x5000000000ull}; }

int vocab_fused_fwd(const float* hs, const float* wv, const float* bv, const long long* target, const long long* s_len,
                    int B, int L, int H, int Vc, float dropout_p, unsigned long long seed, float* loss3, long long* pred,
                    float* lse, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int M = B * L;
  Arena a(ws, ws_bytes);
  FusedWs w;
  carve_fused(a, M, H, Vc, w);
  if (a.failed) { set_last_error("vocab_fused_fwd: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  PVCR_TRY(stage(hs, H, M, H, w.hs_a, 0, nullptr, fused_out_dropout(dropout_p, seed), st));
  PVCR_TRY(prep_weight(wv, H, Vc, H, w.wv, st));
  EpiCeFwd epi{};
  epi.bias = bv; epi.target = target; epi.pmax = w.pmax; epi.psum = w.psum; epi.tgt = w.tgt; epi.pidx = w.pidx;
  epi.M = M; epi.N = Vc; epi.ntiles = w.ntiles;
  GemmCoords gc{M, Vc, (int)w.hs_a.ld, 0, 0, 0, 0};
  PVCR_TRY((launch_gemm_tn<CE_BN, 3, EpiCeFwd>(w.hs_a.view(), w.wv.view(), gc, 1, epi, st)));
  {
    LaunchScope ls_(KC_LOSS, st);
    ce_finalize_rows_kernel<<<cdiv((long long)M * 32, 256), 256, 0, st>>>(w.pmax, w.psum, w.pidx, w.tgt, M, w.ntiles, lse,
                                                                          w.nll, pred);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return loss_finalize(w.nll, pred, target, s_len, B, L, loss3, st);
}

int vocab_fused_bwd(const float* hs, const float* wv, const float* bv, const long long* target, const long long* s_len,
                    int B, int L, int H, int Vc, float dropout_p, unsigned long long seed, const float* gscale,
                    float* d_hs, float* d_wv, float* d_bv, const float* lse, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int M = B * L;
  Arena a(ws, ws_bytes);
  FusedWs w;
  carve_fused(a, M, H, Vc, w);
  if (a.failed) { set_last_error("vocab_fused_bwd: workspace too small"); return PVCR_ERR_WORKSPACE; }
  const Dropout dr = fused_out_dropout(dropout_p, seed);
  {
    LaunchScope ls_(KC_LOSS, st);
    row_weights_kernel<<<cdiv(M, 256), 256, 0, st>>>(s_len, B, L, gscale, w.roww);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  // recompute the logits tile by tile; the epilogue emits bf16 dlogits (row-major and transposed)
  EpiCeBwd epi{};
  epi.bias = bv; epi.target = target; epi.lse = lse; epi.roww = w.roww;
  epi.D = w.D; epi.ldD = w.ldD; epi.DT = w.DT; epi.ldDT = w.ldDT; epi.M = M; epi.N = Vc; epi.Mp = (int)w.ldDT;
  if (w.ldD > Vc)      // chunks lying entirely past Vc are skipped by the GEMM epilogue: their K-padding must read 0
    PVCR_CUDA_CHECK(cudaMemset2DAsync(w.D + Vc, sizeof(bf16) * w.ldD, 0, sizeof(bf16) * (w.ldD - Vc), M, st));
  GemmCoords gc{M, Vc, (int)w.hs_a.ld, 0, 0, 0, 0};
  PVCR_TRY((launch_gemm_tn<CE_BN, 3, EpiCeBwd>(w.hs_a.view(), w.wv.view(), gc, 1, epi, st)));
  // d hs = dlogits W  (K = Vc padded to 64; the padding columns of D are zeros, of W^T planes too)
  PVCR_TRY(prep_weight_T(wv, H, Vc, H, w.wvT, 0, 1, st));
  {
    OperandView dv{w.D, w.ldD, 0, M, 1};
    PVCR_TRY(gemm_planes(dv, w.wvT.view(), M, H, w.wvT.Kp, d_hs, H, nullptr, 0, st));
  }
  if (dropout_p > 0.f) PVCR_TRY(dropout_apply(d_hs, d_hs, (long long)M * H, dr, st));
  // d W = dlogits^T Dropout(hs)
  PVCR_TRY(transpose_split(hs, H, M, H, w.hsT.ptr, w.hsT.ld, w.hsT.Kp, 0, 1, 1, 1, nullptr, nullptr, st, dr));
  {
    OperandView dt{w.DT, w.ldDT, 0, Vc, 1};
    PVCR_TRY(gemm_planes(dt, w.hsT.view(), Vc, H, w.hsT.Kp, d_wv, H, nullptr, 0, st));
  }
  {
    LaunchScope ls_(KC_LOSS, st);
    rowsum_bf16_kernel<<<cdiv((long long)Vc * 32, 256), 256, 0, st>>>(w.DT, w.ldDT, Vc, M, d_bv);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

}  // namespace pvcr
