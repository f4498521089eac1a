// Vocabulary projection fused with the loss contract (bf16 mode): the logits never leave the GEMM.
//   forward : logits tile = hs W^T + b in TMEM -> epilogue keeps per-row running (max, sum-exp, arg-max) and the
//             target logit; one partial per (row, N-tile); a row-wise finalize gives lse / nll / pred.
//   backward: the GEMM is recomputed and its epilogue writes d logits = (softmax - onehot) * w_row directly as bf16,
//             both row-major [M, Vcp] (A operand of dH = dlogits W) and transposed [Vc, Mp] (A operand of
//             dW = dlogits^T hs), so no fp32 logits / dlogits tensor and no staging pass over them exists.
// Reference semantics: Dropout + Linear (model/S2VTAttModel.py:145, model/S2VTModel.py:130), calc_masked_loss,
// calc_masked_accuracy, torch.argmax (train_utils.py:37-71, train.py:38).
#include <cstdlib>

#include "../../include/pvcr_b200.h"
#include "host.h"

namespace pvcr {

constexpr int CE_BN = 256;

// Work in the log2 domain: y = (acc + bias) * log2(e), so that exp(x - m) = ex2(y - m2) costs FADD + MUFU.
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct EpiCeFwd {
  static constexpr int SMEM_PER_WARP = 0;
  const float* bias;
  float *pmax, *psum; int* pidx;
  int M, N, nparts;
  float m_run, s_run; int i_run;
  __device__ __forceinline__ void begin(int, int) { m_run = -INFINITY; s_run = 0.f; i_run = 0x7fffffff; }
  __device__ __forceinline__ void chunk(int row, int col0, int, float (&v)[32]) {
    const bool full = col0 + 32 <= N;
    float cm = -INFINITY;
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (full) b4 = __ldg(reinterpret_cast<const float4*>(bias + col0) + j4);
      else {
        const int c = col0 + j4 * 4;
        if (c < N) b4.x = __ldg(bias + c);
        if (c + 1 < N) b4.y = __ldg(bias + c + 1);
        if (c + 2 < N) b4.z = __ldg(bias + c + 2);
        if (c + 3 < N) b4.w = __ldg(bias + c + 3);
      }
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = j4 * 4 + k;
        float y = (v[j] + bb[k]) * LOG2E;
        if (!full && col0 + j >= N) y = -INFINITY;
        v[j] = y;
        cm = fmaxf(cm, y);
      }
    }
    if (cm > m_run) {                       // new running maximum (rare after the first chunks): locate its column
      int ci = 0;
#pragma unroll
      for (int j = 31; j >= 0; --j)
        if (v[j] == cm) ci = j;             // smallest index among ties
      s_run *= fast_ex2(m_run - cm);
      m_run = cm; i_run = col0 + ci;
    }
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 2) { s0 += fast_ex2(v[j] - m_run); s1 += fast_ex2(v[j + 1] - m_run); }
    s_run += s0 + s1;
  }
  __device__ __forceinline__ void end(int row, int part, int) {
    if (row < M) {
      const long long o = (long long)row * nparts + part;
      pmax[o] = m_run; psum[o] = s_run; pidx[o] = i_run;
    }
  }
};

struct EpiCeBwd {
  // A thread owns one row of the accumulator, so a direct store scatters 32 rows per warp instruction (one
  // L1 wavefront per row).  Each epilogue warp instead transposes its 32 x 32 bf16 chunk through 2 KB of shared
  // memory (16-byte slots XOR-swizzled, conflict free both ways) and stores 8 rows x 64 contiguous bytes per
  // instruction: 4x fewer LSU wavefronts, which were as long as the MMA main loop of a K = 512 tile.
  static constexpr int SMEM_PER_WARP = 2048;
  const float* bias; const long long* target; const float *lse2, *roww;      // lse2 = lse * log2(e)
  const float* gscale;                                                       // device scalar d total / d loss (null = 1)
  bf16* D; long long ldD;
  int M, N;
  float l2, w; int t;
  uint4* stage;
  __device__ __forceinline__ void attach(uint8_t* warp_smem) { stage = reinterpret_cast<uint4*>(warp_smem); }
  __device__ __forceinline__ void begin(int row, int) {
    l2 = 0.f; w = 0.f; t = -1;
    if (row < M) { l2 = lse2[row]; w = roww[row] * (gscale ? __ldg(gscale) : 1.f); t = (int)target[row]; }
  }
  __device__ __forceinline__ void chunk(int row, int col0, int, float (&v)[32]) {
    const bool full = col0 + 32 <= N;
    const int tj = t - col0;
    __align__(16) __nv_bfloat162 o[16];
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (full) b4 = __ldg(reinterpret_cast<const float4*>(bias + col0) + j4);
      else {
        const int c = col0 + j4 * 4;
        if (c < N) b4.x = __ldg(bias + c);
        if (c + 1 < N) b4.y = __ldg(bias + c + 1);
        if (c + 2 < N) b4.z = __ldg(bias + c + 2);
        if (c + 3 < N) b4.w = __ldg(bias + c + 3);
      }
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
      float d[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = j4 * 4 + k;
        float pr = fast_ex2((v[j] + bb[k]) * LOG2E - l2);
        if (j == tj) pr -= 1.f;
        d[k] = pr * w;
        if (!full && col0 + j >= N) d[k] = 0.f;
      }
      o[j4 * 2] = __floats2bfloat162_rn(d[0], d[1]);
      o[j4 * 2 + 1] = __floats2bfloat162_rn(d[2], d[3]);
    }
    const int lane = threadIdx.x & 31, row_base = row - lane;
#pragma unroll
    for (int c = 0; c < 4; ++c) stage[lane * 4 + (c ^ ((lane >> 1) & 3))] = reinterpret_cast<const uint4*>(o)[c];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = i * 8 + (lane >> 2), c = lane & 3;
      const uint4 val = stage[r * 4 + (c ^ ((r >> 1) & 3))];
      if (row_base + r < M) *reinterpret_cast<uint4*>(D + (long long)(row_base + r) * ldD + col0 + c * 8) = val;
    }
    __syncwarp();
  }
  __device__ __forceinline__ void end(int, int, int) {}
};

// target logit per row (log2 domain): tgt2[row] = (hs_bf16[row] . W_bf16[target[row]] + b[target[row]]) * log2(e)
__global__ void __launch_bounds__(256) target_logit_kernel(const bf16* hs, long long ld_hs, const bf16* wv, long long ld_w,
                                                           const float* bias, const long long* target, int M, int H,
                                                           float* tgt2) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const long long t = target[row];
  const bf16* a = hs + (long long)row * ld_hs;
  const bf16* b = wv + t * ld_w;
  float s = 0.f;
  for (int k = lane * 2; k < H; k += 64) {
    const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a + k));
    const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(b + k));
    s += x.x * y.x + x.y * y.y;
  }
  s = warp_sum(s);
  if (lane == 0) tgt2[row] = (s + bias[t]) * LOG2E;
}

// Finalize of the fused forward, ONE launch: block b owns the L tokens of video b (one warp per token at a time):
//   * combines the per-tile partials of the GEMM epilogue into lse / nll / arg-max per token,
//   * writes what the backward needs per token: lse in the log2 domain and the row weight
//     (l < s_len[b]) / (min(s_len[b], L) * B) of calc_masked_loss (train_utils.py:47-51),
//   * reduces the video's masked mean loss / #correct / #mask, and the last block to finish adds the B per-video values in
//     video order (deterministic) into loss3 -- the former ce_finalize_rows + loss_finalize + row_weights launches.
constexpr int CE_MAX_L = 256;
__global__ void __launch_bounds__(1024) ce_finalize_kernel(const float* pmax, const float* psum, const int* pidx,
                                                          const float* tgt, const long long* target, const long long* s_len,
                                                          int B, int L, int ntiles, float* lse_out, float* nll_out,
                                                          long long* pred_out, float* lse2_out, float* roww_out,
                                                          float* token_nll, float* partial, unsigned* counter, float* loss3) {
  __shared__ float s_nll[CE_MAX_L];
  __shared__ float s_ok[CE_MAX_L];
  __shared__ bool last;
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const long long len = s_len[b];
  const float cnt = (float)(len < L ? len : L);
  for (int l = warp; l < L; l += nwarps) {        // one warp per token (L <= 32: all tokens of the video in parallel)
    const int row = b * L + l;
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
      s += psum[(long long)row * ntiles + k] * exp2f(pmax[(long long)row * ntiles + k] - m);
    s = warp_sum(s);
    if (lane == 0) {
      const float lse2 = m + log2f(s);              // log2 domain (see EpiCeFwd)
      const float nll = (lse2 - tgt[row]) * LN2;
      const long long pr = mi == 0x7fffffff ? 0 : mi;
      lse_out[row] = lse2 * LN2;
      nll_out[row] = nll;
      pred_out[row] = pr;
      lse2_out[row] = lse2;
      roww_out[row] = l < len ? 1.f / (cnt * (float)B) : 0.f;
      if (token_nll) token_nll[row] = nll;
      s_nll[l] = nll;
      s_ok[l] = pr == target[row] ? 1.f : 0.f;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sum = 0.f, corr = 0.f, n = 0.f;
    for (int l = 0; l < L && l < len; ++l) { sum += s_nll[l]; corr += s_ok[l]; n += 1.f; }
    partial[3 * b] = sum / cnt; partial[3 * b + 1] = corr; partial[3 * b + 2] = n;
    __threadfence();
    last = atomicAdd(counter, 1u) == (unsigned)(B - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float loss = 0.f, corr = 0.f, n = 0.f;
    for (int i = 0; i < B; ++i) {
      loss += __ldcg(partial + 3 * i); corr += __ldcg(partial + 3 * i + 1); n += __ldcg(partial + 3 * i + 2);
    }
    loss3[0] = loss / (float)B; loss3[1] = corr; loss3[2] = n;
    *counter = 0u;                                  // ready for the next launch (stream order)
  }
}

// out[c] += sum_r in[r*ld + c]  (bf16 in, fp32 accumulate; out pre-zeroed): block = 64 column pairs x 4 row lanes
// over a chunk of CS_ROWS rows, chunk partials merged with atomics
constexpr int CS_ROWS = 256;
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* in, long long ld, int R, int C, float* out) {
  __shared__ float2 part[4][64];
  const int cp = blockIdx.x * 64 + (threadIdx.x & 63), rl = threadIdx.x >> 6, c = cp * 2;
  const int r_begin = blockIdx.y * CS_ROWS, r_end = min(R, r_begin + CS_ROWS);
  float2 s = make_float2(0.f, 0.f);
  if (c < C) {
#pragma unroll 8
    for (int r = r_begin + rl; r < r_end; r += 4) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(in + (long long)r * ld + c));
      s.x += f.x; s.y += f.y;
    }
  }
  part[rl][threadIdx.x & 63] = s;
  __syncthreads();
  if (rl == 0 && c < C) {
    float2 t = part[0][threadIdx.x];
    for (int k = 1; k < 4; ++k) { t.x += part[k][threadIdx.x].x; t.y += part[k][threadIdx.x].y; }
    atomicAdd(out + c, t.x);
    if (c + 1 < C) atomicAdd(out + c + 1, t.y);
  }
}

// PVCR_CE_MC=1: 2-CTA clusters with the W_v tile multicast to both CTAs (gemm_tn_mc2_kernel).  Parity-green, but
// measured equal to the 1-CTA kernel (fwd 101 us, step 2.332 vs 2.333 ms): the K = 512 product is limited by what one
// SM can take in (48 KB of operands per 512 tensor-pipe cycles = 96 B/clk), which multicast does not change - only a
// cta_group::2 tile (each SM holds half of B) would.  Kept as the stepping stone to that kernel; off by default.
static bool ce_multicast() {
  static const bool on = getenv("PVCR_CE_MC") != nullptr;
  return on;
}

// PVCR_CE_2SM=1: CTA-pair (tcgen05 cta_group::2, M = 256) tiles for the two vocabulary GEMMs (gemm_tn_2sm_kernel).
// Parity-green; measured SLOWER than the 1-CTA kernel at this shape (the two GEMMs together +0.11 ms with 16 epilogue
// warps each, 6 x 32 KB stages): kept off, as the base for larger-K products where operand traffic does bind.
static bool ce_pair() {
  static const bool on = getenv("PVCR_CE_2SM") != nullptr;
  return on;
}

// epilogue warps of the 1-CTA fused kernels: 16 (four per SM sub-partition) unless PVCR_CE_EW8 is set (A/B knob)
static bool ce_ew16() {
  static const bool off = getenv("PVCR_CE_EW8") != nullptr;
  return !off && !ce_multicast();
}
static int ce_parts() { return (ce_ew16() || ce_pair()) ? 4 : 2; }

struct FusedWs {
  Planes hs_a, wv;
  float *pmax, *psum, *tgt, *nll, *roww, *lse2, *partial;
  unsigned* counter;              // zero between launches (the finalize kernel resets it); zeroed by the first forward
  int* pidx;
  bf16* D;
  long long ldD;
  int ntiles;
};
static void carve_fused(Arena& a, int M, int H, int Vc, FusedWs& w) {
  w.ntiles = ce_parts() * cdiv(Vc, CE_BN);     // partials per (N tile, column part) of the persistent GEMM epilogue
  w.hs_a = alloc_planes(a, M, H, 1);
  w.wv = alloc_planes(a, Vc, H, 1);
  w.pmax = a.alloc<float>((size_t)M * w.ntiles); w.psum = a.alloc<float>((size_t)M * w.ntiles);
  w.pidx = a.alloc<int>((size_t)M * w.ntiles);
  w.tgt = a.alloc<float>(M); w.nll = a.alloc<float>(M); w.roww = a.alloc<float>(M); w.lse2 = a.alloc<float>(M);
  w.partial = a.alloc<float>((size_t)3 * M); w.counter = a.alloc<unsigned>(32);
  w.ldD = (long long)cdiv(Vc, CE_BN) * CE_BN;
  w.D = a.alloc<bf16>((size_t)M * w.ldD);
}
size_t vocab_fused_workspace(int M, int H, int Vc) {
  Arena a(nullptr, 0);
  FusedWs w;
  carve_fused(a, M, H, Vc, w);
  return a.off + 4096;
}

static Dropout fused_out_dropout(float p, unsigned long long seed) { return make_dropout(p, seed, 0x5000000000ull); }

// Stage the vocabulary weights (fp32 -> bf16 planes) of a coming vocab_fused_fwd on a side lane: a caller that knows
// the projection follows (the tape-free train step) issues this before the encoder / decoder sweeps, which takes the
// 71 MB cast off the critical path between the decoder sweep and the projection.  Leaves a note on the workspace.
int vocab_fused_prepare(const float* wv, int B, int L, int H, int Vc, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (side_mode() == 0) return PVCR_OK;
  Arena a(ws, ws_bytes);
  FusedWs w;
  carve_fused(a, B * L, H, Vc, w);
  if (a.failed) { set_last_error("vocab_fused_prepare: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  cudaStream_t lane;
  PVCR_TRY(side_fork(st, &lane, 2));
  PVCR_TRY(prep_weight(wv, H, Vc, H, w.wv, lane));
  side_note_put(ws, NOTE_VOCAB_WV, wv);
  return PVCR_OK;
}

int vocab_fused_fwd(const float* hs, const float* wv, const float* bv, const long long* target, const long long* s_len,
                    int B, int L, int H, int Vc, float dropout_p, unsigned long long seed, float* loss3, long long* pred,
                    float* lse, float* token_nll, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int M = B * L;
  Arena a(ws, ws_bytes);
  FusedWs w;
  carve_fused(a, M, H, Vc, w);
  if (a.failed) { set_last_error("vocab_fused_fwd: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  PVCR_TRY(stage(hs, H, M, H, w.hs_a, 0, nullptr, fused_out_dropout(dropout_p, seed), st));
  if (side_note_take(ws, NOTE_VOCAB_WV, wv)) PVCR_TRY(side_join(st));       // staged by vocab_fused_prepare on a lane
  else PVCR_TRY(prep_weight(wv, H, Vc, H, w.wv, st));
  // the target logit of every token only feeds the finalize: it runs on a side lane next to the projection GEMM
  cudaStream_t lt = st;
  if (side_site(4)) PVCR_TRY(side_fork(st, &lt, 1));
  {
    LaunchScope ls_(KC_LOSS, lt);
    target_logit_kernel<<<cdiv((long long)M * 32, 256), 256, 0, lt>>>(w.hs_a.ptr, w.hs_a.ld, w.wv.ptr, w.wv.ld, bv, target,
                                                                      M, H, w.tgt);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  PVCR_CUDA_CHECK(cudaMemsetAsync(w.counter, 0, sizeof(unsigned), lt));
  EpiCeFwd epi{};
  epi.bias = bv; epi.pmax = w.pmax; epi.psum = w.psum; epi.pidx = w.pidx;
  epi.M = M; epi.N = Vc; epi.nparts = w.ntiles;
  GemmCoords gc{M, Vc, (int)w.hs_a.ld, 0, 0, 0, 0};
  if (ce_pair()) PVCR_TRY((launch_gemm_tn_2sm<CE_BN, 6, EpiCeFwd, 16>(w.hs_a.view(), w.wv.view(), gc, epi, st)));
  else if (ce_multicast()) PVCR_TRY((launch_gemm_tn_mc2<CE_BN, 4, EpiCeFwd>(w.hs_a.view(), w.wv.view(), gc, epi, st)));
  else if (ce_ew16()) PVCR_TRY((launch_gemm_tn_persistent<CE_BN, 4, EpiCeFwd, false, false, 16>(w.hs_a.view(), w.wv.view(), gc, 1, epi, st)));
  else PVCR_TRY((launch_gemm_tn_persistent<CE_BN, 4, EpiCeFwd>(w.hs_a.view(), w.wv.view(), gc, 1, epi, st)));
  if (lt != st) PVCR_TRY(side_join_lane(st, 1));
  PVCR_REQUIRE(L <= CE_MAX_L, "vocab_fused_fwd: L=%d > %d", L, CE_MAX_L);
  {
    LaunchScope ls_(KC_LOSS, st);
    ce_finalize_kernel<<<B, 32 * (L < 32 ? L : 32), 0, st>>>(w.pmax, w.psum, w.pidx, w.tgt, target, s_len, B, L, w.ntiles, lse, w.nll, pred,
                                          w.lse2, w.roww, token_nll, w.partial, w.counter, loss3);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
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
  // recompute the logits tile by tile; the epilogue emits bf16 dlogits (row-major and transposed).  lse (log2 domain)
  // and the row weights of calc_masked_loss were left in the workspace by the forward's finalize kernel.
  EpiCeBwd epi{};
  epi.bias = bv; epi.target = target; epi.lse2 = w.lse2; epi.roww = w.roww; epi.gscale = gscale;
  epi.D = w.D; epi.ldD = w.ldD; epi.M = M; epi.N = Vc;
  if (w.ldD > Vc)      // chunks lying entirely past Vc are skipped by the GEMM epilogue: their K-padding must read 0
    PVCR_CUDA_CHECK(cudaMemset2DAsync(w.D + Vc, sizeof(bf16) * w.ldD, 0, sizeof(bf16) * (w.ldD - Vc), M, st));
  GemmCoords gc{M, Vc, (int)w.hs_a.ld, 0, 0, 0, 0};
  if (ce_pair()) PVCR_TRY((launch_gemm_tn_2sm<CE_BN, 6, EpiCeBwd, 16>(w.hs_a.view(), w.wv.view(), gc, epi, st)));
  else if (ce_multicast()) PVCR_TRY((launch_gemm_tn_mc2<CE_BN, 4, EpiCeBwd>(w.hs_a.view(), w.wv.view(), gc, epi, st)));
  else if (ce_ew16()) PVCR_TRY((launch_gemm_tn_persistent<CE_BN, 4, EpiCeBwd, false, false, 16>(w.hs_a.view(), w.wv.view(), gc, 1, epi, st)));
  else PVCR_TRY((launch_gemm_tn_persistent<CE_BN, 4, EpiCeBwd>(w.hs_a.view(), w.wv.view(), gc, 1, epi, st)));
  // d W = dlogits^T Dropout(hs) and d b = column sums of dlogits do not feed the rest of the backward: they run on
  // the side lane next to the d hs product (60 output tiles) and whatever the caller enqueues next.
  // Both operands as they are (row-major bf16, MN-major tcgen05 operands); hs_a are the planes staged (with the
  // dropout mask) by the forward pass.
  cudaStream_t lane = st, lane1 = st;
  if (side_site(1)) { PVCR_TRY(side_fork(st, &lane, 0)); PVCR_TRY(side_fork(st, &lane1, 1)); }
  // d hs = dlogits W: A = dlogits (K-major, padding columns written as zeros), B = the forward weight planes [Vc, H]
  // as an MN-major operand (rows past Vc read as zero through the tensor map): no W^T copy
  {
    OperandView dv{w.D, w.ldD, 0, M, 1};
    PVCR_TRY(gemm_kn_store(dv, w.wv.view(), M, H, Vc, d_hs, H, 0, st));
  }
  if (dropout_p > 0.f) PVCR_TRY(dropout_apply(d_hs, d_hs, (long long)M * H, dr, st));
  {
    // A/B knob: cap the CTAs of d W so that it leaves the d hs product (critical path) most of the machine and goes on
    // in the shadow of the decoder sweep that follows (<= SMs the sweep leaves free, or the sweep waits for it)
    static const int dwv_cap = getenv("PVCR_DWV_CAP") ? atoi(getenv("PVCR_DWV_CAP")) : 0;
    CtaCap cap_(lane != st ? dwv_cap : 0);
    OperandView dv{w.D, w.ldD, 0, M, 1};
    PVCR_TRY(gemm_mn_store(dv, w.hs_a.view(), Vc, H, M, d_wv, H, 0, lane));
  }
  PVCR_CUDA_CHECK(cudaMemsetAsync(d_bv, 0, sizeof(float) * Vc, lane1));
  {
    LaunchScope ls_(KC_LOSS, lane1);
    colsum_bf16_kernel<<<dim3(cdiv(cdiv(Vc, 2), 64), cdiv(M, CS_ROWS)), 256, 0, lane1>>>(w.D, w.ldD, M, Vc, d_bv);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  PVCR_TRY(side_call_end(st));
  return PVCR_OK;
}

}  // namespace pvcr
