// Beam-search bookkeeping kernels.  The search itself is defined in
// oracle/captioning_oracle.py: s2vtatt_beam_search (fixed max_len steps like the reference's greedy eval branch,
// model/S2VTAttModel.py:172-191; score = sum of log-softmax; the K best of a video's K x Vc candidates survive, ties to
// the lower flat index beam * Vc + word), so that beam 1 is the reference's greedy decoding.
#include "common.cuh"
#include "kernels.cuh"

namespace pvcr {

constexpr int BEAM_THREADS = 256;
constexpr int BEAM_MAX_K = 8;

// out[(b*K + k), :] = in[b, :]   (rows of `len` floats)
__global__ void repeat_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int K, long long len) {
  const long long r = blockIdx.x;
  const float* src = in + (r / K) * len;
  float* dst = out + r * len;
  for (long long i = threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
}

struct BestCand { float v; int idx; };
__device__ __forceinline__ bool better(float v, int idx, float bv, int bidx) { return v > bv || (v == bv && idx < bidx); }

// One block per video.  logits [B*K, ld] (row b*K + k); score_in [B*K].  Writes score_out / parent / word [B*K].
__global__ void __launch_bounds__(BEAM_THREADS) beam_select_kernel(const float* __restrict__ logits, long long ld, int Vc,
                                                                   int K, int first, const float* __restrict__ score_in,
                                                                   float* __restrict__ score_out, int* __restrict__ parent,
                                                                   long long* __restrict__ word) {
  __shared__ float s_red[BEAM_THREADS / 32];
  __shared__ int s_redi[BEAM_THREADS / 32];
  __shared__ float s_lse[BEAM_MAX_K], s_score[BEAM_MAX_K];
  __shared__ int s_chosen[BEAM_MAX_K];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // log-sum-exp of every beam's row
  for (int k = 0; k < K; ++k) {
    const float* x = logits + (long long)(b * K + k) * ld;
    float m = -INFINITY;
    for (int v = tid; v < Vc; v += BEAM_THREADS) m = fmaxf(m, x[v]);
    m = warp_max(m);
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    m = s_red[0];
    for (int w = 1; w < BEAM_THREADS / 32; ++w) m = fmaxf(m, s_red[w]);
    __syncthreads();
    float s = 0.f;
    for (int v = tid; v < Vc; v += BEAM_THREADS) s += expf(x[v] - m);
    s = warp_sum(s);
    if (lane == 0) s_red[warp] = s;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < BEAM_THREADS / 32; ++w) t += s_red[w];
      s_lse[k] = m + logf(t);
      s_score[k] = score_in[b * K + k];
    }
    __syncthreads();
  }
  const int live = first ? 1 : K;                 // at the first step every beam is the same hypothesis: only beam 0 counts
  const int total = live * Vc;
  for (int j = 0; j < K; ++j) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = tid; c < total; c += BEAM_THREADS) {
      bool taken = false;
      for (int q = 0; q < j; ++q) taken |= (s_chosen[q] == c);
      if (taken) continue;
      const int k = c / Vc, v = c - k * Vc;
      const float val = s_score[k] + (logits[(long long)(b * K + k) * ld + v] - s_lse[k]);
      if (better(val, c, bv, bi)) { bv = val; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_red[warp] = bv; s_redi[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < BEAM_THREADS / 32; ++w)
        if (better(s_red[w], s_redi[w], bv, bi)) { bv = s_red[w]; bi = s_redi[w]; }
      s_chosen[j] = bi;
      const int r = b * K + j;
      if (bi == 0x7fffffff) { score_out[r] = -INFINITY; parent[r] = 0; word[r] = 0; }      // fewer candidates than beams
      else { score_out[r] = bv; parent[r] = bi / Vc; word[r] = bi % Vc; }
    }
    __syncthreads();
  }
}

// Row r = b*K + j takes over hypothesis parent[r] of its video: state and history; the new word is appended.
__global__ void beam_reorder_kernel(const float* __restrict__ h_in, float* __restrict__ h_out, int H,
                                    const long long* __restrict__ hist_in, long long* __restrict__ hist_out, int L, int step,
                                    const int* __restrict__ parent, const long long* __restrict__ word, int K) {
  const int r = blockIdx.x, src = (r / K) * K + parent[r];
  for (int i = threadIdx.x; i < H; i += blockDim.x) h_out[(long long)r * H + i] = h_in[(long long)src * H + i];
  for (int i = threadIdx.x; i < L; i += blockDim.x)
    hist_out[(long long)r * L + i] = i == step ? word[r] : (i < step ? hist_in[(long long)src * L + i] : 0);
}

int repeat_rows(const float* in, float* out, int rows_in, int K, long long len, cudaStream_t st) {
  if (rows_in == 0 || len == 0) return PVCR_OK;
  { LaunchScope ls_(KC_MISC, st);
  repeat_rows_kernel<<<rows_in * K, 256, 0, st>>>(in, out, K, len);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}
int beam_select(const float* logits, long long ld, int B, int Vc, int K, int first, const float* score_in, float* score_out,
                int* parent, long long* word, cudaStream_t st) {
  PVCR_REQUIRE(K >= 1 && K <= BEAM_MAX_K, "beam_select: beam width %d not in 1..%d", K, BEAM_MAX_K);
  PVCR_REQUIRE((long long)K * Vc < 0x7fffffff, "beam_select: K * Vc too large");
  { LaunchScope ls_(KC_MISC, st);
  beam_select_kernel<<<B, BEAM_THREADS, 0, st>>>(logits, ld, Vc, K, first, score_in, score_out, parent, word);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}
int beam_reorder(const float* h_in, float* h_out, int R, int H, const long long* hist_in, long long* hist_out, int L, int step,
                 const int* parent, const long long* word, int K, cudaStream_t st) {
  { LaunchScope ls_(KC_MISC, st);
  beam_reorder_kernel<<<R, 128, 0, st>>>(h_in, h_out, H, hist_in, hist_out, L, step, parent, word, K);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

}  // namespace pvcr
