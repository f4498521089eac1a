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

__device__ __forceinline__ bool better(float v, int idx, float bv, int bidx) { return v > bv || (v == bv && idx < bidx); }

// lse[r] = log sum exp of logits row r; one block per row
__global__ void __launch_bounds__(BEAM_THREADS) beam_lse_kernel(const float* __restrict__ logits, long long ld, int Vc,
                                                                float* __restrict__ lse) {
  __shared__ float s_red[BEAM_THREADS / 32];
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = logits + (long long)r * ld;
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
    lse[r] = m + logf(t);
  }
}

// Block-wide selection of the K best (value desc, flat index asc) among the per-thread sorted lists tv / ti (K entries,
// unused ones -inf / INT_MAX): K rounds, every thread offers the head of its list, the winner pops it.
__device__ __forceinline__ void block_top_k(float (&tv)[BEAM_MAX_K], int (&ti)[BEAM_MAX_K], int K, float* s_v, int* s_i,
                                            float* out_v, int* out_i) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int j = 0; j < K; ++j) {
    float bv = tv[0];
    int bi = ti[0];
    const int mi = bi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
    __syncthreads();
    bv = s_v[0]; bi = s_i[0];
    for (int w = 1; w < nw; ++w)
      if (better(s_v[w], s_i[w], bv, bi)) { bv = s_v[w]; bi = s_i[w]; }
    if (bi != 0x7fffffff && mi == bi) {            // this thread owned the winner (flat indices are unique): pop it
#pragma unroll
      for (int q = 0; q + 1 < BEAM_MAX_K; ++q) { tv[q] = tv[q + 1]; ti[q] = ti[q + 1]; }
      tv[BEAM_MAX_K - 1] = -INFINITY; ti[BEAM_MAX_K - 1] = 0x7fffffff;
    }
    if (tid == 0) { out_v[j] = bv; out_i[j] = bi; }
    __syncthreads();
  }
}
// per-thread list of the BEAM_MAX_K best seen so far, sorted (value desc, flat index asc); static register indexing
__device__ __forceinline__ void list_insert(float (&tv)[BEAM_MAX_K], int (&ti)[BEAM_MAX_K], float v, int idx) {
  if (!better(v, idx, tv[BEAM_MAX_K - 1], ti[BEAM_MAX_K - 1])) return;
  tv[BEAM_MAX_K - 1] = v; ti[BEAM_MAX_K - 1] = idx;
#pragma unroll
  for (int q = BEAM_MAX_K - 1; q > 0; --q) {
    if (better(tv[q], ti[q], tv[q - 1], ti[q - 1])) {
      const float fv = tv[q]; tv[q] = tv[q - 1]; tv[q - 1] = fv;
      const int fi = ti[q]; ti[q] = ti[q - 1]; ti[q - 1] = fi;
    }
  }
}

// Stage 1: CTA (g, b) scans slice g of video b's (live beams x Vc) candidates in ONE pass (per-thread top-K lists) and
// writes its K best to part_v / part_i [b][g][K].
__global__ void __launch_bounds__(BEAM_THREADS) beam_partial_kernel(const float* __restrict__ logits, long long ld, int Vc,
                                                                    int K, int first, const float* __restrict__ score_in,
                                                                    const float* __restrict__ lse, float* __restrict__ part_v,
                                                                    int* __restrict__ part_i) {
  __shared__ float s_v[BEAM_THREADS / 32];
  __shared__ int s_i[BEAM_THREADS / 32];
  const int g = blockIdx.x, G = gridDim.x, b = blockIdx.y;
  const int total = (first ? 1 : K) * Vc;          // at the first step every beam is the same hypothesis: only beam 0 counts
  const int chunk = (total + G - 1) / G, c0 = g * chunk, c1 = min(total, c0 + chunk);
  float tv[BEAM_MAX_K];
  int ti[BEAM_MAX_K];
#pragma unroll
  for (int q = 0; q < BEAM_MAX_K; ++q) { tv[q] = -INFINITY; ti[q] = 0x7fffffff; }
  for (int c = c0 + threadIdx.x; c < c1; c += BEAM_THREADS) {
    const int k = c / Vc, v = c - k * Vc, r = b * K + k;
    const float val = score_in[r] + (logits[(long long)r * ld + v] - lse[r]);
    list_insert(tv, ti, val, c);
  }
  block_top_k(tv, ti, K, s_v, s_i, part_v + ((long long)b * G + g) * K, part_i + ((long long)b * G + g) * K);
}

// Stage 2: one CTA per video merges the G x K partial winners.
__global__ void __launch_bounds__(64) beam_merge_kernel(const float* __restrict__ part_v, const int* __restrict__ part_i, int G,
                                                        int Vc, int K, float* __restrict__ score_out, int* __restrict__ parent,
                                                        long long* __restrict__ word) {
  __shared__ float s_v[2];
  __shared__ int s_i[2];
  __shared__ float o_v[BEAM_MAX_K];
  __shared__ int o_i[BEAM_MAX_K];
  const int b = blockIdx.x;
  float tv[BEAM_MAX_K];
  int ti[BEAM_MAX_K];
#pragma unroll
  for (int q = 0; q < BEAM_MAX_K; ++q) { tv[q] = -INFINITY; ti[q] = 0x7fffffff; }
  for (int c = threadIdx.x; c < G * K; c += 64) {
    const int idx = part_i[(long long)b * G * K + c];
    if (idx != 0x7fffffff) list_insert(tv, ti, part_v[(long long)b * G * K + c], idx);
  }
  block_top_k(tv, ti, K, s_v, s_i, o_v, o_i);
  if (threadIdx.x < K) {
    const int r = b * K + threadIdx.x, bi = o_i[threadIdx.x];
    if (bi == 0x7fffffff) { score_out[r] = -INFINITY; parent[r] = 0; word[r] = 0; }      // fewer candidates than beams
    else { score_out[r] = o_v[threadIdx.x]; parent[r] = bi / Vc; word[r] = bi % Vc; }
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
// scratch: R floats (lse) + B * BEAM_SLICES * K (float + int) partial winners
constexpr int BEAM_SLICES = 16;
size_t beam_select_scratch(int B, int K) { return sizeof(float) * ((size_t)B * K + (size_t)B * BEAM_SLICES * K * 2) + 64; }
int beam_select(const float* logits, long long ld, int B, int Vc, int K, int first, const float* score_in, float* score_out,
                int* parent, long long* word, void* scratch, cudaStream_t st) {
  PVCR_REQUIRE(K >= 1 && K <= BEAM_MAX_K, "beam_select: beam width %d not in 1..%d", K, BEAM_MAX_K);
  PVCR_REQUIRE((long long)K * Vc < 0x7fffffff, "beam_select: K * Vc too large");
  float* lse = static_cast<float*>(scratch);
  float* part_v = lse + (size_t)B * K;
  int* part_i = reinterpret_cast<int*>(part_v + (size_t)B * BEAM_SLICES * K);
  { LaunchScope ls_(KC_MISC, st);
  beam_lse_kernel<<<B * K, BEAM_THREADS, 0, st>>>(logits, ld, Vc, lse);
  }
  { LaunchScope ls_(KC_MISC, st);
  beam_partial_kernel<<<dim3(BEAM_SLICES, B), BEAM_THREADS, 0, st>>>(logits, ld, Vc, K, first, score_in, lse, part_v, part_i);
  }
  { LaunchScope ls_(KC_MISC, st);
  beam_merge_kernel<<<B, 64, 0, st>>>(part_v, part_i, BEAM_SLICES, Vc, K, score_out, parent, word);
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
