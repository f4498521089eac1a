// Row-streaming fp32 product for decoding a handful of videos:  out[b, j] = sum_k x[b, k] W[j, k] (+ bias[j]),  b < B <= 4.
//
// At small batches a decoding step is a matrix-VECTOR product per weight: the tensor-core path (three bf16 terms per
// operand, 128-row tiles of which B are real) still streams 3 x 2 bytes of W_v per weight and step; this kernel reads the
// fp32 parameter itself once (4 bytes per weight, the floor for fp32-equivalent arithmetic) on the CUDA cores:
//   * one warp per weight row at a time, 16-byte loads along K, RU rows' loads in flight per lane,
//   * the B input rows resident in shared memory, B accumulators per lane, butterfly reduction per (row, video),
//   * either plain stores (the [q | gh] product of a step: two stacked weight matrices) or a running (max, first index) per
//     video and CTA -- the per-CTA partials are combined by the gate kernel of the next step like the GEMM epilogue's.
// No operand is rounded: results differ from torch's fp32 Linear by summation order only
// (model/S2VTAttModel.py:140-147 projection + arg-max of the eval branch, :125-137 query / hidden products).
#include <cstdlib>

#include "../../include/pvcr_b200.h"
#include "host.h"

namespace pvcr {

constexpr int GV_THREADS = 256;
constexpr int GV_MAXB = 4;           // measured: 1.00 / 1.05 ms per batch at B = 1 / 2 (tensor path 1.59); at B = 8 the tensor path wins (1.50 vs 1.55)
constexpr int GV_RU = 4;           // rows per warp pass (their loads are issued together)

template <int NB, bool ARGMAX>
__global__ void __launch_bounds__(GV_THREADS) gemv_f32_kernel(const GemvF32 p) {
  extern __shared__ __align__(16) float gv_sx[];                 // [NB][K]
  __shared__ float s_m[GV_THREADS / 32][GV_MAXB];
  __shared__ int s_i[GV_THREADS / 32][GV_MAXB];
  const int K = p.K, K4 = K >> 2, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < NB * K4; i += GV_THREADS) {
    const int b = i / K4, k4 = i - b * K4;
    reinterpret_cast<float4*>(gv_sx)[i] = b < p.B ? __ldg(reinterpret_cast<const float4*>(p.x + (long long)b * p.x_ld) + k4)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int R = p.rows0 + p.rows1;
  const int wg = blockIdx.x * (GV_THREADS / 32) + warp, nwg = gridDim.x * (GV_THREADS / 32);
  float bm[NB];
  int bi[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) { bm[b] = -INFINITY; bi[b] = 0x7fffffff; }
  // rows are dealt to warps in ascending order per warp, so "first maximum wins" holds inside a warp; across warps / CTAs
  // the combine step breaks ties towards the lower index
  for (int j0 = wg * GV_RU; j0 < R; j0 += nwg * GV_RU) {
    float acc[GV_RU][NB];
#pragma unroll
    for (int r = 0; r < GV_RU; ++r)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[r][b] = 0.f;
    for (int k4 = lane; k4 < K4; k4 += 32) {
      float4 w[GV_RU];
#pragma unroll
      for (int r = 0; r < GV_RU; ++r) {
        const int j = j0 + r;
        const float* row = j < p.rows0 ? p.w0 + (long long)j * p.w0_ld : p.w1 + (long long)(j - p.rows0) * p.w1_ld;
        w[r] = j < R ? (p.stream ? __ldcs(reinterpret_cast<const float4*>(row) + k4) : __ldg(reinterpret_cast<const float4*>(row) + k4))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float4 x = reinterpret_cast<const float4*>(gv_sx)[b * K4 + k4];
#pragma unroll
        for (int r = 0; r < GV_RU; ++r) acc[r][b] += w[r].x * x.x + w[r].y * x.y + w[r].z * x.z + w[r].w * x.w;
      }
    }
#pragma unroll
    for (int r = 0; r < GV_RU; ++r) {
      const int j = j0 + r;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float v = warp_sum(acc[r][b]);
        if (j < R && b < p.B) {
          if (p.bias) v += __ldg(p.bias + j);
          if (ARGMAX) {
            if (v > bm[b]) { bm[b] = v; bi[b] = j; }
            if (p.out && lane == 0) __stcs(p.out + (long long)b * p.out_ld + j, v);
          } else if (lane == 0) {
            p.out[(long long)b * p.out_ld + j] = v;
          }
        }
      }
    }
  }
  if (ARGMAX) {
    // (every lane holds the same bm / bi after warp_sum)  combine the 8 warps of the CTA: larger value, then lower index
    if (lane == 0) {
#pragma unroll
      for (int b = 0; b < NB; ++b) { s_m[warp][b] = bm[b]; s_i[warp][b] = bi[b]; }
    }
    __syncthreads();
    if (tid < p.B) {
      float m = s_m[0][tid];
      int mi = s_i[0][tid];
      for (int w = 1; w < GV_THREADS / 32; ++w) {
        const float v = s_m[w][tid];
        const int i = s_i[w][tid];
        if (v > m || (v == m && i < mi)) { m = v; mi = i; }
      }
      p.pmax[(long long)tid * gridDim.x + blockIdx.x] = m;
      p.pidx[(long long)tid * gridDim.x + blockIdx.x] = mi;
    }
  }
}

bool gemv_f32_eligible(int B, int K) { return B >= 1 && B <= GV_MAXB && K % 4 == 0 && K <= 4096; }

// number of (max, index) partials per video the arg-max mode writes (= CTAs launched); scratch: 2 * B * parts * 4 bytes
int gemv_f32_parts(int rows) {
  const int want = cdiv(rows, (GV_THREADS / 32) * GV_RU);
  const int cap = 2 * sm_count();
  return want < cap ? want : cap;
}

template <bool ARGMAX>
static int launch_gemv(const GemvF32& p, int grid, cudaStream_t st) {
  const size_t smem_for = sizeof(float) * (size_t)p.K;
#define PVCR_GV(NB_)                                                                                               \
  {                                                                                                                \
    auto kern = gemv_f32_kernel<NB_, ARGMAX>;                                                                      \
    static bool attr = false;                                                                                      \
    if (!attr) {                                                                                                   \
      PVCR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, NB_ * 4096 * 4));    \
      attr = true;                                                                                                 \
    }                                                                                                              \
    LaunchScope ls_(KC_MISC, st);                                                                                  \
    kern<<<grid, GV_THREADS, NB_ * smem_for, st>>>(p);                                                             \
  }
  if (p.B <= 1) PVCR_GV(1) else if (p.B <= 2) PVCR_GV(2) else PVCR_GV(4)
#undef PVCR_GV
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

int gemv_f32(const GemvF32& p, cudaStream_t st) {
  PVCR_REQUIRE(gemv_f32_eligible(p.B, p.K), "gemv_f32: B=%d K=%d not served (B <= 4, K %% 4 == 0, K <= 4096)", p.B, p.K);
  PVCR_REQUIRE(p.w0 && p.rows0 > 0 && p.w0_ld % 4 == 0 && p.x_ld % 4 == 0 && (p.rows1 == 0 || (p.w1 && p.w1_ld % 4 == 0)) &&
                   ((reinterpret_cast<uintptr_t>(p.w0) | reinterpret_cast<uintptr_t>(p.w1) | reinterpret_cast<uintptr_t>(p.x)) & 15) == 0,
               "gemv_f32: operands must be 16-byte aligned with row strides that are multiples of 4");
  const int rows = p.rows0 + p.rows1;
  if (p.pmax) {
    PVCR_REQUIRE(p.pidx, "gemv_f32: arg-max mode needs both partial buffers");
    return launch_gemv<true>(p, gemv_f32_parts(rows), st);
  }
  PVCR_REQUIRE(p.out, "gemv_f32: no output");
  return launch_gemv<false>(p, gemv_f32_parts(rows), st);
}

}  // namespace pvcr
