// Optimizer step of the reference training loop (train.py:104-105,157-160, train_rationale.py likewise):
//     nn.utils.clip_grad_norm_(model.parameters(), max_norm);  torch.optim.Adam(lr, weight_decay).step()
// as two multi-tensor kernels over a chunk table (no host synchronisation, CUDA-graph capturable):
//   1. per-chunk sum of squares of every gradient (one deterministic partial per chunk);
//   2. every CTA re-reduces the partials in a fixed order (total norm, clip coefficient
//      min(1, max_norm / (norm + 1e-6)) exactly as torch computes it), then applies Adam with the L2-style weight
//      decay torch.optim.Adam uses (grad += wd * param, which is why the embedding gradient is dense):
//          m = b1 m + (1-b1) g ;  v = b2 v + (1-b2) g^2 ;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// The step count lives on the device (incremented by kernel 1) so that graph replays advance the bias correction.
// HBM-bound: reads g twice, p/m/v once, writes p/m/v once = 32 bytes per parameter.
#include "../../include/pvcr_b200.h"
#include "common.cuh"
#include "kernels.cuh"

namespace pvcr {

constexpr int OPT_THREADS = 256;

struct AdamTensor { float* p; float* g; float* m; float* v; long long n; };
static_assert(sizeof(AdamTensor) == sizeof(PvcrAdamTensor), "PvcrAdamTensor layout");

__device__ __forceinline__ float block_sum(float s, float* red) {
  s = warp_sum(s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = lane < OPT_THREADS / 32 ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  t = red[0];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(OPT_THREADS) grad_sumsq_kernel(const AdamTensor* __restrict__ tensors,
                                                                 const int* __restrict__ chunk_tensor,
                                                                 const long long* __restrict__ chunk_off, int chunk_elems,
                                                                 float* __restrict__ partial, long long* step) {
  __shared__ float red[OPT_THREADS / 32];
  const int c = blockIdx.x;
  const AdamTensor t = tensors[chunk_tensor[c]];
  const long long off = chunk_off[c];
  const long long n = min((long long)chunk_elems, t.n - off);
  const float* g = t.g + off;
  float s = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const long long n4 = n >> 2;
    for (long long i = threadIdx.x; i < n4; i += OPT_THREADS) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(g) + i);
      s += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += OPT_THREADS) s += g[i] * g[i];
  } else {
    for (long long i = threadIdx.x; i < n; i += OPT_THREADS) s += g[i] * g[i];
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    partial[c] = s;
    if (c == 0 && step) *step += 1;
  }
}

__global__ void __launch_bounds__(OPT_THREADS) adam_clip_kernel(const AdamTensor* __restrict__ tensors,
                                                                const int* __restrict__ chunk_tensor,
                                                                const long long* __restrict__ chunk_off, int chunk_elems,
                                                                int n_chunks, const float* __restrict__ partial,
                                                                const long long* step, long long step_host, float lr,
                                                                float beta1, float beta2, float eps, float wd,
                                                                float max_norm, float* norm_out) {
  __shared__ float red[OPT_THREADS / 32];
  // total norm: same fixed-order reduction in every CTA (deterministic and identical everywhere)
  float s = 0.f;
  for (int i = threadIdx.x; i < n_chunks; i += OPT_THREADS) s += partial[i];
  const float total = sqrtf(block_sum(s, red));
  float coef = 1.f;
  if (max_norm > 0.f) coef = fminf(max_norm / (total + 1e-6f), 1.f);
  const int c = blockIdx.x;
  if (c == 0 && threadIdx.x == 0 && norm_out) norm_out[0] = total;
  const long long tstep = step ? *step : step_host;
  // bias corrections in double as torch computes them on the host (python floats)
  const double bc1 = 1.0 - pow((double)beta1, (double)tstep);
  const double bc2 = 1.0 - pow((double)beta2, (double)tstep);
  const float step_size = (float)((double)lr / bc1);
  const float sqrt_bc2 = (float)sqrt(bc2);
  const AdamTensor t = tensors[chunk_tensor[c]];
  const long long off = chunk_off[c];
  const long long n = min((long long)chunk_elems, t.n - off);
  float* p = t.p + off; float* g = t.g + off; float* m = t.m + off; float* v = t.v + off;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg = gg * coef + wd * pp;
    mm = mm + (1.f - beta1) * (gg - mm);                 // exp_avg.lerp_(grad, 1 - beta1)
    vv = beta2 * vv + (1.f - beta2) * gg * gg;
    const float denom = sqrtf(vv) / sqrt_bc2 + eps;
    pp -= step_size * (mm / denom);
  };
  const bool al = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                    reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (al) {
    const long long n4 = n >> 2;
    for (long long i = threadIdx.x; i < n4; i += OPT_THREADS) {
      float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
      const float4 G = reinterpret_cast<const float4*>(g)[i];
      upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
      reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(v)[i] = V;
    }
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += OPT_THREADS) upd(p[i], g[i], m[i], v[i]);
  } else {
    for (long long i = threadIdx.x; i < n; i += OPT_THREADS) upd(p[i], g[i], m[i], v[i]);
  }
}

}  // namespace pvcr

using namespace pvcr;

extern "C" {

int pvcr_adam_clip_step(const PvcrAdamTensor* tensors_dev, const int32_t* chunk_tensor_dev, const int64_t* chunk_off_dev,
                        int n_chunks, int chunk_elems, float lr, float beta1, float beta2, float eps, float weight_decay,
                        float max_norm, int64_t* step_dev, int64_t step_host, float* partial_dev, float* norm_out_dev,
                        void* stream) {
  PVCR_REQUIRE(tensors_dev && chunk_tensor_dev && chunk_off_dev && partial_dev, "pvcr_adam_clip_step: null table");
  PVCR_REQUIRE(n_chunks > 0 && chunk_elems > 0, "pvcr_adam_clip_step: n_chunks=%d chunk_elems=%d", n_chunks, chunk_elems);
  PVCR_REQUIRE(step_dev || step_host > 0, "pvcr_adam_clip_step: step must be >= 1 (or a device counter)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const AdamTensor* t = reinterpret_cast<const AdamTensor*>(tensors_dev);
  {
    LaunchScope ls_(KC_MISC, st);
    grad_sumsq_kernel<<<n_chunks, OPT_THREADS, 0, st>>>(t, chunk_tensor_dev, reinterpret_cast<const long long*>(chunk_off_dev),
                                                        chunk_elems, partial_dev, reinterpret_cast<long long*>(step_dev));
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  {
    LaunchScope ls_(KC_MISC, st);
    adam_clip_kernel<<<n_chunks, OPT_THREADS, 0, st>>>(t, chunk_tensor_dev, reinterpret_cast<const long long*>(chunk_off_dev),
                                                       chunk_elems, n_chunks, partial_dev,
                                                       reinterpret_cast<const long long*>(step_dev), (long long)step_host, lr,
                                                       beta1, beta2, eps, weight_decay, max_norm, norm_out_dev);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

}  // extern "C"
