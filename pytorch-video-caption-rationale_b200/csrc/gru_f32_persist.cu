// Persistent GRU forward sweep in plain fp32 arithmetic (CUDA cores), for decoding.
//
// Caption decoding runs in fp32-equivalent arithmetic so that the arg-max word ids match the reference
// (model/S2VTAttModel.py:80-96 encoder GRU, eval branch).  On the tensor cores that costs three bf16 planes per operand
// and six products: the W_hh planes no longer fit beside the activations in shared memory, so the encoder of a decode
// call used to take one split-K GEMM + one gate kernel per frame (25 us per step, 40 steps).  The recurrent product is
// tiny (B x 3H x H MACs per step = 2.8 us of FFMA issue on the whole GPU at B = 128), so this kernel does it on the
// CUDA cores straight from the fp32 parameter: ONE cooperative launch for all T steps,
//   * W_hh rows of 16 hidden units (48 rows x H fp32, padded pitch) resident in shared memory per CTA,
//   * the batch cut into groups of 32 videos served by H/16 CTAs; per step the group's h_{t-1} (32 x H fp32) comes back
//     from L2 into shared memory (cp.async) behind the group's arrive counter (persist.cuh),
//   * thread = (K half, one hidden unit, four videos): all three gates of its unit, so the GRU cell is applied in
//     registers after one shared-memory hand-over between the two K halves,
//   * h_t written as fp32 rows (the encoder output); the caller casts them to the bf16 split planes the following GEMMs
//     consume in one pass after the sweep.
// No rounding of operands anywhere: results differ from torch's fp32 GRU by summation order only.
#include <cstdlib>

#include "../../include/pvcr_b200.h"
#include "host.h"
#include "persist.cuh"

namespace pvcr {

constexpr int GF_U = 16;          // hidden units per CTA
constexpr int GF_BG = 32;         // videos per group
constexpr int GF_THREADS = 256;   // 2 K-halves x 16 units x 8 video quads
constexpr int GF_OP = GF_U + 4;   // pitch of the output staging rows

// Thread (ks, jj, vq): K-half ks, hidden unit jj, videos vq, vq+8, vq+16, vq+24 of the group -- a 3 gates x 4 videos
// register tile, 7 shared-memory loads (16 B) per 48 FMAs.  Row pitch H + 4 floats for both W and h: the 8 lanes of a
// quarter warp (consecutive videos / 4 consecutive units) fall on disjoint banks.
__global__ void __launch_bounds__(GF_THREADS, 1) gru_f32_persist_fwd_kernel(const GruF32Fwd p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int H = p.H, T = p.T, B = p.B, C = p.C, K4 = H >> 2, ld_s = H + 4;
  float* sW = reinterpret_cast<float*>(smem_raw);                 // [3 * GF_U][H + 4]
  float* sH = sW + (size_t)3 * GF_U * ld_s;                       // [GF_BG][H + 4]: h_{t-1} of the group
  float* sP = sH + (size_t)GF_BG * ld_s;                          // [128][12]: partial sums of the upper K half
  float* sO = sP + 128 * 12;                                      // [GF_BG][GF_OP]: h_t of this CTA's units
  const int tid = threadIdx.x, ks = tid >> 7, jj = (tid & 127) >> 3, vq = tid & 7;
  const int grp = blockIdx.x / C, cta = blockIdx.x - grp * C;
  const int j0 = cta * GF_U, j = j0 + jj, b0 = grp * GF_BG;
  unsigned* ctr = p.counters + grp * 32;
  unsigned target = 0;

  // resident weight rows: local row g * 16 + jj = W_hh[g * H + j0 + jj, :]
  for (int i = tid; i < 3 * GF_U * K4; i += GF_THREADS) {
    const int r = i / K4, k4 = i - r * K4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(p.w_hh + ((long long)(r / GF_U) * H + j0 + (r % GF_U)) * H) + k4);
    *reinterpret_cast<float4*>(sW + (size_t)r * ld_s + 4 * k4) = v;
  }
  const float bhr = p.b_hh[j], bhz = p.b_hh[H + j], bhn = p.b_hh[2 * H + j];
  const uint32_t sH_u32 = smem_u32(sH);
  __syncthreads();

  for (int t = 0; t < T; ++t) {
    // this step's input projections (independent of the recurrence: in flight across the barrier)
    float gi[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int b = b0 + vq + 8 * i;
      gi[i][0] = gi[i][1] = gi[i][2] = 0.f;
      if (ks == 0 && b < B) {
        const float* s = p.gi + (long long)t * p.gi_ts + (long long)b * p.gi_ld + j;
        gi[i][0] = __ldg(s); gi[i][1] = __ldg(s + H); gi[i][2] = __ldg(s + 2 * H);
      }
    }
    phase_stamp(p.dbg, t, 0);
    const bool has_prev = t > 0 || p.h0 != nullptr;
    float acc[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.f;
    float hprev[4] = {0.f, 0.f, 0.f, 0.f};
    if (has_prev) {
      const float* src = t > 0 ? p.h + (long long)(t - 1) * p.h_ts : p.h0;
      const long long ld = t > 0 ? p.h_ld : p.h0_ld;
      if (t > 0) { target += (unsigned)C; group_wait(ctr, target); }
      phase_stamp(p.dbg, t, 1);
      // h_{t-1} of the group's 32 videos -> shared memory with cp.async: lanes along a video's row (coalesced), every
      // 16-byte piece in flight at once; rows past the batch are zero-filled
      for (int i = tid; i < GF_BG * K4; i += GF_THREADS) {
        const int v = i / K4, k4 = i - v * K4;
        const bool ok = b0 + v < B;
        const float* g = src + (ok ? (long long)(b0 + v) * ld + 4 * k4 : 0);
        const uint32_t d = sH_u32 + (uint32_t)(v * ld_s + 4 * k4) * 4u;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(g), "r"(ok ? 16 : 0) : "memory");
      }
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      phase_stamp(p.dbg, t, 2);
      const float* wr = sW + (size_t)jj * ld_s + (size_t)ks * (H >> 1);
      const float* wz = wr + (size_t)GF_U * ld_s;
      const float* wn = wz + (size_t)GF_U * ld_s;
      const float* hv = sH + (size_t)vq * ld_s + (size_t)ks * (H >> 1);
      const int K4h = K4 >> 1;
#pragma unroll 4
      for (int k4 = 0; k4 < K4h; ++k4) {
        const float4 a = *reinterpret_cast<const float4*>(wr + 4 * k4);
        const float4 b = *reinterpret_cast<const float4*>(wz + 4 * k4);
        const float4 c = *reinterpret_cast<const float4*>(wn + 4 * k4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 h = *reinterpret_cast<const float4*>(hv + (size_t)(8 * i) * ld_s + 4 * k4);
          acc[i][0] = fmaf(a.w, h.w, fmaf(a.z, h.z, fmaf(a.y, h.y, fmaf(a.x, h.x, acc[i][0]))));
          acc[i][1] = fmaf(b.w, h.w, fmaf(b.z, h.z, fmaf(b.y, h.y, fmaf(b.x, h.x, acc[i][1]))));
          acc[i][2] = fmaf(c.w, h.w, fmaf(c.z, h.z, fmaf(c.y, h.y, fmaf(c.x, h.x, acc[i][2]))));
        }
      }
      if (ks == 1) {
        float* o = sP + (size_t)(tid & 127) * 12;
#pragma unroll
        for (int i = 0; i < 4; ++i) { o[3 * i] = acc[i][0]; o[3 * i + 1] = acc[i][1]; o[3 * i + 2] = acc[i][2]; }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) hprev[i] = sH[(size_t)(vq + 8 * i) * ld_s + j];
      }
      __syncthreads();
    }
    phase_stamp(p.dbg, t, 3);
    // GRU cell (torch gate order r, z, n; same expressions as gru_gate_fwd_kernel) on the lower-K-half threads
    if (ks == 0) {
      const float* o = sP + (size_t)tid * 12;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float gr = acc[i][0], gz = acc[i][1], gn = acc[i][2];
        if (has_prev) { gr += o[3 * i]; gz += o[3 * i + 1]; gn += o[3 * i + 2]; }
        const float r = 1.f / (1.f + expf(-(gi[i][0] + gr + bhr)));
        const float z = 1.f / (1.f + expf(-(gi[i][1] + gz + bhz)));
        const float n = tanhf(gi[i][2] + r * (gn + bhn));
        sO[(vq + 8 * i) * GF_OP + jj] = (1.f - z) * n + z * hprev[i];
      }
    }
    __syncthreads();
    // h_t of this CTA's 16 units: 64 contiguous bytes per video, one float4 per thread
    if (tid < GF_BG * 4) {
      const int v = tid >> 2, seg = tid & 3;
      if (b0 + v < B)
        *reinterpret_cast<float4*>(p.h + (long long)t * p.h_ts + (long long)(b0 + v) * p.h_ld + j0 + 4 * seg) =
            *reinterpret_cast<const float4*>(sO + v * GF_OP + 4 * seg);
    }
    phase_stamp(p.dbg, t, 4);
    if (t + 1 < T) group_arrive(ctr);       // publishes h_t; its CTA barrier also frees sH / sP / sO for the next step
  }
}

static bool plan_f32(int B, int H, int& C, int& G, size_t& smem) {
  if (H % GF_U != 0 || H % 8 != 0 || H > 512) return false;
  C = H / GF_U;
  G = (B + GF_BG - 1) / GF_BG;
  smem = ((size_t)(3 * GF_U + GF_BG) * (H + 4) + 128 * 12 + GF_BG * GF_OP) * sizeof(float);
  return C <= 32 && (long long)G * C <= sm_count() && smem <= 227 * 1024;
}

bool gru_f32_persist_eligible(int B, int H) {
  static const bool off = getenv("PVCR_NO_F32_GRU") != nullptr;        // A/B knob
  int C, G;
  size_t smem;
  return !off && plan_f32(B, H, C, G, smem);
}

int gru_f32_persist_fwd(const GruF32Fwd& p0, cudaStream_t st) {
  int C, G;
  size_t smem;
  PVCR_REQUIRE(plan_f32(p0.B, p0.H, C, G, smem), "gru_f32_persist_fwd: shape B=%d H=%d not supported", p0.B, p0.H);
  PVCR_REQUIRE((reinterpret_cast<uintptr_t>(p0.w_hh) & 15) == 0 && (reinterpret_cast<uintptr_t>(p0.h) & 15) == 0 &&
                   p0.h_ld % 4 == 0 && p0.h_ts % 4 == 0 && (!p0.h0 || ((reinterpret_cast<uintptr_t>(p0.h0) & 15) == 0 && p0.h0_ld % 4 == 0)),
               "gru_f32_persist_fwd: weight / state rows must be 16-byte aligned");
  GruF32Fwd p = p0;
  p.C = C;
  p.dbg = getenv("PVCR_PHASE_GRU_F32") ? debug_phase_buffer() : nullptr;
  const void* kern = (const void*)gru_f32_persist_fwd_kernel;
  PVCR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  PVCR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GF_THREADS, smem));
  PVCR_REQUIRE(per_sm * sm_count() >= G * C, "gru_f32_persist_fwd: %d CTAs cannot be co-resident", G * C);
  PVCR_TRY(fill_zero(p.counters, sizeof(unsigned) * 32 * G, st));
  void* args[] = {&p};
  LaunchScope ls_(KC_GRU_FWD, st);
  PVCR_CUDA_CHECK(cudaLaunchCooperativeKernel(kern, dim3(G * C), dim3(GF_THREADS), args, smem, st));
  return PVCR_OK;
}

}  // namespace pvcr
