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
//   * the product as 6 x 8 register tiles (3 gates of 2 units x 8 videos per thread, the K range dealt over the 8 warps),
//     the 8 partial sums combined through shared memory, then one thread per (unit, video) pair applies the GRU cell,
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
constexpr int GF_THREADS = 256;   // 8 warps = 8 K slices; lane = (unit pair, video quad)
constexpr int GF_OP = GF_U + 4;   // pitch of the output staging rows

// Matvec mapping: warp w owns the K range [w, w+1) * H/8; lane (jp = lane / 4, vq = lane % 4) accumulates the 3 gates of
// units 2jp, 2jp+1 for the 8 videos vq, vq+4, ..., vq+28 over that range: a 6 x 8 register tile, 14 shared-memory loads
// (16 B) per 192 FMAs -- a 128-bit ld.shared occupies the SM's load path ~3.5 cycles per warp instruction whether or not
// its lanes broadcast, so loads per FMA is what bounds this loop (the first version's 3 x 4 tile: 7 loads per 48 FMAs,
// 6.4 us per step).  The 8 partial sums of every (gate, unit, video) go through shared memory; thread q then owns the
// (unit, video) pairs q and q + 256 for the GRU cell.  Row pitch H + 4 floats for both W and h keeps the quarter-warp
// phases of the 16-byte loads on disjoint banks.
__global__ void __launch_bounds__(GF_THREADS, 1) gru_f32_persist_fwd_kernel(const GruF32Fwd p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int H = p.H, T = p.T, B = p.B, C = p.C, K4 = H >> 2, ld_s = H + 4;
  float* sW = reinterpret_cast<float*>(smem_raw);                 // [3 * GF_U][H + 4]
  float* sH = sW + (size_t)3 * GF_U * ld_s;                       // [GF_BG][H + 4]: h_{t-1} of the group
  float* sP = sH + (size_t)GF_BG * ld_s;                          // [8 warps][3 gates][GF_BG][GF_U] partial sums
  float* sO = sP + 8 * 3 * GF_BG * GF_U;                          // [GF_BG][GF_OP]: h_t of this CTA's units
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, jp = lane >> 2, vq = lane & 3;
  const int grp = blockIdx.x / C, cta = blockIdx.x - grp * C;
  const int j0 = cta * GF_U, b0 = grp * GF_BG;
  unsigned* ctr = p.counters + grp * 32;
  unsigned target = 0;

  // resident weight rows: local row g * 16 + jj = W_hh[g * H + j0 + jj, :]
  for (int i = tid; i < 3 * GF_U * K4; i += GF_THREADS) {
    const int r = i / K4, k4 = i - r * K4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(p.w_hh + ((long long)(r / GF_U) * H + j0 + (r % GF_U)) * H) + k4);
    *reinterpret_cast<float4*>(sW + (size_t)r * ld_s + 4 * k4) = v;
  }
  // cell role: pairs (video cv0, unit cj) and (cv0 + 16, cj)
  const int cj = tid & 15, cv0 = tid >> 4, j = j0 + cj;
  const float bhr = p.b_hh[j], bhz = p.b_hh[H + j], bhn = p.b_hh[2 * H + j];
  const uint32_t sH_u32 = smem_u32(sH);
  __syncthreads();

  for (int t = 0; t < T; ++t) {
    // this step's input projections (independent of the recurrence: in flight across the barrier)
    float gi[2][3];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int b = b0 + cv0 + 16 * i;
      gi[i][0] = gi[i][1] = gi[i][2] = 0.f;
      if (b < B) {
        const float* s = p.gi + (long long)t * p.gi_ts + (long long)b * p.gi_ld + j;
        gi[i][0] = __ldg(s); gi[i][1] = __ldg(s + H); gi[i][2] = __ldg(s + 2 * H);
      }
    }
    phase_stamp(p.dbg, t, 0);
    const bool has_prev = t > 0 || p.h0 != nullptr;
    float gh[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    float hprev[2] = {0.f, 0.f};
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
      {
        const int K4w = K4 >> 3;                                  // 16-byte columns of this warp's K slice
        const float* w0 = sW + (size_t)(2 * jp) * ld_s + (size_t)warp * K4w * 4;
        const float* hv = sH + (size_t)vq * ld_s + (size_t)warp * K4w * 4;
        float acc[8][6];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int r = 0; r < 6; ++r) acc[i][r] = 0.f;
#pragma unroll 2
        for (int k4 = 0; k4 < K4w; ++k4) {
          float4 w[6];
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            w[2 * g] = *reinterpret_cast<const float4*>(w0 + (size_t)(g * GF_U) * ld_s + 4 * k4);
            w[2 * g + 1] = *reinterpret_cast<const float4*>(w0 + (size_t)(g * GF_U + 1) * ld_s + 4 * k4);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 h = *reinterpret_cast<const float4*>(hv + (size_t)(4 * i) * ld_s + 4 * k4);
#pragma unroll
            for (int r = 0; r < 6; ++r)
              acc[i][r] = fmaf(w[r].w, h.w, fmaf(w[r].z, h.z, fmaf(w[r].y, h.y, fmaf(w[r].x, h.x, acc[i][r]))));
          }
        }
        // partial sums -> sP[warp][gate][video][unit]
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            float* o = sP + ((size_t)(warp * 3 + g) * GF_BG + (vq + 4 * i)) * GF_U + 2 * jp;
            *reinterpret_cast<float2*>(o) = make_float2(acc[i][2 * g], acc[i][2 * g + 1]);
          }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) hprev[i] = sH[(size_t)(cv0 + 16 * i) * ld_s + j];
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          float s = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) s += sP[((size_t)(w * 3 + g) * GF_BG + cv0 + 16 * i) * GF_U + cj];
          gh[i][g] = s;
        }
    }
    phase_stamp(p.dbg, t, 3);
    // GRU cell (torch gate order r, z, n; same expressions as gru_gate_fwd_kernel)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float r = 1.f / (1.f + expf(-(gi[i][0] + gh[i][0] + bhr)));
      const float z = 1.f / (1.f + expf(-(gi[i][1] + gh[i][1] + bhz)));
      const float n = tanhf(gi[i][2] + r * (gh[i][2] + bhn));
      sO[(cv0 + 16 * i) * GF_OP + cj] = (1.f - z) * n + z * hprev[i];
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
  if (H % GF_U != 0 || H % 32 != 0 || H > 512) return false;
  C = H / GF_U;
  G = (B + GF_BG - 1) / GF_BG;
  smem = ((size_t)(3 * GF_U + GF_BG) * (H + 4) + 8 * 3 * GF_BG * GF_U + GF_BG * GF_OP) * sizeof(float);
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
