// Persistent attention-decoder kernels of S2VTAtt (teacher forced): one cooperative launch runs all L decoder
// steps.  Reference: Decoder.forward / forward_step and Attention.forward, model/S2VTAttModel.py:25-48,125-196.
//
// Work split: groups of bs = C videos served by C CTAs; CTA c of a group
//   * owns hidden units [c*u, (c+1)*u): rows {W_q; W_hh r,z,n} (4u x H) and rows {W_c r,z,n} (3u x H) of the
//     step's two weight matrices stay resident in shared memory (bf16, 128-byte swizzle) for all steps;
//   * owns video c of the group for the attention: the video's proj_key (fp16) and encoder outputs (bf16),
//     N x H each, stay resident in REGISTERS for all steps, so the attention phase touches no HBM/L2 data
//     besides q (in) and ctx / alpha (out).
// Per step:  P1  [q | gh] = [W_q; W_hh] h_{i-1}           (swapped tcgen05 MMA, D1 in TMEM)   -> q exchanged
//            P2  alpha = softmax_n(v . tanh(q + pk_n)),  ctx = sum_n alpha_n enc_n            -> ctx exchanged
//            P3  gi_c = W_c ctx                           (swapped tcgen05 MMA, D2 in TMEM)
//            P4  GRU gates with gi = gi_c + (E W_e^T + b_ih)[i] (hoisted), h kept in fp32 registers -> h exchanged
// Three group barriers per step on one monotonic counter per group.
#include <cuda_fp16.h>

#include <cstdlib>
#include <mutex>

#include "host.h"
#include "persist.cuh"

namespace pvcr {

constexpr int DEC_THREADS = 256;
constexpr int DEC_ITEMS = 4;      // (unit, video) pairs per thread: u * bs <= DEC_ITEMS * DEC_THREADS


// TMA: the exchanged operands (h_{i-1} and ctx of the group's videos, 32 KB each) are fetched by KB bulk-tensor copies
// issued by ONE thread straight into the swizzled operand buffer (tmH0 / tmHs / tmCtx describe enc-final rows, hs_a as
// (k, video, step) and ctx_x as (k, video, step)) instead of 2048 16-byte cp.async requests spread over the CTA.
template <int NF, bool ACC_TANH = false, bool TMA = true>
__global__ void __launch_bounds__(DEC_THREADS, 1) dec_persist_fwd_kernel(const DecPersistFwd p,
                                                                         const __grid_constant__ CUtensorMap tmH0,
                                                                         const __grid_constant__ CUtensorMap tmHs,
                                                                         const __grid_constant__ CUtensorMap tmCtx) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, u = p.u, C = p.C, bsp = p.bsp, N = p.N, L = p.L, B = p.B, KB = H >> 6;
  const int R1 = 4 * u, R3 = 3 * u;
  uint8_t* sW1 = smem;
  uint8_t* sW3 = sW1 + (size_t)KB * R1 * 128;
  uint8_t* sX = sW3 + (size_t)KB * R3 * 128;
  float* sS1 = reinterpret_cast<float*>(sX + (size_t)KB * bsp * 128);
  const int s1_ld = R1 + 1, s3_ld = R3 + 1;
  float* sS3 = sS1 + (size_t)bsp * s1_ld;
  const int DG = H >> 3, FG = DEC_THREADS / DG, PW = DG >= 32 ? DG / 32 : 1;
  float* sP = sS3 + (size_t)bsp * s3_ld;    // [N][PW]
  float* sScore = sP + (size_t)N * PW;      // [N]
  float* sC = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sScore + N) + 15) & ~uintptr_t(15));   // [FG][H], float4 access
  uint64_t* bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sC + (size_t)FG * H) + 15) & ~uintptr_t(7));
  uint64_t* bar_x = bar + 1;                 // TMA: the exchanged operand has landed in sX
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = blockIdx.x / C, c = blockIdx.x % C;
  const int b0 = g * C, j0 = c * u;
  const int bs = min(C, B - b0);            // valid videos of this group
  unsigned* ctr = p.counters + g * 32;

  load_operand_rows(sW1, R1, 0, p.w1, p.w1_ld, j0, u, (long long)4 * H, H);
  for (int q = 0; q < 3; ++q) {
    load_operand_rows(sW1, R1, (q + 1) * u, p.w1, p.w1_ld, (long long)(q + 1) * H + j0, u, (long long)4 * H, H);
    load_operand_rows(sW3, R3, q * u, p.w3, p.w3_ld, (long long)q * H + j0, u, (long long)3 * H, H);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar_x, 1);
    fence_barrier_init();
  }
  uint32_t phase_x = 0;
  const uint32_t x_bytes = (uint32_t)KB * (uint32_t)bsp * 128u;
  const uint32_t ncols = 2 * bsp <= 32 ? 32u : (2 * bsp <= 64 ? 64u : (2 * bsp <= 128 ? 128u : 256u));
  if (warp == 0) {
    tmem_alloc(tmem_slot, ncols);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d1 = *tmem_slot, tmem_d2 = tmem_d1 + (uint32_t)bsp;
  const uint32_t idesc = umma_idesc_bf16(128, bsp);

  // ---- attention residency: video vb, dims [d0, d0+8), frames fg + FG*m --------------------------------
  const int vb = min(b0 + c, B - 1);
  const bool video_ok = (c < bs);
  const int dg = tid % DG, fg = tid / DG, d0 = dg * 8;
  __half2 pkr[NF][4];
  __nv_bfloat162 enr[NF][4];
  float v8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v8[e] = p.v[d0 + e];
#pragma unroll
  for (int m = 0; m < NF; ++m) {
    const int n = fg + FG * m;
#pragma unroll
    for (int e = 0; e < 4; ++e) { pkr[m][e] = __floats2half2_rn(0.f, 0.f); enr[m][e] = __floats2bfloat162_rn(0.f, 0.f); }
    if (n < N) {
      const float4* s4 = reinterpret_cast<const float4*>(p.pk + ((long long)vb * N + n) * H + d0);
      const float4 a = __ldg(s4), bq = __ldg(s4 + 1);
      const float lim = 60000.f;
      pkr[m][0] = __floats2half2_rn(fminf(fmaxf(a.x, -lim), lim), fminf(fmaxf(a.y, -lim), lim));
      pkr[m][1] = __floats2half2_rn(fminf(fmaxf(a.z, -lim), lim), fminf(fmaxf(a.w, -lim), lim));
      pkr[m][2] = __floats2half2_rn(fminf(fmaxf(bq.x, -lim), lim), fminf(fmaxf(bq.y, -lim), lim));
      pkr[m][3] = __floats2half2_rn(fminf(fmaxf(bq.z, -lim), lim), fminf(fmaxf(bq.w, -lim), lim));
      const uint4 ev = __ldg(reinterpret_cast<const uint4*>(p.enc_a + ((long long)vb * N + n) * p.enc_ld + d0));
      enr[m][0] = *reinterpret_cast<const __nv_bfloat162*>(&ev.x);
      enr[m][1] = *reinterpret_cast<const __nv_bfloat162*>(&ev.y);
      enr[m][2] = *reinterpret_cast<const __nv_bfloat162*>(&ev.z);
      enr[m][3] = *reinterpret_cast<const __nv_bfloat162*>(&ev.w);
    }
  }

  // ---- GRU state of this thread's (unit, video) pairs ---------------------------------------------------
  const int n_items = (u * C + DEC_THREADS - 1) / DEC_THREADS;
  float hreg[DEC_ITEMS], bhr[DEC_ITEMS], bhz[DEC_ITEMS], bhn[DEC_ITEMS];
#pragma unroll
  for (int k = 0; k < DEC_ITEMS; ++k) {
    hreg[k] = bhr[k] = bhz[k] = bhn[k] = 0.f;
    if (k < n_items) {
      const int idx = tid + k * DEC_THREADS, jj = idx % u, lb = idx / u;
      if (lb < bs) {
        const int j = j0 + jj, b = b0 + lb;
        bhr[k] = p.b_hh[j]; bhz[k] = p.b_hh[H + j]; bhn[k] = p.b_hh[2 * H + j];
        hreg[k] = p.h0[(long long)b * p.h0_ld + j];
      }
    }
  }
  uint32_t phase = 0;
  unsigned target = 0;

  for (int i = 0; i < L; ++i) {
    // prefetch the hoisted embedding projection of this step
    float epr[DEC_ITEMS], epz[DEC_ITEMS], epn[DEC_ITEMS];
#pragma unroll
    for (int k = 0; k < DEC_ITEMS; ++k) {
      epr[k] = epz[k] = epn[k] = 0.f;
      if (k < n_items) {
        const int idx = tid + k * DEC_THREADS, jj = idx % u, lb = idx / u;
        if (lb < bs) {
          const float* e = p.ep + ((long long)(b0 + lb) * L + i) * 3 * H + j0 + jj;
          epr[k] = __ldg(e); epz[k] = __ldg(e + H); epn[k] = __ldg(e + 2 * H);
        }
      }
    }
    // ---- P1: [q | gh] = [W_q; W_hh] h_{i-1} -------------------------------------------------------------
    phase_stamp(p.dbg, i, 0);
    if (TMA) {
      if (tid == 0) {
        if (i > 0 && ld_acquire_u32(ctr) < target) {
          const long long t0 = clock64();
          while (ld_acquire_u32(ctr) < target) {
            if (clock64() - t0 > 4000000000LL) __trap();
          }
        }
        phase_stamp(p.dbg, i, 1);
        fence_proxy_async();                  // the other CTAs' generic-proxy stores (acquired above) before async-proxy reads
        mbar_arrive_expect_tx(bar_x, x_bytes);
        for (int kb = 0; kb < KB; ++kb)
          tma_load_3d(sX + (size_t)kb * bsp * 128, i > 0 ? &tmHs : &tmH0, bar_x, kb * 64, b0, i > 0 ? i - 1 : 0);
        mbar_wait(bar_x, phase_x);
        phase_stamp(p.dbg, i, 2);
        tc_fence_after();
        issue_swapped_mma(tmem_d1, smem_u32(sW1), R1, smem_u32(sX), bsp, H, idesc, bar);
      }
      phase_x ^= 1;
    } else {
    if (i > 0) {
      group_wait(ctr, target);
      phase_stamp(p.dbg, i, 1);
      load_operand_rows_async(sX, bsp, 0, p.hs_a + (long long)(i - 1) * p.hs_a_ld, (long long)L * p.hs_a_ld, b0, bsp,
                              b0 + bs, H);
    } else {
      load_operand_rows_async(sX, bsp, 0, p.h0_a, p.h0_a_ld, b0, bsp, b0 + bs, H);
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async();
    __syncthreads();
    phase_stamp(p.dbg, i, 2);
    if (tid == 0) {
      tc_fence_after();
      issue_swapped_mma(tmem_d1, smem_u32(sW1), R1, smem_u32(sX), bsp, H, idesc, bar);
    }
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    phase_stamp(p.dbg, i, 3);
    if (tid < 128) tmem_to_smem_cols(tmem_d1, sS1, s1_ld, R1, bsp);
    tc_fence_before();
    __syncthreads();
    {
      float* qout = p.q_all + (long long)i * B * p.q_ld;
      for (int idx = tid; idx < u * bs; idx += DEC_THREADS) {
        const int jj = idx % u, lb = idx / u;
        qout[(long long)(b0 + lb) * p.q_ld + j0 + jj] = sS1[lb * s1_ld + jj];
      }
    }
    if (p.xq == 0) { group_arrive(ctr); target += (unsigned)C; }
    phase_stamp(p.dbg, i, 4);
    // ---- P2: attention of video vb ---------------------------------------------------------------------
    // q of this video arrives in 16-float pieces from the group's CTAs: behind the group barrier (xq 0), or polled on the
    // data itself (persist.cuh) by every thread for its own 8 values (xq 1) or by warp 0 for the CTA (xq 2)
    {
      const float* qrow = p.q_all + (long long)i * B * p.q_ld + (long long)vb * p.q_ld;
      float4 qa, qb;
      if (p.xq == 0) {
        group_wait(ctr, target);
        phase_stamp(p.dbg, i, 5);
        qa = __ldcg(reinterpret_cast<const float4*>(qrow + d0)); qb = __ldcg(reinterpret_cast<const float4*>(qrow + d0) + 1);
      } else if (p.xq == 1) {
        phase_stamp(p.dbg, i, 5);
        poll_f8(qrow + d0, qa, qb);
      } else {
        if (warp == 0) {
          for (int c = tid * 8; c < H; c += 256) {
            float4 a, b;
            poll_f8(qrow + c, a, b);
            *reinterpret_cast<float4*>(sC + c) = a; *reinterpret_cast<float4*>(sC + c + 4) = b;
          }
        }
        __syncthreads();
        phase_stamp(p.dbg, i, 5);
        qa = *reinterpret_cast<const float4*>(sC + d0); qb = *reinterpret_cast<const float4*>(sC + d0 + 4);
      }
      const float q8[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
      if (p.dbg && tid == 0 && blockIdx.x == 0 && q8[0] == 123.456f) p.dbg[0] = 0;   // force the load to complete
      phase_stamp(p.dbg, i, 8);
      const int W = DG < 32 ? DG : 32;
      float sc[NF];
#pragma unroll
      for (int m = 0; m < NF; ++m) {
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 pf = __half22float2(pkr[m][e]);
          if (ACC_TANH) s += v8[2 * e] * fast_tanh(q8[2 * e] + pf.x) + v8[2 * e + 1] * fast_tanh(q8[2 * e + 1] + pf.y);
          else s += v8[2 * e] * tanh_approx(q8[2 * e] + pf.x) + v8[2 * e + 1] * tanh_approx(q8[2 * e + 1] + pf.y);
        }
        sc[m] = s;
      }
      // reduce every frame's partial score over the dims held by the other lanes (levels interleaved over frames)
      if (W == 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int m = 0; m < NF; ++m) sc[m] += __shfl_xor_sync(0xffffffffu, sc[m], o);
        }
      } else {
        for (int o = W >> 1; o > 0; o >>= 1) {
#pragma unroll
          for (int m = 0; m < NF; ++m) sc[m] += __shfl_xor_sync(0xffffffffu, sc[m], o);
        }
      }
#pragma unroll
      for (int m = 0; m < NF; ++m) {
        const int n = fg + FG * m;
        if ((tid % W) == 0 && n < N) sP[n * PW + dg / 32] = sc[m];
      }
      phase_stamp(p.dbg, i, 9);
      __syncthreads();
      // softmax over the N frames by warp 0 (lanes over frames), alpha broadcast through shared memory
      if (warp == 0) {
        float mx = -INFINITY;
        for (int n = tid; n < N; n += 32) {
          float s = 0.f;
          for (int w = 0; w < PW; ++w) s += sP[n * PW + w];
          sScore[n] = s;
          mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float den = 0.f;
        for (int n = tid; n < N; n += 32) {
          const float e = __expf(sScore[n] - mx);
          sScore[n] = e;
          den += e;
        }
        den = warp_sum(den);
        const float inv = 1.f / den;
        for (int n = tid; n < N; n += 32) {
          const float al = sScore[n] * inv;
          sScore[n] = al;
          if (video_ok) p.alpha[((long long)i * B + vb) * N + n] = al;
        }
      }
      __syncthreads();
      phase_stamp(p.dbg, i, 10);
      float c8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int m = 0; m < NF; ++m) {
        const int n = fg + FG * m;
        if (n < N) {
          const float a = sScore[n];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 ef = __bfloat1622float2(enr[m][e]);
            c8[2 * e] += a * ef.x; c8[2 * e + 1] += a * ef.y;
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) sC[fg * H + d0 + e] = c8[e];
      phase_stamp(p.dbg, i, 11);
      __syncthreads();
      phase_stamp(p.dbg, i, 12);
      if (video_ok) {
        for (int d = tid; d < H; d += DEC_THREADS) {
          float cx = 0.f;
          for (int f = 0; f < FG; ++f) cx += sC[f * H + d];
          p.ctx_all[((long long)vb * L + i) * H + d] = cx;
          p.ctx_x[((long long)i * B + vb) * H + d] = __float2bfloat16_rn(cx);
        }
      }
      phase_stamp(p.dbg, i, 13);
    }
    group_arrive(ctr);
    target += (unsigned)C;
    phase_stamp(p.dbg, i, 6);
    // ---- P3: gi_c = W_c ctx -----------------------------------------------------------------------------
    if (TMA) {
      if (tid == 0) {
        if (ld_acquire_u32(ctr) < target) {
          const long long t0 = clock64();
          while (ld_acquire_u32(ctr) < target) {
            if (clock64() - t0 > 4000000000LL) __trap();
          }
        }
        phase_stamp(p.dbg, i, 7);
        fence_proxy_async();
        mbar_arrive_expect_tx(bar_x, x_bytes);
        for (int kb = 0; kb < KB; ++kb) tma_load_3d(sX + (size_t)kb * bsp * 128, &tmCtx, bar_x, kb * 64, b0, i);
        mbar_wait(bar_x, phase_x);
        tc_fence_after();
        issue_swapped_mma(tmem_d2, smem_u32(sW3), R3, smem_u32(sX), bsp, H, idesc, bar);
      }
      phase_x ^= 1;
    } else {
    group_wait(ctr, target);
    phase_stamp(p.dbg, i, 7);
    load_operand_rows_async(sX, bsp, 0, p.ctx_x + (long long)i * B * H, H, b0, bsp, b0 + bs, H);
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_swapped_mma(tmem_d2, smem_u32(sW3), R3, smem_u32(sX), bsp, H, idesc, bar);
    }
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    if (tid < 128) tmem_to_smem_cols(tmem_d2, sS3, s3_ld, R3, bsp);
    tc_fence_before();
    __syncthreads();
    // ---- P4: gates ----------------------------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < DEC_ITEMS; ++k) {
      if (k < n_items) {
        const int idx = tid + k * DEC_THREADS, jj = idx % u, lb = idx / u;
        if (lb < bs) {
          const int j = j0 + jj, b = b0 + lb;
          const float gir = sS3[lb * s3_ld + jj] + epr[k];
          const float giz = sS3[lb * s3_ld + u + jj] + epz[k];
          const float gin = sS3[lb * s3_ld + 2 * u + jj] + epn[k];
          const float ghr = sS1[lb * s1_ld + u + jj] + bhr[k];
          const float ghz = sS1[lb * s1_ld + 2 * u + jj] + bhz[k];
          const float ghn = sS1[lb * s1_ld + 3 * u + jj] + bhn[k];
          const float r = sigmoidf_(gir + ghr);
          const float z = sigmoidf_(giz + ghz);
          const float n = fast_tanh(gin + r * ghn);
          const float hn = (1.f - z) * n + z * hreg[k];
          hreg[k] = hn;
          p.hs[((long long)b * L + i) * H + j] = hn;
          p.hs_a[((long long)b * L + i) * p.hs_a_ld + j] = __float2bfloat16_rn(hn);
          const long long o = ((long long)i * B + b) * H + j;
          p.r[o] = r; p.z[o] = z; p.n[o] = n; p.ghn[o] = ghn;
        }
      }
    }
    group_arrive(ctr);
    target += (unsigned)C;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_d1, ncols);
  }
}

// ---- backward sweep -----------------------------------------------------------------------------------------
// Per step i (reverse):  B1  gate gradients of this CTA's (unit, video) pairs            -> [drp|dzp|dghn|dnp] exchanged
//                        B2  dctx = W_c^T dgi  and the W_hh^T dgh part of dh_{i-1}       (tcgen05, K in chunks of H)
//                        B3  attention gradient of video vb: d alpha, d score, dq         -> dq exchanged
//                        B4  dh_{i-1} += W_q^T dq                                         (tcgen05), carry in registers
// d enc / d proj_key / d v do not feed back into the recurrence and are accumulated after the sweep by
// attn_grad_hoisted from the saved alpha, d score, dctx and q.
// The two products of a step have 16-row weight slices and a long contraction (3H / 4H): they run on warp-level
// mma.sync (K range of every chunk split over the 8 warps, partial tiles reduced through shared memory), which is
// several times faster here than tcgen05.mma's ~70 cycles per K=16 step (persist.cuh); the K-chunks of the exchange
// buffer stream in with cp.async, two in flight, overlapping the MMAs of the previous chunk.
constexpr int DEC_BWD_THREADS = DEC_THREADS;
// ITEMS: (unit, video) pairs per thread (2 at the cfg2 shape: u * C = 512 pairs on 256 threads).
template <int NF, bool ACC_TANH, bool TMA = true, int ITEMS = DEC_ITEMS>
__global__ void __launch_bounds__(DEC_BWD_THREADS, 1) dec_persist_bwd_kernel(const DecPersistBwd p,
                                                                             const __grid_constant__ CUtensorMap tmXg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, u = p.u, C = p.C, bsp = p.bsp, N = p.N, L = p.L, B = p.B, KBH = H >> 6;
  uint8_t* sWA = smem;                                         // 3*KBH k-blocks x (u rows x 128 B)
  uint8_t* sWB = sWA + (size_t)3 * KBH * u * 128;              // 4*KBH k-blocks
  uint8_t* sX0 = sWB + (size_t)4 * KBH * u * 128;              // chunk buffers: KBH x (bsp x 128 B) each
  uint8_t* sX1 = sX0 + (size_t)KBH * bsp * 128;
  float* sR = reinterpret_cast<float*>(sX1 + (size_t)KBH * bsp * 128);     // [8 warps][u][bsp] partial products
  float* sSA = sR + (size_t)(DEC_THREADS / 32) * u * (bsp + 1);
  const int s_ld = u + 1;
  float* sSB = sSA + (size_t)bsp * s_ld;
  const int DG = H >> 3, FG = DEC_THREADS / DG, PW = DG >= 32 ? DG / 32 : 1;
  float* sP = sSB + (size_t)bsp * s_ld;     // [N][PW]
  float* sDa = sP + (size_t)N * PW;         // [N] d alpha -> d score
  float* sAl = sDa + N;                     // [N] alpha
  float* sC = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sAl + N) + 15) & ~uintptr_t(15));      // [FG][H], float4 access
  uint64_t* bar0 = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sC + (size_t)FG * H) + 15) & ~uintptr_t(7));
  uint64_t* bar1 = bar0 + 1;                // TMA: chunk buffer X0 / X1 has landed
  uint32_t ph0 = 0, ph1 = 0;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = blockIdx.x / C, c = blockIdx.x % C;
  const int b0 = g * C, j0 = c * u;
  const int bs = min(C, B - b0);
  unsigned* ctr = p.counters + g * 32;

  load_operand_rows(sWA, u, 0, p.wcT, p.wcT_ld, j0, u, H, 3 * H);
  load_operand_rows(sWB, u, 0, p.wcatT, p.wcatT_ld, j0, u, H, 4 * H);
  if (TMA && tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    fence_barrier_init();
  }
  __syncthreads();
  const uint32_t aWA = smem_u32(sWA), aWB = smem_u32(sWB), aX0 = smem_u32(sX0), aX1 = smem_u32(sX1);
  const int lane = tid & 31, gid = lane >> 2, tig = lane & 3, nwarps = DEC_THREADS / 32;
  const int ksteps = H >> 4;               // k-steps of 16 per chunk
  // one 16-row m-tile (u == 16), bsp / 8 n-tiles of 8 videos
  auto mma_chunk = [&](float (&acc)[4][4], uint32_t aW, int wk0, uint32_t aX) {
    for (int ks = warp; ks < ksteps; ks += nwarps) {
      uint32_t a[4], bq[4];
      load_a_frag(aW, u, 0, wk0 + (ks << 4), a);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        if (np * 16 < bsp) {
          load_b_frag2(aX, bsp, np * 16, ks << 4, bq);
          mma_bf16_16816(acc[2 * np], a, bq[0], bq[1]);
          mma_bf16_16816(acc[2 * np + 1], a, bq[2], bq[3]);
        }
      }
    }
  };
  // partial tiles of the 8 warps -> shared memory [warp][row][bsp + 1] (padded: the summing threads walk rows fastest);
  // after the CTA barrier every thread sums the 8 partials of ITS (unit, video) pairs itself (tile_sum) -- no second
  // staging buffer, no second barrier before the consumer
  const int r_ld = bsp + 1;
  auto spill_tiles = [&](float (&acc)[4][4]) {
    float* r = sR + (size_t)warp * u * r_ld;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (nt * 8 < bsp) {
        const int col = nt * 8 + 2 * tig;
        if (gid < u) { r[gid * r_ld + col] = acc[nt][0]; r[gid * r_ld + col + 1] = acc[nt][1]; }
        if (gid + 8 < u) { r[(gid + 8) * r_ld + col] = acc[nt][2]; r[(gid + 8) * r_ld + col + 1] = acc[nt][3]; }
      }
    }
    __syncthreads();
  };
  auto tile_sum = [&](int row, int col) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < DEC_THREADS / 32; ++w) s += sR[((size_t)w * u + row) * r_ld + col];
    return s;
  };

  // attention residency (as in the forward kernel)
  const int vb = min(b0 + c, B - 1);
  const bool video_ok = (c < bs);
  const int dg = tid % DG, fg = tid / DG, d0 = dg * 8;
  __half2 pkr[NF][4];
  __nv_bfloat162 enr[NF][4];
  float v8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v8[e] = p.v[d0 + e];
#pragma unroll
  for (int m = 0; m < NF; ++m) {
    const int n = fg + FG * m;
#pragma unroll
    for (int e = 0; e < 4; ++e) { pkr[m][e] = __floats2half2_rn(0.f, 0.f); enr[m][e] = __floats2bfloat162_rn(0.f, 0.f); }
    if (n < N) {
      const float4* s4 = reinterpret_cast<const float4*>(p.pk + ((long long)vb * N + n) * H + d0);
      const float4 a = __ldg(s4), bq = __ldg(s4 + 1);
      const float lim = 60000.f;
      pkr[m][0] = __floats2half2_rn(fminf(fmaxf(a.x, -lim), lim), fminf(fmaxf(a.y, -lim), lim));
      pkr[m][1] = __floats2half2_rn(fminf(fmaxf(a.z, -lim), lim), fminf(fmaxf(a.w, -lim), lim));
      pkr[m][2] = __floats2half2_rn(fminf(fmaxf(bq.x, -lim), lim), fminf(fmaxf(bq.y, -lim), lim));
      pkr[m][3] = __floats2half2_rn(fminf(fmaxf(bq.z, -lim), lim), fminf(fmaxf(bq.w, -lim), lim));
      const uint4 ev = __ldg(reinterpret_cast<const uint4*>(p.enc_a + ((long long)vb * N + n) * p.enc_ld + d0));
      enr[m][0] = *reinterpret_cast<const __nv_bfloat162*>(&ev.x);
      enr[m][1] = *reinterpret_cast<const __nv_bfloat162*>(&ev.y);
      enr[m][2] = *reinterpret_cast<const __nv_bfloat162*>(&ev.z);
      enr[m][3] = *reinterpret_cast<const __nv_bfloat162*>(&ev.w);
    }
  }

  const int n_items = (u * C + DEC_THREADS - 1) / DEC_THREADS;
  float dhc[ITEMS];
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) dhc[k] = 0.f;
  unsigned target = 0;
  const long long xrow = (long long)5 * H;

  // Saved activations of a step are HBM-cold (written by the forward pass): they are prefetched into L2 one step
  // ahead, right after the first arrive of the previous step.  (Prefetching into registers instead spilled and cost
  // more in the products than it saved here: measured 19.4 -> 24.8 us per step.)
  auto l2_prefetch = [](const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); };
  auto fetch = [&](int i) {
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      if (k < n_items) {
        const int idx = tid + k * DEC_THREADS, jj = idx % u, lb = idx / u;
        if (lb < bs && (jj & 7) == 0) {              // one request per 32-byte sector
          const int j = j0 + jj, b = b0 + lb;
          l2_prefetch(p.d_hs + ((long long)b * L + i) * H + j);
          const long long o = ((long long)i * B + b) * H + j;
          l2_prefetch(p.r + o); l2_prefetch(p.z + o); l2_prefetch(p.n + o); l2_prefetch(p.ghn + o);
          l2_prefetch(i > 0 ? p.hs + ((long long)b * L + (i - 1)) * H + j : p.h0 + (long long)b * p.h0_ld + j);
        }
      }
    }
    if (tid < N && (tid & 7) == 0) l2_prefetch(p.alpha + ((long long)i * B + vb) * N + tid);
    if (tid >= 32 && tid < 32 + (H >> 3)) l2_prefetch(p.q_all + (long long)i * B * p.q_ld + (long long)vb * p.q_ld + (tid - 32) * 8);
  };

  // Saved activations of the step to process next, in registers: loaded (from L2, where `fetch` put them a step earlier)
  // just before the last group barrier of the previous step, so that their latency hides behind that barrier.
  float sv[ITEMS][6];
  auto load_saved = [&](int i) {
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
#pragma unroll
      for (int a = 0; a < 6; ++a) sv[k][a] = 0.f;
      if (k < n_items) {
        const int idx = tid + k * DEC_THREADS, jj = idx % u, lb = idx / u;
        if (lb < bs) {
          const int j = j0 + jj, b = b0 + lb;
          const long long o = ((long long)i * B + b) * H + j;
          sv[k][0] = __ldcg(p.d_hs + ((long long)b * L + i) * H + j);
          sv[k][1] = __ldcg(p.r + o); sv[k][2] = __ldcg(p.z + o); sv[k][3] = __ldcg(p.n + o); sv[k][4] = __ldcg(p.ghn + o);
          sv[k][5] = i > 0 ? __ldcg(p.hs + ((long long)b * L + (i - 1)) * H + j) : __ldcg(p.h0 + (long long)b * p.h0_ld + j);
        }
      }
    }
  };
  load_saved(L - 1);

  for (int i = L - 1; i >= 0; --i) {
    bf16* xg = p.xg + (size_t)(i & 1) * B * xrow;
    phase_stamp(p.dbg, L - 1 - i, 0);
    // ---- B1: gate gradients ---------------------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      if (k < n_items) {
        const int idx = tid + k * DEC_THREADS, jj = idx % u, lb = idx / u;
        if (lb < bs) {
          const int j = j0 + jj, b = b0 + lb;
          const float dh = dhc[k] + sv[k][0];
          const float r = sv[k][1], z = sv[k][2], n = sv[k][3], ghn = sv[k][4];
          const float hp = sv[k][5];
          const float dn = dh * (1.f - z), dz = dh * (hp - n);
          const float dnp = dn * (1.f - n * n);
          const float dzp = dz * z * (1.f - z);
          const float drp = dnp * ghn * r * (1.f - r);
          const float dghn = dnp * r;
          bf16* x = xg + (long long)b * xrow;        // the exchange operand first: it is what the group waits for
          x[H + j] = __float2bfloat16_rn(drp); x[2 * H + j] = __float2bfloat16_rn(dzp);
          x[3 * H + j] = __float2bfloat16_rn(dghn); x[4 * H + j] = __float2bfloat16_rn(dnp);
          float* dgi = p.dgi_all + ((long long)b * L + i) * 3 * H;
          float* d1 = p.d1_all + ((long long)b * L + i) * 4 * H + H;
          dgi[j] = drp; dgi[H + j] = dzp; dgi[2 * H + j] = dnp;
          d1[j] = drp; d1[H + j] = dzp; d1[2 * H + j] = dghn;
          if (p.dgi_p) {
            bf16* q = p.dgi_p + ((long long)b * L + i) * p.dgi_p_ld;
            q[j] = __float2bfloat16_rn(drp); q[H + j] = __float2bfloat16_rn(dzp); q[2 * H + j] = __float2bfloat16_rn(dnp);
          }
          if (p.d1_p) {
            bf16* q = p.d1_p + ((long long)b * L + i) * p.d1_p_ld + H;
            const float keep = i > 0 ? 1.f : 0.f;
            q[j] = __float2bfloat16_rn(drp * keep); q[H + j] = __float2bfloat16_rn(dzp * keep);
            q[2 * H + j] = __float2bfloat16_rn(dghn * keep);
          }
          dhc[k] = dh * z;
        }
      }
    }
    phase_stamp(p.dbg, L - 1 - i, 1);
    group_arrive(ctr);
    target += (unsigned)C;
    if (i > 0) fetch(i - 1);
    // ---- B2: dctx = W_c^T [drp|dzp|dnp] ; dh part = W_hh^T [drp|dzp|dghn] -------------------------------------
    group_wait(ctr, target);
    phase_stamp(p.dbg, L - 1 - i, 2);
    float accA[4][4], accB[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) { accA[nt][e] = 0.f; accB[nt][e] = 0.f; }
    // chunks drp -> X0, dzp -> X1 in flight; then dnp -> X0, dghn -> X1 behind the MMAs that free the buffers
    if (TMA) {
      // (the group_wait above gave every thread the acquire; its CTA barrier also fences the previous readers of X0 / X1)
      if (tid == 0) {
        tma_fetch_operand(sX0, bsp, 0, &tmXg, bar0, H, KBH, b0, i & 1);
        tma_fetch_operand(sX1, bsp, 0, &tmXg, bar1, 2 * H, KBH, b0, i & 1);
      }
      mbar_wait(bar0, ph0); ph0 ^= 1;
      mma_chunk(accA, aWA, 0, aX0);                 // drp: W_c^T chunk 0, [W_q|W_hh]^T chunk 1
      mma_chunk(accB, aWB, H, aX0);
      __syncthreads();
      if (tid == 0) tma_fetch_operand(sX0, bsp, 0, &tmXg, bar0, 4 * H, KBH, b0, i & 1);
      mbar_wait(bar1, ph1); ph1 ^= 1;
      mma_chunk(accA, aWA, H, aX1);                 // dzp
      mma_chunk(accB, aWB, 2 * H, aX1);
      __syncthreads();
      if (tid == 0) tma_fetch_operand(sX1, bsp, 0, &tmXg, bar1, 3 * H, KBH, b0, i & 1);
      mbar_wait(bar0, ph0); ph0 ^= 1;
      mma_chunk(accA, aWA, 2 * H, aX0);             // dnp: completes dctx
      mbar_wait(bar1, ph1); ph1 ^= 1;
      mma_chunk(accB, aWB, 3 * H, aX1);             // dghn
    } else {
    load_operand_rows_async(sX0, bsp, 0, xg + H, xrow, b0, bsp, b0 + bs, H);
    cp_async_commit();
    load_operand_rows_async(sX1, bsp, 0, xg + 2 * H, xrow, b0, bsp, b0 + bs, H);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    mma_chunk(accA, aWA, 0, aX0);                 // drp: W_c^T chunk 0, [W_q|W_hh]^T chunk 1
    mma_chunk(accB, aWB, H, aX0);
    __syncthreads();
    load_operand_rows_async(sX0, bsp, 0, xg + 4 * H, xrow, b0, bsp, b0 + bs, H);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    mma_chunk(accA, aWA, H, aX1);                 // dzp
    mma_chunk(accB, aWB, 2 * H, aX1);
    __syncthreads();
    load_operand_rows_async(sX1, bsp, 0, xg + 3 * H, xrow, b0, bsp, b0 + bs, H);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    mma_chunk(accA, aWA, 2 * H, aX0);             // dnp: completes dctx
    cp_async_wait<0>();
    __syncthreads();
    mma_chunk(accB, aWB, 3 * H, aX1);             // dghn
    }
    phase_stamp(p.dbg, L - 1 - i, 3);
    phase_stamp(p.dbg, L - 1 - i, 4);
    spill_tiles(accA);
    {
      float* dc = p.dctx_all + (long long)i * B * H;
      for (int idx = tid; idx < u * bs; idx += DEC_THREADS) {
        const int jj = idx % u, lb = idx / u;
        dc[(long long)(b0 + lb) * H + j0 + jj] = tile_sum(jj, lb);
      }
    }
    phase_stamp(p.dbg, L - 1 - i, 5);
    // ---- B3: attention gradient of video vb ------------------------------------------------------------------
    // the saved attention weights of this step do not depend on the exchange: fetched while dctx is on its way
    if (tid < N) sAl[tid] = __ldg(p.alpha + ((long long)i * B + vb) * N + tid);
    phase_stamp(p.dbg, L - 1 - i, 6);
    {
      // dctx of this video arrives in 16-float pieces from the group's CTAs: polled on the data itself, no barrier
      float4 ca, cb;
      const float* drow = p.dctx_all + ((long long)i * B + vb) * H;
      if (p.xd == 1) {
        poll_f8(drow + d0, ca, cb);
      } else {
        if (warp == 0) {
          for (int c = tid * 8; c < H; c += 256) {
            float4 a, b;
            poll_f8(drow + c, a, b);
            *reinterpret_cast<float4*>(sC + c) = a; *reinterpret_cast<float4*>(sC + c + 4) = b;
          }
        }
        __syncthreads();
        ca = *reinterpret_cast<const float4*>(sC + d0); cb = *reinterpret_cast<const float4*>(sC + d0 + 4);
      }
      const float dc8[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
      const float4* q4 = reinterpret_cast<const float4*>(p.q_all + (long long)i * B * p.q_ld + (long long)vb * p.q_ld + d0);
      const float4 qa = __ldg(q4), qb = __ldg(q4 + 1);
      const float q8[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
      const int W = DG < 32 ? DG : 32;
      float da[NF];
#pragma unroll
      for (int m = 0; m < NF; ++m) {
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 ef = __bfloat1622float2(enr[m][e]);
          s += dc8[2 * e] * ef.x + dc8[2 * e + 1] * ef.y;
        }
        da[m] = s;
      }
      if (W == 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int m = 0; m < NF; ++m) da[m] += __shfl_xor_sync(0xffffffffu, da[m], o);
        }
      } else {
        for (int o = W >> 1; o > 0; o >>= 1) {
#pragma unroll
          for (int m = 0; m < NF; ++m) da[m] += __shfl_xor_sync(0xffffffffu, da[m], o);
        }
      }
#pragma unroll
      for (int m = 0; m < NF; ++m) {
        const int n = fg + FG * m;
        if ((tid % W) == 0 && n < N) sP[n * PW + dg / 32] = da[m];
      }
      __syncthreads();
      // d score_n = alpha_n (d alpha_n - sum_m alpha_m d alpha_m): warp 0, lanes over frames
      if (warp == 0) {
        float dot = 0.f;
        for (int n = tid; n < N; n += 32) {
          float s = 0.f;
          for (int w = 0; w < PW; ++w) s += sP[n * PW + w];
          const float al = sAl[n];
          sDa[n] = s;
          dot += al * s;
        }
        dot = warp_sum(dot);
        for (int n = tid; n < N; n += 32) {
          const float ds = sAl[n] * (sDa[n] - dot);
          sDa[n] = ds;
          if (video_ok) p.ds_all[((long long)i * B + vb) * N + n] = ds;
        }
      }
      __syncthreads();
      float dq8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int m = 0; m < NF; ++m) {
        const int n = fg + FG * m;
        if (n < N) {
          const float ds = sDa[n];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 pf = __half22float2(pkr[m][e]);
            if (ACC_TANH) {        // derivative accurate to ~1e-7 relative (ex2 + rcp): see tanh_sech2
              float t0, t1, g0, g1;
              tanh_sech2(q8[2 * e] + pf.x, t0, g0);
              tanh_sech2(q8[2 * e + 1] + pf.y, t1, g1);
              dq8[2 * e] += ds * g0;
              dq8[2 * e + 1] += ds * g1;
            } else {
              const float e0 = tanh_approx(q8[2 * e] + pf.x), e1 = tanh_approx(q8[2 * e + 1] + pf.y);
              dq8[2 * e] += ds * (1.f - e0 * e0);
              dq8[2 * e + 1] += ds * (1.f - e1 * e1);
            }
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) sC[fg * H + d0 + e] = dq8[e] * v8[e];
      __syncthreads();
      if (video_ok) {
        for (int d = tid; d < H; d += DEC_THREADS) {
          float dq = 0.f;
          for (int f = 0; f < FG; ++f) dq += sC[f * H + d];
          p.d1_all[((long long)vb * L + i) * 4 * H + d] = dq;
          xg[(long long)vb * xrow + d] = __float2bfloat16_rn(dq);
          if (p.d1_p) p.d1_p[((long long)vb * L + i) * p.d1_p_ld + d] = __float2bfloat16_rn(i > 0 ? dq : 0.f);
        }
      }
    }
    phase_stamp(p.dbg, L - 1 - i, 7);
    group_arrive(ctr);
    target += (unsigned)C;
    // ---- B4: dh_{i-1} = dh z + W_hh^T dgh + W_q^T dq -----------------------------------------------------------
    if (i > 0) load_saved(i - 1);
    group_wait(ctr, target);
    phase_stamp(p.dbg, L - 1 - i, 8);
    if (TMA) {
      if (tid == 0) tma_fetch_operand(sX0, bsp, 0, &tmXg, bar0, 0, KBH, b0, i & 1);
      mbar_wait(bar0, ph0); ph0 ^= 1;
    } else {
    load_operand_rows_async(sX0, bsp, 0, xg, xrow, b0, bsp, b0 + bs, H);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    }
    mma_chunk(accB, aWB, 0, aX0);                 // dq: completes dh
    phase_stamp(p.dbg, L - 1 - i, 9);
    phase_stamp(p.dbg, L - 1 - i, 10);
    spill_tiles(accB);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      if (k < n_items) {
        const int idx = tid + k * DEC_THREADS, jj = idx % u, lb = idx / u;
        if (lb < bs) dhc[k] += tile_sum(jj, lb);
      }
    }
    __syncthreads();          // sR is rewritten by the next step's B2
  }
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    if (k < n_items) {
      const int idx = tid + k * DEC_THREADS, jj = idx % u, lb = idx / u;
      if (lb < bs) p.dh_carry[(long long)(b0 + lb) * H + j0 + jj] = dhc[k];
    }
  }
}

// One CTA per (video, 64-dim tile): thread = (dim, frame group); frames fg + 4m.  alpha, d score, q and dctx of all L
// steps are staged in shared memory first, so the L x frames inner loops run without global-memory latency.
constexpr int AG_DIMS = 64, AG_FG = 4;
template <int NF, bool ACC_TANH>
__global__ void __launch_bounds__(AG_DIMS * AG_FG) attn_grad_hoisted_kernel(const AttnGradArgs a) {
  extern __shared__ float ag_sm[];
  const int L = a.L, B = a.B, N = a.N, H = a.H;
  float* sAl = ag_sm;                       // [L][N]
  float* sDs = sAl + L * N;                 // [L][N]
  float* sQ = sDs + L * N;                  // [L][AG_DIMS]
  float* sDc = sQ + L * AG_DIMS;            // [L][AG_DIMS]
  float* sV = sDc + L * AG_DIMS;            // [AG_FG][AG_DIMS]
  const int b = blockIdx.x, dl = threadIdx.x % AG_DIMS, d = blockIdx.y * AG_DIMS + dl, fg = threadIdx.x / AG_DIMS;
  for (int i = threadIdx.x; i < L * N; i += blockDim.x) {
    const int l = i / N, n = i - l * N;
    sAl[i] = a.alpha[((long long)l * B + b) * N + n];
    sDs[i] = a.ds[((long long)l * B + b) * N + n];
  }
  for (int i = threadIdx.x; i < L * AG_DIMS; i += blockDim.x) {
    const int l = i / AG_DIMS, dd = blockIdx.y * AG_DIMS + (i - l * AG_DIMS);
    sQ[i] = dd < H ? a.q[(long long)l * B * a.q_ld + (long long)b * a.q_ld + dd] : 0.f;
    sDc[i] = dd < H ? a.dctx[((long long)l * B + b) * H + dd] : 0.f;
  }
  const bool ok = d < H;
  float pk[NF], acc_pk[NF], acc_en[NF];
#pragma unroll
  for (int m = 0; m < NF; ++m) {
    const int n = fg + AG_FG * m;
    // proj_key exactly as the sweep kernels hold it (fp16, clamped): the gradient of the function the forward computed
    pk[m] = (ok && n < N) ? __half2float(__float2half_rn(fminf(fmaxf(a.pk[((long long)b * N + n) * H + d], -60000.f), 60000.f)))
                          : 0.f;
    acc_pk[m] = 0.f; acc_en[m] = 0.f;
  }
  __syncthreads();
  float dv = 0.f;
  for (int l = 0; l < L; ++l) {
    const float q = sQ[l * AG_DIMS + dl], dc = sDc[l * AG_DIMS + dl];
    const float* dsl = sDs + l * N + fg;
    const float* all = sAl + l * N + fg;
#pragma unroll
    for (int m = 0; m < NF; ++m) {
      if (fg + AG_FG * m < N) {
        const float ds = dsl[AG_FG * m];
        float e, g;
        if (ACC_TANH) tanh_sech2(q + pk[m], e, g);   // accurate derivative (ex2 + rcp)
        else { e = tanh_approx(q + pk[m]); g = 1.f - e * e; }     // hardware tanh (2^-11): +1e-3 on dW_k, 2x cheaper
        acc_pk[m] += ds * g;
        acc_en[m] += all[AG_FG * m] * dc;
        dv += ds * e;
      }
    }
  }
  const float vd = ok ? a.v[d] : 0.f;
#pragma unroll
  for (int m = 0; m < NF; ++m) {
    const int n = fg + AG_FG * m;
    if (ok && n < N) {
      a.dpk[((long long)b * N + n) * H + d] = acc_pk[m] * vd;
      if (a.dpk_a) a.dpk_a[((long long)b * N + n) * a.dpk_a_ld + d] = __float2bfloat16_rn(acc_pk[m] * vd);
      a.denc[((long long)b * N + n) * H + d] = acc_en[m];
    }
  }
  sV[fg * AG_DIMS + dl] = dv;
  __syncthreads();
  if (fg == 0 && ok) {
    float s = 0.f;
    for (int f = 0; f < AG_FG; ++f) s += sV[f * AG_DIMS + dl];
    a.dv_part[(long long)b * H + d] = s;
  }
}

int attn_grad_hoisted(const AttnGradArgs& a, cudaStream_t st) {
  PVCR_REQUIRE(a.N <= AG_FG * 20, "attn_grad_hoisted: N=%d > %d frames", a.N, AG_FG * 20);
  const size_t smem = ((size_t)2 * a.L * a.N + (size_t)2 * a.L * AG_DIMS + AG_FG * AG_DIMS) * sizeof(float);
  PVCR_REQUIRE(smem <= 48 * 1024, "attn_grad_hoisted: L=%d N=%d needs %zu B of shared memory", a.L, a.N, smem);
  const dim3 grid(a.B, cdiv(a.H, AG_DIMS));
  LaunchScope ls_(KC_ATTN, st);
  static const bool approx = getenv("PVCR_TANH_ACCURATE_BWD") == nullptr;
  if (approx) {
    if (a.N <= AG_FG * 10) attn_grad_hoisted_kernel<10, false><<<grid, AG_DIMS * AG_FG, smem, st>>>(a);
    else attn_grad_hoisted_kernel<20, false><<<grid, AG_DIMS * AG_FG, smem, st>>>(a);
  } else {
    if (a.N <= AG_FG * 10) attn_grad_hoisted_kernel<10, true><<<grid, AG_DIMS * AG_FG, smem, st>>>(a);
    else attn_grad_hoisted_kernel<20, true><<<grid, AG_DIMS * AG_FG, smem, st>>>(a);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

// ---- host side -------------------------------------------------------------------------------------------
struct DecPlan { int C, u, bsp, G, NF; size_t smem; };

static int dec_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

static bool plan_dec(int B, int N, int H, DecPlan& pl) {
  if (H < 64 || H > 512 || (H & (H - 1)) != 0) return false;
  pl.C = H / 8 < 32 ? H / 8 : 32;
  pl.u = H / pl.C;
  if (pl.u % 8 != 0 || 4 * pl.u > 128) return false;
  pl.bsp = (pl.C + 15) / 16 * 16;
  pl.G = (B + pl.C - 1) / pl.C;
  if ((long long)pl.G * pl.C > dec_num_sms()) return false;
  if (pl.u * pl.C > DEC_ITEMS * DEC_THREADS) return false;
  const int DG = H / 8, FG = DEC_THREADS / DG, PW = DG >= 32 ? DG / 32 : 1;
  const int nf = (N + FG - 1) / FG;
  if (nf > 10) return false;
  pl.NF = nf <= 2 ? 2 : (nf <= 5 ? 5 : 10);
  const size_t KB = H / 64;
  size_t s = KB * 4 * pl.u * 128 + KB * 3 * pl.u * 128 + KB * pl.bsp * 128;
  s += ((size_t)pl.bsp * (4 * pl.u + 1) + (size_t)pl.bsp * (3 * pl.u + 1) + (size_t)N * PW + N + (size_t)FG * H) * 4;
  s += 64 + 16 * 1024 + 1024;      // barrier/slot, over-read tail of the last W3 k-block, alignment slack
  pl.smem = s;
  return s <= 227 * 1024;
}

bool dec_persist_eligible(int B, int N, int H, int nsplit, int Hp) {
  DecPlan pl;
  static const bool off = getenv("PVCR_NO_PERSIST_DEC") != nullptr;     // A/B knob for profiling
  return !off && nsplit == 1 && Hp == H && plan_dec(B, N, H, pl);
}

int dec_persist_fwd(const DecPersistFwd& p0, cudaStream_t st) {
  DecPlan pl;
  PVCR_REQUIRE(plan_dec(p0.B, p0.N, p0.H, pl), "dec_persist_fwd: shape B=%d N=%d H=%d not supported", p0.B, p0.N, p0.H);
  DecPersistFwd p = p0;
  p.C = pl.C; p.u = pl.u; p.bsp = pl.bsp;
  p.dbg = getenv("PVCR_PHASE_DEC_BWD") || getenv("PVCR_PHASE_GRU") ? nullptr : debug_phase_buffer();
  static const bool acc = getenv("PVCR_TANH_ACCURATE_FWD") != nullptr;      // A/B knobs (NF = 10 shape only)
  static const bool no_tma = getenv("PVCR_NO_TMA_XCHG") != nullptr;
  const void* kern = pl.NF == 2 ? (no_tma ? (const void*)dec_persist_fwd_kernel<2, false, false> : (const void*)dec_persist_fwd_kernel<2>)
                     : (pl.NF == 5 ? (no_tma ? (const void*)dec_persist_fwd_kernel<5, false, false> : (const void*)dec_persist_fwd_kernel<5>)
                                   : (acc ? (const void*)dec_persist_fwd_kernel<10, true>
                                          : (no_tma ? (const void*)dec_persist_fwd_kernel<10, false, false>
                                                    : (const void*)dec_persist_fwd_kernel<10>)));
  // tensor maps of the exchanged operands: (k, video, step) views with one (64 x bsp) box per k-block
  CUtensorMap tmH0, tmHs, tmCtx;
  PVCR_TRY(make_tensor_map(&tmH0, OperandView{p.h0_a, p.h0_a_ld, 0, p.B, 1}, p.H, pl.bsp));
  PVCR_TRY(make_tensor_map(&tmHs, OperandView{p.hs_a, (long long)p.L * p.hs_a_ld, p.hs_a_ld, p.B, p.L}, p.H, pl.bsp));
  PVCR_TRY(make_tensor_map(&tmCtx, OperandView{p.ctx_x, (long long)p.H, (long long)p.B * p.H, p.B, p.L}, p.H, pl.bsp));
  PVCR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  int per_sm = 0;
  PVCR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DEC_THREADS, pl.smem));
  const int grid = pl.G * pl.C;
  PVCR_REQUIRE(per_sm * dec_num_sms() >= grid, "dec_persist_fwd: %d CTAs cannot be co-resident", grid);
  PVCR_TRY(fill_zero(p.counters, sizeof(unsigned) * 32 * pl.G, st));
  static const int xq = getenv("PVCR_DEC_FWD_XCHG") ? atoi(getenv("PVCR_DEC_FWD_XCHG")) : 2;     // A/B knob
  p.xq = xq;
  if (p.xq)      // arm the q slots with the sentinel the consumers poll for (0xFF bytes; persist.cuh)
    PVCR_CUDA_CHECK(cudaMemset2DAsync(p.q_all, sizeof(float) * p.q_ld, 0xFF, sizeof(float) * p.H, (size_t)p.L * p.B, st));
  void* args[] = {&p, &tmH0, &tmHs, &tmCtx};
  LaunchScope ls_(KC_DEC_FWD, st);
  PVCR_CUDA_CHECK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(DEC_THREADS), args, pl.smem, st));
  return PVCR_OK;
}

int dec_persist_bwd(const DecPersistBwd& p0, cudaStream_t st) {
  DecPlan pl;
  PVCR_REQUIRE(plan_dec(p0.B, p0.N, p0.H, pl), "dec_persist_bwd: shape B=%d N=%d H=%d not supported", p0.B, p0.N, p0.H);
  DecPersistBwd p = p0;
  p.C = pl.C; p.u = pl.u; p.bsp = pl.bsp;
  p.dbg = getenv("PVCR_PHASE_DEC_BWD") ? debug_phase_buffer() : nullptr;
  const int H = p.H, KBH = H / 64, DG = H / 8, FG = DEC_THREADS / DG, PW = DG >= 32 ? DG / 32 : 1;
  size_t smem = (size_t)7 * KBH * pl.u * 128 + (size_t)2 * KBH * pl.bsp * 128;
  smem += ((size_t)2 * pl.bsp * (pl.u + 1) + (size_t)p.N * PW + 2 * p.N + (size_t)FG * H) * 4 + 64 + 16 + 1024;
  smem += (size_t)(DEC_THREADS / 32) * pl.u * (pl.bsp + 1) * 4;  // partial product tiles of the 8 warps (padded rows)
  PVCR_REQUIRE(pl.u == 16 || pl.u == 8, "dec_persist_bwd: unit slice u=%d not supported by the mma.sync tiling", pl.u);
  PVCR_REQUIRE(smem <= 227 * 1024, "dec_persist_bwd: needs %zu B of shared memory", smem);
  // PVCR_TANH_ACCURATE_BWD=1: ex2 + rcp instead of the single-MUFU hardware tanh (2^-11) in the attention gradient.
  // Measured (B = 128 vs the operand-rounded oracle): no gradient error changes in its first three digits, +0.11 ms per step.
  static const bool approx = getenv("PVCR_TANH_ACCURATE_BWD") == nullptr;
  static const bool no_tma = getenv("PVCR_NO_TMA_XCHG") != nullptr;         // A/B knob: cp.async exchange loads
  const void* kern = no_tma ? (pl.NF == 2 ? (const void*)dec_persist_bwd_kernel<2, false, false>
                               : (pl.NF == 5 ? (const void*)dec_persist_bwd_kernel<5, false, false>
                                             : (const void*)dec_persist_bwd_kernel<10, false, false>))
                     : approx ? (pl.NF == 2 ? (const void*)dec_persist_bwd_kernel<2, false>
                               : (pl.NF == 5 ? (const void*)dec_persist_bwd_kernel<5, false> : (const void*)dec_persist_bwd_kernel<10, false>))
                            : (pl.NF == 2 ? (const void*)dec_persist_bwd_kernel<2, true>
                               : (pl.NF == 5 ? (const void*)dec_persist_bwd_kernel<5, true> : (const void*)dec_persist_bwd_kernel<10, true>));
  if (!no_tma && approx && pl.NF == 10 && pl.u * pl.C <= 2 * DEC_THREADS) kern = (const void*)dec_persist_bwd_kernel<10, false, true, 2>;
  CUtensorMap tmXg;        // exchange buffer [2][B][5H] as (k, video, parity)
  PVCR_TRY(make_tensor_map(&tmXg, OperandView{p.xg, (long long)5 * p.H, (long long)p.B * 5 * p.H, p.B, 2}, 5 * p.H, pl.bsp));
  PVCR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  PVCR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DEC_BWD_THREADS, smem));
  const int grid = pl.G * pl.C;
  PVCR_REQUIRE(per_sm * dec_num_sms() >= grid, "dec_persist_bwd: %d CTAs cannot be co-resident", grid);
  PVCR_TRY(fill_zero(p.counters, sizeof(unsigned) * 32 * pl.G, st));
  static const int xd = getenv("PVCR_DEC_BWD_XCHG") ? atoi(getenv("PVCR_DEC_BWD_XCHG")) : 1;      // A/B knob (1 or 2)
  p.xd = xd == 2 ? 2 : 1;
  // arm the dctx slots with the sentinel the consumers poll for (0xFF bytes; persist.cuh)
  PVCR_CUDA_CHECK(cudaMemsetAsync(p.dctx_all, 0xFF, sizeof(float) * (size_t)p.L * p.B * p.H, st));
  void* args[] = {&p, &tmXg};
  LaunchScope ls_(KC_DEC_BWD, st);
  PVCR_CUDA_CHECK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(DEC_BWD_THREADS), args, smem, st));
  return PVCR_OK;
}

}  // namespace pvcr
