// Persistent GRU sequence kernels (forward and reverse-time backward): one cooperative launch runs all T
// timesteps with the CTA's slice of W_hh (forward) / W_hh^T (backward) resident in shared memory.
// Reference semantics: torch.nn.GRU over a sequence (model/S2VTAttModel.py:88-93, model/S2VTModel.py:84,107);
// gate order r,z,n;  n = tanh(gi_n + r * (W_hn h + b_hn));  h' = (1-z) n + z h.
//
// Work split (persist.cuh): groups of `bs` videos x C CTAs; CTA c of a group owns hidden units [c*u, (c+1)*u).
//   forward : rows {r,z,n} x u of W_hh (3u x H bf16) resident; per step D[3u, bs] = W_slice * h_{t-1}^T, then
//             the gate math for its (unit, video) pairs with h kept in fp32 registers across steps.
//   backward: rows u of W_hh^T (u x 3H bf16) resident; per step the gate gradients of its (unit, video) pairs,
//             exchange of dgh (bf16), then D[u, bs] = W_hh^T slice * dgh^T added into the fp32 dh carry registers.
#include <cstdlib>
#include <mutex>

#include "host.h"
#include "persist.cuh"

namespace pvcr {

#ifndef DBG_STALE_X
#define DBG_STALE_X 0       // experiment knob: read step 0's h every step (stale data) to isolate the load latency
#endif
#ifndef DBG_NO_GI_PREFETCH
#define DBG_NO_GI_PREFETCH 0
#endif
constexpr int MAX_ITEMS = 8;    // (unit, video) pairs per thread: u * bs <= MAX_ITEMS * PERSIST_THREADS

struct GruPersistFwd {
  int T, B, H, bs, C, u;
  const bf16* whh; long long whh_ld;        // [3H, ld] bf16
  const float* b_hh;
  const float* gi; long long gi_ts, gi_ld;  // step t rows: gi + t*gi_ts + b*gi_ld  (includes b_ih)
  const float* gi_b; long long gi_b_ts, gi_b_ld; int gi_b_from;
  const float* gi_bias;
  const float* h0; long long h0_ld;         // nullable
  const bf16* h0p; long long h0p_ld;
  float* h; long long h_ts, h_ld;
  bf16* hp; long long hp_ts, hp_ld;
  float *r, *z, *n, *ghn;                   // [T][B,H]
  unsigned* counters;
  long long* dbg;
};

template <bool TMA>
__global__ void __launch_bounds__(PERSIST_THREADS, 1) gru_persist_fwd_kernel(const GruPersistFwd p,
                                                                             const __grid_constant__ CUtensorMap tmH0,
                                                                             const __grid_constant__ CUtensorMap tmHp) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, u = p.u, bs = p.bs, KB = H >> 6, Rw = 3 * u;
  uint8_t* sW = smem;
  uint8_t* sX = sW + (size_t)KB * Rw * 128;
  float* sS = reinterpret_cast<float*>(sX + (size_t)KB * bs * 128);
  const int s_ld = Rw + 1;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sS + (size_t)bs * s_ld + 2);
  bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bar) + 7) & ~uintptr_t(7));
  uint64_t* bar_x = bar + 1;                  // TMA: the exchanged h_{t-1} rows have landed in sX
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = blockIdx.x / p.C, c = blockIdx.x % p.C;
  const int b0 = g * bs, j0 = c * u;
  unsigned* ctr = p.counters + g * 32;
  uint32_t phase_x = 0;

  // resident weights: local row q*u + jj  <-  W_hh row q*H + j0 + jj
  for (int q = 0; q < 3; ++q)
    load_operand_rows(sW, Rw, q * u, p.whh, p.whh_ld, (long long)q * H + j0, u, (long long)3 * H, H);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar_x, 1);
    fence_barrier_init();
  }
  const uint32_t ncols = bs <= 32 ? 32u : (bs <= 64 ? 64u : 128u);
  if (warp == 0) {
    tmem_alloc(tmem_slot, ncols);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = umma_idesc_bf16(128, bs);

  // this thread's (unit, video) pairs
  const int n_items = (u * bs + PERSIST_THREADS - 1) / PERSIST_THREADS;
  float hreg[MAX_ITEMS], bhr[MAX_ITEMS], bhz[MAX_ITEMS], bhn[MAX_ITEMS];
#pragma unroll
  for (int k = 0; k < MAX_ITEMS; ++k) {
    hreg[k] = 0.f; bhr[k] = 0.f; bhz[k] = 0.f; bhn[k] = 0.f;
    if (k < n_items) {
      const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
      if (lb < bs) {
        const int j = j0 + jj, b = b0 + lb;
        bhr[k] = p.b_hh[j]; bhz[k] = p.b_hh[H + j]; bhn[k] = p.b_hh[2 * H + j];
        if (p.h0 && b < p.B) hreg[k] = p.h0[(long long)b * p.h0_ld + j];
      }
    }
  }
  uint32_t phase = 0;

  for (int t = 0; t < p.T; ++t) {
    // prefetch this step's input projections (independent of the group barrier)
    float gir[MAX_ITEMS], giz[MAX_ITEMS], gin[MAX_ITEMS];
#pragma unroll
    for (int k = 0; k < MAX_ITEMS; ++k) {
      gir[k] = giz[k] = gin[k] = 0.f;
      if (k < n_items) {
        const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
        const int j = j0 + jj, b = b0 + lb;
        if (!DBG_NO_GI_PREFETCH && lb < bs && b < p.B) {
          if (p.gi) {
            const float* gp = p.gi + (long long)t * p.gi_ts + (long long)b * p.gi_ld;
            gir[k] = __ldg(gp + j); giz[k] = __ldg(gp + H + j); gin[k] = __ldg(gp + 2 * H + j);
          }
          if (p.gi_b && t >= p.gi_b_from) {
            const float* gq = p.gi_b + (long long)(t - p.gi_b_from) * p.gi_b_ts + (long long)b * p.gi_b_ld;
            gir[k] += __ldg(gq + j); giz[k] += __ldg(gq + H + j); gin[k] += __ldg(gq + 2 * H + j);
          }
          if (p.gi_bias) { gir[k] += p.gi_bias[j]; giz[k] += p.gi_bias[H + j]; gin[k] += p.gi_bias[2 * H + j]; }
        }
      }
    }
    const bool has_prev = (t > 0) || (p.h0p != nullptr);
    phase_stamp(p.dbg, t, 0);
    if (has_prev) {
      if (TMA) {
        if (tid == 0) {
          if (t > 0) spin_until(ctr, (unsigned)(p.C * t));
          phase_stamp(p.dbg, t, 1);
          tma_fetch_operand(sX, bs, 0, t > 0 ? &tmHp : &tmH0, bar_x, 0, KB, b0, t > 0 ? t - 1 : 0);
          mbar_wait(bar_x, phase_x);
          phase_stamp(p.dbg, t, 7);
          phase_stamp(p.dbg, t, 2);
          tc_fence_after();
          issue_swapped_mma(tmem_base, smem_u32(sW), Rw, smem_u32(sX), bs, H, idesc, bar);
        }
        phase_x ^= 1;
      } else {
      if (t > 0) {
        group_wait(ctr, (unsigned)(p.C * t));
        phase_stamp(p.dbg, t, 1);
        load_operand_rows_async(sX, bs, 0, p.hp + (long long)(DBG_STALE_X ? 0 : t - 1) * p.hp_ts, p.hp_ld, b0, bs, p.B, H);
      } else {
        load_operand_rows_async(sX, bs, 0, p.h0p, p.h0p_ld, b0, bs, p.B, H);
      }
      cp_async_commit();
      cp_async_wait<0>();
      phase_stamp(p.dbg, t, 7);
      fence_proxy_async();
      __syncthreads();
      phase_stamp(p.dbg, t, 2);
      if (tid == 0) {
        tc_fence_after();
        issue_swapped_mma(tmem_base, smem_u32(sW), Rw, smem_u32(sX), bs, H, idesc, bar);
      }
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      phase_stamp(p.dbg, t, 3);
      if (tid < 128) tmem_to_smem_cols(tmem_base, sS, s_ld, Rw, bs);
      tc_fence_before();
      __syncthreads();
      phase_stamp(p.dbg, t, 4);
    }
#pragma unroll
    for (int k = 0; k < MAX_ITEMS; ++k) {
      if (k < n_items) {
        const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
        const int j = j0 + jj, b = b0 + lb;
        if (lb < bs && b < p.B) {
          float ghr = bhr[k], ghz = bhz[k], ghn = bhn[k];
          if (has_prev) {
            ghr += sS[lb * s_ld + jj]; ghz += sS[lb * s_ld + u + jj]; ghn += sS[lb * s_ld + 2 * u + jj];
          }
          const float r = sigmoidf_(gir[k] + ghr);
          const float z = sigmoidf_(giz[k] + ghz);
          const float n = fast_tanh(gin[k] + r * ghn);
          const float hn = (1.f - z) * n + z * hreg[k];
          hreg[k] = hn;
          p.h[(long long)t * p.h_ts + (long long)b * p.h_ld + j] = hn;
          p.hp[(long long)t * p.hp_ts + (long long)b * p.hp_ld + j] = __float2bfloat16_rn(hn);
          const long long o = ((long long)t * p.B + b) * H + j;
          p.r[o] = r; p.z[o] = z; p.n[o] = n; p.ghn[o] = ghn;
        }
      }
    }
    phase_stamp(p.dbg, t, 5);
    group_arrive(ctr);      // also protects sS / sX reuse by the next step
    phase_stamp(p.dbg, t, 6);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ncols);
  }
}

struct GruPersistBwd {
  int T, B, H, bs, C, u;
  const bf16* whhT; long long whhT_ld;      // [H, ld] bf16: element (j, k) = W_hh[k, j], k in [0, 3H)
  const float* dh_ext; long long dh_ext_ts, dh_ext_ld;   // nullable
  float* dh_carry;                          // [B,H] in: gradient on the final state; out: gradient on h0
  const float *r, *z, *n, *ghn;             // [T][B,H]
  const float* h; long long h_ts, h_ld;     // forward states (h_{t-1} = h + (t-1)*h_ts)
  const float* h0; long long h0_ld;         // nullable
  float* dgi; long long dgi_ts, dgi_ld;
  float* dgh; long long dgh_ts, dgh_ld;
  bf16* xch;                                // exchange [2][B][3H] bf16
  bf16* dgi_p; long long dgi_p_ts, dgi_p_ld;  // optional bf16 copies (nullable)
  bf16* dgh_p; long long dgh_p_ts, dgh_p_ld;
  unsigned* counters;
};

template <bool TMA>
__global__ void __launch_bounds__(PERSIST_THREADS, 1) gru_persist_bwd_kernel(const GruPersistBwd p,
                                                                             const __grid_constant__ CUtensorMap tmX) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, u = p.u, bs = p.bs, K = 3 * H, KB = K >> 6;
  uint8_t* sW = smem;                                        // KB x (u rows x 128 B)
  uint8_t* sX = sW + (size_t)KB * u * 128;                   // KB x (bs rows x 128 B)
  float* sR = reinterpret_cast<float*>(sX + (size_t)KB * bs * 128);     // [8 warps][u][bs] K-slice partial products
  const int nwarps = PERSIST_THREADS / 32;
  uint64_t* bar_x = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sR + (size_t)nwarps * u * bs) + 15) & ~uintptr_t(7));
  uint32_t phase_x = 0;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = blockIdx.x / p.C, c = blockIdx.x % p.C;
  const int b0 = g * bs, j0 = c * u;
  unsigned* ctr = p.counters + g * 32;

  load_operand_rows(sW, u, 0, p.whhT, p.whhT_ld, j0, u, H, K);
  if (TMA && tid == 0) {
    mbar_init(bar_x, 1);
    fence_barrier_init();
  }
  __syncthreads();
  const uint32_t aW = smem_u32(sW), aX = smem_u32(sX);
  const int lane = tid & 31, gid = lane >> 2, tig = lane & 3;

  const int n_items = (u * bs + PERSIST_THREADS - 1) / PERSIST_THREADS;
  float dhc[MAX_ITEMS];
#pragma unroll
  for (int k = 0; k < MAX_ITEMS; ++k) {
    dhc[k] = 0.f;
    if (k < n_items) {
      const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
      if (lb < bs && b0 + lb < p.B) dhc[k] = p.dh_carry[(long long)(b0 + lb) * H + j0 + jj];
    }
  }
  unsigned arrivals = 0;

  // Saved activations of a step (written by the forward pass long ago: HBM latency) are fetched one step ahead,
  // right after the arrive of the previous step, so their latency hides behind the exchange and the product.
  float pe[MAX_ITEMS], pr[MAX_ITEMS], pz[MAX_ITEMS], pn[MAX_ITEMS], pg[MAX_ITEMS], ph[MAX_ITEMS];
  auto fetch = [&](int t) {
#pragma unroll
    for (int k = 0; k < MAX_ITEMS; ++k) {
      pe[k] = pr[k] = pz[k] = pn[k] = pg[k] = ph[k] = 0.f;
      if (k < n_items) {
        const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
        const int j = j0 + jj, b = b0 + lb;
        if (lb < bs && b < p.B) {
          if (p.dh_ext) pe[k] = __ldg(p.dh_ext + (long long)t * p.dh_ext_ts + (long long)b * p.dh_ext_ld + j);
          const long long o = ((long long)t * p.B + b) * H + j;
          pr[k] = __ldg(p.r + o); pz[k] = __ldg(p.z + o); pn[k] = __ldg(p.n + o); pg[k] = __ldg(p.ghn + o);
          if (t > 0) ph[k] = __ldg(p.h + (long long)(t - 1) * p.h_ts + (long long)b * p.h_ld + j);
          else if (p.h0) ph[k] = __ldg(p.h0 + (long long)b * p.h0_ld + j);
        }
      }
    }
  };
  fetch(p.T - 1);

  for (int t = p.T - 1; t >= 0; --t) {
    const bool has_prev = (t > 0) || (p.h0 != nullptr);
    bf16* xw = p.xch + (size_t)(t & 1) * p.B * K;
    float zreg[MAX_ITEMS];
#pragma unroll
    for (int k = 0; k < MAX_ITEMS; ++k) {
      zreg[k] = 0.f;
      if (k < n_items) {
        const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
        const int j = j0 + jj, b = b0 + lb;
        if (lb < bs && b < p.B) {
          const float dh = dhc[k] + pe[k];
          const float r = pr[k], z = pz[k], n = pn[k], ghn = pg[k], hp = ph[k];
          const float dn = dh * (1.f - z), dz = dh * (hp - n);
          const float dnp = dn * (1.f - n * n);
          const float dzp = dz * z * (1.f - z);
          const float drp = dnp * ghn * r * (1.f - r);
          const float dghn = dnp * r;
          float* dgi = p.dgi + (long long)t * p.dgi_ts + (long long)b * p.dgi_ld;
          float* dgh = p.dgh + (long long)t * p.dgh_ts + (long long)b * p.dgh_ld;
          if (has_prev) {          // the exchange operand first: it is what the group waits for
            bf16* x = xw + (long long)b * K;
            x[j] = __float2bfloat16_rn(drp); x[H + j] = __float2bfloat16_rn(dzp); x[2 * H + j] = __float2bfloat16_rn(dghn);
          }
          dgi[j] = drp; dgi[H + j] = dzp; dgi[2 * H + j] = dnp;
          dgh[j] = drp; dgh[H + j] = dzp; dgh[2 * H + j] = dghn;
          if (p.dgi_p) {
            bf16* q = p.dgi_p + (long long)t * p.dgi_p_ts + (long long)b * p.dgi_p_ld;
            q[j] = __float2bfloat16_rn(drp); q[H + j] = __float2bfloat16_rn(dzp); q[2 * H + j] = __float2bfloat16_rn(dnp);
          }
          if (p.dgh_p) {      // rows of a step without a previous state (h_{-1} = 0) are written as zeros: the product
            bf16* q = p.dgh_p + (long long)t * p.dgh_p_ts + (long long)b * p.dgh_p_ld;      // with h_{t-1} can then
            const float keep = has_prev ? 1.f : 0.f;                                         // run on shifted h planes
            q[j] = __float2bfloat16_rn(drp * keep); q[H + j] = __float2bfloat16_rn(dzp * keep);
            q[2 * H + j] = __float2bfloat16_rn(dghn * keep);
          }
          dhc[k] = dh * z;
          zreg[k] = 1.f;
        }
      }
    }
    if (has_prev) {
      group_arrive(ctr);
      ++arrivals;
      if (t > 0) fetch(t - 1);
      if (TMA) {
        if (tid == 0) {
          spin_until(ctr, (unsigned)p.C * arrivals);
          tma_fetch_operand(sX, bs, 0, &tmX, bar_x, 0, KB, b0, t & 1);
        }
        mbar_wait(bar_x, phase_x);       // every thread: the async-proxy writes are visible to its ldmatrix reads
        phase_x ^= 1;
      } else {
      group_wait(ctr, (unsigned)p.C * arrivals);
      load_operand_rows_async(sX, bs, 0, xw, K, b0, bs, p.B, K);
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      }
      // D[u, bs] = W_hh^T slice [u, 3H] x dgh^T: warp w takes the k-steps w, w+8, ... (mma.sync m16n8k16, fp32 acc)
      float acc[2][2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
      for (int ks = warp; ks < (K >> 4); ks += nwarps) {
        uint32_t a0[4], a1[4], bq[4];
        load_a_frag(aW, u, 0, ks << 4, a0);
        load_b_frag2(aX, bs, 0, ks << 4, bq);
        mma_bf16_16816(acc[0][0], a0, bq[0], bq[1]);
        mma_bf16_16816(acc[0][1], a0, bq[2], bq[3]);
        if (u > 16) {
          load_a_frag(aW, u, 16, ks << 4, a1);
          mma_bf16_16816(acc[1][0], a1, bq[0], bq[1]);
          mma_bf16_16816(acc[1][1], a1, bq[2], bq[3]);
        }
      }
      {
        float* r = sR + (size_t)warp * u * bs;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (mt * 16 < u) {
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              const int row = mt * 16 + gid, col = nt * 8 + 2 * tig;
              r[row * bs + col] = acc[mt][nt][0]; r[row * bs + col + 1] = acc[mt][nt][1];
              r[(row + 8) * bs + col] = acc[mt][nt][2]; r[(row + 8) * bs + col + 1] = acc[mt][nt][3];
            }
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < MAX_ITEMS; ++k) {
        if (k < n_items && zreg[k] != 0.f) {
          const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
          float s = 0.f;
          for (int w = 0; w < nwarps; ++w) s += sR[(size_t)w * u * bs + jj * bs + lb];
          dhc[k] += s;
        }
      }
      __syncthreads();     // sR / sX are rewritten by the next step
    }
  }
#pragma unroll
  for (int k = 0; k < MAX_ITEMS; ++k) {
    if (k < n_items) {
      const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
      if (lb < bs && b0 + lb < p.B) p.dh_carry[(long long)(b0 + lb) * H + j0 + jj] = dhc[k];
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------
struct PersistPlan { int bs, C, u, G; size_t smem; };

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// units per CTA: largest u <= 32 dividing H with u % 8 == 0; videos per group: 16.
static bool plan_gru(int B, int H, int k_rows_fwd, PersistPlan& pl, bool backward) {
  if (H % 64 != 0 || H < 64) return false;
  int u = 32;
  while (u >= 8 && H % u != 0) u -= 8;
  if (u < 8) return false;
  pl.u = u; pl.C = H / u; pl.bs = 16; pl.G = (B + pl.bs - 1) / pl.bs;
  if (pl.u * pl.bs > MAX_ITEMS * PERSIST_THREADS) return false;
  if ((long long)pl.G * pl.C > num_sms()) return false;
  const size_t K = backward ? (size_t)3 * H : (size_t)H;
  const size_t rows = backward ? (size_t)u : (size_t)3 * u;
  const size_t w = (K / 64) * rows * 128, x = (K / 64) * pl.bs * 128;
  const size_t s = backward ? (size_t)(PERSIST_THREADS / 32) * pl.u * pl.bs * 4 + 64 : (size_t)pl.bs * (rows + 1) * 4 + 64;
  if (backward && (pl.u % 16 != 0 || pl.bs != 16)) return false;      // mma.sync tiling of the backward product
  size_t total = w + x + s;
  // the 128-row MMA tile of the last k-block over-reads (128 - rows) * 128 bytes past the weight slice
  const size_t need_tail = (128 - rows) * 128;
  if (x + s < need_tail) total += need_tail - (x + s);
  pl.smem = total + 1024;
  (void)k_rows_fwd;
  return pl.smem <= 227 * 1024;
}

static bool tma_xchg() {
  static const bool off = getenv("PVCR_NO_TMA_XCHG") != nullptr;      // A/B knob: cp.async exchange loads instead
  return !off;
}

static int coop_launch(const void* kern, int grid, size_t smem, void** args, cudaStream_t st, const char* what,
                       int cls) {
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> g(mu);
    PVCR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int per_sm = 0;
  PVCR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PERSIST_THREADS, smem));
  PVCR_REQUIRE(per_sm * num_sms() >= grid, "%s: %d CTAs cannot be co-resident (%d per SM)", what, grid, per_sm);
  LaunchScope ls_(cls, st);
  PVCR_CUDA_CHECK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(PERSIST_THREADS), args, smem, st));
  return PVCR_OK;
}

bool gru_persist_eligible(const GruSeq& s) {
  PersistPlan pl;
  static const bool off = getenv("PVCR_NO_PERSIST_GRU") != nullptr;     // A/B knob for profiling
  return !off && s.nsplit == 1 && s.sync != nullptr && s.hp != nullptr && s.Hp == s.H && plan_gru(s.B, s.H, 0, pl, false) &&
         plan_gru(s.B, s.H, 0, pl, true) && (s.h0 == nullptr || s.h0_planes != nullptr);
}

int gru_persist_fwd(const GruSeq& s, cudaStream_t st) {
  PersistPlan pl;
  PVCR_REQUIRE(plan_gru(s.B, s.H, 0, pl, false), "gru_persist_fwd: shape B=%d H=%d not supported", s.B, s.H);
  GruPersistFwd p{};
  p.T = s.T; p.B = s.B; p.H = s.H; p.bs = pl.bs; p.C = pl.C; p.u = pl.u;
  p.whh = s.whh.ptr; p.whh_ld = s.whh.ld; p.b_hh = s.b_hh;
  p.gi = s.gi_a; p.gi_ts = s.gi_a_ts; p.gi_ld = s.gi_a_ld;
  p.gi_b = s.gi_b; p.gi_b_ts = s.gi_b_ts; p.gi_b_ld = s.gi_b_ld; p.gi_b_from = s.gi_b_from;
  p.gi_bias = s.gi_bias;
  p.h0 = s.h0; p.h0_ld = s.h0_ld; p.h0p = s.h0 ? s.h0_planes : nullptr; p.h0p_ld = s.h0_planes_ld;
  p.h = s.h; p.h_ts = s.h_ts; p.h_ld = s.h_ld;
  p.hp = s.hp; p.hp_ts = s.hp_ts; p.hp_ld = s.hp_ld;
  p.r = s.r; p.z = s.z; p.n = s.n; p.ghn = s.ghn;
  p.counters = s.sync;
  p.dbg = getenv("PVCR_PHASE_GRU") ? debug_phase_buffer() : nullptr;
  PVCR_TRY(fill_zero(s.sync, sizeof(unsigned) * 32 * pl.G, st));
  // tensor maps of the exchanged h rows: (k, video, step) over the bf16 state planes; the initial state separately
  CUtensorMap tmH0, tmHp;
  PVCR_TRY(make_tensor_map(&tmHp, OperandView{p.hp, p.hp_ld, p.hp_ts, p.B, p.T}, p.H, pl.bs));
  if (p.h0p) PVCR_TRY(make_tensor_map(&tmH0, OperandView{p.h0p, p.h0p_ld, 0, p.B, 1}, p.H, pl.bs));
  else tmH0 = tmHp;
  void* args[] = {&p, &tmH0, &tmHp};
  // measured (cfg2 encoder, 16 KB per step): the bulk-tensor fetch is no faster than cp.async here (4.32 vs 4.04 us per step;
  // both are one L2 round trip), unlike the 32..128 KB exchanges of the other three sweeps -- off unless asked for
  static const bool fwd_tma = getenv("PVCR_GRU_FWD_TMA") != nullptr;
  const void* kern = (fwd_tma && tma_xchg()) ? (const void*)gru_persist_fwd_kernel<true> : (const void*)gru_persist_fwd_kernel<false>;
  return coop_launch(kern, pl.G * pl.C, pl.smem, args, st, "gru_persist_fwd", KC_GRU_FWD);
}

int gru_persist_bwd(const GruSeq& s, const GruSeqGrad& g, cudaStream_t st) {
  PersistPlan pl;
  PVCR_REQUIRE(plan_gru(s.B, s.H, 0, pl, true), "gru_persist_bwd: shape B=%d H=%d not supported", s.B, s.H);
  GruPersistBwd p{};
  p.T = s.T; p.B = s.B; p.H = s.H; p.bs = pl.bs; p.C = pl.C; p.u = pl.u;
  p.whhT = g.whhT.ptr; p.whhT_ld = g.whhT.ld;
  p.dh_ext = g.dh_ext; p.dh_ext_ts = g.dh_ext_ts; p.dh_ext_ld = g.dh_ext_ld;
  p.dh_carry = g.dh_carry;
  p.r = s.r; p.z = s.z; p.n = s.n; p.ghn = s.ghn;
  p.h = s.h; p.h_ts = s.h_ts; p.h_ld = s.h_ld;
  p.h0 = s.h0; p.h0_ld = s.h0_ld;
  p.dgi = g.dgi; p.dgi_ts = g.dgi_ts; p.dgi_ld = g.dgi_ld;
  p.dgh = g.dgh; p.dgh_ts = g.dgh_ts; p.dgh_ld = g.dgh_ld;
  p.xch = g.xch;
  p.dgi_p = g.dgi_p; p.dgi_p_ts = g.dgi_p_ts; p.dgi_p_ld = g.dgi_p_ld;
  p.dgh_p = g.dgh_p; p.dgh_p_ts = g.dgh_p_ts; p.dgh_p_ld = g.dgh_p_ld;
  p.counters = s.sync;
  PVCR_TRY(fill_zero(s.sync, sizeof(unsigned) * 32 * pl.G, st));
  CUtensorMap tmX;         // exchange buffer [2][B][3H] as (k, video, parity)
  PVCR_TRY(make_tensor_map(&tmX, OperandView{p.xch, (long long)3 * p.H, (long long)p.B * 3 * p.H, p.B, 2}, 3 * p.H, pl.bs));
  void* args[] = {&p, &tmX};
  const void* kern = tma_xchg() ? (const void*)gru_persist_bwd_kernel<true> : (const void*)gru_persist_bwd_kernel<false>;
  return coop_launch(kern, pl.G * pl.C, pl.smem, args, st, "gru_persist_bwd", KC_GRU_BWD);
}

}  // namespace pvcr
