// Persistent LSTM sequence kernels for the RationaleNet generator (torch.nn.LSTM, gate order i,f,g,o;
// model/RationaleNet.py:26-27,43): one cooperative launch per direction runs all N timesteps with the CTA's slice of
// W_hh (forward: rows {i,f,g,o} x u units = 4u x H, exactly one 128-row tcgen05 tile at u = 32) or W_hh^T
// (backward: u x 4H) resident in shared memory; c / dh / dc stay in fp32 registers.  Same group / barrier scheme
// as gru_persist.cu.  `rev` selects the time order (the reverse direction walks t = T-1 .. 0).
#include <cstdlib>
#include <mutex>

#include "host.h"
#include "persist.cuh"

namespace pvcr {

constexpr int LSTM_ITEMS = 4;
constexpr int LSTM_XB = 16;      // videos per exchange pass of the backward kernel

struct LstmPersistFwd {
  int T, B, H, bs, C, u, rev;
  const bf16* whh; long long whh_ld;        // [4H, ld]
  const float* b_hh;
  const float* gi; long long gi_ts, gi_ld;  // step t rows: gi + t*gi_ts + b*gi_ld, 4H columns (includes b_ih)
  float* h; long long h_ts, h_ld;
  bf16* hp; long long hp_ts, hp_ld;
  float *si, *sf, *sg, *so, *sc;            // saved [T][B,H]
  unsigned* counters;
};

__global__ void __launch_bounds__(PERSIST_THREADS, 1) lstm_persist_fwd_kernel(const LstmPersistFwd p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, u = p.u, bs = p.bs, KB = H >> 6, Rw = 4 * u;
  uint8_t* sW = smem;
  uint8_t* sX = sW + (size_t)KB * Rw * 128;
  float* sS = reinterpret_cast<float*>(sX + (size_t)KB * bs * 128);
  const int s_ld = Rw + 1;
  uint64_t* bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sS + (size_t)bs * s_ld + 2) + 7) & ~uintptr_t(7));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = blockIdx.x / p.C, c = blockIdx.x % p.C;
  const int b0 = g * bs, j0 = c * u;
  unsigned* ctr = p.counters + g * 32;

  for (int q = 0; q < 4; ++q)
    load_operand_rows(sW, Rw, q * u, p.whh, p.whh_ld, (long long)q * H + j0, u, (long long)4 * H, H);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  const uint32_t ncols = bs <= 32 ? 32u : 64u;
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

  const int n_items = (u * bs + PERSIST_THREADS - 1) / PERSIST_THREADS;
  float creg[LSTM_ITEMS], bh[LSTM_ITEMS][4];
#pragma unroll
  for (int k = 0; k < LSTM_ITEMS; ++k) {
    creg[k] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) bh[k][q] = 0.f;
    if (k < n_items) {
      const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
      if (lb < bs) {
#pragma unroll
        for (int q = 0; q < 4; ++q) bh[k][q] = p.b_hh[q * H + j0 + jj];
      }
    }
  }
  uint32_t phase = 0;

  for (int s = 0; s < p.T; ++s) {
    const int t = p.rev ? p.T - 1 - s : s, tp = p.rev ? t + 1 : t - 1;
    float a[LSTM_ITEMS][4];
#pragma unroll
    for (int k = 0; k < LSTM_ITEMS; ++k) {
#pragma unroll
      for (int q = 0; q < 4; ++q) a[k][q] = 0.f;
      if (k < n_items) {
        const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
        const int b = b0 + lb;
        if (lb < bs && b < p.B) {
          const float* gp = p.gi + (long long)t * p.gi_ts + (long long)b * p.gi_ld + j0 + jj;
#pragma unroll
          for (int q = 0; q < 4; ++q) a[k][q] = __ldg(gp + q * H) + bh[k][q];
        }
      }
    }
    if (s > 0) {
      group_wait(ctr, (unsigned)(p.C * s));
      load_operand_rows_async(sX, bs, 0, p.hp + (long long)tp * p.hp_ts, p.hp_ld, b0, bs, p.B, H);
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        issue_swapped_mma(tmem_base, smem_u32(sW), Rw, smem_u32(sX), bs, H, idesc, bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      if (tid < 128) tmem_to_smem_cols(tmem_base, sS, s_ld, Rw, bs);
      tc_fence_before();
      __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < LSTM_ITEMS; ++k) {
      if (k < n_items) {
        const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
        const int j = j0 + jj, b = b0 + lb;
        if (lb < bs && b < p.B) {
          float x[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) x[q] = a[k][q] + (s > 0 ? sS[lb * s_ld + q * u + jj] : 0.f);
          const float gi_ = sigmoidf_(x[0]), gf = sigmoidf_(x[1]), gg = fast_tanh(x[2]), go = sigmoidf_(x[3]);
          const float cn = gf * creg[k] + gi_ * gg;
          const float hn = go * fast_tanh(cn);
          creg[k] = cn;
          p.h[(long long)t * p.h_ts + (long long)b * p.h_ld + j] = hn;
          p.hp[(long long)t * p.hp_ts + (long long)b * p.hp_ld + j] = __float2bfloat16_rn(hn);
          const long long o = ((long long)t * p.B + b) * H + j;
          p.si[o] = gi_; p.sf[o] = gf; p.sg[o] = gg; p.so[o] = go; p.sc[o] = cn;
        }
      }
    }
    group_arrive(ctr);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ncols);
  }
}

struct LstmPersistBwd {
  int T, B, H, bs, C, u, rev;
  const bf16* whhT; long long whhT_ld;      // [H, ld]: element (j, k) = W_hh[k, j], k in [0, 4H)
  const float* dh_ext; long long dh_ext_ts, dh_ext_ld;
  const float *si, *sf, *sg, *so, *sc;      // saved [T][B,H]
  float* da; long long da_ts, da_ld;        // gate pre-activation gradients, 4H columns per (b, t)
  bf16* xch;                                // exchange [2][B][4H]
  unsigned* counters;
};

template <bool TMA>
__global__ void __launch_bounds__(PERSIST_THREADS, 1) lstm_persist_bwd_kernel(const LstmPersistBwd p,
                                                                              const __grid_constant__ CUtensorMap tmX) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, u = p.u, bs = p.bs, K = 4 * H, KB = K >> 6;
  uint8_t* sW = smem;
  uint8_t* sX = sW + (size_t)KB * u * 128;
  float* sR = reinterpret_cast<float*>(sX + (size_t)KB * LSTM_XB * 128);     // [8 warps][u][bs] K-slice partial products
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
  float dhc[LSTM_ITEMS], dcc[LSTM_ITEMS];
#pragma unroll
  for (int k = 0; k < LSTM_ITEMS; ++k) { dhc[k] = 0.f; dcc[k] = 0.f; }
  unsigned arrivals = 0;

  for (int s = p.T - 1; s >= 0; --s) {
    const int t = p.rev ? p.T - 1 - s : s, tp = p.rev ? t + 1 : t - 1;
    bf16* xw = p.xch + (size_t)(s & 1) * p.B * K;
#pragma unroll
    for (int k = 0; k < LSTM_ITEMS; ++k) {
      if (k < n_items) {
        const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
        const int j = j0 + jj, b = b0 + lb;
        if (lb < bs && b < p.B) {
          float dh = dhc[k];
          if (p.dh_ext) dh += p.dh_ext[(long long)t * p.dh_ext_ts + (long long)b * p.dh_ext_ld + j];
          const long long o = ((long long)t * p.B + b) * H + j;
          const float gi_ = p.si[o], gf = p.sf[o], gg = p.sg[o], go = p.so[o];
          const float tc = fast_tanh(p.sc[o]);
          const float cp = s > 0 ? p.sc[((long long)tp * p.B + b) * H + j] : 0.f;
          const float dc = dcc[k] + dh * go * (1.f - tc * tc);
          float d[4];
          d[0] = dc * gg * gi_ * (1.f - gi_);
          d[1] = dc * cp * gf * (1.f - gf);
          d[2] = dc * gi_ * (1.f - gg * gg);
          d[3] = dh * tc * go * (1.f - go);
          dcc[k] = dc * gf;
          float* da = p.da + (long long)t * p.da_ts + (long long)b * p.da_ld + j;
          bf16* x = xw + (long long)b * K + j;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            da[q * H] = d[q];
            if (s > 0) x[q * H] = __float2bfloat16_rn(d[q]);
          }
          dhc[k] = 0.f;
        }
      }
    }
    if (s > 0) {
      group_arrive(ctr);
      ++arrivals;
      // the group's gate gradients come back 16 videos at a time (64 KB at H = 512): one pass for the 16-video groups,
      // two for the 32-video groups of the paired-direction launch (W_hh^T slice + 32 videos would not fit shared memory)
      for (int hb = 0; hb < bs / LSTM_XB; ++hb) {
      if (TMA) {      // one thread fetches the 64 KB with bulk-tensor copies (persist.cuh)
        if (tid == 0) {
          if (hb == 0) spin_until(ctr, (unsigned)p.C * arrivals);
          tma_fetch_operand(sX, LSTM_XB, 0, &tmX, bar_x, 0, KB, b0 + hb * LSTM_XB, s & 1);
        }
        mbar_wait(bar_x, phase_x);
        phase_x ^= 1;
      } else {
      if (hb == 0) group_wait(ctr, (unsigned)p.C * arrivals);
      load_operand_rows_async(sX, LSTM_XB, 0, xw, K, b0 + hb * LSTM_XB, LSTM_XB, p.B, K);
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      }
      // dh carry [u, 16] = W_hh^T slice [u, 4H] x da^T: mma.sync m16n8k16, k-steps dealt over the 8 warps
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
        load_a_frag(aW, u, 16, ks << 4, a1);
        load_b_frag2(aX, LSTM_XB, 0, ks << 4, bq);
        mma_bf16_16816(acc[0][0], a0, bq[0], bq[1]);
        mma_bf16_16816(acc[0][1], a0, bq[2], bq[3]);
        mma_bf16_16816(acc[1][0], a1, bq[0], bq[1]);
        mma_bf16_16816(acc[1][1], a1, bq[2], bq[3]);
      }
      {
        float* r = sR + (size_t)warp * u * bs + hb * LSTM_XB;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const int row = mt * 16 + gid, col = nt * 8 + 2 * tig;
            r[row * bs + col] = acc[mt][nt][0]; r[row * bs + col + 1] = acc[mt][nt][1];
            r[(row + 8) * bs + col] = acc[mt][nt][2]; r[(row + 8) * bs + col + 1] = acc[mt][nt][3];
          }
      }
      if (hb + 1 < bs / LSTM_XB) __syncthreads();       // every warp is done with sX before the next pass refills it
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < LSTM_ITEMS; ++k) {
        if (k < n_items) {
          const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
          if (lb < bs && b0 + lb < p.B) {
            float s = 0.f;
            for (int w = 0; w < nwarps; ++w) s += sR[(size_t)w * u * bs + jj * bs + lb];
            dhc[k] = s;
          }
        }
      }
      __syncthreads();
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------
static int lstm_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

struct LstmPlan { int bs, C, u, G; size_t smem_f, smem_b; };
// wide: groups of 32 videos instead of 16 -- half the CTAs per direction, so that the two directions of a bidirectional
// LSTM can run side by side (lstm_persist_pair_ok)
static bool plan_lstm(int B, int H, LstmPlan& pl, bool wide = false) {
  if (H % 64 != 0 || H % 32 != 0) return false;
  pl.u = 32; pl.C = H / 32; pl.bs = wide ? 32 : 16; pl.G = (B + pl.bs - 1) / pl.bs;
  if ((long long)pl.G * pl.C > lstm_num_sms()) return false;
  if (pl.u * pl.bs > LSTM_ITEMS * PERSIST_THREADS) return false;
  const size_t KBf = H / 64, KBb = 4 * H / 64;
  pl.smem_f = KBf * 128 * 128 + KBf * pl.bs * 128 + (size_t)pl.bs * 129 * 4 + 64 + 1024;
  pl.smem_b = KBb * pl.u * 128 + KBb * LSTM_XB * 128 + (size_t)(PERSIST_THREADS / 32) * pl.u * pl.bs * 4 + 64 + 1024;
  return pl.smem_f <= 227 * 1024 && pl.smem_b <= 227 * 1024;
}

bool lstm_persist_eligible(int B, int H, int nsplit, int Hp) {
  LstmPlan pl;
  static const bool off = getenv("PVCR_NO_PERSIST_LSTM") != nullptr;
  return !off && nsplit == 1 && Hp == H && plan_lstm(B, H, pl);
}
// both directions co-resident: 2 x (groups of 32 videos x H/32 CTAs) fit the SMs
bool lstm_persist_pair_ok(int B, int H, int nsplit, int Hp) {
  LstmPlan pl;
  static const bool off = getenv("PVCR_NO_LSTM_PAIR") != nullptr;      // A/B knob
  return !off && lstm_persist_eligible(B, H, nsplit, Hp) && plan_lstm(B, H, pl, true) && 2 * pl.G * pl.C <= lstm_num_sms();
}

static int lstm_coop(const void* kern, int grid, size_t smem, void** args, cudaStream_t st, int cls, const char* what) {
  PVCR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  PVCR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PERSIST_THREADS, smem));
  PVCR_REQUIRE(per_sm * lstm_num_sms() >= grid, "%s: %d CTAs cannot be co-resident", what, grid);
  LaunchScope ls_(cls, st);
  PVCR_CUDA_CHECK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(PERSIST_THREADS), args, smem, st));
  return PVCR_OK;
}

int lstm_persist_fwd(const LstmSeqArgs& s, cudaStream_t st, bool wide) {
  LstmPlan pl;
  PVCR_REQUIRE(plan_lstm(s.B, s.H, pl, wide), "lstm_persist_fwd: shape B=%d H=%d not supported", s.B, s.H);
  LstmPersistFwd p{};
  p.T = s.T; p.B = s.B; p.H = s.H; p.bs = pl.bs; p.C = pl.C; p.u = pl.u; p.rev = s.rev;
  p.whh = s.whh.ptr; p.whh_ld = s.whh.ld; p.b_hh = s.b_hh;
  p.gi = s.gi; p.gi_ts = s.gi_ts; p.gi_ld = s.gi_ld;
  p.h = s.h; p.h_ts = s.h_ts; p.h_ld = s.h_ld;
  p.hp = s.hp; p.hp_ts = s.hp_ts; p.hp_ld = s.hp_ld;
  p.si = s.si; p.sf = s.sf; p.sg = s.sg; p.so = s.so; p.sc = s.sc;
  p.counters = s.sync;
  PVCR_TRY(fill_zero(s.sync, sizeof(unsigned) * 32 * pl.G, st));
  void* args[] = {&p};
  return lstm_coop((const void*)lstm_persist_fwd_kernel, pl.G * pl.C, pl.smem_f, args, st, KC_GRU_FWD, "lstm_persist_fwd");
}

int lstm_persist_bwd(const LstmSeqArgs& s, const Planes& whhT, const float* dh_ext, long long dh_ext_ts,
                     long long dh_ext_ld, float* da, long long da_ts, long long da_ld, bf16* xch, cudaStream_t st, bool wide) {
  LstmPlan pl;
  PVCR_REQUIRE(plan_lstm(s.B, s.H, pl, wide), "lstm_persist_bwd: shape B=%d H=%d not supported", s.B, s.H);
  LstmPersistBwd p{};
  p.T = s.T; p.B = s.B; p.H = s.H; p.bs = pl.bs; p.C = pl.C; p.u = pl.u; p.rev = s.rev;
  p.whhT = whhT.ptr; p.whhT_ld = whhT.ld;
  p.dh_ext = dh_ext; p.dh_ext_ts = dh_ext_ts; p.dh_ext_ld = dh_ext_ld;
  p.si = s.si; p.sf = s.sf; p.sg = s.sg; p.so = s.so; p.sc = s.sc;
  p.da = da; p.da_ts = da_ts; p.da_ld = da_ld; p.xch = xch;
  p.counters = s.sync;
  PVCR_TRY(fill_zero(s.sync, sizeof(unsigned) * 32 * pl.G, st));
  // measured (cfg3, 64 KB of gate gradients per step): the bulk-tensor fetch is 4 % SLOWER than cp.async in this kernel
  // (0.295 vs 0.284 ms per direction), unlike the GRU / decoder backward sweeps -- off unless asked for
  static const bool no_tma = getenv("PVCR_LSTM_BWD_TMA") == nullptr || getenv("PVCR_NO_TMA_XCHG") != nullptr;
  CUtensorMap tmX;         // exchange buffer [2][B][4H] as (k, video, parity)
  PVCR_TRY(make_tensor_map(&tmX, OperandView{p.xch, (long long)4 * p.H, (long long)p.B * 4 * p.H, p.B, 2}, 4 * p.H, LSTM_XB));
  void* args[] = {&p, &tmX};
  const void* kern = no_tma ? (const void*)lstm_persist_bwd_kernel<false> : (const void*)lstm_persist_bwd_kernel<true>;
  return lstm_coop(kern, pl.G * pl.C, pl.smem_b, args, st, KC_GRU_BWD, "lstm_persist_bwd");
}

}  // namespace pvcr
