// bf16 x bf16 -> fp32 GEMM for sm_100a:  C[z][m][n] = sum_k A[z][m][k] * B[z][n][k]   ("TN", both K-major)
//
//   * operands are staged global -> shared by TMA (cp.async.bulk.tensor, 128-byte swizzle) through a
//     STAGES-deep mbarrier ring,
//   * one elected thread issues tcgen05.mma (UMMA 128 x BN x 16) with the accumulator in TMEM,
//   * four epilogue warps read the accumulator back with tcgen05.ld (one thread == one output row)
//     and hand 32-column chunks to an epilogue functor (plain store, fused cross-entropy, ...).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue
// (warp w may only touch TMEM lanes 32*(w%4) .. 32*(w%4)+31).
#pragma once
#include "common.cuh"

namespace pvcr {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;   // 64 bf16 = 128 bytes = one swizzle row
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_PERSIST_THREADS = 320;   // TMA warp + MMA warp + 8 epilogue warps (two per TMEM lane quadrant)

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024;   // + alignment slack
};

// TMEM allocations are powers of two >= 32 columns
__host__ __device__ constexpr int tmem_cols_pow2(int c) { return c <= 32 ? 32 : (c <= 64 ? 64 : (c <= 128 ? 128 : (c <= 256 ? 256 : 512))); }

struct GemmCoords {
  int M, N, K;          // K = concatenated (all split planes), multiple of GEMM_BK
  int a_z0, a_zmul;     // slab coordinate of A for grid z:  a_z0 + z * a_zmul
  int b_z0, b_zmul;
  int k_splits;         // > 1: grid z enumerates K ranges instead of slabs (epilogue must accumulate atomically)
  // compact B planes (common.cuh split_term role 2): K-block at virtual column k = p * b_kp + off reads the stored
  // column  term(p) * b_kp + off,  term(p) = (b_terms >> 2p) & 3.  b_kp == 0: B is stored as it is multiplied.
  int b_kp = 0;
  unsigned b_terms = 0;
  // shifted-row taps (persistent kernel, K-major A): the K range is a_taps blocks of a_tap_kb k-blocks; block s reads the A
  // columns of ONE tap-width matrix at rows m0 + a_tap_off[s] -- a 3 x 3 convolution over a flat zero-padded channels-last
  // layout is nine row-shifted products accumulated in the SAME TMEM tile (conv.cu), no im2col and no fp32 accumulation
  // through global memory between the taps
  int a_taps = 0, a_tap_kb = 0;
  int a_tap_off[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  // b_z_koff (persistent kernel, MN-major B, batched over z <= 9): slab z reads B's k rows at kb * 64 + a_tap_off[z] -- the nine
  // weight-gradient products d W_s = dY^T X[. + off_s] of a convolution as ONE batched launch (9 x the tiles of one tap)
  int b_z_koff = 0;
  int b_blocked = 0;     // split3 kernel: B is K-blocked ([term * b_blocked + chunk][row][64], b_blocked = chunks per term)
  int b_prefetch = 0;    // split3 kernel: request every B block of a CTA's tiles into L2 up front (B streamed from HBM)
  int b_evict_last = 0;  // persistent / split3 kernel, K-major B: L2 hint of the B loads (1 = evict_last: re-read by the next
                         // launch and small enough to stay; 2 = evict_first: streamed once, must not displace resident data)
  __host__ __device__ int b_col(int k) const {
    if (b_kp == 0) return k;
    const int p = k / b_kp;
    return (int)((b_terms >> (2 * p)) & 3u) * b_kp + (k - p * b_kp);
  }
};

#ifdef __CUDACC__

template <int BN, int STAGES, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, (BN <= 128 ? 2 : 1))
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmCoords gc,
               Epi epi) {
  using SM = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * GEMM_BM;
  const bool split = gc.k_splits > 1;
  const int z = split ? 0 : blockIdx.z;
  const int total_kb = gc.K / GEMM_BK;
  const int kb_per = split ? (total_kb + gc.k_splits - 1) / gc.k_splits : total_kb;
  const int kb0 = split ? blockIdx.z * kb_per : 0;
  const int num_kb = min(kb_per, total_kb - kb0);      // >= 1 by construction of k_splits on the host

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int az = gc.a_z0 + z * gc.a_zmul, bz = gc.b_z0 + z * gc.b_zmul;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], SM::STAGE_BYTES);
        uint8_t* sa = smem + s * SM::STAGE_BYTES;
        tma_load_3d(sa, &tmA, &full_bar[s], (kb0 + kb) * GEMM_BK, m0, az);
        tma_load_3d(sa + SM::A_BYTES, &tmB, &full_bar[s], gc.b_col((kb0 + kb) * GEMM_BK), n0, bz);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * SM::STAGE_BYTES);
        const uint64_t da = umma_desc_k128(sa);
        const uint64_t db = umma_desc_k128(sa + SM::A_BYTES);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          // advance 16 elements (32 bytes) along K inside the 128-byte swizzle row: +2 in the >>4 address field
          umma_bf16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[s]);     // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full_bar);       // accumulator complete
    }
  } else {
    const int q = warp & 3;             // TMEM lane quadrant of this warp
    const int row = m0 + q * 32 + lane;
    Epi e = epi;                        // per-thread copy: functors may keep running state across chunks
    e.begin(row, split ? (int)blockIdx.z : z);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    float v[32];
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      if (n0 + c >= gc.N) break;        // warp-uniform
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      e.chunk(row, n0 + c, z, v);
    }
    e.end(row, blockIdx.x, z);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// Persistent variant: one CTA per SM loops over output tiles (m fastest, so CTAs running side by side share the
// B tile in L2); the TMA producer runs ahead across tile boundaries through the same shared-memory ring, the
// accumulator is double-buffered in TMEM (2 x BN columns), and the four epilogue warps drain tile i while the
// MMA warp already accumulates tile i+1.
// A_MN / B_MN: that operand is MN-major (a [k][mn] row-major matrix, i.e. it enters the product transposed),
// loaded as 64 x 64 TMA boxes.  A_MN && B_MN: C = A^T B (weight gradients); !A_MN && B_MN: C = A B (data gradients
// dX = dY W with the forward weight planes, no transposed weight copies).
// EW epilogue warps (8 or 16: EW / 4 per TMEM lane quadrant, each draining BN / (EW / 4) columns): epilogues that do real
// arithmetic per element (fused cross entropy) are issue-latency bound with two warps per SM sub-partition.
template <int BN, int STAGES, class Epi, bool A_MN = false, bool B_MN = false, int EW = 8>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_tn_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          GemmCoords gc, int tiles_m, int tiles_n, int num_tiles, Epi epi) {
  using SM = GemmSmem<BN, STAGES>;
  static_assert(2 * BN <= 512, "two accumulators must fit the 512 TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint8_t* epi_smem = smem + SM::BAR_OFFSET + 256;   // Epi::SMEM_PER_WARP bytes of staging per epilogue warp

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int splits = gc.k_splits > 1 ? gc.k_splits : 1;       // work item = (tile, K range); num_tiles counts items
  const int total_kb = gc.K / GEMM_BK;
  const int kb_per = (total_kb + splits - 1) / splits;
  const int per_z = tiles_m * tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], EW);             // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols_pow2(2 * BN));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t b_policy = gc.b_evict_last == 2 ? l2_policy_evict_first() : l2_policy_evict_last();
      int it = 0;                                    // running k-block counter across tiles
      for (int w = blockIdx.x; w < num_tiles; w += gridDim.x) {
        const int t = w / splits, kb0 = (w - t * splits) * kb_per;
        const int num_kb = min(kb_per, total_kb - kb0);
        const int z = t / per_z, r = t - z * per_z;
        const int m0 = (r % tiles_m) * GEMM_BM, n0 = (r / tiles_m) * BN;
        const int az = gc.a_z0 + z * gc.a_zmul, bz = gc.b_z0 + z * gc.b_zmul;
        for (int kq = 0; kq < num_kb; ++kq, ++it) {
          const int kb = kb0 + kq;
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], SM::STAGE_BYTES);
          uint8_t* sa = smem + s * SM::STAGE_BYTES;
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < GEMM_BM / 64; ++i) tma_load_3d(sa + i * 8192, &tmA, &full_bar[s], m0 + 64 * i, kb * GEMM_BK, az);
          } else if (gc.a_taps) {
            const int tap = kb / gc.a_tap_kb;
            tma_load_3d(sa, &tmA, &full_bar[s], (kb - tap * gc.a_tap_kb) * GEMM_BK, m0 + gc.a_tap_off[tap], az);
          } else {
            tma_load_3d(sa, &tmA, &full_bar[s], kb * GEMM_BK, m0, az);
          }
          if (B_MN) {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_3d(sa + SM::A_BYTES + i * 8192, &tmB, &full_bar[s], n0 + 64 * i,
                          kb * GEMM_BK + (gc.b_z_koff ? gc.a_tap_off[z] : 0), bz);
          } else if (gc.b_evict_last) {
            tma_load_3d_hint(sa + SM::A_BYTES, &tmB, &full_bar[s], gc.b_col(kb * GEMM_BK), n0, bz, b_policy);
          } else {
            tma_load_3d(sa + SM::A_BYTES, &tmB, &full_bar[s], gc.b_col(kb * GEMM_BK), n0, bz);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int it = 0, ti = 0;
      for (int w = blockIdx.x; w < num_tiles; w += gridDim.x, ++ti) {
        const int kb0 = (w % splits) * kb_per;
        const int num_kb = min(kb_per, total_kb - kb0);
        const int acc = ti & 1;
        mbar_wait(&tmem_empty_bar[acc], ((ti >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * SM::STAGE_BYTES);
          // MN-major: 64-element mn blocks 8192 B apart (one 64 x 64 box each), 8-row k groups 1024 B apart, 16 k rows
          // per instruction = 2048 B;  K-major: 16 k elements = 32 B inside the 128-byte swizzle row
          const uint64_t da = A_MN ? umma_desc_mn128(sa, 8192, 1024) : umma_desc_k128(sa);
          const uint64_t db = B_MN ? umma_desc_mn128(sa + SM::A_BYTES, 8192, 1024) : umma_desc_k128(sa + SM::A_BYTES);
          constexpr uint64_t ka = A_MN ? 128 : 2, kbs = B_MN ? 128 : 2;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_bf16(tmem_d, da + (uint64_t)k * ka, db + (uint64_t)k * kbs, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    // warp w drains TMEM lanes 32*(w%4).. of the column half (w-2)/4: two warps per SM sub-partition hide each
    // other's dependency stalls in the epilogue arithmetic
    const int q = warp & 3, half = (warp - 2) >> 2;
    constexpr int PARTS = EW / 4, HALF = BN / PARTS;
    int ti = 0;
    for (int w = blockIdx.x; w < num_tiles; w += gridDim.x, ++ti) {
      const int t = w / splits;
      const int z = t / per_z, r = t - z * per_z;
      const int m0 = (r % tiles_m) * GEMM_BM, n0 = (r / tiles_m) * BN;
      const int acc = ti & 1;
      const int row = m0 + q * 32 + lane;
      Epi e = epi;
      if constexpr (Epi::SMEM_PER_WARP > 0) e.attach(epi_smem + (warp - 2) * Epi::SMEM_PER_WARP);
      e.begin(row, splits > 1 ? w - t * splits : z);
      mbar_wait(&tmem_full_bar[acc], (ti >> 1) & 1);
      tc_fence_after();
      float v[32];
#pragma unroll 1
      for (int c = half * HALF; c < (half + 1) * HALF; c += 32) {
        if (n0 + c >= gc.N) break;
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c), v);
        e.chunk(row, n0 + c, z, v);
      }
      e.end(row, (r / tiles_m) * PARTS + half, z);  // part index: (N tile, column part)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols_pow2(2 * BN));
  }
}


// ---- split-precision (bf16x3) product with every operand block loaded ONCE per K chunk ---------------------------------
// C = sum over the six term pairs (a2,b0) (a1,b1) (a0,b2) (a1,b0) (a0,b1) (a0,b0) of A_term B_term^T.  The generic kernels
// above walk the six virtual planes one after the other, so every stored term of both operands crosses L2 -> SM two to three
// times (48 KB per 128 x 256 x 64 product block).  Here a pipeline stage holds the three A terms and the three B terms of one
// 64-column K chunk (3 x 16 KB + 3 x BN x 128 B) and the MMA warp issues the six products from it: half the operand traffic
// for the skinny decoding products (M = batch <= 128 rows: the vocabulary projection streams W_v's 70 MB once per step and is
// bound by exactly that).  A: split planes in A-role order (term t at plane 2 - t, common.cuh), a_kp columns per plane;
// B: compact planes (term t at plane t), b_kp columns per plane; gc.K = columns of ONE plane (multiple of 64).
// BK = 64 (128-byte swizzle rows) or 32 (64-byte swizzle rows: half-size stages, twice as many of them in flight)
template <int BN, int BK = GEMM_BK>
struct GemmSplit3Smem {
  static_assert(BK == 64 || BK == 32, "K chunk");
  static constexpr int A_BYTES = GEMM_BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 3 * A_BYTES + 3 * B_BYTES;
  static constexpr int STAGES = (227 * 1024 - 2048) / STAGE_BYTES > 6 ? 6 : (227 * 1024 - 2048) / STAGE_BYTES;
  static_assert(STAGES >= 2, "tile too wide for two pipeline stages");
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;          // barriers / slot + alignment slack
};

template <int BN, class Epi, int EW, int BK = GEMM_BK>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_split3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmCoords gc,
                   int a_kp, int b_kp, int tiles_m, int tiles_n, int num_tiles, Epi epi) {
  using SM = GemmSplit3Smem<BN, BK>;
  constexpr int STAGES = SM::STAGES;
  static_assert(2 * BN <= 512 && BN % 16 == 0 && (BN * 128) % 1024 == 0, "tile width");
  static_assert(BN % (EW / 4) == 0 && (BN / (EW / 4)) % 32 == 0, "each epilogue warp drains whole 32-column chunks");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int chunks = gc.K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], EW);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols_pow2(2 * BN));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // launched programmatically (decode loop): everything above overlapped the predecessor's tail; its output (the A planes)
  // may be read, and C written, only from here on
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t b_policy = gc.b_evict_last == 2 ? l2_policy_evict_first() : l2_policy_evict_last();
      int it = 0;
      // B streamed from HBM (b_prefetch): the bytes in flight of the shared-memory ring (STAGES x 3 x B_BYTES per SM) are far
      // below bandwidth x latency, so every B block of this CTA's tiles is first requested into L2 -- the ring then runs at L2
      // latency while DRAM streams at its own pace.  (B does not depend on the predecessor: this could even precede pdl_wait.)
      if (gc.b_prefetch && tiles_m == 1) {                // (several row tiles share a B tile: nothing to stream ahead)
        for (int w = blockIdx.x; w < num_tiles; w += gridDim.x) {
          const int n0 = w * BN;
          for (int c = 0; c < chunks; ++c)
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              const int bk = gc.b_blocked ? (c * BK) % GEMM_BK : t * b_kp + c * BK;
              const int bz = gc.b_blocked ? t * gc.b_blocked + (c * BK) / GEMM_BK : gc.b_z0;
              tma_prefetch_l2_3d(&tmB, bk, n0, bz);
            }
        }
      }
      for (int w = blockIdx.x; w < num_tiles; w += gridDim.x) {
        const int m0 = (w % tiles_m) * GEMM_BM, n0 = (w / tiles_m) * BN;
        for (int c = 0; c < chunks; ++c, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], SM::STAGE_BYTES);
          uint8_t* sa = smem + s * SM::STAGE_BYTES;
          uint8_t* sb = sa + 3 * SM::A_BYTES;
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            tma_load_3d(sa + t * SM::A_BYTES, &tmA, &full_bar[s], (2 - t) * a_kp + c * BK, m0, gc.a_z0);
            const int bk = gc.b_blocked ? (c * BK) % GEMM_BK : t * b_kp + c * BK;
            const int bz = gc.b_blocked ? t * gc.b_blocked + (c * BK) / GEMM_BK : gc.b_z0;
            if (gc.b_evict_last) tma_load_3d_hint(sb + t * SM::B_BYTES, &tmB, &full_bar[s], bk, n0, bz, b_policy);
            else tma_load_3d(sb + t * SM::B_BYTES, &tmB, &full_bar[s], bk, n0, bz);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
      int it = 0, ti = 0;
      for (int w = blockIdx.x; w < num_tiles; w += gridDim.x, ++ti) {
        const int acc = ti & 1;
        mbar_wait(&tmem_empty_bar[acc], ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int c = 0; c < chunks; ++c, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * SM::STAGE_BYTES);
          const uint32_t sb = sa + 3 * SM::A_BYTES;
          // smallest products first (the order of the six virtual planes of the generic kernels)
          constexpr int TA[6] = {2, 1, 0, 1, 0, 0}, TB[6] = {0, 1, 2, 0, 1, 0};
#pragma unroll
          for (int p = 0; p < 6; ++p) {
            const uint64_t da = BK == 64 ? umma_desc_k128(sa + TA[p] * SM::A_BYTES) : umma_desc_k64(sa + TA[p] * SM::A_BYTES);
            const uint64_t db = BK == 64 ? umma_desc_k128(sb + TB[p] * SM::B_BYTES) : umma_desc_k64(sb + TB[p] * SM::B_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (c | p | k) != 0);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    constexpr int PARTS = EW / 4, HALF = BN / PARTS;
    int ti = 0;
    for (int w = blockIdx.x; w < num_tiles; w += gridDim.x, ++ti) {
      const int m0 = (w % tiles_m) * GEMM_BM, n0 = (w / tiles_m) * BN;
      const int acc = ti & 1;
      const int row = m0 + q * 32 + lane;
      Epi e = epi;
      e.begin(row, 0);
      mbar_wait(&tmem_full_bar[acc], (ti >> 1) & 1);
      tc_fence_after();
      float v[32];
#pragma unroll 1
      for (int c = half * HALF; c < (half + 1) * HALF; c += 32) {
        if (n0 + c >= gc.N) break;
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c), v);
        e.chunk(row, n0 + c, 0, v);
      }
      e.end(row, (w / tiles_m) * PARTS + half, 0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols_pow2(2 * BN));
  }
}

// ---- 2-CTA cluster variant: the B tile is shared by TMA multicast ------------------------------------------------
// With K = 512 (vocabulary projection) a 128 x 256 tile needs 384 KB of operands for 4096 tensor-pipe cycles: 148 CTAs
// ask L2 for ~10.5 TB/s, at the chip's L2 -> SM limit, and the tensor pipe idles half of the time.  Here two CTAs of a
// cluster work on the two 128-row tiles (m0, m0 + 128) of the same 256 columns: each loads its own A tile and HALF of
// the B tile, multicast to both (cp.async.bulk.tensor ... .multicast::cluster signals the full barrier at the same
// offset in both CTAs), so a tile costs 256 KB of L2 reads instead of 384 KB.  A stage may be refilled only when BOTH
// CTAs' MMAs have read it: the empty barriers count two arrivals and every tcgen05.commit on them is multicast.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}

template <int BN, int STAGES, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_PERSIST_THREADS, 1)
gemm_tn_mc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh, GemmCoords gc,
                   int pairs_m, int tiles_n, int num_items, Epi epi) {
  using SM = GemmSmem<BN, STAGES>;
  static_assert(2 * BN <= 512, "two accumulators must fit the 512 TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint8_t* epi_smem = smem + SM::BAR_OFFSET + 256;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int total_kb = gc.K / GEMM_BK;
  constexpr int HALF_ROWS = BN / 2;                  // rows of the B tile this CTA loads (for both)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmBh);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 2);                   // this CTA's and the peer's MMA have read the stage
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int w = cluster_id; w < num_items; w += num_clusters) {
        const int m0 = (2 * (w % pairs_m) + rank) * GEMM_BM, n0 = (w / pairs_m) * BN;
        for (int kb = 0; kb < total_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], SM::STAGE_BYTES);
          uint8_t* sa = smem + s * SM::STAGE_BYTES;
          tma_load_3d(sa, &tmA, &full_bar[s], kb * GEMM_BK, m0, gc.a_z0);
          tma_load_3d_mc(sa + SM::A_BYTES + rank * HALF_ROWS * 128, &tmBh, &full_bar[s], kb * GEMM_BK,
                         n0 + rank * HALF_ROWS, gc.b_z0, (uint16_t)3);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
      int it = 0, ti = 0;
      for (int w = cluster_id; w < num_items; w += num_clusters, ++ti) {
        const int acc = ti & 1;
        mbar_wait(&tmem_empty_bar[acc], ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < total_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * SM::STAGE_BYTES);
          const uint64_t da = umma_desc_k128(sa), db = umma_desc_k128(sa + SM::A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_bf16(tmem_d, da + (uint64_t)k * 2, db + (uint64_t)k * 2, idesc, (kb | k) != 0);
          umma_commit_mc(&empty_bar[s], (uint16_t)3);
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    constexpr int HALF = BN / 2;
    int ti = 0;
    for (int w = cluster_id; w < num_items; w += num_clusters, ++ti) {
      const int nt = w / pairs_m;
      const int m0 = (2 * (w % pairs_m) + rank) * GEMM_BM, n0 = nt * BN;
      const int acc = ti & 1;
      const int row = m0 + q * 32 + lane;
      Epi e = epi;
      if constexpr (Epi::SMEM_PER_WARP > 0) e.attach(epi_smem + (warp - 2) * Epi::SMEM_PER_WARP);
      e.begin(row, 0);
      mbar_wait(&tmem_full_bar[acc], (ti >> 1) & 1);
      tc_fence_after();
      float v[32];
#pragma unroll 1
      for (int c = half * HALF; c < (half + 1) * HALF; c += 32) {
        if (n0 + c >= gc.N) break;
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c), v);
        e.chunk(row, n0 + c, 0, v);
      }
      e.end(row, nt * 2 + half, 0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // no CTA leaves while the peer may still signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ---- CTA-pair variant (tcgen05 cta_group::2): one 256 x BN UMMA tile per pair ---------------------------------------
// Each CTA of the pair stages its own 128 rows of A and its own BN/2 rows of B (32 KB per k-block instead of 48 KB),
// the leader CTA issues tcgen05.mma.cta_group::2 (M = 256) which reads both CTAs' shared memory, and each CTA ends up
// with the 128 x BN accumulator of its rows in its own TMEM.  One SM then takes in 64 B per tensor-pipe cycle at
// K-block granularity instead of 96 B: the K = 512 vocabulary GEMMs were limited by exactly that.
//   full[s]      (leader)   2 producer arrivals (.expect_tx, the peer's is remote) + both CTAs' TMA bytes
//   empty[s]     (each CTA) the leader's tcgen05.commit, multicast to both
//   tmem_full[a] (each CTA) the leader's tcgen05.commit, multicast to both
//   tmem_empty[a](leader)   16 arrivals: the 8 epilogue warps of both CTAs (the peer's are remote)
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}

template <int BN, int STAGES>
struct Gemm2SmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;          // this CTA's 128 rows of A
  static constexpr int B_BYTES = (BN / 2) * GEMM_BK * 2;         // this CTA's BN/2 rows of B
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;          // barriers / slot / epilogue offset + alignment slack
};

template <int BN, int STAGES, class Epi, int EW = 8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
gemm_tn_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh, GemmCoords gc,
                   int pairs_m, int tiles_n, int num_items, Epi epi) {
  using SM = Gemm2SmSmem<BN, STAGES>;
  static_assert(2 * BN <= 512, "two accumulators must fit the 512 TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint8_t* epi_smem = smem + SM::BAR_OFFSET + 512;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int total_kb = gc.K / GEMM_BK;
  constexpr int HALF_ROWS = BN / 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmBh);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);                    // used in the leader only
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 2 * EW);         // used in the leader only
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(2 * BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int w = cluster_id; w < num_items; w += num_clusters) {
        const int m0 = (2 * (w % pairs_m) + rank) * GEMM_BM, n0 = (w / pairs_m) * BN + rank * HALF_ROWS;
        for (int kb = 0; kb < total_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t lbar = mapa_rank(smem_u32(&full_bar[s]), 0);
          mbar_arrive_expect_tx_cluster(lbar, SM::STAGE_BYTES);
          uint8_t* sa = smem + s * SM::STAGE_BYTES;
          tma_load_3d_2sm(sa, &tmA, lbar, kb * GEMM_BK, m0, gc.a_z0);
          tma_load_3d_2sm(sa + SM::A_BYTES, &tmBh, lbar, kb * GEMM_BK, n0, gc.b_z0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * GEMM_BM, BN);
      int it = 0, ti = 0;
      for (int w = cluster_id; w < num_items; w += num_clusters, ++ti) {
        const int acc = ti & 1;
        mbar_wait(&tmem_empty_bar[acc], ((ti >> 1) & 1) ^ 1);      // both CTAs' epilogues drained it
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < total_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * SM::STAGE_BYTES);
          const uint64_t da = umma_desc_k128(sa), db = umma_desc_k128(sa + SM::A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_bf16_2sm(tmem_d, da + (uint64_t)k * 2, db + (uint64_t)k * 2, idesc, (kb | k) != 0);
          umma_commit_2sm(&empty_bar[s], (uint16_t)3);
        }
        umma_commit_2sm(&tmem_full_bar[acc], (uint16_t)3);
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    constexpr int PARTS = EW / 4, HALF = BN / PARTS;
    int ti = 0;
    for (int w = cluster_id; w < num_items; w += num_clusters, ++ti) {
      const int nt = w / pairs_m;
      const int m0 = (2 * (w % pairs_m) + rank) * GEMM_BM, n0 = nt * BN;
      const int acc = ti & 1;
      const int row = m0 + q * 32 + lane;
      Epi e = epi;
      if constexpr (Epi::SMEM_PER_WARP > 0) e.attach(epi_smem + (warp - 2) * Epi::SMEM_PER_WARP);
      e.begin(row, 0);
      mbar_wait(&tmem_full_bar[acc], (ti >> 1) & 1);
      tc_fence_after();
      float v[32];
#pragma unroll 1
      for (int c = half * HALF; c < (half + 1) * HALF; c += 32) {
        if (n0 + c >= gc.N) break;
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c), v);
        e.chunk(row, n0 + c, 0, v);
      }
      e.end(row, nt * PARTS + half, 0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&tmem_empty_bar[acc]), 0));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)) : "memory");
  }
}

// Plain store epilogue: C = acc (+ bias[n]) (+ C).  fp32 output, arbitrary ldc.
struct EpiStore {
  static constexpr int SMEM_PER_WARP = 0;   // shared-memory staging the persistent kernels reserve per epilogue warp
  float* C;
  long long ldc, c_zstride;
  const float* bias;
  long long bias_zstride;
  int accumulate, M, N;
  int atomic;             // split-K: every K range adds its partial product with red.global.add (C pre-zeroed)
  int first;              // set in begin(): this CTA's K range is the first one (adds the bias)
  __device__ __forceinline__ void begin(int, int zsplit) { first = (zsplit == 0); }
  __device__ __forceinline__ void chunk(int row, int col0, int z, float (&v)[32]) {
    if (row >= M) return;
    if (atomic) {
      float* c = C + (long long)row * ldc + col0;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < N) atomicAdd(c + j, v[j] + ((bias && first) ? __ldg(bias + col0 + j) : 0.f));
      return;
    }
    float* c = C + (long long)z * c_zstride + (long long)row * ldc + col0;
    const float* b = bias ? bias + (long long)z * bias_zstride + col0 : nullptr;
    const bool vec = (col0 + 32 <= N) && ((reinterpret_cast<uintptr_t>(c) & 15) == 0);
    if (vec) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        if (b) { o.x += __ldg(b + j); o.y += __ldg(b + j + 1); o.z += __ldg(b + j + 2); o.w += __ldg(b + j + 3); }
        if (accumulate) {
          const float4 p = *reinterpret_cast<const float4*>(c + j);
          o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
        }
        *reinterpret_cast<float4*>(c + j) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (col0 + j < N) {
          float o = v[j] + (b ? __ldg(b + j) : 0.f);
          if (accumulate) o += c[j];
          c[j] = o;
        }
      }
    }
  }
  __device__ __forceinline__ void end(int, int, int) {}
};

#endif  // __CUDACC__

// ---- host side ------------------------------------------------------------------------------
// Operand view: bf16, element (z, r, k) at ptr[z*slab_stride + r*ld + k]; k in [0,K).
struct OperandView {
  const bf16* ptr;
  long long ld, slab_stride;
  int rows, slabs;
  int kp = 0, terms = 0;      // compact B-role planes: `terms` stored planes of kp columns each (0: plain)
  const bf16* blocked = nullptr;   // optional second copy in the K-blocked layout [terms * kp / 64][rows][64] (block_planes)
};

// where gemm_argmax left the per-row (max, index) partials of its column parts
struct ArgmaxParts { const float* pmax; const int* pidx; int nparts; };

// Upper bound on the CTAs of a persistent GEMM launched while a CtaCap is alive on this host thread (0 = all SMs):
// GEMMs forked onto the side lane next to a persistent recurrent sweep take only the SMs the sweep leaves free.
int gemm_cta_cap();
struct CtaCap {
  int prev;
  explicit CtaCap(int cap);
  ~CtaCap();
};

// box_k = 64: 128-byte swizzle rows (default);  32: 64-byte swizzle rows
int make_tensor_map(CUtensorMap* out, const OperandView& v, int K, int box_rows, int box_k = GEMM_BK);
// MN-major operand: bf16 matrix [k_rows, mn_cols] (row stride v.ld, v.rows = k_rows); boxes of 64 (mn) x 64 (k).
int make_tensor_map_mn(CUtensorMap* out, const OperandView& v, int mn_cols);

#ifdef __CUDACC__
template <int BN, int STAGES, class Epi>
int launch_gemm_tn(const OperandView& a, const OperandView& b, const GemmCoords& gc, int grid_z, const Epi& epi,
                   cudaStream_t stream) {
  using SM = GemmSmem<BN, STAGES>;
  PVCR_REQUIRE(gc.K > 0 && gc.K % GEMM_BK == 0, "gemm: K=%d must be a positive multiple of %d", gc.K, GEMM_BK);
  PVCR_REQUIRE(gc.M > 0 && gc.N > 0 && grid_z > 0, "gemm: empty problem M=%d N=%d z=%d", gc.M, gc.N, grid_z);
  CUtensorMap ta, tb;
  PVCR_TRY(make_tensor_map(&ta, a, gc.K, GEMM_BM));
  PVCR_TRY(make_tensor_map(&tb, b, b.kp ? b.kp * b.terms : gc.K, BN));
  PVCR_REQUIRE((b.kp != 0) == (gc.b_kp != 0), "gemm: compact B planes need GemmCoords::b_kp (and vice versa)");
  auto kern = gemm_tn_kernel<BN, STAGES, Epi>;
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    PVCR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    attr_set = true;
  }
  PVCR_REQUIRE(gc.k_splits <= 1 || grid_z == 1, "gemm: split-K and batched slabs are exclusive");
  dim3 grid(cdiv(gc.N, BN), cdiv(gc.M, GEMM_BM), gc.k_splits > 1 ? gc.k_splits : grid_z);
  {
    LaunchScope ls_(KC_GEMM, stream, 2.0 * gc.M * gc.N * (double)gc.K * grid_z);
    kern<<<grid, GEMM_THREADS, SM::TOTAL, stream>>>(ta, tb, gc, epi);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

// MN = true: C[M,N] = A^T B with A = a [K rows, M cols], B = b [K rows, N cols] (row-major bf16, a.rows = b.rows = K
// need not be padded: rows past the end read as zero through the tensor map).
template <int BN, int STAGES, class Epi, bool A_MN = false, bool B_MN = false, int EW = 8>
int launch_gemm_tn_persistent(const OperandView& a, const OperandView& b, const GemmCoords& gc, int grid_z,
                              const Epi& epi, cudaStream_t stream) {
  using SM = GemmSmem<BN, STAGES>;
  PVCR_REQUIRE(gc.K > 0 && gc.K % GEMM_BK == 0, "gemm: K=%d must be a positive multiple of %d", gc.K, GEMM_BK);
  PVCR_REQUIRE(gc.M > 0 && gc.N > 0 && grid_z > 0, "gemm: empty problem M=%d N=%d z=%d", gc.M, gc.N, grid_z);
  CUtensorMap ta, tb;
  if (A_MN) PVCR_TRY(make_tensor_map_mn(&ta, a, gc.M));
  else PVCR_TRY(make_tensor_map(&ta, a, gc.a_taps ? gc.a_tap_kb * GEMM_BK : gc.K, GEMM_BM));
  PVCR_REQUIRE(!gc.a_taps || (!A_MN && gc.k_splits <= 1 && gc.a_taps <= 9 && gc.K == gc.a_taps * gc.a_tap_kb * GEMM_BK),
               "gemm: shifted-row taps need a K-major A, no split-K and K = taps x tap width");
  if (B_MN) PVCR_TRY(make_tensor_map_mn(&tb, b, gc.N)); else PVCR_TRY(make_tensor_map(&tb, b, b.kp ? b.kp * b.terms : gc.K, BN));
  PVCR_REQUIRE((b.kp != 0) == (gc.b_kp != 0) && !(B_MN && b.kp), "gemm: compact B planes need GemmCoords::b_kp (K-major B only)");
  auto kern = gemm_tn_persistent_kernel<BN, STAGES, Epi, A_MN, B_MN, EW>;
  static bool attr_set = false;
  static int sms = 0;
  if (!attr_set) {
    PVCR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL + 64 + 256 + EW * Epi::SMEM_PER_WARP));
    int dev = 0;
    PVCR_CUDA_CHECK(cudaGetDevice(&dev));
    PVCR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  const int tiles_m = cdiv(gc.M, GEMM_BM), tiles_n = cdiv(gc.N, BN);
  PVCR_REQUIRE(gc.k_splits <= 1 || grid_z == 1, "gemm: split-K and batched slabs are exclusive");
  const long long num_tiles = (long long)tiles_m * tiles_n * grid_z * (gc.k_splits > 1 ? gc.k_splits : 1);
  int grid = (int)(num_tiles < sms ? num_tiles : sms);
  if (gemm_cta_cap() > 0 && grid > gemm_cta_cap()) grid = gemm_cta_cap();
  {
    LaunchScope ls_(KC_GEMM, stream, 2.0 * gc.M * gc.N * (double)gc.K * grid_z);
    kern<<<grid, 64 + 32 * EW, SM::TOTAL + 64 + 256 + EW * Epi::SMEM_PER_WARP, stream>>>(ta, tb, gc, tiles_m, tiles_n, (int)num_tiles, epi);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

// bf16x3 product from A-role split planes (6 planes of a_kp columns) and compact B planes (3 terms of b.kp columns):
// gc.K = columns of one plane.  Epilogues without per-warp shared memory only.
template <int BN, class Epi, int EW, int BK = GEMM_BK>
int launch_gemm_split3(const OperandView& a, const OperandView& b, GemmCoords gc, int a_kp, const Epi& epi, cudaStream_t stream) {
  using SM = GemmSplit3Smem<BN, BK>;
  static_assert(Epi::SMEM_PER_WARP == 0, "split3 kernel: register-only epilogues");
  PVCR_REQUIRE(gc.K > 0 && gc.K % GEMM_BK == 0 && gc.K <= a_kp && gc.K <= b.kp, "gemm (split3): K=%d vs planes of %d / %d columns", gc.K, a_kp, b.kp);
  PVCR_REQUIRE(gc.M > 0 && gc.N > 0 && b.terms == 3 && gc.k_splits <= 1, "gemm (split3): needs three compact B terms, no split-K");
  CUtensorMap ta, tb;
  PVCR_TRY(make_tensor_map(&ta, a, 6 * a_kp, GEMM_BM, BK));
  static const bool blocked_off = getenv("PVCR_NO_WV_BLOCKED") != nullptr;      // A/B knob
  if (b.blocked && !blocked_off) {
    PVCR_REQUIRE(b.kp % GEMM_BK == 0 && b.slabs == 1, "gemm (split3): blocked B planes need kp %% 64 == 0 and one slab");
    OperandView bb{b.blocked, GEMM_BK, (long long)b.rows * GEMM_BK, b.rows, b.terms * b.kp / GEMM_BK};
    PVCR_TRY(make_tensor_map(&tb, bb, GEMM_BK, BN, BK));
    gc.b_blocked = b.kp / GEMM_BK;
  } else {
    PVCR_TRY(make_tensor_map(&tb, b, b.kp * b.terms, BN, BK));
  }
  auto kern = gemm_split3_kernel<BN, Epi, EW, BK>;
  static bool attr_set = false;
  static int sms = 0;
  if (!attr_set) {
    PVCR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    int dev = 0;
    PVCR_CUDA_CHECK(cudaGetDevice(&dev));
    PVCR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  const int tiles_m = cdiv(gc.M, GEMM_BM), tiles_n = cdiv(gc.N, BN);
  const long long num_tiles = (long long)tiles_m * tiles_n;
  int grid = (int)(num_tiles < sms ? num_tiles : sms);
  if (gemm_cta_cap() > 0 && grid > gemm_cta_cap()) grid = gemm_cta_cap();
  {
    LaunchScope ls_(KC_GEMM, stream, 2.0 * gc.M * gc.N * (double)gc.K * 6);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 + 32 * EW); cfg.dynamicSmemBytes = SM::TOTAL; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    if (pdl_enabled()) { cfg.attrs = at; cfg.numAttrs = 1; }
    PVCR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, ta, tb, gc, a_kp, (int)b.kp, tiles_m, tiles_n, (int)num_tiles, epi));
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

// K-major x K-major product on the 2-CTA multicast kernel (single slab, no split-K).
template <int BN, int STAGES, class Epi>
int launch_gemm_tn_mc2(const OperandView& a, const OperandView& b, const GemmCoords& gc, const Epi& epi,
                       cudaStream_t stream) {
  using SM = GemmSmem<BN, STAGES>;
  PVCR_REQUIRE(gc.K > 0 && gc.K % GEMM_BK == 0, "gemm: K=%d must be a positive multiple of %d", gc.K, GEMM_BK);
  PVCR_REQUIRE(gc.M > 0 && gc.N > 0 && gc.k_splits <= 1, "gemm (multicast): empty problem or split-K");
  CUtensorMap ta, tb;
  PVCR_TRY(make_tensor_map(&ta, a, gc.K, GEMM_BM));
  PVCR_TRY(make_tensor_map(&tb, b, gc.K, BN / 2));
  auto kern = gemm_tn_mc2_kernel<BN, STAGES, Epi>;
  static bool attr_set = false;
  static int sms = 0;
  constexpr int SMEM = SM::TOTAL + 64 + 256 + 8 * Epi::SMEM_PER_WARP;
  if (!attr_set) {
    PVCR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    int dev = 0;
    PVCR_CUDA_CHECK(cudaGetDevice(&dev));
    PVCR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  const int pairs_m = cdiv(cdiv(gc.M, GEMM_BM), 2), tiles_n = cdiv(gc.N, BN);
  const long long num_items = (long long)pairs_m * tiles_n;
  int clusters = sms / 2;
  if (gemm_cta_cap() > 1 && clusters > gemm_cta_cap() / 2) clusters = gemm_cta_cap() / 2;
  if (num_items < clusters) clusters = (int)num_items;
  {
    LaunchScope ls_(KC_GEMM, stream, 2.0 * gc.M * gc.N * (double)gc.K);
    kern<<<2 * clusters, GEMM_PERSIST_THREADS, SMEM, stream>>>(ta, tb, gc, pairs_m, tiles_n, (int)num_items, epi);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

// K-major x K-major product on the CTA-pair (cta_group::2) kernel (single slab, no split-K).
template <int BN, int STAGES, class Epi, int EW = 8>
int launch_gemm_tn_2sm(const OperandView& a, const OperandView& b, const GemmCoords& gc, const Epi& epi,
                       cudaStream_t stream) {
  using SM = Gemm2SmSmem<BN, STAGES>;
  PVCR_REQUIRE(gc.K > 0 && gc.K % GEMM_BK == 0, "gemm: K=%d must be a positive multiple of %d", gc.K, GEMM_BK);
  PVCR_REQUIRE(gc.M > 0 && gc.N > 0 && gc.k_splits <= 1, "gemm (2-CTA): empty problem or split-K");
  CUtensorMap ta, tb;
  PVCR_TRY(make_tensor_map(&ta, a, gc.K, GEMM_BM));
  PVCR_TRY(make_tensor_map(&tb, b, gc.K, BN / 2));
  auto kern = gemm_tn_2sm_kernel<BN, STAGES, Epi, EW>;
  static bool attr_set = false;
  static int sms = 0;
  constexpr int SMEM = SM::TOTAL + EW * Epi::SMEM_PER_WARP;
  if (!attr_set) {
    PVCR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    int dev = 0;
    PVCR_CUDA_CHECK(cudaGetDevice(&dev));
    PVCR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  const int pairs_m = cdiv(cdiv(gc.M, GEMM_BM), 2), tiles_n = cdiv(gc.N, BN);
  const long long num_items = (long long)pairs_m * tiles_n;
  int clusters = sms / 2;
  if (gemm_cta_cap() > 1 && clusters > gemm_cta_cap() / 2) clusters = gemm_cta_cap() / 2;
  if (num_items < clusters) clusters = (int)num_items;
  {
    LaunchScope ls_(KC_GEMM, stream, 2.0 * gc.M * gc.N * (double)gc.K);
    kern<<<2 * clusters, 64 + 32 * EW, SMEM, stream>>>(ta, tb, gc, pairs_m, tiles_n, (int)num_items, epi);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}
#endif

}  // namespace pvcr
