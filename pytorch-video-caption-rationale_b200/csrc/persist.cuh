// Building blocks of the persistent recurrent kernels (sm_100a).
//
// A persistent kernel owns a slice of the recurrent weight matrices in shared memory for the whole sequence
// and loops over the timesteps inside one cooperative launch.  The batch is cut into independent groups of
// `bs` videos; a group is served by C CTAs, each owning the weight rows of `u` hidden units, so the only
// inter-CTA communication is the per-step exchange of the group's activation vectors through global memory
// (L2-resident), guarded by a monotonic arrive counter per group.
//
// The per-step product is a "swapped" tcgen05 MMA:  D[row, video] = sum_k W[row, k] * X[video, k]
//   A = weight slice (rows x K, K-major, 128-byte swizzle, resident in shared memory),
//   B = the group's activations (bs x K, K-major, 128-byte swizzle, refreshed every step),
//   D in TMEM: lane = weight row, column = video of the group.
#pragma once
#include "common.cuh"

namespace pvcr {

constexpr int PERSIST_THREADS = 256;

// Optional in-kernel phase timing (tuning aid): when non-null, CTA 0 / thread 0 of a persistent kernel stores
// clock64() at up to PHASE_SLOTS points of every step into this device buffer ([step][slot]).
constexpr int PHASE_SLOTS = 16;
long long* debug_phase_buffer();       // null unless pvcr_debug_phase_timing(1) was called

#ifdef __CUDACC__

__device__ __forceinline__ void phase_stamp(long long* dbg, int step, int slot) {
  if (dbg && blockIdx.x == 0 && threadIdx.x == 0) dbg[step * PHASE_SLOTS + slot] = clock64();
}

// Byte offset of element (row, k) inside a K-major SWIZZLE_128B operand whose 64-column k-blocks hold
// `rows_alloc` rows each (rows at 128 B pitch; 16-byte chunk index XORed with row & 7), for k % 8 == 0.
__device__ __forceinline__ uint32_t sw128_offset(int row, int k, int rows_alloc) {
  const int kb = k >> 6, c = (k >> 3) & 7;
  return (uint32_t)kb * (uint32_t)rows_alloc * 128u + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4);
}

// Copy rows [row0, row0+nrows) x K bf16 from global (row stride ld elements, rows >= row_limit read as zero) into a
// SW128 operand at local rows [lrow0, lrow0+nrows).  All threads of the CTA participate; 16-byte L2 loads (.cg).
__device__ __forceinline__ void load_operand_rows(uint8_t* dst, int rows_alloc, int lrow0, const bf16* src,
                                                  long long ld, long long row0, int nrows, long long row_limit,
                                                  int K, int nthr = 0) {
  const int chunks = K >> 3, total = nrows * chunks, step = nthr ? nthr : (int)blockDim.x;
  // (row, chunk) of element i = threadIdx.x + j*step advanced incrementally: no divisions inside the loops.
  const int dlr = step / chunks, dch = step - dlr * chunks;
  int lr = threadIdx.x / chunks, ch = threadIdx.x - lr * chunks;
  // batches of 8 independent 16-byte loads per thread: one L2 round trip per batch instead of one per chunk
  for (int base = threadIdx.x; base < total; base += 8 * step) {
    uint4 v[8];
    int lr_j = lr, ch_j = ch;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = make_uint4(0u, 0u, 0u, 0u);
      if (base + j * step < total && row0 + lr_j < row_limit)
        v[j] = __ldcg(reinterpret_cast<const uint4*>(src + (row0 + lr_j) * ld + ch_j * 8));
      lr_j += dlr; ch_j += dch;
      if (ch_j >= chunks) { ch_j -= chunks; ++lr_j; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (base + j * step < total)
        *reinterpret_cast<uint4*>(dst + sw128_offset(lrow0 + lr, ch * 8, rows_alloc)) = v[j];
      lr += dlr; ch += dch;
      if (ch >= chunks) { ch -= chunks; ++lr; }
    }
  }
}

// Same copy with cp.async (LDGSTS): no register staging, every 16-byte chunk of the operand in flight at once, so a
// whole operand (or several K-chunks of it) costs one L2 round trip.  Rows >= row_limit are zero-filled (src-size 0).
// Completion: cp_async_commit() + cp_async_wait<N>() by the issuing thread, then fence.proxy.async before the MMA.
__device__ __forceinline__ void load_operand_rows_async(uint8_t* dst, int rows_alloc, int lrow0, const bf16* src,
                                                        long long ld, long long row0, int nrows, long long row_limit,
                                                        int K, int nthr = 0) {
  const int chunks = K >> 3, total = nrows * chunks, step = nthr ? nthr : (int)blockDim.x;
  const int dlr = step / chunks, dch = step - dlr * chunks;
  int lr = threadIdx.x / chunks, ch = threadIdx.x - lr * chunks;
  const uint32_t dbase = smem_u32(dst);
  for (int i = threadIdx.x; i < total; i += step) {
    const bool ok = row0 + lr < row_limit;
    const bf16* g = src + (ok ? (row0 + lr) * ld + ch * 8 : 0);
    const uint32_t d = dbase + sw128_offset(lrow0 + lr, ch * 8, rows_alloc);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(g), "r"(ok ? 16 : 0) : "memory");
    lr += dlr; ch += dch;
    if (ch >= chunks) { ch -= chunks; ++lr; }
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Group barrier on a monotonic counter.  arrive: all of this CTA's global writes of the step are published;
// wait: the counter has reached `target` arrivals.  Bounded spin: a protocol bug traps instead of hanging the GPU.
// worker-only CTA barrier (named barrier 1) for kernels that keep a dedicated MMA-issue warp out of the step loop
template <int NWORKERS>
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NWORKERS) : "memory"); }
template <int NWORKERS>
__device__ __forceinline__ void group_arrive_w(unsigned* ctr) {
  worker_sync<NWORKERS>();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
  }
}
template <int NWORKERS>
__device__ __forceinline__ void group_wait_w(const unsigned* ctr, unsigned target) {
  if (threadIdx.x == 0) {
    if (ld_acquire_u32(ctr) < target) {
      const long long t0 = clock64();
      while (ld_acquire_u32(ctr) < target) {
        if (clock64() - t0 > 4000000000LL) __trap();
      }
    }
  }
  worker_sync<NWORKERS>();
}

__device__ __forceinline__ void group_arrive(unsigned* ctr) {
  __syncthreads();
  // release-reduction: publishes (cumulatively, through the CTA barrier above) every thread's writes of this phase
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
}
__device__ __forceinline__ void group_wait(const unsigned* ctr, unsigned target) {
  if (threadIdx.x == 0) {
    if (ld_acquire_u32(ctr) < target) {
      const long long t0 = clock64();
      while (ld_acquire_u32(ctr) < target) {
        if (clock64() - t0 > 4000000000LL) __trap();
      }
    }
  }
  __syncthreads();      // orders every thread's later loads after thread 0's acquire
}

// ---- small all-to-one exchanges by polling the DATA --------------------------------------------------------------
// When a phase hands a few hundred floats to ONE consumer thread block (q and dctx of a video: 16 values from each of a
// group's 32 CTAs), the consumer threads poll the words they need until none carries the sentinel the buffer was filled
// with before the launch (0xFFFFFFFF, a NaN the kernels never produce).  No release fence, counter update, counter poll
// and CTA barrier on the way: one store propagation + one load round trip instead of ~1.3 us.
constexpr unsigned XCH_SENTINEL = 0xFFFFFFFFu;
__device__ __forceinline__ float4 ld_volatile_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.gpu.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool has_sentinel(const float4& a, const float4& b) {
  return __float_as_uint(a.x) == XCH_SENTINEL || __float_as_uint(a.y) == XCH_SENTINEL || __float_as_uint(a.z) == XCH_SENTINEL ||
         __float_as_uint(a.w) == XCH_SENTINEL || __float_as_uint(b.x) == XCH_SENTINEL || __float_as_uint(b.y) == XCH_SENTINEL ||
         __float_as_uint(b.z) == XCH_SENTINEL || __float_as_uint(b.w) == XCH_SENTINEL;
}
// 8 consecutive floats at p (32-byte aligned), polled until all have been written
__device__ __forceinline__ void poll_f8(const float* p, float4& a, float4& b) {
  a = ld_volatile_f4(p); b = ld_volatile_f4(p + 4);
  if (has_sentinel(a, b)) {
    const long long t0 = clock64();
    do {
      a = ld_volatile_f4(p); b = ld_volatile_f4(p + 4);
      if (clock64() - t0 > 4000000000LL) __trap();
    } while (has_sentinel(a, b));
  }
}

// ---- exchanged operands by TMA ------------------------------------------------------------------------------------
// One thread waits for the group counter and then fetches the operand with a few bulk-tensor copies (one 64-column
// k-block x `rows_alloc` rows box each) straight into the SW128 buffer the MMAs read; completion is an mbarrier
// transaction count.  Replaces `rows * K / 8` 16-byte cp.async requests spread over the CTA plus a CTA barrier.
__device__ __forceinline__ void spin_until(const unsigned* ctr, unsigned target) {
  if (ld_acquire_u32(ctr) < target) {
    const long long t0 = clock64();
    while (ld_acquire_u32(ctr) < target) {
      if (clock64() - t0 > 4000000000LL) __trap();
    }
  }
}
// k0: first column (multiple of 64), nkb k-blocks -> dst k-blocks [dst_kb0, dst_kb0 + nkb); rows [row0, row0 + rows_alloc)
// of slab `slab` (rows past the tensor's extent arrive as zeros)
__device__ __forceinline__ void tma_fetch_operand(uint8_t* dst, int rows_alloc, int dst_kb0, const CUtensorMap* tm,
                                                  uint64_t* bar, int k0, int nkb, int row0, int slab) {
  // the acquire that observed the producers' generic-proxy stores orders before the async-proxy reads below
  fence_proxy_async();
  mbar_arrive_expect_tx(bar, (uint32_t)nkb * (uint32_t)rows_alloc * 128u);
  for (int kb = 0; kb < nkb; ++kb)
    tma_load_3d(dst + (size_t)(dst_kb0 + kb) * rows_alloc * 128, tm, bar, k0 + kb * 64, row0, slab);
}

// D[128 x N] (TMEM, fp32) = A[128 x K] * B[N x K]^T, both SW128 K-major in shared memory; issued by one thread.
__device__ __forceinline__ void issue_swapped_mma(uint32_t tmem_d, uint32_t a_base, int a_rows_alloc, uint32_t b_base,
                                                  int b_rows_alloc, int K, uint32_t idesc, uint64_t* done_bar) {
  const int KB = K >> 6;
  for (int kb = 0; kb < KB; ++kb) {
    const uint64_t da = umma_desc_k128(a_base + (uint32_t)kb * (uint32_t)a_rows_alloc * 128u);
    const uint64_t db = umma_desc_k128(b_base + (uint32_t)kb * (uint32_t)b_rows_alloc * 128u);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
  }
  umma_commit(done_bar);
}

// Same product over one K-chunk, without the commit: A k-blocks [a_kb0, a_kb0 + Kc/64), B k-blocks [0, Kc/64).
__device__ __forceinline__ void issue_mma_chunk(uint32_t tmem_d, uint32_t a_base, int a_rows_alloc, int a_kb0,
                                                uint32_t b_base, int b_rows_alloc, int Kc, uint32_t idesc,
                                                bool accumulate) {
  const int KB = Kc >> 6;
  for (int kb = 0; kb < KB; ++kb) {
    const uint64_t da = umma_desc_k128(a_base + (uint32_t)(a_kb0 + kb) * (uint32_t)a_rows_alloc * 128u);
    const uint64_t db = umma_desc_k128(b_base + (uint32_t)kb * (uint32_t)b_rows_alloc * 128u);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, accumulate || (kb | k) != 0);
  }
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// M = 64 variant (tcgen05.mma with a 64-row A tile, cta_group::1): the accumulator row r lives in TMEM lane
// 32*(r/16) + r%16, i.e. every warp quadrant holds 16 rows in its first 16 lanes.
__device__ __forceinline__ void tmem64_to_smem_cols(uint32_t tmem_base, float* S, int s_ld, int nrows, int ncols) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = warp * 16 + lane;
  float v[16];
  for (int c0 = 0; c0 < ncols; c0 += 16) {
    tmem_ld_32x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    if (lane < 16 && row < nrows) {
#pragma unroll
      for (int j = 0; j < 16; ++j) S[(c0 + j) * s_ld + row] = v[j];
    }
  }
}

// Move the accumulator D[row = TMEM lane, col < ncols] to shared memory as S[col * s_ld + row] for rows < nrows.
// Executed by warps 0..3 (thread == lane).  ncols is a multiple of 16.
__device__ __forceinline__ void tmem_to_smem_cols(uint32_t tmem_base, float* S, int s_ld, int nrows, int ncols) {
  const int warp = threadIdx.x >> 5, row = threadIdx.x;
  float v[16];
  for (int c0 = 0; c0 < ncols; c0 += 16) {
    tmem_ld_32x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    if (row < nrows) {
#pragma unroll
      for (int j = 0; j < 16; ++j) S[(c0 + j) * s_ld + row] = v[j];
    }
  }
}

// ---- warp-level mma.sync path for the skinny backward products -------------------------------------------------
// tcgen05.mma costs ~70 cycles per instruction on these shapes whatever M (64/128), N (16/32), the number of
// accumulators or issuing threads (measured), i.e. K/16 x 70 cycles per product.  A 16..32-row weight slice
// times 16..32 videos is only a handful of m16n8k16 tiles, so for the long-K backward products the legacy warp MMA
// with the K range split over the 8 warps is several times faster; operands stay in the same swizzled layout.
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// A fragment (16 rows x 16 k) of a SW128 operand: rows m0.., k0.. (k0 % 16 == 0)
__device__ __forceinline__ void load_a_frag(uint32_t base, int rows_alloc, int m0, int k0, uint32_t (&a)[4]) {
  const int lane = threadIdx.x & 31;
  const int row = m0 + (lane & 15), k = k0 + ((lane >> 4) << 3);
  ldmatrix_x4(base + sw128_offset(row, k, rows_alloc), a);      // a0: rows 0-7 k 0-7, a1: rows 8-15 k 0-7, a2/a3: k 8-15
}
// B fragments for two n-tiles (16 "videos" x 16 k) of a SW128 operand stored [video][k]:
// b[0], b[1] = (k 0-7, k 8-15) of videos n0..n0+7;  b[2], b[3] = same for videos n0+8..n0+15
__device__ __forceinline__ void load_b_frag2(uint32_t base, int rows_alloc, int n0, int k0, uint32_t (&b)[4]) {
  const int lane = threadIdx.x & 31;
  const int row = n0 + (lane & 7) + ((lane >> 4) << 3), k = k0 + (((lane >> 3) & 1) << 3);
  ldmatrix_x4(base + sw128_offset(row, k, rows_alloc), b);
}

#endif  // __CUDACC__

}  // namespace pvcr
