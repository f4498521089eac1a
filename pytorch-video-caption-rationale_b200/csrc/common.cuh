// Common device helpers for the sm_100a captioning kernels: PTX wrappers for mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (TMEM alloc / mma / commit / ld) and small math utilities.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pvcr {

typedef __nv_bfloat16 bf16;

#define PVCR_OK 0
#define PVCR_ERR_ARG -1
#define PVCR_ERR_CUDA -2
#define PVCR_ERR_WORKSPACE -3
#define PVCR_ERR_DRIVER -4

void set_last_error(const char* fmt, ...);
int sm_count();          // multiprocessors of the current device (148 on B200); grids and split heuristics are sized from it

#define PVCR_CUDA_CHECK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      pvcr::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return PVCR_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

#define PVCR_REQUIRE(cond, ...)                    \
  do {                                             \
    if (!(cond)) {                                 \
      pvcr::set_last_error(__VA_ARGS__);           \
      return PVCR_ERR_ARG;                         \
    }                                              \
  } while (0)

#define PVCR_TRY(expr)            \
  do {                            \
    int _r = (expr);              \
    if (_r != PVCR_OK) return _r; \
  } while (0)

// ---- launch accounting (pvcr_prof_* in include/pvcr_b200.h) ---------------------------------------
// Every kernel launcher opens a LaunchScope: it counts the launch per kernel class and, when timing is
// enabled (bench.py's roofline leg), brackets it with a CUDA-event pair on the launching stream.
enum KernelClass { KC_GEMM = 0, KC_STAGE, KC_GATE, KC_ATTN, KC_LOSS, KC_GRU_FWD, KC_GRU_BWD, KC_DEC_FWD, KC_DEC_BWD, KC_MISC,
                   KC_COUNT };
struct LaunchScope {
  int cls; cudaStream_t st; void* rec; bool ext;
  LaunchScope(int cls, cudaStream_t st, double work = 0.0);
  ~LaunchScope();
};

// While a PdlScope(true) is alive on this host thread, the launchers that support it (split3 GEMM, projected-value attention
// step) launch with programmatic stream serialization: the decode loop's dependent launch chain then overlaps every kernel's
// launch latency and prologue with the tail of its predecessor.
bool pdl_enabled();
struct PdlScope {
  bool prev;
  explicit PdlScope(bool on);
  ~PdlScope();
};

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (surfacing as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while
// its predecessor in the stream still runs; pdl_wait() blocks until that predecessor has completed and its writes are
// visible (a no-op for a normal launch), pdl_launch_dependents() lets the successor's CTAs in early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// L2 eviction-priority hint for operands that are re-read launch after launch (decoding: the W_v planes of every step)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_3d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 int c2, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}

// TMA prefetch of a box into L2 (no shared-memory destination, no barrier): DRAM fetches far ahead of the pipeline
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, one CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp receives TMEM lane (base_lane + i), columns c..c+31.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte swizzled shared-memory operand descriptor (rows at 128 B pitch, 8-row groups
// 1024 B apart).  Field layout follows the sm_100 shared-memory matrix descriptor:
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// K-major, 64-byte swizzled operand (rows of 32 bf16 at 64 B pitch, 8-row groups 512 B apart; layout 4 = SWIZZLE_64B)
__device__ __forceinline__ uint64_t umma_desc_k64(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// MN-major, 128-byte swizzled operand (the matrix is stored [k][mn] with mn contiguous, e.g. dY for dW = dY^T X):
// swizzle atom = 8 k-rows x 64 mn-elements (128 B per row, 1024 B per atom); LBO = byte distance between 64-element
// mn blocks, SBO = byte distance between 8-row k groups.
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor: bf16 A/B (K-major), fp32 accumulate, tile M x N.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- math
// tanh via one fast exponential: absolute error ~1e-7 (used inside the attention score, where only the absolute
// error matters); saturates correctly for |x| large.
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }
// 2^x on the special-function unit (ex2.approx: 2 ulp; ex2(-inf) = 0)
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// single-MUFU hardware tanh (max relative error ~2^-11)
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// tanh(x) and 1 - tanh(x)^2 from one exponential and one reciprocal, accurate to ~1e-7 relative even where
// 1 - tanh^2 is tiny (the attention backward needs the derivative's RELATIVE accuracy; tanh.approx's 2^-11 is enough
// for the forward score only)
__device__ __forceinline__ void tanh_sech2(float x, float& th, float& sech2) {
  const float t = __expf(-2.f * fabsf(x));
  const float r = __fdividef(1.f, 1.f + t);
  th = copysignf((1.f - t) * r, x);
  sech2 = 4.f * t * r * r;
}
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Split an fp32 value into up to three bf16 terms: x ~= t0 + t1 + t2 (each RN of the running residual).
__device__ __forceinline__ void split3(float x, bf16& t0, bf16& t1, bf16& t2) {
  t0 = __float2bfloat16_rn(x);
  float r = x - __bfloat162float(t0);
  t1 = __float2bfloat16_rn(r);
  r -= __bfloat162float(t1);
  t2 = __float2bfloat16_rn(r);
}

#endif  // __CUDACC__

// Split-precision plane tables.  A GEMM with `nsplit` in {1,2,3} concatenates P = {1,3,6} planes
// along K: plane p of the A-role operand holds term A_TERM[p] of the activation and plane p of the
// B-role operand holds term B_TERM[p] of the weight, so that sum_p A_p * B_p^T reproduces all
// products a_i * b_j with i + j <= nsplit + 1 (smallest terms first, fp32 accumulation).
__host__ __device__ inline int split_planes(int nsplit) { return nsplit == 1 ? 1 : (nsplit == 2 ? 3 : 6); }
// 0-based term index held by plane p; tables packed 2 bits per plane:
//   nsplit 2: A = {1,0,0}  B = {0,1,0}        nsplit 3: A = {2,1,0,1,0,0}  B = {0,1,2,0,1,0}
// role_b == 2: COMPACT weight planes -- the nsplit distinct terms once each (plane t = term t); the GEMM producer maps the
// K-block of virtual plane p to stored plane B_TERM[p] (GemmCoords::b_kp / b_terms), so a weight is stored and streamed
// nsplit times instead of {1,3,6} times.
__host__ __device__ inline int split_planes_role(int nsplit, int role_b) { return role_b == 2 ? nsplit : split_planes(nsplit); }
__host__ __device__ inline unsigned split_b_terms(int nsplit) { return nsplit == 1 ? 0u : (nsplit == 2 ? 4u : 292u); }
__host__ __device__ inline int split_term(int nsplit, int role_b, int p) {
  if (role_b == 2) return p;
  const unsigned code = nsplit == 1 ? 0u : (nsplit == 2 ? (role_b ? 4u : 1u) : (role_b ? 292u : 70u));
  return (int)((code >> (2 * p)) & 3u);
}

}  // namespace pvcr
