// Host side of the tcgen05 GEMM: TMA tensor-map construction (cached) and the plain-store entry point.
#include <mutex>
#include <unordered_map>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "gemm_sm100.cuh"

namespace pvcr {

static thread_local char g_last_error[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_last_error; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct MapKey {
  const void* ptr;
  long long ld, slab_stride;
  int rows, slabs, K, box_rows;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return (size_t)h;
  }
};
static_assert(sizeof(MapKey) % 8 == 0, "MapKey must be 8-byte granular");

// 3-D tensor map over (k, row, slab), box = 64 x box_rows x 1, 128-byte swizzle, OOB elements read as zero.
int make_tensor_map(CUtensorMap* out, const OperandView& v, int K, int box_rows, int box_k) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = v.ptr; key.ld = v.ld; key.slab_stride = v.slab_stride; key.rows = v.rows; key.slabs = v.slabs;
  key.K = K; key.box_rows = box_rows + (box_k == GEMM_BK ? 0 : 1 << 20);
  PVCR_REQUIRE(box_k == GEMM_BK || box_k == 32, "tensor map: box of %d columns not supported", box_k);
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return PVCR_OK; }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_last_error("cuTensorMapEncodeTiled not available from the driver"); return PVCR_ERR_DRIVER; }
  PVCR_REQUIRE((reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0, "tensor map: base %p not 16-byte aligned", v.ptr);
  PVCR_REQUIRE(v.ld % 8 == 0 && v.ld >= K, "tensor map: ld=%lld must be a multiple of 8 and >= K=%d", v.ld, K);
  const int slabs = v.slabs > 0 ? v.slabs : 1;
  long long slab_stride = v.slab_stride;
  if (slabs == 1 && slab_stride <= 0) slab_stride = v.ld * (long long)v.rows;
  PVCR_REQUIRE(slab_stride % 8 == 0, "tensor map: slab stride %lld must be a multiple of 8", slab_stride);
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)v.rows, (cuuint64_t)slabs};
  cuuint64_t gstr[2] = {(cuuint64_t)v.ld * 2, (cuuint64_t)slab_stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(v.ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, box_k == GEMM_BK ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): ptr=%p K=%d rows=%d slabs=%d ld=%lld slab=%lld", (int)r,
                   v.ptr, K, v.rows, slabs, v.ld, slab_stride);
    return PVCR_ERR_DRIVER;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 65536) cache.clear();
  cache[key] = *out;
  return PVCR_OK;
}

// 3-D tensor map over (mn, k row, slab) of a row-major [k_rows, mn_cols] bf16 matrix, box = 64 x 64 x 1.
int make_tensor_map_mn(CUtensorMap* out, const OperandView& v, int mn_cols) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = v.ptr; key.ld = v.ld; key.slab_stride = v.slab_stride; key.rows = v.rows; key.slabs = v.slabs;
  key.K = mn_cols; key.box_rows = -64;
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return PVCR_OK; }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_last_error("cuTensorMapEncodeTiled not available from the driver"); return PVCR_ERR_DRIVER; }
  PVCR_REQUIRE((reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0, "tensor map (mn): base %p not 16-byte aligned", v.ptr);
  PVCR_REQUIRE(v.ld % 8 == 0 && v.ld >= mn_cols, "tensor map (mn): ld=%lld must be a multiple of 8 and >= %d", v.ld, mn_cols);
  const int slabs = v.slabs > 0 ? v.slabs : 1;
  long long slab_stride = v.slab_stride;
  if (slabs == 1 && slab_stride <= 0) slab_stride = v.ld * (long long)v.rows;
  cuuint64_t gdim[3] = {(cuuint64_t)mn_cols, (cuuint64_t)v.rows, (cuuint64_t)slabs};
  cuuint64_t gstr[2] = {(cuuint64_t)v.ld * 2, (cuuint64_t)slab_stride * 2};
  cuuint32_t box[3] = {64, 64, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(v.ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (mn) failed (%d): ptr=%p cols=%d rows=%d ld=%lld", (int)r, v.ptr, mn_cols,
                   v.rows, v.ld);
    return PVCR_ERR_DRIVER;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 65536) cache.clear();
  cache[key] = *out;
  return PVCR_OK;
}

// C[M,N] (+)= A^T B for row-major bf16 A [K, M], B [K, N] (no transposed copies): weight gradients dW = dY^T X.
int gemm_mn_store(const OperandView& a, const OperandView& b, int M, int N, int K, float* C, long long ldc,
                  int accumulate, cudaStream_t stream) {
  EpiStore epi{C, ldc, 0, nullptr, 0, accumulate, M, N, 0, 1};
  GemmCoords gc{M, N, (int)round_up(K, GEMM_BK), 0, 0, 0, 0};
  // few output tiles, long contraction (the recurrent weight gradients): spread K ranges over the idle SMs
  const long long tiles = (long long)cdiv(M, GEMM_BM) * cdiv(N, 256);
  const int total_kb = gc.K / GEMM_BK;
  static const bool no_split = getenv("PVCR_NO_SPLITK") != nullptr;
  if (tiles <= 8 && total_kb >= 16 && !no_split) {      // measured: beyond a handful of tiles the atomics cost more than they save
    int splits = (int)(sm_count() / tiles);
    if (splits > total_kb / 8) splits = total_kb / 8;
    if (splits > 1) {
      const int kb_per = (total_kb + splits - 1) / splits;
      gc.k_splits = (total_kb + kb_per - 1) / kb_per;
      if (!accumulate) PVCR_CUDA_CHECK(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, stream));
      epi.atomic = 1;
    }
  }
  return launch_gemm_tn_persistent<256, 4, EpiStore, true, true>(a, b, gc, 1, epi, stream);
}

// Nine products sharing A:  C_s[M,N] = A^T B[koff[s] + k, :]  (s < taps <= 9), one batched launch (C_s = C + s * c_stride).
// a: [K rows, M cols], b: [>= max koff + K rows, N cols], both row-major bf16; rows of `a` past its end read as zero.
int gemm_mn_taps_store(const OperandView& a, const OperandView& b, int M, int N, int K, int taps, const int* koff, float* C,
                       long long ldc, long long c_stride, cudaStream_t stream) {
  PVCR_REQUIRE(taps >= 1 && taps <= 9, "gemm_mn_taps_store: %d taps", taps);
  EpiStore epi{C, ldc, c_stride, nullptr, 0, 0, M, N, 0, 1};
  GemmCoords gc{M, N, (int)round_up(K, GEMM_BK), 0, 0, 0, 0};
  gc.b_z_koff = 1;
  for (int s = 0; s < taps; ++s) gc.a_tap_off[s] = koff[s];
  return launch_gemm_tn_persistent<256, 4, EpiStore, true, true>(a, b, gc, taps, epi, stream);
}

// C[M,N] (+)= A B for bf16 A [M, K] (K-major) and row-major B [K rows, N cols] (MN-major): data gradients dX = dY W
// straight from the forward weight planes.  K = rows of b actually present (rows beyond read as zero).
int gemm_kn_store(const OperandView& a, const OperandView& b, int M, int N, int K, float* C, long long ldc,
                  int accumulate, cudaStream_t stream) {
  EpiStore epi{C, ldc, 0, nullptr, 0, accumulate, M, N, 0, 1};
  GemmCoords gc{M, N, (int)round_up(K, GEMM_BK), 0, 0, 0, 0};
  // long contraction, fewer tiles than half the SMs (d hs = d logits W_v: 60 tiles, K = 23 040): two K ranges per tile
  static const int kn_split = getenv("PVCR_KN_SPLIT") ? atoi(getenv("PVCR_KN_SPLIT")) : 1;
  const long long tiles = (long long)cdiv(M, GEMM_BM) * cdiv(N, 256);
  if (kn_split > 1 && tiles * kn_split <= sm_count() && gc.K / GEMM_BK >= 64 * kn_split) {
    gc.k_splits = kn_split;
    if (!accumulate) PVCR_CUDA_CHECK(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, stream));
    epi.atomic = 1;
  }
  return launch_gemm_tn_persistent<256, 4, EpiStore, false, true>(a, b, gc, 1, epi, stream);
}

// C[z] = A[z] * B[z]^T (+bias) (+C).  Tile choice: 128x256 when N is wide enough to fill the machine, else 128x128.
int gemm_store(const OperandView& a, const OperandView& b, const GemmCoords& gc, int grid_z, float* C, long long ldc,
               long long c_zstride, const float* bias, long long bias_zstride, int accumulate,
               cudaStream_t stream) {
  EpiStore epi{C, ldc, c_zstride, bias, bias_zstride, accumulate, gc.M, gc.N, 0, 1};
  // skinny bf16x3 products on compact weight planes (decoding: [q | gh] = h [W_q; W_hh]^T with M = batch <= 128): the split3
  // kernel on 64-column tiles instead of six split-K ranges + atomics.  A/B knob: PVCR_NO_SPLIT3_STORE=1.
  static const bool s3_off = getenv("PVCR_NO_SPLIT3_STORE") != nullptr;
  if (!s3_off && b.kp && b.terms == 3 && gc.K == 6 * b.kp && grid_z == 1 && gc.M <= GEMM_BM && gc.k_splits <= 1) {
    GemmCoords g3 = gc;
    g3.K = b.kp; g3.b_kp = 0; g3.b_terms = 0; g3.b_evict_last = 1;
    static const int bk_knob = getenv("PVCR_SPLIT3_STORE_BK") ? atoi(getenv("PVCR_SPLIT3_STORE_BK")) : 64;
    if (bk_knob == 32) return launch_gemm_split3<64, EpiStore, 4, 32>(a, b, g3, b.kp, epi, stream);
    return launch_gemm_split3<64, EpiStore, 4>(a, b, g3, b.kp, epi, stream);
  }
  // several row tiles of a bf16x3 product on compact planes (decoding B > 128: [q | gh], W_c ctx): 256-column tiles, 32-column
  // K chunks -- half the operand traffic per tile of the six-plane walk.  A/B knob: PVCR_NO_SPLIT3_STORE_BIG=1.
  static const bool s3_big_off = getenv("PVCR_NO_SPLIT3_STORE_BIG") != nullptr;
  if (!s3_off && !s3_big_off && b.kp && b.terms == 3 && gc.K == 6 * b.kp && grid_z == 1 && gc.M > GEMM_BM && gc.N >= 256 &&
      gc.k_splits <= 1) {
    GemmCoords g3 = gc;
    g3.K = b.kp; g3.b_kp = 0; g3.b_terms = 0;
    return launch_gemm_split3<256, EpiStore, 8, 32>(a, b, g3, b.kp, epi, stream);
  }
  if (gc.a_taps) {       // shifted-row taps: persistent kernel only (its producer maps k-blocks to (tap, column block))
    PVCR_REQUIRE(gc.N >= 256 && grid_z == 1 && gc.K % (gc.a_taps * GEMM_BK) == 0, "gemm (taps): N=%d K=%d taps=%d", gc.N, gc.K, gc.a_taps);
    return launch_gemm_tn_persistent<256, 4, EpiStore>(a, b, gc, 1, epi, stream);
  }
  const long long tiles256 = (long long)cdiv(gc.N, 256) * cdiv(gc.M, GEMM_BM) * grid_z;
  static const bool no_persist = getenv("PVCR_NO_PERSIST_GEMM") != nullptr;
  if (gc.N >= 256 && tiles256 >= 64 && !no_persist)
    return launch_gemm_tn_persistent<256, 4, EpiStore>(a, b, gc, grid_z, epi, stream);
  if (gc.N >= 256 && tiles256 >= sm_count()) return launch_gemm_tn<256, 4, EpiStore>(a, b, gc, grid_z, epi, stream);
  // few output tiles but a long contraction (weight / data gradients over all timesteps): split K across CTAs
  const long long tiles128 = (long long)cdiv(gc.N, 128) * cdiv(gc.M, GEMM_BM);
  const int total_kb = gc.K / GEMM_BK;
  static const bool no_split = getenv("PVCR_NO_SPLITK") != nullptr;
  if (grid_z == 1 && tiles128 <= 24 && total_kb >= 32 && !no_split) {     // measured: atomics lose beyond ~24 tiles
    int splits = (int)((2 * sm_count() + tiles128 - 1) / tiles128);
    if (splits > total_kb / 8) splits = total_kb / 8;
    if (splits > 1) {
      const int kb_per = (total_kb + splits - 1) / splits;
      splits = (total_kb + kb_per - 1) / kb_per;          // no empty K range
      if (!accumulate)
        PVCR_CUDA_CHECK(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * gc.N, gc.M, stream));
      GemmCoords g2 = gc;
      g2.k_splits = splits;
      epi.atomic = 1;
      return launch_gemm_tn<128, 3, EpiStore>(a, b, g2, 1, epi, stream);
    }
  }
  return launch_gemm_tn<128, 3, EpiStore>(a, b, gc, grid_z, epi, stream);
}

// ---- projection + row arg-max in one pass (greedy decoding: logits_i = h_i W_v^T + b, word_{i+1} = argmax) -------------
// The epilogue keeps a running (max, first index) per accumulator row and column part while it adds the bias (and stores
// the logits when the caller wants them); a one-warp-per-row kernel combines the parts.  The logits are not read back.
struct EpiArgmax {
  static constexpr int SMEM_PER_WARP = 0;
  float* C; long long ldc;              // nullable: logits not materialised
  const float* bias;
  float* pmax; int* pidx;
  int M, N, nparts;
  float m_run; int i_run;
  __device__ __forceinline__ void begin(int, int) { m_run = -INFINITY; i_run = 0x7fffffff; }
  __device__ __forceinline__ void chunk(int row, int col0, int, float (&v)[32]) {
    if (row >= M) return;
    const bool full = col0 + 32 <= N;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (full || col0 + j < N) {
        v[j] += bias ? __ldg(bias + col0 + j) : 0.f;
        if (v[j] > m_run) { m_run = v[j]; i_run = col0 + j; }      // ascending columns: the first maximum wins
      }
    }
    if (C) {
      float* c = C + (long long)row * ldc + col0;
      if (full && ((reinterpret_cast<uintptr_t>(c) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) __stcs(reinterpret_cast<float4*>(c + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
      } else {                                  // streaming stores: the logits are not read back inside the decode loop
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < N) __stcs(c + j, v[j]);
      }
    }
  }
  __device__ __forceinline__ void end(int row, int part, int) {
    if (row < M) { pmax[(long long)row * nparts + part] = m_run; pidx[(long long)row * nparts + part] = i_run; }
  }
};
__global__ void __launch_bounds__(128) argmax_parts_kernel(const float* __restrict__ pmax, const int* __restrict__ pidx, int R,
                                                           int nparts, long long* out, long long out_stride, long long* next) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  float m = -INFINITY;
  int mi = 0x7fffffff;
  for (int j = lane; j < nparts; j += 32) {
    const float v = pmax[(long long)row * nparts + j];
    const int i = pidx[(long long)row * nparts + j];
    if (v > m || (v == m && i < mi)) { m = v; mi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
  }
  if (lane == 0) {
    if (mi == 0x7fffffff) mi = 0;                      // all-NaN row: torch.argmax returns an index too
    out[(long long)row * out_stride] = mi;
    if (next) next[row] = mi;
  }
}
int gemv_f32_parts(int rows);
size_t gemm_argmax_scratch(int M, int N) {
  const int parts = cdiv(N, 128) * 2 > gemv_f32_parts(N) ? cdiv(N, 128) * 2 : gemv_f32_parts(N);
  return (size_t)M * parts * (sizeof(float) + sizeof(int)) + 256;
}
int argmax_combine(const float* pmax, const int* pidx, int R, int nparts, long long* out, long long out_stride, long long* next,
                   cudaStream_t st) {
  { LaunchScope ls_(KC_LOSS, st);
    argmax_parts_kernel<<<cdiv(R, 4), 128, 0, st>>>(pmax, pidx, R, nparts, out, out_stride, next);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}
// logits (nullable) [M, ldc] = A B^T + bias;  out[row * out_stride] = next[row] = argmax_n.  scratch: gemm_argmax_scratch bytes.
// out == nullptr: the combine pass is left to the consumer (gru_gate_fwd folds it into the next step's gate kernel);
// parts (nullable) receives where the (max, index) partials are and how many there are per row.
int gemm_argmax(const OperandView& a, const OperandView& b, int M, int N, int Kcat, const float* bias, float* logits,
                long long ldc, long long* out, long long out_stride, long long* next, void* scratch, cudaStream_t st,
                ArgmaxParts* parts) {
  GemmCoords gc{M, N, Kcat, 0, 0, 0, 0};
  if (b.kp) { gc.b_kp = b.kp; gc.b_terms = split_b_terms(b.terms); }
  // the W_v planes are read again by the next step's launch: ask L2 to evict them last (A/B knob PVCR_DECODE_WV_HINT=0)
  static const int wv_hint = getenv("PVCR_DECODE_WV_HINT") ? atoi(getenv("PVCR_DECODE_WV_HINT")) : -1;
  // Tile width (A/B knob PVCR_ARGMAX_BN=192): Vc = 23 000 gives 90 tiles of 256 columns (61 % of the SMs) or 120 of 192;
  // measured, the narrower tiles are SLOWER inside the decode loop (3.04 -> 3.17 ms per batch): the SMs the 256-wide
  // tiling leaves free are what the overlapped query / attention / context half of the next step runs on.
  static const int bn_knob = getenv("PVCR_ARGMAX_BN") ? atoi(getenv("PVCR_ARGMAX_BN")) : 0;
  const int bn = bn_knob == 192 ? 192 : 256;
  // bf16x3 with compact W_v planes: the split3 kernel loads every term block once per K chunk (half the L2 -> SM traffic of
  // walking the six virtual planes) on 160-column tiles (144 CTAs for Vc = 23 000).  A/B knob: PVCR_ARGMAX_SPLIT3=0 / 128.
  static const int s3_knob = getenv("PVCR_ARGMAX_SPLIT3") ? atoi(getenv("PVCR_ARGMAX_SPLIT3")) : 128;
  // (one row tile only: beyond 128 rows the product turns compute-bound -- B = 1024 decodes 103 k captions/s on the generic
  // 256-column tiles, 90 k on this kernel -- and the wide tiles re-read less of A)
  // 256: 256-column tiles on 32-column K chunks (3 stages of 72 KB): the least operand traffic per FLOP of all variants
  // (1.15 MB per 128 x 256 tile against 2.3 MB on the generic kernel) -- for several row tiles (B > 128), where the step is
  // bound by what an SM can take in, and as an A/B variant at B <= 128
  static const int s3_big = getenv("PVCR_ARGMAX_SPLIT3_BIG") ? atoi(getenv("PVCR_ARGMAX_SPLIT3_BIG")) : 1;
  const bool s3_ok = s3_knob != 0 && b.kp && b.terms == 3 && Kcat == 6 * b.kp && bn_knob == 0;
  const bool s3_256 = s3_ok && (s3_knob == 256 || (M > GEMM_BM && s3_big));
  const bool s3 = s3_ok && (M <= GEMM_BM || s3_256);
  const int nparts = s3 ? (s3_256 ? cdiv(N, 256) * 2 : (s3_knob == 128 ? cdiv(N, 128) * 2 : cdiv(N, 160))) : cdiv(N, bn) * 2;
  // L2 policy of the W_v loads: 70 MB of planes do not stay in L2 from one step to the next anyway (measured: 73 MB of DRAM
  // reads per launch with any policy), so the split3 kernel, which reads every byte once, streams them evict_first and leaves
  // the L2 to what the other half of the step re-reads (projected frames, keys); the generic kernel re-reads planes within a
  // launch and gives no hint (evict_last measured 1 % slower there).  PVCR_DECODE_WV_HINT = 0 / 1 / 2 overrides.
  gc.b_evict_last = wv_hint >= 0 ? wv_hint : (s3 && M <= GEMM_BM ? 2 : 0);      // (row tiles side by side share a W_v tile)
  EpiArgmax epi{};
  epi.C = logits; epi.ldc = ldc; epi.bias = bias; epi.M = M; epi.N = N; epi.nparts = nparts;
  epi.pmax = reinterpret_cast<float*>(scratch);
  epi.pidx = reinterpret_cast<int*>(epi.pmax + (size_t)M * nparts);
  if (s3) {
    GemmCoords g3 = gc;
    g3.K = b.kp; g3.b_kp = 0; g3.b_terms = 0;
    // L2 prefetch of the CTA's W_v blocks ahead of the ring: measured SLOWER (33.4 -> 37.3 us alone, 1.95 -> 2.06 ms per batch:
    // the prefetches queue in front of the ring's own loads in the TMA unit); off unless PVCR_WV_PREFETCH=1
    static const bool prefetch_on = getenv("PVCR_WV_PREFETCH") != nullptr;
    g3.b_prefetch = prefetch_on && g3.b_evict_last == 2;
    // 128-column tiles on at most 90 CTAs (two tiles each at Vc = 23 000): the other SMs run the recurrent half of the next
    // step next to it (measured: 144 CTAs of 160 columns 2.22 / 2.04 ms per batch, 90 CTAs of 128 columns 2.13 / 1.86)
    static const int cta_knob = getenv("PVCR_ARGMAX_CTAS") ? atoi(getenv("PVCR_ARGMAX_CTAS")) : 90;
    CtaCap cap(M > GEMM_BM ? gemm_cta_cap() : (cta_knob > 0 ? cta_knob : gemm_cta_cap()));
    static const int bk_knob = getenv("PVCR_ARGMAX_BK") ? atoi(getenv("PVCR_ARGMAX_BK")) : 64;
    if (s3_256) PVCR_TRY((launch_gemm_split3<256, EpiArgmax, 8, 32>(a, b, g3, b.kp, epi, st)));
    else if (s3_knob == 128 && bk_knob == 32) PVCR_TRY((launch_gemm_split3<128, EpiArgmax, 8, 32>(a, b, g3, b.kp, epi, st)));
    else if (s3_knob == 128) PVCR_TRY((launch_gemm_split3<128, EpiArgmax, 8>(a, b, g3, b.kp, epi, st)));
    else PVCR_TRY((launch_gemm_split3<160, EpiArgmax, 4>(a, b, g3, b.kp, epi, st)));
  }
  else if (bn == 192) PVCR_TRY((launch_gemm_tn_persistent<192, 4, EpiArgmax>(a, b, gc, 1, epi, st)));
  else PVCR_TRY((launch_gemm_tn_persistent<256, 4, EpiArgmax>(a, b, gc, 1, epi, st)));
  if (parts) { parts->pmax = epi.pmax; parts->pidx = epi.pidx; parts->nparts = nparts; }
  if (!out) return PVCR_OK;
  { LaunchScope ls_(KC_LOSS, st);
    argmax_parts_kernel<<<cdiv(M, 4), 128, 0, st>>>(epi.pmax, epi.pidx, M, nparts, out, out_stride, next);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

}  // namespace pvcr
