// Non-GEMM kernels of the captioning path: operand staging (fp32 -> bf16 split planes, transposes,
// embedding gather / scatter), GRU / LSTM gate math (forward and backward), the additive attention
// step (forward and backward), cross entropy on materialised logits, and the RationaleNet generator head.
#include <cstdlib>

#include "kernels.cuh"

namespace pvcr {

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based RNG) for dropout masks and Gumbel noise
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float philox_uniform(unsigned long long seed, unsigned long long idx) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  // 23 random bits + 0.5: the largest value is (2^23 - 0.5) / 2^23 = 1 - 2^-24, exactly representable in fp32, the
  // smallest 2^-24 -- strictly inside (0,1), so -log(u) is finite and > 0 (as torch's exponential_() guarantees).
  // (24 bits + 0.5 would round 16777215.5 up to 2^24 and return exactly 1.0 once in 2^24 draws.)
  return ((float)(r.x >> 9) + 0.5f) * (1.0f / 8388608.0f);
}
__device__ __forceinline__ unsigned long long stepped_seed(unsigned long long seed, const unsigned long long* step) {
  return step ? seed + __ldg(step) * 0x9E3779B97F4A7C15ull : seed;
}
__device__ __forceinline__ float dropout_scale(const Dropout& d, unsigned long long idx) {
  if (d.p <= 0.f) return 1.f;
  return philox_uniform(stepped_seed(d.seed, d.step), d.offset + idx) >= d.p ? 1.f / (1.f - d.p) : 0.f;
}

// Test hook (pvcr_debug_philox_minmax): smallest and largest uniform over indices [idx0, idx0 + n): both must lie
// strictly inside (0,1) for the in-kernel Exp(1) / Gumbel draws to be finite.
__global__ void philox_minmax_kernel(unsigned long long seed, unsigned long long idx0, unsigned long long n,
                                     unsigned int* minmax_bits) {
  float mn = 2.f, mx = -1.f;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const float u = philox_uniform(seed, idx0 + i);
    mn = fminf(mn, u); mx = fmaxf(mx, u);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {      // positive floats order like their bit patterns
    atomicMin(minmax_bits, __float_as_uint(mn));
    atomicMax(minmax_bits + 1, __float_as_uint(mx));
  }
}
int philox_minmax(unsigned long long seed, unsigned long long idx0, unsigned long long n, float* minmax, cudaStream_t st) {
  const unsigned int init[2] = {0x7f7fffffu, 0u};
  PVCR_CUDA_CHECK(cudaMemcpyAsync(minmax, init, sizeof(init), cudaMemcpyHostToDevice, st));
  philox_minmax_kernel<<<sm_count() * 8, 256, 0, st>>>(seed, idx0, n, reinterpret_cast<unsigned int*>(minmax));
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

__device__ __forceinline__ unsigned long long drop_index(const Dropout& d, long long r, int C, int c) {
  return (unsigned long long)((r * d.row_mul + d.row_add) * C + c);
}

__device__ __forceinline__ void write_split(bf16* row_out, int Kp, int k, int nsplit, int role_b, float x) {
  bf16 t[3];
  split3(x, t[0], t[1], t[2]);
  const int P = split_planes_role(nsplit, role_b);
  for (int p = 0; p < P; ++p) row_out[(long long)p * Kp + k] = t[split_term(nsplit, role_b, p)];
}

// ------------------------------------------------------------------------------------------------
// operand staging
// ------------------------------------------------------------------------------------------------
__global__ void cast_split_kernel(const float* __restrict__ in, long long ld_in, int R, int C, bf16* __restrict__ out,
                                  long long ld_out, int Cp, int nsplit, int role_b,
                                  const float* __restrict__ row_scale, Dropout drop) {
  const int groups = Cp / 8;
  const long long total = (long long)R * groups;
  const int P = split_planes_role(nsplit, role_b);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / groups), c0 = (int)(i % groups) * 8;
    const float* src = in + (long long)r * ld_in + c0;
    float x[8];
    if (c0 + 8 <= C && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (c0 + j < C) ? __ldg(src + j) : 0.f;
    }
    const float rs = row_scale ? __ldg(row_scale + r) : 1.f;
    bf16 t[3][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float val = x[j] * rs;
      if (drop.p > 0.f) val *= dropout_scale(drop, drop_index(drop, r, C, c0 + j));
      split3(val, t[0][j], t[1][j], t[2][j]);
    }
    bf16* dst = out + (long long)r * ld_out + c0;
    for (int p = 0; p < P; ++p) {
      const int term = split_term(nsplit, role_b, p);
      *reinterpret_cast<uint4*>(dst + (long long)p * Cp) = *reinterpret_cast<const uint4*>(t[term]);
    }
  }
}

// Fast path of cast_split: one bf16 plane, no dropout, C a multiple of 8 (== Cp), 16-byte aligned rows.  Streaming
// kernel: every thread converts UNR chunks of 8 floats, all 2*UNR 16-byte loads issued before the first use (the
// generic kernel above keeps only two loads in flight per thread and reaches ~30 % of the HBM bandwidth).
constexpr int CAST_UNR = 4;
__global__ void __launch_bounds__(256) cast_bf16_stream_kernel(const float* __restrict__ in, long long ld_in, unsigned R,
                                                               unsigned groups, bf16* __restrict__ out, long long ld_out,
                                                               const float* __restrict__ row_scale) {
  const unsigned total = R * groups, T = gridDim.x * blockDim.x;
  for (unsigned base = blockIdx.x * blockDim.x + threadIdx.x; base < total; base += CAST_UNR * T) {
    float4 a[CAST_UNR], b[CAST_UNR];
    unsigned r[CAST_UNR], g[CAST_UNR];
#pragma unroll
    for (int k = 0; k < CAST_UNR; ++k) {
      const unsigned i = base + k * T;
      r[k] = i / groups; g[k] = i - r[k] * groups;
      if (i < total) {
        const float4* src = reinterpret_cast<const float4*>(in + (long long)r[k] * ld_in + g[k] * 8);
        a[k] = __ldcs(src); b[k] = __ldcs(src + 1);          // streamed once: evict-first
      }
    }
#pragma unroll
    for (int k = 0; k < CAST_UNR; ++k) {
      const unsigned i = base + k * T;
      if (i < total) {
        const float sc = row_scale ? __ldg(row_scale + r[k]) : 1.f;
        __align__(16) __nv_bfloat162 o[4];
        o[0] = __floats2bfloat162_rn(a[k].x * sc, a[k].y * sc); o[1] = __floats2bfloat162_rn(a[k].z * sc, a[k].w * sc);
        o[2] = __floats2bfloat162_rn(b[k].x * sc, b[k].y * sc); o[3] = __floats2bfloat162_rn(b[k].z * sc, b[k].w * sc);
        *reinterpret_cast<uint4*>(out + (long long)r[k] * ld_out + g[k] * 8) = *reinterpret_cast<const uint4*>(o);
      }
    }
  }
}

// K-blocked copy of operand planes: [R, ld] (ld a multiple of 64) -> [ld / 64][R][64], so that the 64-column block of any
// row range is ONE contiguous run of memory.  A weight matrix that is streamed from HBM launch after launch (decoding: the
// W_v planes, 70 MB per step) is then fetched in multi-KB runs instead of one 128-byte piece per row.
__global__ void __launch_bounds__(256) block_planes_kernel(const uint4* __restrict__ in, long long groups, long long R,
                                                           uint4* __restrict__ out) {
  const long long total = R * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups, g = i - r * groups;
    out[((g >> 3) * R + r) * 8 + (g & 7)] = in[i];
  }
}
int block_planes(const bf16* in, long long ld, int R, bf16* out, cudaStream_t st) {
  PVCR_REQUIRE(ld % 64 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
               "block_planes: ld=%lld must be a multiple of 64 and the buffers 16-byte aligned", ld);
  if (R == 0) return PVCR_OK;
  const long long total = (long long)R * (ld / 8);
  const int blocks = (int)((total + 255) / 256 > sm_count() * 16 ? sm_count() * 16 : (total + 255) / 256);
  { LaunchScope ls_(KC_STAGE, st);
  block_planes_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(in), ld / 8, R, reinterpret_cast<uint4*>(out));
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

int cast_split(const float* in, long long ld_in, int R, int C, bf16* out, long long ld_out, int Cp, int nsplit,
               int role_b, const float* row_scale, Dropout drop, cudaStream_t st) {
  PVCR_REQUIRE(Cp % 8 == 0 && Cp >= C && ld_out % 8 == 0, "cast_split: bad padding Cp=%d C=%d ld_out=%lld", Cp, C, ld_out);
  if (R == 0) return PVCR_OK;
  static const bool fast_off = getenv("PVCR_NO_FAST_CAST") != nullptr;      // A/B knob
  if (!fast_off && nsplit == 1 && drop.p <= 0.f && C == Cp && (ld_in & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
      (long long)R * (Cp / 8) < (1ll << 31)) {
    const unsigned groups = (unsigned)(Cp / 8);
    const long long total = (long long)R * groups;
    const long long want = (total + 256 * CAST_UNR - 1) / (256 * CAST_UNR);
    const int blocks = (int)(want > sm_count() * 8 ? sm_count() * 8 : want);
    { LaunchScope ls_(KC_STAGE, st);
    cast_bf16_stream_kernel<<<blocks, 256, 0, st>>>(in, ld_in, (unsigned)R, groups, out, ld_out, row_scale);
    }
    PVCR_CUDA_CHECK(cudaGetLastError());
    return PVCR_OK;
  }
  const long long total = (long long)R * (Cp / 8);
  const int blocks = (int)((total + 255) / 256 > sm_count() * 16 ? sm_count() * 16 : (total + 255) / 256);
  { LaunchScope ls_(KC_STAGE, st);
  cast_split_kernel<<<blocks, 256, 0, st>>>(in, ld_in, R, C, out, ld_out, Cp, nsplit, role_b, row_scale, drop);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

__global__ void transpose_split_kernel(const float* __restrict__ in, long long ld_in, int R, int C,
                                       bf16* __restrict__ out, long long ld_out, int Rp, int r_off, int r_end,
                                       int nsplit, int role_b, const long long* __restrict__ row_ids,
                                       const float* __restrict__ row_scale, Dropout drop) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < R && c < C) {
      const long long src_row = row_ids ? row_ids[r] : r;
      v = __ldg(in + src_row * ld_in + c);
      if (row_scale) v *= __ldg(row_scale + src_row);
      if (drop.p > 0.f) v *= dropout_scale(drop, drop_index(drop, r, C, c));
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < r_end) write_split(out + (long long)c * ld_out, Rp, r_off + r, nsplit, role_b, tile[threadIdx.x][i]);
  }
}

// Fast path (single plane, no gather / dropout): 64 x 64 tiles, 16-byte loads along the input rows, fp32 tile in
// shared memory, 16-byte bf16 stores along the output rows.
__global__ void __launch_bounds__(256) transpose_cast_kernel(const float* __restrict__ in, long long ld_in, int R, int C,
                                                             bf16* __restrict__ out, long long ld_out, int r_off,
                                                             int r_end, const float* __restrict__ row_scale) {
  __shared__ float tile[64][65];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64, t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + (t >> 4) + 16 * i, c = c0 + (t & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < R) {
      const float* src = in + (long long)r * ld_in + c;
      if (c + 3 < C) v = __ldg(reinterpret_cast<const float4*>(src));
      else {
        if (c < C) v.x = __ldg(src);
        if (c + 1 < C) v.y = __ldg(src + 1);
        if (c + 2 < C) v.z = __ldg(src + 2);
      }
      if (row_scale) { const float s = __ldg(row_scale + r); v.x *= s; v.y *= s; v.z *= s; v.w *= s; }
    }
    float* d = &tile[(t >> 4) + 16 * i][(t & 15) * 4];
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int cl = (t >> 3) + 32 * i, rl = (t & 7) * 8;
    const int c = c0 + cl, r = r0 + rl;
    if (c < C && r < r_end) {
      __align__(16) __nv_bfloat162 o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = __floats2bfloat162_rn(tile[rl + 2 * k][cl], tile[rl + 2 * k + 1][cl]);
      bf16* dst = out + (long long)c * ld_out + r_off + r;
      if (r + 8 <= r_end) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(o);
      else {
        const bf16* ob = reinterpret_cast<const bf16*>(o);
        for (int k = 0; k < 8 && r + k < r_end; ++k) dst[k] = ob[k];
      }
    }
  }
}

int transpose_split(const float* in, long long ld_in, int R, int C, bf16* out, long long ld_out, int Rp, int r_off,
                    int zero_pad, int nsplit, int role_b, const long long* row_ids, const float* row_scale,
                    cudaStream_t st, Dropout drop) {
  PVCR_REQUIRE(r_off + R <= Rp, "transpose_split: r_off=%d R=%d exceed Rp=%d", r_off, R, Rp);
  if (C == 0) return PVCR_OK;
  const int r_end = zero_pad ? Rp - r_off : R;       // rows >= R read as zero
  if (r_end == 0) return PVCR_OK;
  const bool fast = nsplit == 1 && !row_ids && drop.p <= 0.f && ld_in % 4 == 0 && r_off % 8 == 0 && ld_out % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (fast) {
    dim3 grid(cdiv(r_end, 64), cdiv(C, 64));
    { LaunchScope ls_(KC_STAGE, st);
    transpose_cast_kernel<<<grid, 256, 0, st>>>(in, ld_in, R, C, out, ld_out, r_off, r_end, row_scale);
    }
    PVCR_CUDA_CHECK(cudaGetLastError());
    return PVCR_OK;
  }
  dim3 grid(cdiv(r_end, 32), cdiv(C, 32));
  { LaunchScope ls_(KC_STAGE, st);
  transpose_split_kernel<<<grid, dim3(32, 8), 0, st>>>(in, ld_in, R, C, out, ld_out, Rp, r_off, r_end, nsplit, role_b,
                                                       row_ids, row_scale, drop);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

__global__ void gather_split_kernel(const float* __restrict__ table, int E, const long long* __restrict__ ids,
                                    int n_ids, bf16* __restrict__ out, long long ld_out, int Ep, int nsplit,
                                    Dropout drop) {
  const long long total = (long long)n_ids * Ep;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / Ep), e = (int)(i % Ep);
    float val = 0.f;
    if (e < E) {
      val = __ldg(table + ids[r] * E + e);
      if (drop.p > 0.f) val *= dropout_scale(drop, drop_index(drop, r, E, e));
    }
    write_split(out + (long long)r * ld_out, Ep, e, nsplit, 0, val);
  }
}
int gather_split(const float* table, int E, const long long* ids, int n_ids, bf16* out, long long ld_out, int Ep,
                 int nsplit, Dropout drop, cudaStream_t st) {
  if (n_ids == 0) return PVCR_OK;
  const long long total = (long long)n_ids * Ep;
  const int blocks = (int)((total + 255) / 256 > sm_count() * 8 ? sm_count() * 8 : (total + 255) / 256);
  { LaunchScope ls_(KC_STAGE, st);
  gather_split_kernel<<<blocks, 256, 0, st>>>(table, E, ids, n_ids, out, ld_out, Ep, nsplit, drop);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

__global__ void scatter_add_rows_kernel(const float* __restrict__ rows, long long ld, const long long* __restrict__ ids,
                                        int n_ids, int E, float* __restrict__ grad, Dropout drop) {
  const long long total = (long long)n_ids * E;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / E), e = (int)(i % E);
    float v = rows[(long long)r * ld + e];
    if (drop.p > 0.f) v *= dropout_scale(drop, drop_index(drop, r, E, e));
    atomicAdd(grad + ids[r] * E + e, v);
  }
}
int scatter_add_rows(const float* rows, long long ld, const long long* ids, int n_ids, int E, float* table_grad,
                     Dropout drop, cudaStream_t st) {
  if (n_ids == 0) return PVCR_OK;
  const long long total = (long long)n_ids * E;
  const int blocks = (int)((total + 255) / 256 > sm_count() * 8 ? sm_count() * 8 : (total + 255) / 256);
  { LaunchScope ls_(KC_MISC, st);
  scatter_add_rows_kernel<<<blocks, 256, 0, st>>>(rows, ld, ids, n_ids, E, table_grad, drop);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

// column sums: block = 32 columns x 32 row-lanes; deterministic (fixed tree, no atomics)
__global__ void colsum_kernel(const float* __restrict__ in, long long ld, int R, int C, float* __restrict__ out,
                              int accumulate) {
  __shared__ float part[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < C)
    for (int r = threadIdx.y; r < R; r += 32) acc += in[(long long)r * ld + c];
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += part[i][threadIdx.x];
    out[c] = accumulate ? out[c] + s : s;
  }
}
int colsum(const float* in, long long ld, int R, int C, float* out, int accumulate, cudaStream_t st) {
  if (C == 0) return PVCR_OK;
  { LaunchScope ls_(KC_MISC, st);
  colsum_kernel<<<cdiv(C, 32), dim3(32, 32), 0, st>>>(in, ld, R, C, out, accumulate);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

__global__ void dropout_apply_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, Dropout drop) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = in[i] * dropout_scale(drop, (unsigned long long)i);
}
int dropout_apply(const float* in, float* out, long long n, Dropout drop, cudaStream_t st) {
  if (n == 0) return PVCR_OK;
  const int blocks = (int)((n + 255) / 256 > sm_count() * 8 ? sm_count() * 8 : (n + 255) / 256);
  { LaunchScope ls_(KC_MISC, st);
  dropout_apply_kernel<<<blocks, 256, 0, st>>>(in, out, n, drop);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ in, long long ld, int C,
                                                         long long* out, long long out_stride, long long* next,
                                                         const long long* teacher, long long teacher_stride,
                                                         int use_teacher) {
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  const int row = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* x = in + (long long)row * ld;
  float m = -INFINITY;
  int mi = 0x7fffffff;
  for (int j = tid; j < C; j += 256) {
    const float v = x[j];
    if (v > m) { m = v; mi = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
  }
  if (lane == 0) { s_val[warp] = m; s_idx[warp] = mi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w)
      if (s_val[w] > m || (s_val[w] == m && s_idx[w] < mi)) { m = s_val[w]; mi = s_idx[w]; }
    if (mi == 0x7fffffff) mi = 0;                      // all-NaN row: torch.argmax returns an index too
    out[(long long)row * out_stride] = mi;
    if (next) next[row] = use_teacher ? teacher[(long long)row * teacher_stride] : (long long)mi;
  }
}
int argmax_rows(const float* in, long long ld, int R, int C, long long* out, long long out_stride, long long* next,
                const long long* teacher, long long teacher_stride, int use_teacher, cudaStream_t st) {
  if (R == 0) return PVCR_OK;
  { LaunchScope ls_(KC_LOSS, st);
  argmax_rows_kernel<<<R, 256, 0, st>>>(in, ld, C, out, out_stride, next, teacher, teacher_stride, use_teacher);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}
__global__ void fill_i64_kernel(long long* p, long long v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
int fill_i64(long long* p, long long v, int n, cudaStream_t st) {
  if (n == 0) return PVCR_OK;
  { LaunchScope ls_(KC_MISC, st);
  fill_i64_kernel<<<cdiv(n, 256), 256, 0, st>>>(p, v, n);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}
__global__ void add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += x[i];
}
int add_inplace(float* y, const float* x, long long n, cudaStream_t st) {
  if (n <= 0) return PVCR_OK;
  add_inplace_kernel<<<(int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, st>>>(y, x, n);
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}
int fill_zero(void* p, size_t bytes, cudaStream_t st) {
  if (bytes) PVCR_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, st));
  return PVCR_OK;
}

// ------------------------------------------------------------------------------------------------
// GRU gates (torch gate order r,z,n):  r = s(gi_r+gh_r)  z = s(gi_z+gh_z)  n = tanh(gi_n + r*gh_n)
//                                      h' = (1-z)*n + z*h
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gru_gate_fwd_kernel(GruFwdArgs a) {
  __shared__ long long s_word[9];                      // H >= 32 (host check): a block of 256 threads touches <= 9 videos
  pdl_launch_dependents();                             // a programmatically launched successor may set itself up now
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int b_first = 0;
  if (a.am_pmax) {
    // fed-back word of every video this block touches: warp w combines the partial arg-max of video b_first + w, ...
    // (same rule as argmax_parts_kernel: larger value, then lower index; an all-NaN row gives index 0)
    b_first = (int)(((long long)blockIdx.x * blockDim.x) / a.H);
    int b_last = (int)(((long long)blockIdx.x * blockDim.x + blockDim.x - 1) / a.H);
    if (b_last >= a.B) b_last = a.B - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int bb = b_first + warp; bb <= b_last; bb += 8) {
      float m = -INFINITY;
      int mi = 0x7fffffff;
      for (int k = lane; k < a.am_nparts; k += 32) {
        const float v = a.am_pmax[(long long)bb * a.am_nparts + k];
        const int i = a.am_pidx[(long long)bb * a.am_nparts + k];
        if (v > m || (v == m && i < mi)) { m = v; mi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
      }
      if (lane == 0) {
        if (mi == 0x7fffffff) mi = 0;
        s_word[bb - b_first] = mi;
        // the block that holds unit 0 of the video publishes the id
        if ((long long)bb * a.H >= (long long)blockIdx.x * blockDim.x && a.am_out) a.am_out[(long long)bb * a.am_out_stride] = mi;
      }
    }
    __syncthreads();
  }
  if (idx >= a.B * a.H) return;
  const int b = idx / a.H, j = idx % a.H, H = a.H;
  float gi[3], gh[3];
  const long long rb = a.am_pmax ? s_word[b - b_first] : (a.gi_b_rows ? a.gi_b_rows[b] : (long long)b);
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    float x = a.gi_a ? a.gi_a[(long long)b * a.gi_a_ld + g * H + j] : 0.f;
    if (a.gi_b) x += a.gi_b[rb * a.gi_b_ld + g * H + j];
    if (a.gi_bias) x += a.gi_bias[g * H + j];
    gi[g] = x;
    gh[g] = a.b_hh[g * H + j] + (a.gh ? a.gh[(long long)b * a.gh_ld + g * H + j] : 0.f);
  }
  const float hp = a.h_prev ? a.h_prev[(long long)b * a.h_prev_ld + j] : 0.f;
  const float r = 1.f / (1.f + expf(-(gi[0] + gh[0])));
  const float z = 1.f / (1.f + expf(-(gi[1] + gh[1])));
  const float n = tanhf(gi[2] + r * gh[2]);
  const float h = (1.f - z) * n + z * hp;
  a.h_out[(long long)b * a.h_out_ld + j] = h;
  if (a.h_planes) write_split(a.h_planes + (long long)b * a.h_planes_ld, a.Hp, j, a.nsplit, 0, h);
  if (a.r) {
    a.r[idx] = r; a.z[idx] = z; a.n[idx] = n; a.ghn[idx] = gh[2];
  }
}
int gru_gate_fwd(const GruFwdArgs& a, cudaStream_t st) {
  const int total = a.B * a.H;
  if (total == 0) return PVCR_OK;
  PVCR_REQUIRE(!a.am_pmax || (a.H >= 32 && a.gi_b && a.am_pidx && a.am_nparts > 0),
               "gru_gate_fwd: folded arg-max needs H >= 32, the word table and the partials (H=%d)", a.H);
  { LaunchScope ls_(KC_GATE, st);
  gru_gate_fwd_kernel<<<cdiv(total, 256), 256, 0, st>>>(a);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

__global__ void gru_gate_bwd_kernel(GruBwdArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.B * a.H) return;
  const int b = idx / a.H, j = idx % a.H, H = a.H;
  float dh = 0.f;
  if (a.dh_a) dh += a.dh_a[(long long)b * a.dh_a_ld + j];
  if (a.dh_b) dh += a.dh_b[(long long)b * a.dh_b_ld + j];
  const float r = a.r[idx], z = a.z[idx], n = a.n[idx], ghn = a.ghn[idx];
  const float hp = a.h_prev ? a.h_prev[(long long)b * a.h_prev_ld + j] : 0.f;
  const float dn = dh * (1.f - z), dz = dh * (hp - n);
  const float dnp = dn * (1.f - n * n);
  const float dzp = dz * z * (1.f - z);
  const float drp = dnp * ghn * r * (1.f - r);
  const float dghn = dnp * r;
  float* dgi = a.dgi + (long long)b * a.dgi_ld;
  float* dgh = a.dgh + (long long)b * a.dgh_ld;
  dgi[j] = drp; dgi[H + j] = dzp; dgi[2 * H + j] = dnp;
  dgh[j] = drp; dgh[H + j] = dzp; dgh[2 * H + j] = dghn;
  if (a.dgi_planes) {
    bf16* o = a.dgi_planes + (long long)b * a.dgi_planes_ld;
    write_split(o, a.dgi_Kp, a.dgi_col0 + j, a.nsplit, 0, drp);
    write_split(o, a.dgi_Kp, a.dgi_col0 + H + j, a.nsplit, 0, dzp);
    write_split(o, a.dgi_Kp, a.dgi_col0 + 2 * H + j, a.nsplit, 0, dnp);
  }
  if (a.dgh_planes) {
    bf16* o = a.dgh_planes + (long long)b * a.dgh_planes_ld;
    write_split(o, a.dgh_Kp, a.dgh_col0 + j, a.nsplit, 0, drp);
    write_split(o, a.dgh_Kp, a.dgh_col0 + H + j, a.nsplit, 0, dzp);
    write_split(o, a.dgh_Kp, a.dgh_col0 + 2 * H + j, a.nsplit, 0, dghn);
  }
  a.dh_direct[(long long)b * a.dh_direct_ld + j] = dh * z;
}
int gru_gate_bwd(const GruBwdArgs& a, cudaStream_t st) {
  const int total = a.B * a.H;
  if (total == 0) return PVCR_OK;
  { LaunchScope ls_(KC_GATE, st);
  gru_gate_bwd_kernel<<<cdiv(total, 256), 256, 0, st>>>(a);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

// ------------------------------------------------------------------------------------------------
// LSTM gates (torch gate order i,f,g,o):  c' = s(f)*c + s(i)*tanh(g)   h' = s(o)*tanh(c')
// ------------------------------------------------------------------------------------------------
__global__ void lstm_gate_fwd_kernel(LstmFwdArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.B * a.H) return;
  const int b = idx / a.H, j = idx % a.H, H = a.H;
  float x[4];
#pragma unroll
  for (int g = 0; g < 4; ++g)
    x[g] = a.gi[(long long)b * a.gi_ld + g * H + j] + a.b_hh[g * H + j] +
           (a.gh ? a.gh[(long long)b * a.gh_ld + g * H + j] : 0.f);
  const float i = 1.f / (1.f + expf(-x[0])), f = 1.f / (1.f + expf(-x[1]));
  const float g = tanhf(x[2]), o = 1.f / (1.f + expf(-x[3]));
  const float cp = a.c_prev ? a.c_prev[idx] : 0.f;
  const float c = f * cp + i * g;
  const float h = o * tanhf(c);
  a.h_out[(long long)b * a.h_out_ld + j] = h;
  if (a.h_planes) write_split(a.h_planes + (long long)b * a.h_planes_ld, a.Hp, j, a.nsplit, 0, h);
  a.i[idx] = i; a.f[idx] = f; a.g[idx] = g; a.o[idx] = o; a.c[idx] = c;
}
int lstm_gate_fwd(const LstmFwdArgs& a, cudaStream_t st) {
  const int total = a.B * a.H;
  if (total == 0) return PVCR_OK;
  { LaunchScope ls_(KC_GATE, st);
  lstm_gate_fwd_kernel<<<cdiv(total, 256), 256, 0, st>>>(a);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

__global__ void lstm_gate_bwd_kernel(LstmBwdArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.B * a.H) return;
  const int b = idx / a.H, j = idx % a.H, H = a.H;
  float dh = 0.f;
  if (a.dh_a) dh += a.dh_a[(long long)b * a.dh_a_ld + j];
  if (a.dh_b) dh += a.dh_b[(long long)b * a.dh_b_ld + j];
  const float i = a.i[idx], f = a.f[idx], g = a.g[idx], o = a.o[idx];
  const float tc = tanhf(a.c[idx]);
  const float cp = a.c_prev ? a.c_prev[idx] : 0.f;
  const float dO = dh * tc;
  const float dc = a.dc[idx] + dh * o * (1.f - tc * tc);
  float d[4];
  d[0] = dc * g * i * (1.f - i);
  d[1] = dc * cp * f * (1.f - f);
  d[2] = dc * i * (1.f - g * g);
  d[3] = dO * o * (1.f - o);
  a.dc[idx] = dc * f;
  float* da = a.da + (long long)b * a.da_ld;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    da[q * H + j] = d[q];
    if (a.da_planes) write_split(a.da_planes + (long long)b * a.da_planes_ld, a.Kp, q * H + j, a.nsplit, 0, d[q]);
  }
}
int lstm_gate_bwd(const LstmBwdArgs& a, cudaStream_t st) {
  const int total = a.B * a.H;
  if (total == 0) return PVCR_OK;
  { LaunchScope ls_(KC_GATE, st);
  lstm_gate_bwd_kernel<<<cdiv(total, 256), 256, 0, st>>>(a);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

// ------------------------------------------------------------------------------------------------
// additive attention step: one CTA per video.
//   scores[n] = v . tanh(q + pk[n]);  alpha = softmax_n(scores) (unmasked);  ctx = sum_n alpha[n] enc[n]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_fwd_kernel(AttnFwdArgs a) {
  extern __shared__ float sm[];
  const int H = a.H, N = a.N, b = blockIdx.x;
  float* sq = sm;            // [H]
  float* sv = sm + H;        // [H]
  float* ss = sm + 2 * H;    // [N]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  for (int h = tid; h < H; h += blockDim.x) {
    sq[h] = a.q[(long long)b * a.q_ld + h];
    sv[h] = a.v[h];
  }
  __syncthreads();
  const float* pk = a.pk + (long long)b * N * H;
  const float* enc = a.enc + (long long)b * N * H;
  for (int n = warp; n < N; n += nw) {
    float acc = 0.f;
    for (int h = lane; h < H; h += 32) acc += sv[h] * tanhf(sq[h] + pk[(long long)n * H + h]);
    acc = warp_sum(acc);
    if (lane == 0) ss[n] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    float m = -INFINITY;
    for (int n = lane; n < N; n += 32) m = fmaxf(m, ss[n]);
    m = warp_max(m);
    float s = 0.f;
    for (int n = lane; n < N; n += 32) {
      const float e = expf(ss[n] - m);
      ss[n] = e;
      s += e;
    }
    s = warp_sum(s);
    for (int n = lane; n < N; n += 32) {
      const float al = ss[n] / s;
      ss[n] = al;
      a.alpha[(long long)b * N + n] = al;
    }
  }
  __syncthreads();
  for (int h = tid; h < H; h += blockDim.x) {
    float c = 0.f;
    for (int n = 0; n < N; ++n) c += ss[n] * enc[(long long)n * H + h];
    a.ctx[(long long)b * a.ctx_ld + h] = c;
    if (a.ctx_planes) write_split(a.ctx_planes + (long long)b * a.ctx_planes_ld, a.Hp, h, a.nsplit, 0, c);
  }
}
// Same step with thread = (8 consecutive dims, frame group): every thread issues the 128-bit loads of all its frames
// before it uses them (one memory round trip per phase instead of one per frame), scores are reduced by warp shuffles.
// The step-wise decoders (greedy, beam search, split-precision training) were spending 50 of their ~195 us per step in
// the kernel above.  Accurate tanhf / expf as above: this path serves the fp32-equivalent modes.
template <int NF, bool PROJ = false>
__global__ void __launch_bounds__(256) attn_fwd_vec_kernel(AttnFwdArgs a) {
  extern __shared__ float sm[];
  const int H = a.H, N = a.N, b = blockIdx.x;
  const int DG = H >> 3, FG = 256 / DG, WPF = DG >> 5;      // dim groups, frame groups, warps per frame group (DG % 32 == 0)
  float* sP = sm;                 // [N][WPF]
  float* ss = sP + N * WPF;       // [N]
  float* sC = ss + N;             // [FG][H]
  const int tid = threadIdx.x, lane = tid & 31, dg = tid % DG, fg = tid / DG, d0 = dg * 8;
  const float4* q4 = reinterpret_cast<const float4*>(a.q + (long long)b * a.q_ld + d0);
  const float4* v4 = reinterpret_cast<const float4*>(a.v + d0);
  const float4 va = v4[0], vb = v4[1];
  const float v8[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
  const float* pk = a.pk + (long long)b * N * H + d0;
  const float* enc = a.enc + (long long)b * N * H + d0;
  float4 x[NF][2];
#pragma unroll
  for (int m = 0; m < NF; ++m) {
    const int n = fg + FG * m;
    x[m][0] = x[m][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) {
      const float4* p4 = reinterpret_cast<const float4*>(pk + (long long)n * H);
      x[m][0] = __ldg(p4); x[m][1] = __ldg(p4 + 1);
    }
  }
  // the keys and v do not depend on the step: when launched programmatically (decode loop) they were fetched while the
  // predecessor, which produces the query, was still running
  pdl_wait();
  pdl_launch_dependents();
  const float4 qa = q4[0], qb = q4[1];
  const float q8[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
  for (int m = 0; m < NF; ++m) {
    const int n = fg + FG * m;
    const float p8[8] = {x[m][0].x, x[m][0].y, x[m][0].z, x[m][0].w, x[m][1].x, x[m][1].y, x[m][1].z, x[m][1].w};
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v8[e] * tanhf(q8[e] + p8[e]);
    s = warp_sum(s);
    if (lane == 0 && n < N) sP[n * WPF + (dg >> 5)] = s;
  }
  // the encoder rows of the context phase are independent of the softmax: fetch them now
  if (!PROJ) {
#pragma unroll
    for (int m = 0; m < NF; ++m) {
      const int n = fg + FG * m;
      x[m][0] = x[m][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < N) {
        const float4* e4 = reinterpret_cast<const float4*>(enc + (long long)n * H);
        x[m][0] = __ldg(e4); x[m][1] = __ldg(e4 + 1);
      }
    }
  }
  __syncthreads();
  if (tid < 32) {
    float mx = -INFINITY;
    for (int n = lane; n < N; n += 32) {
      float s = 0.f;
      for (int w = 0; w < WPF; ++w) s += sP[n * WPF + w];
      ss[n] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float den = 0.f;
    for (int n = lane; n < N; n += 32) {
      const float e = expf(ss[n] - mx);
      ss[n] = e;
      den += e;
    }
    den = warp_sum(den);
    for (int n = lane; n < N; n += 32) {
      const float al = ss[n] / den;
      ss[n] = al;
      a.alpha[(long long)b * N + n] = al;
    }
  }
  __syncthreads();
  if (PROJ) {
    // out[b, :] = sum_n alpha_n val[b, n, :]: a thread owns 4 consecutive output columns, 8 frames' 16-byte loads in flight.
    // (Measured alternative: items of (4 columns, half of the frames) with 20 loads in flight -- 14.9 instead of 16.9 us alone,
    // but 146 registers per thread: one CTA per SM, and the 128 CTAs no longer fit the SMs the vocabulary GEMM running next
    // to this kernel leaves free: 1.95 instead of 1.77 ms per batch.)
    const int W4 = a.W >> 2;
    const float4* val = reinterpret_cast<const float4*>(a.val + (long long)b * N * a.W);
    for (int c4 = tid; c4 < W4; c4 += 256) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int n0 = 0; n0 < N; n0 += 8) {
        float4 y[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = n0 + k < N ? __ldg(val + (long long)(n0 + k) * W4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float al = n0 + k < N ? ss[n0 + k] : 0.f;
          acc.x += al * y[k].x; acc.y += al * y[k].y; acc.z += al * y[k].z; acc.w += al * y[k].w;
        }
      }
      *reinterpret_cast<float4*>(a.out + (long long)b * a.out_ld + 4 * c4) = acc;
    }
    return;
  }
  float c8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int m = 0; m < NF; ++m) {
    const int n = fg + FG * m;
    if (n < N) {
      const float al = ss[n];
      c8[0] += al * x[m][0].x; c8[1] += al * x[m][0].y; c8[2] += al * x[m][0].z; c8[3] += al * x[m][0].w;
      c8[4] += al * x[m][1].x; c8[5] += al * x[m][1].y; c8[6] += al * x[m][1].z; c8[7] += al * x[m][1].w;
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) sC[fg * H + d0 + e] = c8[e];
  __syncthreads();
  for (int h = tid; h < H; h += 256) {
    float c = 0.f;
    for (int f = 0; f < FG; ++f) c += sC[f * H + h];
    a.ctx[(long long)b * a.ctx_ld + h] = c;
    if (a.ctx_planes) write_split(a.ctx_planes + (long long)b * a.ctx_planes_ld, a.Hp, h, a.nsplit, 0, c);
  }
}

bool attn_fwd_projected_ok(int N, int H, int W) {
  const int DG = H >> 3, FG = DG > 0 && DG <= 256 ? 256 / DG : 0;
  return H % 256 == 0 && H <= 2048 && FG >= 1 && (N + FG - 1) / FG <= 10 && W % 4 == 0;
}

int attn_fwd(const AttnFwdArgs& a, cudaStream_t st) {
  if (a.B == 0) return PVCR_OK;
  if (a.val) {
    PVCR_REQUIRE(attn_fwd_projected_ok(a.N, a.H, a.W) && a.out && a.out_ld % 4 == 0 && a.q_ld % 4 == 0 &&
                     ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.pk) | reinterpret_cast<uintptr_t>(a.val) |
                       reinterpret_cast<uintptr_t>(a.out) | reinterpret_cast<uintptr_t>(a.v)) & 15) == 0,
                 "attn_fwd (projected values): shape N=%d H=%d W=%d / alignment not supported", a.N, a.H, a.W);
    const int DG = a.H >> 3, FG = 256 / DG, nf = (a.N + FG - 1) / FG;
    const size_t smem = ((size_t)a.N * (DG >> 5) + a.N + (size_t)FG * a.H) * sizeof(float);
    PVCR_REQUIRE(smem <= 48 * 1024, "attn_fwd: H=%d N=%d needs %zu B of shared memory", a.H, a.N, smem);
    { LaunchScope ls_(KC_ATTN, st);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(a.B); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    if (pdl_enabled()) { cfg.attrs = at; cfg.numAttrs = 1; }
    if (nf <= 5) PVCR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_fwd_vec_kernel<5, true>, a));
    else PVCR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_fwd_vec_kernel<10, true>, a));
    }
    PVCR_CUDA_CHECK(cudaGetLastError());
    return PVCR_OK;
  }
  // vectorised variant: H a multiple of 256 up to 2048 (whole warps per frame group), aligned rows, <= 10 frames per group
  const int DG = a.H >> 3, FG = DG > 0 && DG <= 256 ? 256 / DG : 0;
  static const bool vec_off = getenv("PVCR_NO_VEC_ATTN") != nullptr;
  const bool vec = !vec_off && a.H % 256 == 0 && a.H <= 2048 && FG >= 1 && (a.N + FG - 1) / FG <= 10 && a.q_ld % 4 == 0 &&
                   ((reinterpret_cast<uintptr_t>(a.q) | reinterpret_cast<uintptr_t>(a.pk) | reinterpret_cast<uintptr_t>(a.enc) |
                     reinterpret_cast<uintptr_t>(a.v)) & 15) == 0;
  if (vec) {
    const int nf = (a.N + FG - 1) / FG;
    const size_t smem = ((size_t)a.N * (DG >> 5) + a.N + (size_t)FG * a.H) * sizeof(float);
    PVCR_REQUIRE(smem <= 48 * 1024, "attn_fwd: H=%d N=%d needs %zu B of shared memory", a.H, a.N, smem);
    { LaunchScope ls_(KC_ATTN, st);
    if (nf <= 5) attn_fwd_vec_kernel<5><<<a.B, 256, smem, st>>>(a);
    else attn_fwd_vec_kernel<10><<<a.B, 256, smem, st>>>(a);
    }
    PVCR_CUDA_CHECK(cudaGetLastError());
    return PVCR_OK;
  }
  const size_t smem = (size_t)(2 * a.H + a.N) * sizeof(float);
  PVCR_REQUIRE(smem <= 48 * 1024, "attn_fwd: H=%d N=%d needs %zu B of shared memory", a.H, a.N, smem);
  { LaunchScope ls_(KC_ATTN, st);
  attn_fwd_kernel<<<a.B, 256, smem, st>>>(a);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

__global__ void __launch_bounds__(256) attn_bwd_kernel(AttnBwdArgs a) {
  extern __shared__ float sm[];
  const int H = a.H, N = a.N, b = blockIdx.x;
  float* sq = sm;              // [H]
  float* sv = sm + H;          // [H]
  float* sd = sm + 2 * H;      // [H] dctx
  float* sa = sm + 3 * H;      // [N] alpha
  float* sds = sm + 3 * H + N; // [N] da -> dscore
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  for (int h = tid; h < H; h += blockDim.x) {
    sq[h] = a.q[(long long)b * a.q_ld + h];
    sv[h] = a.v[h];
    sd[h] = a.dctx[(long long)b * a.dctx_ld + h];
  }
  for (int n = tid; n < N; n += blockDim.x) sa[n] = a.alpha[(long long)b * N + n];
  __syncthreads();
  const float* pk = a.pk + (long long)b * N * H;
  const float* enc = a.enc + (long long)b * N * H;
  float* dpk = a.dpk + (long long)b * N * H;
  float* denc = a.denc + (long long)b * N * H;
  for (int n = warp; n < N; n += nw) {
    float acc = 0.f;
    const float al = sa[n];
    for (int h = lane; h < H; h += 32) {
      acc += sd[h] * enc[(long long)n * H + h];
      denc[(long long)n * H + h] += al * sd[h];
    }
    acc = warp_sum(acc);
    if (lane == 0) sds[n] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    float dot = 0.f;
    for (int n = lane; n < N; n += 32) dot += sa[n] * sds[n];
    dot = warp_sum(dot);
    for (int n = lane; n < N; n += 32) sds[n] = sa[n] * (sds[n] - dot);
  }
  __syncthreads();
  for (int h = tid; h < H; h += blockDim.x) {
    float dq = 0.f, dv = 0.f;
    const float qh = sq[h], vh = sv[h];
    for (int n = 0; n < N; ++n) {
      const float e = tanhf(qh + pk[(long long)n * H + h]);
      const float de = sds[n] * vh * (1.f - e * e);
      dpk[(long long)n * H + h] += de;
      dq += de;
      dv += sds[n] * e;
    }
    a.dq[(long long)b * a.dq_ld + h] = dq;
    if (a.dq_planes) write_split(a.dq_planes + (long long)b * a.dq_planes_ld, a.dq_Kp, h, a.nsplit, 0, dq);
    a.dv_part[(long long)b * H + h] += dv;
  }
}
int attn_bwd(const AttnBwdArgs& a, cudaStream_t st) {
  if (a.B == 0) return PVCR_OK;
  const size_t smem = (size_t)(3 * a.H + 2 * a.N) * sizeof(float);
  PVCR_REQUIRE(smem <= 48 * 1024, "attn_bwd: H=%d N=%d needs %zu B of shared memory", a.H, a.N, smem);
  { LaunchScope ls_(KC_ATTN, st);
  attn_bwd_kernel<<<a.B, 256, smem, st>>>(a);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

// ------------------------------------------------------------------------------------------------
// cross entropy on materialised logits (generic-precision path and the logits-returning module API)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ logits, long long ld, int B, int L,
                                                      int Vc, const long long* __restrict__ target,
                                                      const long long* __restrict__ s_len, float* lse_out,
                                                      float* nll_out, long long* pred_out, float* dlogits,
                                                      long long ld_d, const float* gscale) {
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  __shared__ float s_bcast[2];
  const int row = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* x = logits + (long long)row * ld;
  float m = -INFINITY;
  int mi = 0x7fffffff;
  for (int j = tid; j < Vc; j += 256) {
    const float v = x[j];
    if (v > m) { m = v; mi = j; }       // strided scan keeps the smallest index per thread on ties
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
  }
  if (lane == 0) { s_val[warp] = m; s_idx[warp] = mi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w)
      if (s_val[w] > m || (s_val[w] == m && s_idx[w] < mi)) { m = s_val[w]; mi = s_idx[w]; }
    s_bcast[0] = m;
    pred_out[row] = mi;
  }
  __syncthreads();
  m = s_bcast[0];
  float s = 0.f;
  for (int j = tid; j < Vc; j += 256) s += expf(x[j] - m);
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) s_val[warp] = s;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_val[w];
    const float lse = m + logf(t);
    s_bcast[1] = lse;
    lse_out[row] = lse;
    nll_out[row] = lse - x[target[row]];
  }
  __syncthreads();
  if (dlogits) {
    const float lse = s_bcast[1];
    const int b = row / L, l = row % L;
    const long long len = s_len[b];
    const float cnt = (float)(len < L ? len : L);      // mask.sum(dim=1) of the reference (train_utils.py:50-51)
    const float w = (l < len ? 1.f / (cnt * (float)B) : 0.f) * (gscale ? gscale[0] : 1.f);
    const long long t = target[row];
    float* d = dlogits + (long long)row * ld_d;
    for (int j = tid; j < Vc; j += 256) d[j] = (expf(x[j] - lse) - (j == t ? 1.f : 0.f)) * w;
  }
}
int ce_rows(const float* logits, long long ld, int B, int L, int Vc, const long long* target, const long long* s_len,
            float* lse, float* nll, long long* pred, float* dlogits, long long ld_d, const float* gscale,
            cudaStream_t st) {
  if (B * L == 0) return PVCR_OK;
  { LaunchScope ls_(KC_LOSS, st);
  ce_rows_kernel<<<B * L, 256, 0, st>>>(logits, ld, B, L, Vc, target, s_len, lse, nll, pred, dlogits, ld_d, gscale);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

__global__ void __launch_bounds__(256) loss_finalize_kernel(const float* nll, const long long* pred,
                                                            const long long* target, const long long* s_len, int B,
                                                            int L, float* out3) {
  __shared__ float red[3][256];
  float loss = 0.f, corr = 0.f, cnt = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) {
    const long long len = s_len[b];
    float s = 0.f;
    for (int l = 0; l < L && l < len; ++l) {
      s += nll[b * L + l];
      corr += (pred[b * L + l] == target[b * L + l]) ? 1.f : 0.f;
      cnt += 1.f;
    }
    loss += s / (float)(len < L ? len : L);     // mask.sum(dim=1) == min(s_len, L) (train_utils.py:50-51)
  }
  red[0][threadIdx.x] = loss; red[1][threadIdx.x] = corr; red[2][threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < 3; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out3[0] = red[0][0] / (float)B;
    out3[1] = red[1][0];
    out3[2] = red[2][0];
  }
}
int loss_finalize(const float* nll, const long long* pred, const long long* target, const long long* s_len, int B,
                  int L, float* out3, cudaStream_t st) {
  { LaunchScope ls_(KC_LOSS, st);
  loss_finalize_kernel<<<1, 256, 0, st>>>(nll, pred, target, s_len, B, L, out3);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

// ------------------------------------------------------------------------------------------------
// RationaleNet generator head: Linear(2H -> 2) + 2-class Gumbel-softmax + penalties
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gumbel_fwd_kernel(GumbelArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int B = a.B, N = a.N, H = a.H;
  if (warp >= B * N) return;
  const int b = warp / N, n = warp % N;
  const float* hf = a.hf + (long long)n * a.h_ts + (long long)b * a.h_bs;
  const float* hb = a.hb + (long long)n * a.h_ts + (long long)b * a.h_bs;
  float l0 = 0.f, l1 = 0.f;
  const unsigned long long e0 = (unsigned long long)warp * 2 * H;
  for (int k = lane; k < 2 * H; k += 32) {
    float x = k < H ? hf[k] : hb[k - H];
    if (a.drop.p > 0.f) x *= dropout_scale(a.drop, e0 + k);
    l0 += x * a.w[k];
    l1 += x * a.w[2 * H + k];
  }
  l0 = warp_sum(l0) + a.bias[0];
  l1 = warp_sum(l1) + a.bias[1];
  if (lane == 0) {
    float u0, u1;
    if (a.noise) { u0 = a.noise[warp * 2]; u1 = a.noise[warp * 2 + 1]; }
    else {  // Exp(1) draws
      const unsigned long long sd = stepped_seed(a.seed, a.seed_step);
      u0 = -logf(philox_uniform(sd, (unsigned long long)warp * 2));
      u1 = -logf(philox_uniform(sd, (unsigned long long)warp * 2 + 1));
    }
    const float y0 = (l0 - logf(u0)) / a.tau, y1 = (l1 - logf(u1)) / a.tau;
    const float m = fmaxf(y0, y1);
    const float e0f = expf(y0 - m), e1f = expf(y1 - m);
    const float s0 = e0f / (e0f + e1f), s1 = e1f / (e0f + e1f);
    a.y[warp * 2] = s0; a.y[warp * 2 + 1] = s1;
    if (a.hard) {     // straight-through value: (onehot - y) + y, argmax ties -> index 0
      const int idx = s1 > s0 ? 1 : 0;
      a.probs[warp * 2] = ((idx == 0 ? 1.f : 0.f) - s0) + s0;
      a.probs[warp * 2 + 1] = ((idx == 1 ? 1.f : 0.f) - s1) + s1;
    } else {
      a.probs[warp * 2] = s0; a.probs[warp * 2 + 1] = s1;
    }
    if (a.p1) a.p1[warp] = a.probs[warp * 2 + 1];
  }
}
__global__ void __launch_bounds__(256) penalties_kernel(const float* probs, int B, int N, float* pen) {
  __shared__ float red[2][256];
  float brev = 0.f, cont = 0.f;
  for (int i = threadIdx.x; i < B * N; i += 256) {
    const float p = probs[i * 2 + 1];
    brev += p;
    if (i % N) cont += fabsf(p - probs[(i - 1) * 2 + 1]);
  }
  red[0][threadIdx.x] = brev; red[1][threadIdx.x] = cont;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { red[0][threadIdx.x] += red[0][threadIdx.x + o]; red[1][threadIdx.x] += red[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    pen[0] = red[0][0] / (float)B;
    pen[1] = N > 1 ? red[1][0] / ((float)B * (float)(N - 1)) : nanf("");
  }
}
__global__ void __launch_bounds__(256) penalties_bwd_kernel(const float* probs, int B, int N, const float* g_pen,
                                                            float* dprobs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int n = i % N;
  const float p = probs[i * 2 + 1];
  float d1 = g_pen[0] / (float)B;
  if (N > 1) {
    const float sc = g_pen[1] / ((float)B * (float)(N - 1));
    if (n > 0) { const float df = p - probs[(i - 1) * 2 + 1]; d1 += sc * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f)); }
    if (n < N - 1) { const float df = probs[(i + 1) * 2 + 1] - p; d1 -= sc * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f)); }
  }
  dprobs[i * 2] = 0.f;
  dprobs[i * 2 + 1] = d1;
}
int penalties_fwd(const float* probs, int B, int N, float* pen, cudaStream_t st) {
  { LaunchScope ls_(KC_MISC, st);
  penalties_kernel<<<1, 256, 0, st>>>(probs, B, N, pen);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}
int penalties_bwd(const float* probs, int B, int N, const float* g_pen, float* dprobs, cudaStream_t st) {
  if (B * N == 0) return PVCR_OK;
  { LaunchScope ls_(KC_MISC, st);
  penalties_bwd_kernel<<<cdiv(B * N, 256), 256, 0, st>>>(probs, B, N, g_pen, dprobs);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

int gumbel_select_fwd(const GumbelArgs& a, cudaStream_t st) {
  const int rows = a.B * a.N;
  if (rows == 0) return PVCR_OK;
  { LaunchScope ls_(KC_MISC, st);
  gumbel_fwd_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, st>>>(a);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  if (a.pen) {
    { LaunchScope ls_(KC_MISC, st);
    penalties_kernel<<<1, 256, 0, st>>>(a.probs, a.B, a.N, a.pen);
    }
    PVCR_CUDA_CHECK(cudaGetLastError());
  }
  return PVCR_OK;
}

__global__ void __launch_bounds__(256) gumbel_bwd_rows_kernel(GumbelBwdArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int B = a.B, N = a.N, H = a.H;
  if (warp >= B * N) return;
  const int b = warp / N, n = warp % N;
  const float y0 = a.y[warp * 2], y1 = a.y[warp * 2 + 1];
  float d0 = a.dprobs ? a.dprobs[warp * 2] : 0.f;
  float d1 = a.dprobs ? a.dprobs[warp * 2 + 1] : 0.f;
  if (a.dp1_sel) d1 += a.dp1_sel[warp];
  const float g_brev = a.g_pen ? a.g_pen[0] : 0.f, g_cont = a.g_pen ? a.g_pen[1] : 0.f;
  d1 += g_brev / (float)B;
  if (N > 1 && g_cont != 0.f) {
    const float sc = g_cont / ((float)B * (float)(N - 1));
    if (n > 0) { const float df = y1 - a.y[(warp - 1) * 2 + 1]; d1 += sc * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f)); }
    if (n < N - 1) { const float df = a.y[(warp + 1) * 2 + 1] - y1; d1 -= sc * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f)); }
  }
  const float dot = y0 * d0 + y1 * d1;
  const float dl0 = y0 * (d0 - dot) / a.tau, dl1 = y1 * (d1 - dot) / a.tau;
  if (lane == 0) { a.scratch[warp * 2] = dl0; a.scratch[warp * 2 + 1] = dl1; }
  float* dhf = a.dhf + (long long)n * a.h_ts + (long long)b * a.h_bs;
  float* dhb = a.dhb + (long long)n * a.h_ts + (long long)b * a.h_bs;
  const unsigned long long e0 = (unsigned long long)warp * 2 * H;
  for (int k = lane; k < 2 * H; k += 32) {
    float g = dl0 * a.w[k] + dl1 * a.w[2 * H + k];
    if (a.drop.p > 0.f) g *= dropout_scale(a.drop, e0 + k);
    if (k < H) dhf[k] = g; else dhb[k - H] = g;
  }
}
// d W[c][k] = sum_rows dlogit[row][c] * Dropout(out)[row][k]: block = (64-row chunk, 256-column tile), partial sums
// merged with atomics (dw / dbias zeroed by the launcher)
constexpr int GW_ROWS = 64;
__global__ void __launch_bounds__(256) gumbel_bwd_w_kernel(GumbelBwdArgs a) {
  __shared__ float sdl[GW_ROWS][2];
  const int B = a.B, N = a.N, H = a.H, R = B * N;
  const int row0 = blockIdx.x * GW_ROWS, k = blockIdx.y * 256 + threadIdx.x;
  for (int i = threadIdx.x; i < GW_ROWS * 2; i += 256) {
    const int row = row0 + (i >> 1);
    sdl[i >> 1][i & 1] = row < R ? a.scratch[row * 2 + (i & 1)] : 0.f;
  }
  __syncthreads();
  if (k < 2 * H) {
    float g0 = 0.f, g1 = 0.f;
    for (int i = 0; i < GW_ROWS && row0 + i < R; ++i) {
      const int row = row0 + i, b = row / N, n = row % N;
      const long long o = (long long)n * a.h_ts + (long long)b * a.h_bs;
      float x = k < H ? a.hf[o + k] : a.hb[o + k - H];
      if (a.drop.p > 0.f) x *= dropout_scale(a.drop, (unsigned long long)row * 2 * H + k);
      g0 += sdl[i][0] * x;
      g1 += sdl[i][1] * x;
    }
    atomicAdd(a.dw + k, g0);
    atomicAdd(a.dw + 2 * H + k, g1);
  }
  if (blockIdx.y == 0 && threadIdx.x < 2) {
    float s = 0.f;
    for (int i = 0; i < GW_ROWS; ++i) s += sdl[i][threadIdx.x];
    atomicAdd(a.dbias + threadIdx.x, s);
  }
}
int gumbel_select_bwd(const GumbelBwdArgs& a, cudaStream_t st) {
  const int rows = a.B * a.N;
  if (rows == 0) return PVCR_OK;
  { LaunchScope ls_(KC_MISC, st);
  gumbel_bwd_rows_kernel<<<cdiv((long long)rows * 32, 256), 256, 0, st>>>(a);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  PVCR_CUDA_CHECK(cudaMemsetAsync(a.dw, 0, sizeof(float) * 4 * a.H, st));
  PVCR_CUDA_CHECK(cudaMemsetAsync(a.dbias, 0, sizeof(float) * 2, st));
  { LaunchScope ls2_(KC_MISC, st);
  gumbel_bwd_w_kernel<<<dim3(cdiv(rows, GW_ROWS), cdiv(2 * a.H, 256)), 256, 0, st>>>(a);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ x, const float* __restrict__ d, int R,
                                                     int C, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const float* xr = x + (long long)warp * C;
  const float* dr = d + (long long)warp * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c] * dr[c];
  s = warp_sum(s);
  if (lane == 0) out[warp] = s;
}
int rowdot(const float* x, const float* dsel, int R, int C, float* out, cudaStream_t st) {
  if (R == 0) return PVCR_OK;
  { LaunchScope ls_(KC_MISC, st);
  rowdot_kernel<<<cdiv((long long)R * 32, 256), 256, 0, st>>>(x, dsel, R, C, out);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

}  // namespace pvcr
