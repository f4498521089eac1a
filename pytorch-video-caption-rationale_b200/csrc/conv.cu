// SpatialNet front: two Conv3x3 (stride 1, pad 1) + BatchNorm2d + ReLU blocks on the K x K grid features of every frame
// (reference model/SpatialNet.py:76-86,106: self.conv = Sequential(Conv2d(F, H, 3, 1, 1), BatchNorm2d(H), ReLU(),
// Conv2d(H, H, 3, 1, 1), BatchNorm2d(H), ReLU()) applied to vid_feats.view(B*N, F, K, K)), forward and hand-written backward.
//
// Convolution = implicit GEMM on the tcgen05 GEMM kernels over a FLAT ZERO-PADDED channels-last layout with SHARED padding:
// image i, position (y, x) in [0, K+1)^2 is row G + i*P + y*(K+1) + x of a [rows, channels] matrix (P = (K+1)^2, G guard rows
// of zeros at both ends); cells y < K, x < K are the image, column x = K and line y = K are zeros.  The zero column is the right
// padding of its line AND the left padding of the next one, the zero line the bottom padding of its image AND the top padding
// of the next (the first image's top / left neighbours are the guard rows).  In that layout a 3x3 tap (dy, dx) is a ROW
// OFFSET dy*(K+1) + dx, so
//     Y[r, :] = b + sum_{taps s} X[r + off_s, :] W_s^T            (valid for the image rows; zero rows/columns are discarded)
// is nine GEMMs on shifted views of the same bf16 planes accumulating into one fp32 output -- no im2col tensor.  The same
// holds backwards: dX[r, :] = sum_s dY[r - off_s, :] W_s on the zero-bordered dY, and dW_s = X[. + off_s]^T dY is a
// contraction over the rows (MN-major tcgen05 operands: both matrices as they lie).  Cost of the padding: (K+1)^2 / K^2 of
// the useful FLOPs (1.36x at K = 6; a private border per image, (K+2)^2 rows, cost 1.78x).  BatchNorm uses batch statistics
// over the image rows in training (biased variance for the normalisation, unbiased for the running estimate, momentum 0.1,
// eps 1e-5 -- torch.nn.BatchNorm2d defaults).
#include <cstdlib>

#include "../../include/pvcr_b200.h"
#include "host.h"

namespace pvcr {

namespace {

struct Geo {
  int I, K, Kp, P, G;          // images, grid, line width K + 1, positions per image (K + 1)^2, guard rows
  long long R, Rtot;           // I*P rows that the GEMMs produce; R + 2G rows allocated
  // interior test in 32-bit arithmetic for row r0 + k of a chunk whose first row r0 sits at position p0 = r0 % P of its image
  // (the kernels compute p0 and r0 / P once per thread): dimg = images past r0's, cell = interior cell index
  __host__ __device__ bool interior_at(int p0, int k, int& cell, int& dimg) const {
    const int pk = p0 + k;
    dimg = pk / P;
    const int p = pk - dimg * P, y = p / Kp, x = p - y * Kp;
    cell = y * K + x;
    return y < K && x < K;
  }
};
Geo make_geo(int I, int K) {
  Geo g;
  g.I = I; g.K = K; g.Kp = K + 1; g.P = g.Kp * g.Kp; g.G = (int)round_up(g.Kp + 1, 8);
  g.R = (long long)I * g.P; g.Rtot = g.R + 2 * g.G;
  return g;
}

// vid [I, C, K, K] fp32 (channels first, as the reference hands it over) -> xp [Rtot, C] fp32 padded channels-last (borders
// of every image zero; guards zeroed by the caller) and, optionally, feats [I*K*K, C] fp32 channels-last (the values the
// spatial attention averages, model/SpatialNet.py:109-112).  Block = (image, 64-channel tile), transposed through smem.
// Single-plane mode: xb (bf16 planes of the padded matrix, row stride ldb) is written INSTEAD of the fp32 xp -- the convolution's A
// operand directly, no 2.7 GB fp32 padded copy written and re-read by a cast at cfg4.
__global__ void __launch_bounds__(256) nchw_to_padded_cl_kernel(const float* __restrict__ vid, int C, Geo g,
                                                                float* __restrict__ xp, float* __restrict__ feats,
                                                                bf16* __restrict__ xb, long long ldb) {
  extern __shared__ float tile[];                       // [KK][65] cell-major, then the cell of every padded position (-1: border)
  const int img = blockIdx.x, c0 = blockIdx.y * 64, KK = g.K * g.K;
  int* tab = reinterpret_cast<int*>(tile + KK * 65);
  const float* src = vid + ((long long)img * C + c0) * KK;
  const int nch = min(64, C - c0);
  for (int p = threadIdx.x; p < g.P; p += 256) {
    const int y = p / g.Kp, x = p - y * g.Kp;
    tab[p] = (y < g.K && x < g.K) ? y * g.K + x : -1;
  }
  {                                                     // element i = (channel i / KK, cell i % KK), advanced without divisions
    int ch = threadIdx.x / KK, ce = threadIdx.x - ch * KK;
    const int dq = 256 / KK, dr = 256 - dq * KK;
    for (int i = threadIdx.x; i < nch * KK; i += 256) {
      tile[ce * 65 + ch] = __ldcs(src + i);
      ch += dq; ce += dr;
      if (ce >= KK) { ce -= KK; ++ch; }
    }
  }
  __syncthreads();
  const long long row0 = (long long)g.G + (long long)img * g.P;
  if (nch == 64 && xb && (C & 1) == 0) {                // two channels per thread: 128-byte bf16 / 256-byte fp32 rows per warp
    for (int i = threadIdx.x; i < g.P * 32; i += 256) {
      const int p = i >> 5, c = (i & 31) * 2, cell = tab[p];
      float v0 = 0.f, v1 = 0.f;
      if (cell >= 0) { v0 = tile[cell * 65 + c]; v1 = tile[cell * 65 + c + 1]; }
      *reinterpret_cast<__nv_bfloat162*>(xb + (row0 + p) * ldb + c0 + c) = __floats2bfloat162_rn(v0, v1);
      if (cell >= 0 && feats) *reinterpret_cast<float2*>(feats + ((long long)img * KK + cell) * C + c0 + c) = make_float2(v0, v1);
    }
    return;
  }
  for (int i = threadIdx.x; i < g.P * 64; i += 256) {
    const int p = i >> 6, c = i & 63, cell = tab[p];
    if (c >= nch) continue;
    const float v = cell >= 0 ? tile[cell * 65 + c] : 0.f;
    if (xb) xb[(row0 + p) * ldb + c0 + c] = __float2bfloat16_rn(v);
    else xp[(row0 + p) * C + c0 + c] = v;
    if (cell >= 0 && feats) feats[((long long)img * KK + cell) * C + c0 + c] = v;
  }
}

// w [Co, Ci, 3, 3] -> ws [9][Co, Ci] (tap-major), or back (gradient): w[(co*Ci + ci)*9 + s] <-> ws[(s*Co + co)*Ci + ci]
__global__ void conv_weight_taps_kernel(const float* __restrict__ w, float* __restrict__ ws, long long n, int Co, int Ci,
                                        int to_taps) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i % 9);
    const long long cc = i / 9;              // co*Ci + ci
    const long long j = (long long)s * Co * Ci + cc;
    if (to_taps) ws[j] = w[i];
    else const_cast<float*>(w)[i] = ws[j];
  }
}

// Per-channel sums over the INTERIOR rows of y [R, C]: sums[0][c] += sum y, sums[1][c] += sum y^2 (double accumulators).
// Block = 32 channel quads (16-byte loads) x 8 row lanes over a chunk of rows, four rows in flight per thread; border rows (44 % of
// the padded layout at K = 6) are never read.  (The first version read 4 bytes per thread and dependent iteration: these
// reductions over 0.67 GB ran at ~1 TB/s and were 5 ms of the cfg4 step.)
constexpr int BN_ROWS = 1024;
constexpr int BN_UNROLL = 4;
__device__ __forceinline__ void bn_block_reduce(float4 s, float4 q, bool two, int C, int c, double* __restrict__ sums) {
  __shared__ float4 ps[8][33], pq[8][33];
  const int rl = threadIdx.x >> 5, l = threadIdx.x & 31;
  ps[rl][l] = s;
  if (two) pq[rl][l] = q;
  __syncthreads();
  if (rl == 0 && c < C) {
    float4 a = ps[0][l], b = two ? pq[0][l] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      const float4 x = ps[i][l];
      a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
      if (two) { const float4 y = pq[i][l]; b.x += y.x; b.y += y.y; b.z += y.z; b.w += y.w; }
    }
    atomicAdd(sums + c, (double)a.x); atomicAdd(sums + c + 1, (double)a.y);
    atomicAdd(sums + c + 2, (double)a.z); atomicAdd(sums + c + 3, (double)a.w);
    if (two) {
      atomicAdd(sums + C + c, (double)b.x); atomicAdd(sums + C + c + 1, (double)b.y);
      atomicAdd(sums + C + c + 2, (double)b.z); atomicAdd(sums + C + c + 3, (double)b.w);
    }
  }
}
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ y, int C, Geo g, double* __restrict__ sums) {
  const int c = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4, rl = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.y * BN_ROWS, r1 = min(g.R, r0 + BN_ROWS);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  if (c < C) {
    const int p0 = (int)(r0 % g.P), n = (int)(r1 - r0);
    const float* yb = y + r0 * C + c;
    for (int k0 = rl; k0 < n; k0 += 8 * BN_UNROLL) {
      float4 v[BN_UNROLL];
#pragma unroll
      for (int u = 0; u < BN_UNROLL; ++u) {
        const int k = k0 + 8 * u;
        int cell, dimg;
        v[u] = (k < n && g.interior_at(p0, k, cell, dimg)) ? __ldg(reinterpret_cast<const float4*>(yb + (long long)k * C))
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < BN_UNROLL; ++u) {
        s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w;
        q.x += v[u].x * v[u].x; q.y += v[u].y * v[u].y; q.z += v[u].z * v[u].z; q.w += v[u].w * v[u].w;
      }
    }
  }
  bn_block_reduce(s, q, true, C, c, sums);
}

// mean / invstd from the sums (training) or from the running estimates (eval); training also updates the running estimates
__global__ void bn_finalize_kernel(const double* sums, int C, double count, float eps, float momentum, int training,
                                   float* running_mean, float* running_var, float* mean, float* invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (training) {
    const double m = sums[c] / count;
    double var = sums[C + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  } else {
    mean[c] = running_mean[c];
    invstd[c] = rsqrtf(running_var[c] + eps);
  }
}

constexpr int APPLY_ROWS = 32;      // rows per block of the two apply kernels (32-bit index arithmetic inside a block)
// z = relu((y - mean) * invstd * gamma + beta) on the interior rows.  padded_out [Rtot, C]: zero on border rows (the next
// convolution's zero padding); compact_out [I*K*K, C]: interior rows only (what the attention consumes).
__global__ void bn_relu_apply_kernel(const float* __restrict__ y, int C, Geo g, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ padded_out,
                                     float* __restrict__ compact_out, bf16* __restrict__ padded_bf, long long ld_bf) {
  const int C4 = C / 4;
  const long long r0 = (long long)blockIdx.x * APPLY_ROWS;
  const int n = (int)min((long long)APPLY_ROWS, g.R - r0), p0 = (int)(r0 % g.P);
  const long long img0 = r0 / g.P;
  for (int i = threadIdx.x; i < n * C4; i += blockDim.x) {
    const int k = i / C4, c = (i - k * C4) * 4;
    const long long r = r0 + k;
    int cell, dimg;
    const bool in = g.interior_at(p0, k, cell, dimg);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in) {
      const float4 v = *reinterpret_cast<const float4*>(y + r * C + c);
      const float4 m = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
      const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
      o.x = fmaxf((v.x - m.x) * is.x * ga.x + be.x, 0.f); o.y = fmaxf((v.y - m.y) * is.y * ga.y + be.y, 0.f);
      o.z = fmaxf((v.z - m.z) * is.z * ga.z + be.z, 0.f); o.w = fmaxf((v.w - m.w) * is.w * ga.w + be.w, 0.f);
    }
    if (padded_out) *reinterpret_cast<float4*>(padded_out + ((long long)g.G + r) * C + c) = o;
    if (padded_bf) {               // single-plane mode: the next convolution's bf16 A operand directly
      const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
      uint2 pk;
      pk.x = *reinterpret_cast<const unsigned*>(&lo); pk.y = *reinterpret_cast<const unsigned*>(&hi);
      *reinterpret_cast<uint2*>(padded_bf + ((long long)g.G + r) * ld_bf + c) = pk;
    }
    if (compact_out && in) *reinterpret_cast<float4*>(compact_out + ((img0 + dimg) * (g.K * g.K) + cell) * C + c) = o;
  }
}

// Backward of BN (batch statistics) + ReLU, pass 1: sums[0][c] = sum dyhat (= d beta), sums[1][c] = sum dyhat * xhat
// (= d gamma) over the interior rows, with dyhat = dz * [bn(y) > 0].  dz comes compact ([I*K*K, C]) or padded ([., C] with
// row offset dz_row0, interior rows used).
__global__ void __launch_bounds__(256) bn_relu_bwd_stats_kernel(const float* __restrict__ y, int C, Geo g,
                                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                const float* __restrict__ dz, int dz_compact,
                                                                double* __restrict__ sums) {
  const int c = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4, rl = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.y * BN_ROWS, r1 = min(g.R, r0 + BN_ROWS);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  if (c < C) {
    const float4 m = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
    const int p0 = (int)(r0 % g.P), n = (int)(r1 - r0);
    const long long img0 = r0 / g.P;
    for (int k0 = rl; k0 < n; k0 += 8 * BN_UNROLL) {
      float4 v[BN_UNROLL], d[BN_UNROLL];
      bool in[BN_UNROLL];
#pragma unroll
      for (int u = 0; u < BN_UNROLL; ++u) {
        const int k = k0 + 8 * u;
        int cell = 0, dimg = 0;
        in[u] = k < n && g.interior_at(p0, k, cell, dimg);
        if (in[u]) {
          const long long rr = r0 + k;
          const long long dr = dz_compact ? (img0 + dimg) * (g.K * g.K) + cell : rr;
          v[u] = __ldg(reinterpret_cast<const float4*>(y + rr * C + c));
          d[u] = __ldg(reinterpret_cast<const float4*>(dz + dr * C + c));
        }
      }
#pragma unroll
      for (int u = 0; u < BN_UNROLL; ++u) {
        if (!in[u]) continue;
        const float hx = (v[u].x - m.x) * is.x, hy = (v[u].y - m.y) * is.y, hz = (v[u].z - m.z) * is.z, hw = (v[u].w - m.w) * is.w;
        const float dx = (hx * ga.x + be.x > 0.f) ? d[u].x : 0.f, dy = (hy * ga.y + be.y > 0.f) ? d[u].y : 0.f;
        const float dzz = (hz * ga.z + be.z > 0.f) ? d[u].z : 0.f, dw = (hw * ga.w + be.w > 0.f) ? d[u].w : 0.f;
        s.x += dx; s.y += dy; s.z += dzz; s.w += dw;
        q.x += dx * hx; q.y += dy * hy; q.z += dzz * hz; q.w += dw * hw;
      }
    }
  }
  bn_block_reduce(s, q, true, C, c, sums);
}
// pass 2: dy = gamma * invstd * (dyhat - dbeta / M - xhat * dgamma / M) on the interior rows (training; eval: gamma * invstd *
// dyhat), zero on border rows, written at rows G + r of dy [Rtot, C]; also d gamma / d beta out (fp32).
__global__ void bn_relu_bwd_apply_kernel(const float* __restrict__ y, int C, Geo g, const float* __restrict__ mean,
                                         const float* __restrict__ invstd, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, const float* __restrict__ dz, int dz_compact,
                                         const double* __restrict__ sums, double count, int training,
                                         float* __restrict__ dy) {
  const int C4 = C / 4;
  const long long r0 = (long long)blockIdx.x * APPLY_ROWS;
  const int n = (int)min((long long)APPLY_ROWS, g.R - r0), p0 = (int)(r0 % g.P);
  const long long img0 = r0 / g.P;
  for (int i = threadIdx.x; i < n * C4; i += blockDim.x) {
    const int k = i / C4, c = (i - k * C4) * 4;
    const long long r = r0 + k;
    int cell, dimg;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g.interior_at(p0, k, cell, dimg)) {
      const long long dr = dz_compact ? (img0 + dimg) * (g.K * g.K) + cell : r;
      const float4 v = __ldg(reinterpret_cast<const float4*>(y + r * C + c));
      const float4 d = __ldg(reinterpret_cast<const float4*>(dz + dr * C + c));
      const float4 m = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
      const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
      const float vv[4] = {v.x, v.y, v.z, v.w}, dd[4] = {d.x, d.y, d.z, d.w}, mm[4] = {m.x, m.y, m.z, m.w};
      const float ii[4] = {is.x, is.y, is.z, is.w}, gg[4] = {ga.x, ga.y, ga.z, ga.w}, bb[4] = {be.x, be.y, be.z, be.w};
      float oo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xh = (vv[e] - mm[e]) * ii[e];
        const float dh = (xh * gg[e] + bb[e] > 0.f) ? dd[e] : 0.f;
        oo[e] = training ? gg[e] * ii[e] * (dh - (float)(sums[c + e] / count) - xh * (float)(sums[C + c + e] / count)) : gg[e] * ii[e] * dh;
      }
      o = make_float4(oo[0], oo[1], oo[2], oo[3]);
    }
    *reinterpret_cast<float4*>(dy + ((long long)g.G + r) * C + c) = o;
  }
}
__global__ void bn_param_grads_kernel(const double* sums, int C, float* dgamma, float* dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { dbeta[c] = (float)sums[c]; dgamma[c] = (float)sums[C + c]; }
}
// sums[c] += sum over the INTERIOR rows of in[r, c] (double atomics; the conv bias gradient on the zero-bordered dY: its border
// rows are zero and are not read)
__global__ void __launch_bounds__(256) colsum_rows_kernel(const float* __restrict__ in, int C, Geo g, double* __restrict__ sums) {
  const int c = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4, rl = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.y * BN_ROWS, r1 = min(g.R, r0 + BN_ROWS);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    const int p0 = (int)(r0 % g.P), n = (int)(r1 - r0);
    const float* ib = in + r0 * C + c;
    for (int k0 = rl; k0 < n; k0 += 8 * BN_UNROLL) {
      float4 v[BN_UNROLL];
#pragma unroll
      for (int u = 0; u < BN_UNROLL; ++u) {
        const int k = k0 + 8 * u;
        int cell, dimg;
        v[u] = (k < n && g.interior_at(p0, k, cell, dimg)) ? __ldg(reinterpret_cast<const float4*>(ib + (long long)k * C))
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < BN_UNROLL; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
  }
  bn_block_reduce(s, s, false, C, c, sums);
}
__global__ void double_to_float_kernel(const double* in, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

int grid_for(long long n) { const long long b = (n + 255) / 256; return (int)(b > (long long)sm_count() * 16 ? (long long)sm_count() * 16 : b); }

struct Layer {                 // one Conv3x3 + BN + ReLU block
  int Ci, Co;
  float* xpf;                  // [Rtot, Ci] fp32 padded input (zero borders / guards)
  Planes xp;                   // its bf16 planes (A role), Rtot rows
  Planes w[9];                 // taps, B role [Co, Ci]
  Planes wT[9];                // transposed taps [Ci, Co] (B role of dX = dY W)   (backward)
  Planes w_all, wT_all;        // single-plane mode: the nine taps side by side along K ([Co, 9 Ci], [Ci, 9 Co]) for the fused launch
  float* wtaps;                // [9][Co, Ci] fp32
  float* y;                    // [R, Co] fp32 conv output (pre-BN)
  float *mean, *invstd;        // [Co]
  double* sums;                // [2*Co] scratch
  float* dy;                   // [Rtot, Co] fp32 zero-bordered gradient on y        (backward)
  float* dwtaps;               // [9][Co, Ci]                                        (backward)
};
struct FrontWs {
  Geo g;
  Layer l1, l2;
  float* dz1;                  // [Rtot, H]: gradient on the first block's output in padded rows (backward)
};

void carve_layer(Arena& a, const Geo& g, int Ci, int Co, int ns, Layer& l) {
  l.Ci = Ci; l.Co = Co;
  l.xpf = a.alloc<float>((size_t)g.Rtot * Ci);
  l.xp = alloc_planes(a, (int)g.Rtot, Ci, ns);
  for (int s = 0; s < 9; ++s) l.w[s] = alloc_planes(a, Co, Ci, ns);
  for (int s = 0; s < 9; ++s) l.wT[s] = alloc_planes(a, Ci, Co, ns);
  l.w_all = alloc_planes(a, Co, 9 * Ci, 1);
  l.wT_all = alloc_planes(a, Ci, 9 * Co, 1);
  l.wtaps = a.alloc<float>((size_t)9 * Co * Ci);
  l.y = a.alloc<float>((size_t)g.R * Co);
  l.mean = a.alloc<float>(Co); l.invstd = a.alloc<float>(Co);
  l.sums = a.alloc<double>((size_t)2 * Co);
  l.dy = a.alloc<float>((size_t)g.Rtot * Co);
  l.dwtaps = a.alloc<float>((size_t)9 * Co * Ci);
}
size_t front_scratch(const Geo& g, int F, int H, int ns) {
  // transient planes of the backward: dY planes (A role, Rtot rows) + grad_w's own scratch for one tap
  Arena a(nullptr, 0);
  alloc_planes(a, (int)g.Rtot, H, ns);
  const size_t gw = grad_w_scratch((int)g.R, H, F > H ? F : H, ns);
  return a.off + gw + 4096;
}
void carve_front(Arena& a, int I, int K, int F, int H, int ns, FrontWs& w) {
  w.g = make_geo(I, K);
  carve_layer(a, w.g, F, H, ns, w.l1);
  carve_layer(a, w.g, H, H, ns, w.l2);
  w.dz1 = a.alloc<float>((size_t)w.g.Rtot * H);
}

int off_of(const Geo& g, int s) { return (s / 3 - 1) * g.Kp + (s % 3 - 1); }
// single-plane (bf16) mode with 64-aligned channel counts: the nine taps of a convolution in ONE launch, accumulated in TMEM
// (gemm_taps); otherwise nine launches accumulating through the fp32 output.  A/B knob: PVCR_NO_CONV_FUSED_TAPS=1.
bool fused_taps(int ns, int Ci, int Co) {
  static const bool off = getenv("PVCR_NO_CONV_FUSED_TAPS") != nullptr;
  return !off && ns == 1 && Ci % 64 == 0 && Co % 64 == 0 && Co >= 256 && Ci >= 256;
}

// y[R, Co] = bias + sum_s X[. + off_s] W_s^T
// single-plane mode with 64-aligned channels: the producer of a layer's padded input writes its bf16 planes directly (no fp32 xpf)
bool direct_planes(int ns, int Ci) {
  static const bool off = getenv("PVCR_NO_FRONT_DIRECT_BF16") != nullptr;       // A/B knob
  return !off && ns == 1 && Ci % 64 == 0;
}
int zero_plane_guards(const Geo& g, const Layer& l, cudaStream_t st) {
  PVCR_TRY(fill_zero(l.xp.ptr, sizeof(bf16) * (size_t)g.G * l.xp.ld, st));
  return fill_zero(l.xp.ptr + ((size_t)g.G + g.R) * l.xp.ld, sizeof(bf16) * (size_t)g.G * l.xp.ld, st);
}

int conv_forward(const Geo& g, const Layer& l, const float* w, const float* bias, int ns, cudaStream_t st) {
  const long long n = (long long)9 * l.Co * l.Ci;
  conv_weight_taps_kernel<<<grid_for(n), 256, 0, st>>>(w, l.wtaps, n, l.Co, l.Ci, 1);
  PVCR_CUDA_CHECK(cudaGetLastError());
  if (!direct_planes(ns, l.Ci)) PVCR_TRY(stage(l.xpf, l.Ci, (int)g.Rtot, l.Ci, l.xp, 0, nullptr, NO_DROPOUT, st));
  if (fused_taps(ns, l.Ci, l.Co)) {
    int off[9];
    for (int s = 0; s < 9; ++s) {
      // tap s -> columns [s Ci, (s + 1) Ci) of the side-by-side weight plane
      PVCR_TRY(cast_split(l.wtaps + (size_t)s * l.Co * l.Ci, l.Ci, l.Co, l.Ci, l.w_all.ptr + (size_t)s * l.Ci, l.w_all.ld, l.Ci, 1, 1,
                          nullptr, NO_DROPOUT, st));
      off[s] = g.G + off_of(g, s);
    }
    const OperandView av{l.xp.ptr, l.xp.ld, 0, (int)g.Rtot, 1};
    return gemm_taps(av, l.w_all.view(), (int)g.R, l.Co, l.Ci, 9, off, l.y, l.Co, bias, st);
  }
  for (int s = 0; s < 9; ++s) {
    PVCR_TRY(prep_weight(l.wtaps + (size_t)s * l.Co * l.Ci, l.Ci, l.Co, l.Ci, l.w[s], st));
    const OperandView av{l.xp.ptr + ((long long)g.G + off_of(g, s)) * l.xp.ld, l.xp.ld, 0, (int)g.R, 1};
    PVCR_TRY(gemm_planes(av, l.w[s].view(), (int)g.R, l.Co, (int)l.xp.ld, l.y, l.Co, s == 0 ? bias : nullptr, s > 0, st));
  }
  return PVCR_OK;
}

int bn_forward(const Geo& g, const Layer& l, const float* gamma, const float* beta, float* running_mean, float* running_var,
               int training, float eps, float momentum, float* padded_out, float* compact_out, cudaStream_t st,
               bf16* padded_bf = nullptr, long long ld_bf = 0) {
  const int C = l.Co;
  PVCR_REQUIRE(C % 4 == 0, "spatial front: channel count %d must be a multiple of 4", C);
  if (training) {
    PVCR_TRY(fill_zero(l.sums, sizeof(double) * 2 * C, st));
    { LaunchScope ls_(KC_MISC, st);
    bn_stats_kernel<<<dim3(cdiv(C, 128), (unsigned)cdiv(g.R, BN_ROWS)), 256, 0, st>>>(l.y, C, g, l.sums);
    }
    PVCR_CUDA_CHECK(cudaGetLastError());
  }
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>(l.sums, C, (double)g.I * g.K * g.K, eps, momentum, training, running_mean,
                                                   running_var, l.mean, l.invstd);
  PVCR_CUDA_CHECK(cudaGetLastError());
  { LaunchScope ls_(KC_MISC, st);
  bn_relu_apply_kernel<<<(unsigned)cdiv(g.R, APPLY_ROWS), 256, 0, st>>>(l.y, C, g, l.mean, l.invstd, gamma, beta, padded_out, compact_out, padded_bf, ld_bf);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

// BN + ReLU backward into l.dy (zero-bordered, guards zeroed here), d gamma / d beta out
int bn_backward(const Geo& g, const Layer& l, const float* gamma, const float* beta, const float* dz, int dz_compact,
                int training, float* dgamma, float* dbeta, cudaStream_t st) {
  const int C = l.Co;
  PVCR_REQUIRE(C % 4 == 0, "spatial front: channel count %d must be a multiple of 4", C);
  PVCR_TRY(fill_zero(l.sums, sizeof(double) * 2 * C, st));
  { LaunchScope ls_(KC_MISC, st);
  bn_relu_bwd_stats_kernel<<<dim3(cdiv(C, 128), (unsigned)cdiv(g.R, BN_ROWS)), 256, 0, st>>>(l.y, C, g, l.mean, l.invstd, gamma, beta,
                                                                                          dz, dz_compact, l.sums);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  PVCR_TRY(fill_zero(l.dy, sizeof(float) * (size_t)g.G * C, st));
  PVCR_TRY(fill_zero(l.dy + ((size_t)g.G + g.R) * C, sizeof(float) * (size_t)g.G * C, st));
  { LaunchScope ls_(KC_MISC, st);
  bn_relu_bwd_apply_kernel<<<(unsigned)cdiv(g.R, APPLY_ROWS), 256, 0, st>>>(l.y, C, g, l.mean, l.invstd, gamma, beta, dz, dz_compact, l.sums,
                                                             (double)g.I * g.K * g.K, training, l.dy);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  bn_param_grads_kernel<<<cdiv(C, 128), 128, 0, st>>>(l.sums, C, dgamma, dbeta);
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

// conv backward from l.dy: d bias, d W (all nine taps), and optionally dX[Rtot-guarded rows] = sum_s dY[. - off_s] W_s
int conv_backward(Arena& a, const Geo& g, const Layer& l, int ns, float* dw, float* dbias, float* dx_padded, cudaStream_t st) {
  const int Co = l.Co, Ci = l.Ci;
  // d bias = column sums of dY (borders are zero)
  PVCR_TRY(fill_zero(l.sums, sizeof(double) * Co, st));
  { LaunchScope ls_(KC_MISC, st);
  colsum_rows_kernel<<<dim3(cdiv(Co, 128), (unsigned)cdiv(g.R, BN_ROWS)), 256, 0, st>>>(l.dy + (size_t)g.G * Co, Co, g, l.sums);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  double_to_float_kernel<<<cdiv(Co, 128), 128, 0, st>>>(l.sums, dbias, Co);
  PVCR_CUDA_CHECK(cudaGetLastError());
  const size_t m = a.mark();
  Planes dyp = alloc_planes(a, (int)g.Rtot, Co, ns);
  if (a.failed) { set_last_error("spatial front backward: workspace too small (dY planes)"); return PVCR_ERR_WORKSPACE; }
  PVCR_TRY(stage(l.dy, Co, (int)g.Rtot, Co, dyp, 0, nullptr, NO_DROPOUT, st));
  // d W_s = dY^T X[. + off_s]: contraction over the R rows.  Single-plane mode with more output tiles per tap than the split-K
  // path serves: the nine taps as ONE batched launch (32 tiles per tap on 148 SMs otherwise).  Knob: PVCR_NO_CONV_FUSED_TAPS.
  const bool dw_batched = fused_taps(ns, Ci, Co) && (long long)cdiv(Co, GEMM_BM) * cdiv(Ci, 256) > 8;
  if (dw_batched) {
    int koff[9];
    for (int s = 0; s < 9; ++s) koff[s] = g.G + off_of(g, s);
    const OperandView dv{dyp.ptr + (long long)g.G * dyp.ld, dyp.ld, 0, (int)g.R, 1};
    const OperandView xv{l.xp.ptr, l.xp.ld, 0, (int)g.Rtot, 1};
    PVCR_TRY(gemm_mn_taps_store(dv, xv, Co, Ci, (int)g.R, 9, koff, l.dwtaps, Ci, (long long)Co * Ci, st));
  } else
  for (int s = 0; s < 9; ++s) {
    float* dws = l.dwtaps + (size_t)s * Co * Ci;
    if (ns == 1) {
      const OperandView dv{dyp.ptr + (long long)g.G * dyp.ld, dyp.ld, 0, (int)g.R, 1};
      const OperandView xv{l.xp.ptr + ((long long)g.G + off_of(g, s)) * l.xp.ld, l.xp.ld, 0, (int)g.R, 1};
      PVCR_TRY(gemm_mn_store(dv, xv, Co, Ci, (int)g.R, dws, Ci, 0, st));
    } else {
      PVCR_TRY(grad_w(a, l.dy + (size_t)g.G * Co, Co, (int)g.R, Co, l.xpf + ((long long)g.G + off_of(g, s)) * Ci, Ci, Ci, nullptr,
                      nullptr, dws, Ci, 0, ns, st));
    }
  }
  const long long n = (long long)9 * Co * Ci;
  conv_weight_taps_kernel<<<grid_for(n), 256, 0, st>>>(dw, l.dwtaps, n, Co, Ci, 0);
  PVCR_CUDA_CHECK(cudaGetLastError());
  if (dx_padded) {
    // dX[r] = sum_s dY[r - off_s] W_s   (B operand: transposed tap planes [Ci, Co], contraction over Co)
    if (fused_taps(ns, Co, Ci)) {
      int off[9];
      for (int s = 0; s < 9; ++s) {
        PVCR_TRY(prep_weight_T(l.wtaps + (size_t)s * Co * Ci, Ci, Co, Ci, l.wT_all, s * Co, 0, st));
        off[s] = g.G - off_of(g, s);
      }
      const OperandView dv{dyp.ptr, dyp.ld, 0, (int)g.Rtot, 1};
      PVCR_TRY(gemm_taps(dv, l.wT_all.view(), (int)g.R, Ci, Co, 9, off, dx_padded, Ci, nullptr, st));
    } else
    for (int s = 0; s < 9; ++s) {
      PVCR_TRY(prep_weight_T(l.wtaps + (size_t)s * Co * Ci, Ci, Co, Ci, l.wT[s], 0, 1, st));
      const OperandView dv{dyp.ptr + ((long long)g.G - off_of(g, s)) * dyp.ld, dyp.ld, 0, (int)g.R, 1};
      PVCR_TRY(gemm_planes(dv, l.wT[s].view(), (int)g.R, Ci, (int)dyp.ld, dx_padded, Ci, nullptr, s > 0, st));
    }
  }
  a.release(m);
  return PVCR_OK;
}

}  // namespace

size_t spatial_front_workspace(int I, int K, int F, int H, int nsplit) {
  Arena a(nullptr, 0);
  FrontWs w;
  carve_front(a, I, K, F, H, nsplit, w);
  return a.off + front_scratch(w.g, F, H, nsplit) + 4096;
}

int spatial_front_fwd(int I, int K, int F, int H, int nsplit, const float* vid, const PvcrSpatialFrontParams& p,
                      float* running1_mean, float* running1_var, float* running2_mean, float* running2_var, int training,
                      float eps, float momentum, float* conv_feats, float* feats_cl, void* ws, size_t ws_bytes, cudaStream_t st) {
  PVCR_REQUIRE(I > 0 && K > 0 && F > 0 && H > 0 && nsplit >= 1 && nsplit <= 3, "spatial_front_fwd: I=%d K=%d F=%d H=%d nsplit=%d",
               I, K, F, H, nsplit);
  Arena a(ws, ws_bytes);
  FrontWs w;
  carve_front(a, I, K, F, H, nsplit, w);
  if (a.failed) { set_last_error("spatial_front_fwd: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  const Geo& g = w.g;
  // guards of both padded inputs
  for (Layer* l : {&w.l1, &w.l2}) {
    if (direct_planes(nsplit, l->Ci)) { PVCR_TRY(zero_plane_guards(g, *l, st)); continue; }
    PVCR_TRY(fill_zero(l->xpf, sizeof(float) * (size_t)g.G * l->Ci, st));
    PVCR_TRY(fill_zero(l->xpf + ((size_t)g.G + g.R) * l->Ci, sizeof(float) * (size_t)g.G * l->Ci, st));
  }
  const bool d1 = direct_planes(nsplit, F), d2 = direct_planes(nsplit, H);
  { LaunchScope ls_(KC_MISC, st);
  nchw_to_padded_cl_kernel<<<dim3(I, cdiv(F, 64)), 256, sizeof(float) * (65 * K * K + g.P), st>>>(vid, F, g, d1 ? nullptr : w.l1.xpf, feats_cl,
                                                                                               d1 ? w.l1.xp.ptr : nullptr, w.l1.xp.ld);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  PVCR_TRY(conv_forward(g, w.l1, p.conv1_w, p.conv1_b, nsplit, st));
  PVCR_TRY(bn_forward(g, w.l1, p.bn1_w, p.bn1_b, running1_mean, running1_var, training, eps, momentum, d2 ? nullptr : w.l2.xpf, nullptr, st,
                      d2 ? w.l2.xp.ptr : nullptr, w.l2.xp.ld));
  PVCR_TRY(conv_forward(g, w.l2, p.conv2_w, p.conv2_b, nsplit, st));
  PVCR_TRY(bn_forward(g, w.l2, p.bn2_w, p.bn2_b, running2_mean, running2_var, training, eps, momentum, nullptr, conv_feats, st));
  return PVCR_OK;
}

// d_conv_feats [I*K*K, H] -> gradients of both blocks' parameters (the input features need none).  Needs the workspace of
// the matching spatial_front_fwd untouched.
int spatial_front_bwd(int I, int K, int F, int H, int nsplit, const PvcrSpatialFrontParams& p, int training,
                      const float* d_conv_feats, PvcrSpatialFrontParams& gr, void* ws, size_t ws_bytes, cudaStream_t st) {
  Arena a(ws, ws_bytes);
  FrontWs w;
  carve_front(a, I, K, F, H, nsplit, w);
  if (a.failed || a.off + front_scratch(w.g, F, H, nsplit) > ws_bytes) {
    set_last_error("spatial_front_bwd: workspace too small (%zu bytes)", ws_bytes);
    return PVCR_ERR_WORKSPACE;
  }
  const Geo& g = w.g;
  PVCR_TRY(bn_backward(g, w.l2, p.bn2_w, p.bn2_b, d_conv_feats, 1, training, const_cast<float*>(gr.bn2_w), const_cast<float*>(gr.bn2_b), st));
  PVCR_TRY(conv_backward(a, g, w.l2, nsplit, const_cast<float*>(gr.conv2_w), const_cast<float*>(gr.conv2_b), w.dz1 + (size_t)g.G * H, st));
  // dz1 rows are the padded rows of the first block's output: interior rows carry its gradient
  PVCR_TRY(bn_backward(g, w.l1, p.bn1_w, p.bn1_b, w.dz1 + (size_t)g.G * H, 0, training, const_cast<float*>(gr.bn1_w),
                       const_cast<float*>(gr.bn1_b), st));
  PVCR_TRY(conv_backward(a, g, w.l1, nsplit, const_cast<float*>(gr.conv1_w), const_cast<float*>(gr.conv1_b), nullptr, st));
  return PVCR_OK;
}

// ---- per-frame spatial attention over the K*K cells (model/SpatialNet.py:27-53) -----------------------------------------
// scores[b,c] = v . tanh(q[b] + pk[b,c]);  alpha = softmax_c;  ctx[b] = sum_c alpha[b,c] feats[b,c]  -- keys of width H, values of
// width Fv (the frame's input features), unlike the temporal attention of the decoder where both are H wide.
// One CTA per video: warps over cells for the scores, threads over the value columns for the context (coalesced rows).
constexpr int SA_MAX_CELLS = 256;
__global__ void __launch_bounds__(256) spatial_attn_fwd_kernel(int Kc, int H, int Fv, const float* __restrict__ q, long long q_ld,
                                                               const float* __restrict__ pk, long long pk_bs,
                                                               const float* __restrict__ feats, long long feats_bs,
                                                               const float* __restrict__ v, float* __restrict__ alpha,
                                                               float* __restrict__ ctx) {
  extern __shared__ float sm[];
  float* sq = sm;            // [H]
  float* sv = sq + H;        // [H]
  float* sa = sv + H;        // [Kc]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int d = tid; d < H; d += 256) { sq[d] = q[(long long)b * q_ld + d]; sv[d] = v[d]; }
  __syncthreads();
  const float* pkb = pk + (long long)b * pk_bs;
  for (int c = warp; c < Kc; c += 8) {
    float s = 0.f;
    for (int d = lane; d < H; d += 32) s += sv[d] * tanhf(sq[d] + pkb[(long long)c * H + d]);
    s = warp_sum(s);
    if (lane == 0) sa[c] = s;
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int c = lane; c < Kc; c += 32) mx = fmaxf(mx, sa[c]);
    mx = warp_max(mx);
    float den = 0.f;
    for (int c = lane; c < Kc; c += 32) { const float e = expf(sa[c] - mx); sa[c] = e; den += e; }
    den = warp_sum(den);
    const float inv = 1.f / den;
    for (int c = lane; c < Kc; c += 32) { const float al = sa[c] * inv; sa[c] = al; alpha[(long long)b * Kc + c] = al; }
  }
  __syncthreads();
  const float* fb = feats + (long long)b * feats_bs;
  for (int f = tid; f < Fv; f += 256) {
    float acc = 0.f;
    for (int c = 0; c < Kc; ++c) acc += sa[c] * fb[(long long)c * Fv + f];
    ctx[(long long)b * Fv + f] = acc;
  }
}

// Backward: d alpha[c] = dctx . feats[c];  d score = alpha (d alpha - sum alpha d alpha);  with e = tanh(q + pk[c]):
// dq[d] = sum_c ds[c] v[d] (1 - e^2),  dpk[c,d] = ds[c] v[d] (1 - e^2),  dv_part[b,d] = sum_c ds[c] e   (features need no gradient)
__global__ void __launch_bounds__(256) spatial_attn_bwd_kernel(int Kc, int H, int Fv, const float* __restrict__ dctx,
                                                               const float* __restrict__ q, long long q_ld,
                                                               const float* __restrict__ pk, long long pk_bs,
                                                               const float* __restrict__ feats, long long feats_bs,
                                                               const float* __restrict__ v, const float* __restrict__ alpha,
                                                               float* __restrict__ dq, long long dq_ld, float* __restrict__ dpk,
                                                               long long dpk_bs, float* __restrict__ dv_part) {
  extern __shared__ float sm[];
  float* sds = sm;           // [Kc] d alpha -> d score
  float* sal = sds + Kc;     // [Kc]
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* fb = feats + (long long)b * feats_bs;
  const float* dc = dctx + (long long)b * Fv;
  for (int c = warp; c < Kc; c += 8) {
    float s = 0.f;
    for (int f = lane; f < Fv; f += 32) s += dc[f] * fb[(long long)c * Fv + f];
    s = warp_sum(s);
    if (lane == 0) { sds[c] = s; sal[c] = alpha[(long long)b * Kc + c]; }
  }
  __syncthreads();
  if (warp == 0) {
    float dot = 0.f;
    for (int c = lane; c < Kc; c += 32) dot += sal[c] * sds[c];
    dot = warp_sum(dot);
    for (int c = lane; c < Kc; c += 32) sds[c] = sal[c] * (sds[c] - dot);
  }
  __syncthreads();
  const float* pkb = pk + (long long)b * pk_bs;
  float* dpkb = dpk + (long long)b * dpk_bs;
  for (int d = tid; d < H; d += 256) {
    const float qd = q[(long long)b * q_ld + d], vd = v[d];
    float aq = 0.f, av = 0.f;
    for (int c = 0; c < Kc; ++c) {
      const float e = tanhf(qd + pkb[(long long)c * H + d]);
      const float g = sds[c] * vd * (1.f - e * e);
      dpkb[(long long)c * H + d] = g;
      aq += g; av += sds[c] * e;
    }
    dq[(long long)b * dq_ld + d] = aq;
    dv_part[(long long)b * H + d] = av;
  }
}


// ---- vectorised variants (H a multiple of 256 up to 1024, Fv a multiple of 4, 16-byte aligned rows) ---------------------------
// The kernels above walk their cells with one 4-byte load in flight per thread and dependent iteration; at cfg4 (36 cells, keys
// 512 wide, values 2048 wide, 369 KB per video and frame, every byte from HBM) that ran at 0.7 TB/s.  Here every thread issues
// the 16-byte loads of up to 10-12 cells before it uses them, scores are reduced by warp shuffles (as attn_fwd_vec_kernel).
// NC = cells per thread and pass in the score / key phases (template parameter: 10 with 256 threads, 5 with 512)
constexpr int SAV_CC = 12;       // cells per pass in the value phases
template <int NT, int NC>
__global__ void __launch_bounds__(NT) spatial_attn_fwd_vec_kernel(int Kc, int H, int Fv, const float* __restrict__ q, long long q_ld,
                                                                   const float* __restrict__ pk, long long pk_bs,
                                                                   const float* __restrict__ feats, long long feats_bs,
                                                                   const float* __restrict__ v, float* __restrict__ alpha,
                                                                   float* __restrict__ ctx) {
  extern __shared__ float sm[];
  const int DG = H >> 3, FG = NT / DG, WPF = DG >> 5;          // dim groups of 8, cell groups, warps per cell group
  float* sP = sm;                  // [Kc][WPF]
  float* sa = sP + Kc * WPF;       // [Kc]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, dg = tid % DG, fg = tid / DG, d0 = dg * 8;
  const float4* q4 = reinterpret_cast<const float4*>(q + (long long)b * q_ld + d0);
  const float4* v4 = reinterpret_cast<const float4*>(v + d0);
  const float4 qa = q4[0], qb = q4[1], va = v4[0], vb = v4[1];
  const float q8[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
  const float v8[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
  const float* pkb = pk + (long long)b * pk_bs + d0;
  for (int c0 = 0; c0 < Kc; c0 += FG * NC) {
    float4 x[NC][2];
#pragma unroll
    for (int m = 0; m < NC; ++m) {
      const int c = c0 + fg + FG * m;
      x[m][0] = x[m][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < Kc) {
        const float4* p4 = reinterpret_cast<const float4*>(pkb + (long long)c * H);
        x[m][0] = __ldg(p4); x[m][1] = __ldg(p4 + 1);
      }
    }
#pragma unroll
    for (int m = 0; m < NC; ++m) {
      const int c = c0 + fg + FG * m;
      const float p8[8] = {x[m][0].x, x[m][0].y, x[m][0].z, x[m][0].w, x[m][1].x, x[m][1].y, x[m][1].z, x[m][1].w};
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v8[e] * tanhf(q8[e] + p8[e]);
      s = warp_sum(s);
      if (lane == 0 && c < Kc) sP[c * WPF + ((dg >> 5))] = s;
    }
  }
  __syncthreads();
  if (tid < 32) {
    float mx = -INFINITY;
    for (int c = lane; c < Kc; c += 32) {
      float s = 0.f;
      for (int w = 0; w < WPF; ++w) s += sP[c * WPF + w];
      sa[c] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float den = 0.f;
    for (int c = lane; c < Kc; c += 32) { const float e = expf(sa[c] - mx); sa[c] = e; den += e; }
    den = warp_sum(den);
    const float inv = 1.f / den;
    for (int c = lane; c < Kc; c += 32) { const float al = sa[c] * inv; sa[c] = al; alpha[(long long)b * Kc + c] = al; }
  }
  __syncthreads();
  const int F4 = Fv >> 2;
  const float4* fb = reinterpret_cast<const float4*>(feats + (long long)b * feats_bs);
  for (int f4 = tid; f4 < F4; f4 += NT) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c0 = 0; c0 < Kc; c0 += SAV_CC) {
      float4 y[SAV_CC];
#pragma unroll
      for (int k = 0; k < SAV_CC; ++k) y[k] = c0 + k < Kc ? __ldcs(fb + (long long)(c0 + k) * F4 + f4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < SAV_CC; ++k) {
        const float al = c0 + k < Kc ? sa[c0 + k] : 0.f;
        acc.x += al * y[k].x; acc.y += al * y[k].y; acc.z += al * y[k].z; acc.w += al * y[k].w;
      }
    }
    *reinterpret_cast<float4*>(ctx + (long long)b * Fv + 4 * f4) = acc;
  }
}

template <int NT, int NC>
__global__ void __launch_bounds__(NT) spatial_attn_bwd_vec_kernel(int Kc, int H, int Fv, const float* __restrict__ dctx,
                                                                   const float* __restrict__ q, long long q_ld,
                                                                   const float* __restrict__ pk, long long pk_bs,
                                                                   const float* __restrict__ feats, long long feats_bs,
                                                                   const float* __restrict__ v, const float* __restrict__ alpha,
                                                                   float* __restrict__ dq, long long dq_ld, float* __restrict__ dpk,
                                                                   long long dpk_bs, float* __restrict__ dv_part) {
  extern __shared__ float sm[];
  const int Q4 = H >> 2, CH = NT / Q4;                          // 16-byte dim columns, cell groups of the key phase
  float* sds = sm;                 // [Kc] d alpha -> d score
  float* sal = sds + Kc;           // [Kc]
  float* sW = sal + Kc;            // [8][Kc] per-warp partial d alpha
  constexpr int NW = NT / 32;
  float* sR = sW + NW * Kc + ((4 - (((2 + NW) * Kc) & 3)) & 3);        // [2][CH][H] partial dq / dv of the cell groups (16-byte aligned)
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // d alpha[c] = dctx . feats[c]
  const int F4 = Fv >> 2;
  const float4* fb = reinterpret_cast<const float4*>(feats + (long long)b * feats_bs);
  const float4* dc4 = reinterpret_cast<const float4*>(dctx + (long long)b * Fv);
  for (int c0 = 0; c0 < Kc; c0 += SAV_CC) {
    float part[SAV_CC];
#pragma unroll
    for (int k = 0; k < SAV_CC; ++k) part[k] = 0.f;
    for (int f4 = tid; f4 < F4; f4 += NT) {
      const float4 d4 = __ldg(dc4 + f4);
      float4 y[SAV_CC];
#pragma unroll
      for (int k = 0; k < SAV_CC; ++k) y[k] = c0 + k < Kc ? __ldcs(fb + (long long)(c0 + k) * F4 + f4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < SAV_CC; ++k) part[k] += d4.x * y[k].x + d4.y * y[k].y + d4.z * y[k].z + d4.w * y[k].w;
    }
#pragma unroll
    for (int k = 0; k < SAV_CC; ++k) {
      const float s = warp_sum(part[k]);
      if (lane == 0 && c0 + k < Kc) sW[warp * Kc + c0 + k] = s;
    }
  }
  __syncthreads();
  for (int c = tid; c < Kc; c += NT) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += sW[w * Kc + c];
    sds[c] = s;
    sal[c] = alpha[(long long)b * Kc + c];
  }
  __syncthreads();
  if (warp == 0) {
    float dot = 0.f;
    for (int c = lane; c < Kc; c += 32) dot += sal[c] * sds[c];
    dot = warp_sum(dot);
    for (int c = lane; c < Kc; c += 32) sds[c] = sal[c] * (sds[c] - dot);
  }
  __syncthreads();
  // keys: thread = (4 dims, cell group); cells ch, ch + CH, ...
  const int d4i = tid % Q4, ch = tid / Q4;
  const float4 qd = __ldg(reinterpret_cast<const float4*>(q + (long long)b * q_ld) + d4i);
  const float4 vd = __ldg(reinterpret_cast<const float4*>(v) + d4i);
  const float4* pkb = reinterpret_cast<const float4*>(pk + (long long)b * pk_bs);
  float4* dpkb = reinterpret_cast<float4*>(dpk + (long long)b * dpk_bs);
  float4 aq = make_float4(0.f, 0.f, 0.f, 0.f), av = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c0 = ch; c0 < Kc; c0 += CH * NC) {
    float4 x[NC];
#pragma unroll
    for (int m = 0; m < NC; ++m) {
      const int c = c0 + CH * m;
      x[m] = c < Kc ? __ldg(pkb + (long long)c * Q4 + d4i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int m = 0; m < NC; ++m) {
      const int c = c0 + CH * m;
      if (c < Kc) {
        const float ds = sds[c];
        const float ex = tanhf(qd.x + x[m].x), ey = tanhf(qd.y + x[m].y), ez = tanhf(qd.z + x[m].z), ew = tanhf(qd.w + x[m].w);
        const float4 g = make_float4(ds * vd.x * (1.f - ex * ex), ds * vd.y * (1.f - ey * ey), ds * vd.z * (1.f - ez * ez),
                                     ds * vd.w * (1.f - ew * ew));
        dpkb[(long long)c * Q4 + d4i] = g;
        aq.x += g.x; aq.y += g.y; aq.z += g.z; aq.w += g.w;
        av.x += ds * ex; av.y += ds * ey; av.z += ds * ez; av.w += ds * ew;
      }
    }
  }
  reinterpret_cast<float4*>(sR)[ch * Q4 + d4i] = aq;
  reinterpret_cast<float4*>(sR)[(CH + ch) * Q4 + d4i] = av;
  __syncthreads();
  if (ch == 0) {
    float4 a = aq, w = av;
    for (int k = 1; k < CH; ++k) {
      const float4 a2 = reinterpret_cast<const float4*>(sR)[k * Q4 + d4i], w2 = reinterpret_cast<const float4*>(sR)[(CH + k) * Q4 + d4i];
      a.x += a2.x; a.y += a2.y; a.z += a2.z; a.w += a2.w;
      w.x += w2.x; w.y += w2.y; w.z += w2.z; w.w += w2.w;
    }
    reinterpret_cast<float4*>(dq + (long long)b * dq_ld)[d4i] = a;
    reinterpret_cast<float4*>(dv_part + (long long)b * H)[d4i] = w;
  }
}

// Forward: 512 threads per video (16 warps: twice the loads in flight per SM -- one CTA per video leaves the SM to a single CTA -- and
// half the serialized load batches per phase) when the key width splits evenly.  A/B knob: PVCR_SPATIAL_ATTN_NT=256.
static bool spatial_attn_wide(int H) {
  static const int nt = getenv("PVCR_SPATIAL_ATTN_NT") ? atoi(getenv("PVCR_SPATIAL_ATTN_NT")) : 512;
  return nt == 512 && (H == 256 || H == 512 || H == 1024);
}
static bool spatial_attn_vec_ok(int Kc, int H, int Fv, long long q_ld, long long pk_bs, long long feats_bs, const void* p0,
                                const void* p1, const void* p2, const void* p3, const void* p4, const void* p5) {
  static const bool off = getenv("PVCR_NO_VEC_SPATIAL_ATTN") != nullptr;       // A/B knob
  const uintptr_t bits = reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(p2) |
                         reinterpret_cast<uintptr_t>(p3) | reinterpret_cast<uintptr_t>(p4) | reinterpret_cast<uintptr_t>(p5);
  return !off && H % 256 == 0 && H <= 1024 && Fv % 4 == 0 && q_ld % 4 == 0 && pk_bs % 4 == 0 && feats_bs % 4 == 0 && (bits & 15) == 0 &&
         Kc <= SA_MAX_CELLS;
}

// launchers (also used by the fused frame sweep, spatial_sweep.cu): dq rows dq_ld apart, dproj_key videos dpk_bs apart
int spatial_attn_fwd_launch(int B, int Kc, int H, int Fv, const float* q, long long q_ld, const float* proj_key, long long pk_batch_stride,
                            const float* feats, long long feats_batch_stride, const float* v, float* alpha, float* ctx, cudaStream_t stream) {
  if (B <= 0 || Kc <= 0 || Kc > SA_MAX_CELLS || H <= 0 || Fv <= 0) { set_last_error("pvcr_spatial_attn_fwd: B=%d Kc=%d H=%d Fv=%d", B, Kc, H, Fv); return PVCR_ERR_ARG; }
  if (spatial_attn_vec_ok(Kc, H, Fv, q_ld, pk_batch_stride, feats_batch_stride, q, proj_key, feats, v, ctx, ctx)) {
    const size_t smem_v = sizeof(float) * ((size_t)Kc * ((H >> 3) >> 5) + Kc);
    { LaunchScope ls_(KC_ATTN, stream);
    if (spatial_attn_wide(H))
      spatial_attn_fwd_vec_kernel<512, 5><<<B, 512, smem_v, stream>>>(Kc, H, Fv, q, q_ld, proj_key, pk_batch_stride, feats,
                                                                     feats_batch_stride, v, alpha, ctx);
    else
      spatial_attn_fwd_vec_kernel<256, 10><<<B, 256, smem_v, stream>>>(Kc, H, Fv, q, q_ld, proj_key, pk_batch_stride, feats,
                                                                      feats_batch_stride, v, alpha, ctx);
    }
    PVCR_CUDA_CHECK(cudaGetLastError());
    return PVCR_OK;
  }
  const size_t smem = sizeof(float) * ((size_t)2 * H + Kc);
  if (smem > 48 * 1024) { set_last_error("pvcr_spatial_attn_fwd: H=%d too large", H); return PVCR_ERR_ARG; }
  { LaunchScope ls_(KC_ATTN, stream);
  spatial_attn_fwd_kernel<<<B, 256, smem, stream>>>(Kc, H, Fv, q, q_ld, proj_key, pk_batch_stride, feats, feats_batch_stride, v, alpha, ctx);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}
int spatial_attn_bwd_launch(int B, int Kc, int H, int Fv, const float* dctx, const float* q, long long q_ld, const float* proj_key,
                            long long pk_batch_stride, const float* feats, long long feats_batch_stride, const float* v,
                            const float* alpha, float* dq, long long dq_ld, float* dproj_key, long long dpk_batch_stride,
                            float* dv_part, cudaStream_t stream) {
  if (B <= 0 || Kc <= 0 || Kc > SA_MAX_CELLS || H <= 0 || Fv <= 0) { set_last_error("pvcr_spatial_attn_bwd: B=%d Kc=%d H=%d Fv=%d", B, Kc, H, Fv); return PVCR_ERR_ARG; }
  if (spatial_attn_vec_ok(Kc, H, Fv, q_ld, pk_batch_stride, feats_batch_stride, q, proj_key, feats, v, dctx, dproj_key) &&
      ((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dv_part)) & 15) == 0 && dq_ld % 4 == 0 && dpk_batch_stride % 4 == 0) {
    // measured at cfg4 (tests/gpu_probe_spatial_attn.py): forward 45.2 -> 27.1 us per launch with 512 threads, backward 24.6 -> 33.6 us:
    // the backward stays on 256 threads unless PVCR_SPATIAL_ATTN_BWD_NT=512
    static const bool bwd_wide = getenv("PVCR_SPATIAL_ATTN_BWD_NT") && atoi(getenv("PVCR_SPATIAL_ATTN_BWD_NT")) == 512;
    int nt = bwd_wide && spatial_attn_wide(H) ? 512 : 256;
    if (sizeof(float) * ((size_t)(2 + nt / 32) * Kc + 4 + (size_t)2 * (nt / (H >> 2)) * H) > 48 * 1024) nt = 256;
    const size_t smem_v = sizeof(float) * ((size_t)(2 + nt / 32) * Kc + 4 + (size_t)2 * (nt / (H >> 2)) * H);
    { LaunchScope ls_(KC_ATTN, stream);
    if (nt == 512) {
      spatial_attn_bwd_vec_kernel<512, 5><<<B, 512, smem_v, stream>>>(Kc, H, Fv, dctx, q, q_ld, proj_key, pk_batch_stride, feats,
                                                                     feats_batch_stride, v, alpha, dq, dq_ld, dproj_key, dpk_batch_stride, dv_part);
    } else {
      spatial_attn_bwd_vec_kernel<256, 10><<<B, 256, smem_v, stream>>>(Kc, H, Fv, dctx, q, q_ld, proj_key, pk_batch_stride, feats,
                                                                      feats_batch_stride, v, alpha, dq, dq_ld, dproj_key, dpk_batch_stride, dv_part);
    }
    }
    PVCR_CUDA_CHECK(cudaGetLastError());
    return PVCR_OK;
  }
  { LaunchScope ls_(KC_ATTN, stream);
  spatial_attn_bwd_kernel<<<B, 256, sizeof(float) * 2 * Kc, stream>>>(Kc, H, Fv, dctx, q, q_ld, proj_key, pk_batch_stride, feats,
                                                                     feats_batch_stride, v, alpha, dq, dq_ld, dproj_key, dpk_batch_stride, dv_part);
  }
  PVCR_CUDA_CHECK(cudaGetLastError());
  return PVCR_OK;
}

}  // namespace pvcr

using namespace pvcr;

extern "C" {

int pvcr_spatial_attn_fwd(int B, int Kc, int H, int Fv, const float* q, int64_t q_ld, const float* proj_key, int64_t pk_batch_stride,
                          const float* feats, int64_t feats_batch_stride, const float* v, float* alpha, float* ctx, void* stream) {
  return spatial_attn_fwd_launch(B, Kc, H, Fv, q, q_ld, proj_key, pk_batch_stride, feats, feats_batch_stride, v, alpha, ctx, (cudaStream_t)stream);
}
int pvcr_spatial_attn_bwd(int B, int Kc, int H, int Fv, const float* dctx, const float* q, int64_t q_ld, const float* proj_key,
                          int64_t pk_batch_stride, const float* feats, int64_t feats_batch_stride, const float* v,
                          const float* alpha, float* dq, float* dproj_key, float* dv_part, void* stream) {
  return spatial_attn_bwd_launch(B, Kc, H, Fv, dctx, q, q_ld, proj_key, pk_batch_stride, feats, feats_batch_stride, v, alpha, dq, H,
                                 dproj_key, (long long)Kc * H, dv_part, (cudaStream_t)stream);
}

size_t pvcr_spatial_front_workspace(int I, int K, int F, int H, int nsplit) { return spatial_front_workspace(I, K, F, H, nsplit); }
int pvcr_spatial_front_fwd(int I, int K, int F, int H, int nsplit, const float* vid_feats, const PvcrSpatialFrontParams* p,
                           float* bn1_running_mean, float* bn1_running_var, float* bn2_running_mean, float* bn2_running_var,
                           int training, float eps, float momentum, float* conv_feats, float* feats_cl, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (!p || !vid_feats || !conv_feats) { set_last_error("pvcr_spatial_front_fwd: null argument"); return PVCR_ERR_ARG; }
  return spatial_front_fwd(I, K, F, H, nsplit, vid_feats, *p, bn1_running_mean, bn1_running_var, bn2_running_mean,
                           bn2_running_var, training, eps, momentum, conv_feats, feats_cl, workspace, workspace_bytes,
                           (cudaStream_t)stream);
}
int pvcr_spatial_front_bwd(int I, int K, int F, int H, int nsplit, const PvcrSpatialFrontParams* p, int training,
                           const float* d_conv_feats, PvcrSpatialFrontParams* grads, void* workspace, size_t workspace_bytes,
                           void* stream) {
  if (!p || !grads || !d_conv_feats) { set_last_error("pvcr_spatial_front_bwd: null argument"); return PVCR_ERR_ARG; }
  return spatial_front_bwd(I, K, F, H, nsplit, *p, training, d_conv_feats, *grads, workspace, workspace_bytes,
                           (cudaStream_t)stream);
}

}  // extern "C"
