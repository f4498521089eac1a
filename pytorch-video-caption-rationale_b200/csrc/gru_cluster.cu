// Persistent GRU forward sweep with the state exchange in DISTRIBUTED SHARED MEMORY (thread-block clusters).
//
// gru_persist.cu hands h_t from the CTAs of a group to each other through global memory: stores, a release-add on the
// group counter, an acquire poll, a CTA barrier and a 16 KB load from L2 -- 1.45 of the 4.15 us of a step (phase table in
// profiles/r02_phase_tables.md).  Here the C CTAs that serve a group of 16 videos form ONE cluster (C = H/32 <= 16; 16 is
// the non-portable maximum, one cluster per GPC on a B200): after the GRU cell every CTA writes its 16 x 32 slice of h_t
// (bf16) straight into the MMA operand buffer of EVERY CTA of the cluster (st.shared::cluster, 128-byte-swizzled
// K-major layout, double-buffered by step parity) and arrives on the consumers' mbarriers (release.cluster); the MMA-issuing
// thread waits for its C arrivals (acquire.cluster, bounded spin), fences the async proxy and issues the step's product.
// The global stores of h_t / the saved gates leave the critical path: they are issued after the push.
// Work split, arithmetic and outputs are those of gru_persist_fwd_kernel (same parity tests).
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "host.h"
#include "persist.cuh"

namespace pvcr {

constexpr int GC_ITEMS = 4;       // (unit, video) pairs per thread: u * bs <= 4 * 256 (bs = 16 or 32 videos per cluster)

struct GruClusterFwd {
  int T, B, H, C, u, bs;
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
  long long* dbg;
};

__global__ void __launch_bounds__(PERSIST_THREADS, 1) gru_cluster_fwd_kernel(const GruClusterFwd p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int H = p.H, u = p.u, C = p.C, KB = H >> 6, Rw = 3 * u, bs = p.bs;
  uint8_t* sW = smem;
  uint8_t* sX0 = sW + (size_t)KB * Rw * 128;                    // two operand buffers (step parity)
  const uint32_t x_bytes = (uint32_t)KB * bs * 128;
  float* sS = reinterpret_cast<float*>(sX0 + 2 * (size_t)x_bytes);
  const int s_ld = Rw + 1;
  bf16* sO = reinterpret_cast<bf16*>(sS + (size_t)bs * s_ld + 4);          // [bs][u] this CTA's slice of h_t
  uint64_t* bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sO + (size_t)bs * u) + 15) & ~uintptr_t(7));
  uint64_t* bar_in = bar + 1;                                    // [2]: the C slices of h_{t-1} have landed in sX[parity]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = blockIdx.x / C;
  const int c = (int)cluster_ctarank();
  const int b0 = g * bs, j0 = c * u;

  // resident weights: local row q*u + jj  <-  W_hh row q*H + j0 + jj
  for (int q = 0; q < 3; ++q)
    load_operand_rows(sW, Rw, q * u, p.whh, p.whh_ld, (long long)q * H + j0, u, (long long)3 * H, H);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar_in, (uint32_t)C);
    mbar_init(bar_in + 1, (uint32_t)C);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 32u);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();                      // every CTA's barriers and buffers exist before anybody pushes into them
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = umma_idesc_bf16(128, bs);

  // this thread's (unit, video) pairs
  const int n_items = (u * bs + PERSIST_THREADS - 1) / PERSIST_THREADS;
  float hreg[GC_ITEMS], bhr[GC_ITEMS], bhz[GC_ITEMS], bhn[GC_ITEMS];
#pragma unroll
  for (int k = 0; k < GC_ITEMS; ++k) {
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
  // push role of this thread: destination CTA tid / 16, video rows tid % 16 (+ 16)
  const int pd = tid >> 4, prow = tid & 15;

  for (int t = 0; t < p.T; ++t) {
    uint8_t* sX = sX0 + (size_t)(t & 1) * x_bytes;
    // prefetch this step's input projections (independent of the exchange)
    float gir[GC_ITEMS], giz[GC_ITEMS], gin[GC_ITEMS];
#pragma unroll
    for (int k = 0; k < GC_ITEMS; ++k) {
      gir[k] = giz[k] = gin[k] = 0.f;
      if (k < n_items) {
        const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
        const int j = j0 + jj, b = b0 + lb;
        if (lb < bs && b < p.B) {
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
      if (t == 0) {           // caller-given initial state: from global memory, once
        load_operand_rows_async(sX, bs, 0, p.h0p, p.h0p_ld, b0, bs, p.B, H);
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async();
        __syncthreads();
      }
      if (tid == 0) {
        if (t > 0) {
          mbar_wait_cluster(bar_in + (t & 1), (uint32_t)(((t - 1) >> 1) & 1));
          fence_proxy_async();                 // the peers' generic-proxy stores into sX before the tensor core reads them
        }
        phase_stamp(p.dbg, t, 1);
        tc_fence_after();
        issue_swapped_mma(tmem_base, smem_u32(sW), Rw, smem_u32(sX), bs, H, idesc, bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      phase_stamp(p.dbg, t, 2);
      if (tid < 128) tmem_to_smem_cols(tmem_base, sS, s_ld, Rw, bs);
      tc_fence_before();
      __syncthreads();
      phase_stamp(p.dbg, t, 3);
    }
    float sr[GC_ITEMS], sz[GC_ITEMS], sn[GC_ITEMS], sg[GC_ITEMS];
#pragma unroll
    for (int k = 0; k < GC_ITEMS; ++k) {
      sr[k] = sz[k] = sn[k] = sg[k] = 0.f;
      if (k < n_items) {
        const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
        if (lb < bs) {
          float ghr = bhr[k], ghz = bhz[k], ghn = bhn[k];
          if (has_prev) {
            ghr += sS[lb * s_ld + jj]; ghz += sS[lb * s_ld + u + jj]; ghn += sS[lb * s_ld + 2 * u + jj];
          }
          const float r = sigmoidf_(gir[k] + ghr);
          const float z = sigmoidf_(giz[k] + ghz);
          const float n = fast_tanh(gin[k] + r * ghn);
          const float hn = (1.f - z) * n + z * hreg[k];
          hreg[k] = hn;
          sr[k] = r; sz[k] = z; sn[k] = n; sg[k] = ghn;
          sO[lb * u + jj] = __float2bfloat16_rn(b0 + lb < p.B ? hn : 0.f);
        }
      }
    }
    __syncthreads();
    phase_stamp(p.dbg, t, 4);
    if (t + 1 < p.T) {
      // push this CTA's 16 x u slice of h_t into the next step's operand buffer of every CTA of the cluster
      if (pd < C) {
        const uint32_t dst0 = mapa_rank(smem_u32(sX0 + (size_t)((t + 1) & 1) * x_bytes), (uint32_t)pd);
        for (int row = prow; row < bs; row += 16) {
          for (int q = 0; q < (u >> 3); ++q) {
            const uint4 v = *reinterpret_cast<const uint4*>(sO + row * u + 8 * q);
            const uint32_t a = dst0 + sw128_offset(row, j0 + 8 * q, bs);
            asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
          }
        }
      }
      __syncthreads();
      if (pd < C && prow == 0) mbar_arrive_cluster(mapa_rank(smem_u32(bar_in + ((t + 1) & 1)), (uint32_t)pd));
    }
    phase_stamp(p.dbg, t, 5);
    // outputs for the callers and the backward pass: off the exchange's critical path
#pragma unroll
    for (int k = 0; k < GC_ITEMS; ++k) {
      if (k < n_items) {
        const int idx = tid + k * PERSIST_THREADS, jj = idx % u, lb = idx / u;
        const int j = j0 + jj, b = b0 + lb;
        if (lb < bs && b < p.B) {
          p.h[(long long)t * p.h_ts + (long long)b * p.h_ld + j] = hreg[k];
          p.hp[(long long)t * p.hp_ts + (long long)b * p.hp_ld + j] = __float2bfloat16_rn(hreg[k]);
          const long long o = ((long long)t * p.B + b) * H + j;
          p.r[o] = sr[k]; p.z[o] = sz[k]; p.n[o] = sn[k]; p.ghn[o] = sg[k];
        }
      }
    }
    phase_stamp(p.dbg, t, 6);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // nobody leaves while a peer may still write into its shared memory
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32u);
  }
}

static int gc_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

struct GcPlan { int u, C, G, bs; size_t smem; };
static bool plan_gc(int B, int H, int bs, GcPlan& pl) {
  if (H % 64 != 0 || H < 64 || H % 32 != 0) return false;
  pl.u = 32; pl.C = H / 32; pl.bs = bs; pl.G = (B + bs - 1) / bs;
  if (pl.C > 16 || pl.C < 1) return false;
  if (pl.u * bs > GC_ITEMS * PERSIST_THREADS) return false;
  if ((long long)pl.G * pl.C > gc_num_sms()) return false;
  const size_t KB = H / 64, Rw = 3 * pl.u;
  size_t total = KB * Rw * 128 + 2 * KB * bs * 128 + ((size_t)bs * (Rw + 1) + 4) * 4 + (size_t)bs * pl.u * 2 + 64 + 32;
  pl.smem = total + 1024 + (128 - Rw) * 128;      // + alignment slack + the 128-row MMA tile's over-read past the weight slice
  return pl.smem <= 227 * 1024;
}

static int cluster_launch_config(const void* kern, const GcPlan& pl, cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr,
                                 cudaStream_t st) {
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> g(mu);
    PVCR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    if (pl.C > 8) PVCR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  }
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(pl.G * pl.C);
  cfg.blockDim = dim3(PERSIST_THREADS);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pl.C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return PVCR_OK;
}

// All G clusters must be able to run side by side (one group's sweep is as long as the whole kernel): asked of the
// occupancy calculator once per shape.  16-CTA clusters fit one per GPC, and not every GPC of a B200 has 16 SMs to give
// (measured: 7 of the 8 needed at B = 128, H = 512) -- then the groups are widened to 32 videos (4 clusters).
static int cluster_group_videos(const GruSeq& s) {
  static const bool on = getenv("PVCR_GRU_CLUSTER") != nullptr;        // opt-in while it is being measured
  if (!on || s.nsplit != 1 || s.hp == nullptr || s.Hp != s.H) return 0;
  if (s.h0 != nullptr && s.h0_planes == nullptr) return 0;
  static std::mutex mu;
  static int cached_key = -1, cached_bs = 0;
  std::lock_guard<std::mutex> g(mu);
  const int key = s.B * 4096 + s.H;
  if (key == cached_key) return cached_bs;
  cached_key = key;
  cached_bs = 0;
  for (int bs = 16; bs <= 32 && !cached_bs; bs += 16) {
    GcPlan pl;
    if (!plan_gc(s.B, s.H, bs, pl)) continue;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    if (cluster_launch_config((const void*)gru_cluster_fwd_kernel, pl, cfg, attr, nullptr) != PVCR_OK) continue;
    int n = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&n, (const void*)gru_cluster_fwd_kernel, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    if (getenv("PVCR_GRU_CLUSTER_VERBOSE"))
      fprintf(stderr, "gru_cluster: B=%d H=%d %d videos per cluster of %d CTAs, %zu B smem: max active clusters %d (need %d) [%s]\n",
              s.B, s.H, bs, pl.C, pl.smem, n, pl.G, cudaGetErrorString(e));
    if (n >= pl.G) cached_bs = bs;
  }
  return cached_bs;
}
bool gru_cluster_eligible(const GruSeq& s) { return cluster_group_videos(s) != 0; }

int gru_cluster_fwd(const GruSeq& s, cudaStream_t st) {
  GcPlan pl;
  const int bs = cluster_group_videos(s);
  PVCR_REQUIRE(bs && plan_gc(s.B, s.H, bs, pl), "gru_cluster_fwd: shape B=%d H=%d not supported", s.B, s.H);
  GruClusterFwd p{};
  p.T = s.T; p.B = s.B; p.H = s.H; p.C = pl.C; p.u = pl.u; p.bs = bs;
  p.whh = s.whh.ptr; p.whh_ld = s.whh.ld; p.b_hh = s.b_hh;
  p.gi = s.gi_a; p.gi_ts = s.gi_a_ts; p.gi_ld = s.gi_a_ld;
  p.gi_b = s.gi_b; p.gi_b_ts = s.gi_b_ts; p.gi_b_ld = s.gi_b_ld; p.gi_b_from = s.gi_b_from;
  p.gi_bias = s.gi_bias;
  p.h0 = s.h0; p.h0_ld = s.h0_ld; p.h0p = s.h0 ? s.h0_planes : nullptr; p.h0p_ld = s.h0_planes_ld;
  p.h = s.h; p.h_ts = s.h_ts; p.h_ld = s.h_ld;
  p.hp = s.hp; p.hp_ts = s.hp_ts; p.hp_ld = s.hp_ld;
  p.r = s.r; p.z = s.z; p.n = s.n; p.ghn = s.ghn;
  p.dbg = getenv("PVCR_PHASE_GRU") ? debug_phase_buffer() : nullptr;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  PVCR_TRY(cluster_launch_config((const void*)gru_cluster_fwd_kernel, pl, cfg, attr, st));
  LaunchScope ls_(KC_GRU_FWD, st);
  PVCR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gru_cluster_fwd_kernel, p));
  return PVCR_OK;
}

}  // namespace pvcr
