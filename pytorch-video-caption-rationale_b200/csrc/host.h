// Host-side building blocks shared by the model-level entry points: workspace arena, prepared
// operand planes, and GEMM wrappers over them.
#pragma once
#include "gemm_sm100.cuh"
#include "kernels.cuh"

namespace pvcr {

const char* last_error();

// ---- side lane (side.cu): second, lower-priority stream for work off the step's critical path --------------------
int side_mode();                                        // 0 off, 1 join at the end of each call, 2 deferred join
bool side_site(int bit);                                // fork site enabled (PVCR_SIDE_MASK tuning aid; false when off)
int side_fork(cudaStream_t main, cudaStream_t* lane, int id = 0);   // lane `id` waits for main's current point (*lane == main when off)
int side_join_lane(cudaStream_t main, int id);          // main waits for lane `id`'s current point only
int side_join(cudaStream_t main);                       // main waits for every lane's current point
int side_call_end(cudaStream_t main);                   // join unless mode 2
int side_milestone(int id, cudaStream_t lane);          // record milestone `id` at the lane's current point (side.cu)
enum SideNote { NOTE_ATT_BWD_WEIGHTS = 1, NOTE_VOCAB_WV = 2 };
// one-shot notes between the calls of a step (side.cu), keyed by (workspace, tag) and the data they are about
void side_note_put(const void* key, int tag, const void* what = nullptr);
bool side_note_take(const void* key, int tag, const void* what = nullptr);

// Bump allocator over a caller-provided workspace.  With base == nullptr it only measures (used by the
// *_workspace_bytes queries, so sizing and carving share one code path).
struct StageCache;
struct Arena {
  char* base;
  size_t cap, off;
  bool failed;
  StageCache* cache = nullptr;   // optional: bf16 planes already staged in this call, keyed by their fp32 source
  Arena(void* b, size_t c) : base(static_cast<char*>(b)), cap(c), off(0), failed(false) {}
  template <class T>
  T* alloc(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : reinterpret_cast<T*>(uintptr_t(256));
    off += n * sizeof(T);
    if (base && off > cap) failed = true;
    return p;
  }
  bool measuring() const { return base == nullptr; }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }     // stream order makes reuse of released scratch safe
};

// bf16 split planes of a [rows, K] fp32 matrix: element (r, p, k) at ptr[r*ld + p*Kp + k], ld = P*Kp.
struct Planes {
  bf16* ptr;
  long long ld;
  int rows, K, Kp, nsplit;
  bool compact = false;          // B-role planes holding each term once (common.cuh split_term role 2)
  int P() const { return split_planes(nsplit); }
  OperandView view() const {
    OperandView v{ptr, ld, 0, rows, 1};
    if (compact) { v.kp = Kp; v.terms = nsplit; }
    return v;
  }
  // rows [r0, r0+nr)
  OperandView view_rows(int r0, int nr) const { return OperandView{ptr + (long long)r0 * ld, ld, 0, nr, 1}; }
};
// weight planes in the compact layout (a no-op distinction for nsplit == 1)
inline Planes alloc_planes_compact(Arena& a, int rows, int K, int nsplit) {
  Planes p;
  p.rows = rows; p.K = K; p.Kp = (int)round_up(K, 64); p.nsplit = nsplit; p.compact = nsplit > 1;
  p.ld = (long long)nsplit * p.Kp;
  p.ptr = a.alloc<bf16>((size_t)rows * p.ld);
  return p;
}
inline Planes alloc_planes(Arena& a, int rows, int K, int nsplit) {
  Planes p;
  p.rows = rows; p.K = K; p.Kp = (int)round_up(K, 64); p.nsplit = nsplit;
  p.ld = (long long)split_planes(nsplit) * p.Kp;
  p.ptr = a.alloc<bf16>((size_t)rows * p.ld);
  return p;
}

int gemm_mn_taps_store(const OperandView& a, const OperandView& b, int M, int N, int K, int taps, const int* koff, float* C,
                       long long ldc, long long c_stride, cudaStream_t stream);
int gemm_store(const OperandView& a, const OperandView& b, const GemmCoords& gc, int grid_z, float* C, long long ldc,
               long long c_zstride, const float* bias, long long bias_zstride, int accumulate, cudaStream_t stream);

// projection + row arg-max in one pass (gemm.cu): logits (nullable) = A B^T + bias, out[row * out_stride] = next[row] = argmax
size_t gemm_argmax_scratch(int M, int N);
int gemm_argmax(const OperandView& a, const OperandView& b, int M, int N, int Kcat, const float* bias, float* logits,
                long long ldc, long long* out, long long out_stride, long long* next, void* scratch, cudaStream_t st,
                ArgmaxParts* parts = nullptr);

// bf16 (single-plane) copies of fp32 matrices staged during one backward call: the hoisted gradient GEMMs share
// operands (d gi feeds dW_c, dW_e and d emb; the forward pass already staged vid / enc / embedded words), so each
// source is cast once.  Key = (source pointer, leading dimension, rows, cols); gathers are keyed by their id array.
struct StageCache {
  struct Entry { const void* src; long long ld; int R, C; Planes p; };
  Entry e[24];
  int n = 0;
  const Planes* find(const void* src, long long ld, int R, int C) const {
    for (int i = 0; i < n; ++i)
      if (e[i].src == src && e[i].ld == ld && e[i].R == R && e[i].C == C) return &e[i].p;
    return nullptr;
  }
  void put(const void* src, long long ld, int R, int C, const Planes& p) {
    if (n < 24) e[n++] = Entry{src, ld, R, C, p};
  }
};

// C[M,N] (+)= A^T B for row-major bf16 A [K rows, M cols], B [K rows, N cols] (MN-major tcgen05 operands)
int gemm_mn_store(const OperandView& a, const OperandView& b, int M, int N, int K, float* C, long long ldc,
                  int accumulate, cudaStream_t stream);

// C[M,N] (+)= A B for bf16 A [M, K] (K-major planes) and row-major bf16 B [K rows, N cols] (MN-major operand)
int gemm_kn_store(const OperandView& a, const OperandView& b, int M, int N, int K, float* C, long long ldc,
                  int accumulate, cudaStream_t stream);

// C[M,N] (ldc) = A * B^T (+bias) (+C); A = a_view (M rows), B = b_view (N rows), both planes over the same K.
inline int gemm_planes(const OperandView& a, const OperandView& b, int M, int N, int Kcat, float* C, long long ldc,
                       const float* bias, int accumulate, cudaStream_t st) {
  GemmCoords gc{M, N, Kcat, 0, 0, 0, 0};
  if (b.kp) {                               // compact weight planes: Kcat is the contraction the A planes span
    gc.b_kp = b.kp;
    gc.b_terms = split_b_terms(b.terms);
  }
  return gemm_store(a, b, gc, 1, C, ldc, 0, bias, 0, accumulate, st);
}

// C[M,N] (ldc) = sum_s A[. + off[s]] * B_s^T (+bias): `a` is ONE K-major matrix of Kt columns (all its rows: the shifted reads
// stay inside a.rows, rows outside read as zero), `b` holds the taps side by side along K ([N, taps * Kt]); off[s] is the row
// of `a` that output row 0 reads for tap s.  One launch, the taps accumulate in TMEM (gemm_sm100.cuh: GemmCoords::a_taps).
inline int gemm_taps(const OperandView& a, const OperandView& b, int M, int N, int Kt, int taps, const int* off, float* C,
                     long long ldc, const float* bias, cudaStream_t st) {
  GemmCoords gc{M, N, taps * Kt, 0, 0, 0, 0};
  gc.a_taps = taps; gc.a_tap_kb = Kt / GEMM_BK;
  for (int s = 0; s < taps; ++s) gc.a_tap_off[s] = off[s];
  return gemm_store(a, b, gc, 1, C, ldc, 0, bias, 0, 0, st);
}

static const Dropout NO_DROPOUT = {0.f, 0ull, 0ull};

// fp32 [R,C] -> planes (role A or B), optional row scale / dropout
inline int stage(const float* in, long long ld_in, int R, int C, const Planes& p, int role_b, const float* row_scale,
                 Dropout drop, cudaStream_t st, int row0 = 0) {
  return cast_split(in, ld_in, R, C, p.ptr + (long long)row0 * p.ld, p.ld, p.Kp, p.nsplit, role_b, row_scale, drop, st);
}


// weight [N,K] fp32 -> B-role planes [N, P*Kp]
inline int prep_weight(const float* w, long long ldw, int N, int K, const Planes& p, cudaStream_t st, int row0 = 0) {
  return stage(w, ldw, N, K, p, p.compact ? 2 : 1, nullptr, NO_DROPOUT, st, row0);
}
// weight [N,K] fp32 -> transposed B-role planes [K, P*Np] occupying contraction columns n_off .. n_off+N-1
inline int prep_weight_T(const float* w, long long ldw, int N, int K, const Planes& p, int n_off, int zero_pad,
                         cudaStream_t st) {
  return transpose_split(w, ldw, N, K, p.ptr, p.ld, p.Kp, n_off, zero_pad, p.nsplit, 1, nullptr, nullptr, st, NO_DROPOUT);
}

// scratch bytes grad_w needs for (R, N, K): the larger of its transposed-plane and row-major-plane layouts
inline size_t grad_w_scratch(int R, int N, int K, int nsplit) {
  Arena t(nullptr, 0);
  alloc_planes(t, N, R, nsplit); alloc_planes(t, K, R, nsplit);
  Arena u(nullptr, 0);
  alloc_planes(u, R, N, 1); alloc_planes(u, R, K, 1);
  return (t.off > u.off ? t.off : u.off) + 512;
}
// dw[N,K] (+)= dy^T x over R rows.  Scratch planes come from `a` and are released on return.
int grad_w(Arena& a, const float* dy, long long lddy, int R, int N, const float* x, long long ldx, int K,
           const long long* x_row_ids, const float* x_row_scale, float* dw, long long lddw, int accumulate,
           int nsplit, cudaStream_t st, Dropout x_drop = NO_DROPOUT);
// dx[R,K] (+)= dy wT^T where wT are the transposed B-role planes of w ([K, P*Np]).
int grad_x(Arena& a, const float* dy, long long lddy, int R, int N, const Planes& wT, float* dx, long long lddx,
           int accumulate, cudaStream_t st);

// bf16 mode: dx[R,Kout] (+)= dy W with the forward weight planes w ([N rows, >= Kout cols]) as MN-major operand.
int grad_x_fwdw(Arena& a, const float* dy, long long lddy, int R, int N, const Planes& w, int Kout, float* dx,
                long long lddx, int accumulate, cudaStream_t st);

// ---- GRU layer over a sequence (per-step launches: tcgen05 GEMM h W_hh^T + fused gate kernel) ----------
struct GruSeq {
  int T, B, H, nsplit;
  // input projections for step t (bias b_ih included): gi_a + t*gi_a_ts (row stride gi_a_ld); optional gi_b
  // from step gi_b_from on; optional constant bias vector
  const float* gi_a; long long gi_a_ts, gi_a_ld;
  const float* gi_b; long long gi_b_ts, gi_b_ld; int gi_b_from;
  const float* gi_bias;
  const float* b_hh;
  Planes whh;                              // B-role planes of W_hh [3H, P*Hp]
  const float* h0; long long h0_ld;        // initial state (null => zeros)
  const bf16* h0_planes; long long h0_planes_ld;
  float* h; long long h_ts, h_ld;          // h_t at h + t*h_ts
  bf16* hp; long long hp_ts, hp_ld; int Hp;   // A-role planes of h_t
  float* gh;                               // scratch [B,3H]
  float *r, *z, *n, *ghn;                  // saved [T][B,H]
  unsigned* sync;                          // >= ceil(B/16) counters for the persistent kernels (null: per-step path)
};
int gru_seq_fwd(const GruSeq& s, cudaStream_t st);
// cluster / distributed-shared-memory variant of the persistent forward sweep (gru_cluster.cu)
bool gru_cluster_eligible(const GruSeq& s);
int gru_cluster_fwd(const GruSeq& s, cudaStream_t st);

// persistent GRU forward sweep in plain fp32 arithmetic (gru_f32_persist.cu): decoding, where operands may not be rounded
struct GruF32Fwd {
  int T, B, H, C;                              // C (CTAs per group of 32 videos) is filled in by the launcher
  const float* gi; long long gi_ts, gi_ld;     // input projections of step t (b_ih included): gi + t*gi_ts, row b at b*gi_ld
  const float* w_hh;                           // [3H, H] fp32, the parameter itself
  const float* b_hh;
  const float* h0; long long h0_ld;            // initial state (null => zeros)
  float* h; long long h_ts, h_ld;              // h_t at h + t*h_ts, row b at b*h_ld
  unsigned* counters;                          // >= 32 * ceil(B/32), zeroed by the launcher
  long long* dbg = nullptr;                    // phase timestamps (tuning aid)
};
// row-streaming fp32 product for decoding <= 4 videos (gemv_f32.cu): out[b, j] = x[b, :] . W[j, :] (+ bias[j]) over the rows of
// w0 stacked on w1; arg-max mode (pmax / pidx set) leaves gemv_f32_parts(rows) (max, first index) partials per video
struct GemvF32 {
  int B, K;
  const float* x; long long x_ld;
  const float* w0; long long w0_ld; int rows0;
  const float* w1; long long w1_ld; int rows1;
  const float* bias;
  float* out; long long out_ld;          // nullable in arg-max mode
  float* pmax; int* pidx;                // [B][parts]
  int stream;                            // weights read once per launch and too large for L2: evict-first loads
};
bool gemv_f32_eligible(int B, int K);
int gemv_f32_parts(int rows);
int gemv_f32(const GemvF32& p, cudaStream_t st);
// arg-max over per-row (max, index) partials (gemm.cu): out[row * out_stride] = next[row] = index
int argmax_combine(const float* pmax, const int* pidx, int R, int nparts, long long* out, long long out_stride, long long* next,
                   cudaStream_t st);
bool gru_f32_persist_eligible(int B, int H);
int gru_f32_persist_fwd(const GruF32Fwd& p, cudaStream_t st);

struct GruSeqGrad {
  const float* dh_ext; long long dh_ext_ts, dh_ext_ld;   // gradient on every h_t (nullable)
  float* dh_carry;                         // [B,H]: in = gradient on the final state, out = gradient on h0
  float* dgi; long long dgi_ts, dgi_ld;    // [B,3H] per step
  float* dgh; long long dgh_ts, dgh_ld;
  Planes dgh_a;                            // scratch A-role planes [B, P*(3H)p]
  Planes whhT;                             // transposed B-role planes of W_hh [H, P*(3H)p]
  bf16* xch;                               // [2][B][3H] bf16 exchange buffer of the persistent kernel (nullable)
  // optional bf16 copies of dgi / dgh written by the persistent kernel (operands of the hoisted gradient GEMMs)
  bf16* dgi_p = nullptr; long long dgi_p_ts = 0, dgi_p_ld = 0;
  bf16* dgh_p = nullptr; long long dgh_p_ts = 0, dgh_p_ld = 0;
};
// Persistent single-launch implementations (gru_persist.cu); used by gru_seq_* when eligible.
bool gru_persist_eligible(const GruSeq& s);
int gru_persist_fwd(const GruSeq& s, cudaStream_t st);
int gru_persist_bwd(const GruSeq& s, const GruSeqGrad& g, cudaStream_t st);
int gru_seq_bwd(const GruSeq& s, const GruSeqGrad& g, cudaStream_t st);

// ---- persistent LSTM sequence (lstm_persist.cu; one direction of the RationaleNet generator) ---------------------
struct LstmSeqArgs {
  int T, B, H, rev;                        // rev: walk t = T-1 .. 0
  Planes whh;                              // [4H, Hp] bf16
  const float* b_hh;
  const float* gi; long long gi_ts, gi_ld; // 4H columns per (b, t), includes b_ih
  float* h; long long h_ts, h_ld;
  bf16* hp; long long hp_ts, hp_ld;
  float *si, *sf, *sg, *so, *sc;           // saved gates / cell [T][B,H]
  unsigned* sync;
};
bool lstm_persist_eligible(int B, int H, int nsplit, int Hp);
bool lstm_persist_pair_ok(int B, int H, int nsplit, int Hp);      // both directions fit the SMs side by side (32-video groups)
int lstm_persist_fwd(const LstmSeqArgs& s, cudaStream_t st, bool wide = false);
int lstm_persist_bwd(const LstmSeqArgs& s, const Planes& whhT, const float* dh_ext, long long dh_ext_ts,
                     long long dh_ext_ld, float* da, long long da_ts, long long da_ld, bf16* xch, cudaStream_t st,
                     bool wide = false);

// ---- persistent attention decoder (dec_persist.cu) ---------------------------------------------------------
struct DecPersistFwd {
  int L, B, N, H, C, u, bsp;                // bsp = videos per group rounded up to 16 (MMA N)
  const bf16* w1; long long w1_ld;          // [4H, ld]: rows [0,H) = W_q, [H,4H) = W_hh
  const bf16* w3; long long w3_ld;          // [3H, ld]: W_c = W_ih[:, :H]
  const float* b_hh;
  const float* v;                           // [H]
  const float* pk;                          // [B,N,H] fp32
  const bf16* enc_a; long long enc_ld;      // rows b*N + n
  const float* enc;                         // [B,N,H] fp32
  const float* h0; long long h0_ld;         // initial state, rows b (forward(): frame N-1 of enc; decode(): encoder_final)
  const bf16* h0_a; long long h0_a_ld;      // its bf16 operand rows
  const float* ep;                          // [B,L,3H] hoisted embedding projection (+ b_ih)
  float* q_all; long long q_ld;             // step i: q_all + i*B*q_ld, rows b
  bf16* ctx_x;                              // [L][B][H]
  float* ctx_all;                           // rows b*L + i, H
  float* alpha;                             // [L][B][N]
  float* hs;                                // [B,L,H]
  bf16* hs_a; long long hs_a_ld;            // rows b*L + i
  float *r, *z, *n, *ghn;                   // [L][B,H]
  unsigned* counters;
  long long* dbg;                           // phase timestamps (tuning aid, nullable)
  int xq = 0;                               // q hand-over P1 -> P2: 0 group barrier, 1 every thread polls its words, 2 warp 0 polls
};
struct DecPersistBwd {
  int L, B, N, H, C, u, bsp;
  const bf16* wcT; long long wcT_ld;        // [H, ld]: element (j, k) = W_c[k, j],            k in [0, 3H)
  const bf16* wcatT; long long wcatT_ld;    // [H, ld]: element (j, k) = [W_q; W_hh][k, j],    k in [0, 4H)
  const float* v;                           // [H]
  const float* pk;                          // [B,N,H]
  const bf16* enc_a; long long enc_ld;
  const float* enc;                         // [B,N,H]
  const float* h0; long long h0_ld;         // h_prev of step 0, rows b
  const float* hs;                          // [B,L,H] forward states
  const float* d_hs;                        // [B,L,H] incoming gradient
  const float* q_all; long long q_ld;       // saved q (step i: q_all + i*B*q_ld)
  const float* alpha;                       // [L][B][N]
  const float *r, *z, *n, *ghn;             // [L][B,H]
  float* dgi_all;                           // rows b*L + i, 3H   (d gi = d(ctx W_c^T + ep))
  float* d1_all;                            // rows b*L + i, 4H   ([dq | d gh])
  float* dctx_all;                          // [L][B][H]
  float* ds_all;                            // [L][B][N] d(score)
  float* dh_carry;                          // [B,H] out: gradient on the initial state (encoder final)
  bf16* xg;                                 // exchange [2][B][5H]: [dq | drp | dzp | dghn | dnp]
  // optional bf16 operand planes of the hoisted gradient GEMMs, rows b*L + i (nullable):
  //   dgi_p [.,3H] = d gi;  d1_p [.,4H] = [dq | d gh] with the i = 0 rows ZERO, so that the products with h_{i-1}
  //   can run on the forward's hs planes shifted by one row (the i = 0 rows pair with the encoder state separately)
  bf16* dgi_p = nullptr; long long dgi_p_ld = 0;
  bf16* d1_p = nullptr; long long d1_p_ld = 0;
  unsigned* counters;
  long long* dbg;
  int xd = 1;                               // dctx hand-over B2 -> B3 by polling the data: 1 every thread its words, 2 warp 0 for the CTA
};
// dpk / denc / dv from the per-step quantities saved by the backward sweep (hoisted out of the time loop)
struct AttnGradArgs {
  int L, B, N, H;
  const float *alpha, *ds, *dctx;           // [L][B][N], [L][B][N], [L][B][H]
  const float* q; long long q_ld;           // step i: q + i*B*q_ld
  const float *pk, *v;
  float *dpk, *denc, *dv_part;              // [B,N,H], [B,N,H], [B,H]  (all overwritten)
  bf16* dpk_a = nullptr; long long dpk_a_ld = 0;   // optional bf16 copy of dpk (rows b*N + n)
};
int attn_grad_hoisted(const AttnGradArgs& a, cudaStream_t st);
int dec_persist_bwd(const DecPersistBwd& p, cudaStream_t st);
bool dec_persist_eligible(int B, int N, int H, int nsplit, int Hp);
int dec_persist_fwd(const DecPersistFwd& p, cudaStream_t st);

}  // namespace pvcr
