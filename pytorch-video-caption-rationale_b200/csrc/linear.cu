// Dense layer building blocks in split-bf16 tensor-core arithmetic:  y = x W^T + b  and its gradients.
// (Reference call sites: every nn.Linear / nn.GRU / nn.LSTM input projection on the hot path,
//  SURVEY.md section 2.4 rows K1, K2, K8, K11, K17.)
#include "host.h"

namespace pvcr {

size_t linear_fwd_workspace(int M, int N, int K, int nsplit) {
  Arena a(nullptr, 0);
  alloc_planes(a, M, K, nsplit);
  alloc_planes(a, N, K, nsplit);
  return a.off + 256;
}

int linear_fwd(const float* x, long long ldx, const float* w, long long ldw, const float* bias, float* y,
               long long ldy, int M, int N, int K, int nsplit, void* ws, size_t ws_bytes, cudaStream_t st) {
  PVCR_REQUIRE(nsplit >= 1 && nsplit <= 3, "linear_fwd: nsplit=%d not in 1..3", nsplit);
  if (M == 0 || N == 0) return PVCR_OK;
  Arena a(ws, ws_bytes);
  Planes xa = alloc_planes(a, M, K, nsplit);
  Planes wb = alloc_planes(a, N, K, nsplit);
  if (a.failed) { set_last_error("linear_fwd: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  PVCR_TRY(stage(x, ldx, M, K, xa, 0, nullptr, NO_DROPOUT, st));
  PVCR_TRY(stage(w, ldw, N, K, wb, 1, nullptr, NO_DROPOUT, st));
  return gemm_planes(xa.view(), wb.view(), M, N, (int)xa.ld, y, ldy, bias, 0, st);
}

size_t linear_bwd_workspace(int M, int N, int K, int nsplit) {
  Arena a(nullptr, 0);
  alloc_planes(a, M, N, nsplit);   // dY   (A role)      for dX
  alloc_planes(a, K, N, nsplit);   // W^T  (B role)      for dX
  alloc_planes(a, N, M, nsplit);   // dY^T (A role)      for dW
  alloc_planes(a, K, M, nsplit);   // X^T  (B role)      for dW
  return a.off + 256;
}

// dx[M,K] = dy W ; dw[N,K] (+)= dy^T x ; db[N] (+)= colsum(dy).  Any of dx/dw/db may be null.
int linear_bwd(const float* dy, long long lddy, const float* x, long long ldx, const float* w, long long ldw,
               float* dx, long long lddx, float* dw, long long lddw, float* db, int M, int N, int K, int nsplit,
               int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  PVCR_REQUIRE(nsplit >= 1 && nsplit <= 3, "linear_bwd: nsplit=%d not in 1..3", nsplit);
  if (M == 0 || N == 0 || K == 0) return PVCR_OK;
  Arena a(ws, ws_bytes);
  Planes dya = alloc_planes(a, M, N, nsplit);
  Planes wtb = alloc_planes(a, K, N, nsplit);
  Planes dyta = alloc_planes(a, N, M, nsplit);
  Planes xtb = alloc_planes(a, K, M, nsplit);
  if (a.failed) { set_last_error("linear_bwd: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  if (dx) {
    PVCR_TRY(stage(dy, lddy, M, N, dya, 0, nullptr, NO_DROPOUT, st));
    PVCR_TRY(transpose_split(w, ldw, N, K, wtb.ptr, wtb.ld, wtb.Kp, 0, 1, nsplit, 1, nullptr, nullptr, st, NO_DROPOUT));
    PVCR_TRY(gemm_planes(dya.view(), wtb.view(), M, K, (int)dya.ld, dx, lddx, nullptr, 0, st));
  }
  if (dw) {
    PVCR_TRY(transpose_split(dy, lddy, M, N, dyta.ptr, dyta.ld, dyta.Kp, 0, 1, nsplit, 0, nullptr, nullptr, st, NO_DROPOUT));
    PVCR_TRY(transpose_split(x, ldx, M, K, xtb.ptr, xtb.ld, xtb.Kp, 0, 1, nsplit, 1, nullptr, nullptr, st, NO_DROPOUT));
    PVCR_TRY(gemm_planes(dyta.view(), xtb.view(), N, K, (int)dyta.ld, dw, lddw, nullptr, accumulate, st));
  }
  if (db) PVCR_TRY(colsum(dy, lddy, M, N, db, accumulate, st));
  return PVCR_OK;
}

// dw[N,K] (+)= dy^T x with MN-major tensor-core operands: dy [R,N] and x [R,K] are only cast to bf16 (row-major,
// no transposed copies); the contraction runs over their rows.
size_t wgrad_mn_workspace(int R, int N, int K) {
  Arena a(nullptr, 0);
  alloc_planes(a, R, N, 1);
  alloc_planes(a, R, K, 1);
  return a.off + 256;
}
int wgrad_mn(const float* dy, long long lddy, const float* x, long long ldx, float* dw, long long lddw, int R, int N,
             int K, int accumulate, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (R == 0 || N == 0 || K == 0) return PVCR_OK;
  Arena a(ws, ws_bytes);
  Planes dya = alloc_planes(a, R, N, 1);
  Planes xa = alloc_planes(a, R, K, 1);
  if (a.failed) { set_last_error("wgrad_mn: workspace too small (%zu < %zu)", ws_bytes, a.off); return PVCR_ERR_WORKSPACE; }
  PVCR_TRY(stage(dy, lddy, R, N, dya, 0, nullptr, NO_DROPOUT, st));
  PVCR_TRY(stage(x, ldx, R, K, xa, 0, nullptr, NO_DROPOUT, st));
  return gemm_mn_store(dya.view(), xa.view(), N, K, R, dw, lddw, accumulate, st);
}

}  // namespace pvcr
