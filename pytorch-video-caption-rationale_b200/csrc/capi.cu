// extern "C" surface declared in include/pvcr_b200.h.
#include "../../include/pvcr_b200.h"
#include "host.h"

namespace pvcr {
size_t linear_fwd_workspace(int M, int N, int K, int nsplit);
int linear_fwd(const float*, long long, const float*, long long, const float*, float*, long long, int, int, int, int,
               void*, size_t, cudaStream_t);
size_t linear_bwd_workspace(int M, int N, int K, int nsplit);
int linear_bwd(const float*, long long, const float*, long long, const float*, long long, float*, long long, float*,
               long long, float*, int, int, int, int, int, void*, size_t, cudaStream_t);
}  // namespace pvcr

using namespace pvcr;

extern "C" {

const char* pvcr_last_error(void) { return last_error(); }
int pvcr_version(void) { return 100; }

size_t pvcr_linear_fwd_workspace(int M, int N, int K, int nsplit) { return linear_fwd_workspace(M, N, K, nsplit); }
int pvcr_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y,
                    int64_t ldy, int M, int N, int K, int nsplit, void* workspace, size_t workspace_bytes,
                    void* stream) {
  return linear_fwd(x, ldx, w, ldw, bias, y, ldy, M, N, K, nsplit, workspace, workspace_bytes, (cudaStream_t)stream);
}
size_t pvcr_linear_bwd_workspace(int M, int N, int K, int nsplit) { return linear_bwd_workspace(M, N, K, nsplit); }
int pvcr_linear_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* w, int64_t ldw,
                    float* dx, int64_t lddx, float* dw, int64_t lddw, float* db, int M, int N, int K, int nsplit,
                    int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  return linear_bwd(dy, lddy, x, ldx, w, ldw, dx, lddx, dw, lddw, db, M, N, K, nsplit, accumulate, workspace,
                    workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
