/* pvcr_b200 — C ABI of the B200-native captioning hot path.
 *
 * Drop-in boundary for p-kar/pytorch-video-caption-rationale.  The reference has no FFI: its boundary is
 * the Python nn.Module surface (model/S2VTAttModel.py:199-264, model/S2VTModel.py:12-202,
 * model/RationaleNet.py:57-106) plus the loss contract (train_utils.py:22-95).  Each entry point below
 * names the reference call it replaces; INTEGRATION.md shows the ctypes binding.
 *
 * Conventions: every pointer is a DEVICE pointer unless stated; matrices are row-major fp32 with an
 * explicit leading dimension in elements; token ids / lengths are int64 (torch.long); `stream` is a
 * cudaStream_t passed as void*.  No hidden allocation: scratch comes from the caller through
 * (workspace, workspace_bytes) sized by the matching *_workspace query.  All work is enqueued on
 * `stream`; nothing synchronises.  Return value 0 = ok, negative = error (pvcr_last_error()).
 *
 * `nsplit` selects the tensor-core arithmetic: 1 = plain bf16 operands with fp32 accumulation (training
 * throughput mode), 2 / 3 = each fp32 operand split into 2 / 3 bf16 terms (3 / 6 tcgen05 products per
 * logical product; nsplit = 3 reproduces fp32 arithmetic and is what greedy decoding uses so that token ids
 * match the fp32 reference bit for bit).
 */
#ifndef PVCR_B200_H
#define PVCR_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* pvcr_last_error(void);
int pvcr_version(void);

/* y[M,N] = x[M,K] w[N,K]^T + bias[N]      (torch.nn.Linear / F.linear; bias may be NULL) */
size_t pvcr_linear_fwd_workspace(int M, int N, int K, int nsplit);
int pvcr_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, float* y,
                    int64_t ldy, int M, int N, int K, int nsplit, void* workspace, size_t workspace_bytes,
                    void* stream);
/* dx = dy w ; dw (+)= dy^T x ; db (+)= colsum(dy)   (autograd of F.linear; any output may be NULL) */
size_t pvcr_linear_bwd_workspace(int M, int N, int K, int nsplit);
int pvcr_linear_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* w, int64_t ldw,
                    float* dx, int64_t lddx, float* dw, int64_t lddw, float* db, int M, int N, int K, int nsplit,
                    int accumulate, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
