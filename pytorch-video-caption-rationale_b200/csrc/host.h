// Host-side building blocks shared by the model-level entry points: workspace arena, prepared
// operand planes, and GEMM wrappers over them.
#pragma once
#include "gemm_sm100.cuh"
#include "kernels.cuh"

namespace pvcr {

const char* last_error();

// Bump allocator over a caller-provided workspace.  With base == nullptr it only measures (used by the
// *_workspace_bytes queries, so sizing and carving share one code path).
struct Arena {
  char* base;
  size_t cap, off;
  bool failed;
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
};

// bf16 split planes of a [rows, K] fp32 matrix: element (r, p, k) at ptr[r*ld + p*Kp + k], ld = P*Kp.
struct Planes {
  bf16* ptr;
  long long ld;
  int rows, K, Kp, nsplit;
  int P() const { return split_planes(nsplit); }
  OperandView view() const { return OperandView{ptr, ld, 0, rows, 1}; }
  // rows [r0, r0+nr)
  OperandView view_rows(int r0, int nr) const { return OperandView{ptr + (long long)r0 * ld, ld, 0, nr, 1}; }
};
inline Planes alloc_planes(Arena& a, int rows, int K, int nsplit) {
  Planes p;
  p.rows = rows; p.K = K; p.Kp = (int)round_up(K, 64); p.nsplit = nsplit;
  p.ld = (long long)split_planes(nsplit) * p.Kp;
  p.ptr = a.alloc<bf16>((size_t)rows * p.ld);
  return p;
}

int gemm_store(const OperandView& a, const OperandView& b, const GemmCoords& gc, int grid_z, float* C, long long ldc,
               long long c_zstride, const float* bias, long long bias_zstride, int accumulate, cudaStream_t stream);

// C[M,N] (ldc) = A * B^T (+bias) (+C); A = a_view (M rows), B = b_view (N rows), both planes over the same K.
inline int gemm_planes(const OperandView& a, const OperandView& b, int M, int N, int Kcat, float* C, long long ldc,
                       const float* bias, int accumulate, cudaStream_t st) {
  GemmCoords gc{M, N, Kcat, 0, 0, 0, 0};
  return gemm_store(a, b, gc, 1, C, ldc, 0, bias, 0, accumulate, st);
}

// fp32 [R,C] -> planes (role A or B), optional row scale / dropout
inline int stage(const float* in, long long ld_in, int R, int C, const Planes& p, int role_b, const float* row_scale,
                 Dropout drop, cudaStream_t st, int row0 = 0) {
  return cast_split(in, ld_in, R, C, p.ptr + (long long)row0 * p.ld, p.ld, p.Kp, p.nsplit, role_b, row_scale, drop, st);
}

static const Dropout NO_DROPOUT = {0.f, 0ull, 0ull};

}  // namespace pvcr
