// Decomposed-layer forward y = W2 (W1 x) + b (reference: the nn.Sequential built at F:84-95 /
// D:74-85, K7 of SURVEY.md).
//
// Unfused path (any dtype / rank): two passes of the tcgen05 GEMM engine with the [n, k]
// intermediate H staged in workspace (bf16 for bf16 models, bf16x3 split for fp32 models).
#include "lowrank.cuh"

#include <cuda_bf16.h>

#include <cstdint>

#include "elementwise.cuh"
#include "gemm_tc.cuh"

namespace ptd {

namespace {
inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

struct Carve {
  uint8_t* base;
  size_t off = 0;
  void* take(size_t bytes) {
    off = (off + 255) & ~static_cast<size_t>(255);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

// bf16: use in place when TMA-friendly, else padded copy; fp32: bf16x3 split.
int stage(const void* src, int is_bf16, long long rows, int cols, long long ld, Carve& cv,
          GemmOperand* op, cudaStream_t st, bool dry) {
  op->mn_major = 0;
  if (is_bf16 && aligned16(src) && (ld % 8) == 0 && !dry) {
    op->ptr = static_cast<const __nv_bfloat16*>(src);
    op->ld = ld;
    op->nseg = 1;
    op->seg_stride = 0;
    return 0;
  }
  const int nseg = is_bf16 ? 1 : 3;
  const long long ldp = round_up(cols, 8);
  __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(cv.take(2ull * nseg * rows * ldp));
  if (dry) return 0;
  if (dst == nullptr) return -12;
  int rc = split_rows(src, is_bf16, ld, rows, cols, nullptr, 0.f, dst, ldp, nseg, rows * ldp, st);
  if (rc) return rc;
  op->ptr = dst;
  op->ld = ldp;
  op->nseg = nseg;
  op->seg_stride = rows * ldp;
  return 0;
}
}  // namespace

size_t lowrank_workspace_bytes(int is_bf16, long long n, int in_f, int k, int out_f) {
  Carve cv{nullptr};
  GemmOperand op;
  stage(nullptr, is_bf16, n, in_f, in_f, cv, &op, nullptr, true);
  stage(nullptr, is_bf16, k, in_f, in_f, cv, &op, nullptr, true);
  stage(nullptr, is_bf16, out_f, k, k, cv, &op, nullptr, true);
  const int nseg = is_bf16 ? 1 : 3;
  cv.take(2ull * nseg * n * round_up(k, 8));
  return cv.off + 512;
}

int lowrank_forward(const void* X, long long ldx, const void* W1, long long ldw1, const void* W2,
                    long long ldw2, const float* bias, void* Y, long long ldy, int is_bf16,
                    long long n, int in_f, int k, int out_f, void* ws, size_t ws_bytes,
                    cudaStream_t st) {
  if (X == nullptr || W1 == nullptr || W2 == nullptr || Y == nullptr) return -22;
  if (n <= 0 || in_f <= 0 || k <= 0 || out_f <= 0 || n > 0x7fffffffLL) return -22;
  if (ldx < in_f || ldw1 < in_f || ldw2 < k || ldy < out_f) return -22;
  if (ws == nullptr || ws_bytes < lowrank_workspace_bytes(is_bf16, n, in_f, k, out_f)) return -12;
  uint8_t* base = static_cast<uint8_t*>(ws);
  base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(base) + 255) & ~uintptr_t(255));
  Carve cv{base};
  GemmOperand x, w1, w2;
  int rc;
  if ((rc = stage(X, is_bf16, n, in_f, ldx, cv, &x, st, false))) return rc;
  if ((rc = stage(W1, is_bf16, k, in_f, ldw1, cv, &w1, st, false))) return rc;
  if ((rc = stage(W2, is_bf16, out_f, k, ldw2, cv, &w2, st, false))) return rc;
  const int nseg = is_bf16 ? 1 : 3;
  const long long kp = round_up(k, 8);
  __nv_bfloat16* H = static_cast<__nv_bfloat16*>(cv.take(2ull * nseg * n * kp));
  // H = X W1^T  (K-major x K-major), written as bf16 (bf16 models) or bf16x3 split (fp32 models)
  GemmEpilogue e1;
  if (is_bf16) {
    e1.Cb = H;
    e1.ldcb = kp;
  } else {
    e1.Cs = H;
    e1.ldcs = kp;
    e1.cs_seg = n * kp;
  }
  if ((rc = gemm_tc(x, w1, static_cast<int>(n), k, in_f, -1, e1, st))) return rc;
  GemmOperand h{H, 0, kp, nseg, n * kp};
  GemmEpilogue e2;
  e2.bias = bias;
  if (is_bf16) {
    e2.Cb = static_cast<__nv_bfloat16*>(Y);
    e2.ldcb = ldy;
  } else {
    e2.C = static_cast<float*>(Y);
    e2.ldc = ldy;
  }
  return gemm_tc(h, w2, static_cast<int>(n), out_f, k, -1, e2, st);
}

}  // namespace ptd
