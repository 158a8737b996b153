// extern "C" boundary (include/ptdeco_b200.h). Argument checking + composition of the kernels.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "../../include/ptdeco_b200.h"
#include "eigh.cuh"
#include "elementwise.cuh"
#include "gemm_tc.cuh"
#include "lowrank.cuh"

namespace {

inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }
inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Stage `src` ([rows][ld] of dtype) as a TMA-friendly bf16 operand. bf16 data that is already
// aligned is used in place; otherwise it is copied/padded (nseg=1) or split bf16x3 (fp32, nseg=3)
// into `ws`. Returns bytes of workspace consumed via *used (256-byte rounded).
int stage_operand(const void* src, int dtype, long long rows, int cols, long long ld,
                  const float* sub, int mn_major, uint8_t* ws, size_t ws_bytes, size_t* used,
                  ptd::GemmOperand* op, cudaStream_t st) {
  *used = 0;
  op->mn_major = mn_major;
  if (dtype == PTDECO_BF16 && sub == nullptr && aligned16(src) && (ld % 8) == 0) {
    op->ptr = static_cast<const __nv_bfloat16*>(src);
    op->ld = ld;
    op->nseg = 1;
    op->seg_stride = 0;
    return 0;
  }
  const int nseg = (dtype == PTDECO_F32) ? 3 : 1;
  const long long ldp = round_up(cols, 8);
  const size_t need = static_cast<size_t>(round_up(2LL * nseg * rows * ldp, 256));
  if (ws == nullptr || ws_bytes < need) return -12;  // ENOMEM
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(ws);
  int rc = ptd::split_rows(src, dtype == PTDECO_BF16, ld, rows, cols, sub, 1.f, dst, ldp, nseg,
                           rows * ldp, st);
  if (rc) return rc;
  op->ptr = dst;
  op->ld = ldp;
  op->nseg = nseg;
  op->seg_stride = rows * ldp;
  *used = need;
  return 0;
}

size_t staged_bytes(int dtype, long long rows, int cols) {
  const int nseg = (dtype == PTDECO_F32) ? 3 : 1;
  return static_cast<size_t>(round_up(2LL * nseg * rows * round_up(cols, 8), 256));
}

}  // namespace

extern "C" {

int ptdeco_version(void) { return 100; }

const char* ptdeco_strerror(int code) {
  switch (code) {
    case 0: return "ok";
    case -5: return "CUDA launch failed (EIO)";
    case -12: return "workspace too small or missing (ENOMEM)";
    case -22: return "invalid argument / unsupported alignment (EINVAL)";
    case -34: return "numerical failure: no convergence (ERANGE)";
    case -38: return "driver entry point cuTensorMapEncodeTiled unavailable (ENOSYS)";
    default: return code <= -1000 ? "cuTensorMapEncodeTiled failed (CUresult = -code-1000)" : "unknown error";
  }
}

size_t ptdeco_syrk_workspace_bytes(int dtype, long long n_tokens, int d) {
  return staged_bytes(dtype, n_tokens, d);
}

int ptdeco_syrk_accumulate(const void* Y, int dtype, long long n_tokens, int d, long long ldy,
                           const float* sub, float* C, long long ldc, float* colsum, float alpha,
                           void* workspace, size_t workspace_bytes, void* stream) {
  return ptdeco_syrk_accumulate_ex(Y, dtype, n_tokens, d, ldy, sub, C, ldc, colsum, alpha, workspace,
                                   workspace_bytes, stream, 0u);
}

int ptdeco_syrk_accumulate_ex(const void* Y, int dtype, long long n_tokens, int d, long long ldy,
                              const float* sub, float* C, long long ldc, float* colsum, float alpha,
                              void* workspace, size_t workspace_bytes, void* stream, unsigned flags) {
  if (dtype != PTDECO_F32 && dtype != PTDECO_BF16) return -22;
  if (d <= 0 || n_tokens < 0 || Y == nullptr || C == nullptr || ldc < d || ldy < d) return -22;
  if (n_tokens == 0) return 0;
  if (n_tokens > 0x7fffffffLL) return -22;
  cudaStream_t st = as_stream(stream);
  if (colsum != nullptr) {
    int rc = ptd::colsum(Y, dtype == PTDECO_BF16, ldy, n_tokens, d, alpha, sub, 1.f, colsum, st);
    if (rc) return rc;
  }
  ptd::GemmOperand op;
  size_t used = 0;
  int rc = stage_operand(Y, dtype, n_tokens, d, ldy, sub, /*mn_major=*/1,
                         static_cast<uint8_t*>(workspace), workspace_bytes, &used, &op, st);
  if (rc) return rc;
  ptd::GemmEpilogue ep;
  ep.alpha = alpha;
  ep.C = C;
  ep.ldc = ldc;
  ep.accumulate = 1;
  ep.lower_only = 1;
  ep.deterministic = (flags & PTDECO_FLAG_DETERMINISTIC) ? 1 : 0;
  return ptd::gemm_tc(op, op, d, d, static_cast<int>(n_tokens), -1, ep, st);
}

int ptdeco_cov_finalize(float* C, long long ldc, int d, const float* colsum, int n_steps,
                        int use_mean, float damp_factor, float* damp_out, void* stream) {
  if (C == nullptr || d <= 0 || ldc < d || n_steps <= 0) return -22;
  return ptd::cov_finalize(C, ldc, d, colsum, 1.f / static_cast<float>(n_steps), use_mean,
                           damp_factor, damp_out, as_stream(stream));
}

size_t ptdeco_gemm_workspace_bytes(int a_dtype, int b_dtype, int M, int N, int K) {
  // upper bound valid for either storage order of each operand
  auto bound = [](int dtype, long long a, long long b) {
    const int nseg = (dtype == PTDECO_F32) ? 3 : 1;
    return static_cast<size_t>(round_up(2LL * nseg * round_up(a, 8) * round_up(b, 8), 256));
  };
  return bound(a_dtype, M, K) + bound(b_dtype, N, K);
}

int ptdeco_gemm(const void* A, int a_dtype, int a_mn_major, long long lda, const void* B,
                int b_dtype, int b_mn_major, long long ldb, int M, int N, int K, float alpha,
                const float* bias, void* C, int c_dtype, long long ldc, int accumulate,
                void* workspace, size_t workspace_bytes, void* stream) {
  return ptdeco_gemm_ex(A, a_dtype, a_mn_major, lda, B, b_dtype, b_mn_major, ldb, M, N, K, alpha, bias,
                        C, c_dtype, ldc, accumulate, workspace, workspace_bytes, stream, 0u);
}

int ptdeco_gemm_ex(const void* A, int a_dtype, int a_mn_major, long long lda, const void* B,
                   int b_dtype, int b_mn_major, long long ldb, int M, int N, int K, float alpha,
                   const float* bias, void* C, int c_dtype, long long ldc, int accumulate,
                   void* workspace, size_t workspace_bytes, void* stream, unsigned flags) {
  if (A == nullptr || B == nullptr || C == nullptr || M <= 0 || N <= 0 || K <= 0) return -22;
  if (c_dtype == PTDECO_BF16 && accumulate) return -22;
  cudaStream_t st = as_stream(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  size_t left = workspace_bytes, used = 0;
  ptd::GemmOperand a, b;
  int rc = stage_operand(A, a_dtype, a_mn_major ? K : M, a_mn_major ? M : K, lda, nullptr,
                         a_mn_major, ws, left, &used, &a, st);
  if (rc) return rc;
  if (ws) ws += used;
  left -= used;
  rc = stage_operand(B, b_dtype, b_mn_major ? K : N, b_mn_major ? N : K, ldb, nullptr, b_mn_major,
                     ws, left, &used, &b, st);
  if (rc) return rc;
  ptd::GemmEpilogue ep;
  ep.alpha = alpha;
  ep.bias = bias;
  ep.accumulate = accumulate;
  ep.deterministic = (flags & PTDECO_FLAG_DETERMINISTIC) ? 1 : 0;
  if (c_dtype == PTDECO_BF16) {
    ep.Cb = static_cast<__nv_bfloat16*>(C);
    ep.ldcb = ldc;
  } else {
    ep.C = static_cast<float*>(C);
    ep.ldc = ldc;
  }
  return ptd::gemm_tc(a, b, M, N, K, -1, ep, st);
}

size_t ptdeco_eigh_workspace_bytes(int d, int k) {
  if (d <= 0 || k < 1 || k > d) return 0;
  return ptd::eigh_workspace_bytes(d, k);
}

int ptdeco_eigh(const float* A, int d, long long lda, int k, float* evals, float* U, long long ldu,
                void* workspace, size_t workspace_bytes, void* stream) {
  return ptd::eigh(A, d, lda, k, evals, U, ldu, workspace, workspace_bytes, as_stream(stream), 0u);
}

int ptdeco_eigh_ex(const float* A, int d, long long lda, int k, float* evals, float* U, long long ldu,
                   void* workspace, size_t workspace_bytes, void* stream, unsigned flags) {
  return ptd::eigh(A, d, lda, k, evals, U, ldu, workspace, workspace_bytes, as_stream(stream), flags);
}

size_t ptdeco_lowrank_workspace_bytes(int dtype, long long n, int in_features, int k,
                                      int out_features) {
  return ptd::lowrank_workspace_bytes(dtype == PTDECO_BF16, n, in_features, k, out_features);
}

int ptdeco_lowrank_forward(const void* X, long long ldx, const void* W1, long long ldw1,
                           const void* W2, long long ldw2, const float* bias, void* Y,
                           long long ldy, int dtype, long long n, int in_features, int k,
                           int out_features, void* workspace, size_t workspace_bytes,
                           void* stream) {
  if (dtype != PTDECO_F32 && dtype != PTDECO_BF16) return -22;
  return ptd::lowrank_forward(X, ldx, W1, ldw1, W2, ldw2, bias, Y, ldy, dtype == PTDECO_BF16, n,
                              in_features, k, out_features, workspace, workspace_bytes,
                              as_stream(stream));
}

size_t ptdeco_nsr_workspace_bytes(long long channels) {
  return static_cast<size_t>(channels) * 3 * sizeof(double);
}

int ptdeco_nsr_metric(const void* x, const void* y, int dtype, long long rows, long long channels,
                      double eps, void* workspace, size_t workspace_bytes, float* out,
                      void* stream) {
  if (x == nullptr || y == nullptr || out == nullptr) return -22;
  if (workspace == nullptr || workspace_bytes < ptdeco_nsr_workspace_bytes(channels)) return -12;
  return ptd::nsr_metric(x, y, dtype == PTDECO_BF16, rows, channels, eps,
                         static_cast<double*>(workspace), out, as_stream(stream));
}

int ptdeco_kl_metric(const void* student, const void* teacher, int dtype, long long rows,
                     long long classes, float* out, void* stream) {
  if (student == nullptr || teacher == nullptr || out == nullptr) return -22;
  return ptd::kl_metric(student, teacher, dtype == PTDECO_BF16, rows, classes, out,
                        as_stream(stream));
}

void ptdeco_debug_set(int key, long long value) {
  if (key == 100) ptd::eigh_debug_profile(static_cast<int>(value));
  else if (key == 101) ptd::eigh_debug_sym_min_m(static_cast<int>(value));
  else if (key == 102) ptd::eigh_debug_resident(static_cast<int>(value), 0, -1);
  else if (key == 103) ptd::eigh_debug_resident(1, static_cast<int>(value), -1);
  else if (key == 104) ptd::eigh_debug_resident(1, 0, static_cast<int>(value));
  else if (key == 105) ptd::eigh_debug_bisect_narrow(static_cast<int>(value));
  else if (key == 106) ptd::eigh_debug_sturm_ratio(static_cast<int>(value));
  else if (key == 107) ptd::eigh_debug_small(static_cast<int>(value));
  else if (key >= 200 && key < 210) ptd::lowrank_debug_set(key - 200, value);
  else ptd::gemm_tc_debug_set(key, value);
}
long long ptdeco_debug_get(int key) {
  if (key >= 100 && key < 116) return ptd::eigh_debug_phase_cycles(key - 100);
  if (key >= 210 && key < 218) return ptd::lowrank_debug_get(key - 210);
  return ptd::gemm_tc_last_launch_info(key);
}

}  // extern "C"
