// Internal interface of the tcgen05 GEMM engine (gemm_tc.cu). Not part of the C-ABI.
//
// One warp-specialised, persistent sm_100a kernel computes
//     C[M,N] (+)= alpha * sum_{pairs (sa,sb)} opA_sa^T-or-not * opB_sb     (+ bias[N])
// with bf16 operands staged by TMA into 128B-swizzled shared memory, tcgen05.mma accumulating
// fp32 in TMEM, and an epilogue that reads TMEM with tcgen05.ld. Every dense contraction of the
// hot path is an instance of it:
//   SYRK  C += Y^T Y / N           A = B = Y[K=tokens, MN=features]   (MN-major / MN-major)
//   factor W1 = Uk^T W             A = Uk[K=out, M=k], B = W[K=out, N=in] (MN / MN)
//   forward Y = X W^T              A = X[M=tokens, K=in], B = W[N=out, K=in] (K / K)
//   eigh trailing updates / back-transform: mixes of the above
// fp32-grade products use the "bf16x3" split (x = h + m + l, six h/m/l segment pairs).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptd {

// A bf16 operand in global memory, seen as nseg stacked 2-D row-major arrays.
//   mn_major = 1: each segment is [K rows][MN cols]   (the reduction index is the row)
//   mn_major = 0: each segment is [MN rows][K cols]   (the reduction index is the column)
struct GemmOperand {
  const __nv_bfloat16* ptr;
  int mn_major;
  long long ld;          // row pitch in elements (multiple of 8)
  int nseg;              // 1 (plain bf16) or 3 (h/m/l split of fp32)
  long long seg_stride;  // elements between segments
};

struct GemmEpilogue {
  float alpha = 1.f;
  float* C = nullptr;  // fp32 output, row-major [M][ldc]
  long long ldc = 0;
  __nv_bfloat16* Cb = nullptr;  // optional bf16 output (store mode only)
  long long ldcb = 0;
  // Optional bf16x3 split output (store mode): Cs[seg][M][ldcs], seg stride = cs_seg.
  __nv_bfloat16* Cs = nullptr;
  long long ldcs = 0, cs_seg = 0;
  const float* bias = nullptr;  // optional, length N, added after alpha scaling
  int accumulate = 0;           // 0: C = result, 1: C += result (red.global.add)
  int lower_only = 0;           // 1: only tiles touching the lower triangle (M == N)
  int deterministic = 0;        // 1: no split-K (every output element has one producer, fixed order)
};

// Returns 0 on success, negative errno-style on bad arguments / launch failure.
int gemm_tc(const GemmOperand& A, const GemmOperand& B, int M, int N, int K, int full_pairs,
            const GemmEpilogue& ep, cudaStream_t stream);

// 2-D bf16 tensor map {inner, rows} with row pitch ld (elements), box {64, box_rows}, 128B swizzle,
// zero fill out of bounds. Shared with the fused low-rank kernel. 0 on success.
int make_tma_2d_bf16(CUtensorMap* map, const void* ptr, long long inner, long long rows,
                     long long ld, int box_rows);
int device_sm_count();

// Debug overrides used by the descriptor sweep in tests/tools (0 = default).
void gemm_tc_debug_set(int key, long long value);
long long gemm_tc_last_launch_info(int key);

}  // namespace ptd
