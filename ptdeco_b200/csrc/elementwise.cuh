// Internal interface of the memory-bound helper kernels (elementwise.cu). Not part of the C-ABI.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace ptd {

// dst[seg][r][c] = split_seg(src[r][c] - sub_scale*sub[c]); columns in [cols, ldd) are zeroed.
// nseg = 1: plain bf16 copy/convert; nseg = 3: bf16x3 split (h, m, l).
int split_rows(const void* src, int src_is_bf16, long long lds, long long rows, int cols,
               const float* sub, float sub_scale, __nv_bfloat16* dst, long long ldd, int nseg,
               long long seg_stride, cudaStream_t st);

// out[c] += scale * sum_r (src[r][c] - sub_scale*sub[c])
int colsum(const void* src, int src_is_bf16, long long lds, long long rows, int cols, float scale,
           const float* sub, float sub_scale, float* out, cudaStream_t st);

// In place on the fp32 accumulator (lower triangle authoritative):
//   C = C*inv_steps ; if use_mean: C -= Ey Ey^T with Ey = colsum*inv_steps ; mirror ; diag += damp
int cov_finalize(float* C, long long ldc, int d, const float* colsum_v, float inv_steps,
                 int use_mean, float damp_factor, float* damp_out, cudaStream_t st);

int nsr_metric(const void* x, const void* y, int is_bf16, long long rows, long long ch, double eps,
               double* scratch, float* out, cudaStream_t st);
int kl_metric(const void* s, const void* t, int is_bf16, long long rows, long long ch, float* out,
              cudaStream_t st);

}  // namespace ptd
