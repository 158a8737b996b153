// Memory-bound helper kernels of the hot path (all HBM-bound; coalesced, vectorised where aligned):
//   split_rows      fp32/bf16 [rows][cols] -> bf16 h/m/l segments (bf16x3 split) or padded bf16 copy
//   colsum          column sums of the activation block (falor use_mean, F:161)
//   cov_finalize    scale by 1/steps, optional centring, mirror lower->upper, damping (F:192-205, D:158-160)
//   nsr / kl        rank-search metrics (U/l:10-22, U/l:48-63)
#include "elementwise.cuh"

#include <cuda_bf16.h>

namespace ptd {

namespace {

__device__ __forceinline__ void split3(float x, __nv_bfloat16& h, __nv_bfloat16& m,
                                       __nv_bfloat16& l) {
  h = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h);
  m = __float2bfloat16_rn(r1);
  l = __float2bfloat16_rn(r1 - __bfloat162float(m));
}

// One thread per (row, 8-column group). dst row pitch ldd (multiple of 8); columns >= cols are
// zero filled so the padded copy is safe to feed to TMA.
template <typename T, int NSEG>
__global__ void split_rows_kernel(const T* __restrict__ src, long long lds, long long rows,
                                  int cols, float sub_scale, const float* __restrict__ sub,
                                  __nv_bfloat16* __restrict__ dst, long long ldd,
                                  long long seg_stride) {
  const int groups = static_cast<int>(ldd / 8);
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * groups) return;
  const long long r = idx / groups;
  const int c0 = static_cast<int>(idx % groups) * 8;
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    float v = 0.f;
    if (c < cols) {
      v = static_cast<float>(src[r * lds + c]);
      if (sub != nullptr) v -= sub_scale * sub[c];
    }
    x[j] = v;
  }
  __align__(16) __nv_bfloat16 o[NSEG][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (NSEG == 1) {
      o[0][j] = __float2bfloat16_rn(x[j]);
    } else {
      split3(x[j], o[0][j], o[1 % NSEG][j], o[2 % NSEG][j]);
    }
  }
#pragma unroll
  for (int s = 0; s < NSEG; ++s)
    *reinterpret_cast<uint4*>(dst + s * seg_stride + r * ldd + c0) =
        *reinterpret_cast<const uint4*>(o[s]);
}

// Column sums: grid (col blocks of 32*4, row chunks); fp32 atomics into out[cols].
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ src, long long lds, long long rows, int cols,
                              float scale, const float* __restrict__ sub, float sub_scale,
                              float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const long long rows_per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = rows_per * blockIdx.y;
  const long long r1 = min(rows, r0 + rows_per);
  if (c >= cols) return;
  float acc = 0.f, comp = 0.f;  // Kahan: long columns at fp32
  for (long long r = r0; r < r1; ++r) {
    const float v = static_cast<float>(src[r * lds + c]) - comp;
    const float t = acc + v;
    comp = (t - acc) - v;
    acc = t;
  }
  if (r1 > r0) {
    float v = acc * scale;
    if (sub != nullptr) v -= sub_scale * sub[c] * static_cast<float>(r1 - r0) * scale;
    atomicAdd(out + c, v);
  }
}

// 32x32 tile of the lower triangle (bi >= bj): scale, centre, write (i,j) and the mirror (j,i).
__global__ void cov_scale_mirror_kernel(float* __restrict__ C, long long ldc, int d,
                                        const float* __restrict__ colsum, float inv_steps,
                                        int use_mean) {
  __shared__ float tile[32][33];
  // linear block index over lower-triangular tile pairs
  const int t = blockIdx.x;
  int bi = static_cast<int>((sqrtf(8.f * static_cast<float>(t) + 1.f) - 1.f) * 0.5f);
  while (bi * (bi + 1) / 2 > t) --bi;
  while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
  const int bj = t - bi * (bi + 1) / 2;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int yy = ty; yy < 32; yy += 8) {
    const int i = bi * 32 + yy, j = bj * 32 + tx;
    float v = 0.f;
    if (i < d && j < d) {
      // the source of truth is the lower triangle; inside a diagonal tile read the mirrored
      // element when (i, j) lies above the diagonal
      const int si = (i >= j) ? i : j, sj = (i >= j) ? j : i;
      v = C[static_cast<long long>(si) * ldc + sj] * inv_steps;
      if (use_mean) v -= (colsum[si] * inv_steps) * (colsum[sj] * inv_steps);
    }
    tile[yy][tx] = v;
  }
  __syncthreads();
  for (int yy = ty; yy < 32; yy += 8) {
    const int i = bi * 32 + yy, j = bj * 32 + tx;
    if (i < d && j < d) C[static_cast<long long>(i) * ldc + j] = tile[yy][tx];
  }
  if (bi != bj) {
    for (int yy = ty; yy < 32; yy += 8) {
      const int i = bj * 32 + yy, j = bi * 32 + tx;  // mirrored tile, row = old column
      if (i < d && j < d) C[static_cast<long long>(i) * ldc + j] = tile[tx][yy];
    }
  }
}

// Single block: damp = factor * mean(diag) ; diag += damp ; optionally report trace.
__global__ void cov_damp_kernel(float* __restrict__ C, long long ldc, int d, float factor,
                                float* __restrict__ damp_out) {
  __shared__ double red[32];
  __shared__ float damp_s;
  double acc = 0.0;
  for (int i = threadIdx.x; i < d; i += blockDim.x) acc += C[static_cast<long long>(i) * ldc + i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
    damp_s = static_cast<float>(static_cast<double>(factor) * (s / d));
    if (damp_out) *damp_out = damp_s;
  }
  __syncthreads();
  const float damp = damp_s;
  if (factor != 0.f)
    for (int i = threadIdx.x; i < d; i += blockDim.x) C[static_cast<long long>(i) * ldc + i] += damp;
}

// ------------------------------------------------------------------ metrics
// Per-channel sums over rows for NSR: out[c] = {sum y, sum y^2, sum (x-y)^2} in fp64.
// Rows are split over blockIdx.y; channels over blockIdx.x (coalesced along channels).
template <typename T>
__global__ void nsr_partial_kernel(const T* __restrict__ x, const T* __restrict__ y,
                                   long long rows, long long ch, double* __restrict__ part) {
  const long long c = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= ch) return;
  const long long rows_per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = rows_per * blockIdx.y, r1 = min(rows, r0 + rows_per);
  double sy = 0.0, syy = 0.0, sd = 0.0;
  for (long long r = r0; r < r1; ++r) {
    const double yv = static_cast<double>(static_cast<float>(y[r * ch + c]));
    const double xv = static_cast<double>(static_cast<float>(x[r * ch + c]));
    sy += yv;
    syy += yv * yv;
    sd += (xv - yv) * (xv - yv);
  }
  atomicAdd(part + 3 * c + 0, sy);
  atomicAdd(part + 3 * c + 1, syy);
  atomicAdd(part + 3 * c + 2, sd);
}

// mean_c( mean_r (x-y)^2 / (var_unbiased_r(y) + eps) ) -> out[0]
__global__ void nsr_final_kernel(const double* __restrict__ part, long long rows, long long ch,
                                 double eps, float* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (long long c = threadIdx.x; c < ch; c += blockDim.x) {
    const double sy = part[3 * c], syy = part[3 * c + 1], sd = part[3 * c + 2];
    const double mean = sy / rows;
    // rows == 1 -> 0/0 = NaN, as torch.std(unbiased) gives
    double var = (syy - sy * mean) / static_cast<double>(rows - 1);
    if (var < 0.0) var = 0.0;  // rounding only; NaN (rows == 1) stays NaN
    const double msd = sd / rows;
    acc += msd / (var + eps);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
    out[0] = static_cast<float>(s / ch);
  }
}

__device__ __forceinline__ double block_reduce(double v, double* red, bool is_max) {
  for (int o = 16; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmax(v, w) : v + w;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) r = is_max ? fmax(r, red[w]) : r + red[w];
  return r;
}

// One block per row: symmetric-max KL of softmax(s) vs softmax(t) over the last dim,
// accumulated as a mean over rows into out[0] (U/l:48-63).
template <typename T>
__global__ void kl_rows_kernel(const T* __restrict__ s, const T* __restrict__ t, long long rows,
                               long long ch, float* __restrict__ out) {
  __shared__ double red[32];
  const long long r = blockIdx.x;
  const T* sr = s + r * ch;
  const T* tr = t + r * ch;
  double ms = -1e300, mt = -1e300;
  for (long long c = threadIdx.x; c < ch; c += blockDim.x) {
    ms = fmax(ms, static_cast<double>(static_cast<float>(sr[c])));
    mt = fmax(mt, static_cast<double>(static_cast<float>(tr[c])));
  }
  ms = block_reduce(ms, red, true);
  mt = block_reduce(mt, red, true);
  double zs = 0.0, zt = 0.0;
  for (long long c = threadIdx.x; c < ch; c += blockDim.x) {
    zs += exp(static_cast<double>(static_cast<float>(sr[c])) - ms);
    zt += exp(static_cast<double>(static_cast<float>(tr[c])) - mt);
  }
  zs = block_reduce(zs, red, false);
  zt = block_reduce(zt, red, false);
  const double ls = ms + log(zs), lt = mt + log(zt);
  double kl_ts = 0.0, kl_st = 0.0;  // KL(t||s), KL(s||t)
  for (long long c = threadIdx.x; c < ch; c += blockDim.x) {
    const double a = static_cast<double>(static_cast<float>(sr[c])) - ls;  // log p_s
    const double b = static_cast<double>(static_cast<float>(tr[c])) - lt;  // log p_t
    kl_ts += exp(b) * (b - a);
    kl_st += exp(a) * (a - b);
  }
  kl_ts = block_reduce(kl_ts, red, false);
  kl_st = block_reduce(kl_st, red, false);
  if (threadIdx.x == 0) atomicAdd(out, static_cast<float>(fmax(kl_ts, kl_st) / rows));
}

template <typename T>
int split_rows_t(const T* src, long long lds, long long rows, int cols, const float* sub,
                 float sub_scale, __nv_bfloat16* dst, long long ldd, int nseg,
                 long long seg_stride, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return 0;
  if (ldd % 8) return -22;
  const long long total = rows * (ldd / 8);
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  if (blocks > 0x7fffffffLL) return -22;
  if (nseg == 1)
    split_rows_kernel<T, 1><<<static_cast<unsigned>(blocks), threads, 0, st>>>(
        src, lds, rows, cols, sub_scale, sub, dst, ldd, seg_stride);
  else
    split_rows_kernel<T, 3><<<static_cast<unsigned>(blocks), threads, 0, st>>>(
        src, lds, rows, cols, sub_scale, sub, dst, ldd, seg_stride);
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace

int split_rows(const void* src, int src_is_bf16, long long lds, long long rows, int cols,
               const float* sub, float sub_scale, __nv_bfloat16* dst, long long ldd, int nseg,
               long long seg_stride, cudaStream_t st) {
  if (src_is_bf16)
    return split_rows_t(static_cast<const __nv_bfloat16*>(src), lds, rows, cols, sub, sub_scale,
                        dst, ldd, nseg, seg_stride, st);
  return split_rows_t(static_cast<const float*>(src), lds, rows, cols, sub, sub_scale, dst, ldd,
                      nseg, seg_stride, st);
}

int colsum(const void* src, int src_is_bf16, long long lds, long long rows, int cols, float scale,
           const float* sub, float sub_scale, float* out, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return 0;
  const int threads = 128;
  dim3 grid((cols + threads - 1) / threads,
            static_cast<unsigned>(min(static_cast<long long>(512), (rows + 255) / 256)));
  if (src_is_bf16)
    colsum_kernel<<<grid, threads, 0, st>>>(static_cast<const __nv_bfloat16*>(src), lds, rows, cols,
                                            scale, sub, sub_scale, out);
  else
    colsum_kernel<<<grid, threads, 0, st>>>(static_cast<const float*>(src), lds, rows, cols, scale,
                                            sub, sub_scale, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int cov_finalize(float* C, long long ldc, int d, const float* colsum_v, float inv_steps,
                 int use_mean, float damp_factor, float* damp_out, cudaStream_t st) {
  if (d <= 0) return -22;
  if (use_mean && colsum_v == nullptr) return -22;
  const int nb = (d + 31) / 32;
  const long long tiles = static_cast<long long>(nb) * (nb + 1) / 2;
  cov_scale_mirror_kernel<<<static_cast<unsigned>(tiles), dim3(32, 8), 0, st>>>(
      C, ldc, d, colsum_v, inv_steps, use_mean);
  cov_damp_kernel<<<1, 1024, 0, st>>>(C, ldc, d, damp_factor, damp_out);
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int nsr_metric(const void* x, const void* y, int is_bf16, long long rows, long long ch, double eps,
               double* scratch /* 3*ch doubles */, float* out, cudaStream_t st) {
  if (rows <= 0 || ch <= 0) return -22;
  cudaMemsetAsync(scratch, 0, sizeof(double) * 3 * ch, st);
  const int threads = 128;
  dim3 grid(static_cast<unsigned>((ch + threads - 1) / threads),
            static_cast<unsigned>(max(1LL, min(static_cast<long long>(256), rows / 64))));
  if (is_bf16)
    nsr_partial_kernel<<<grid, threads, 0, st>>>(static_cast<const __nv_bfloat16*>(x),
                                                 static_cast<const __nv_bfloat16*>(y), rows, ch,
                                                 scratch);
  else
    nsr_partial_kernel<<<grid, threads, 0, st>>>(static_cast<const float*>(x),
                                                 static_cast<const float*>(y), rows, ch, scratch);
  nsr_final_kernel<<<1, 1024, 0, st>>>(scratch, rows, ch, eps, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

int kl_metric(const void* s, const void* t, int is_bf16, long long rows, long long ch, float* out,
              cudaStream_t st) {
  if (rows <= 0 || ch <= 0) return -22;
  cudaMemsetAsync(out, 0, sizeof(float), st);
  if (is_bf16)
    kl_rows_kernel<<<static_cast<unsigned>(rows), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(s), static_cast<const __nv_bfloat16*>(t), rows, ch, out);
  else
    kl_rows_kernel<<<static_cast<unsigned>(rows), 256, 0, st>>>(
        static_cast<const float*>(s), static_cast<const float*>(t), rows, ch, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace ptd
