// Latencies that bound the one-SM / few-row kernels of eigh.cu, measured on the part itself:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/sm_lat tools/micro/sm_latency_bench.cu && /tmp/sm_lat
// One CTA of 512 threads. Cycles (clock64) per: dependent shared-memory load (4 B and 16 B),
// dependent shuffle, dependent FFMA, __syncthreads with all 16 warps arriving together, and one
// "symv pass" shaped like sytd2_small_kernel's (96 x 96, 16-byte vectors, 8 lanes per row).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) lat_kernel(long long* out, int n) {
  __shared__ int chase[1024];
  __shared__ __align__(16) float As[96 * 100];
  __shared__ __align__(16) float v[132];
  __shared__ float pvec[132];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 1024; i += 512) chase[i] = (i + 33) & 1023;
  for (int i = tid; i < 96 * 100; i += 512) As[i] = 1.f / (1 + i);
  for (int i = tid; i < 132; i += 512) v[i] = 0.5f;
  __syncthreads();
  long long t0, t1;
  // (1) dependent LDS.32
  int p = lane;
  t0 = clock64();
  for (int i = 0; i < n; ++i) p = chase[p];
  t1 = clock64();
  if (tid == 0) out[0] = (t1 - t0);
  // (2) dependent LDS.128
  int q = (lane * 4) & 1020;
  t0 = clock64();
  for (int i = 0; i < n; ++i) {
    const int4 x = *reinterpret_cast<const int4*>(chase + q);
    q = x.x & 1020;
  }
  t1 = clock64();
  if (tid == 0) out[1] = (t1 - t0);
  // (3) dependent shuffle
  float s = p + q;
  t0 = clock64();
  for (int i = 0; i < n; ++i) s += __shfl_xor_sync(0xffffffffu, s, 1);
  t1 = clock64();
  if (tid == 0) out[2] = (t1 - t0);
  // (4) dependent FFMA
  float f = s;
  t0 = clock64();
  for (int i = 0; i < n; ++i) f = fmaf(f, 1.0001f, 0.5f);
  t1 = clock64();
  if (tid == 0) out[3] = (t1 - t0);
  // (5) __syncthreads, all warps together
  __syncthreads();
  t0 = clock64();
  for (int i = 0; i < n; ++i) __syncthreads();
  t1 = clock64();
  if (tid == 0) out[4] = (t1 - t0);
  // (6) one symv pass like the small kernel's: rows r0 = warp*4 + rg (+64), 8 lanes per row
  const int rg = lane >> 3, ch = lane & 7;
  float acc = 0.f;
  __syncthreads();
  t0 = clock64();
  for (int it = 0; it < n; ++it) {
    for (int r0 = warp * 4; r0 < 96; r0 += 64) {
      const int r = r0 + rg;
      float part = 0.f;
      if (r < 96) {
        const float* arow = As + r * 100;
        for (int c = ch * 4; c < 96; c += 32) {
          const float4 a4 = *reinterpret_cast<const float4*>(arow + c);
          const float4 v4 = *reinterpret_cast<const float4*>(v + c);
          part += a4.x * v4.x + a4.y * v4.y + a4.z * v4.z + a4.w * v4.w;
        }
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      if (ch == 0 && r < 96) pvec[r] = part;
    }
    __syncthreads();
    acc += pvec[lane];
    v[lane] = acc * 1e-9f + 0.5f;  // keep the passes dependent
    __syncthreads();
  }
  t1 = clock64();
  if (tid == 0) out[5] = (t1 - t0);
  if (p + q + f + acc == 12345.678f) out[7] = 1;
}

int main() {
  long long* d;
  cudaMalloc(&d, 8 * sizeof(long long));
  const int n = 2000;
  lat_kernel<<<1, 512>>>(d, n);
  lat_kernel<<<1, 512>>>(d, n);
  cudaDeviceSynchronize();
  long long h[8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[6] = {"dependent LDS.32", "dependent LDS.128", "dependent SHFL + FADD", "dependent FFMA",
                          "__syncthreads (16 warps)", "symv pass 96x96 + 2 barriers"};
  for (int i = 0; i < 6; ++i) printf("%-32s %8.1f cycles\n", names[i], double(h[i]) / n);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
