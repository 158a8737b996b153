// Internal interface of the decomposed-layer forward (lowrank.cu). Not part of the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace ptd {

size_t lowrank_workspace_bytes(int dtype_is_bf16, long long n, int in_f, int k, int out_f);

// Y[n][out] = (X[n][in] W1[k][in]^T) W2[out][k]^T + bias[out]; X, W1, W2, Y all bf16 or all fp32.
int lowrank_forward(const void* X, long long ldx, const void* W1, long long ldw1, const void* W2,
                    long long ldw2, const float* bias, void* Y, long long ldy, int dtype_is_bf16,
                    long long n, int in_f, int k, int out_f, void* ws, size_t ws_bytes,
                    cudaStream_t st);

// Runtime knobs 0..9 (see lowrank.cu); reached through ptdeco_debug_set keys 200..209. Knob 9
// switches on the fused kernel's phase-cycle counters, read back through ptdeco_debug_get 210..217.
void lowrank_debug_set(int key, long long value);
long long lowrank_debug_get(int k);

}  // namespace ptd
