// Internal interface of the symmetric eigensolver (eigh.cu). Not part of the C-ABI.
//
//   d <= 32   one-CTA parallel cyclic Jacobi in fp64 shared memory (all eigenpairs, ~1e-15)
//   d  > 32   (1) Householder tridiagonalisation in two regimes:
//                 trailing block > ~2560 rows: blocked -- a cooperative persistent panel kernel
//                 (2 grid barriers per column: column update + reflector, then the symv against
//                 the trailing matrix in L2 / HBM) and a tcgen05 rank-2nb trailing update
//                 A -= V W^T + W V^T (bf16x3 split, fp32-grade);
//                 trailing block <= ~2560 rows (the whole matrix for d <= 2560): ONE cooperative
//                 launch with the block resident in the shared memory of all SMs, the rank-2
//                 update fused into the next column's symv pass, one exchange per column;
//             (2) tridiagonal eigenproblem in fp64: splitting, warp-parallel multisection on the
//                 Sturm count, eigenvectors by twisted factorisation (one thread per vector),
//                 Gram-Schmidt inside numerically tight clusters;
//             (3) back-transformation of the k wanted vectors with compact-WY block reflectors,
//                 two tcgen05 GEMMs per panel.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace ptd {

size_t eigh_workspace_bytes(int d, int k);

// A [d][lda] fp32 symmetric (lower triangle authoritative, like torch.linalg.eigh's UPLO='L');
// not modified. evals[d] ascending. U [d][ldu]: column c holds the eigenvector of eigenvalue
// number d-k+c (the k largest, ascending). Everything is enqueued on `st`; nothing syncs.
// flags: bit 0 = deterministic (no split-K in the tensor-core stages).
int eigh(const float* A, int d, long long lda, int k, float* evals, float* U, long long ldu,
         void* ws, size_t ws_bytes, cudaStream_t st, unsigned flags = 0);

// Debug: per-phase cycle counters of the panel kernel (CTA 0), see eigh.cu.
void eigh_debug_profile(int enable);
long long eigh_debug_phase_cycles(int k);
// Debug: trailing-matrix size from which the panel symv reads only the lower triangle (0 = never).
void eigh_debug_sym_min_m(int m);
// Debug: resident (shared-memory) tridiagonalisation on/off, its target rows per CTA (<= 0 keeps),
// and the largest d solved by the one-CTA Jacobi kernel (< 0 keeps).
void eigh_debug_resident(int enable, int rows_target, int jacobi_max);
// Debug: smallest d whose multisection uses 4 lanes per eigenvalue (<= 0: never, the default).
void eigh_debug_bisect_narrow(int d);
// Debug: 1 = ratio-form (division) Sturm count instead of the product form.
void eigh_debug_sturm_ratio(int on);
void eigh_debug_small(int on);

}  // namespace ptd
