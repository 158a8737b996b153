/* ptdeco_b200 C-ABI: the drop-in boundary of the falor/dwain decomposition hot path.
 *
 * The reference (TCLResearchEurope/ptdeco) is pure Python on torch; it has no FFI of its own.
 * Each entry point below replaces one torch library call site of the reference (cited as
 * file:line relative to the reference tree, F = src/ptdeco/falor/decomposition.py,
 * D = src/ptdeco/dwain/decomposition.py, U/l = src/ptdeco/utils/losses_primitives.py) and is what
 * a maintainer would bind with ctypes from those call sites.
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless named host_*
 *   - row-major everywhere; ld* = row pitch in elements
 *   - dtype codes: PTDECO_F32 = 0, PTDECO_BF16 = 1
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing syncs
 *     unless stated
 *   - return 0 on success, negative errno-style code on failure (ptdeco_strerror); no exceptions
 *   - the caller owns every buffer; scratch comes from the *_workspace_bytes queries. A workspace
 *     (and, for ptdeco_lowrank_forward's decode path, the control words inside it) belongs to ONE
 *     call at a time: concurrent calls on different streams need different workspaces
 *   - entry points are re-entrant across streams under that rule; process-wide state is limited to
 *     a lazily resolved driver entry point, per-device kernel attributes, and the debug knobs
 *     below. Numerics-affecting options are per call (the `flags` of the *_ex variants), never
 *     global
 */
#ifndef PTDECO_B200_H_
#define PTDECO_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTDECO_F32 0
#define PTDECO_BF16 1

/* flags of the *_ex entry points */
#define PTDECO_FLAG_DETERMINISTIC 1u /* no split-K: every output element is accumulated by one CTA in
                                        a fixed order (default: short-and-wide reductions are split
                                        over CTAs and combined with fp32 red.add, whose arrival order
                                        varies from run to run) */

int ptdeco_version(void);
const char* ptdeco_strerror(int code);

/* ---- K1 (+K1b): covariance accumulate --------------------------------------------------------
 * Replaces  F:160-161  Eyyt += einsum("bp,bq->pq", y, y) / N ; Ey += y.mean(0)
 *           D:152      Eyyt += einsum("bp,bq->pq", y, y) / N
 * C[d][ldc] (fp32) += alpha * (Y - 1*sub^T)^T (Y - 1*sub^T); only tiles touching the LOWER
 * triangle are updated (the strict upper triangle is unspecified until ptdeco_cov_finalize).
 * colsum[d] (fp32, may be NULL) += alpha * column sums of (Y - 1*sub^T).
 * sub[d] (fp32, may be NULL) is subtracted from every row first (layer bias, when Y is the hooked
 * layer output rather than x @ W^T). Y is [n_tokens][ldy] of `dtype`.
 * bf16 input: products exact, fp32 accumulation. fp32 input: bf16x3 split (fp32-grade products). */
size_t ptdeco_syrk_workspace_bytes(int dtype, long long n_tokens, int d);
int ptdeco_syrk_accumulate(const void* Y, int dtype, long long n_tokens, int d, long long ldy,
                           const float* sub, float* C, long long ldc, float* colsum, float alpha,
                           void* workspace, size_t workspace_bytes, void* stream);

int ptdeco_syrk_accumulate_ex(const void* Y, int dtype, long long n_tokens, int d, long long ldy,
                              const float* sub, float* C, long long ldc, float* colsum, float alpha,
                              void* workspace, size_t workspace_bytes, void* stream, unsigned flags);

/* ---- K2: covariance finalize -----------------------------------------------------------------
 * Replaces  F:192-205 (Ey,Eyyt /= steps; cov = Eyyt - outer(Ey,Ey); diag += 0.01*mean(diag))
 *           D:158-160, D:207, D:242
 * In place: C *= 1/n_steps; if use_mean: C -= Ey Ey^T with Ey = colsum/n_steps; lower triangle is
 * mirrored to the upper; diag += damp_factor * mean(diag). damp_out (device float, may be NULL)
 * receives the damping value that was added. */
int ptdeco_cov_finalize(float* C, long long ldc, int d, const float* colsum, int n_steps,
                        int use_mean, float damp_factor, float* damp_out, void* stream);

/* ---- generic tensor-core contraction (engine entry; used by K4 and by tests) -------------------
 * C[M][ldc] (fp32, or bf16 when c_dtype = PTDECO_BF16) = or += alpha * op(A) op(B) (+ bias[N])
 *   a_mn_major = 1: A is stored [K][lda] with M contiguous   (contraction index = row)
 *   a_mn_major = 0: A is stored [M][lda] with K contiguous
 *   same for B with N. dtype of A and B: bf16 used directly, fp32 split bf16x3 into workspace. */
size_t ptdeco_gemm_workspace_bytes(int a_dtype, int b_dtype, int M, int N, int K);
int ptdeco_gemm(const void* A, int a_dtype, int a_mn_major, long long lda, const void* B,
                int b_dtype, int b_mn_major, long long ldb, int M, int N, int K, float alpha,
                const float* bias, void* C, int c_dtype, long long ldc, int accumulate,
                void* workspace, size_t workspace_bytes, void* stream);

int ptdeco_gemm_ex(const void* A, int a_dtype, int a_mn_major, long long lda, const void* B,
                   int b_dtype, int b_mn_major, long long ldb, int M, int N, int K, float alpha,
                   const float* bias, void* C, int c_dtype, long long ldc, int accumulate,
                   void* workspace, size_t workspace_bytes, void* stream, unsigned flags);

/* ---- K3: symmetric eigendecomposition ------------------------------------------------------------
 * Replaces  F:207 / D:162  _, u = torch.linalg.eigh(cov)   (and the uk = u[:, d-k:] slice of F:346,
 * D:425): A[d][lda] fp32 symmetric (lower triangle authoritative, torch's UPLO='L'), not modified.
 * evals[d] receives ALL eigenvalues ascending; U[d][ldu] receives in column c the eigenvector of
 * eigenvalue number d-k+c, i.e. the k largest in ascending order (k = d gives torch's full `u`).
 * d <= 32: one-CTA fp64 Jacobi. d > 32: Householder tridiagonalisation -- while the trailing block
 * exceeds the chip's shared memory (d > ~2560) a cooperative blocked panel kernel with tcgen05
 * trailing updates, then ONE cooperative launch that keeps the trailing block resident in the
 * shared memory of all SMs (one grid-wide exchange per column) --, fp64 multisection +
 * twisted-factorisation eigenvectors, tcgen05 compact-WY back-transformation of the k wanted vectors. */
size_t ptdeco_eigh_workspace_bytes(int d, int k);
int ptdeco_eigh(const float* A, int d, long long lda, int k, float* evals, float* U, long long ldu,
                void* workspace, size_t workspace_bytes, void* stream);

int ptdeco_eigh_ex(const float* A, int d, long long lda, int k, float* evals, float* U,
                   long long ldu, void* workspace, size_t workspace_bytes, void* stream,
                   unsigned flags);

/* ---- K7: decomposed-layer forward -----------------------------------------------------------------
 * Replaces the forward of the nn.Sequential(Linear(in->k, no bias), Linear(k->out, bias)) built at
 * F:84-95 / D:74-85 (and its 1x1-conv twin on NHWC rows):
 * Y[n][ldy] = (X[n][ldx] W1[k][ldw1]^T) W2[out][ldw2]^T + bias[out]; X, W1, W2, Y of `dtype`,
 * bias fp32 or NULL.
 * Three kernels behind the one entry point (bf16): n <= 128 with >= 2 MB of factors runs the
 * single-launch weight-streaming kernel (needs the workspace: fp32 partial sums, bf16 intermediate,
 * grid-barrier / ticket words; zeroed by the call); k <= 256 runs the fused kernel that keeps the
 * [128, k] intermediate on chip (no workspace); everything else, and fp32, runs two tcgen05 GEMMs
 * with the intermediate in the workspace. */
size_t ptdeco_lowrank_workspace_bytes(int dtype, long long n, int in_features, int k,
                                      int out_features);
int ptdeco_lowrank_forward(const void* X, long long ldx, const void* W1, long long ldw1,
                           const void* W2, long long ldw2, const float* bias, void* Y,
                           long long ldy, int dtype, long long n, int in_features, int k,
                           int out_features, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K6: rank-search metrics -------------------------------------------------------------------
 * ptdeco_nsr_metric replaces U/l:10-22 calc_per_channel_noise_to_signal_ratio for channel-last
 * inputs viewed as [rows][channels]: out[0] = mean_c( mean_r (x-y)^2 / (var_unbiased_r(y) + eps) ).
 * ptdeco_kl_metric replaces U/l:57-63 calc_kl_loss for logits [rows][classes]:
 * out[0] = mean_r max(KL(t||s), KL(s||t)). `out` is a device float. */
size_t ptdeco_nsr_workspace_bytes(long long channels);
int ptdeco_nsr_metric(const void* x, const void* y, int dtype, long long rows, long long channels,
                      double eps, void* workspace, size_t workspace_bytes, float* out, void* stream);
int ptdeco_kl_metric(const void* student, const void* teacher, int dtype, long long rows,
                     long long classes, float* out, void* stream);

/* ---- debug knobs (descriptor sweeps, forced tile shapes); not for production use -------------- */
void ptdeco_debug_set(int key, long long value);
long long ptdeco_debug_get(int key);

#ifdef __cplusplus
}
#endif
#endif /* PTDECO_B200_H_ */
