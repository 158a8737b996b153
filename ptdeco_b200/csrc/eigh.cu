// Symmetric eigensolver for sm_100a. See eigh.cuh for the algorithm outline.
//
// Replaces torch.linalg.eigh at the reference's F:207 / D:162. Only the eigenvectors of the k
// largest eigenvalues are back-transformed (the rank search never slices more than
// ceil(min(in,out)/2) of them).
#include "eigh.cuh"

#include <cuda_bf16.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>

#include "gemm_tc.cuh"

namespace ptd {

namespace {

constexpr int JACOBI_MAX = 96;      // largest d the one-CTA Jacobi kernel can hold
constexpr int JACOBI_DEFAULT = 32;  // largest d it is USED for: beyond, the resident path is faster
                                    // (B200: d = 96 0.72 vs 4.85 ms; at d = 32 0.28 vs 0.40 ms, but
                                    // the fp64 Jacobi's 1e-8 orthogonality is what keeps the
                                    // reference's full-rank reconstruction test under its 1e-6)
constexpr int NB = 64;  // Householder panel width
constexpr int PANEL_THREADS = 512;
constexpr int PANEL_WARPS = PANEL_THREADS / 32;
constexpr int GP_STRIDE = 2 * NB + 2;  // floats per CTA in the partial-dot exchange buffer

inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void split3(float x, __nv_bfloat16& h, __nv_bfloat16& m,
                                       __nv_bfloat16& l) {
  h = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h);
  m = __float2bfloat16_rn(r1);
  l = __float2bfloat16_rn(r1 - __bfloat162float(m));
}

// ============================================================================ small matrices
// Parallel cyclic Jacobi (round-robin ordering) on one CTA, everything in fp64 shared memory.
__global__ void __launch_bounds__(512, 1)
jacobi_kernel(const float* __restrict__ A, long long lda, int d, int k, float* __restrict__ evals,
              float* __restrict__ U, long long ldu) {
  extern __shared__ double smd[];
  const int ld = d + 1;
  const int n = d + (d & 1);  // players of the tournament (one dummy when d is odd)
  const int half = n / 2;
  double* a = smd;             // [d][ld]
  double* v = a + d * ld;      // [d][ld]
  double* cs = v + d * ld;     // [half][2]
  int* pp = reinterpret_cast<int*>(cs + 2 * half);
  int* qq = pp + half;
  int* rnk = qq + half;        // [d]
  __shared__ double red[32];
  __shared__ double s_off, s_diag;
  const int tid = threadIdx.x, nt = blockDim.x;

  for (int idx = tid; idx < d * d; idx += nt) {
    const int r = idx / d, c = idx % d;
    const int hi = r > c ? r : c, lo = r > c ? c : r;
    a[r * ld + c] = static_cast<double>(A[static_cast<long long>(hi) * lda + lo]);
    v[r * ld + c] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();

  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int idx = tid; idx < d * d; idx += nt) {
      const int r = idx / d, c = idx % d;
      const double x = a[r * ld + c];
      if (r == c) dg += x * x; else off += x * x;
    }
    off = warp_sum(off);
    dg = warp_sum(dg);
    if ((tid & 31) == 0) red[tid >> 5] = off;
    __syncthreads();
    if (tid == 0) { double s = 0.0; for (int w = 0; w < (nt >> 5); ++w) s += red[w]; s_off = s; }
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = dg;
    __syncthreads();
    if (tid == 0) { double s = 0.0; for (int w = 0; w < (nt >> 5); ++w) s += red[w]; s_diag = s; }
    __syncthreads();
    if (s_off <= 1e-31 * s_diag || s_off == 0.0) break;  // uniform across the CTA

    for (int round = 0; round < n - 1; ++round) {
      if (tid < half) {
        int p, q;
        if (tid == 0) { p = n - 1; q = round; }
        else { p = (round + tid) % (n - 1); q = (round - tid + (n - 1)) % (n - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < d) {
          const double apq = a[p * ld + q];
          if (fabs(apq) > 1e-300) {
            const double theta = (a[q * ld + q] - a[p * ld + p]) / (2.0 * apq);
            const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
            c = 1.0 / sqrt(t * t + 1.0);
            s = t * c;
          }
        }
        pp[tid] = p; qq[tid] = q; cs[2 * tid] = c; cs[2 * tid + 1] = s;
      }
      __syncthreads();
      for (int idx = tid; idx < half * d; idx += nt) {  // columns p,q of A and V
        const int pr = idx / d, r = idx % d;
        const double s = cs[2 * pr + 1];
        if (s == 0.0) continue;
        const double c = cs[2 * pr];
        const int p = pp[pr], q = qq[pr];
        const double arp = a[r * ld + p], arq = a[r * ld + q];
        a[r * ld + p] = c * arp - s * arq;
        a[r * ld + q] = s * arp + c * arq;
        const double vrp = v[r * ld + p], vrq = v[r * ld + q];
        v[r * ld + p] = c * vrp - s * vrq;
        v[r * ld + q] = s * vrp + c * vrq;
      }
      __syncthreads();
      for (int idx = tid; idx < half * d; idx += nt) {  // rows p,q of A
        const int pr = idx / d, r = idx % d;
        const double s = cs[2 * pr + 1];
        if (s == 0.0) continue;
        const double c = cs[2 * pr];
        const int p = pp[pr], q = qq[pr];
        const double apr = a[p * ld + r], aqr = a[q * ld + r];
        a[p * ld + r] = c * apr - s * aqr;
        a[q * ld + r] = s * apr + c * aqr;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  for (int i = tid; i < d; i += nt) {
    const double li = a[i * ld + i];
    int cnt = 0;
    for (int j = 0; j < d; ++j) {
      const double lj = a[j * ld + j];
      cnt += (lj < li) || (lj == li && j < i);
    }
    rnk[i] = cnt;
    evals[cnt] = static_cast<float>(li);
  }
  __syncthreads();
  for (int idx = tid; idx < d * d; idx += nt) {
    const int r = idx / d, i = idx % d;
    const int c = rnk[i] - (d - k);
    if (c >= 0) U[static_cast<long long>(r) * ldu + c] = static_cast<float>(v[r * ld + i]);
  }
}

// ============================================================================ copy-in
// W[r][c] = A[max(r,c)][min(r,c)] (lower triangle authoritative), zero in the pad columns.
__global__ void copy_sym_kernel(const float* __restrict__ A, long long lda, int d,
                                float* __restrict__ W, long long ldw) {
  __shared__ float tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  if (bi >= bj) {
    for (int yy = ty; yy < 32; yy += 8) {
      const int i = bi * 32 + yy, j = bj * 32 + tx;
      if (i < d && j < ldw) {
        float x = 0.f;
        if (j < d) {
          const int hi = i > j ? i : j, lo = i > j ? j : i;
          x = A[static_cast<long long>(hi) * lda + lo];
        }
        W[static_cast<long long>(i) * ldw + j] = x;
      }
    }
  } else {
    for (int yy = ty; yy < 32; yy += 8) {
      const int i = bj * 32 + yy, j = bi * 32 + tx;  // source tile (bj, bi), below the diagonal
      tile[yy][tx] = (i < d && j < d) ? A[static_cast<long long>(i) * lda + j] : 0.f;
    }
    __syncthreads();
    for (int yy = ty; yy < 32; yy += 8) {
      const int i = bi * 32 + yy, j = bj * 32 + tx;
      if (i < d && j < ldw) W[static_cast<long long>(i) * ldw + j] = tile[tx][yy];
    }
  }
}

// In place: A[i][j] = A[j][i] for j > i over an m x m block (the lower triangle is authoritative).
// Run once when the panels switch from the lower-triangle symv back to full rows.
__global__ void mirror_lower_kernel(float* __restrict__ A, long long lda, int m) {
  __shared__ float tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int yy = ty; yy < 32; yy += 8) {
    const int i = bj * 32 + yy, j = bi * 32 + tx;  // source tile (bj, bi): on or below the diagonal
    tile[yy][tx] = (i < m && j < m) ? A[static_cast<long long>(i) * lda + j] : 0.f;
  }
  __syncthreads();
  for (int yy = ty; yy < 32; yy += 8) {
    const int i = bi * 32 + yy, j = bj * 32 + tx;
    if (i < m && j < m && j > i) A[static_cast<long long>(i) * lda + j] = tile[tx][yy];
  }
}

// ============================================================================ sytrd panel
struct PanelArgs {
  float* A;
  long long ldA;
  int d, j0, ncols, rows_per_cta;
  float* Vp;      // [d][NB] indexed by global row
  float* Wp;      // [d][NB]
  float* colbuf;  // [d] (local row index)
  double* nacc;   // [NB] per-column sum of squares, accumulated with atomics (pre-zeroed)
  double* gacc;   // [NB][GP_STRIDE] per-column W^T v | V^T v | v^T A v; fp64 atomics make the sum
                  // independent of arrival order to ~1e-16, i.e. reproducible after rounding
  float* ptop;    // [NB]
  float* dvec;    // [d]
  float* evec;    // [d]
  float* taus;    // [d]
  float* Tmat;    // [NB][NB] of this panel (pre-zeroed)
  unsigned* bar;  // grid barrier counter (pre-zeroed)
  int sym;        // 1: symv reads only the lower triangle (see sym_symv)
  int sym_tile;   // tile edge of the triangle partition (multiple of 4, <= 256)
  double* pglob;  // [2][ldp] symv result of the current / next column (sym mode; pre-zeroed)
  long long ldp;
  int prof;       // 1: CTA 0 accumulates per-phase cycle counts into g_phase_cycles
};

// Panel kernel, cycles of the middle CTA summed over columns: 0 P1 | 1 barrier A | 6 column gather +
// scalars | 2 v fill + V writes | 7 symv row items | 8 symv segment sums | 3 partial dots + atomics |
// 4 barrier B | 9 P3 gather | 5 P3 w rows   (the resident kernel uses the same slots for its own phases)
__device__ unsigned long long g_phase_cycles[16];

// Grid barrier of the cooperative panel kernel: release-arrive on one counter, acquire-poll.
// (Measured alternatives, both slower: per-CTA release flags written by the last arriver, and a
// two-level arrival over groups of 12 CTAs -- the cost is dependent L2 round trips, not
// contention on the counter.)
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++epoch;
    const unsigned target = epoch * gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned v;
    unsigned polls = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (++polls > (1u << 27)) __trap();  // never hang the box on a lost CTA
    } while (v < target);
  }
  __syncthreads();
}

// Sum over the 32 lanes of 16 per-lane values with 16 shuffles instead of 80: each stage halves
// the values a lane carries and the lanes a value still has to visit. Returns, in every lane, the
// full sum of v[8*b4 + 4*b3 + 2*b2 + b1] (b_k = bit k of the lane index).
__device__ __forceinline__ float reduce16_packed(const float (&v)[16], int lane) {
  float a8[8], a4[4], a2[2];
  const bool u4 = (lane & 16) != 0, u3 = (lane & 8) != 0, u2 = (lane & 4) != 0, u1 = (lane & 2) != 0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    a8[j] = (u4 ? v[j + 8] : v[j]) + __shfl_xor_sync(0xffffffffu, u4 ? v[j] : v[j + 8], 16);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    a4[j] = (u3 ? a8[j + 4] : a8[j]) + __shfl_xor_sync(0xffffffffu, u3 ? a8[j] : a8[j + 4], 8);
#pragma unroll
  for (int j = 0; j < 2; ++j)
    a2[j] = (u2 ? a4[j + 2] : a4[j]) + __shfl_xor_sync(0xffffffffu, u2 ? a4[j] : a4[j + 2], 4);
  float a1 = (u1 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, u1 ? a2[0] : a2[1], 2);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
  return a1;
}

// Symmetric matrix-vector product from the LOWER triangle only (large d: the trailing matrix no
// longer fits the L2 and the one-stage symv is HBM-bound, so reading half of it is the lever).
// The triangle is cut into T x T tiles dealt round-robin to the CTAs (fixed for the whole panel);
// an element A[r][c], c < r, feeds p[r] += A v[c] (row part) and p[c] += A v[r] (column part).
// All (tile, 128-column strip, 16-row chunk) items of a CTA are walked by its 16 warps in one
// flat loop -- every lane keeps 8 independent 16-byte loads in flight, no barrier between tiles --
// with row dots (packed warp reduction) and column sums (registers) gathered per tile in shared
// memory, then flushed once with fp64 atomics into the global p (2T atomics per T^2 elements;
// fp64 keeps the sum independent of arrival order after rounding).
// Returns this thread's partial of v^T A v.
constexpr int SYM_MAX_SLOTS = 12;  // tiles per CTA
__device__ __forceinline__ void sym_tile_of(int t, int T, int m, int L, int& ta, int& tb, int& ca,
                                            int& cb, bool& diag) {
  int bi = static_cast<int>((sqrtf(8.f * static_cast<float>(t) + 1.f) - 1.f) * 0.5f);
  while (bi * (bi + 1) / 2 > t) --bi;
  while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
  const int bj = t - bi * (bi + 1) / 2;
  ta = bi * T;
  tb = min(m, ta + T);
  ca = bj * T;
  cb = min(L, ca + T);
  diag = bi == bj;
}

__device__ __forceinline__ float sym_symv(const float* __restrict__ Abase, long long ldA, int i,
                                          const float* vs, int T, double* pcur, float* trow,
                                          float* tcol, const int* tiles, int nslots, int tid) {
  // tiles[slot * 4 + {0,1,2,3}] = ta, tb, ca, cb of the CTA's slot-th tile (diagonal iff ta == ca),
  // decoded once per panel launch
  const int warp = tid >> 5, lane = tid & 31;
  const int nstrip = (T + 127) >> 7, nchunk = (T + 15) >> 4;
  const int per_tile = nstrip * nchunk;
  float vav = 0.f;
  for (int idx = tid; idx < nslots * 256; idx += PANEL_THREADS) {
    trow[idx] = 0.f;
    tcol[idx] = 0.f;
  }
  __syncthreads();
  for (int item = warp; item < nslots * per_tile; item += PANEL_WARPS) {
    const int slot = item / per_tile, rem = item - slot * per_tile;
    const int ta = tiles[slot * 4], tb = tiles[slot * 4 + 1], ca = tiles[slot * 4 + 2],
              cb = tiles[slot * 4 + 3];
    const bool diag = ta == ca;
    if (tb <= i + 1 || cb <= i + 1) continue;  // no active row / column in this tile
    const int sidx = rem % nstrip, ch = rem / nstrip;
    const int c = ca + (sidx << 7) + 4 * lane;  // first of this lane's four columns
    const int rbase = ta + (ch << 4);
    if (rbase >= tb || ca + (sidx << 7) >= cb) continue;          // edge tile: nothing here
    if (diag && ca + (sidx << 7) > rbase + 15) continue;          // strip entirely above the diagonal
    const bool cvalid = c < cb;
    const float4 x = cvalid ? *reinterpret_cast<const float4*>(vs + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 colacc = make_float4(0.f, 0.f, 0.f, 0.f);
    float rd[16];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 a[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = rbase + 8 * h + u;
        a[u] = (cvalid && r < tb)
                   ? __ldg(reinterpret_cast<const float4*>(Abase + static_cast<long long>(r) * ldA + c))
                   : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = rbase + 8 * h + u;
        const float vr = (r < tb) ? vs[r] : 0.f;
        float4 e = a[u];  // entries with c <= r (diagonal included)
        float4 f = a[u];  // entries with c <  r
        if (diag) {
          e.x = (c <= r) ? e.x : 0.f;
          e.y = (c + 1 <= r) ? e.y : 0.f;
          e.z = (c + 2 <= r) ? e.z : 0.f;
          e.w = (c + 3 <= r) ? e.w : 0.f;
          f.x = (c < r) ? f.x : 0.f;
          f.y = (c + 1 < r) ? f.y : 0.f;
          f.z = (c + 2 < r) ? f.z : 0.f;
          f.w = (c + 3 < r) ? f.w : 0.f;
        }
        const float dot = e.x * x.x + e.y * x.y + e.z * x.z + e.w * x.w;
        rd[8 * h + u] = dot;
        colacc.x += f.x * vr;
        colacc.y += f.y * vr;
        colacc.z += f.z * vr;
        colacc.w += f.w * vr;
        if (diag) {
          const float off = f.x * x.x + f.y * x.y + f.z * x.z + f.w * x.w;
          vav += vr * (off + dot);  // 2 * off + (dot - off)
        } else {
          vav += 2.f * vr * dot;
        }
      }
    }
    const float rsum = reduce16_packed(rd, lane);
    if ((lane & 1) == 0) {
      const int rl = rbase - ta + (((lane >> 4) & 1) << 3) + (((lane >> 3) & 1) << 2) +
                     (((lane >> 2) & 1) << 1) + ((lane >> 1) & 1);
      if (rl < tb - ta && rsum != 0.f) atomicAdd(trow + slot * 256 + rl, rsum);
    }
    if (cvalid) {
      float* tc = tcol + slot * 256 + (c - ca);
      atomicAdd(tc, colacc.x);
      atomicAdd(tc + 1, colacc.y);
      atomicAdd(tc + 2, colacc.z);
      atomicAdd(tc + 3, colacc.w);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < nslots * 512; idx += PANEL_THREADS) {
    const int slot = idx >> 9, within = idx & 511;
    const int ta = tiles[slot * 4], tb = tiles[slot * 4 + 1], ca = tiles[slot * 4 + 2],
              cb = tiles[slot * 4 + 3];
    if (tb <= i + 1 || cb <= i + 1) continue;
    if (within < 256) {
      if (within < tb - ta) {
        const float sr = trow[slot * 256 + within];
        if (sr != 0.f) atomicAdd(pcur + ta + within, static_cast<double>(sr));
      }
    } else if (within - 256 < cb - ca) {
      const float sc = tcol[slot * 256 + within - 256];
      if (sc != 0.f) atomicAdd(pcur + ca + within - 256, static_cast<double>(sc));
    }
  }
  return vav;
}

// One launch = one panel of up to NB Householder columns (LAPACK latrd, lower variant).
// Cooperative: all CTAs are co-resident; row slices of the trailing matrix are owned by CTAs.
template <bool SYM>
__global__ void __launch_bounds__(PANEL_THREADS, 1) sytrd_panel_kernel(const PanelArgs g) {
  extern __shared__ float smf[];
  const int m = g.d - g.j0;
  const int L = static_cast<int>(g.ldA - g.j0);  // padded row length (multiple of 4)
  float* vs = smf;                           // [L] current Householder vector, local index
  float* Vt = vs + L;                        // [NB][NB+1] top block of V (rows < NB)
  float* Wt = Vt + NB * (NB + 1);            // [NB][NB+1] top block of W
  float* pbuf = Wt + NB * (NB + 1);          // [rows_per_cta] symv result of own rows
  float* gWs = pbuf + g.rows_per_cta;        // [NB]  W^T v
  float* gVs = gWs + NB;                     // [NB]  V^T v
  float* red = gVs + NB;                     // [PANEL_WARPS][2*NB]
  float* pts = red + PANEL_WARPS * 2 * NB;   // [NB] symv results of the top rows (all CTAs)
  float* Ts = pts + NB;                      // [NB][NB+1] compact-WY T (CTA 0 only)
  float* pseg = Ts + NB * (NB + 1);          // [rows_per_cta][segments] symv partials
  __shared__ double sred[PANEL_WARPS];
  __shared__ double s_scal[4];
  __shared__ float s_trow[SYM ? SYM_MAX_SLOTS * 256 : 1], s_tcol[SYM ? SYM_MAX_SLOTS * 256 : 1];
  __shared__ int s_tiles[SYM ? SYM_MAX_SLOTS * 4 : 1];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  const int r0 = cta * g.rows_per_cta;
  const int r1 = min(m, r0 + g.rows_per_cta);
  unsigned epoch = 0;
  const float* Abase = g.A + static_cast<long long>(g.j0) * g.ldA + g.j0;
  float* Vp = g.Vp + static_cast<long long>(g.j0) * NB;
  float* Wp = g.Wp + static_cast<long long>(g.j0) * NB;

  int sym_nslots = 0;
  if (SYM) {  // this CTA's tiles of the lower-triangle partition: fixed for the whole panel
    const int T = g.sym_tile;
    const int kk = (m + T - 1) / T;
    const int ntiles = kk * (kk + 1) / 2;
    sym_nslots = (ntiles > cta) ? (ntiles - cta + G - 1) / G : 0;
    if (tid < sym_nslots) {
      int ta, tb, ca, cb;
      bool diag;
      sym_tile_of(cta + tid * G, T, m, L, ta, tb, ca, cb, diag);
      s_tiles[tid * 4] = ta;
      s_tiles[tid * 4 + 1] = tb;
      s_tiles[tid * 4 + 2] = ca;
      s_tiles[tid * 4 + 3] = cb;
    }
  }
  for (int idx = tid; idx < 2 * NB * (NB + 1); idx += PANEL_THREADS) Vt[idx] = 0.f;
  for (int idx = tid; idx < NB * (NB + 1); idx += PANEL_THREADS) Ts[idx] = 0.f;
  __syncthreads();

  const bool prof = g.prof && cta == G / 2 && tid == 0;  // a CTA that keeps its rows to the end
  long long tp = prof ? clock64() : 0;
#define PTD_PHASE(k)                 \
  if (prof) {                        \
    const long long tn = clock64();  \
    pc[k] += tn - tp;                \
    tp = tn;                         \
  }
  long long pc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < g.ncols; ++i) {
    const int j = g.j0 + i;
    // ---------------------------------------------------------------- P1: updated column i
    // 64 rows per pass, 8 threads per row, each thread one contiguous 8-column chunk of the row's
    // V / W entries (two 16-byte loads per array, all independent: one L2 round trip per pass)
    double nrm = 0.0;
    {
      const int sub = tid & 7, c0 = sub * 8;
      for (int rb = max(r0, i); rb < r1; rb += PANEL_THREADS / 8) {
        const int r = rb + (tid >> 3);
        float s = 0.f;
        if (r < r1 && c0 < i) {
          const float4* vp = reinterpret_cast<const float4*>(Vp + static_cast<long long>(r) * NB + c0);
          const float4* wp = reinterpret_cast<const float4*>(Wp + static_cast<long long>(r) * NB + c0);
          const float4 v0 = vp[0], v1 = vp[1], w0 = wp[0], w1 = wp[1];
          const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
          const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (c0 + e < i) s += vv[e] * Wt[i * (NB + 1) + c0 + e] + ww[e] * Vt[i * (NB + 1) + c0 + e];
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (sub == 0 && r < r1) {
          // column i of the trailing matrix: from row i of the (mirrored) upper triangle, or, when
          // only the lower triangle is kept current (SYM panels), from column i itself
          const float a = (SYM ? Abase[static_cast<long long>(r) * g.ldA + i]
                               : Abase[static_cast<long long>(i) * g.ldA + r]) - s;
          g.colbuf[r] = a;
          if (r >= i + 2) nrm += static_cast<double>(a) * a;
        }
      }
    }
    nrm = warp_sum(nrm);
    if (lane == 0) sred[warp] = nrm;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < PANEL_WARPS; ++w) s += sred[w];
      if (s != 0.0) atomicAdd(g.nacc + i, s);
    }
    PTD_PHASE(0)
    grid_barrier(g.bar, epoch);
    PTD_PHASE(1)

    // ---------------------------------------------------------------- P2: reflector + symv
    // The column and its norm were produced by other CTAs: read through L2 (ld.cg). The column
    // loads are issued before the norm is consumed so both latencies overlap.
    float raw[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = tid + u * PANEL_THREADS;
      raw[u] = (c > i + 1 && c < m) ? __ldcg(g.colbuf + c) : 0.f;
    }
    const double xnorm2 = __ldcg(g.nacc + i);
    const float dj = __ldcg(g.colbuf + i);
    if (i + 1 >= m) {  // last diagonal element of the matrix: no reflector (uniform branch)
      if (cta == 0 && tid == 0) g.dvec[j] = dj;
      break;
    }
    const float alpha = __ldcg(g.colbuf + i + 1);
    float tau = 0.f, beta = alpha, scale = 0.f;
    if (xnorm2 > 0.0) {
      const double bt = -copysign(sqrt(static_cast<double>(alpha) * alpha + xnorm2),
                                  static_cast<double>(alpha));
      beta = static_cast<float>(bt);
      tau = static_cast<float>((bt - alpha) / bt);
      scale = static_cast<float>(1.0 / (static_cast<double>(alpha) - bt));
    }
    if (cta == 0 && tid == 0) {
      g.dvec[j] = dj;
      g.evec[j] = beta;
      g.taus[j] = tau;
    }
    if (prof && tau == 123.456f) pc[9] += 1;  // (profile only) the stamp below waits for the gathered scalars
    PTD_PHASE(6)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = tid + u * PANEL_THREADS;
      if (c < L) vs[c] = (c == i + 1) ? 1.f : raw[u] * scale;
    }
    for (int c = tid + 8 * PANEL_THREADS; c < L; c += PANEL_THREADS) {
      float x = 0.f;
      if (c == i + 1) x = 1.f;
      else if (c > i + 1 && c < m) x = __ldcg(g.colbuf + c) * scale;
      vs[c] = x;
    }
    __syncthreads();
    for (int r = r0 + tid; r < r1; r += PANEL_THREADS) Vp[r * NB + i] = vs[r];
    for (int r = tid; r < NB; r += PANEL_THREADS) Vt[r * (NB + 1) + i] = (r < L) ? vs[r] : 0.f;
    PTD_PHASE(2)

    if (SYM) {
      // The other parity buffer of p was last read in P3 of the previous column (before barrier
      // A) and is next accumulated after the next barrier A: zero it here, tau or not.
      double* pnext = g.pglob + static_cast<long long>((i + 1) & 1) * g.ldp;
      for (int c = cta * PANEL_THREADS + tid; c < L; c += G * PANEL_THREADS) pnext[c] = 0.0;
    }
    if (tau != 0.f) {
      const int rbeg = max(r0, i + 1);
      const int R = max(0, r1 - rbeg);
      float vav_sym = 0.f;
      if (SYM) {
        double* pcur = g.pglob + static_cast<long long>(i & 1) * g.ldp;
        vav_sym = sym_symv(Abase, g.ldA, i, vs, g.sym_tile, pcur, s_trow, s_tcol, s_tiles, sym_nslots, tid);
      } else {
        // symv over the CTA's row block, split into (row, 1024-float segment) items so that every
        // lane keeps 8 independent 16-byte loads in flight even when the block has few rows
        const int n4 = L >> 2;
        const float4* v4 = reinterpret_cast<const float4*>(vs);
        const int seg0 = (i + 1) >> 10;  // segments entirely left of column i+1 multiply zeros
        const int nseg = ((n4 + 255) >> 8) - seg0;
        for (int item = warp; item < R * nseg; item += PANEL_WARPS) {
          const int rr = item / nseg, sg = item - rr * nseg + seg0;
          const float4* arow =
              reinterpret_cast<const float4*>(Abase + static_cast<long long>(rbeg + rr) * g.ldA);
          const int cbase = (sg << 8) + lane;
          float4 a[8];
  #pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int c4 = cbase + 32 * u;
            a[u] = (c4 < n4) ? __ldg(arow + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          float s0 = 0.f, s1 = 0.f;
  #pragma unroll
          for (int u = 0; u < 8; u += 2) {
            const int c4 = cbase + 32 * u;
            if (c4 < n4) {
              const float4 x = v4[c4];
              s0 += a[u].x * x.x + a[u].y * x.y + a[u].z * x.z + a[u].w * x.w;
            }
            if (c4 + 32 < n4) {
              const float4 x = v4[c4 + 32];
              s1 += a[u + 1].x * x.x + a[u + 1].y * x.y + a[u + 1].z * x.z + a[u + 1].w * x.w;
            }
          }
          const float s = warp_sum(s0 + s1);
          if (lane == 0) pseg[rr * nseg + (sg - seg0)] = s;
        }
        PTD_PHASE(7)
        __syncthreads();
        for (int rr = tid; rr < R; rr += PANEL_THREADS) {
          float s = 0.f;
          for (int sg = 0; sg < nseg; ++sg) s += pseg[rr * nseg + sg];
          const int r = rbeg + rr;
          pbuf[r - r0] = s;
          if (r < NB) g.ptop[r] = s;
        }
        __syncthreads();
        PTD_PHASE(8)
      }
      float aw0 = 0.f, aw1 = 0.f, av0 = 0.f, av1 = 0.f;
      double vp = 0.0;
      for (int r = max(r0, i + 1) + warp; r < r1; r += PANEL_WARPS) {
        const float vr = vs[r];
        if (lane < i) {
          aw0 += Wp[r * NB + lane] * vr;
          av0 += Vp[r * NB + lane] * vr;
        }
        if (lane + 32 < i) {
          aw1 += Wp[r * NB + lane + 32] * vr;
          av1 += Vp[r * NB + lane + 32] * vr;
        }
        if (lane == 0 && !SYM) vp += static_cast<double>(vr) * pbuf[r - r0];
      }
      red[warp * 2 * NB + lane] = aw0;
      red[warp * 2 * NB + 32 + lane] = aw1;
      red[warp * 2 * NB + NB + lane] = av0;
      red[warp * 2 * NB + NB + 32 + lane] = av1;
      if (SYM) vp = static_cast<double>(warp_sum(vav_sym));
      if (lane == 0) sred[warp] = vp;
      __syncthreads();
      if (tid < 2 * NB) {
        if ((tid & (NB - 1)) < i && R > 0) {
          float s = 0.f;
          for (int w = 0; w < PANEL_WARPS; ++w) s += red[w * 2 * NB + tid];
          atomicAdd(g.gacc + i * GP_STRIDE + tid, static_cast<double>(s));
        }
      } else if (tid == 2 * NB && (R > 0 || SYM)) {
        double s = 0.0;
        for (int w = 0; w < PANEL_WARPS; ++w) s += sred[w];
        atomicAdd(g.gacc + i * GP_STRIDE + 2 * NB, s);
      }
    }
    PTD_PHASE(3)
    grid_barrier(g.bar, epoch);
    PTD_PHASE(4)

    // ---------------------------------------------------------------- P3: w column
    if (tau != 0.f) {
      if (tid <= 2 * NB) {
        const double s = __ldcg(g.gacc + i * GP_STRIDE + tid);
        if (tid < NB) gWs[tid] = static_cast<float>(s);
        else if (tid < 2 * NB) gVs[tid - NB] = static_cast<float>(s);
        else s_scal[1] = s;
      } else if (tid >= 256 && tid < 256 + NB) {  // one parallel fetch of the top rows' symv results
        const int r = tid - 256;
        if (SYM)
          pts[r] = (r > i && r < m)
                       ? static_cast<float>(__ldcg(g.pglob + static_cast<long long>(i & 1) * g.ldp + r))
                       : 0.f;
        else
          pts[r] = (r > i && r < m) ? __ldcg(g.ptop + r) : 0.f;
      }
      if (SYM) {  // own rows of A v, summed over all CTAs' tiles
        const double* pcur = g.pglob + static_cast<long long>(i & 1) * g.ldp;
        for (int r = max(r0, i + 1) + tid; r < r1; r += PANEL_THREADS)
          pbuf[r - r0] = static_cast<float>(__ldcg(pcur + r));
      }
      __syncthreads();
      PTD_PHASE(9)
      if (warp == 0) {
        double s = 0.0;
        for (int c = lane; c < i; c += 32) s += static_cast<double>(gVs[c]) * gWs[c];
        s = warp_sum(s);
        if (lane == 0) {
          const double dot = static_cast<double>(tau) * (s_scal[1] - 2.0 * s);
          s_scal[2] = -0.5 * static_cast<double>(tau) * dot;
        }
      }
      __syncthreads();
      const float alpha2 = static_cast<float>(s_scal[2]);
      {  // own rows: 64 per pass, 8 threads per row (see P1)
        const int sub = tid & 7, c0 = sub * 8;
        for (int rb = max(r0, i + 1); rb < r1; rb += PANEL_THREADS / 8) {
          const int r = rb + (tid >> 3);
          float s = 0.f;
          if (r < r1 && c0 < i) {
            const float4* vp = reinterpret_cast<const float4*>(Vp + static_cast<long long>(r) * NB + c0);
            const float4* wp = reinterpret_cast<const float4*>(Wp + static_cast<long long>(r) * NB + c0);
            const float4 v0 = vp[0], v1 = vp[1], w0 = wp[0], w1 = wp[1];
            const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (c0 + e < i) s += vv[e] * gWs[c0 + e] + ww[e] * gVs[c0 + e];
          }
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          if (sub == 0 && r < r1) Wp[r * NB + i] = tau * (pbuf[r - r0] - s) + alpha2 * vs[r];
        }
      }
      for (int r = r0 + tid; r < min(r1, i + 1); r += PANEL_THREADS) Wp[r * NB + i] = 0.f;
      {  // top block, redundantly in every CTA: all NB rows in one pass, 8 threads per row
        static_assert(PANEL_THREADS == 8 * NB, "one 8-thread group per top-block row");
        const int r = tid >> 3, sub = tid & 7;
        const bool live = r > i && r < m;
        float s = 0.f;
        if (live)
          for (int c = sub; c < i; c += 8)
            s += Vt[r * (NB + 1) + c] * gWs[c] + Wt[r * (NB + 1) + c] * gVs[c];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (sub == 0) Wt[r * (NB + 1) + i] = live ? tau * (pts[r] - s) + alpha2 * vs[r] : 0.f;
      }
      if (cta == 0 && tid < NB) {  // compact-WY T column: T[0:i,i] = -tau T[0:i,0:i] (V^T v)
        if (tid < i) {
          float s = 0.f;
          for (int c2 = tid; c2 < i; ++c2) s += Ts[tid * (NB + 1) + c2] * gVs[c2];
          Ts[tid * (NB + 1) + i] = -tau * s;
        } else if (tid == i) {
          Ts[i * (NB + 1) + i] = tau;
        }
      }
    } else {
      for (int r = r0 + tid; r < r1; r += PANEL_THREADS) Wp[r * NB + i] = 0.f;
      for (int r = tid; r < NB; r += PANEL_THREADS) Wt[r * (NB + 1) + i] = 0.f;
    }
    __syncthreads();
    PTD_PHASE(5)
  }
  if (prof)
    for (int k2 = 0; k2 < 10; ++k2) atomicAdd(&g_phase_cycles[k2], static_cast<unsigned long long>(pc[k2]));
#undef PTD_PHASE
  if (cta == 0) {
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += PANEL_THREADS)
      g.Tmat[idx] = Ts[(idx / NB) * (NB + 1) + (idx % NB)];
  }
}

// bf16x3 split of the finished panel: reflectors into the global store Vs (used by the
// back-transformation) and the [V|W], [W|V] operands of the rank-2nb trailing update.
__global__ void panel_split_kernel(const float* __restrict__ Vp, const float* __restrict__ Wp,
                                   int j0, int m, int ncols, __nv_bfloat16* __restrict__ Vs,
                                   long long ldv, long long vs_seg, __nv_bfloat16* __restrict__ VW,
                                   __nv_bfloat16* __restrict__ WV, long long vw_seg) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int groups = NB / 8;
  if (idx >= m * groups) return;
  const int r = idx / groups, c0 = (idx % groups) * 8;
  const long long gr = j0 + r;
  __align__(16) __nv_bfloat16 vh[3][8], wh[3][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c0 + e;
    const float v = (c < ncols) ? Vp[gr * NB + c] : 0.f;
    const float w = (c < ncols) ? Wp[gr * NB + c] : 0.f;
    split3(v, vh[0][e], vh[1][e], vh[2][e]);
    split3(w, wh[0][e], wh[1][e], wh[2][e]);
  }
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    if (c0 < ncols)
      *reinterpret_cast<uint4*>(Vs + s * vs_seg + gr * ldv + j0 + c0) =
          *reinterpret_cast<const uint4*>(vh[s]);
    __nv_bfloat16* vw = VW + s * vw_seg + gr * (2 * NB);
    __nv_bfloat16* wv = WV + s * vw_seg + gr * (2 * NB);
    *reinterpret_cast<uint4*>(vw + c0) = *reinterpret_cast<const uint4*>(vh[s]);
    *reinterpret_cast<uint4*>(vw + NB + c0) = *reinterpret_cast<const uint4*>(wh[s]);
    *reinterpret_cast<uint4*>(wv + c0) = *reinterpret_cast<const uint4*>(wh[s]);
    *reinterpret_cast<uint4*>(wv + NB + c0) = *reinterpret_cast<const uint4*>(vh[s]);
  }
}

// ============================================================================ resident sytd2
// Tridiagonalisation of a trailing block that FITS THE CHIP'S SHARED MEMORY (m0 <= ~2560 on 148
// SMs): rows are dealt cyclically to the CTAs of one cooperative launch and stay in shared memory
// for the whole reduction, so a Householder column costs no trip over the matrix through L2/HBM.
// The unblocked recurrence (LAPACK sytd2) is restructured so that ONE grid-wide exchange per
// column suffices (the blocked panel kernel needs two):
//   * the rank-2 update of column i-1 (v_{i-1}, w_{i-1}) is applied on the fly while the symv of
//     column i streams over the rows (one read-modify-write pass over shared memory),
//   * every CTA then publishes its rows of p = A v_i, and the owner of row i+1 publishes that row;
//     one release-arrive on the column's own counter, ONE thread per CTA polls it (measured on a
//     B200, tools/micro/exchange_bench.cu: 1.9-2.0 us per round whatever m; self-validating tagged
//     packets polled by every thread cost 5-6.6 us -- 75k pollers swamp the L2),
//   * every CTA reads p and the row back through L2 and REDUNDANTLY computes w_i, column i+1 and
//     the next reflector with identical arithmetic, so all CTAs agree bit for bit without a
//     second exchange.
// The exchange buffers are double-buffered by column parity: a CTA can publish column i+1 only
// after passing the barrier of column i, i.e. after every CTA finished reading column i-1.
constexpr int RES_THREADS = 512;
constexpr int RES_WARPS = RES_THREADS / 32;
constexpr int RES_MAX_NLOC = 32;   // local rows per CTA
constexpr int RES_SLOTS = 2;       // 4-column chunks per thread: covers row lengths up to 4096
constexpr int RES_MAX_L = 4 * RES_THREADS * RES_SLOTS;

struct ResArgs {
  const float* A;       // working copy (both triangles valid); the block starts at (j0, j0)
  long long ldA;
  int j0, m0, L, nloc;  // L = m0 rounded up to 4; nloc = ceil(m0 / grid)
  float* VT;            // [m0][ldvt]: row i = reflector of column j0 + i (local row index), pre-zeroed
  long long ldvt;
  float* dvec;          // [d]
  float* evec;          // [d]
  float* taus;          // [d]
  float* xP;            // [2][L] exchanged symv results
  float* xR;            // [2][L] exchanged row
  unsigned* ctr;        // [m0] arrival counter of every column (pre-zeroed)
  int prof;
};

// Sum over the CTA; every thread gets the same value (identical order in every CTA).
__device__ __forceinline__ double res_block_sum(float v, float* buf, int warp, int lane) {
  v = warp_sum(v);
  if (lane == 0) buf[warp] = v;
  __syncthreads();
  // fixed pairwise tree in fp32: a chain of 16 dependent fp64 adds cost ~400 cycles per call here
  float t[RES_WARPS];
#pragma unroll
  for (int w = 0; w < RES_WARPS; ++w) t[w] = buf[w];
#pragma unroll
  for (int span = RES_WARPS / 2; span > 0; span >>= 1)
#pragma unroll
    for (int w = 0; w < span; ++w) t[w] += t[w + span];
  return static_cast<double>(t[0]);
}

// Sums over the 32 lanes of 8 per-lane values with 9 shuffles instead of 40 (see reduce16_packed):
// every lane returns the full sum of v[4*b4 + 2*b3 + b2] (b_k = bit k of the lane index).
__device__ __forceinline__ float reduce8_packed(const float (&v)[8], int lane) {
  float a4[4], a2[2];
  const bool u4 = (lane & 16) != 0, u3 = (lane & 8) != 0, u2 = (lane & 4) != 0;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    a4[j] = (u4 ? v[j + 4] : v[j]) + __shfl_xor_sync(0xffffffffu, u4 ? v[j] : v[j + 4], 16);
#pragma unroll
  for (int j = 0; j < 2; ++j)
    a2[j] = (u3 ? a4[j + 2] : a4[j]) + __shfl_xor_sync(0xffffffffu, u3 ? a4[j] : a4[j + 2], 8);
  float a1 = (u2 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, u2 ? a2[0] : a2[1], 4);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
  a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
  return a1;
}

// Reflector of column n from the full column a (this thread's chunks): d_n = a[n], alpha = a[n+1],
// x = a[n+2:]. Writes v (1 at n+1, zeros up to n) into vnext, returns tau. Uniform over the CTA.
struct ResReflector { float tau, beta, dn; };

__global__ void __launch_bounds__(RES_THREADS, 1) sytd2_resident_kernel(const ResArgs g) {
  extern __shared__ float smf[];
  const int L = g.L, m0 = g.m0;
  const int G = gridDim.x, cta = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* As = smf;                                        // [nloc][L]
  float* vb = As + static_cast<size_t>(g.nloc) * L;       // [3][L] rotating v_{i-1}, v_i, v_{i+1}
  float* wprev = vb + 3 * L;                              // [L]
  float2* rowvw = reinterpret_cast<float2*>(wprev + L);   // [RES_MAX_NLOC] (v_{i-1}[r], w_{i-1}[r]) of own rows
  float* red = reinterpret_cast<float*>(rowvw + RES_MAX_NLOC);  // [RES_WARPS][RES_MAX_NLOC]
  __shared__ float sred[2][RES_WARPS];
  __shared__ float s_b[4];
  const int nrows = (m0 > cta) ? (m0 - cta + G - 1) / G : 0;  // rows cta, cta + G, ...
  const int nq = L >> 2;

  const bool prof = g.prof && cta == 0 && tid == 0;
  long long tp = prof ? clock64() : 0;
#define RES_PHASE(k)                                             \
  if (prof) {                                                    \
    const long long tn = clock64();                              \
    atomicAdd(&g_phase_cycles[k], (unsigned long long)(tn - tp)); \
    tp = tn;                                                     \
  }

  // ---- own rows into shared memory
  for (int lr = 0; lr < nrows; ++lr) {
    const float* src = g.A + static_cast<long long>(g.j0 + cta + lr * G) * g.ldA + g.j0;
    for (int q = tid; q < nq; q += RES_THREADS) {
      const int c = q << 2;
      float4 x;
      if (c + 3 < m0) {
        x = __ldg(reinterpret_cast<const float4*>(src + c));
      } else {
        x.x = (c < m0) ? src[c] : 0.f;
        x.y = (c + 1 < m0) ? src[c + 1] : 0.f;
        x.z = (c + 2 < m0) ? src[c + 2] : 0.f;
        x.w = 0.f;
      }
      *reinterpret_cast<float4*>(As + static_cast<size_t>(lr) * L + c) = x;
    }
  }
  for (int c = tid; c < 3 * L + L; c += RES_THREADS) vb[c] = 0.f;  // v buffers and wprev
  __syncthreads();

  // Reflector of column n from this thread's elements a[s][e] of the full column (chunk base
  // columns cb[s]); publishes d / e / tau / v through CTA (n mod G). Returns tau.
  auto make_reflector = [&](const float (&a)[RES_SLOTS][4], const int (&cb)[RES_SLOTS],
                            const bool (&live)[RES_SLOTS], int n, float* vnext, int redbuf) -> float {
    float part = 0.f;
#pragma unroll
    for (int s = 0; s < RES_SLOTS; ++s)
      if (live[s]) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = cb[s] + e;
          if (c == n) s_b[1] = a[s][e];
          if (c == n + 1) s_b[2] = a[s][e];
          if (c >= n + 2 && c < m0) part += a[s][e] * a[s][e];
        }
      }
    const float xnorm2 = static_cast<float>(res_block_sum(part, sred[redbuf], warp, lane));
    const float dn = s_b[1];
    float tau = 0.f, beta = 0.f, scale = 0.f;
    if (n + 1 < m0) {
      // working-precision reflector (LAPACK slarfg); 512 threads doing this in fp64 would cost
      // more than the whole symv pass
      const float alpha = s_b[2];
      beta = alpha;
      if (xnorm2 > 0.f) {
        beta = -copysignf(sqrtf(alpha * alpha + xnorm2), alpha);
        tau = (beta - alpha) / beta;
        scale = 1.f / (alpha - beta);
      }
    }
    const bool writer = cta == (n % G);
#pragma unroll
    for (int s = 0; s < RES_SLOTS; ++s)
      if (live[s]) {
        float4 v;
        float* vv = reinterpret_cast<float*>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = cb[s] + e;
          vv[e] = (c >= n + 2 && c < m0) ? a[s][e] * scale : ((c == n + 1 && c < m0) ? 1.f : 0.f);
        }
        *reinterpret_cast<float4*>(vnext + cb[s]) = v;
        if (writer && n + 1 < m0)
          *reinterpret_cast<float4*>(g.VT + static_cast<long long>(n) * g.ldvt + cb[s]) = v;
      }
    if (writer && tid == 0) {
      g.dvec[g.j0 + n] = dn;
      if (n + 1 < m0) {
        g.evec[g.j0 + n] = beta;
        g.taus[g.j0 + n] = tau;
      }
    }
    return tau;
  };

  int cur = 0;  // vb[cur] = v_i, vb[(cur+2)%3] = v_{i-1}, vb[(cur+1)%3] receives v_{i+1}
  float tau;
  {  // ---- reflector of column 0 straight from the global copy (row j0 == column j0)
    float a[RES_SLOTS][4];
    int cb[RES_SLOTS];
    bool live[RES_SLOTS];
    const float* src = g.A + static_cast<long long>(g.j0) * g.ldA + g.j0;
#pragma unroll
    for (int s = 0; s < RES_SLOTS; ++s) {
      const int q = tid + s * RES_THREADS;
      live[s] = q < nq;
      cb[s] = q << 2;
#pragma unroll
      for (int e = 0; e < 4; ++e) a[s][e] = (live[s] && cb[s] + e < m0) ? src[cb[s] + e] : 0.f;
    }
    tau = make_reflector(a, cb, live, 0, vb + cur * L, 0);
  }
  __syncthreads();
  RES_PHASE(0)

  for (int i = 0; i + 1 < m0; ++i) {
    const int q0 = (i + 1) >> 2;
    float* vcur = vb + cur * L;
    const float* vprev = vb + ((cur + 2) % 3) * L;
    float* vnext = vb + ((cur + 1) % 3) * L;
    const int lr0 = (i + 1 > cta) ? (i + 1 - cta + G - 1) / G : 0;  // first local row >= i + 1
    float* P = g.xP + static_cast<size_t>(i & 1) * L;
    float* R = g.xR + static_cast<size_t>(i & 1) * L;

    if (tid < nrows) {
      const int r = cta + tid * G;
      rowvw[tid] = make_float2(vprev[r], wprev[r]);
    }
    float4 v4[RES_SLOTS], vp4[RES_SLOTS], wp4[RES_SLOTS];
    int cb[RES_SLOTS];
    bool live[RES_SLOTS];
#pragma unroll
    for (int s = 0; s < RES_SLOTS; ++s) {
      const int q = q0 + tid + s * RES_THREADS;
      live[s] = q < nq;
      cb[s] = q << 2;
      if (live[s]) {
        v4[s] = *reinterpret_cast<const float4*>(vcur + cb[s]);
        vp4[s] = *reinterpret_cast<const float4*>(vprev + cb[s]);
        wp4[s] = *reinterpret_cast<const float4*>(wprev + cb[s]);
      } else {
        v4[s] = vp4[s] = wp4[s] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();
    RES_PHASE(1)

    // ---- fused pass: pending rank-2 update of column i-1, then the symv row dots against v_i
    const bool own_next = lr0 < nrows && cta + lr0 * G == i + 1;  // this CTA holds row i+1
    for (int lr8 = lr0; lr8 < nrows; lr8 += 8) {
      float part[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int lr = lr8 + u;
        part[u] = 0.f;
        if (lr < nrows) {
          const float2 rw = rowvw[lr];
          float* arow = As + static_cast<size_t>(lr) * L;
#pragma unroll
          for (int s = 0; s < RES_SLOTS; ++s)
            if (live[s]) {
              float4 a4 = *reinterpret_cast<float4*>(arow + cb[s]);
              a4.x -= rw.x * wp4[s].x + rw.y * vp4[s].x;
              a4.y -= rw.x * wp4[s].y + rw.y * vp4[s].y;
              a4.z -= rw.x * wp4[s].z + rw.y * vp4[s].z;
              a4.w -= rw.x * wp4[s].w + rw.y * vp4[s].w;
              *reinterpret_cast<float4*>(arow + cb[s]) = a4;
              part[u] += a4.x * v4[s].x + a4.y * v4[s].y + a4.z * v4[s].z + a4.w * v4[s].w;
              if (own_next && lr == lr0)  // row i+1 (updated through column i-1) for everybody
                __stcg(reinterpret_cast<float4*>(R + cb[s]), a4);
            }
        }
      }
      const float rsum = reduce8_packed(part, lane);
      if ((lane & 3) == 0) {
        const int lr = lr8 + (((lane >> 4) & 1) << 2) + (((lane >> 3) & 1) << 1) + ((lane >> 2) & 1);
        if (lr < nrows) red[warp * RES_MAX_NLOC + lr] = rsum;
      }
    }
    RES_PHASE(2)
    __syncthreads();
    if (warp == 0) {  // nrows <= 32: lane = local row
      if (lane >= lr0 && lane < nrows) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < RES_WARPS; ++w) s += red[w * RES_MAX_NLOC + lane];
        __stcg(P + cta + lane * G, s);
      }
      __syncwarp();
      if (lane == 0) {  // arrive (release: the row stores above were ordered by the barrier) and wait
        unsigned* c = g.ctr + i;
        RES_PHASE(3)
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(c) : "memory");
        RES_PHASE(4)
        unsigned v, polls = 0;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
          if (++polls > (1u << 26)) __trap();  // never hang the box on a lost CTA
        } while (v < static_cast<unsigned>(G));
        RES_PHASE(5)
      }
    }
    __syncthreads();

    // ---- gather: p (all live rows) and row i+1, for this thread's columns (through L2)
    float pr[RES_SLOTS][4], ar[RES_SLOTS][4];
#pragma unroll
    for (int s = 0; s < RES_SLOTS; ++s) {
      float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f), r4 = p4;
      if (live[s]) {
        p4 = __ldcg(reinterpret_cast<const float4*>(P + cb[s]));
        r4 = __ldcg(reinterpret_cast<const float4*>(R + cb[s]));
      }
      const float* pp = reinterpret_cast<const float*>(&p4);
      const float* rr = reinterpret_cast<const float*>(&r4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = cb[s] + e;
        const bool in = live[s] && c >= i + 1 && c < m0;  // the rest is stale / padding
        pr[s][e] = in ? pp[e] : 0.f;
        ar[s][e] = in ? rr[e] : 0.f;
      }
    }

    // ---- w_i = tau p - (tau^2/2)(p^T v) v ; column i+1 = row i+1 - w_i - w_i[i+1] v_i
    float part = 0.f;
#pragma unroll
    for (int s = 0; s < RES_SLOTS; ++s) {
      const float* vv = reinterpret_cast<const float*>(&v4[s]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        part += pr[s][e] * vv[e];
        if (live[s] && cb[s] + e == i + 1) s_b[0] = pr[s][e];
      }
    }
    RES_PHASE(6)
    const float pv = static_cast<float>(res_block_sum(part, sred[0], warp, lane));
    RES_PHASE(7)
    const float alpha = -0.5f * tau * tau * pv;
    const float w1 = tau * s_b[0] + alpha;  // v_i[i+1] = 1
    float a[RES_SLOTS][4];
#pragma unroll
    for (int s = 0; s < RES_SLOTS; ++s) {
      const float* vv = reinterpret_cast<const float*>(&v4[s]);
      float4 w4;
      float* ww = reinterpret_cast<float*>(&w4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = cb[s] + e;
        const bool in = live[s] && c >= i + 1 && c < m0;
        ww[e] = in ? tau * pr[s][e] + alpha * vv[e] : 0.f;
        a[s][e] = in ? ar[s][e] - ww[e] - w1 * vv[e] : 0.f;
      }
      if (live[s]) *reinterpret_cast<float4*>(wprev + cb[s]) = w4;
    }
    tau = make_reflector(a, cb, live, i + 1, vnext, 1);
    cur = (cur + 1) % 3;
    RES_PHASE(8)
    __syncthreads();
    RES_PHASE(9)
  }
#undef RES_PHASE
}

// ---------------------------------------------------------------------------- d <= 128: one CTA
// The whole matrix fits one SM's shared memory, so the reduction needs no grid-wide exchange at
// all: unblocked sytd2 (LAPACK, lower) with three block barriers per column. One SM is bound by
// instruction issue and shared-memory bandwidth (the scalar version: 4.9 k cycles per column at
// d = 96, 16 warps x ~300 instructions through 4 schedulers), so the symv dots and the rank-2 update
// work on 16-byte vectors: a warp instruction covers 4 rows x 32 columns (8 lanes per row; every
// group of 8 lanes reads 128 contiguous bytes: conflict-free for any row stride). Same outputs as
// the resident kernel (d / e / tau, reflectors as rows of VT). fp32 working precision like the
// other two kernels.
constexpr int SMALL_MAX = 128;
constexpr int SMALL_THREADS = 512;
constexpr int SMALL_LD = SMALL_MAX + 4;
__device__ __forceinline__ int small_ld(int n) { return ((n + 3) & ~3) + 4; }

__global__ void __launch_bounds__(SMALL_THREADS, 1)
sytd2_small_kernel(const float* __restrict__ A, long long ldA, int d, int L, float* __restrict__ VT,
                   long long ldvt, float* __restrict__ dvec, float* __restrict__ evec,
                   float* __restrict__ taus, int prof_on) {
  extern __shared__ __align__(16) float sma[];
  const int LD = small_ld(d);      // multiple of 4: rows are 16-byte aligned; columns >= d hold zeros
  float* As = sma;                 // [d][LD]
  float* v = As + d * LD;          // [SMALL_LD]  zero outside (i, d)
  float* pvec = v + SMALL_LD;      // [SMALL_LD]
  __shared__ float s_tau;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rg = lane >> 3, ch = lane & 7;  // row within the warp's group of 4, 16-byte chunk of 32 columns
  constexpr int NW = SMALL_THREADS / 32;
  for (int idx = tid; idx < d * LD; idx += SMALL_THREADS) {
    const int r = idx / LD, c = idx - r * LD;
    As[idx] = c < d ? A[static_cast<long long>(r) * ldA + c] : 0.f;
  }
  for (int idx = tid; idx < 2 * SMALL_LD; idx += SMALL_THREADS) v[idx] = 0.f;
  __syncthreads();
  const bool prof = prof_on && tid == 32;  // a thread outside warp 0: sees warp 0's phase as barrier wait
  long long tp = prof ? clock64() : 0, pc[6] = {0, 0, 0, 0, 0, 0};
#define SM_PHASE(kk)                \
  if (prof) {                       \
    const long long tn = clock64(); \
    pc[kk] += tn - tp;              \
    tp = tn;                        \
  }
  for (int i = 0; i + 1 < d; ++i) {
    // ---- reflector of column i (warp 0; column i below the diagonal == row i right of it: the
    // trailing square is kept fully symmetric), v into shared memory and into row i of VT
    if (warp == 0) {
      const float* rowi = As + i * LD;
      float part = 0.f;
      for (int r = i + 2 + lane; r < d; r += 32) {
        const float a = rowi[r];
        part += a * a;
      }
      const float xnorm2 = warp_sum(part);
      const float alpha = rowi[i + 1];
      float beta = alpha, tau = 0.f, scale = 0.f;
      if (xnorm2 > 0.f) {
        beta = -copysignf(sqrtf(alpha * alpha + xnorm2), alpha);
        tau = (beta - alpha) / beta;
        scale = 1.f / (alpha - beta);
      }
      for (int c = lane; c < L; c += 32) {
        const float x = (c >= i + 2 && c < d) ? rowi[c] * scale : (c == i + 1 ? 1.f : 0.f);
        v[c] = x;
        VT[static_cast<long long>(i) * ldvt + c] = x;
      }
      if (lane == 0) {
        dvec[i] = rowi[i];
        evec[i] = beta;
        taus[i] = tau;
        s_tau = tau;
      }
    }
    __syncthreads();
    SM_PHASE(0)
    const float tau = s_tau;
    const int cb = (i + 1) & ~3;  // first 16-byte chunk with a live column (v is zero left of i + 1)
    if (tau != 0.f) {
      // ---- p = A22 v
      for (int r0 = i + 1 + warp * 4; r0 < d; r0 += NW * 4) {
        const int r = r0 + rg;
        float part = 0.f;
        if (r < d) {
          const float* arow = As + r * LD;
          for (int c = cb + ch * 4; c < d; c += 32) {
            const float4 a4 = *reinterpret_cast<const float4*>(arow + c);
            const float4 v4 = *reinterpret_cast<const float4*>(v + c);
            part += a4.x * v4.x + a4.y * v4.y + a4.z * v4.z + a4.w * v4.w;
          }
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        part += __shfl_xor_sync(0xffffffffu, part, 4);
        if (ch == 0 && r < d) pvec[r] = part;
      }
      SM_PHASE(1)
      __syncthreads();
      SM_PHASE(2)
      // ---- w = tau p - (tau^2 / 2)(p^T v) v, redundantly per warp; A22 -= v w^T + w v^T
      float pv = 0.f;
      for (int c = i + 1 + lane; c < d; c += 32) pv += pvec[c] * v[c];
      pv = warp_sum(pv);
      const float alpha2 = -0.5f * tau * tau * pv;
      for (int r0 = i + 1 + warp * 4; r0 < d; r0 += NW * 4) {
        const int r = r0 + rg;
        if (r < d) {
          const float vr = v[r], wr = tau * pvec[r] + alpha2 * vr;
          float* arow = As + r * LD;
          for (int c = cb + ch * 4; c < d; c += 32) {
            float4 a4 = *reinterpret_cast<float4*>(arow + c);
            const float4 v4 = *reinterpret_cast<const float4*>(v + c);
            const float4 p4 = *reinterpret_cast<const float4*>(pvec + c);
            // columns left of i + 1 (same chunk) and right of d - 1 stay as they are: v = 0 there,
            // and w is masked (pvec holds stale / no values outside the live range)
            const float w0 = (c >= i + 1 && c < d) ? tau * p4.x + alpha2 * v4.x : 0.f;
            const float w1 = (c + 1 >= i + 1 && c + 1 < d) ? tau * p4.y + alpha2 * v4.y : 0.f;
            const float w2 = (c + 2 >= i + 1 && c + 2 < d) ? tau * p4.z + alpha2 * v4.z : 0.f;
            const float w3 = (c + 3 >= i + 1 && c + 3 < d) ? tau * p4.w + alpha2 * v4.w : 0.f;
            a4.x -= vr * w0 + wr * v4.x;
            a4.y -= vr * w1 + wr * v4.y;
            a4.z -= vr * w2 + wr * v4.z;
            a4.w -= vr * w3 + wr * v4.w;
            *reinterpret_cast<float4*>(arow + c) = a4;
          }
        }
      }
    }
    SM_PHASE(3)
    __syncthreads();
    SM_PHASE(4)
  }
  if (prof)
    for (int k2 = 0; k2 < 5; ++k2) atomicAdd(&g_phase_cycles[k2], static_cast<unsigned long long>(pc[k2]));
#undef SM_PHASE
  if (tid == 0) dvec[d - 1] = As[(d - 1) * LD + (d - 1)];
}

// Back-transformation for d <= 128 in one CTA: U <- H_0 H_1 ... H_{d-2} Z with the reflectors (rows
// of VT) and Z = U both resident in shared memory, one reflector at a time (w = v^T Z, Z -= tau v w^T;
// 2 d^2 k FLOP in all: 2.4 MFLOP at d = k = 96). Replaces, at these sizes, the T factors, the
// bf16x3 reflector store and four launches per panel of the compact-WY path, which cost ~100 us
// of launch latency and tiny GEMMs for the same arithmetic. Warps own rows {warp, warp + 16, ...};
// a lane owns four columns (16-byte accesses).
__global__ void __launch_bounds__(SMALL_THREADS, 1)
backtransform_small_kernel(const float* __restrict__ VT, long long ldvt, const float* __restrict__ taus,
                           int d, int k, float* __restrict__ U, long long ldu) {
  extern __shared__ __align__(16) float smb[];
  constexpr int NW = SMALL_THREADS / 32;
  const int LDV = d | 1, LDZ = small_ld(k);
  float* Zs = smb;                    // [d][LDZ]   (first: 16-byte aligned rows)
  float* part = Zs + d * LDZ;         // [NW][SMALL_MAX]
  float* wv = part + NW * SMALL_MAX;  // [SMALL_MAX]
  float* ts = wv + SMALL_MAX;         // [SMALL_MAX] taus (a global load per reflector would be an L2
                                      // round trip on the critical path of every iteration)
  float* Vs = ts + SMALL_MAX;         // [d][LDV]   reflector i in row i (scalar broadcast reads)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < d; idx += SMALL_THREADS) ts[idx] = taus[idx];
  for (int idx = tid; idx < d * d; idx += SMALL_THREADS) {
    const int r = idx / d, c = idx - r * d;
    Vs[r * LDV + c] = (r + 1 < d) ? VT[static_cast<long long>(r) * ldvt + c] : 0.f;
  }
  for (int idx = tid; idx < d * LDZ; idx += SMALL_THREADS) {
    const int r = idx / LDZ, c = idx - r * LDZ;
    Zs[idx] = c < k ? U[static_cast<long long>(r) * ldu + c] : 0.f;
  }
  __syncthreads();
  const int c4 = lane * 4;
  const bool colive = c4 < k;
  for (int i = d - 2; i >= 0; --i) {
    const float tau = ts[i];
    if (tau == 0.f) continue;  // uniform
    const float* vi = Vs + i * LDV;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (colive)
      for (int r = i + 1 + warp; r < d; r += NW) {
        const float vr = vi[r];
        const float4 z = *reinterpret_cast<const float4*>(Zs + r * LDZ + c4);
        s.x += vr * z.x;
        s.y += vr * z.y;
        s.z += vr * z.z;
        s.w += vr * z.w;
      }
    *reinterpret_cast<float4*>(part + warp * SMALL_MAX + c4) = s;
    __syncthreads();
    if (tid < SMALL_MAX) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) t += part[w * SMALL_MAX + tid];
      wv[tid] = tau * t;
    }
    __syncthreads();
    if (colive) {
      const float4 w4 = *reinterpret_cast<const float4*>(wv + c4);
      for (int r = i + 1 + warp; r < d; r += NW) {
        const float vr = vi[r];
        float4 z = *reinterpret_cast<float4*>(Zs + r * LDZ + c4);
        z.x -= vr * w4.x;
        z.y -= vr * w4.y;
        z.z -= vr * w4.z;
        z.w -= vr * w4.w;
        *reinterpret_cast<float4*>(Zs + r * LDZ + c4) = z;
      }
    }
    // the next reflector's dots read rows written by other warps
    __syncthreads();
  }
  for (int idx = tid; idx < d * k; idx += SMALL_THREADS) {
    const int r = idx / k, cc = idx - r * k;
    U[static_cast<long long>(r) * ldu + cc] = Zs[r * LDZ + cc];
  }
}

// Compact-WY factor of one panel of reflectors stored as ROWS of VT (the resident kernel's
// layout): T[a][b] = -tau_b * sum_{c=a}^{b-1} T[a][c] (v_c . v_b), T[b][b] = tau_b (LAPACK larft,
// forward / columnwise; the panel kernel builds the same T on the fly). One CTA per panel.
__global__ void __launch_bounds__(256) larft_rows_kernel(const float* __restrict__ VT, long long ldvt,
                                                         int m0, int L, const float* __restrict__ taus,
                                                         float* __restrict__ Tmats) {
  __shared__ float tile[NB][NB + 1];  // staging of VT columns, then T
  __shared__ float Gs[NB][NB + 1];
  float (*Ts)[NB + 1] = tile;
  const int i0 = blockIdx.x * NB;
  const int nc = min(NB, m0 - i0);
  const int tid = threadIdx.x;
  const int ta = (tid >> 4) << 2, tb = (tid & 15) << 2;
  float acc[4][4];
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
  for (int c0 = (i0 / NB) * NB; c0 < L; c0 += NB) {  // reflector i0 + a is zero up to column i0 + a
    for (int idx = tid; idx < NB * NB; idx += 256) {
      const int a = idx / NB, c = idx % NB;
      tile[a][c] = (a < nc && c0 + c < L) ? VT[static_cast<long long>(i0 + a) * ldvt + c0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < NB; ++c) {
      float av[4], bv[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) { av[x] = tile[ta + x][c]; bv[x] = tile[tb + x][c]; }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] += av[x] * bv[y];
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) { Gs[ta + x][tb + y] = acc[x][y]; Ts[ta + x][tb + y] = 0.f; }
  __syncthreads();
  for (int b = 0; b < nc; ++b) {
    const float tau = taus[i0 + b];
    if (tid < b) {
      float s = 0.f;
      for (int c = tid; c < b; ++c) s += Ts[tid][c] * Gs[c][b];
      Ts[tid][b] = -tau * s;
    } else if (tid == b) {
      Ts[b][b] = tau;
    }
    __syncthreads();
  }
  float* T = Tmats + static_cast<size_t>(blockIdx.x) * NB * NB;
  for (int idx = tid; idx < NB * NB; idx += 256) T[idx] = Ts[idx / NB][idx % NB];
}

// Vs[seg][(j0 + c)][(j0 + i)] = split(VT[i][c]): the reflectors as the bf16x3 COLUMNS the
// back-transformation reads (same layout panel_split_kernel produces).
__global__ void vt_split_kernel(const float* __restrict__ VT, long long ldvt, int m0, int j0,
                                __nv_bfloat16* __restrict__ Vs, long long ldv, long long vs_seg) {
  __shared__ float tile[32][33];
  const int bi = blockIdx.y * 32, bc = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int yy = ty; yy < 32; yy += 8) {
    const int i = bi + yy, c = bc + tx;
    tile[yy][tx] = (i < m0 && c < m0) ? VT[static_cast<long long>(i) * ldvt + c] : 0.f;
  }
  __syncthreads();
  for (int yy = ty; yy < 32; yy += 8) {
    const int c = bc + yy, i = bi + tx;
    if (c < m0 && i < m0) {
      __nv_bfloat16 h, mm, l;
      split3(tile[tx][yy], h, mm, l);
      const long long o = static_cast<long long>(j0 + c) * ldv + j0 + i;
      Vs[o] = h;
      Vs[vs_seg + o] = mm;
      Vs[2 * vs_seg + o] = l;
    }
  }
}

// ============================================================================ tridiagonal stage
struct TriBufs {
  double* D;     // [d]
  double* E;     // [d] off-diagonals, E[i] couples i and i+1 (0 at splits and at d-1)
  double* E2;    // [d]
  int* blo;      // [d] first row of the unreduced block containing row i
  int* bhi;      // [d] last row
  double* lamA;  // [d] bracket of eigenvalue number (i - blo[i]) of i's block
  double* lamB;  // [d]
  int* sel;      // [k] row id of the eigenvalue that becomes output column c
  double* Dp;    // [d][k]
  double* Dm;    // [d][k]
  double* znorm; // [k]
  double* clam;  // [k] eigenvalue of column c
  double* cbn;   // [k] norm of its block
  int* clo;      // [k]
  int* chi;      // [k]
};

__global__ void __launch_bounds__(1024, 1)
tri_prep_kernel(const float* __restrict__ dvec, const float* __restrict__ evec, int d, TriBufs b) {
  __shared__ double redm[32];
  __shared__ int lastf[1024];
  __shared__ double s_tol;
  const int tid = threadIdx.x;
  double mx = 0.0;
  for (int i = tid; i < d; i += 1024) {
    const double di = dvec[i];
    const double el = i > 0 ? fabs(static_cast<double>(evec[i - 1])) : 0.0;
    const double er = i < d - 1 ? fabs(static_cast<double>(evec[i])) : 0.0;
    mx = fmax(mx, fabs(di) + el + er);
    b.D[i] = di;
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) redm[tid >> 5] = mx;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < 32; ++w) s = fmax(s, redm[w]);
    // Off-diagonals below fp32 noise of the tridiagonalisation are treated as exact splits: the
    // perturbation (<= 2^-27 ||T||) is smaller than the sytrd's own backward error.
    s_tol = s * 7.450580596923828e-09;
  }
  __syncthreads();
  const double tol = s_tol;
  for (int i = tid; i < d; i += 1024) {
    double e = (i < d - 1) ? static_cast<double>(evec[i]) : 0.0;
    if (fabs(e) <= tol) e = 0.0;
    b.E[i] = e;
    b.E2[i] = e * e;
  }
  __syncthreads();
  // block starts: row i starts a block iff i == 0 or E[i-1] == 0
  const int chunk = (d + 1023) / 1024;
  const int c0 = tid * chunk, c1 = min(d, c0 + chunk);
  int last = -1;
  for (int i = c0; i < c1; ++i)
    if (i == 0 || b.E[i - 1] == 0.0) last = i;
  lastf[tid] = last;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int t = 0; t < 1024; ++t) {
      const int mine = lastf[t];
      lastf[t] = run;  // block start in force when entering chunk t
      if (mine >= 0) run = mine;
    }
  }
  __syncthreads();
  int cur = lastf[tid];
  for (int i = c0; i < c1; ++i) {
    if (i == 0 || b.E[i - 1] == 0.0) cur = i;
    b.blo[i] = cur;
  }
  __syncthreads();
  // block ends: row i ends a block iff i == d-1 or E[i] == 0
  last = -1;
  for (int i = c1 - 1; i >= c0; --i)
    if (i == d - 1 || b.E[i] == 0.0) last = i;
  lastf[tid] = last;
  __syncthreads();
  if (tid == 0) {
    int run = d - 1;
    for (int t = 1023; t >= 0; --t) {
      const int mine = lastf[t];
      lastf[t] = run;
      if (mine >= 0) run = mine;
    }
  }
  __syncthreads();
  cur = lastf[tid];
  for (int i = c1 - 1; i >= c0; --i) {
    if (i == d - 1 || b.E[i] == 0.0) cur = i;
    b.bhi[i] = cur;
  }
}

// Number of eigenvalues of the block below x. Two forms:
//  * ratio form (LAPACK dstebz): q_i = (d_i - x) - e_{i-1}^2 / q_{i-1}, count q_i < 0. One fp64
//    DIVISION on the dependent chain per row: ~180 cycles per row on this part, and the chain is
//    what the multisection waits for (measured: 22 % of the eigensolve at d = 768, 15 % at 4096).
//  * product form (the classical Sturm sequence): p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2},
//    count sign changes of consecutive p. The chain is ONE fp64 FMA per row (e^2 p_{i-2} is ready a
//    step earlier). Its known weakness, overflow / underflow of p, is handled by working on the
//    matrix scaled to unit norm (growth <= ~3 per row) and renormalising every 8 rows; an exact
//    zero takes the sign opposite to its predecessor (the ratio form's q = -pivmin). Both forms
//    count the eigenvalues of a matrix within rounding of T - x: checked against each other and
//    against numpy on random / graded / clustered / Wilkinson / covariance tridiagonals, they
//    differ only for x within 1e-15 ||T|| of an eigenvalue.
__device__ __forceinline__ int sturm_count_ratio(const double* __restrict__ D,
                                                 const double* __restrict__ E2, int lo, int hi,
                                                 double x, double pivmin) {
  double q = __ldg(D + lo) - x;
  int cnt = q < 0.0;
  for (int i = lo + 1; i <= hi; ++i) {
    if (fabs(q) < pivmin) q = -pivmin;
    q = (__ldg(D + i) - x) - __ldg(E2 + i - 1) / q;
    cnt += (q < 0.0);
  }
  return cnt;
}

__device__ __forceinline__ int sturm_count(const double* __restrict__ D,
                                           const double* __restrict__ E2, int lo, int hi, double x,
                                           double inv_norm) {
  // The dependent chain is exactly one DFMA per row: the signs are read off the high words by the
  // integer pipe and never feed back into p. An exact zero p_i counts as a sign opposite to its
  // predecessor's and is left in place: the next term is then -e^2 p_{i-1}, which carries that same
  // opposite sign (e^2 > 0 inside an unreduced block; tri_prep_kernel zeroes off-diagonals below
  // 2^-27 ||T||, so the scaled e^2 cannot underflow), i.e. one change is counted across the zero.
  const double inv2 = inv_norm * inv_norm;
  double p0 = 1.0, p1 = (__ldg(D + lo) - x) * inv_norm;
  int n1 = (__double2hiint(p1) < 0) || (p1 == 0.0);
  int cnt = n1;
  auto step = [&](double dx, double e2) {
    const double p2 = fma(dx, p1, -(e2 * p0));
    const int n2 = (p2 == 0.0) ? (n1 ^ 1) : static_cast<int>(static_cast<unsigned>(__double2hiint(p2)) >> 31);
    cnt += n2 ^ n1;
    p0 = p1;
    p1 = p2;
    n1 = n2;
  };
  int i = lo + 1;
  for (; i + 7 <= hi; i += 8) {
    double dx[8], e2[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {  // loads and scalings are off the dependent chain
      dx[u] = (__ldg(D + i + u) - x) * inv_norm;
      e2[u] = __ldg(E2 + i + u - 1) * inv2;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) step(dx[u], e2[u]);
    // growth is at most ~3x per row on the unit-norm matrix: renormalise by the exponent of the
    // larger term every 8 rows (integer test on the high word, a power-of-two scale: exact)
    const int ex = max((__double2hiint(p0) >> 20) & 0x7ff, (__double2hiint(p1) >> 20) & 0x7ff);
    if (ex > 1023 + 300 || ex < 1023 - 300) {
      const double sc = scalbn(1.0, 1023 - ex);
      p0 *= sc;
      p1 *= sc;
    }
  }
  for (; i <= hi; ++i) step((__ldg(D + i) - x) * inv_norm, __ldg(E2 + i - 1) * inv2);
  return cnt;
}

// LPE lanes per eigenvalue: (LPE+1)-way multisection on the Sturm count, log2(LPE+1) bits per pass
// (16 lanes: 8 + 7 passes instead of 10 + 9 with 8 lanes; measured 1-3 % of the eigensolve).
// 8 lanes cost 2.5x the arithmetic of plain bisection (32 lanes: 6.4x) and still give the FP64 pipe
// 8 d independent recurrences to overlap.
// first = 1: start from the Gershgorin interval of the block; else continue from lamA/lamB.
// Small d: 16 lanes (the chain of d dependent divisions per pass is the cost, so few passes);
// d >= 1536: 4 lanes (the fp64 pipe is the cost: 5-way multisection needs 96 Sturm counts per
// eigenvalue for the two launches below against 240 with 17-way).
template <int LPE>
__global__ void bisect_kernel(TriBufs b, int d, int count, const int* __restrict__ sel, int first,
                              int passes, int ratio_form) {
  const int gthread = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = gthread / LPE;
  const int sub = threadIdx.x & (LPE - 1);                  // lane within the eigenvalue's group
  const int gshift = (threadIdx.x & 31) & ~(LPE - 1);       // first lane of the group in the warp
  const unsigned gmask = ((1u << LPE) - 1u) << gshift;
  const bool active = w < count;
  const int t = active ? (sel ? sel[w] : w) : 0;
  const int lo = b.blo[t], hi = b.bhi[t];
  const int q = t - lo;
  if (lo == hi) {
    if (active && sub == 0) { b.lamA[t] = b.D[lo]; b.lamB[t] = b.D[lo]; }
  }
  double gl = DBL_MAX, gu = -DBL_MAX, e2max = 0.0;
  for (int i = lo + sub; i <= hi; i += LPE) {
    const double el = i > lo ? fabs(b.E[i - 1]) : 0.0;
    const double er = i < hi ? fabs(b.E[i]) : 0.0;
    gl = fmin(gl, b.D[i] - el - er);
    gu = fmax(gu, b.D[i] + el + er);
    e2max = fmax(e2max, er * er);
  }
  for (int o = LPE / 2; o > 0; o >>= 1) {
    gl = fmin(gl, __shfl_xor_sync(0xffffffffu, gl, o));
    gu = fmax(gu, __shfl_xor_sync(0xffffffffu, gu, o));
    e2max = fmax(e2max, __shfl_xor_sync(0xffffffffu, e2max, o));
  }
  const double bnorm = fmax(fabs(gl), fabs(gu));
  const double pivmin = DBL_MIN * fmax(1.0, e2max);
  const double inv_norm = bnorm > 0.0 ? 1.0 / bnorm : 1.0;
  double a, bb;
  if (first) {
    const double slack = 2.1 * bnorm * DBL_EPSILON * (hi - lo + 1) + 4.2 * pivmin;
    a = gl - slack;
    bb = gu + slack;
  } else {
    a = b.lamA[t];
    bb = b.lamB[t];
  }
  const bool work = active && lo != hi;
  for (int pass = 0; pass < passes; ++pass) {
    // every lane of the warp runs every pass (the ballot is warp-wide); converged groups idle
    const bool done = !work || (bb - a <= 2.0 * DBL_EPSILON * fmax(fabs(a), fabs(bb)) + 2.0 * pivmin);
    const double wdt = (bb - a) / (LPE + 1);
    const double x = a + wdt * (sub + 1);
    const int cnt = done ? 0
                         : (ratio_form ? sturm_count_ratio(b.D, b.E2, lo, hi, x, pivmin)
                                       : sturm_count(b.D, b.E2, lo, hi, x, inv_norm));
    const unsigned mask = (__ballot_sync(0xffffffffu, !done && cnt >= q + 1) & gmask) >> gshift;
    if (!done) {
      if (mask == 0u) {
        a = a + wdt * LPE;
      } else {
        const int f = __ffs(mask) - 1;
        const double na = a + wdt * f, nb = a + wdt * (f + 1);
        a = na;
        bb = nb;
      }
    }
  }
  if (work && sub == 0) { b.lamA[t] = a; b.lamB[t] = bb; }
}

// rank of every eigenvalue in the global ascending order (ties by row id); scatter eigenvalues
// and the row ids of the k largest.
__global__ void rank_kernel(TriBufs b, int d, int k, float* __restrict__ evals) {
  __shared__ double tile[256];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const double my = t < d ? 0.5 * (b.lamA[t] + b.lamB[t]) : 0.0;
  int cnt = 0;
  for (int base = 0; base < d; base += 256) {
    const int s = base + threadIdx.x;
    tile[threadIdx.x] = s < d ? 0.5 * (b.lamA[s] + b.lamB[s]) : 0.0;
    __syncthreads();
    const int nn = min(256, d - base);
    for (int u = 0; u < nn; ++u) {
      const double v = tile[u];
      cnt += (v < my) || (v == my && (base + u) < t);
    }
    __syncthreads();
  }
  if (t < d) {
    evals[cnt] = static_cast<float>(my);
    if (cnt >= d - k) b.sel[cnt - (d - k)] = t;
  }
}

__device__ __forceinline__ double guard_piv(double x, double pivmin) {
  return fabs(x) < pivmin ? (x < 0.0 ? -pivmin : pivmin) : x;
}

// Twisted factorisation of T_block - lambda I for 32 wanted eigenvectors per block (lane = vector):
// forward L D+ L^T and backward U D- U^T are independent recurrences (one fp64 division per row
// each), so warp 0 runs the forward one while warp 1 runs the backward one; both then scan half of
// the rows for the twist index argmin |gamma|, and the two halves of z (below / above the twist)
// are again swept concurrently. Un-normalised z goes into Dp.
__global__ void __launch_bounds__(64) eigvec_kernel(TriBufs b, int d, int k) {
  __shared__ double s_best[2][32];
  __shared__ int s_r[2][32];
  __shared__ double s_nn[2][32];
  const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;  // role 0: forward / lower, 1: backward / upper
  const int c = blockIdx.x * 32 + lane;
  const bool act = c < k;
  const int t = act ? b.sel[c] : 0;
  const int lo = act ? b.blo[t] : 0, hi = act ? b.bhi[t] : 0;
  const double lam = act ? 0.5 * (b.lamA[t] + b.lamB[t]) : 0.0;
  const long long K = k;
  const bool single = lo == hi;
  double e2max = 0.0, bn = 0.0;
  if (act && !single) {
    for (int i = lo; i <= hi; ++i) {
      const double er = i < hi ? fabs(b.E[i]) : 0.0;
      const double el = i > lo ? fabs(b.E[i - 1]) : 0.0;
      e2max = fmax(e2max, er * er);
      bn = fmax(bn, fabs(b.D[i]) + el + er);
    }
  }
  const double pivmin = fmax(DBL_MIN * fmax(1.0, e2max), 1e-300);
  if (act && role == 0) {
    b.clam[c] = lam;
    b.clo[c] = lo;
    b.chi[c] = hi;
    b.cbn[c] = single ? fabs(lam) : bn;
    if (single) {
      b.Dp[lo * K + c] = 1.0;
      b.znorm[c] = 1.0;
    }
  }
  const bool work = act && !single;
  // ---- phase A: the two pivot recurrences
  if (work) {
    // D / E are fetched eight rows ahead through the read-only path: the recurrence is one
    // dependent fp64 division per row, and a load issued inside it would add its latency to every
    // row (the stores to Dp / Dm keep the compiler from hoisting plain loads)
    if (role == 0) {
      double dp = __ldg(b.D + lo) - lam;
      b.Dp[lo * K + c] = dp;
      for (int i0 = lo; i0 < hi; i0 += 8) {
        double ev[8], dv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool ok = i0 + u < hi;
          ev[u] = ok ? __ldg(b.E + i0 + u) : 0.0;
          dv[u] = ok ? __ldg(b.D + i0 + u + 1) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (i0 + u < hi) {
            const double l = ev[u] / guard_piv(dp, pivmin);
            dp = (dv[u] - lam) - l * ev[u];
            b.Dp[(i0 + u + 1) * K + c] = dp;
          }
        }
      }
    } else {
      double dm = __ldg(b.D + hi) - lam;
      b.Dm[hi * K + c] = dm;
      for (int i0 = hi - 1; i0 >= lo; i0 -= 8) {
        double ev[8], dv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool ok = i0 - u >= lo;
          ev[u] = ok ? __ldg(b.E + i0 - u) : 0.0;
          dv[u] = ok ? __ldg(b.D + i0 - u) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (i0 - u >= lo) {
            const double uu = ev[u] / guard_piv(dm, pivmin);
            dm = (dv[u] - lam) - uu * ev[u];
            b.Dm[(i0 - u) * K + c] = dm;
          }
        }
      }
    }
  }
  __threadfence_block();
  __syncthreads();
  // ---- phase B: twist index = argmin |D+ + D- - (D - lambda)|, ties to the larger row
  {
    double best = DBL_MAX;
    int r = hi;
    if (work) {
      const int mid = lo + (hi - lo) / 2;
      const int i_hi = role == 0 ? mid : hi, i_lo = role == 0 ? lo : mid + 1;
      for (int i0 = i_hi; i0 >= i_lo; i0 -= 8) {
        double dpv[8], dmv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool ok = i0 - u >= i_lo;
          dpv[u] = ok ? b.Dp[(i0 - u) * K + c] : 0.0;
          dmv[u] = ok ? b.Dm[(i0 - u) * K + c] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 - u;
          if (i >= i_lo) {
            const double gam = fabs(dpv[u] + dmv[u] - (__ldg(b.D + i) - lam));
            if (gam < best) { best = gam; r = i; }
          }
        }
      }
    }
    s_best[role][lane] = best;
    s_r[role][lane] = r;
  }
  __syncthreads();
  int r = hi;
  if (work) {  // the upper half wins ties (larger row), like a single downward scan
    r = (s_best[1][lane] <= s_best[0][lane]) ? s_r[1][lane] : s_r[0][lane];
    if (s_best[0][lane] == DBL_MAX && s_best[1][lane] == DBL_MAX) r = hi;
  }
  // ---- phase C: the two halves of z
  double nn = 0.0;
  if (work) {
    double zz = 1.0;
    if (role == 0) {
      for (int i0 = r - 1; i0 >= lo; i0 -= 8) {
        double dpv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) dpv[u] = (i0 - u >= lo) ? b.Dp[(i0 - u) * K + c] : 1.0;
        double ev[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) ev[u] = (i0 - u >= lo) ? __ldg(b.E + i0 - u) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 - u;
          if (i >= lo) {
            zz = -(ev[u] / guard_piv(dpv[u], pivmin)) * zz;
            b.Dp[i * K + c] = zz;
            nn += zz * zz;
          }
        }
      }
    } else {
      for (int i0 = r; i0 < hi; i0 += 8) {
        double dmv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) dmv[u] = (i0 + u < hi) ? b.Dm[(i0 + u + 1) * K + c] : 1.0;
        double ev[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) ev[u] = (i0 + u < hi) ? __ldg(b.E + i0 + u) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u;
          if (i < hi) {
            zz = -(ev[u] / guard_piv(dmv[u], pivmin)) * zz;
            b.Dp[(i + 1) * K + c] = zz;
            nn += zz * zz;
          }
        }
      }
    }
  }
  s_nn[role][lane] = nn;
  __syncthreads();
  if (work && role == 0) {
    b.Dp[r * K + c] = 1.0;
    b.znorm[c] = sqrt(1.0 + s_nn[0][lane] + s_nn[1][lane]);
  }
}

__device__ __forceinline__ bool same_cluster(const TriBufs& b, int c1, int c2) {
  return b.clo[c1] == b.clo[c2] && b.clo[c1] != b.chi[c1] &&
         fabs(b.clam[c2] - b.clam[c1]) <= 1e-9 * b.cbn[c1];
}

__device__ double block_sum_256(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}

// Safety net for eigenvalues of one unreduced block that agree to 1e-9 ||T_b||: the independently
// computed vectors are re-orthogonalised (modified Gram-Schmidt, twice) in fp64. One CTA per
// cluster head; every other CTA exits at once.
__global__ void __launch_bounds__(256) cluster_fix_kernel(TriBufs b, int k) {
  __shared__ double red[8];
  const int c = blockIdx.x;
  if (c > 0 && same_cluster(b, c - 1, c)) return;
  if (!(c + 1 < k && same_cluster(b, c, c + 1))) return;
  const int lo = b.clo[c], hi = b.chi[c];
  const long long K = k;
  const int tid = threadIdx.x;
  {
    const double inv = 1.0 / b.znorm[c];
    for (int i = lo + tid; i <= hi; i += 256) b.Dp[i * K + c] *= inv;
    __syncthreads();
    if (tid == 0) b.znorm[c] = 1.0;
  }
  for (int mcol = c + 1; mcol < k && same_cluster(b, mcol - 1, mcol); ++mcol) {
    double inv = 1.0 / b.znorm[mcol];
    for (int i = lo + tid; i <= hi; i += 256) b.Dp[i * K + mcol] *= inv;
    __syncthreads();
    for (int rep = 0; rep < 2; ++rep) {
      for (int p = c; p < mcol; ++p) {
        double dot = 0.0;
        for (int i = lo + tid; i <= hi; i += 256) dot += b.Dp[i * K + mcol] * b.Dp[i * K + p];
        dot = block_sum_256(dot, red);
        for (int i = lo + tid; i <= hi; i += 256) b.Dp[i * K + mcol] -= dot * b.Dp[i * K + p];
        __syncthreads();
      }
      double nn = 0.0;
      for (int i = lo + tid; i <= hi; i += 256) nn += b.Dp[i * K + mcol] * b.Dp[i * K + mcol];
      nn = sqrt(block_sum_256(nn, red));
      if (nn < 1e-10 && rep == 0) {
        // numerically identical vectors: restart from a unit vector of the block
        const int pick = lo + (mcol - c) % (hi - lo + 1);
        for (int i = lo + tid; i <= hi; i += 256) b.Dp[i * K + mcol] = (i == pick) ? 1.0 : 0.0;
        __syncthreads();
        rep = -1;  // orthogonalise the replacement twice as well
        continue;
      }
      inv = nn > 0.0 ? 1.0 / nn : 0.0;
      for (int i = lo + tid; i <= hi; i += 256) b.Dp[i * K + mcol] *= inv;
      __syncthreads();
    }
    if (tid == 0) b.znorm[mcol] = 1.0;
    __syncthreads();
  }
}

// U[i][c] = z / ||z|| inside the block of column c, 0 outside.
__global__ void z_to_u_kernel(TriBufs b, int d, int k, float* __restrict__ U, long long ldu) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(d) * k) return;
  const int i = static_cast<int>(idx / k), c = static_cast<int>(idx % k);
  float x = 0.f;
  if (i >= b.clo[c] && i <= b.chi[c]) x = static_cast<float>(b.Dp[idx] / b.znorm[c]);
  U[static_cast<long long>(i) * ldu + c] = x;
}

// ============================================================================ back-transform
// Zs[seg][r][c] = split(U[row0 + r][c]) for r < m, c < kp (zero beyond k).
__global__ void split_z_kernel(const float* __restrict__ U, long long ldu, int row0, int m, int k,
                               int kp, __nv_bfloat16* __restrict__ Zs, long long seg) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int groups = kp / 8;
  if (idx >= static_cast<long long>(m) * groups) return;
  const int r = static_cast<int>(idx / groups), c0 = static_cast<int>(idx % groups) * 8;
  __align__(16) __nv_bfloat16 o[3][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c0 + e;
    const float x = c < k ? U[static_cast<long long>(row0 + r) * ldu + c] : 0.f;
    split3(x, o[0][e], o[1][e], o[2][e]);
  }
#pragma unroll
  for (int s = 0; s < 3; ++s)
    *reinterpret_cast<uint4*>(Zs + s * seg + static_cast<long long>(r) * kp + c0) =
        *reinterpret_cast<const uint4*>(o[s]);
}

// Xs[seg][a][n] = split( sum_{b >= a} T[a][b] X1[b][n] ), a < NB (rows >= nc give zero).
__global__ void tmul_split_kernel(const float* __restrict__ T, const float* __restrict__ X1, int nc,
                                  int kp, __nv_bfloat16* __restrict__ Xs, long long seg) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int a = blockIdx.y;
  if (n >= kp) return;
  float s = 0.f;
  if (a < nc)
    for (int bb = a; bb < nc; ++bb) s += T[a * NB + bb] * X1[static_cast<long long>(bb) * kp + n];
  __nv_bfloat16 h, mm, l;
  split3(s, h, mm, l);
  Xs[static_cast<long long>(a) * kp + n] = h;
  Xs[seg + static_cast<long long>(a) * kp + n] = mm;
  Xs[2 * seg + static_cast<long long>(a) * kp + n] = l;
}

__global__ void tiny_1x1_kernel(const float* A, float* evals, float* U) {
  evals[0] = A[0];
  U[0] = 1.f;
}

// ---------------------------------------------------------------------------- workspace carving
struct Carver {
  uint8_t* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t count) {
    off = (off + 255) & ~static_cast<size_t>(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

struct Plan {
  float* Aw; long long ldA;
  __nv_bfloat16* Vs; long long ldv;
  float *Vp, *Wp, *colbuf, *ptop, *dvec, *evec, *taus, *Tmats, *X1;
  double *nacc, *gacc, *pglob;
  unsigned* bars;
  __nv_bfloat16 *VW, *WV, *Zs, *Xs;
  TriBufs tb;
  int kp, npanels;
  float* VT; long long ldvt; int res_cap;  // resident stage: reflector rows, exchange buffers, counters
  float* xbuf; unsigned* rctr;
  size_t bytes;
};

// Where the resident kernel takes over: the first panel boundary j0 from which the trailing block
// fits the shared memory of one cooperative grid (j0 = 0: the whole matrix).
struct ResShape { int valid, j0, m0, L, G, nloc; size_t smem; };
constexpr size_t RES_SMEM_LIMIT = 227 * 1024 - 512;  // opt-in maximum minus the kernel's static part
size_t res_smem_bytes(int nloc, int L) {
  return (static_cast<size_t>(nloc) * L + 4 * static_cast<size_t>(L) + 2 * RES_MAX_NLOC +
          RES_WARPS * RES_MAX_NLOC) * sizeof(float);
}
ResShape resident_shape(int d, int sms, int rows_target) {
  ResShape rs{};
  for (int j0 = 0; j0 < d; j0 += NB) {
    const int m0 = d - j0;
    const int L = static_cast<int>(round_up(m0, 4));
    if (L > RES_MAX_L) continue;
    int G = std::min(sms, std::max(1, (m0 + rows_target - 1) / rows_target));
    int nloc = (m0 + G - 1) / G;
    if (nloc > RES_MAX_NLOC) {
      G = (m0 + RES_MAX_NLOC - 1) / RES_MAX_NLOC;
      if (G > sms) continue;
      nloc = (m0 + G - 1) / G;
    }
    const size_t smem = res_smem_bytes(nloc, L);
    if (smem > RES_SMEM_LIMIT) continue;
    rs = ResShape{1, j0, m0, L, G, nloc, smem};
    return rs;
  }
  return rs;
}

int max_grid() { return device_sm_count(); }

// cudaFuncSetAttribute is per device: remember which devices were configured.
bool* attr_flag(int which) {
  static bool done[4][64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  return &done[which][dev & 63];
}

Plan make_plan(void* ws, int d, int k) {
  Plan p;
  Carver cv{static_cast<uint8_t*>(ws)};
  const size_t D = d, K = k;
  p.ldA = round_up(d, 8);
  p.ldv = round_up(d, 8);
  p.kp = static_cast<int>(round_up(k, 8));
  p.npanels = (d + NB - 1) / NB;
  p.Aw = cv.take<float>(D * p.ldA);
  p.Vs = cv.take<__nv_bfloat16>(3 * D * p.ldv);
  p.Vp = cv.take<float>(D * NB);
  p.Wp = cv.take<float>(D * NB);
  p.colbuf = cv.take<float>(D);
  // per-panel accumulators for the cross-CTA reductions, zeroed once for all panels
  p.nacc = cv.take<double>(static_cast<size_t>(p.npanels) * NB);
  p.gacc = cv.take<double>(static_cast<size_t>(p.npanels) * NB * GP_STRIDE);
  p.ptop = cv.take<float>(NB);
  p.pglob = cv.take<double>(2 * static_cast<size_t>(p.ldA));
  p.dvec = cv.take<float>(D);
  p.evec = cv.take<float>(D);
  p.taus = cv.take<float>(D);
  p.Tmats = cv.take<float>(static_cast<size_t>(p.npanels) * NB * NB);
  p.bars = cv.take<unsigned>(p.npanels);
  p.VW = cv.take<__nv_bfloat16>(3 * D * 2 * NB);
  p.WV = cv.take<__nv_bfloat16>(3 * D * 2 * NB);
  p.tb.D = cv.take<double>(D);
  p.tb.E = cv.take<double>(D);
  p.tb.E2 = cv.take<double>(D);
  p.tb.blo = cv.take<int>(D);
  p.tb.bhi = cv.take<int>(D);
  p.tb.lamA = cv.take<double>(D);
  p.tb.lamB = cv.take<double>(D);
  p.tb.sel = cv.take<int>(K);
  p.tb.Dp = cv.take<double>(D * K);
  p.tb.Dm = cv.take<double>(D * K);
  p.tb.znorm = cv.take<double>(K);
  p.tb.clam = cv.take<double>(K);
  p.tb.cbn = cv.take<double>(K);
  p.tb.clo = cv.take<int>(K);
  p.tb.chi = cv.take<int>(K);
  p.res_cap = static_cast<int>(std::min<long long>(d, RES_MAX_L));
  p.ldvt = round_up(p.res_cap, 4);
  p.VT = cv.take<float>(static_cast<size_t>(p.res_cap) * p.ldvt);
  p.xbuf = cv.take<float>(4 * static_cast<size_t>(p.ldvt));
  p.rctr = cv.take<unsigned>(static_cast<size_t>(p.res_cap));
  p.Zs = cv.take<__nv_bfloat16>(3 * D * p.kp);
  p.X1 = cv.take<float>(static_cast<size_t>(NB) * p.kp);
  p.Xs = cv.take<__nv_bfloat16>(3 * static_cast<size_t>(NB) * p.kp);
  p.bytes = cv.off + 256;
  return p;
}

#define PTD_CHECK_LAUNCH()                                   \
  do {                                                       \
    if (cudaGetLastError() != cudaSuccess) return -5;        \
  } while (0)

}  // namespace

static int g_panel_prof = 0;
static int g_res_enable = 1;  // 0: blocked panel kernel all the way (round-1 path)
static int g_res_rows = 4;    // target rows per CTA of the resident kernel (grid = m0 / rows, <= SMs)
static int g_jacobi_max = JACOBI_DEFAULT;
static int g_small_enable = 1;  // d <= 128: single-CTA tridiagonalisation instead of the resident kernel
void eigh_debug_small(int on) { g_small_enable = on; }
// From this d on the multisection would use 4 lanes per eigenvalue instead of 16. Measured on a
// B200 (d = 2048 / 4096): 4 lanes are 1.5-2 ms SLOWER -- the stage is bound by the chain of d
// dependent fp64 divisions per pass, not by the fp64 pipe, so fewer, wider passes win. Kept as a
// knob (ptdeco_debug_set key 105).
static int g_bisect_narrow_d = 1 << 30;
static int g_sturm_ratio = 0;  // 1: the division-based Sturm count (LAPACK's form), for A/B checks
void eigh_debug_resident(int enable, int rows_target, int jacobi_max) {
  g_res_enable = enable;
  if (rows_target > 0) g_res_rows = rows_target;
  if (jacobi_max >= 0) g_jacobi_max = std::min(jacobi_max, JACOBI_MAX);
}
void eigh_debug_bisect_narrow(int d) { g_bisect_narrow_d = d > 0 ? d : (1 << 30); }
void eigh_debug_sturm_ratio(int on) { g_sturm_ratio = on ? 1 : 0; }
static int g_sym_min_m = 5120;  // trailing size from which the symv reads only the lower triangle
void eigh_debug_sym_min_m(int m) { g_sym_min_m = m; }
void eigh_debug_profile(int enable) {
  g_panel_prof = enable;
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
}
long long eigh_debug_phase_cycles(int k) {
  unsigned long long v[16];
  cudaMemcpyFromSymbol(v, g_phase_cycles, sizeof(v));
  return (k >= 0 && k < 16) ? static_cast<long long>(v[k]) : 0;
}

size_t eigh_workspace_bytes(int d, int k) {
  if (d <= g_jacobi_max) return 256;
  return make_plan(nullptr, d, k).bytes;
}

int eigh(const float* A, int d, long long lda, int k, float* evals, float* U, long long ldu,
         void* ws, size_t ws_bytes, cudaStream_t st, unsigned flags) {
  const int deterministic = (flags & 1u) ? 1 : 0;
  if (A == nullptr || evals == nullptr || U == nullptr || d <= 0 || k < 1 || k > d || lda < d ||
      ldu < k)
    return -22;
  if (d == 1) {
    tiny_1x1_kernel<<<1, 1, 0, st>>>(A, evals, U);
    PTD_CHECK_LAUNCH();
    return 0;
  }
  if (d <= g_jacobi_max) {
    const int n = d + (d & 1);
    const size_t smem = (2 * static_cast<size_t>(d) * (d + 1) + n) * sizeof(double) +
                        (static_cast<size_t>(n) + d) * sizeof(int) + 64;
    bool* attr = attr_flag(0);
    if (!*attr) {
      if (cudaFuncSetAttribute(jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               200 * 1024) != cudaSuccess)
        return -12;
      *attr = true;
    }
    jacobi_kernel<<<1, 512, smem, st>>>(A, lda, d, k, evals, U, ldu);
    PTD_CHECK_LAUNCH();
    return 0;
  }
  if (ws == nullptr || ws_bytes < eigh_workspace_bytes(d, k)) return -12;
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return -22;
  Plan p = make_plan(ws, d, k);

  // ---- copy-in (mirrors the lower triangle, zero pad columns)
  {
    dim3 grid(static_cast<unsigned>((p.ldA + 31) / 32), static_cast<unsigned>((d + 31) / 32));
    copy_sym_kernel<<<grid, dim3(32, 8), 0, st>>>(A, lda, d, p.Aw, p.ldA);
    PTD_CHECK_LAUNCH();
  }
  cudaMemsetAsync(p.Tmats, 0, static_cast<size_t>(p.npanels) * NB * NB * sizeof(float), st);
  cudaMemsetAsync(p.bars, 0, static_cast<size_t>(p.npanels) * sizeof(unsigned), st);
  cudaMemsetAsync(p.nacc, 0, static_cast<size_t>(p.npanels) * NB * sizeof(double), st);
  cudaMemsetAsync(p.gacc, 0, static_cast<size_t>(p.npanels) * NB * GP_STRIDE * sizeof(double), st);
  cudaMemsetAsync(p.Vp, 0, static_cast<size_t>(d) * NB * sizeof(float), st);
  cudaMemsetAsync(p.Wp, 0, static_cast<size_t>(d) * NB * sizeof(float), st);
  cudaMemsetAsync(p.evec, 0, static_cast<size_t>(d) * sizeof(float), st);
  cudaMemsetAsync(p.taus, 0, static_cast<size_t>(d) * sizeof(float), st);
  cudaMemsetAsync(p.pglob, 0, 2 * static_cast<size_t>(p.ldA) * sizeof(double), st);

  // ---- (1) tridiagonalisation
  bool* panel_attr = attr_flag(1);
  if (!*panel_attr) {
    if (cudaFuncSetAttribute(sytrd_panel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             220 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(sytrd_panel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             200 * 1024) != cudaSuccess ||  // + 24 KB of static tile accumulators
        cudaFuncSetAttribute(sytd2_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(RES_SMEM_LIMIT)) != cudaSuccess)
      return -12;
    *panel_attr = true;
  }
  const int sms = max_grid();
  const ResShape rs = g_res_enable ? resident_shape(d, sms, g_res_rows) : ResShape{};
  const int blocked_panels = rs.valid ? rs.j0 / NB : p.npanels;
  struct PanelShape { int rows_per_cta, grid, sym, sym_tile; size_t smem; };
  // Lower-triangle symv while the trailing matrix is too large for the L2 (full rows otherwise:
  // L2-resident, and the per-tile atomics would only add latency). Needs its tile accumulators
  // (24 KB static) next to the dynamic shared memory and at most SYM_MAX_SLOTS tiles per CTA.
  auto shape_of = [&](int pi) {
    PanelShape ps;
    const int j0 = pi * NB;
    const int m = d - j0;
    ps.rows_per_cta = std::max(PANEL_WARPS, (m + sms - 1) / sms);
    ps.grid = (m + ps.rows_per_cta - 1) / ps.rows_per_cta;
    const long long L = p.ldA - j0;
    const size_t nseg_max = static_cast<size_t>((L / 4 + 255) / 256);
    ps.smem = (static_cast<size_t>(L) + 3 * NB * (NB + 1) + ps.rows_per_cta + 3 * NB +
               PANEL_WARPS * 2 * NB + ps.rows_per_cta * nseg_max) * sizeof(float);
    ps.sym = (g_sym_min_m > 0 && m >= g_sym_min_m) ? 1 : 0;
    ps.sym_tile = static_cast<int>(std::min<long long>(256, round_up((m + 32) / 33, 4)));
    const long long kk = (m + ps.sym_tile - 1) / ps.sym_tile;
    const long long tiles_per_cta = (kk * (kk + 1) / 2 + ps.grid - 1) / ps.grid;
    if (tiles_per_cta > SYM_MAX_SLOTS || ps.smem > 200 * 1024) ps.sym = 0;
    return ps;
  };
  for (int pi = 0; pi < blocked_panels; ++pi) {
    const int j0 = pi * NB;
    const int m = d - j0;
    const PanelShape ps = shape_of(pi);
    PanelArgs g;
    g.A = p.Aw; g.ldA = p.ldA; g.d = d; g.j0 = j0; g.ncols = std::min(NB, m);
    g.rows_per_cta = ps.rows_per_cta;
    g.Vp = p.Vp; g.Wp = p.Wp; g.colbuf = p.colbuf;
    g.nacc = p.nacc + static_cast<size_t>(pi) * NB;
    g.gacc = p.gacc + static_cast<size_t>(pi) * NB * GP_STRIDE;
    g.ptop = p.ptop; g.dvec = p.dvec; g.evec = p.evec; g.taus = p.taus;
    g.Tmat = p.Tmats + static_cast<size_t>(pi) * NB * NB;
    g.bar = p.bars + pi;
    g.pglob = p.pglob;
    g.ldp = p.ldA;
    g.prof = g_panel_prof;
    const int grid = ps.grid;
    const size_t smem = ps.smem;
    if (smem > 220 * 1024) return -22;  // d beyond what one SM's shared memory can stage
    g.sym = ps.sym;
    g.sym_tile = ps.sym_tile;
    void* args[] = {&g};
    void* kern = g.sym ? reinterpret_cast<void*>(sytrd_panel_kernel<true>)
                       : reinterpret_cast<void*>(sytrd_panel_kernel<false>);
    if (cudaLaunchCooperativeKernel(kern, dim3(grid),
                                    dim3(PANEL_THREADS), args, smem, st) != cudaSuccess)
      return -5;
    {
      const int total = m * (NB / 8);
      panel_split_kernel<<<(total + 255) / 256, 256, 0, st>>>(
          p.Vp, p.Wp, j0, m, g.ncols, p.Vs, p.ldv, static_cast<long long>(d) * p.ldv, p.VW, p.WV,
          static_cast<long long>(d) * 2 * NB);
      PTD_CHECK_LAUNCH();
    }
    if (m > NB) {
      const int mt = m - NB;
      GemmOperand a{p.VW + static_cast<long long>(j0 + NB) * 2 * NB, 0, 2 * NB, 3,
                    static_cast<long long>(d) * 2 * NB};
      GemmOperand b{p.WV + static_cast<long long>(j0 + NB) * 2 * NB, 0, 2 * NB, 3,
                    static_cast<long long>(d) * 2 * NB};
      GemmEpilogue ep;
      ep.alpha = -1.f;
      ep.C = p.Aw + static_cast<long long>(j0 + NB) * p.ldA + (j0 + NB);
      ep.ldc = p.ldA;
      ep.accumulate = 1;
      // While the following panel reads only the lower triangle, only that half is updated
      // (V W^T + W V^T is symmetric); when the panels switch back to full rows the block is
      // mirrored once.
      const int next_sym = (pi + 1 < p.npanels) ? shape_of(pi + 1).sym : 0;
      ep.lower_only = next_sym;
      ep.deterministic = deterministic;
      const int rc = gemm_tc(a, b, mt, mt, 2 * NB, -1, ep, st);
      if (rc) return rc;
      if (g.sym && !next_sym) {
        dim3 mgrid(static_cast<unsigned>((mt + 31) / 32), static_cast<unsigned>((mt + 31) / 32));
        mirror_lower_kernel<<<mgrid, dim3(32, 8), 0, st>>>(ep.C, p.ldA, mt);
        PTD_CHECK_LAUNCH();
      }
    }
  }

  const bool small_path = rs.valid && rs.j0 == 0 && d <= SMALL_MAX && g_small_enable;
  if (rs.valid) {
    // The rest of the reduction in ONE launch with the trailing block resident in shared memory.
    cudaMemsetAsync(p.VT, 0, static_cast<size_t>(rs.m0) * p.ldvt * sizeof(float), st);
    cudaMemsetAsync(p.rctr, 0, static_cast<size_t>(rs.m0) * sizeof(unsigned), st);
    ResArgs g;
    g.A = p.Aw; g.ldA = p.ldA; g.j0 = rs.j0; g.m0 = rs.m0; g.L = rs.L; g.nloc = rs.nloc;
    g.VT = p.VT; g.ldvt = p.ldvt;
    g.dvec = p.dvec; g.evec = p.evec; g.taus = p.taus;
    g.xP = p.xbuf; g.xR = p.xbuf + 2 * static_cast<size_t>(rs.L); g.ctr = p.rctr;
    g.prof = g_panel_prof;
    void* args[] = {&g};
    if (small_path) {
      const size_t smem = (static_cast<size_t>(d) * (((d + 3) & ~3) + 4) + 2 * SMALL_LD) * sizeof(float);
      bool* sattr = attr_flag(2);
      if (!*sattr) {
        if (cudaFuncSetAttribute(sytd2_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (SMALL_MAX * SMALL_LD + 2 * SMALL_LD) * 4) != cudaSuccess)
          return -12;
        *sattr = true;
      }
      sytd2_small_kernel<<<1, SMALL_THREADS, smem, st>>>(p.Aw, p.ldA, d, rs.L, p.VT, p.ldvt, p.dvec,
                                                         p.evec, p.taus, g_panel_prof);
      PTD_CHECK_LAUNCH();
    } else if (cudaLaunchCooperativeKernel(reinterpret_cast<void*>(sytd2_resident_kernel), dim3(rs.G),
                                           dim3(RES_THREADS), args, rs.smem, st) != cudaSuccess) {
      return -5;
    }
    if (!small_path) {  // (the one-CTA back-transformation reads VT and the taus directly)
      const int rpanels = (rs.m0 + NB - 1) / NB;
      larft_rows_kernel<<<rpanels, 256, 0, st>>>(p.VT, p.ldvt, rs.m0, rs.L, p.taus + rs.j0,
                                                 p.Tmats + static_cast<size_t>(rs.j0 / NB) * NB * NB);
      PTD_CHECK_LAUNCH();
      dim3 sgrid(static_cast<unsigned>((rs.m0 + 31) / 32), static_cast<unsigned>((rs.m0 + 31) / 32));
      vt_split_kernel<<<sgrid, dim3(32, 8), 0, st>>>(p.VT, p.ldvt, rs.m0, rs.j0, p.Vs, p.ldv,
                                                     static_cast<long long>(d) * p.ldv);
      PTD_CHECK_LAUNCH();
    }
  }

  // ---- (2) tridiagonal eigenproblem in fp64
  tri_prep_kernel<<<1, 1024, 0, st>>>(p.dvec, p.evec, d, p.tb);
  PTD_CHECK_LAUNCH();
  {
    // all eigenvalues to ~2^-31 of the block norm (enough for ordering and the fp32 output), then
    // the k wanted ones to full fp64 precision (the pass loop stops at convergence)
    const int tpb = 128;
    const bool wide = d < g_bisect_narrow_d;
    const int pb = tpb / (wide ? 16 : 4);
    if (wide) bisect_kernel<16><<<(d + pb - 1) / pb, tpb, 0, st>>>(p.tb, d, d, nullptr, 1, 8, g_sturm_ratio);
    else bisect_kernel<4><<<(d + pb - 1) / pb, tpb, 0, st>>>(p.tb, d, d, nullptr, 1, 14, g_sturm_ratio);
    PTD_CHECK_LAUNCH();
    rank_kernel<<<(d + 255) / 256, 256, 0, st>>>(p.tb, d, k, evals);
    PTD_CHECK_LAUNCH();
    if (wide) bisect_kernel<16><<<(k + pb - 1) / pb, tpb, 0, st>>>(p.tb, d, k, p.tb.sel, 0, 7, g_sturm_ratio);
    else bisect_kernel<4><<<(k + pb - 1) / pb, tpb, 0, st>>>(p.tb, d, k, p.tb.sel, 0, 11, g_sturm_ratio);
    PTD_CHECK_LAUNCH();
  }
  eigvec_kernel<<<(k + 31) / 32, 64, 0, st>>>(p.tb, d, k);
  PTD_CHECK_LAUNCH();
  cluster_fix_kernel<<<k, 256, 0, st>>>(p.tb, k);
  PTD_CHECK_LAUNCH();
  {
    const long long total = static_cast<long long>(d) * k;
    z_to_u_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(p.tb, d, k, U, ldu);
    PTD_CHECK_LAUNCH();
  }

  // ---- (3) back-transformation: U <- (I - V_p T_p V_p^T) U for p = last .. first
  if (small_path) {
    const size_t smem = (static_cast<size_t>(d) * (d | 1) + static_cast<size_t>(d) * (((k + 3) & ~3) + 4) +
                         (SMALL_THREADS / 32 + 2) * SMALL_MAX) * sizeof(float);
    bool* battr = attr_flag(3);
    if (!*battr) {
      if (cudaFuncSetAttribute(backtransform_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (SMALL_MAX * (SMALL_MAX | 1) + SMALL_MAX * SMALL_LD +
                                (SMALL_THREADS / 32 + 2) * SMALL_MAX) * 4) != cudaSuccess)
        return -12;
      *battr = true;
    }
    backtransform_small_kernel<<<1, SMALL_THREADS, smem, st>>>(p.VT, p.ldvt, p.taus, d, k, U, ldu);
    PTD_CHECK_LAUNCH();
    return 0;
  }
  const long long zseg = static_cast<long long>(d) * p.kp;
  const long long xseg = static_cast<long long>(NB) * p.kp;
  for (int pi = p.npanels - 1; pi >= 0; --pi) {
    const int j0 = pi * NB;
    const int m = d - j0;
    const int nc = std::min(NB, m);
    if (m < 2) continue;  // a 1x1 trailing block has no reflector
    {
      const long long total = static_cast<long long>(m) * (p.kp / 8);
      split_z_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
          U, ldu, j0, m, k, p.kp, p.Zs, zseg);
      PTD_CHECK_LAUNCH();
    }
    cudaMemsetAsync(p.X1, 0, static_cast<size_t>(NB) * p.kp * sizeof(float), st);
    const __nv_bfloat16* vblock = p.Vs + static_cast<long long>(j0) * p.ldv + j0;
    {
      GemmOperand a{vblock, 1, p.ldv, 3, static_cast<long long>(d) * p.ldv};
      GemmOperand b{p.Zs, 1, p.kp, 3, zseg};
      GemmEpilogue ep;
      ep.C = p.X1;
      ep.ldc = p.kp;
      ep.accumulate = 1;
      ep.deterministic = deterministic;
      const int rc = gemm_tc(a, b, nc, k, m, -1, ep, st);
      if (rc) return rc;
    }
    {
      dim3 grid(static_cast<unsigned>((p.kp + 127) / 128), NB);
      tmul_split_kernel<<<grid, 128, 0, st>>>(p.Tmats + static_cast<size_t>(pi) * NB * NB, p.X1,
                                              nc, p.kp, p.Xs, xseg);
      PTD_CHECK_LAUNCH();
    }
    {
      GemmOperand a{vblock, 0, p.ldv, 3, static_cast<long long>(d) * p.ldv};
      GemmOperand b{p.Xs, 1, p.kp, 3, xseg};
      GemmEpilogue ep;
      ep.alpha = -1.f;
      ep.C = U + static_cast<long long>(j0) * ldu;
      ep.ldc = ldu;
      ep.accumulate = 1;
      ep.deterministic = deterministic;
      const int rc = gemm_tc(a, b, m, k, nc, -1, ep, st);
      if (rc) return rc;
    }
  }
  return 0;
}

}  // namespace ptd
