// tcgen05 GEMM engine for sm_100a. See gemm_tc.cuh for what it computes and who uses it.
//
// Kernel anatomy (one CTA per SM, persistent over work units = output tile x k-split):
//   warp 0        TMA producer: cp.async.bulk.tensor (128B swizzle) into a STAGES-deep smem ring
//   warp 1        TMEM allocation + single-thread tcgen05.mma issue, tcgen05.commit to mbarriers
//   warps 2..9    epilogue: tcgen05.ld TMEM -> fp32 register sums -> global (store / red.add / bf16)
// Accumulators are double-buffered in TMEM (2 x TILE_N fp32 columns): the tensor core fills one
// buffer for CHUNK_ITERS k-blocks while the epilogue warps drain the other into registers.
#include "gemm_tc.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ptx.cuh"

namespace ptd {

namespace {

constexpr int TILE_M = 128;
constexpr int BLOCK_K = 64;  // bf16 elements per k-block = one 128-byte swizzle span
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;  // producer + MMA + 8 epilogue warps
// The tensor core adds into its fp32 TMEM accumulator with truncation, which biases long
// reductions (measured 1.8e-5 relative at K = 8192). TMEM accumulation is therefore bounded to
// CHUNK_ITERS k-blocks (64 MMA k-steps); completed chunks are summed in fp32 registers (RN) by
// the epilogue warps while the tensor core fills the other TMEM buffer.
constexpr int CHUNK_ITERS = 16;
// fp32-grade (bf16x3) contractions feed the eigensolver and the factors: 4 k-blocks per chunk
// measures 3.3e-7 relative (vs 1.0e-6 at 16) for 10 % of the throughput.
constexpr int CHUNK_ITERS_SPLIT = 4;
constexpr int MAX_PAIRS = 6;
constexpr int A_BYTES = TILE_M * BLOCK_K * 2;
constexpr int GROUP_BYTES = 64 * BLOCK_K * 2;  // one 64-wide MN group of an MN-major tile

template <int TILE_N>
struct Cfg {
  static constexpr int B_BYTES = TILE_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (TILE_N == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * TILE_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct KArgs {
  int M, N;
  int kblocks;  // k-blocks per segment pair
  int npairs;
  int pair_a[MAX_PAIRS], pair_b[MAX_PAIRS];
  int tiles_n, ntiles, splitk;
  int chunk_iters;
  int lower, accumulate;
  float alpha;
  float* C;
  long long ldc;
  __nv_bfloat16* Cb;
  long long ldcb;
  __nv_bfloat16* Cs;
  long long ldcs, cs_seg;
  const float* bias;
  uint32_t lbo_mn, sbo_mn, lbo_k, sbo_k;
  uint32_t sleep_ns;  // back-off of the long waits (epilogue / producer)
  int band;           // raster band (tile rows) of the plain-GEMM / SYRK tile order in the pair kernel
};

template <int TILE_N>
__device__ __forceinline__ void decode_tile(const KArgs& g, int t, int& m0, int& n0) {
  if (!g.lower) {
    m0 = (t / g.tiles_n) * TILE_M;
    n0 = (t % g.tiles_n) * TILE_N;
    return;
  }
  if (TILE_N == 128) {
    // row block i holds tiles j = 0..i ; cumulative i(i+1)/2
    int i = static_cast<int>((sqrtf(8.f * static_cast<float>(t) + 1.f) - 1.f) * 0.5f);
    while (i * (i + 1) / 2 > t) --i;
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    m0 = i * TILE_M;
    n0 = (t - i * (i + 1) / 2) * TILE_N;
  } else {
    // 256-wide column blocks: row block i (128 rows) owns column blocks j <= i/2. Tiles are
    // enumerated by super-rows of SUPER row blocks, column-major inside a super-row, so that
    // the ~148 tiles in flight share ~16 A slabs and ~9 B slabs (~70 MB, L2-resident) instead
    // of sweeping the whole operand once per row block (measured 4x the algorithmic DRAM reads).
    constexpr int SUPER = 16;
    const int tiles_m = (g.M + TILE_M - 1) / TILE_M;
    int i = 0, j = 0;
    for (int lo = 0; lo < tiles_m; lo += SUPER) {
      const int hi = min(tiles_m, lo + SUPER);
      const int nrows = hi - lo;
      const int full_cols = lo / 2;  // column blocks every row block of the super-row owns
      const int cnt_full = full_cols * nrows;
      if (t < cnt_full) {
        j = t / nrows;
        i = lo + t - j * nrows;
        break;
      }
      t -= cnt_full;
      bool found = false;
      for (int jj = full_cols; 2 * jj < hi; ++jj) {  // the staircase at the diagonal
        const int first = max(lo, 2 * jj);
        const int c = hi - first;
        if (t < c) {
          j = jj;
          i = first + t;
          found = true;
          break;
        }
        t -= c;
      }
      if (found) break;
    }
    m0 = i * TILE_M;
    n0 = j * TILE_N;
  }
}

// Epilogue write-out of one thread's row slice: sum[COLS] holds row m, columns [nbase, nbase+COLS)
// of alpha-unscaled fp32 results. Scales, adds the bias and writes every requested output form.
template <int COLS>
__device__ __forceinline__ void write_out(const KArgs& g, float (&sum)[COLS], int m, int nbase,
                                          int niter, bool vec_ok) {
  if (m < g.M && nbase < g.N && niter > 0) {
#pragma unroll
    for (int j = 0; j < COLS; ++j) sum[j] *= g.alpha;
    if (g.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < COLS; ++j)
        if (nbase + j < g.N) sum[j] += __ldg(g.bias + nbase + j);
    }
    if (g.C != nullptr) {
      float* crow = g.C + static_cast<long long>(m) * g.ldc + nbase;
      if (vec_ok && nbase + COLS <= g.N) {
        if (g.accumulate) {
#pragma unroll
          for (int j = 0; j < COLS; j += 4)
            red_add_v4(crow + j, sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < COLS; j += 4)
            *reinterpret_cast<float4*>(crow + j) =
                make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < COLS; ++j) {
          if (nbase + j < g.N) {
            if (g.accumulate)
              atomicAdd(crow + j, sum[j]);
            else
              crow[j] = sum[j];
          }
        }
      }
    }
    if (g.Cb != nullptr) {
      __nv_bfloat16* brow = g.Cb + static_cast<long long>(m) * g.ldcb + nbase;
      const bool bvec = ((g.ldcb & 7) == 0) && ((reinterpret_cast<uintptr_t>(g.Cb) & 15) == 0) &&
                        nbase + COLS <= g.N;
      if (bvec) {
#pragma unroll
        for (int j = 0; j < COLS; j += 8) {
          __align__(16) __nv_bfloat16 o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = __float2bfloat16_rn(sum[j + e]);
          *reinterpret_cast<uint4*>(brow + j) = *reinterpret_cast<const uint4*>(o);
        }
      } else {
#pragma unroll
        for (int j = 0; j < COLS; ++j)
          if (nbase + j < g.N) brow[j] = __float2bfloat16_rn(sum[j]);
      }
    }
    if (g.Cs != nullptr) {
      __nv_bfloat16* srow = g.Cs + static_cast<long long>(m) * g.ldcs + nbase;
#pragma unroll
      for (int j = 0; j < COLS; ++j) {
        if (nbase + j < g.N) {
          const float x = sum[j];
          const __nv_bfloat16 h = __float2bfloat16_rn(x);
          const float r1 = x - __bfloat162float(h);
          const __nv_bfloat16 mm = __float2bfloat16_rn(r1);
          const __nv_bfloat16 l = __float2bfloat16_rn(r1 - __bfloat162float(mm));
          srow[j] = h;
          srow[g.cs_seg + j] = mm;
          srow[2 * g.cs_seg + j] = l;
        }
      }
    }
  }
}

template <bool A_MN, bool B_MN, int TILE_N>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const KArgs g) {
  using C_ = Cfg<TILE_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C_::STAGES * C_::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C_::STAGES;
  uint64_t* acc_full = bars + 2 * C_::STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C_::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], NUM_EPI_WARPS);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C_::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nunits = g.ntiles * g.splitk;
  const int kb_per_split = (g.kblocks + g.splitk - 1) / g.splitk;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    // The whole warp walks the loop converged (uniform control flow); one elected lane issues.
    int stage = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      int m0, n0;
      decode_tile<TILE_N>(g, u / g.splitk, m0, n0);
      const int sp = u % g.splitk;
      const int kb0 = sp * kb_per_split;
      const int kb1 = min(g.kblocks, kb0 + kb_per_split);
      for (int p = 0; p < g.npairs; ++p) {
        const int sa = g.pair_a[p], sb = g.pair_b[p];
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sA = smem + stage * C_::STAGE_BYTES;
            uint8_t* sB = sA + A_BYTES;
            mbar_expect_tx(&full[stage], C_::STAGE_BYTES);
            if (A_MN) {
#pragma unroll
              for (int gi = 0; gi < TILE_M / 64; ++gi)
                tma_load_3d(sA + gi * GROUP_BYTES, &tmA, &full[stage], m0 + 64 * gi, kb * BLOCK_K,
                            sa);
            } else {
              tma_load_3d(sA, &tmA, &full[stage], kb * BLOCK_K, m0, sa);
            }
            if (B_MN) {
#pragma unroll
              for (int gi = 0; gi < TILE_N / 64; ++gi)
                tma_load_3d(sB + gi * GROUP_BYTES, &tmB, &full[stage], n0 + 64 * gi, kb * BLOCK_K,
                            sb);
            } else {
              tma_load_3d(sB, &tmB, &full[stage], kb * BLOCK_K, n0, sb);
            }
          }
          __syncwarp();
          if (++stage == C_::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    // Converged warp, one elected lane issues. Descriptors: the high word (LBO/SBO/version/
    // swizzle) is a kernel constant; only the 14-bit start-address field moves, by adding
    // (byte offset >> 4) to the low word (shared memory is < 256 KB, so no carry leaves the field).
    constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, TILE_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t adesc0 = A_MN ? umma_smem_desc_sw128(smem_base, g.lbo_mn, g.sbo_mn)
                                 : umma_smem_desc_sw128(smem_base, g.lbo_k, g.sbo_k);
    const uint64_t bdesc0 = B_MN ? umma_smem_desc_sw128(smem_base + A_BYTES, g.lbo_mn, g.sbo_mn)
                                 : umma_smem_desc_sw128(smem_base + A_BYTES, g.lbo_k, g.sbo_k);
    constexpr uint32_t KSTEP_A = (A_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
    constexpr uint32_t KSTEP_B = (B_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      const int sp = u % g.splitk;
      const int kb0 = sp * kb_per_split;
      const int kb1 = min(g.kblocks, kb0 + kb_per_split);
      const int niter = g.npairs * max(0, kb1 - kb0);
      for (int it0 = 0; it0 < niter; it0 += g.chunk_iters) {
        const int it1 = min(niter, it0 + g.chunk_iters);
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TILE_N;
        for (int it = it0; it < it1; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t soff = static_cast<uint32_t>(stage * C_::STAGE_BYTES) >> 4;
            const uint64_t ad = adesc0 + soff;
            const uint64_t bd = bdesc0 + soff;
            umma_bf16(d_tmem, ad, bd, idesc, it > it0 ? 1u : 0u);
#pragma unroll
            for (int ks = 1; ks < BLOCK_K / UMMA_K; ++ks)
              umma_bf16_acc(d_tmem, ad + ks * KSTEP_A, bd + ks * KSTEP_B, idesc);
            umma_commit(&empty[stage]);  // frees this smem stage once its MMAs retire
          }
          __syncwarp();
          if (++stage == C_::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&acc_full[acc]);  // chunk complete -> epilogue
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue (warps 2..9)
    constexpr int COLS = TILE_N / 2;           // columns owned by one epilogue warp
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;          // which column half of the tile
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      int m0, n0;
      decode_tile<TILE_N>(g, u / g.splitk, m0, n0);
      const int sp = u % g.splitk;
      const int kb0 = sp * kb_per_split;
      const int kb1 = min(g.kblocks, kb0 + kb_per_split);
      const int niter = g.npairs * max(0, kb1 - kb0);
      float sum[COLS];
#pragma unroll
      for (int j = 0; j < COLS; ++j) sum[j] = 0.f;
      for (int it0 = 0; it0 < niter; it0 += g.chunk_iters) {
        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + acc * TILE_N + half * COLS +
                               (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t_row + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[c * 32 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      write_out<COLS>(g, sum, m0 + q * 32 + lane, n0 + half * COLS, niter, vec_ok);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C_::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ CTA-pair kernel
// Same pipeline on a 2-CTA cluster (cta_group::2): the pair owns a 256 x 256 output tile, each CTA
// stages its own 128 A rows and HALF of the B tile (128 columns), and the leader CTA issues one
// 256 x 256 x 16 tcgen05.mma that reads both CTAs' shared memory and writes each CTA's 128 rows
// into that CTA's TMEM. Per SM and k-block that is 32 KB of operands instead of 48 KB: the TMA
// fill + tensor-core read traffic of the 128 x 256 single-CTA tile oversubscribes the 128 B/clk
// shared-memory port (measured 73 % tensor-pipe active with the issue thread never starved), the
// pair tile fits it, and the ring is 6 deep instead of 4.
constexpr int P_TILE = 256;
constexpr int P_HALF = 128;
constexpr int P_B_BYTES = P_HALF * BLOCK_K * 2;
constexpr int P_STAGE_BYTES = A_BYTES + P_B_BYTES;
constexpr int P_STAGES = 6;
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ void decode_tile2(const KArgs& g, int t, int& bi, int& bj) {
  // Tiles are enumerated in bands of g.band row blocks, column-major inside a band, so the CTA
  // pairs in flight share a few A slabs (the band) and a few B slabs: L2-resident reuse instead of
  // streaming one operand once per tile row (DRAM traffic is power the tensor cores cannot use).
  const int tiles_m = (g.M + P_TILE - 1) / P_TILE;
  const int band = g.band;
  bi = bj = 0;
  if (!g.lower) {
    const int per_band = band * g.tiles_n;
    const int b = t / per_band;
    const int lo = b * band;
    const int nrows = min(band, tiles_m - lo);
    t -= b * per_band;
    bj = t / nrows;
    bi = lo + t - bj * nrows;
    return;
  }
  // lower triangle: row block i owns column blocks j <= i
  for (int lo = 0; lo < tiles_m; lo += band) {
    const int hi = min(tiles_m, lo + band);
    const int nrows = hi - lo;
    const int cnt_full = lo * nrows;
    if (t < cnt_full) {
      bj = t / nrows;
      bi = lo + t - bj * nrows;
      return;
    }
    t -= cnt_full;
    for (int jj = lo; jj < hi; ++jj) {  // the staircase at the diagonal
      const int c = hi - jj;
      if (t < c) {
        bj = jj;
        bi = jj + t;
        return;
      }
      t -= c;
    }
  }
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const KArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_STAGES * P_STAGE_BYTES);
  uint64_t* full = bars;  // only the leader's are used: both CTAs' TMA bytes are credited there
  uint64_t* empty = bars + P_STAGES;
  uint64_t* acc_full = bars + 2 * P_STAGES;
  uint64_t* acc_empty = acc_full + 2;  // only the leader's: both CTAs' epilogue warps arrive there
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = warp_idx_uniform();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = static_cast<int>(cluster_id_x());
  const int npair = static_cast<int>(cluster_count_x());

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 2 * NUM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, 2 * P_TILE);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs exist before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nunits = g.ntiles * g.splitk;
  const int kb_per_split = (g.kblocks + g.splitk - 1) / g.splitk;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    for (int u = pair; u < nunits; u += npair) {
      int bi, bj;
      decode_tile2(g, u / g.splitk, bi, bj);
      const int m0 = bi * P_TILE + static_cast<int>(rank) * TILE_M;
      const int n0 = bj * P_TILE + static_cast<int>(rank) * P_HALF;
      const int sp = u % g.splitk;
      const int kb0 = sp * kb_per_split;
      const int kb1 = min(g.kblocks, kb0 + kb_per_split);
      for (int p = 0; p < g.npairs; ++p) {
        const int sa = g.pair_a[p], sb = g.pair_b[p];
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_relaxed(&empty[stage], phase ^ 1, g.sleep_ns);
          if (elect_one()) {
            uint8_t* sA = smem + stage * P_STAGE_BYTES;
            uint8_t* sB = sA + A_BYTES;
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * P_STAGE_BYTES);
            if (A_MN) {
#pragma unroll
              for (int gi = 0; gi < TILE_M / 64; ++gi)
                tma_load_3d_pair(sA + gi * GROUP_BYTES, &tmA, &full[stage], m0 + 64 * gi,
                                 kb * BLOCK_K, sa);
            } else {
              tma_load_3d_pair(sA, &tmA, &full[stage], kb * BLOCK_K, m0, sa);
            }
            if (B_MN) {
#pragma unroll
              for (int gi = 0; gi < P_HALF / 64; ++gi)
                tma_load_3d_pair(sB + gi * GROUP_BYTES, &tmB, &full[stage], n0 + 64 * gi,
                                 kb * BLOCK_K, sb);
            } else {
              tma_load_3d_pair(sB, &tmB, &full[stage], kb * BLOCK_K, n0, sb);
            }
          }
          __syncwarp();
          if (++stage == P_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (leader CTA only)
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * TILE_M, P_TILE, A_MN ? 1 : 0, B_MN ? 1 : 0);
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t adesc0 = A_MN ? umma_smem_desc_sw128(smem_base, g.lbo_mn, g.sbo_mn)
                                   : umma_smem_desc_sw128(smem_base, g.lbo_k, g.sbo_k);
      const uint64_t bdesc0 = B_MN ? umma_smem_desc_sw128(smem_base + A_BYTES, g.lbo_mn, g.sbo_mn)
                                   : umma_smem_desc_sw128(smem_base + A_BYTES, g.lbo_k, g.sbo_k);
      constexpr uint32_t KSTEP_A = (A_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
      constexpr uint32_t KSTEP_B = (B_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = pair; u < nunits; u += npair) {
        const int sp = u % g.splitk;
        const int kb0 = sp * kb_per_split;
        const int kb1 = min(g.kblocks, kb0 + kb_per_split);
        const int niter = g.npairs * max(0, kb1 - kb0);
        for (int it0 = 0; it0 < niter; it0 += g.chunk_iters) {
          const int it1 = min(niter, it0 + g.chunk_iters);
          mbar_wait(&acc_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * P_TILE;
          for (int it = it0; it < it1; ++it) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t soff = static_cast<uint32_t>(stage * P_STAGE_BYTES) >> 4;
              const uint64_t ad = adesc0 + soff;
              const uint64_t bd = bdesc0 + soff;
              umma_bf16_pair(d_tmem, ad, bd, idesc, it > it0 ? 1u : 0u);
#pragma unroll
              for (int ks = 1; ks < BLOCK_K / UMMA_K; ++ks)
                umma_bf16_pair_acc(d_tmem, ad + ks * KSTEP_A, bd + ks * KSTEP_B, idesc);
              umma_commit_pair(&empty[stage]);  // frees this stage in BOTH CTAs
            }
            __syncwarp();
            if (++stage == P_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (elect_one()) umma_commit_pair(&acc_full[acc]);  // chunk complete -> both epilogues
          __syncwarp();
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue (warps 2..9, both CTAs)
    constexpr int COLS = P_TILE / 2;
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
    for (int u = pair; u < nunits; u += npair) {
      int bi, bj;
      decode_tile2(g, u / g.splitk, bi, bj);
      const int m0 = bi * P_TILE + static_cast<int>(rank) * TILE_M;
      const int n0 = bj * P_TILE;
      const int sp = u % g.splitk;
      const int kb0 = sp * kb_per_split;
      const int kb1 = min(g.kblocks, kb0 + kb_per_split);
      const int niter = g.npairs * max(0, kb1 - kb0);
      float sum[COLS];
#pragma unroll
      for (int j = 0; j < COLS; ++j) sum[j] = 0.f;
      for (int it0 = 0; it0 < niter; it0 += g.chunk_iters) {
        mbar_wait_relaxed(&acc_full[acc], acc_phase, g.sleep_ns);
        tc_fence_after();
        const uint32_t t_row = tmem_base + acc * P_TILE + half * COLS +
                               (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t_row + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sum[c * 32 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0)
            mbar_arrive(&acc_empty[acc]);
          else
            mbar_arrive_cluster(&acc_empty[acc], 0);
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      write_out<COLS>(g, sum, m0 + q * 32 + lane, n0 + half * COLS, niter, vec_ok);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // nobody leaves while the peer can still signal or read this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 2 * P_TILE);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
          cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 3-D map {inner, rows, seg}; box {64, box_rows, 1}; 128B swizzle; zero fill out of bounds.
int make_map(CUtensorMap* map, const GemmOperand& op, long long inner, long long rows,
             int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return -38;  // ENOSYS
  if ((reinterpret_cast<uintptr_t>(op.ptr) & 15) || (op.ld & 7) || (op.nseg > 1 && (op.seg_stride & 7)))
    return -22;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(rows),
                        static_cast<cuuint64_t>(op.nseg)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(op.ld) * 2,
                           static_cast<cuuint64_t>(op.nseg > 1 ? op.seg_stride : op.ld * rows) * 2};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(op.ptr),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(1000 + r);
}

long long g_dbg[16] = {0};
long long g_info[16] = {0};

// Per-device caches: the SM count and "cudaFuncSetAttribute done" flags (a function attribute
// belongs to the device that was current when it was set).
int num_sms() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& slot = n[dev & 63];
  if (slot == 0) {
    cudaDeviceGetAttribute(&slot, cudaDevAttrMultiProcessorCount, dev);
    if (slot <= 0) slot = 148;
  }
  return slot;
}
bool& attr_flag(bool (&flags)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  return flags[dev & 63];
}

template <bool A_MN, bool B_MN, int TILE_N>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const KArgs& g, int grid,
           cudaStream_t stream) {
  using C_ = Cfg<TILE_N>;
  auto kern = gemm_tc_kernel<A_MN, B_MN, TILE_N>;
  static bool attr_flags[64] = {};
  bool& attr_done = attr_flag(attr_flags);
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES) !=
        cudaSuccess)
      return -12;
    attr_done = true;
  }
  kern<<<grid, NUM_THREADS, C_::SMEM_BYTES, stream>>>(ta, tb, g);
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

template <bool A_MN, bool B_MN>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const KArgs& g, int npairs_grid,
            cudaStream_t stream) {
  auto kern = gemm_tc2_kernel<A_MN, B_MN>;
  static bool attr_flags[64] = {};
  bool& attr_done = attr_flag(attr_flags);
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES) !=
        cudaSuccess)
      return -12;
    attr_done = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * npairs_grid, 1, 1);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = P_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, ta, tb, g) != cudaSuccess) return -5;
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace

int make_tma_2d_bf16(CUtensorMap* map, const void* ptr, long long inner, long long rows,
                     long long ld, int box_rows) {
  GemmOperand op{static_cast<const __nv_bfloat16*>(ptr), 0, ld, 1, 0};
  return make_map(map, op, inner, rows, box_rows);
}
int device_sm_count() { return num_sms(); }

void gemm_tc_debug_set(int key, long long value) {
  if (key >= 0 && key < 16) g_dbg[key] = value;
}
long long gemm_tc_last_launch_info(int key) { return (key >= 0 && key < 16) ? g_info[key] : 0; }

int gemm_tc(const GemmOperand& A, const GemmOperand& B, int M, int N, int K, int pair_order,
            const GemmEpilogue& ep, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return -22;
  if (ep.lower_only && M != N) return -22;
  if ((A.nseg != 1 && A.nseg != 3) || (B.nseg != 1 && B.nseg != 3)) return -22;

  // dbg[0]: force TILE_N (128/256); dbg[1]: force splitk; dbg[2..5]: lbo_mn, sbo_mn, lbo_k, sbo_k
  int tile_n = (N > 128) ? 256 : 128;
  if (g_dbg[0] == 128 || g_dbg[0] == 256) tile_n = static_cast<int>(g_dbg[0]);
  // dbg[7]: 1 = never use the CTA-pair kernel, 2 = use it whenever the tile shape allows
  const int sms = num_sms();
  bool pair = tile_n == 256 && M > 128 && sms >= 2;
  if (g_dbg[7] == 1) pair = false;

  KArgs g;
  memset(&g, 0, sizeof(g));
  g.M = M;
  g.N = N;
  g.kblocks = (K + BLOCK_K - 1) / BLOCK_K;
  if (pair_order < 0) pair_order = (A.nseg == 3 || B.nseg == 3) ? 2 : 0;
  g.npairs = 0;
  for (int s = 0; s <= pair_order; ++s)  // most significant products first
    for (int sa = 0; sa <= s; ++sa) {
      const int sb = s - sa;
      if (sa < A.nseg && sb < B.nseg && g.npairs < MAX_PAIRS) {
        g.pair_a[g.npairs] = sa;
        g.pair_b[g.npairs] = sb;
        ++g.npairs;
      }
    }
  const int tiles_m = pair ? (M + P_TILE - 1) / P_TILE : (M + TILE_M - 1) / TILE_M;
  g.tiles_n = (N + tile_n - 1) / tile_n;
  if (ep.lower_only) {
    long long nt = 0;
    for (int i = 0; i < tiles_m; ++i) nt += (pair || tile_n == 128) ? (i + 1) : (i / 2 + 1);
    g.ntiles = static_cast<int>(nt);
  } else {
    g.ntiles = tiles_m * g.tiles_n;
  }
  g.lower = ep.lower_only;
  g.accumulate = ep.accumulate;
  g.alpha = ep.alpha;
  g.C = ep.C;
  g.ldc = ep.ldc;
  g.Cb = ep.Cb;
  g.ldcb = ep.ldcb;
  g.Cs = ep.Cs;
  g.ldcs = ep.ldcs;
  g.cs_seg = ep.cs_seg;
  g.bias = ep.bias;
  g.chunk_iters = g_dbg[6] > 0 ? static_cast<int>(g_dbg[6]) : (g.npairs > 1 ? CHUNK_ITERS_SPLIT : CHUNK_ITERS);
  g.lbo_mn = g_dbg[2] ? static_cast<uint32_t>(g_dbg[2]) : GROUP_BYTES;  // 64-wide MN group stride
  g.sbo_mn = g_dbg[3] ? static_cast<uint32_t>(g_dbg[3]) : 1024;         // 8 k-rows x 128 B
  g.lbo_k = g_dbg[4] ? static_cast<uint32_t>(g_dbg[4]) : 16;            // unused for swizzled K-major
  g.sbo_k = g_dbg[5] ? static_cast<uint32_t>(g_dbg[5]) : 1024;          // 8 MN-rows x 128 B
  // dbg[8]: back-off ns of the long waits (default 0 for now), dbg[9]: raster band (tile rows)
  g.sleep_ns = g_dbg[8] > 0 ? static_cast<uint32_t>(g_dbg[8]) : 0;
  g.band = g_dbg[9] > 0 ? static_cast<int>(g_dbg[9]) : 8;

  // k-split: only when accumulating with atomics (a plain store cannot be split). Cost model per
  // unit: max(main loop, epilogue) with the epilogue (TILE_M x TILE_N red.add) worth ~E k-block
  // times; pick the split minimising waves x unit cost.
  const int workers = pair ? sms / 2 : sms;  // CTAs, or CTA pairs
  int splitk = 1;
  if (ep.accumulate && ep.Cb == nullptr && ep.Cs == nullptr && ep.bias == nullptr) {
    const double epi_kb = 40.0;  // measured: a 128x256 red.add epilogue ~ 20k cycles ~ 40 k-blocks
    double best = 1e300;
    const int max_split = std::max(1, std::min(g.kblocks / 4, 64));
    for (int s = 1; s <= max_split; ++s) {
      const long long units = static_cast<long long>(g.ntiles) * s;
      const long long waves = (units + workers - 1) / workers;
      const double kb_unit = static_cast<double>((g.kblocks + s - 1) / s) * g.npairs;
      const double cost = static_cast<double>(waves) * std::max(kb_unit, epi_kb) + epi_kb;
      if (cost < best * 0.97) {  // prefer fewer splits on near-ties
        best = cost;
        splitk = s;
      }
    }
  }
  if (g_dbg[1] > 0) splitk = static_cast<int>(std::min<long long>(g_dbg[1], g.kblocks));
  if (!ep.accumulate || ep.deterministic) splitk = 1;
  g.splitk = splitk;
  // make every split non-empty
  {
    const int per = (g.kblocks + g.splitk - 1) / g.splitk;
    g.splitk = (g.kblocks + per - 1) / per;
  }

  CUtensorMap ta, tb;
  int rc;
  // MN-major: inner = MN extent, rows = K. K-major: inner = K, rows = MN extent.
  rc = A.mn_major ? make_map(&ta, A, M, K, BLOCK_K) : make_map(&ta, A, K, M, TILE_M);
  if (rc) return rc;
  rc = B.mn_major ? make_map(&tb, B, N, K, BLOCK_K) : make_map(&tb, B, K, N, pair ? P_HALF : tile_n);
  if (rc) return rc;

  const long long units = static_cast<long long>(g.ntiles) * g.splitk;
  const int grid = static_cast<int>(std::min<long long>(units, workers));
  g_info[0] = tile_n;
  g_info[1] = g.splitk;
  g_info[2] = g.ntiles;
  g_info[3] = grid;
  g_info[4] = g.npairs;
  g_info[5] = pair ? 1 : 0;

  if (pair) {
    if (A.mn_major && B.mn_major) return launch2<true, true>(ta, tb, g, grid, stream);
    if (!A.mn_major && !B.mn_major) return launch2<false, false>(ta, tb, g, grid, stream);
    if (!A.mn_major && B.mn_major) return launch2<false, true>(ta, tb, g, grid, stream);
    return launch2<true, false>(ta, tb, g, grid, stream);
  }

#define PTD_LAUNCH(AM, BM)                                                      \
  (tile_n == 256 ? launch<AM, BM, 256>(ta, tb, g, grid, stream)                 \
                 : launch<AM, BM, 128>(ta, tb, g, grid, stream))
  if (A.mn_major && B.mn_major) return PTD_LAUNCH(true, true);
  if (!A.mn_major && !B.mn_major) return PTD_LAUNCH(false, false);
  if (!A.mn_major && B.mn_major) return PTD_LAUNCH(false, true);
  return PTD_LAUNCH(true, false);
#undef PTD_LAUNCH
}

}  // namespace ptd
