// Decomposed-layer forward y = W2 (W1 x) + b (reference: the nn.Sequential built at F:84-95 /
// D:74-85, K7 of SURVEY.md).
//
// Fused path (bf16, k <= 256, TMA-friendly pointers): ONE kernel per 128-token tile computes
// H = X W1^T into TMEM (tcgen05), drains it once to a 128B-swizzled bf16 tile in shared memory
// and feeds that tile straight back to the tensor core as the A operand of Y = H W2^T; the rank-k
// intermediate never leaves the SM. Warp-specialised: warp 0 TMA producer (one 3-stage ring
// carries first X/W1 k-blocks, then W2 tiles), warp 1 issues tcgen05.mma, warps 2-9 drain TMEM
// (H -> smem, Y -> +bias -> bf16 -> global) with double-buffered Y accumulators. When there are
// fewer token tiles than SMs the `out` dimension is split over CTAs of the same token tile
// (a cost model trades the recomputed first GEMM against idle SMs).
//
// Unfused path (fp32 models, k > 256): two passes of the tcgen05 GEMM engine with the [n, k]
// intermediate H staged in workspace (bf16, or bf16x3 split for fp32 models).
#include "lowrank.cuh"

#include <cuda_bf16.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "elementwise.cuh"
#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace ptd {

namespace {

// Runtime knobs (set through ptdeco_debug_set keys 200..206; the Python loader maps the
// PTDECO_B200_* environment variables onto them ONCE at load time -- nothing on the forward path
// reads the environment): 0 force the decode kernel, 1 no decode kernel, 2 no fused kernel,
// 3 no persistent kernel, 4 no TMA store, 5 tile rotation (-1 = default), 6 pipeline stages,
// 7 forced token rows per tile (multiple of 8 in [8, 128]; 0 = cost model).
long long g_knob[10] = {0, 0, 0, 0, 0, -1, 0, 0, 0, 0};  // [8]: 1 = no k-split clusters, 2 = always when possible

// cudaFuncSetAttribute is per device: remember which devices were configured.
bool* lr_attr_flag(int which) {
  static bool done[3][64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  return &done[which][dev & 63];
}
inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

struct Carve {
  uint8_t* base;
  size_t off = 0;
  void* take(size_t bytes) {
    off = (off + 255) & ~static_cast<size_t>(255);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

// bf16: use in place when TMA-friendly, else padded copy; fp32: bf16x3 split.
int stage(const void* src, int is_bf16, long long rows, int cols, long long ld, Carve& cv,
          GemmOperand* op, cudaStream_t st, bool dry) {
  op->mn_major = 0;
  if (is_bf16 && aligned16(src) && (ld % 8) == 0 && !dry) {
    op->ptr = static_cast<const __nv_bfloat16*>(src);
    op->ld = ld;
    op->nseg = 1;
    op->seg_stride = 0;
    return 0;
  }
  const int nseg = is_bf16 ? 1 : 3;
  const long long ldp = round_up(cols, 8);
  __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(cv.take(2ull * nseg * rows * ldp));
  if (dry) return 0;
  if (dst == nullptr) return -12;
  int rc = split_rows(src, is_bf16, ld, rows, cols, nullptr, 0.f, dst, ldp, nseg, rows * ldp, st);
  if (rc) return rc;
  op->ptr = dst;
  op->ld = ldp;
  op->nseg = nseg;
  op->seg_stride = rows * ldp;
  return 0;
}
// ------------------------------------------------------------------------------ fused kernel
constexpr int F_TILE_M = 128;
constexpr int F_BK = 64;
constexpr int F_TILE_N = 256;
constexpr int F_MAX_STAGES = 5;
constexpr int F_MAX_STAGES1 = 8;  // GEMM-1 ring of the narrow-rank plan (see below)
constexpr int F_XBYTES = F_TILE_M * F_BK * 2;    // X k-block: 128 rows x 128 B
constexpr int F_WBYTES = 256 * F_BK * 2;         // W2 tile: 256 rows x 128 B
constexpr int F_HBLOCK = F_TILE_M * 128;         // one 64-wide k-block of H: 16 KB
constexpr int F_EPI_WARPS = 16;  // 4 per TMEM lane quarter: one 64-column slice each
constexpr int F_THREADS = 64 + 32 * F_EPI_WARPS;
constexpr int F_STG_BYTES = F_EPI_WARPS * 4096;  // output staging, 4 KB per epilogue warp
// Shared-memory plans (the ring must keep ~2 us of TMA latency covered, so it is as deep as fits):
//   kp <= 128: 4 slots of [X 16 KB | W1 16 KB]; a 32 KB W2 tile fills a whole slot; H 32 KB;
//              separate 64 KB output staging (4 KB per epilogue warp)         -> 225 KB
//              GEMM 1 streams 2 MB per token tile through ONE SM; with 4 slots (128 KB in flight)
//              Little's law caps that at ~75 GB/s (measured: the phase took 27 of the kernel's
//              47 us at N = 8192). H and the output staging are idle until GEMM 1 is over, so
//              during GEMM 1 the ring extends over them: 7 slots (6 at kp = 64), its own barriers;
//              GEMM 2 restarts on the first 4 slots once GEMM 1 has committed.
//   kp  > 128: 3 slots of [X 16 KB | W 32 KB]; W2 tiles land in the W part; H 64 KB; the idle X
//              parts of the three slots plus 16 KB are the output staging     -> 224 KB

struct FusedArgs {
  int n, in_f, k, out_f, kp;
  int groups, tiles_per_group, out_tiles;
  __nv_bfloat16* Y;
  long long ldy;
  const float* bias;
  int tma_store;  // 1: Y tiles leave through shared memory + cp.async.bulk.tensor stores
  int stages, slot_bytes, w2_off, stg_separate, rotate;
  int stages1;  // slots of the GEMM-1 ring (== stages: one ring for both phases)
  // Token rows per tile (multiple of 8, <= 128). The tensor core always works on 128 rows; when
  // the op is HBM-bound (small k) the rows are what costs, so the host sizes the tiles to fill
  // whole waves of SMs (N = 32768: 293 tiles of 112 rows = 2 full waves instead of 1.73; N = 8192:
  // 147 tiles of 56 rows on 147 SMs instead of 64 tiles). X boxes / Y stores are clipped to
  // tile_m rows; the tile's remaining accumulator rows hold don't-care values (rows are
  // independent in both GEMMs) and are never stored.
  int tile_m;
  // 1: launched as 2-CTA clusters, the two CTAs being the out-groups 2c and 2c+1 of one token tile.
  // Without it each of them runs the WHOLE GEMM 1 of that tile (X tile and all of W1 streamed
  // twice, the per-SM TMA stream being what bounds small-N launches); with it each accumulates
  // H over half of `in`, the fp32 partials cross through distributed shared memory (64 KB each
  // way, into the output staging area that is idle until the first Y tile), and both add the two
  // halves in the same order -- so every CTA streams half of X's tile and half of W1, and issues
  // half of GEMM 1's MMAs. kp <= 128 only (the partial has to fit the staging area).
  int ksplit;
  int prof;  // debug: accumulate phase cycles into g_fused_prof
};

// (debug) cycles of the middle CTA's first epilogue warp: setup | GEMM 1 | H hand-over | first Y tile |
// remaining Y tiles | store drain; switched on by lowrank_debug_set(9, 1), read by lowrank_debug_get
__device__ unsigned long long g_fused_prof[8];

__global__ void __launch_bounds__(F_THREADS, 1)
lowrank_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmY,
                     const FusedArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int F_STAGES = g.stages, F_STAGE = g.slot_bytes;
  uint8_t* hbuf = smem + F_STAGES * F_STAGE;
  uint8_t* stg_base = hbuf + (g.kp / F_BK) * F_HBLOCK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + (g.stg_separate ? F_STG_BYTES : 4 * 4096));
  uint64_t* full = bars;
  uint64_t* empty = bars + F_MAX_STAGES;
  uint64_t* h_full = bars + 2 * F_MAX_STAGES;
  uint64_t* h_ready = h_full + 1;
  uint64_t* y_full = h_ready + 1;
  uint64_t* y_empty = y_full + 2;
  uint64_t* full1 = y_empty + 2;               // GEMM-1 ring when it is longer than the GEMM-2 ring
  uint64_t* empty1 = full1 + F_MAX_STAGES1;
  uint64_t* land_free = empty1 + F_MAX_STAGES1;  // k-split: the PEER's staging area may be written
  uint64_t* land_full = land_free + 1;           // k-split: the peer's partial H has landed here
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(land_full + 1);
  const int F_STAGES1 = g.stages1;
  const bool two_rings = F_STAGES1 != F_STAGES;
  uint64_t* f1 = two_rings ? full1 : full;
  uint64_t* e1 = two_rings ? empty1 : empty;

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const bool prof = g.prof && blockIdx.x == gridDim.x / 2 && threadIdx.x == 64;
  long long tp = prof ? clock64() : 0;
#define LR_PHASE(k)                                                              \
  if (prof) {                                                                    \
    const long long tn = clock64();                                              \
    atomicAdd(&g_fused_prof[k], static_cast<unsigned long long>(tn - tp));       \
    tp = tn;                                                                     \
  }
  const int rt = blockIdx.x / g.groups, og = blockIdx.x % g.groups;
  const int m0 = rt * g.tile_m;
  const int t0 = og * g.tiles_per_group;
  const int t1 = min(g.out_tiles, t0 + g.tiles_per_group);
  const int kb_all = (g.in_f + F_BK - 1) / F_BK;
  // k-split: groups is even, so blockIdx.x and og have the same parity = the rank in the cluster
  const uint32_t crank = g.ksplit ? static_cast<uint32_t>(og & 1) : 0u;
  const int kb_lo = g.ksplit ? static_cast<int>(crank) * ((kb_all + 1) / 2) : 0;
  const int kb1 = g.ksplit ? (crank == 0 ? (kb_all + 1) / 2 : kb_all - (kb_all + 1) / 2) : kb_all;
  const int kb2 = g.kp / F_BK;
  // Every CTA streams the SAME W1 k-blocks and W2 tiles; started in lockstep they all hit the
  // same L2 lines at the same time. Each CTA therefore starts at its own offset of the k loop
  // (a sum: order-free) and of its tile list.
  const int ntl = t1 - t0;
  const int rot1 = g.rotate ? static_cast<int>((blockIdx.x * 7u) % static_cast<unsigned>(kb1)) : 0;
  const int rot3 = (g.rotate && ntl > 0) ? static_cast<int>(rt % static_cast<unsigned>(ntl)) : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int s = 0; s < F_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    if (two_rings)
      for (int s = 0; s < F_STAGES1; ++s) {
        mbar_init(&full1[s], 1);
        mbar_init(&empty1[s], 1);
      }
    mbar_init(h_full, 1);
    mbar_init(h_ready, F_EPI_WARPS);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&y_full[a], 1);
      mbar_init(&y_empty[a], F_EPI_WARPS);
    }
    mbar_init(land_free, F_EPI_WARPS);
    mbar_init(land_full, F_EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (g.ksplit) cluster_sync_all();  // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t h_tmem = tmem_base + 256;  // H accumulates in the upper half; Y tile 0 uses the lower

  // Producer and MMA warps walk their loops converged (uniform control flow); one elected lane
  // issues, so TMA / tcgen05 instructions stay on the uniform datapath (see gemm_tc.cu).
  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int kk = 0; kk < kb1; ++kk) {  // GEMM 1 operands
      const int kb = kb_lo + (kk + rot1) % kb1;
      mbar_wait(&e1[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* sA = smem + stage * F_STAGE;
        mbar_expect_tx(&f1[stage], g.tile_m * 128 + g.kp * 128);
        tma_load_3d(sA, &tmX, &f1[stage], kb * F_BK, m0, 0);
        tma_load_3d(sA + F_XBYTES, &tmW1, &f1[stage], kb * F_BK, 0, 0);
      }
      __syncwarp();
      if (++stage == F_STAGES1) { stage = 0; phase ^= 1; }
    }
    if (two_rings) {  // the long ring overlaps H and the staging: GEMM 2 starts on fresh barriers
      mbar_wait(h_full, 0);  // every GEMM-1 MMA has read its slot
      stage = 0;
      phase = 0;
    }
    for (int tt = 0; tt < ntl; ++tt) {  // GEMM 2: W2 tiles
      const int t = t0 + (tt + rot3) % ntl;
      for (int kb = 0; kb < kb2; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sB = smem + stage * F_STAGE + g.w2_off;
          mbar_expect_tx(&full[stage], F_WBYTES);
          tma_load_3d(sB, &tmW2, &full[stage], kb * F_BK, t * F_TILE_N, 0);
        }
        __syncwarp();
        if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc1 = umma_idesc_bf16(F_TILE_M, g.kp, 0, 0);
    const uint32_t idesc2 = umma_idesc_bf16(F_TILE_M, F_TILE_N, 0, 0);
    const uint64_t desc0 = umma_smem_desc_sw128(0, 16, 1024);  // + (address >> 4) in the low word
    const uint32_t smem_base = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < kb1; ++kb) {
      mbar_wait(&f1[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sA = smem_base + stage * F_STAGE;
        const uint64_t ad = desc0 + (sA >> 4), bd = desc0 + ((sA + F_XBYTES) >> 4);
        umma_bf16(h_tmem, ad, bd, idesc1, kb > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 1; ks < F_BK / 16; ++ks) umma_bf16_acc(h_tmem, ad + 2 * ks, bd + 2 * ks, idesc1);
        umma_commit(&e1[stage]);
      }
      __syncwarp();
      if (++stage == F_STAGES1) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(h_full);
    __syncwarp();
    if (two_rings) {
      stage = 0;
      phase = 0;
    }
    mbar_wait(h_ready, 0);  // H is in shared memory as a swizzled bf16 K-major tile
    tc_fence_after();
    const uint32_t sH = smem_u32(hbuf);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = t0; t < t1; ++t) {
      mbar_wait(&y_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * F_TILE_N;
      for (int kb = 0; kb < kb2; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sB = smem_base + stage * F_STAGE + g.w2_off;
          const uint32_t sA = sH + kb * F_HBLOCK;
          const uint64_t ad = desc0 + (sA >> 4), bd = desc0 + (sB >> 4);
          umma_bf16(d_tmem, ad, bd, idesc2, kb > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 1; ks < F_BK / 16; ++ks) umma_bf16_acc(d_tmem, ad + 2 * ks, bd + 2 * ks, idesc2);
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == F_STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&y_full[acc]);
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quarter of this warp
    const int cg = (warp - 2) >> 2;         // column group 0..3
    const int row = q * 32 + lane;
    const uint32_t lane_bits = static_cast<uint32_t>(q * 32) << 16;
    // ---- H: TMEM -> bf16 -> swizzled shared memory (the layout TMA would have produced)
    LR_PHASE(0)
    mbar_wait(h_full, 0);
    tc_fence_after();
    LR_PHASE(1)
    if (g.ksplit) {
      // kp <= 128: at most one 32 x 32 chunk per warp. Partial sums cross as fp32, stored
      // [column][row] so that a warp's 32 rows of one column are one 128-byte line.
      const uint32_t peer = crank ^ 1u;
      const int col0 = cg * 32;
      const bool has = col0 < g.kp;  // warp-uniform
      uint32_t r[32];
      if (has) {
        tmem_ld_32x32(h_tmem + col0 + lane_bits, r);
        tmem_ld_wait();
      }
      // my GEMM 1 no longer reads its ring (h_full), which overlays my staging area: the peer may write it
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(land_free, peer);
      mbar_wait_cluster(land_free, 0);
      LR_PHASE(6)
      // landing zone: [column][row] fp32, so the 32 rows a warp holds of one column are one 128-byte
      // line (measured: 4-byte remote stores in this layout take 4.3k cycles for the 64 KB, 16-byte
      // vector stores into a row-major, XOR-swizzled zone 7.4k)
      float* zone = reinterpret_cast<float*>(stg_base);
      if (has) {
        const uint32_t dst = mapa_u32(smem_u32(zone + col0 * F_TILE_M + row), peer);
#pragma unroll
        for (int j = 0; j < 32; ++j) st_cluster_f32(dst + j * F_TILE_M * 4, __uint_as_float(r[j]));
      }
      fence_acq_rel_cluster();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(land_full, peer);
      LR_PHASE(7)
      mbar_wait_cluster(land_full, 0);
      fence_acq_rel_cluster();
      if (has) {
        // both CTAs add (rank 0's half) + (rank 1's half) in that order: identical H in both
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float other = zone[(col0 + j) * F_TILE_M + row];
          const float mine = __uint_as_float(r[j]);
          r[j] = __float_as_uint(crank == 0 ? mine + other : other + mine);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __align__(16) __nv_bfloat16 o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = __float2bfloat16_rn(__uint_as_float(r[8 * j + e]));
          const int col = col0 + 8 * j;
          const int chunk = (col & 63) >> 3;
          uint8_t* dst = hbuf + (col >> 6) * F_HBLOCK + row * 128 + ((chunk ^ (row & 7)) << 4);
          *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(o);
        }
      }
    } else
    for (int col0 = cg * 32; col0 < g.kp; col0 += 4 * 32) {
      uint32_t r[32];
      tmem_ld_32x32(h_tmem + col0 + lane_bits, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __align__(16) __nv_bfloat16 o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = __float2bfloat16_rn(__uint_as_float(r[8 * j + e]));
        const int col = col0 + 8 * j;
        const int chunk = (col & 63) >> 3;
        uint8_t* dst = hbuf + (col >> 6) * F_HBLOCK + row * 128 + ((chunk ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(o);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(h_ready);
    LR_PHASE(2)
    // ---- Y tiles: this warp owns rows [32q, 32q+32) x columns [64cg, 64cg+64) of every tile
    int acc = 0;
    uint32_t acc_phase = 0;
    const int m = m0 + row;
    const bool yvec = ((g.ldy & 7) == 0) && ((reinterpret_cast<uintptr_t>(g.Y) & 15) == 0);
    const bool bias_vec = (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0;
    // 32 rows x 128 B, swizzled like a TMA box. Wide-rank plan: the X parts of the ring slots are
    // idle during GEMM 2 (the producer only refills the W parts) and host 12 of the 16 slices.
    const int ew = warp - 2;
    uint8_t* stg = g.stg_separate ? stg_base + ew * 4096
                   : (ew < 12 ? smem + (ew >> 2) * F_STAGE + (ew & 3) * 4096 : stg_base + (ew - 12) * 4096);
    for (int tt = 0; tt < ntl; ++tt) {
      const int t = t0 + (tt + rot3) % ntl;
      mbar_wait(&y_full[acc], acc_phase);
      tc_fence_after();
      if (tt == 0) { LR_PHASE(3) }
      const int n0 = t * F_TILE_N + cg * 64;
      uint32_t ra[32], rb[32];
      const uint32_t taddr = tmem_base + acc * F_TILE_N + cg * 64 + lane_bits;
      tmem_ld_32x32(taddr, ra);
      tmem_ld_32x32(taddr + 32, rb);
      tmem_ld_wait();
      // the accumulator is in registers: hand the TMEM buffer back before the stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&y_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (g.bias != nullptr) {  // same 64 columns for every lane: broadcast 16-byte loads (L1 hits)
        if (bias_vec && n0 + 64 <= g.out_f) {
          const float4* b4 = reinterpret_cast<const float4*>(g.bias + n0);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 lo = __ldg(b4 + e), hi = __ldg(b4 + 8 + e);
            ra[4 * e + 0] = __float_as_uint(__uint_as_float(ra[4 * e + 0]) + lo.x);
            ra[4 * e + 1] = __float_as_uint(__uint_as_float(ra[4 * e + 1]) + lo.y);
            ra[4 * e + 2] = __float_as_uint(__uint_as_float(ra[4 * e + 2]) + lo.z);
            ra[4 * e + 3] = __float_as_uint(__uint_as_float(ra[4 * e + 3]) + lo.w);
            rb[4 * e + 0] = __float_as_uint(__uint_as_float(rb[4 * e + 0]) + hi.x);
            rb[4 * e + 1] = __float_as_uint(__uint_as_float(rb[4 * e + 1]) + hi.y);
            rb[4 * e + 2] = __float_as_uint(__uint_as_float(rb[4 * e + 2]) + hi.z);
            rb[4 * e + 3] = __float_as_uint(__uint_as_float(rb[4 * e + 3]) + hi.w);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (n0 + e < g.out_f) ra[e] = __float_as_uint(__uint_as_float(ra[e]) + __ldg(g.bias + n0 + e));
            if (n0 + 32 + e < g.out_f)
              rb[e] = __float_as_uint(__uint_as_float(rb[e]) + __ldg(g.bias + n0 + 32 + e));
          }
        }
      }
      if (g.tma_store) {
        if (lane == 0) bulk_wait_read_all();  // the previous tile's store has drained the slice
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __align__(16) __nv_bfloat16 o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            o[e] = __float2bfloat16_rn(__uint_as_float(j < 4 ? ra[8 * j + e] : rb[8 * (j - 4) + e]));
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
              *reinterpret_cast<const uint4*>(o);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && n0 < g.out_f) {  // 8-row boxes: only rows of THIS tile leave
#pragma unroll
          for (int r8 = 0; r8 < 4; ++r8) {
            const int rl = q * 32 + r8 * 8;
            if (rl < g.tile_m && m0 + rl < g.n) tma_store_3d(&tmY, stg + r8 * 1024, n0, m0 + rl, 0);
          }
          bulk_commit();
        }
      } else if (m < g.n && row < g.tile_m && n0 < g.out_f) {
        __nv_bfloat16* yrow = g.Y + static_cast<long long>(m) * g.ldy + n0;
        if (yvec && n0 + 64 <= g.out_f) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __align__(16) __nv_bfloat16 o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              o[e] = __float2bfloat16_rn(__uint_as_float(j < 4 ? ra[8 * j + e] : rb[8 * (j - 4) + e]));
            *reinterpret_cast<uint4*>(yrow + 8 * j) = *reinterpret_cast<const uint4*>(o);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 64; ++e)
            if (n0 + e < g.out_f)
              yrow[e] = __float2bfloat16_rn(__uint_as_float(e < 32 ? ra[e] : rb[e - 32]));
        }
      }
    }
    LR_PHASE(4)
    if (g.tma_store && lane == 0) bulk_wait_read_all();  // staging must outlive the last store
    LR_PHASE(5)
  }
#undef LR_PHASE

  tc_fence_before();
  __syncthreads();
  if (g.ksplit) cluster_sync_all();  // nobody leaves while the peer could still address this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------ persistent kernel
// Large-N variant (token tiles >= SMs, k <= 128). One CTA per SM walks over its token tiles and
// OVERLAPS the two phases of consecutive tiles: while the epilogue warps stream the Y tiles of
// token tile j to HBM (store-bound), the producer and the tensor core already accumulate
// H = X W1^T of token tile j+1 (load-bound), so HBM reads and writes are in flight together
// instead of all SMs reading, then all SMs writing. Producer and MMA issuer follow the same
// static op schedule:  K1(0,*) | for j: lead K1(j+1,*) , { T(j,t) , a share of K1(j+1,*) }_t
//   TMEM   H accumulators double-buffered [0,128) [128,256); Y accumulator [256,512)
//   smem   4 slots x 32 KB ([X 16|W1 16] or one 32 KB W2 k-block) | Hs 32 KB | staging 64 KB
constexpr int P_STAGES = 4;
constexpr int P_SLOT = 32768;
constexpr int P_LEAD = 8;
constexpr int P_SMEM = P_STAGES * P_SLOT + 2 * F_HBLOCK + F_STG_BYTES + 1024 + 256;

__global__ void __launch_bounds__(F_THREADS, 1)
lowrank_persistent_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                          const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmY,
                          const FusedArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* hbuf = smem + P_STAGES * P_SLOT;
  uint8_t* stg_base = hbuf + 2 * F_HBLOCK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + F_STG_BYTES);
  uint64_t* full = bars;                 // [P_STAGES]
  uint64_t* empty = bars + P_STAGES;     // [P_STAGES]
  uint64_t* h_full = bars + 2 * P_STAGES;  // [2] GEMM 1 of a token tile complete (per H buffer)
  uint64_t* h_ready = h_full + 2;        // Hs holds the bf16 tile
  uint64_t* hs_free = h_ready + 1;       // GEMM 2 of the token tile finished reading Hs
  uint64_t* y_full = hs_free + 1;
  uint64_t* y_empty = y_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_empty + 1);

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int row_tiles = (g.n + g.tile_m - 1) / g.tile_m;
  const int J = (row_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                static_cast<int>(gridDim.x);  // token tiles of this CTA: blockIdx.x + j * gridDim.x
  const int kb1 = (g.in_f + F_BK - 1) / F_BK;
  const int kb2 = g.kp / F_BK;
  const int ntiles = g.out_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&h_full[0], 1);
    mbar_init(&h_full[1], 1);
    mbar_init(h_ready, F_EPI_WARPS);
    mbar_init(hs_free, 1);
    mbar_init(y_full, 1);
    mbar_init(y_empty, F_EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t y_tmem = tmem_base + 256;

  if (warp == 0) {
    // converged warp, one elected lane issues (see gemm_tc.cu)
    int stage = 0;
    uint32_t phase = 0;
    auto load_k1 = [&](int j, int kb) {
      const int m0 = (static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x)) * g.tile_m;
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* slot = smem + stage * P_SLOT;
        mbar_expect_tx(&full[stage], g.tile_m * 128 + g.kp * 128);
        tma_load_3d(slot, &tmX, &full[stage], kb * F_BK, m0, 0);
        tma_load_3d(slot + F_XBYTES, &tmW1, &full[stage], kb * F_BK, 0, 0);
      }
      __syncwarp();
      if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
    };
    auto load_tile = [&](int t) {
      for (int kb = 0; kb < kb2; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[stage], F_WBYTES);
          tma_load_3d(smem + stage * P_SLOT, &tmW2, &full[stage], kb * F_BK, t * F_TILE_N, 0);
        }
        __syncwarp();
        if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
      }
    };
    for (int kb = 0; kb < kb1; ++kb) load_k1(0, kb);
    for (int j = 0; j < J; ++j) {
      const bool has_next = j + 1 < J;
      int kk = 0;
      if (has_next)
        for (; kk < min(kb1, P_LEAD); ++kk) load_k1(j + 1, kk);
      for (int t = 0; t < ntiles; ++t) {
        load_tile(t);
        if (has_next) {
          const int quota = (kb1 - kk + (ntiles - t) - 1) / (ntiles - t);
          for (int q = 0; q < quota; ++q) load_k1(j + 1, kk++);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc1 = umma_idesc_bf16(F_TILE_M, g.kp, 0, 0);
    const uint32_t idesc2 = umma_idesc_bf16(F_TILE_M, F_TILE_N, 0, 0);
    const uint64_t desc0 = umma_smem_desc_sw128(0, 16, 1024);  // + (address >> 4) in the low word
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t sH = smem_u32(hbuf);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t ytile = 0;  // running count of Y tiles (parity of y_full / y_empty)
    auto mma_k1 = [&](int j, int kb) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sA = smem_base + stage * P_SLOT;
        const uint64_t ad = desc0 + (sA >> 4), bd = desc0 + ((sA + F_XBYTES) >> 4);
        const uint32_t d_tmem = tmem_base + (j & 1) * 128;
        umma_bf16(d_tmem, ad, bd, idesc1, kb > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 1; ks < F_BK / 16; ++ks) umma_bf16_acc(d_tmem, ad + 2 * ks, bd + 2 * ks, idesc1);
        umma_commit(&empty[stage]);
        if (kb == kb1 - 1) umma_commit(&h_full[j & 1]);
      }
      __syncwarp();
      if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
    };
    auto mma_tile = [&](int j, int t) {
      if (t == 0) {
        mbar_wait(h_ready, j & 1);  // Hs holds token tile j
        tc_fence_after();
      }
      mbar_wait(y_empty, (ytile & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < kb2; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sB = smem_base + stage * P_SLOT;
          const uint32_t sA = sH + kb * F_HBLOCK;
          const uint64_t ad = desc0 + (sA >> 4), bd = desc0 + (sB >> 4);
          umma_bf16(y_tmem, ad, bd, idesc2, kb > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 1; ks < F_BK / 16; ++ks) umma_bf16_acc(y_tmem, ad + 2 * ks, bd + 2 * ks, idesc2);
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) {
        umma_commit(y_full);
        if (t == ntiles - 1) umma_commit(hs_free);
      }
      __syncwarp();
      ++ytile;
    };
    for (int kb = 0; kb < kb1; ++kb) mma_k1(0, kb);
    for (int j = 0; j < J; ++j) {
      const bool has_next = j + 1 < J;
      int kk = 0;
      if (has_next)
        for (; kk < min(kb1, P_LEAD); ++kk) mma_k1(j + 1, kk);
      for (int t = 0; t < ntiles; ++t) {
        mma_tile(j, t);
        if (has_next) {
          const int quota = (kb1 - kk + (ntiles - t) - 1) / (ntiles - t);
          for (int q = 0; q < quota; ++q) mma_k1(j + 1, kk++);
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_bits = static_cast<uint32_t>(q * 32) << 16;
    const bool bias_vec = (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0;
    uint8_t* stg = stg_base + (warp - 2) * 4096;
    uint32_t ytile = 0;
    for (int j = 0; j < J; ++j) {
      const int m0 = (static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x)) * g.tile_m;
      // ---- drain H of token tile j into Hs (bf16, 128B-swizzled K-major tile)
      mbar_wait(&h_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      if (j > 0) mbar_wait(hs_free, (j - 1) & 1);  // GEMM 2 of tile j-1 no longer reads Hs
      for (int col0 = cg * 32; col0 < g.kp; col0 += 4 * 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (j & 1) * 128 + col0 + lane_bits, r);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          __align__(16) __nv_bfloat16 o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = __float2bfloat16_rn(__uint_as_float(r[8 * jj + e]));
          const int col = col0 + 8 * jj;
          const int chunk = (col & 63) >> 3;
          uint8_t* dst = hbuf + (col >> 6) * F_HBLOCK + row * 128 + ((chunk ^ (row & 7)) << 4);
          *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(o);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(h_ready);
      // ---- Y tiles of token tile j
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(y_full, ytile & 1);
        tc_fence_after();
        ++ytile;
        const int n0 = t * F_TILE_N + cg * 64;
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(y_tmem + cg * 64 + lane_bits, ra);
        tmem_ld_32x32(y_tmem + cg * 64 + 32 + lane_bits, rb);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(y_empty);
        if (g.bias != nullptr) {
          if (bias_vec && n0 + 64 <= g.out_f) {
            const float4* b4 = reinterpret_cast<const float4*>(g.bias + n0);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 lo = __ldg(b4 + e), hi = __ldg(b4 + 8 + e);
              ra[4 * e + 0] = __float_as_uint(__uint_as_float(ra[4 * e + 0]) + lo.x);
              ra[4 * e + 1] = __float_as_uint(__uint_as_float(ra[4 * e + 1]) + lo.y);
              ra[4 * e + 2] = __float_as_uint(__uint_as_float(ra[4 * e + 2]) + lo.z);
              ra[4 * e + 3] = __float_as_uint(__uint_as_float(ra[4 * e + 3]) + lo.w);
              rb[4 * e + 0] = __float_as_uint(__uint_as_float(rb[4 * e + 0]) + hi.x);
              rb[4 * e + 1] = __float_as_uint(__uint_as_float(rb[4 * e + 1]) + hi.y);
              rb[4 * e + 2] = __float_as_uint(__uint_as_float(rb[4 * e + 2]) + hi.z);
              rb[4 * e + 3] = __float_as_uint(__uint_as_float(rb[4 * e + 3]) + hi.w);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              if (n0 + e < g.out_f) ra[e] = __float_as_uint(__uint_as_float(ra[e]) + __ldg(g.bias + n0 + e));
              if (n0 + 32 + e < g.out_f)
                rb[e] = __float_as_uint(__uint_as_float(rb[e]) + __ldg(g.bias + n0 + 32 + e));
            }
          }
        }
        if (lane == 0) bulk_wait_read_all();
        __syncwarp();
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          __align__(16) __nv_bfloat16 o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            o[e] = __float2bfloat16_rn(__uint_as_float(jj < 4 ? ra[8 * jj + e] : rb[8 * (jj - 4) + e]));
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((jj ^ (lane & 7)) << 4)) =
              *reinterpret_cast<const uint4*>(o);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && n0 < g.out_f) {  // 8-row boxes: only rows of THIS tile leave
#pragma unroll
          for (int r8 = 0; r8 < 4; ++r8) {
            const int rl = q * 32 + r8 * 8;
            if (rl < g.tile_m && m0 + rl < g.n) tma_store_3d(&tmY, stg + r8 * 1024, n0, m0 + rl, 0);
          }
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait_read_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------ decode kernel
// Small token counts (N <= 128: decode / short prefill). The op is pure weight streaming
// (2 k (in + out) bytes), so the weights take the M side of the tensor core ("swap AB") and the
// tokens the N side (padded to a multiple of 16 instead of to a 128-row tile):
//   phase 1   H^T[k, N] = W1[k, in] X^T : units = (128-row tiles of W1) x (splits of `in`), so every
//             SM streams a slice of W1; partial sums meet in an fp32 H[N, k] with red.add; the
//             last unit of a row tile (ticket counter) rounds its 128 columns of H to bf16
//   grid barrier (cooperative launch)
//   phase 2   Y^T[out, N] = W2[out, k] H^T : units = 128-row tiles of W2, H k-blocks ride the same
//             TMA ring as the W2 k-blocks; bias + bf16 store, coalesced along `out`
// One launch, one grid barrier, one memset of the (small) H / ticket workspace before it.
constexpr int D_TILE = 128;
constexpr int D_BK = 64;
constexpr int D_WBYTES = D_TILE * D_BK * 2;  // 16 KB weight k-block
constexpr int D_STAGES = 6;
constexpr int D_THREADS = 64 + 128;  // producer, MMA, 4 epilogue warps (one per TMEM lane quarter)
constexpr int D_CTRL_BYTES = 4096;   // [0] grid barrier, [16 + t] ticket of row tile t

struct DecodeArgs {
  int n, npad, in_f, k, out_f;
  int kb_in, kb_k;         // k-blocks of the two reductions
  int rt1, rt2;            // 128-row tiles of W1 / W2
  int s1, kb_per_split;    // phase 1 split of `in`
  long long ldh;           // row pitch of H (fp32) and Hb (bf16), elements (multiple of 128)
  unsigned* ctrl;
  float* H;
  __nv_bfloat16* Hb;
  __nv_bfloat16* Y;
  long long ldy;
  const float* bias;
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(D_THREADS, 1)
lowrank_decode_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmX,
                      const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmH,
                      const DecodeArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int stage_bytes = D_WBYTES + g.npad * 128;  // weight k-block + token-side k-block
  // stages are placed 32 KB apart so every operand stays 1024-byte aligned
  constexpr int D_SLOT = 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + D_STAGES * D_SLOT);
  uint64_t* full = bars;
  uint64_t* empty = bars + D_STAGES;
  uint64_t* acc_full = bars + 2 * D_STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  int* last_flag = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int G = static_cast<int>(gridDim.x), cta = static_cast<int>(blockIdx.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmH);
    for (int s = 0; s < D_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);  // two accumulators of up to 128 columns
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int units1 = g.rt1 * g.s1;
  // pipeline state carried across the two phases by each role
  int stage = 0;
  uint32_t phase = 0;
  int acc = 0;
  uint32_t acc_phase = 0;

  // ================================================================== phase 1
  if (warp == 0) {
    for (int u = cta; u < units1; u += G) {
      const int rt = u / g.s1, sp = u % g.s1;
      const int kb0 = sp * g.kb_per_split, kb1 = min(g.kb_in, kb0 + g.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sW = smem + stage * D_SLOT;
          mbar_expect_tx(&full[stage], stage_bytes);
          tma_load_3d(sW, &tmW1, &full[stage], kb * D_BK, rt * D_TILE, 0);
          tma_load_3d(sW + D_WBYTES, &tmX, &full[stage], kb * D_BK, 0, 0);
        }
        __syncwarp();
        if (++stage == D_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(D_TILE, g.npad, 0, 0);
    const uint64_t desc0 = umma_smem_desc_sw128(0, 16, 1024);
    const uint32_t smem_base = smem_u32(smem);
    for (int u = cta; u < units1; u += G) {
      const int sp = u % g.s1;
      const int kb0 = sp * g.kb_per_split, kb1 = min(g.kb_in, kb0 + g.kb_per_split);
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 128;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sW = smem_base + stage * D_SLOT;
          const uint64_t ad = desc0 + (sW >> 4), bd = desc0 + ((sW + D_WBYTES) >> 4);
          umma_bf16(d_tmem, ad, bd, idesc, kb > kb0 ? 1u : 0u);
#pragma unroll
          for (int ks = 1; ks < D_BK / 16; ++ks) umma_bf16_acc(d_tmem, ad + 2 * ks, bd + 2 * ks, idesc);
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == D_STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&acc_full[acc]);
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter (warps 2..5 -> 2,3,0,1)
    const int row = q * 32 + lane;
    const uint32_t lane_bits = static_cast<uint32_t>(q * 32) << 16;
    for (int u = cta; u < units1; u += G) {
      const int rt = u / g.s1;
      const int kcol = rt * D_TILE + row;  // column of H this thread owns
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      for (int c0 = 0; c0 < g.npad; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x16(tmem_base + acc * 128 + c0 + lane_bits, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)  // lanes of a warp hit 32 consecutive floats of row (c0 + j)
          atomicAdd(g.H + static_cast<long long>(c0 + j) * g.ldh + kcol, __uint_as_float(r[j]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      // ticket: the unit that completes a row tile rounds its 128 columns of H to bf16
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        const unsigned old = atomicAdd(g.ctrl + 16 + rt, 1u);
        *last_flag = (old == static_cast<unsigned>(g.s1 - 1)) ? 1 : 0;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (*last_flag) {
        __threadfence();
        if ((g.npad & 63) == 0) {  // 64 independent loads in flight, then the stores
          for (int n0 = 0; n0 < g.npad; n0 += 64) {
            float v[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = __ldcg(g.H + static_cast<long long>(n0 + j) * g.ldh + kcol);
#pragma unroll
            for (int j = 0; j < 64; ++j)
              g.Hb[static_cast<long long>(n0 + j) * g.ldh + kcol] = __float2bfloat16_rn(v[j]);
          }
        } else if ((g.npad & 31) == 0) {  // 32 independent loads in flight, then the stores
          for (int n0 = 0; n0 < g.npad; n0 += 32) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __ldcg(g.H + static_cast<long long>(n0 + j) * g.ldh + kcol);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              g.Hb[static_cast<long long>(n0 + j) * g.ldh + kcol] = __float2bfloat16_rn(v[j]);
          }
        } else {
          for (int n0 = 0; n0 < g.npad; n0 += 16) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __ldcg(g.H + static_cast<long long>(n0 + j) * g.ldh + kcol);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              g.Hb[static_cast<long long>(n0 + j) * g.ldh + kcol] = __float2bfloat16_rn(v[j]);
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // last_flag is rewritten by the next unit
    }
  }

  // ================================================================== grid barrier
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");  // Hb (generic stores) -> TMA loads of phase 2
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(g.ctrl, 1u);
    unsigned v = 0;
    unsigned polls = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(g.ctrl) : "memory");
      if (++polls > (1u << 28)) __trap();  // never hang the box on a lost CTA
    } while (v < static_cast<unsigned>(G));
    __threadfence();
  }
  __syncthreads();
  asm volatile("fence.proxy.async;" ::: "memory");

  // ================================================================== phase 2
  if (warp == 0) {
    for (int t = cta; t < g.rt2; t += G) {
      for (int kb = 0; kb < g.kb_k; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sW = smem + stage * D_SLOT;
          mbar_expect_tx(&full[stage], stage_bytes);
          tma_load_3d(sW, &tmW2, &full[stage], kb * D_BK, t * D_TILE, 0);
          tma_load_3d(sW + D_WBYTES, &tmH, &full[stage], kb * D_BK, 0, 0);
        }
        __syncwarp();
        if (++stage == D_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(D_TILE, g.npad, 0, 0);
    const uint64_t desc0 = umma_smem_desc_sw128(0, 16, 1024);
    const uint32_t smem_base = smem_u32(smem);
    for (int t = cta; t < g.rt2; t += G) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 128;
      for (int kb = 0; kb < g.kb_k; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sW = smem_base + stage * D_SLOT;
          const uint64_t ad = desc0 + (sW >> 4), bd = desc0 + ((sW + D_WBYTES) >> 4);
          umma_bf16(d_tmem, ad, bd, idesc, kb > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 1; ks < D_BK / 16; ++ks) umma_bf16_acc(d_tmem, ad + 2 * ks, bd + 2 * ks, idesc);
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == D_STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&acc_full[acc]);
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_bits = static_cast<uint32_t>(q * 32) << 16;
    for (int t = cta; t < g.rt2; t += G) {
      const int orow = t * D_TILE + row;  // output feature this thread owns
      const float b = (g.bias != nullptr && orow < g.out_f) ? __ldg(g.bias + orow) : 0.f;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      for (int c0 = 0; c0 < g.npad; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x16(tmem_base + acc * 128 + c0 + lane_bits, r);
        tmem_ld_wait();
        if (orow < g.out_f) {
#pragma unroll
          for (int j = 0; j < 16; ++j)  // a warp writes 32 consecutive bf16 of token row (c0 + j)
            if (c0 + j < g.n)
              g.Y[static_cast<long long>(c0 + j) * g.ldy + orow] =
                  __float2bfloat16_rn(__uint_as_float(r[j]) + b);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

size_t decode_workspace_bytes(long long n, int k) {
  const long long npad = round_up(n, 16);
  const long long ldh = round_up(k, 128);
  return D_CTRL_BYTES + static_cast<size_t>(npad * ldh) * 6 + 1024;
}

// Weight streaming only pays once there are weights to stream: below ~2 MB of factors the fused /
// two-launch paths (no grid barrier, no workspace memset) are as fast or faster (measured).
bool decode_eligible(const void* X, long long ldx, const void* W1, long long ldw1, const void* W2,
                     long long ldw2, int is_bf16, long long n, int in_f, int k, int out_f) {
  const bool big = static_cast<long long>(k) * (in_f + out_f) >= (1ll << 20) ||
                   g_knob[0] != 0;
  return is_bf16 && big && n <= 128 && k >= 16 && aligned16(X) && aligned16(W1) && aligned16(W2) &&
         (ldx % 8) == 0 && (ldw1 % 8) == 0 && (ldw2 % 8) == 0 &&
         (round_up(k, 128) / 128) <= (D_CTRL_BYTES / 4 - 16);
}

int launch_decode(const void* X, long long ldx, const void* W1, long long ldw1, const void* W2,
                  long long ldw2, const float* bias, void* Y, long long ldy, long long n, int in_f,
                  int k, int out_f, void* ws, cudaStream_t st) {
  DecodeArgs g;
  memset(&g, 0, sizeof(g));
  g.n = static_cast<int>(n);
  g.npad = static_cast<int>(round_up(n, 16));
  g.in_f = in_f;
  g.k = k;
  g.out_f = out_f;
  g.kb_in = (in_f + D_BK - 1) / D_BK;
  g.kb_k = static_cast<int>(round_up(k, D_BK)) / D_BK;
  g.rt1 = (k + D_TILE - 1) / D_TILE;
  g.rt2 = (out_f + D_TILE - 1) / D_TILE;
  g.ldh = round_up(k, 128);
  const int sms = device_sm_count();
  int s1 = std::max(1, std::min(g.kb_in, sms / g.rt1));
  g.kb_per_split = (g.kb_in + s1 - 1) / s1;
  g.s1 = (g.kb_in + g.kb_per_split - 1) / g.kb_per_split;
  uint8_t* base = static_cast<uint8_t*>(ws);
  base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(base) + 255) & ~uintptr_t(255));
  g.ctrl = reinterpret_cast<unsigned*>(base);
  g.H = reinterpret_cast<float*>(base + D_CTRL_BYTES);
  g.Hb = reinterpret_cast<__nv_bfloat16*>(base + D_CTRL_BYTES + static_cast<size_t>(g.npad) * g.ldh * 4);
  g.Y = static_cast<__nv_bfloat16*>(Y);
  g.ldy = ldy;
  g.bias = bias;
  // zero the barrier / tickets and the fp32 H (Hb is fully rewritten by the kernel)
  if (cudaMemsetAsync(base, 0, D_CTRL_BYTES + static_cast<size_t>(g.npad) * g.ldh * 4, st) != cudaSuccess)
    return -5;
  CUtensorMap tw1, tx, tw2, th;
  int rc;
  if ((rc = make_tma_2d_bf16(&tw1, W1, in_f, k, ldw1, D_TILE))) return rc;
  if ((rc = make_tma_2d_bf16(&tx, X, in_f, n, ldx, g.npad))) return rc;
  if ((rc = make_tma_2d_bf16(&tw2, W2, k, out_f, ldw2, D_TILE))) return rc;
  if ((rc = make_tma_2d_bf16(&th, g.Hb, g.ldh, g.npad, g.ldh, g.npad))) return rc;
  constexpr int D_SMEM = D_STAGES * 32768 + 1024 + 256;
  bool& attr = *lr_attr_flag(0);
  if (!attr) {
    if (cudaFuncSetAttribute(lowrank_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             D_SMEM) != cudaSuccess)
      return -12;
    attr = true;
  }
  const int units = std::max(g.rt1 * g.s1, g.rt2);
  const int grid = std::min(sms, units);
  void* params[] = {&tw1, &tx, &tw2, &th, &g};
  if (cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lowrank_decode_kernel), dim3(grid),
                                  dim3(D_THREADS), params, D_SMEM, st) != cudaSuccess)
    return -5;
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

bool fused_eligible(const void* X, long long ldx, const void* W1, long long ldw1, const void* W2,
                    long long ldw2, int is_bf16, int k) {
  return is_bf16 && k <= 256 && aligned16(X) && aligned16(W1) && aligned16(W2) && (ldx % 8) == 0 &&
         (ldw1 % 8) == 0 && (ldw2 % 8) == 0;
}

int launch_fused(const void* X, long long ldx, const void* W1, long long ldw1, const void* W2,
                 long long ldw2, const float* bias, void* Y, long long ldy, long long n, int in_f,
                 int k, int out_f, cudaStream_t st) {
  FusedArgs g;
  g.n = static_cast<int>(n);
  g.in_f = in_f;
  g.k = k;
  g.out_f = out_f;
  g.kp = static_cast<int>(round_up(k, 64));
  g.out_tiles = (out_f + F_TILE_N - 1) / F_TILE_N;
  g.Y = static_cast<__nv_bfloat16*>(Y);
  g.ldy = ldy;
  g.bias = bias;
  const int sms = device_sm_count();
  // Shape of the launch: rows per token tile (multiple of 8) and, when token tiles alone cannot
  // fill the GPU, a split of `out` over the CTAs of one token tile. Two-term model per candidate:
  // tensor time = waves x (128-row MMAs of one unit; the tensor core always works on 128 rows),
  // memory time = bytes / HBM bandwidth (X re-read by the CTAs sharing a token tile counted at half:
  // mostly L2 hits), plus a per-wave prologue / drain overhead.
  // Per-SM streaming rate: what one CTA's TMA ring sustains (bytes in flight / latency), measured
  // ~75 GB/s with 4 x 32 KB slots and ~1.3x that with the 7-slot GEMM-1 ring. Every unit streams
  // its X rows AND the weights (W1 whole, its share of W2), which is what makes small tiles lose.
  const double mma_rate = 1.6e15 / sms, hbm = 6.5e12, sm_stream = 95e9, wave_overhead = 3e-6;
  double best = 1e300;
  g.groups = 1;
  g.tiles_per_group = g.out_tiles;
  g.tile_m = F_TILE_M;
  g.ksplit = 0;
  g.prof = g_knob[9] != 0;
  const int tm_lo = g_knob[7] > 0 ? static_cast<int>(g_knob[7]) : 32;
  const int tm_hi = g_knob[7] > 0 ? static_cast<int>(g_knob[7]) : F_TILE_M;
  // k-split clusters (see FusedArgs::ksplit): pairs of out-groups share one GEMM 1. Needs an even
  // number of groups, the partial H in the 64 KB staging area (kp <= 128) and two k-blocks.
  const bool ks_possible = g.kp <= 128 && in_f > F_BK && g_knob[8] != 1;
  const double exchange = 4e-6;  // measured ~5 us (1.5 us waiting for the peer, 2.3 us of remote stores, 1.2 us back); 4 keeps the N = 2048, k = 32 win
  for (int tm = tm_hi; tm >= tm_lo; tm -= 8) {
    const long long rtiles = (n + tm - 1) / tm;
    for (int G = 1; G <= g.out_tiles; ++G) {
      const int tpg = (g.out_tiles + G - 1) / G;
      const int geff = (g.out_tiles + tpg - 1) / tpg;
      const long long units = rtiles * geff;
      const long long waves = (units + sms - 1) / sms;
      for (int ks = 0; ks <= ((ks_possible && geff % 2 == 0) ? 1 : 0); ++ks) {
        if (g_knob[8] == 2 && ks_possible && geff % 2 == 0 && ks == 0) continue;
        const double in_eff = ks ? 0.5 * in_f : static_cast<double>(in_f);
        const double readers = ks ? 0.5 * geff : static_cast<double>(geff);  // CTAs reading all of X's tile
        const double t_mma = static_cast<double>(waves) * 2.0 * F_TILE_M * g.kp *
                             (in_eff + static_cast<double>(tpg) * F_TILE_N) / mma_rate;
        const double bytes = 2.0 * static_cast<double>(n) * (in_f * (1.0 + 0.5 * (readers - 1.0)) + out_f);
        const double unit_bytes = 2.0 * (static_cast<double>(tm) * in_eff + static_cast<double>(g.kp) * in_eff +
                                         static_cast<double>(tpg) * F_TILE_N * (g.kp + tm));
        const double t_mem = std::max(bytes / hbm, static_cast<double>(waves) * unit_bytes / sm_stream);
        const double cost = std::max(t_mma, t_mem) + (wave_overhead + (ks ? exchange : 0.0)) * static_cast<double>(waves);
        if (cost < best * 0.98) {
          best = cost;
          g.groups = geff;
          g.tiles_per_group = tpg;
          g.tile_m = tm;
          g.ksplit = ks;
        }
      }
    }
  }
  const int row_tiles = static_cast<int>((n + g.tile_m - 1) / g.tile_m);
  CUtensorMap tx, tw1, tw2, ty;
  int rc;
  const bool persistent = g.kp <= 128 && row_tiles >= sms && aligned16(Y) && (ldy % 8) == 0 &&
                          g_knob[3] == 0;
  g.tma_store = (aligned16(Y) && (ldy % 8) == 0 && g_knob[4] == 0) ? 1 : 0;
  if (g.tma_store) {
    if ((rc = make_tma_2d_bf16(&ty, Y, out_f, n, ldy, 8))) return rc;
  } else {
    memset(&ty, 0, sizeof(ty));
  }
  if ((rc = make_tma_2d_bf16(&tx, X, in_f, n, ldx, g.tile_m))) return rc;
  if ((rc = make_tma_2d_bf16(&tw1, W1, in_f, k, ldw1, g.kp))) return rc;
  if ((rc = make_tma_2d_bf16(&tw2, W2, k, out_f, ldw2, F_TILE_N))) return rc;
  if (g.kp <= 128) {
    g.stages = 4; g.slot_bytes = 32768; g.w2_off = 0; g.stg_separate = 1;
    // GEMM-1 ring over the slots + the (then idle) H block and output staging
    g.stages1 = g_knob[6] == 1 ? g.stages
                               : g.stages + ((g.kp / F_BK) * F_HBLOCK + F_STG_BYTES) / g.slot_bytes;
  } else {
    g.stages = 3; g.slot_bytes = 49152; g.w2_off = F_XBYTES; g.stg_separate = 0;
    // one more GEMM-1 slot over the idle H block (+ the 16 KB staging tail)
    g.stages1 = g_knob[6] == 1 ? g.stages
                               : g.stages + ((g.kp / F_BK) * F_HBLOCK + 4 * 4096) / g.slot_bytes;
  }
  g.rotate = 1;
  if (g_knob[5] >= 0) g.rotate = static_cast<int>(g_knob[5]);
  if (g_knob[6] >= 2 && g_knob[6] <= g.stages) {  // experiment knob: a shorter single ring
    g.stages = static_cast<int>(g_knob[6]);
    g.stages1 = g.stages;
  }
  const int F_SMEM = g.stages * g.slot_bytes + (g.kp / F_BK) * F_HBLOCK +
                     (g.stg_separate ? F_STG_BYTES : 4 * 4096) + 1024 + 512;
  bool& attr = *lr_attr_flag(1);
  if (!attr) {
    if (cudaFuncSetAttribute(lowrank_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024) != cudaSuccess)
      return -12;
    attr = true;
  }
  if (persistent) {
    bool& pattr = *lr_attr_flag(2);
    if (!pattr) {
      if (cudaFuncSetAttribute(lowrank_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               227 * 1024) != cudaSuccess)
        return -12;
      pattr = true;
    }
    g.groups = 1;
    g.tiles_per_group = g.out_tiles;
    g.ksplit = 0;
    lowrank_persistent_kernel<<<sms, F_THREADS, P_SMEM, st>>>(tx, tw1, tw2, ty, g);
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
  }
  const long long grid = static_cast<long long>(row_tiles) * g.groups;
  if (grid > 0x7fffffffLL) return -22;
  if (g.ksplit) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(F_THREADS);
    cfg.dynamicSmemBytes = F_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, lowrank_fused_kernel, tx, tw1, tw2, ty, g) != cudaSuccess) return -5;
    return cudaGetLastError() == cudaSuccess ? 0 : -5;
  }
  lowrank_fused_kernel<<<static_cast<unsigned>(grid), F_THREADS, F_SMEM, st>>>(tx, tw1, tw2, ty, g);
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}

}  // namespace

long long lowrank_debug_get(int k) {
  unsigned long long v[8];
  if (k < 0 || k >= 8 || cudaMemcpyFromSymbol(v, g_fused_prof, sizeof(v)) != cudaSuccess) return -1;
  return static_cast<long long>(v[k]);
}

void lowrank_debug_set(int key, long long value) {
  if (key == 9) {
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(g_fused_prof, z, sizeof(z));
  }
  if (key >= 0 && key < 10) g_knob[key] = value;
}

size_t lowrank_workspace_bytes(int is_bf16, long long n, int in_f, int k, int out_f) {
  Carve cv{nullptr};
  GemmOperand op;
  stage(nullptr, is_bf16, n, in_f, in_f, cv, &op, nullptr, true);
  stage(nullptr, is_bf16, k, in_f, in_f, cv, &op, nullptr, true);
  stage(nullptr, is_bf16, out_f, k, k, cv, &op, nullptr, true);
  const int nseg = is_bf16 ? 1 : 3;
  cv.take(2ull * nseg * n * round_up(k, 8));
  size_t need = cv.off + 512;
  if (is_bf16 && n <= 128) need = std::max(need, decode_workspace_bytes(n, k));
  return need;
}

int lowrank_forward(const void* X, long long ldx, const void* W1, long long ldw1, const void* W2,
                    long long ldw2, const float* bias, void* Y, long long ldy, int is_bf16,
                    long long n, int in_f, int k, int out_f, void* ws, size_t ws_bytes,
                    cudaStream_t st) {
  if (X == nullptr || W1 == nullptr || W2 == nullptr || Y == nullptr) return -22;
  if (n <= 0 || in_f <= 0 || k <= 0 || out_f <= 0 || n > 0x7fffffffLL) return -22;
  if (ldx < in_f || ldw1 < in_f || ldw2 < k || ldy < out_f) return -22;
  if (decode_eligible(X, ldx, W1, ldw1, W2, ldw2, is_bf16, n, in_f, k, out_f) && ws != nullptr &&
      ws_bytes >= decode_workspace_bytes(n, k) + 256 && (reinterpret_cast<uintptr_t>(Y) & 1) == 0 &&
      g_knob[1] == 0)
    return launch_decode(X, ldx, W1, ldw1, W2, ldw2, bias, Y, ldy, n, in_f, k, out_f, ws, st);
  if (fused_eligible(X, ldx, W1, ldw1, W2, ldw2, is_bf16, k) && g_knob[2] == 0)
    return launch_fused(X, ldx, W1, ldw1, W2, ldw2, bias, Y, ldy, n, in_f, k, out_f, st);
  if (ws == nullptr || ws_bytes < lowrank_workspace_bytes(is_bf16, n, in_f, k, out_f)) return -12;
  uint8_t* base = static_cast<uint8_t*>(ws);
  base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(base) + 255) & ~uintptr_t(255));
  Carve cv{base};
  GemmOperand x, w1, w2;
  int rc;
  if ((rc = stage(X, is_bf16, n, in_f, ldx, cv, &x, st, false))) return rc;
  if ((rc = stage(W1, is_bf16, k, in_f, ldw1, cv, &w1, st, false))) return rc;
  if ((rc = stage(W2, is_bf16, out_f, k, ldw2, cv, &w2, st, false))) return rc;
  const int nseg = is_bf16 ? 1 : 3;
  const long long kp = round_up(k, 8);
  __nv_bfloat16* H = static_cast<__nv_bfloat16*>(cv.take(2ull * nseg * n * kp));
  // H = X W1^T  (K-major x K-major), written as bf16 (bf16 models) or bf16x3 split (fp32 models)
  GemmEpilogue e1;
  if (is_bf16) {
    e1.Cb = H;
    e1.ldcb = kp;
  } else {
    e1.Cs = H;
    e1.ldcs = kp;
    e1.cs_seg = n * kp;
  }
  if ((rc = gemm_tc(x, w1, static_cast<int>(n), k, in_f, -1, e1, st))) return rc;
  GemmOperand h{H, 0, kp, nseg, n * kp};
  GemmEpilogue e2;
  e2.bias = bias;
  if (is_bf16) {
    e2.Cb = static_cast<__nv_bfloat16*>(Y);
    e2.ldcb = ldy;
  } else {
    e2.C = static_cast<float*>(Y);
    e2.ldc = ldy;
  }
  return gemm_tc(h, w2, static_cast<int>(n), out_f, k, -1, e2, st);
}

}  // namespace ptd
