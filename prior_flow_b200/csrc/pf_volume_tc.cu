// (a) All-pairs correlation volume with the avg-pool pyramid fused into the epilogue —
// tcgen05 / TMEM / TMA kernel for sm_100a.
// Replaces PriOr_RAFT.corr (PriOr-RAFT/core/prior_raft.py:69-75: fp32 cuBLAS GEMM + a separate
// full-volume `div` pass) and DCCL.build_pyramid (core/corr.py:99-111: three avg_pool2d passes).
//
// Arithmetic.  tcgen05 has no fp32-input kind, so the fp32 contraction is done as an fp16 hi/lo
// split of both operands (x*s = hi + lo, s a power of two chosen from the tensor's absmax so that
// hi never overflows fp16) and three kind::f16 MMAs into one fp32 TMEM accumulator:
//     A*B ~= hi_a*hi_b + hi_a*lo_b + lo_a*hi_b          (dropped term lo*lo ~ 2^-22 relative)
// fp16 carries the same 11 significant bits as TF32, so this is the "3xTF32" scheme at twice the
// tensor rate and half the operand bytes.  PF_VOL_F16 issues the first product only.
//
// Structure (persistent, one CTA per SM, 10 warps):
//   prep kernels : absmax -> power-of-two scale ; transpose [C, N] fp32 -> K-major [N, C] fp16 hi/lo
//   warp 0       : TMA producer.  A tile = 128 query rows x 64 k (2-D map); B tile = 8x32 *patch*
//                  of target pixels x 64 k (4-D map over [B, h, w, C]) — the N-tile is a spatial
//                  patch so that the 2x2 / 4x4 / 8x8 pools are tile-local.  SWIZZLE_128B, 96 KB/stage.
//   warp 1       : TMEM alloc (512 cols = 2 accumulator stages of 128 x 256 fp32) and the single
//                  thread issuing tcgen05.mma (M128 N256 K16), tcgen05.commit -> mbarriers.
//   warps 2..9   : epilogue, two groups of 4 warps alternating tiles (one per TMEM accumulator stage).  tcgen05.ld (thread = query row, registers = one patch row of 32
//                  targets), scale by 1/(sqrt(C) s_a s_b), level 0 staged in swizzled smem and
//                  written with TMA (3-D map over [B*N, h, w]); levels 1..3 pooled in registers
//                  in avg_pool2d's order ((a+b)+c+d)/4 from the rounded finer level and stored
//                  as 64/32/16-byte row segments.
// HBM traffic is the compulsory 1.33 x N^2 x 4 B of pyramid writes; the operands (fp16 planes,
// 4 x 4 MiB per view at 512x1024) stay in L2.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "pf_common.cuh"
#include "pf_tc.cuh"

namespace pf {

constexpr int BM = 128;                 // query rows per tile == TMEM lanes
constexpr int BN = 256;                 // target pixels per tile
constexpr int PATCH_H = 8, PATCH_W = 32;
constexpr int BK = 64;                  // fp16 per k-block: 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_PLANE = BM * BK * 2;    // 16 KiB
constexpr int B_PLANE = BN * BK * 2;    // 32 KiB
constexpr int OUT_STAGE = BM * 32 * 4;  // 16 KiB: 128 rows x one 32-float patch row
constexpr int kOutGroups = 2;          // epilogue groups, one staging buffer (or two, see Cfg::kOutBufs) each
constexpr int kTcThreads = 320;         // TMA warp, MMA warp, 2 x 4 epilogue warps
constexpr int kAccCols = BN;            // fp32 accumulator columns per stage
constexpr uint32_t kTmemCols = 512;

// kTwoSm: one tcgen05.mma.cta_group::2 spans the CTA pair (M = 256); a CTA stages its own 128 query rows and only ITS HALF
// of the patch, so a stage is 64 KiB instead of 96 (three stages fit) and a third less operand data crosses L2 -> SM.
template <bool kSplit, bool kTwoSm = false>
struct Cfg {
  static constexpr int kBBytes = kTwoSm ? B_PLANE / 2 : B_PLANE;
  static constexpr int kPlaneBytes = A_PLANE + kBBytes;
  static constexpr int kStageBytes = (kSplit ? 2 : 1) * kPlaneBytes;
  // 2-SM: the smaller stages leave room for DOUBLE-BUFFERED level-0 staging (the TMA store of patch row c is still reading
  // its buffer while row c + 1 is being staged into the other one)
  static constexpr int kOutBufs = kTwoSm ? 2 : 1;
  static constexpr int kStages = kTwoSm ? (kSplit ? 2 : 4) : (kSplit ? 2 : 4);
  static constexpr int kUsedBytes = kStages * kStageBytes + kOutGroups * kOutBufs * OUT_STAGE + 256 /*barriers*/ + 2048 /*level-3 exchange*/;
  // the split mode uses 226.25 KiB of the 227 KiB a CTA may have: the 1 KiB alignment slack does not fit entirely.  The dynamic
  // window starts 1 KiB aligned when the kernel has no static shared memory (it has none); the kernel traps if it ever does not.
  static constexpr int kSmemBytes = (kUsedBytes + 1024 <= 232448) ? kUsedBytes + 1024 : 232448;
};

// kind::f16 instruction descriptor: D = F32, A = B = F16, both K-major, M = 128, N = 256.
constexpr uint32_t kIdesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t kIdesc2Sm = umma_idesc_f16(2 * BM, BN);   // cta_group::2: M = 256 across the pair

// Power-of-two scale that maps absmax into [2^13, 2^14): hi = fp16(x*s) cannot overflow.
__device__ __forceinline__ float split_scale(uint32_t amax_bits) {
  int e = (int)((amax_bits >> 23) & 0xff) - 127;  // floor(log2(absmax))
  if ((amax_bits & 0x7fffffffu) == 0u) e = 13;    // all-zero tensor: s = 1
  int se = 13 - e;
  se = se < -100 ? -100 : (se > 100 ? 100 : se);
  return __uint_as_float((uint32_t)(se + 127) << 23);
}

// ---------------------------------------------------------------------------------- prep kernels
// blockIdx.y selects the operand (0: fmap1, 1: fmap2); float4 loads, one atomicMax per warp.
__global__ void __launch_bounds__(256) absmax_kernel(const float *__restrict__ x0, const float *__restrict__ x1, long long n,
                                                      uint32_t *__restrict__ out) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // split_transpose_kernel may start loading its tiles
  const float *x = blockIdx.y ? x1 : x0;
  const float4 *x4 = reinterpret_cast<const float4 *>(x);
  const long long n4 = n >> 2;
  uint32_t m = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x4 + i);
    m = max(max(m, __float_as_uint(fabsf(v.x))), __float_as_uint(fabsf(v.y)));
    m = max(max(m, __float_as_uint(fabsf(v.z))), __float_as_uint(fabsf(v.w)));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) m = max(m, __float_as_uint(fabsf(x[(n4 << 2) + threadIdx.x])));
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out + blockIdx.y, m);
}

// [B, C, N] fp32  ->  K-major [B, N, C] fp16 hi (and lo) planes.  64(c) x 32(n) tile through smem;
// blockIdx.z = operand * B + batch.
__global__ void __launch_bounds__(256) split_transpose_kernel(const float *__restrict__ x0, const float *__restrict__ x1,
                                                              __half *__restrict__ hi0, __half *__restrict__ lo0,
                                                              __half *__restrict__ hi1, __half *__restrict__ lo1, int B,
                                                              int C, int N, const uint32_t *__restrict__ amax_bits,
                                                              int want_lo) {
  __shared__ float t[64][33];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // lets volume_tc_kernel's prologue overlap this grid's tail
  const int which = blockIdx.z / B, b = blockIdx.z - which * B;
  const float *x = which ? x1 : x0;
  __half *hi = which ? hi1 : hi0, *lo = which ? lo1 : lo0;
  const int c0 = blockIdx.y * 64, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float *xb = x + (long long)b * C * N;
  float raw[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + ty + i * 8;
    raw[i] = (c < C && n0 + tx < N) ? xb[(long long)c * N + n0 + tx] : 0.f;
  }
  // programmatic dependent of absmax_kernel: the tile above is an input; only the scale needs that grid's result
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const float s = split_scale(amax_bits[which]);
#pragma unroll
  for (int i = 0; i < 8; ++i) t[ty + i * 8][tx] = raw[i] * s;
  __syncthreads();
  // each thread writes one half2 (channels 2*tx, 2*tx+1) for rows ty, ty+8, ...
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty + i * 8;
    const int c = c0 + 2 * tx;
    if (n < N && c + 1 < C) {
      const float v0 = t[2 * tx][ty + i * 8], v1 = t[2 * tx + 1][ty + i * 8];
      const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
      const long long o = ((long long)b * N + n) * C + c;
      *reinterpret_cast<__half2 *>(hi + o) = __halves2half2(h0, h1);
      if (want_lo)
        *reinterpret_cast<__half2 *>(lo + o) =
            __halves2half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
    }
  }
}

// ---------------------------------------------------------------------------------- main kernel
struct TcParams {
  int B, C, N, h, w;
  int num_levels;
  int tiles_m, patches_x, patches_y;  // per batch
  int contiguous;                     // work-item assignment (see the kernel)
  long long total_tiles;
  float inv_sqrt_c;
  const uint32_t *amax_bits;  // [2]: fmap1, fmap2
  float *lvl0, *lvl1, *lvl2, *lvl3;
};

// kCluster = 2: CTA pairs (thread-block cluster 2x1x1) work on two M-tiles of the SAME target patch in lockstep; each
// CTA fetches half of the B (patch) tile and TMA-multicasts it to both, so the L2 -> SM operand traffic per tile drops
// from A + B to A + B/2 (ncu r01a: that traffic, not the tensor pipe or HBM, bounded the 1-CTA kernel).  A stage is
// recycled only when BOTH CTAs' MMAs have retired it (commit multicast to both empty barriers, count 2).
// kShare (r03): BOTH epilogue groups drain the SAME tile (group g takes patch rows 4g .. 4g+3), so the MMA of tile t + 1 runs
// into the other accumulator while tile t is drained.  With one accumulator per group (r02) a group idles for the whole MMA
// of its next tile: 6.9 tiles x (3.3 us MMA + 14.5 us epilogue) per group in the fp32 split mode.
// kStg (r03): level 0 leaves the staging buffer with coalesced st.global.v4 (8 lanes per 128-byte line, four full lines per
// warp instruction) instead of one TMA store of 128 separate 128-byte rows per patch row: scripts/probe/write_probe.cu shows
// that pattern sustaining 4.1 TB/s, while the TMA-store epilogue topped out at ~3.0 TB/s in every variant of the kernel.
template <bool kSplit, int kCluster, bool kTwoSm = false, bool kShare = false, bool kStg = false>
__global__ void __launch_bounds__(kTcThreads, 1)
volume_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                 const __grid_constant__ CUtensorMap map_out, const TcParams p) {
  static_assert(!kTwoSm || kCluster == 2, "cta_group::2 needs the CTA pair");
  using C_ = Cfg<kSplit, kTwoSm>;
  constexpr int kStages = C_::kStages;
  constexpr int kPlaneBytes = C_::kPlaneBytes;   // [A 16 KiB | B (half)] of one plane; the lo plane follows the hi plane
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  if (smem + C_::kUsedBytes > smem_raw + C_::kSmemBytes) __trap();   // see Cfg::kSmemBytes
  uint8_t *out_stage = smem + kStages * C_::kStageBytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(out_stage + kOutGroups * C_::kOutBufs * OUT_STAGE);
  // bars: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]; then the TMEM base address
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kStages);
  const uint32_t bar_tfull = smem_u32(bars + 2 * kStages), bar_tempty = smem_u32(bars + 2 * kStages + 2);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = p.C / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, kTwoSm ? 1 : kCluster);   // 2-SM: ONE multicast commit arrives per CTA
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, (kTwoSm ? 4 * kCluster : 4) * (kShare ? 2 : 1));   // 2-SM: both CTAs' epilogue warps release the leader's MMA thread
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (kTwoSm) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (kCluster > 1)
    cluster_sync_all();   // the peer's barriers must be initialised before anything is multicast at them
  else
    __syncthreads();
  tc_fence_after();
  // programmatic dependent launch: the prologue above (barrier init, TMEM allocation, cluster sync) touches no global
  // data and overlaps the tail of split_transpose_kernel; the operand planes and the absmax words are read only after
  // that grid has completed (no-op when launched without the attribute)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t cta_rank = kCluster > 1 ? cluster_ctarank() : 0;
  // work items: kCluster consecutive M-tiles of one patch per cluster, so the peers share the B tile
  // Work items (kCluster M-tiles x one patch) in M-major, patch-minor order.  PF_VOLUME_ORDER (p.contiguous): a cluster takes a
  // CONTIGUOUS range of them, so its consecutive tiles are neighbouring patches of the same query planes and every plane row
  // (512 B = four patches) is completed within a few tiles — DRAM sees runs instead of isolated 128-byte lines; otherwise the
  // r02 round-robin assignment.
  const long long total_items = p.total_tiles / kCluster;
  const long long n_clusters = gridDim.x / kCluster, cluster_id = blockIdx.x / kCluster;
  const long long first_item = p.contiguous ? total_items * cluster_id / n_clusters : cluster_id;
  const long long end_item = p.contiguous ? total_items * (cluster_id + 1) / n_clusters : total_items;
  const long long item_stride = p.contiguous ? 1 : n_clusters;
  const int tiles_m_items = p.tiles_m / kCluster;
  constexpr uint16_t kMask = (1u << kCluster) - 1;

  if (warp == 0) {
    // ================================================================= TMA producer (one thread)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long item = first_item; item < end_item; item += item_stride) {
        const int per_b = tiles_m_items * p.patches_x * p.patches_y;
        const int b = (int)(item / per_b);
        int r = (int)(item - (long long)b * per_b);
        const int mt = (r / (p.patches_x * p.patches_y)) * kCluster + (int)cta_rank;
        r %= p.patches_x * p.patches_y;
        const int py = r / p.patches_x, px = r - py * p.patches_x;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          const uint32_t sbase = smem_u32(smem + stage * C_::kStageBytes);
          const int row = b * p.N + mt * BM;
          if (kTwoSm) {
            // both CTAs' loads complete on the LEADER's barrier: it expects the pair's bytes
            if (cta_rank == 0) mbar_arrive_expect_tx(full, 2u * (uint32_t)C_::kStageBytes);
            constexpr int kHalfRows = PATCH_H / 2;
            const int y = py * PATCH_H + (int)cta_rank * kHalfRows;
            tma_load_2d_2sm(sbase, &map_a_hi, full, kb * BK, row);
            tma_load_4d_2sm(sbase + A_PLANE, &map_b_hi, full, kb * BK, px * PATCH_W, y, b);
            if (kSplit) {
              tma_load_2d_2sm(sbase + kPlaneBytes, &map_a_lo, full, kb * BK, row);
              tma_load_4d_2sm(sbase + kPlaneBytes + A_PLANE, &map_b_lo, full, kb * BK, px * PATCH_W, y, b);
            }
          } else {
            mbar_arrive_expect_tx(full, (uint32_t)C_::kStageBytes);
            tma_load_2d(sbase, &map_a_hi, full, kb * BK, row);
            if (kSplit) tma_load_2d(sbase + A_PLANE + B_PLANE, &map_a_lo, full, kb * BK, row);
            if (kCluster == 1) {
              tma_load_4d(sbase + A_PLANE, &map_b_hi, full, kb * BK, px * PATCH_W, py * PATCH_H, b);
              if (kSplit) tma_load_4d(sbase + 2 * A_PLANE + B_PLANE, &map_b_lo, full, kb * BK, px * PATCH_W, py * PATCH_H, b);
            } else {
              // my half of the patch rows (box height PATCH_H / kCluster), delivered to both CTAs
              constexpr int kHalfRows = PATCH_H / kCluster, kHalfBytes = B_PLANE / kCluster;
              const int y = py * PATCH_H + (int)cta_rank * kHalfRows;
              tma_load_4d_mc(sbase + A_PLANE + cta_rank * kHalfBytes, &map_b_hi, full, kb * BK, px * PATCH_W, y, b, kMask);
              if (kSplit)
                tma_load_4d_mc(sbase + 2 * A_PLANE + B_PLANE + cta_rank * kHalfBytes, &map_b_lo, full, kb * BK, px * PATCH_W, y, b,
                               kMask);
            }
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer (one thread)
    if (lane == 0 && (!kTwoSm || cta_rank == 0)) {     // 2-SM: only the leader CTA issues; its MMAs write both CTAs' TMEM
      int stage = 0;
      uint32_t phase = 0, acc = 0, acc_phase = 0;
      for (long long item = first_item; item < end_item; item += item_stride) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kAccCols;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sbase = smem_u32(smem + stage * C_::kStageBytes);
          // stage layout: multicast path [A.hi | B.hi | A.lo | B.lo] with whole patches; 2-SM path the same with HALF patches
          const uint64_t a_hi = make_smem_desc(sbase), b_hi = make_smem_desc(sbase + A_PLANE);
          const uint64_t a_lo = make_smem_desc(sbase + kPlaneBytes);
          const uint64_t b_lo = make_smem_desc(sbase + kPlaneBytes + A_PLANE);
          auto mma = [&](uint64_t da, uint64_t db, uint32_t accum) {
            if (kTwoSm) umma_f16_2sm(tmem_d, da, db, kIdesc2Sm, accum); else umma_f16(tmem_d, da, db, kIdesc, accum);
          };
          if (kSplit) {
            // small cross terms first, then the leading product
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) mma(a_lo + 2 * k, b_hi + 2 * k, (kb | k) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) mma(a_hi + 2 * k, b_lo + 2 * k, 1u);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) mma(a_hi + 2 * k, b_hi + 2 * k, 1u);
          } else {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) mma(a_hi + 2 * k, b_hi + 2 * k, (kb | k) ? 1u : 0u);
          }
          // frees the smem stage when these MMAs retire — in both CTAs (multicast path: the peer multicasts into my smem too;
          // 2-SM path: the MMA read both CTAs' stages)
          if (kTwoSm)
            umma_commit_2sm(bar_empty + 8 * stage, kMask);
          else if (kCluster > 1)
            umma_commit_mc(bar_empty + 8 * stage, kMask);
          else
            umma_commit(bar_empty + 8 * stage);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator complete -> epilogue (2-SM: of both CTAs)
        if (kTwoSm) umma_commit_2sm(bar_tfull + 8 * acc, kMask); else umma_commit(bar_tfull + 8 * acc);
        if ((acc ^= 1) == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ================================================================= epilogue (2 groups x 4 warps)
    // Group g drains TMEM accumulator stage g, i.e. every other tile of this CTA, so two tiles' epilogues overlap
    // (ncu r01a stall sampling: a single 4-warp epilogue was the critical path, ~12k cycles per tile in both the
    // fp32-split and the f16 mode).  Each group owns one 16 KiB staging buffer and walks the 8 patch rows one at a
    // time; the TMEM load of the next row is in flight while the current one is staged, stored and pooled.
    const int group = (warp - 2) >> 2;     // 0: warps 2-5, 1: warps 6-9
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;   // query row inside the tile
    const bool leader = (threadIdx.x & 127) == 64;   // thread 64 (group 0) / 192 (group 1): first lane of a group's first warp
    const float scale = p.inv_sqrt_c / (split_scale(p.amax_bits[0]) * split_scale(p.amax_bits[1]));
    uint32_t acc = kShare ? 0u : (uint32_t)group;
    uint32_t acc_phase = 0;
    // kShare: the level-3 pool spans both groups' rows: group 0 hands its half of the sum, (a + b) of avg_pool2d's
    // ((a + b) + c + d) / 4, to group 1 through this double-buffered exchange (the tail of the barrier block)
    float *l3_xchg = reinterpret_cast<float *>(bars) + 64;     // [128 rows][4] floats = 2 KiB behind the barriers
    uint8_t *stage_base = out_stage + group * C_::kOutBufs * OUT_STAGE;
    const int w1 = p.w >> 1, h1 = p.h >> 1, w2 = p.w >> 2, h2 = p.h >> 2, w3 = p.w >> 3, h3 = p.h >> 3;
    for (long long item = first_item + (kShare ? 0 : group) * item_stride; item < end_item; item += (kShare ? 1 : 2) * item_stride) {
      const int per_b = tiles_m_items * p.patches_x * p.patches_y;
      const int b = (int)(item / per_b);
      int r = (int)(item - (long long)b * per_b);
      const int mt = (r / (p.patches_x * p.patches_y)) * kCluster + (int)cta_rank;
      r %= p.patches_x * p.patches_y;
      const int py = r / p.patches_x, px = r - py * p.patches_x;
      const long long qrow = (long long)b * p.N + mt * BM + row;  // global query index (plane index)

      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * kAccCols + ((uint32_t)(quarter * 32) << 16);
      constexpr int kRows = kShare ? PATCH_H / 2 : PATCH_H;      // patch rows this group drains
      const int c0 = kShare ? group * kRows : 0;
      float *xchg = l3_xchg + row * 4;
      uint32_t un[32];
      tmem_ld32(taddr + c0 * 32, un);
      float vp[32], l1_prev[16], l2_prev[8];
#pragma unroll
      for (int cc = 0; cc < kRows; ++cc) {
        const int c = c0 + cc;
        tmem_ld_wait();
        float vc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) vc[j] = __uint_as_float(un[j]) * scale;
        if (cc + 1 < kRows) {
          tmem_ld32(taddr + (c + 1) * 32, un);   // prefetch the next patch row
        } else {                                  // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kTwoSm) mbar_arrive_cluster(bar_tempty + 8 * acc, 0); else mbar_arrive(bar_tempty + 8 * acc);
          }
        }
        // ---- level 0: one patch row through swizzled staging + TMA store
        uint8_t *stage_buf = stage_base + (C_::kOutBufs == 2 ? (c & 1) * OUT_STAGE : 0);
        if (!kStg && leader) {
          if (C_::kOutBufs == 2) tma_store_wait_read1(); else tma_store_wait_read0();   // the store that last used THIS buffer has read it
        }
        epi_bar_sync(group);
        {
          uint8_t *r0 = stage_buf + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4 *>(r0 + ((j ^ (row & 7)) << 4)) = make_float4(vc[4 * j], vc[4 * j + 1], vc[4 * j + 2], vc[4 * j + 3]);
        }
        if (!kStg) fence_async_smem();
        epi_bar_sync(group);
        if (kStg) {
          // 128 rows x 128 B: thread t handles 16-byte chunks t, t + 128, ...: 8 consecutive lanes = one full line
          const int t = threadIdx.x & 127;
          float *l0 = p.lvl0 + (((long long)b * p.N + mt * BM) * p.h + (py * PATCH_H + c)) * p.w + px * PATCH_W;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int idx = i * 128 + t, rr = idx >> 3, ch = idx & 7;
            const float4 v = *reinterpret_cast<const float4 *>(stage_buf + rr * 128 + ((ch ^ (rr & 7)) << 4));
            *reinterpret_cast<float4 *>(l0 + (long long)rr * p.h * p.w + ch * 4) = v;
          }
        } else if (leader) {
          tma_store_3d(&map_out, smem_u32(stage_buf), px * PATCH_W, py * PATCH_H + c, b * p.N + mt * BM);
          tma_store_commit();
        }
        if (c & 1) {
          const int cp = c >> 1;
          // ---- level 1: ((a + b) + c + d) / 4 over the 2x2 window, row-major order (avg_pool2d)
          if (p.num_levels > 1) {
            float l1[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              l1[j] = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(vp[2 * j], vp[2 * j + 1]), vc[2 * j]), vc[2 * j + 1]), 0.25f);
            float *d1 = p.lvl1 + (qrow * h1 + (py * (PATCH_H / 2) + cp)) * w1 + px * (PATCH_W / 2);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<float4 *>(d1 + 4 * j) = make_float4(l1[4 * j], l1[4 * j + 1], l1[4 * j + 2], l1[4 * j + 3]);
            if (p.num_levels > 2) {
              if (cp & 1) {
                float l2[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  l2[j] = __fmul_rn(
                      __fadd_rn(__fadd_rn(__fadd_rn(l1_prev[2 * j], l1_prev[2 * j + 1]), l1[2 * j]), l1[2 * j + 1]), 0.25f);
                float *d2 = p.lvl2 + (qrow * h2 + (py * (PATCH_H / 4) + (cp >> 1))) * w2 + px * (PATCH_W / 4);
                *reinterpret_cast<float4 *>(d2) = make_float4(l2[0], l2[1], l2[2], l2[3]);
                *reinterpret_cast<float4 *>(d2 + 4) = make_float4(l2[4], l2[5], l2[6], l2[7]);
                if (p.num_levels > 3) {
                  if (kShare) {
                    if (cp == 1) {      // group 0: rows 0-3 done
                      *reinterpret_cast<float4 *>(xchg) = make_float4(__fadd_rn(l2[0], l2[1]), __fadd_rn(l2[2], l2[3]),
                                                                      __fadd_rn(l2[4], l2[5]), __fadd_rn(l2[6], l2[7]));
                    }
                    asm volatile("bar.sync 3, 256;" ::: "memory");   // both groups, once per tile: the partial sums are visible
                  }
                  if (cp == 3) {
                    float l3[4];
                    if (kShare) {
                      const float4 top = *reinterpret_cast<const float4 *>(xchg);
                      l3[0] = __fmul_rn(__fadd_rn(__fadd_rn(top.x, l2[0]), l2[1]), 0.25f);
                      l3[1] = __fmul_rn(__fadd_rn(__fadd_rn(top.y, l2[2]), l2[3]), 0.25f);
                      l3[2] = __fmul_rn(__fadd_rn(__fadd_rn(top.z, l2[4]), l2[5]), 0.25f);
                      l3[3] = __fmul_rn(__fadd_rn(__fadd_rn(top.w, l2[6]), l2[7]), 0.25f);
                    } else {
#pragma unroll
                      for (int j = 0; j < 4; ++j)
                        l3[j] = __fmul_rn(
                            __fadd_rn(__fadd_rn(__fadd_rn(l2_prev[2 * j], l2_prev[2 * j + 1]), l2[2 * j]), l2[2 * j + 1]),
                            0.25f);
                    }
                    float *d3 = p.lvl3 + (qrow * h3 + py) * w3 + px * (PATCH_W / 8);
                    *reinterpret_cast<float4 *>(d3) = make_float4(l3[0], l3[1], l3[2], l3[3]);
                  } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) l2_prev[j] = l2[j];
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) l1_prev[j] = l1[j];
              }
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) vp[j] = vc[j];
        }
      }
      if (kShare) {
        if ((acc ^= 1) == 0) acc_phase ^= 1;
      } else {
        acc_phase ^= 1;
      }
    }
    if (leader) tma_store_wait_all();
  }

  tc_fence_before();
  if (kCluster > 1)
    cluster_sync_all();   // no CTA may exit while its peer can still multicast into it or signal its barriers
  else
    __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kTwoSm)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------- host side
static long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

long long volume_tc_workspace_bytes(int batch, int channels, int h, int w, int mode) {
  const long long plane = align_up((long long)batch * h * w * channels * 2, 1024);
  const int planes = (mode == PF_VOL_FP32_3XF16) ? 4 : 2;
  return 1024 + planes * plane;
}

int volume_build_tc(const pf_volume_args *a, cudaStream_t st) {
  const int B = a->batch, C = a->channels, h = a->h, w = a->w, N = h * w;
  const bool split = a->mode == PF_VOL_FP32_3XF16;
  PF_REQUIRE(w % PATCH_W == 0 && h % PATCH_H == 0 && N % BM == 0 && C % BK == 0,
             "pf_volume_build(tcgen05): need w %% 32 == 0, h %% 8 == 0, h*w %% 128 == 0, C %% 64 == 0 (got %dx%d, C=%d); "
             "use PF_VOL_FP32_SIMT for other shapes",
             h, w, C);
  PF_REQUIRE((long long)B * N < (1LL << 31), "pf_volume_build: batch*h*w too large");
  const long long need = volume_tc_workspace_bytes(B, C, h, w, a->mode);
  PF_REQUIRE(a->workspace && a->workspace_bytes >= need, "pf_volume_build: workspace too small (%lld < %lld bytes)",
             a->workspace_bytes, need);
  PF_REQUIRE(((uintptr_t)a->workspace & 1023) == 0, "pf_volume_build: workspace must be 1 KiB aligned");
  for (int l = 0; l < a->num_levels; ++l)
    PF_REQUIRE(((uintptr_t)a->level[l] & 15) == 0, "pf_volume_build: level[%d] must be 16-byte aligned", l);

  uint8_t *ws = reinterpret_cast<uint8_t *>(a->workspace);
  uint32_t *amax = reinterpret_cast<uint32_t *>(ws);
  const long long plane = align_up((long long)B * N * C * 2, 1024);
  __half *a_hi = reinterpret_cast<__half *>(ws + 1024);
  __half *b_hi = reinterpret_cast<__half *>(ws + 1024 + plane);
  __half *a_lo = split ? reinterpret_cast<__half *>(ws + 1024 + 2 * plane) : a_hi;
  __half *b_lo = split ? reinterpret_cast<__half *>(ws + 1024 + 3 * plane) : b_hi;

  // ---- operand preparation
  if (cudaMemsetAsync(amax, 0, 2 * sizeof(uint32_t), st) != cudaSuccess) return check_launch("pf_volume_build(memset)");
  const long long total = (long long)B * C * N;
  PF_REQUIRE((((uintptr_t)a->fmap1 | (uintptr_t)a->fmap2) & 15) == 0, "pf_volume_build: fmaps must be 16-byte aligned");
  const unsigned rblocks = (unsigned)((total / 4 + 255) / 256 < 592 ? (total / 4 + 255) / 256 : 592);
  absmax_kernel<<<dim3(rblocks, 2), 256, 0, st>>>(a->fmap1, a->fmap2, total, amax);
  dim3 tgrid(ceil_div(N, 32), ceil_div(C, 64), 2 * B);
  {
    static const bool pdl_prep = !(getenv("PF_VOLUME_PDL") != nullptr && getenv("PF_VOLUME_PDL")[0] == '0');
    cudaLaunchConfig_t pc = {};
    pc.gridDim = tgrid;
    pc.blockDim = dim3(256);
    pc.stream = st;
    cudaLaunchAttribute pa[1];
    pa[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pa[0].val.programmaticStreamSerializationAllowed = 1;
    pc.attrs = pa;
    pc.numAttrs = pdl_prep ? 1 : 0;
    const uint32_t *amax_c = amax;
    if (cudaLaunchKernelEx(&pc, split_transpose_kernel, a->fmap1, a->fmap2, a_hi, a_lo, b_hi, b_lo, B, C, N, amax_c, split ? 1 : 0) !=
        cudaSuccess) {
      (void)cudaGetLastError();
      split_transpose_kernel<<<tgrid, 256, 0, st>>>(a->fmap1, a->fmap2, a_hi, a_lo, b_hi, b_lo, B, C, N, amax, split ? 1 : 0);
    }
  }
  if (int e = check_launch("pf_volume_build(prep)")) return e;

  // ---- tensor maps
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int cluster = ((N / BM) % 2 == 0 && sms % 2 == 0 && getenv("PF_VOLUME_NO_CLUSTER") == nullptr) ? 2 : 1;
  CUtensorMap m_a_hi, m_a_lo, m_b_hi, m_b_lo, m_out;
  {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)B * N};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {BK, BM};
    if (int e = encode(&m_a_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a_hi, dims, strides, box, "A.hi")) return e;
    if (int e = encode(&m_a_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a_lo, dims, strides, box, "A.lo")) return e;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)w * C * 2, (cuuint64_t)h * w * C * 2};
    cuuint32_t box[4] = {BK, PATCH_W, (cuuint32_t)(PATCH_H / cluster), 1};
    if (int e = encode(&m_b_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, b_hi, dims, strides, box, "B.hi")) return e;
    if (int e = encode(&m_b_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, b_lo, dims, strides, box, "B.lo")) return e;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B * N};
    cuuint64_t strides[2] = {(cuuint64_t)w * 4, (cuuint64_t)h * w * 4};
    cuuint32_t box[3] = {32, 1, BM};
    if (int e = encode(&m_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, a->level[0], dims, strides, box, "level0")) return e;
  }

  TcParams p;
  p.B = B;
  p.C = C;
  p.N = N;
  p.h = h;
  p.w = w;
  p.num_levels = a->num_levels;
  p.tiles_m = N / BM;
  p.patches_x = w / PATCH_W;
  p.patches_y = h / PATCH_H;
  p.total_tiles = (long long)B * p.tiles_m * p.patches_x * p.patches_y;
  p.inv_sqrt_c = 1.0f / sqrtf((float)C);
  static const bool contiguous = getenv("PF_VOLUME_ORDER") != nullptr && getenv("PF_VOLUME_ORDER")[0] == '1';
  p.contiguous = contiguous;
  p.amax_bits = amax;
  p.lvl0 = a->level[0];
  p.lvl1 = a->num_levels > 1 ? a->level[1] : nullptr;
  p.lvl2 = a->num_levels > 2 ? a->level[2] : nullptr;
  p.lvl3 = a->num_levels > 3 ? a->level[3] : nullptr;

  unsigned grid = (unsigned)(p.total_tiles < sms ? p.total_tiles : sms);
  grid -= grid % cluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kTcThreads);
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = !(getenv("PF_VOLUME_PDL") != nullptr && getenv("PF_VOLUME_PDL")[0] == '0');
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  auto launch = [&](auto kern, int smem) -> cudaError_t {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cfg.dynamicSmemBytes = smem;
    return cudaLaunchKernelEx(&cfg, kern, m_a_hi, m_a_lo, m_b_hi, m_b_lo, m_out, p);
  };
  cudaError_t err;
  // PF_VOLUME_2SM=1: one tcgen05.mma.cta_group::2 per CTA pair (each CTA stages only its half of the patch, double-buffered level-0
  // staging) instead of TMA multicast + two cta_group::1 MMAs.  Parity-green and measured equal (129.0 vs 131.1 us per view, r03):
  // the kernel is bound by the per-group chain (MMA, then an epilogue paced by HBM writes), not by operand delivery — DESIGN §4.1.
  static const bool two_sm = getenv("PF_VOLUME_2SM") != nullptr && getenv("PF_VOLUME_2SM")[0] == '1';
  // PF_VOLUME_SHARE=1: both epilogue groups drain the same tile (MMA of the next tile overlaps) — parity-green, measured equal
  // to the default (131.1 vs 131.1 us, r03), so the default stays the r02 epilogue (each group drains every other tile)
  static const bool share = getenv("PF_VOLUME_SHARE") != nullptr && getenv("PF_VOLUME_SHARE")[0] == '1';
  // PF_VOLUME_STG=1: level 0 written with coalesced st.global instead of TMA stores (A/B, r03)
  static const bool stg = getenv("PF_VOLUME_STG") != nullptr && getenv("PF_VOLUME_STG")[0] == '1';
  if (cluster == 2 && !two_sm && stg) {
    err = split ? launch(volume_tc_kernel<true, 2, false, false, true>, Cfg<true>::kSmemBytes) : launch(volume_tc_kernel<false, 2, false, false, true>, Cfg<false>::kSmemBytes);
  } else if (cluster == 2 && !two_sm && share && a->num_levels == 4) {
    err = split ? launch(volume_tc_kernel<true, 2, false, true>, Cfg<true>::kSmemBytes) : launch(volume_tc_kernel<false, 2, false, true>, Cfg<false>::kSmemBytes);
  } else if (split) {
    if (cluster == 2 && two_sm) err = launch(volume_tc_kernel<true, 2, true>, Cfg<true, true>::kSmemBytes);
    else if (cluster == 2) err = launch(volume_tc_kernel<true, 2>, Cfg<true>::kSmemBytes);
    else err = launch(volume_tc_kernel<true, 1>, Cfg<true>::kSmemBytes);
  } else {
    if (cluster == 2 && two_sm) err = launch(volume_tc_kernel<false, 2, true>, Cfg<false, true>::kSmemBytes);
    else if (cluster == 2) err = launch(volume_tc_kernel<false, 2>, Cfg<false>::kSmemBytes);
    else err = launch(volume_tc_kernel<false, 1>, Cfg<false>::kSmemBytes);
  }
  if (err != cudaSuccess) {
    set_error("pf_volume_build(tcgen05): launch failed: %s", cudaGetErrorString(err));
    return 2;
  }
  return check_launch("pf_volume_build(tcgen05)");
}

}  // namespace pf
