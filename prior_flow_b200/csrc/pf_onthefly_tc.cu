// (c) On-the-fly lookup with the dot products on tensor cores.
//
// The r01 kernel (pf_onthefly.cu) evaluates, per query, the correlation "plane" on the bounding box of its window with one
// warp per target pixel — every query re-reads ~100 feature vectors of 1 KB from L2, 6.5 GB per DCCL call at 64x128, and the
// kernel runs at the L2's bandwidth (725 us).  Neighbouring queries look at neighbouring targets, so the same work is a small
// dense contraction per TILE of queries:
//   otf_box_kernel    per query: the sampler's coordinate chain (bit-exact, shared with the blend below) and the integer
//                     bounding box of its taps; atomically merged into the box of its 8 x 16 query tile
//   otf_alloc_kernel  one CTA: a prefix sum over the tiles gives every box its place in a pool of "local planes" (128 queries x box
//                     pixels of fp32, in segments of 128 queries x 32 pixels = 16 KiB; the pool is O(N): 32 KiB per query and
//                     view by default) and cuts the boxes into work items of up to four MMA passes; boxes wider than 256
//                     columns or beyond the pool go on the work list of the CUDA-core kernel
//   otf_dots_kernel   persistent, one CTA per SM over the work items: D[128 queries, 256 box pixels] = F1[tile] . F2_l[box]^T on
//                     tcgen05 — both operands are TMA boxes of pre-split fp16 hi/lo K-major planes ([B, h, w, C], the
//                     channels-last convention of `alt_cuda_corr`), three products into fp32 TMEM like the volume kernel,
//                     written to the tile's local planes with one 16 KiB bulk store per segment
//   otf_blend_kernel  per query: the taps blend from the local planes in ATen's order (zeros for tiles no tap touches)
//   otf_fallback_kernel  persistent over its work list: the r01 CUDA-core path, query by query
// The ERP seam: the sampler wraps x, so a window across the seam touches columns at both ends of the plane.  Boxes are therefore
// kept in two "unwrapped" column numberings — u0(x) = x and u1(x) = x + W for x < W/2 — a window narrower than W/2 is contiguous
// in at least one of them, and the target planes are stored twice side by side ([B, Hl, 2 Wl, C]) so that a box in either
// numbering is one TMA box.
// Outputs, layouts and tolerances are those of pf_lookup_onthefly (1e-5 of max|ref| against the materialised lookup).
#include <climits>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "pf_tc.cuh"

namespace pf {

constexpr int OT_TH = 8, OT_TW = 16;                 // query tile: 8 rows x 16 columns = 128 queries = TMEM lanes
constexpr int OT_SEG = 128 * 32;                     // floats of a pool segment: 128 queries x 32 box pixels
constexpr int OT_BK = 64;
constexpr int OT_APLANE = 128 * OT_BK * 2;           // 16 KiB
constexpr int OT_BPLANE = 256 * OT_BK * 2;           // 32 KiB
constexpr int OT_STAGE = 2 * (OT_APLANE + OT_BPLANE);   // hi + lo: 96 KiB
constexpr int OT_STAGES = 2;
constexpr int OT_THREADS = 192;                      // TMA warp, MMA warp, 4 epilogue warps
constexpr int OT_OUT = 128 * 32 * 4;                 // 16 KiB staging for one box-row segment of all queries, two of them
constexpr int OT_SMEM = OT_STAGES * OT_STAGE + 2 * OT_OUT + 256 + 1024;

constexpr int kBlendThreads = 256;
constexpr int kBlendQueries = 16;    // one row of a query tile
constexpr int kBoxDots = 324;    // per-query boxes up to 4 dots per tap are evaluated as a plane, larger ones tap by tap
constexpr int kMaxBox = kBoxDots;
constexpr int kMaxTaps = 81;
constexpr int kMaxVec = 4;

struct OtfTcParams {
  int B, N, h, w, C, L, div_mode;
  const float *coords;
  const float *f1[2];                      // [B, N, C] fp32 (CUDA-core fallback)
  const float *f2[2][PF_MAX_LEVELS];       // [B, Hl*Wl, C] fp32
  Axis axW[PF_MAX_LEVELS], axH[PF_MAX_LEVELS], ax_gw, ax_gh;
  const float *grid_w2c;
  long long grid_bs;
  float scale;                             // 1 / sqrt(C)
  int tiles_x, tiles, T;                   // query tiles per row / per image; T = views * L * B * tiles entries e = ((view * L + lvl) * B + b) * tiles + tile
  int *box_lo, *box_hi;                    // [T][4]: (u0, u1, y, -) min / max of the tile's taps
  int *ctr;                                // [0] CUDA-core tiles, [1] their cursor, [2] work items of the dots kernel, [3] pool segments in use
  int *alloc;                              // [T] first pool segment of the tile's local planes; -1: CUDA-core path, -2: no tap touches the plane
  int *fb_list;                            // [T] entries of the CUDA-core path
  int2 *items;                             // [max_items] (entry, first MMA pass)
  float *pool;                             // [pool_segs][128][32]: row r of a segment is query r of the tile, 16-byte groups XOR-swizzled by r & 7
  int pool_segs, max_items;
  int item_chunks;                         // MMA passes (256 box pixels each) per work item
  const uint32_t *amax[2];                 // per view: absmax bits of {f1, f2} (split scales of the fp16 planes)
  float *out_own, *out_raw;
  float2 *tapxy;                           // [L][B][N][81] sample coordinates of the other view's taps: written by the box kernel,
                                           // read by the blend (the chain through the rotation grid is evaluated once)
  int own_cl;                              // own view channels-last [B, N, L*81] like out_raw (for pf_dccl_conv), no rotate pass
};

__device__ __forceinline__ int tile_of(const OtfTcParams &p, int n) {
  const int y = n / p.w, x = n - y * p.w;
  return (y / OT_TH) * p.tiles_x + x / OT_TW;
}
__device__ __forceinline__ int entry_of(const OtfTcParams &p, int branch, int lvl, int b, int tile) {
  return ((branch * p.L + lvl) * p.B + b) * p.tiles + tile;
}
struct Entry {
  int branch, lvl, b, tile;
};
__device__ __forceinline__ Entry decode_entry(const OtfTcParams &p, int e) {
  Entry r;
  r.tile = e % p.tiles;
  e /= p.tiles;
  r.b = e % p.B;
  e /= p.B;
  r.lvl = e % p.L;
  r.branch = e / p.L;
  return r;
}

__device__ __forceinline__ float ot_split_scale(uint32_t amax_bits) {
  int e = (int)((amax_bits >> 23) & 0xff) - 127;
  if ((amax_bits & 0x7fffffffu) == 0u) e = 13;
  int se = 13 - e;
  se = se < -100 ? -100 : (se > 100 ? 100 : se);
  return __uint_as_float((uint32_t)(se + 127) << 23);
}

// The sampler's coordinates (ix, iy) of all 81 taps of the CTA's 8 queries, same arithmetic as pf_onthefly.cu / lookup_kernel.  The
// chains separate by axis — x depends on the tap's column offset only, y on its row offset — so step 1 evaluates 18 chains per
// query (own view: the final coordinates; other view: the coordinates into the rotation grid) and step 2 visits the 81 taps
// (other view: the two grid samples and the final chains).  `use(q, t, ix, iy)` consumes a tap.  Contains one __syncthreads.
// step 1 for item i = 18 q + j of the CTA: j < 9 the x chain of column offset j - 4, else the y chain of row offset j - 13
__device__ __forceinline__ float tap_axis_coord(const OtfTcParams &p, int branch, int lvl, int b, int n, int j) {
  const bool xaxis = j < 9;
  const float c = __fmul_rn(__ldg(p.coords + ((long long)b * 2 + (xaxis ? 0 : 1)) * p.N + n), 1.0f / (float)(1 << lvl));
  const float pv = __fadd_rn(c, (float)((xaxis ? j : j - 9) - 4));
  if (branch) return xaxis ? to_sample_coord(remainder_pos(pv, p.ax_gw.size), p.ax_gw, p.div_mode) : to_sample_coord(pv, p.ax_gh, p.div_mode);
  return xaxis ? to_sample_coord(remainder_pos(pv, p.axW[lvl].size), p.axW[lvl], p.div_mode) : to_sample_coord(pv, p.axH[lvl], p.div_mode);
}
template <class F>
__device__ __forceinline__ void cta_tap_coords(const OtfTcParams &p, int branch, int lvl, int b, int n0, float *s_axis /*[16][18]*/, F use) {
  const Axis axW = p.axW[lvl], axH = p.axH[lvl];
  for (int i = threadIdx.x; i < kBlendQueries * 18; i += kBlendThreads) {
    const int q = i / 18, j = i - q * 18, n = n0 + q;
    if (n < p.N) s_axis[i] = tap_axis_coord(p, branch, lvl, b, n, j);
  }
  __syncthreads();
  const float *gridx = p.grid_w2c ? p.grid_w2c + (long long)b * p.grid_bs : nullptr, *gridy = gridx ? gridx + p.N : nullptr;
  for (int i = threadIdx.x; i < kBlendQueries * kMaxTaps; i += kBlendThreads) {
    const int q = i / kMaxTaps, t = i - q * kMaxTaps, aa = t / 9, bb = t - aa * 9;
    if (n0 + q >= p.N) continue;
    float ix = s_axis[q * 18 + aa], iy = s_axis[q * 18 + 9 + bb];
    if (branch) {
      const Taps tg = make_taps(ix, iy);
      const float sx = blend_zeros(gridx, p.h, p.w, tg), sy = blend_zeros(gridy, p.h, p.w, tg);
      ix = to_sample_coord(remainder_pos(sx, axW.size), axW, p.div_mode);
      iy = to_sample_coord(sy, axH, p.div_mode);
    }
    use(q, t, ix, iy);
  }
}

// ------------------------------------------------------------------------------------------------ boxes
// s_box = {u0 lo, u0 hi, u1 lo, u1 hi, y lo, y hi} of the corners a tap touches (clamped to the plane like blend_zeros)
__device__ __forceinline__ void box_reset(int *s_box) {
  s_box[0] = s_box[2] = s_box[4] = INT_MAX;
  s_box[1] = s_box[3] = s_box[5] = INT_MIN;
}
__device__ __forceinline__ int unwrap1(int x, int Wl) { return x < (Wl >> 1) ? x + Wl : x; }
__device__ __forceinline__ void box_add_tap(int *s_box, int x0, int y0, int Wl, int Hl) {
  if (x0 + 1 >= 0 && x0 < Wl && y0 + 1 >= 0 && y0 < Hl) {  // tap touches the plane
    const int xa = max(x0, 0), xb = min(x0 + 1, Wl - 1);
    atomicMin(&s_box[0], xa);
    atomicMax(&s_box[1], xb);
    const int ua = unwrap1(xa, Wl), ub = unwrap1(xb, Wl);
    atomicMin(&s_box[2], min(ua, ub));
    atomicMax(&s_box[3], max(ua, ub));
    atomicMin(&s_box[4], max(y0, 0));
    atomicMax(&s_box[5], min(y0 + 1, Hl - 1));
  }
}
__global__ void __launch_bounds__(kBlendThreads) otf_box_kernel(const OtfTcParams p) {
  __shared__ int s_box[6];
  __shared__ float s_axis[kBlendQueries * 18];
  const int lvl = blockIdx.y % p.L, branch = blockIdx.y / p.L, b = blockIdx.z;
  const int Hl = p.h >> lvl, Wl = p.w >> lvl;
  const int n0 = blockIdx.x * kBlendQueries;
  // the 16 queries of a CTA are one row of a query tile
  if (threadIdx.x == 0) box_reset(s_box);
  if (branch == 0) {
    // own view: the taps are a product of 9 columns and 9 rows, so the box is the columns any query touches x the rows any
    // query touches (a superset of the taps' union when some query misses the plane in one axis only: harmless)
    __syncthreads();
    for (int i = threadIdx.x; i < kBlendQueries * 18; i += kBlendThreads) {
      const int q = i / 18, j = i - q * 18, n = n0 + q;
      if (n >= p.N) continue;
      const int c0 = (int)floorf(tap_axis_coord(p, 0, lvl, b, n, j));
      if (j < 9) {
        if (c0 + 1 >= 0 && c0 < Wl) {
          const int xa = max(c0, 0), xb = min(c0 + 1, Wl - 1), ua = unwrap1(xa, Wl), ub = unwrap1(xb, Wl);
          atomicMin(&s_box[0], xa), atomicMax(&s_box[1], xb), atomicMin(&s_box[2], min(ua, ub)), atomicMax(&s_box[3], max(ua, ub));
        }
      } else if (c0 + 1 >= 0 && c0 < Hl) {
        atomicMin(&s_box[4], max(c0, 0)), atomicMax(&s_box[5], min(c0 + 1, Hl - 1));
      }
    }
    __syncthreads();
    if (s_box[0] > s_box[1]) return;      // no column touched: leave the tile's box as it is (uniform)
  } else {
    float2 *xy = p.tapxy + (((long long)lvl * p.B + b) * p.N + n0) * kMaxTaps;
    // a thread folds the corners of its own taps into registers; one warp reduction and six shared atomics per warp at the end
    int v[6] = {INT_MAX, INT_MIN, INT_MAX, INT_MIN, INT_MAX, INT_MIN};
    cta_tap_coords(p, branch, lvl, b, n0, s_axis, [&](int q, int t, float ix, float iy) {
      xy[q * kMaxTaps + t] = make_float2(ix, iy);
      const int x0 = (int)floorf(ix), y0 = (int)floorf(iy);
      if (x0 + 1 >= 0 && x0 < Wl && y0 + 1 >= 0 && y0 < Hl) {
        const int xa = max(x0, 0), xb = min(x0 + 1, Wl - 1), ua = unwrap1(xa, Wl), ub = unwrap1(xb, Wl);
        v[0] = min(v[0], xa), v[1] = max(v[1], xb), v[2] = min(v[2], min(ua, ub)), v[3] = max(v[3], max(ua, ub));
        v[4] = min(v[4], max(y0, 0)), v[5] = max(v[5], min(y0 + 1, Hl - 1));
      }
    });
#pragma unroll
    for (int i = 0; i < 6; i += 2) v[i] = __reduce_min_sync(0xffffffffu, v[i]), v[i + 1] = __reduce_max_sync(0xffffffffu, v[i + 1]);
    if ((threadIdx.x & 31) == 0 && v[4] <= v[5]) {
#pragma unroll
      for (int i = 0; i < 6; i += 2) atomicMin(&s_box[i], v[i]), atomicMax(&s_box[i + 1], v[i + 1]);
    }
    __syncthreads();
  }
  if (threadIdx.x < 3 && s_box[4] <= s_box[5]) {
    const long long e = (long long)entry_of(p, branch, lvl, b, tile_of(p, n0)) * 4 + threadIdx.x;
    atomicMin(p.box_lo + e, s_box[2 * threadIdx.x]);
    atomicMax(p.box_hi + e, s_box[2 * threadIdx.x + 1]);
  }
}

// The tile's box: which numbering (mode), origin, rows, and the pitch of its local planes (32 ... 256 columns, 1 ... 8 segments).
struct TileBox {
  int mode, X0, Y0, rows, pitch;
  bool empty;      // no tap of the tile touches the plane: the outputs are zero
  __device__ __forceinline__ int spr() const { return pitch >> 5; }                 // segments per box row
  __device__ __forceinline__ int rpc() const { return 256 / pitch; }                // box rows per MMA pass
  __device__ __forceinline__ int nseg() const { return rows * (pitch >> 5); }
  __device__ __forceinline__ int nchunks() const { return (rows + rpc() - 1) / rpc(); }
};
// false: empty, or wider than one MMA pass (256 columns)
__device__ __forceinline__ bool read_box(const OtfTcParams &p, int e, TileBox &tb) {
  const int4 lo = *reinterpret_cast<const int4 *>(p.box_lo + 4ll * e), hi = *reinterpret_cast<const int4 *>(p.box_hi + 4ll * e);
  tb.empty = lo.z > hi.z || lo.x > hi.x;
  if (tb.empty) return false;
  const int w0 = hi.x - lo.x + 1, w1 = hi.y - lo.y + 1;
  tb.mode = w1 < w0 ? 1 : 0;
  const int bw = tb.mode ? w1 : w0;
  tb.X0 = tb.mode ? lo.y : lo.x;
  tb.Y0 = lo.z;
  tb.rows = hi.z - lo.z + 1;
  tb.pitch = bw <= 32 ? 32 : (bw <= 64 ? 64 : (bw <= 128 ? 128 : 256));
  return bw <= 256;
}

// ------------------------------------------------------------------------------------------------ pool allocation, work items
__device__ __forceinline__ int block_exclusive_scan(int v, int *s_warp /*[32]*/, int &total) {   // 1024 threads
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  const int base = warp ? s_warp[warp - 1] : 0;
  total = s_warp[31];
  __syncthreads();       // s_warp is reused by the next scan
  return base + x - v;
}

// counters to zero, boxes to (+large, -large): one launch in front of the box kernel
__global__ void __launch_bounds__(256) otf_init_kernel(const OtfTcParams p) {
  const int n = 4 * p.T;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p.box_lo[i] = INT_MAX, p.box_hi[i] = INT_MIN;
  if (blockIdx.x == 0 && threadIdx.x < 16) p.ctr[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(1024) otf_alloc_kernel(const OtfTcParams p) {
  __shared__ int s_scan[32];
  const int per = (p.T + 1023) / 1024, e0 = threadIdx.x * per, e1 = min(e0 + per, p.T);
  // 1. pool segments of the boxes one MMA pass can cover; a box beyond the pool's end goes to the CUDA-core path
  int segs = 0;
  for (int e = e0; e < e1; ++e) {
    TileBox tb;
    if (read_box(p, e, tb)) segs += tb.nseg();
  }
  int total, total_segs;
  int off = block_exclusive_scan(segs, s_scan, total_segs);
  int items = 0;
  for (int e = e0; e < e1; ++e) {
    TileBox tb;
    const bool shape_ok = read_box(p, e, tb);
    int a = tb.empty ? -2 : -1;
    if (shape_ok) {
      if (off + tb.nseg() <= p.pool_segs) a = off, items += (tb.nchunks() + p.item_chunks - 1) / p.item_chunks;
      off += tb.nseg();
    }
    p.alloc[e] = a;
    if (a == -1) p.fb_list[atomicAdd(p.ctr, 1)] = e;
  }
  // 2. work items of the dots kernel
  int it = block_exclusive_scan(items, s_scan, total);
  if (threadIdx.x == 0) p.ctr[2] = min(total, p.max_items), p.ctr[3] = total_segs;      // [3]: what the boxes asked for (diagnostics)
  for (int e = e0; e < e1; ++e) {
    if (p.alloc[e] < 0) continue;
    TileBox tb;
    read_box(p, e, tb);
    for (int c = 0; c < tb.nchunks(); c += p.item_chunks, ++it)
      if (it < p.max_items) p.items[it] = make_int2(e, c);
  }
}

// ------------------------------------------------------------------------------------------------ dots (tcgen05)
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

struct OtfViewMaps {
  CUtensorMap f1_hi, f1_lo;                                    // [C, w, h, B] fp16, box {64, 16, 8, 1}
  CUtensorMap f2_hi[4][PF_MAX_LEVELS], f2_lo[4][PF_MAX_LEVELS];   // [C, 2 Wl, Hl, B] fp16, box {64, pitch, 256 / pitch, 1}, pitch 32 ... 256
};
struct OtfMaps {
  OtfViewMaps view[2];     // ~8.5 KiB of kernel parameters (CUDA 12.1+: up to 32 KiB)
};
__device__ __forceinline__ int pitch_index(int pitch) { return pitch == 32 ? 0 : (pitch == 64 ? 1 : (pitch == 128 ? 2 : 3)); }

__global__ void __launch_bounds__(OT_THREADS, 1) otf_dots_kernel(const __grid_constant__ OtfMaps all_maps, const OtfTcParams p) {
  const int nitems = p.ctr[2];
  if ((int)blockIdx.x >= nitems) return;

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *out_stage = smem + OT_STAGES * OT_STAGE;
  uint64_t *bars = reinterpret_cast<uint64_t *>(out_stage + 2 * OT_OUT);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + OT_STAGES);
  const uint32_t bar_tfull = smem_u32(bars + 2 * OT_STAGES), bar_tempty = smem_u32(bars + 2 * OT_STAGES + 2);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * OT_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = p.C / OT_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < OT_STAGES; ++s) mbar_init(bar_full + 8 * s, 1), mbar_init(bar_empty + 8 * s, 1);
    for (int a = 0; a < 2; ++a) mbar_init(bar_tfull + 8 * a, 1), mbar_init(bar_tempty + 8 * a, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // every role walks the same static item sequence; the smem / TMEM pipelines run on across items
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
        const int2 item = p.items[it];
        const Entry en = decode_entry(p, item.x);
        TileBox tb;
        read_box(p, item.x, tb);
        const OtfViewMaps &maps = all_maps.view[en.branch];
        const int pi = pitch_index(tb.pitch), rpc = tb.rpc(), c1 = min(item.y + p.item_chunks, tb.nchunks());
        const int ty = en.tile / p.tiles_x, tx = en.tile - ty * p.tiles_x;
        for (int c = item.y; c < c1; ++c)
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t full = bar_full + 8 * stage;
            mbar_arrive_expect_tx(full, (uint32_t)OT_STAGE);
            const uint32_t sbase = smem_u32(smem + stage * OT_STAGE);
            // layout of a stage: [A.hi 16K | B.hi 32K | A.lo 16K | B.lo 32K]
            tma_load_4d(sbase, &maps.f1_hi, full, kb * OT_BK, tx * OT_TW, ty * OT_TH, en.b);
            tma_load_4d(sbase + OT_APLANE, &maps.f2_hi[pi][en.lvl], full, kb * OT_BK, tb.X0, tb.Y0 + c * rpc, en.b);
            tma_load_4d(sbase + OT_APLANE + OT_BPLANE, &maps.f1_lo, full, kb * OT_BK, tx * OT_TW, ty * OT_TH, en.b);
            tma_load_4d(sbase + 2 * OT_APLANE + OT_BPLANE, &maps.f2_lo[pi][en.lvl], full, kb * OT_BK, tb.X0, tb.Y0 + c * rpc, en.b);
            if (++stage == OT_STAGES) stage = 0, phase ^= 1;
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, acc = 0, acc_phase = 0;
      for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
        const int2 item = p.items[it];
        TileBox tb;
        read_box(p, item.x, tb);
        const int rpc = tb.rpc(), c1 = min(item.y + p.item_chunks, tb.nchunks());
        for (int c = item.y; c < c1; ++c) {
          mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * 256;
          const int ncols = min(256, (tb.rows - c * rpc) * tb.pitch);       // the last pass needs only the box rows that are left
          const uint32_t idesc = umma_idesc_f16(128, ncols);
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t sbase = smem_u32(smem + stage * OT_STAGE);
            const uint64_t a_hi = make_smem_desc(sbase), b_hi = make_smem_desc(sbase + OT_APLANE);
            const uint64_t a_lo = make_smem_desc(sbase + OT_APLANE + OT_BPLANE), b_lo = make_smem_desc(sbase + 2 * OT_APLANE + OT_BPLANE);
#pragma unroll
            for (int k = 0; k < OT_BK / 16; ++k) umma_f16(tmem_d, a_lo + 2 * k, b_hi + 2 * k, idesc, (kb | k) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < OT_BK / 16; ++k) umma_f16(tmem_d, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
#pragma unroll
            for (int k = 0; k < OT_BK / 16; ++k) umma_f16(tmem_d, a_hi + 2 * k, b_hi + 2 * k, idesc, 1u);
            umma_commit(bar_empty + 8 * stage);
            if (++stage == OT_STAGES) stage = 0, phase ^= 1;
          }
          umma_commit(bar_tfull + 8 * acc);
          if ((acc ^= 1) == 0) acc_phase ^= 1;
        }
      }
    }
  } else {
    // epilogue: 4 warps, TMEM lane = query of the tile (rx fastest), columns = box pixels of the pass in segments of 32.  Segment g
    // of pass c is pool segment 8 c + g of the tile (box row (8 c + g) / spr, columns 32 ((8 c + g) % spr) ...): staged as
    // [128 queries][32 floats] with the 16-byte groups XOR-swizzled (conflict-free float4 writes; the blend un-swizzles) and
    // written with one 16 KiB bulk store; two staging buffers, a store drains while the next segment is staged
    const int quarter = warp & 3, row = quarter * 32 + lane;
    const bool leader = threadIdx.x == 64;
    uint32_t acc = 0, acc_phase = 0, buf = 0;
    for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
      const int2 item = p.items[it];
      const Entry en = decode_entry(p, item.x);
      TileBox tb;
      read_box(p, item.x, tb);
      const int rpc = tb.rpc(), spr = tb.spr(), c1 = min(item.y + p.item_chunks, tb.nchunks());
      const float scale = p.scale / (ot_split_scale(p.amax[en.branch][0]) * ot_split_scale(p.amax[en.branch][1]));
      float *planes = p.pool + (long long)p.alloc[item.x] * OT_SEG;
      for (int c = item.y; c < c1; ++c) {
        mbar_wait(bar_tfull + 8 * acc, acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * 256 + ((uint32_t)(quarter * 32) << 16);
        const int nseg = min(8, (tb.rows - c * rpc) * spr);      // segments of this pass that hold box rows
        uint32_t un[32];
        tmem_ld32(taddr, un);
#pragma unroll 1
        for (int g = 0; g < nseg; ++g) {
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(un[j]) * scale;
          if (g + 1 < nseg) {
            tmem_ld32(taddr + (g + 1) * 32, un);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
          }
          if (leader) tma_store_wait_read1();      // the store issued from this buffer two segments ago has read it
          asm volatile("bar.sync 1, 128;" ::: "memory");
          uint8_t *stage = out_stage + buf * OT_OUT;
          {
            uint8_t *r0 = stage + row * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4 *>(r0 + ((j ^ (row & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          fence_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (leader) {
            bulk_store(planes + (long long)(8 * c + g) * OT_SEG, smem_u32(stage), OT_OUT);
            tma_store_commit();
          }
          buf ^= 1;
        }
        if ((acc ^= 1) == 0) acc_phase ^= 1;
      }
    }
    if (leader) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ blend
// Four dot products of the query vector at once: all loads in flight together, then one 6-shuffle reduction for the four sums
// (fold 4 -> 2 -> 1 values per lane over xor 16 / 8, then a 3-step butterfly).  Every lane of group g = lane >> 3 ends with sum g.
// Null pointers give 0.  The per-pixel latency (L2 round trip + reduction) is what bounds the CUDA-core path.
__device__ __forceinline__ float ot_dot4(const float *const (&ptr)[4], const float4 (&q)[kMaxVec], int nvec, int lane) {
  float4 v[4][kMaxVec];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j)
      v[c][j] = (j < nvec && ptr[c]) ? __ldg(reinterpret_cast<const float4 *>(ptr[c]) + j * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  float a[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      if (j < nvec) {
        acc = fmaf(q[j].x, v[c][j].x, acc);
        acc = fmaf(q[j].y, v[c][j].y, acc);
        acc = fmaf(q[j].z, v[c][j].z, acc);
        acc = fmaf(q[j].w, v[c][j].w, acc);
      }
    }
    a[c] = acc;
  }
  const bool up16 = lane & 16, up8 = lane & 8;
  // lanes 0-15 keep sums {0, 1}, lanes 16-31 keep {2, 3}
  const float s0 = (up16 ? a[2] : a[0]) + __shfl_xor_sync(0xffffffffu, up16 ? a[0] : a[2], 16);
  const float s1 = (up16 ? a[3] : a[1]) + __shfl_xor_sync(0xffffffffu, up16 ? a[1] : a[3], 16);
  // within each half, lanes with bit 3 clear keep the first, set keep the second
  float r = (up8 ? s1 : s0) + __shfl_xor_sync(0xffffffffu, up8 ? s0 : s1, 8);
  r += __shfl_xor_sync(0xffffffffu, r, 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

// s_out [tap][query] -> own view: [K2][8 queries] = 32-byte row segments of the [B, L*K2, N] output; other view: channels-last
// [B, N, L*K2] pre-rotation map (contiguous per query), see pf_lookup.cu
__device__ __forceinline__ void write_taps(const OtfTcParams &p, const float (*s_out)[kBlendQueries + 1], int branch, int lvl, int b, int n0) {
  constexpr int K2 = kMaxTaps;      // (N is a multiple of 16: every query of the CTA exists)
  if (branch == 0 && !p.own_cl) {
    float *out = p.out_own + ((long long)b * p.L + lvl) * K2 * (long long)p.N + n0;
    for (int i = threadIdx.x; i < K2 * kBlendQueries; i += kBlendThreads) {
      const int ch = i / kBlendQueries, q = i - ch * kBlendQueries;
      out[(unsigned)ch * (unsigned)p.N + q] = s_out[ch][q];
    }
  } else {
    float *out = (branch ? p.out_raw : p.out_own) + (((long long)b * p.N + n0) * p.L + lvl) * K2;
    const int qstride = p.L * K2;
    for (int i = threadIdx.x; i < K2 * kBlendQueries; i += kBlendThreads) {
      const int q = i / K2, ch = i - q * K2;
      out[q * qstride + ch] = s_out[ch][q];
    }
  }
}

// Tiles on the tensor-core path: the tile's dots are in its local planes (otf_dots_kernel); every tap reads its four corners from
// there.  All taps of the CTA's 16 queries are in flight together.  Tiles no tap touches: zeros.
__global__ void __launch_bounds__(kBlendThreads) otf_blend_kernel(const OtfTcParams p) {
  __shared__ float s_out[kMaxTaps][kBlendQueries + 1];
  const int lvl = blockIdx.y % p.L, branch = blockIdx.y / p.L, b = blockIdx.z;
  const int Hl = p.h >> lvl, Wl = p.w >> lvl;
  const int n0 = blockIdx.x * kBlendQueries;
  const int e = entry_of(p, branch, lvl, b, tile_of(p, n0));
  const int first = p.alloc[e];                       // block-uniform
  if (first == -1) return;                            // otf_fallback_kernel's
  if (first == -2) {
    for (int i = threadIdx.x; i < kMaxTaps * (kBlendQueries + 1); i += kBlendThreads) (&s_out[0][0])[i] = 0.f;
    __syncthreads();
    write_taps(p, s_out, branch, lvl, b, n0);
    return;
  }
  TileBox tb;
  read_box(p, e, tb);
  const int spr = tb.spr();
  const int qrow0 = ((n0 / p.w) % OT_TH) * OT_TW;     // row of query n0 in its tile's segments
  const float *planes = p.pool + (long long)first * OT_SEG;
  // element (box row y, column c) of query r (row of the tile's segments): segment y * spr + c / 32, float
  // r * 32 + (((c % 32) / 4) ^ (r & 7)) * 4 + c % 4 of it.  Offsets inside a tile's planes fit an int.
  const int down = spr * OT_SEG;
  auto col_off = [&](int c, int r) { return (c >> 5) * OT_SEG + ((((c & 31) >> 2) ^ (r & 7)) << 2) + (c & 3); };
  // column of the west corner in the box's numbering (mode 1: unwrapped); the east corner is the next column — a tap
  // whose corners straddle the numbering's cut would have made the box as wide as the plane, and then mode 0 is chosen
  auto west_col = [&](int x0, bool xin0, bool xin1) {
    return ((tb.mode && xin0) ? unwrap1(x0, Wl) : ((tb.mode && xin1) ? unwrap1(x0 + 1, Wl) - 1 : x0)) - tb.X0;
  };
  if (branch) {
    // other view: the coordinates were stashed by otf_box_kernel (each tap went through the rotation grid)
    const float2 *xy = p.tapxy + (((long long)lvl * p.B + b) * p.N + n0) * kMaxTaps;
    for (int i = threadIdx.x; i < kBlendQueries * kMaxTaps; i += kBlendThreads) {
      const int q = i / kMaxTaps, t = i - q * kMaxTaps;
      const float2 c = __ldg(xy + i);
      const Taps tp = make_taps(c.x, c.y);
      const bool xin0 = (unsigned)tp.x0 < (unsigned)Wl, xin1 = (unsigned)(tp.x0 + 1) < (unsigned)Wl;
      const bool yin0 = (unsigned)tp.y0 < (unsigned)Hl, yin1 = (unsigned)(tp.y0 + 1) < (unsigned)Hl;
      const int r = qrow0 + q, cw = west_col(tp.x0, xin0, xin1);
      const float *qn = planes + (tp.y0 - tb.Y0) * down + r * 32;
      const int off_w = col_off(cw, r), off_e = col_off(cw + 1, r);
      const float v_nw = (yin0 && xin0) ? __ldg(qn + off_w) : 0.f;
      const float v_ne = (yin0 && xin1) ? __ldg(qn + off_e) : 0.f;
      const float v_sw = (yin1 && xin0) ? __ldg(qn + down + off_w) : 0.f;
      const float v_se = (yin1 && xin1) ? __ldg(qn + down + off_e) : 0.f;
      float acc = __fmul_rn(v_nw, tp.nw);
      acc = __fmaf_rn(v_ne, tp.ne, acc);
      acc = __fmaf_rn(v_sw, tp.sw, acc);
      acc = __fmaf_rn(v_se, tp.se, acc);
      s_out[t][q] = acc;
    }
  } else {
    // own view: the 81 taps of a query are 9 columns x 9 rows, so the column half (corner offsets, dx weights) and the row half of
    // make_taps / the corner addresses are evaluated 9 + 9 times per query and the taps only combine them (same roundings:
    // nw = dxe * dys, ne = dxw * dys, sw = dxe * dyn, se = dxw * dyn as in make_taps)
    __shared__ int4 s_half[kBlendQueries * 18];      // (offset of corner 0 or -1, offset of corner 1 or -1, weight toward 0, toward 1)
    for (int i = threadIdx.x; i < kBlendQueries * 18; i += kBlendThreads) {
      const int q = i / 18, j = i - q * 18, r = qrow0 + q;
      const float v = tap_axis_coord(p, 0, lvl, b, n0 + q, j);
      const float f0 = floorf(v), f1 = __fadd_rn(f0, 1.f);
      const float w0 = __fsub_rn(f1, v), w1 = __fsub_rn(v, f0);      // dxe / dys, dxw / dyn
      const int i0 = (int)f0;
      int o0, o1;
      if (j < 9) {
        const bool in0 = (unsigned)i0 < (unsigned)Wl, in1 = (unsigned)(i0 + 1) < (unsigned)Wl;
        const int cw = west_col(i0, in0, in1);
        o0 = in0 ? col_off(cw, r) + r * 32 : -1, o1 = in1 ? col_off(cw + 1, r) + r * 32 : -1;
      } else {
        const bool in0 = (unsigned)i0 < (unsigned)Hl, in1 = (unsigned)(i0 + 1) < (unsigned)Hl;
        o0 = in0 ? (i0 - tb.Y0) * down : -1, o1 = in1 ? (i0 + 1 - tb.Y0) * down : -1;
      }
      s_half[i] = make_int4(o0, o1, __float_as_int(w0), __float_as_int(w1));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kBlendQueries * kMaxTaps; i += kBlendThreads) {
      const int q = i / kMaxTaps, t = i - q * kMaxTaps, aa = t / 9, bb = t - aa * 9;
      const int4 X = s_half[q * 18 + aa], Y = s_half[q * 18 + 9 + bb];
      const float dxe = __int_as_float(X.z), dxw = __int_as_float(X.w), dys = __int_as_float(Y.z), dyn = __int_as_float(Y.w);
      const float v_nw = (Y.x >= 0 && X.x >= 0) ? __ldg(planes + Y.x + X.x) : 0.f;
      const float v_ne = (Y.x >= 0 && X.y >= 0) ? __ldg(planes + Y.x + X.y) : 0.f;
      const float v_sw = (Y.y >= 0 && X.x >= 0) ? __ldg(planes + Y.y + X.x) : 0.f;
      const float v_se = (Y.y >= 0 && X.y >= 0) ? __ldg(planes + Y.y + X.y) : 0.f;
      float acc = __fmul_rn(v_nw, __fmul_rn(dxe, dys));
      acc = __fmaf_rn(v_ne, __fmul_rn(dxw, dys), acc);
      acc = __fmaf_rn(v_sw, __fmul_rn(dxe, dyn), acc);
      acc = __fmaf_rn(v_se, __fmul_rn(dxw, dyn), acc);
      s_out[t][q] = acc;
    }
  }
  __syncthreads();
  write_taps(p, s_out, branch, lvl, b, n0);
}

// Tiles whose box is too wide or beyond the pool (the work list otf_alloc_kernel wrote): the CUDA-core path of the r01 kernel, query
// by query — the plane on the query's own box if that is fewer dot products than four per tap, else tap by tap.  Persistent: a
// CTA takes (tile, row of 16 queries) items off the list until it is empty; these are the long items of the call.
__global__ void __launch_bounds__(kBlendThreads) otf_fallback_kernel(const OtfTcParams p) {
  __shared__ float s_ix[kBlendQueries * kMaxTaps], s_iy[kBlendQueries * kMaxTaps];
  __shared__ float s_axis[kBlendQueries * 18];
  __shared__ float s_dots[kMaxBox];
  __shared__ float s_out[kMaxTaps][kBlendQueries + 1];
  __shared__ int s_box[kBlendQueries][6];
  __shared__ int s_item;
  constexpr int K2 = kMaxTaps;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.ctr[0] * OT_TH;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_item = atomicAdd(p.ctr + 1, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= items) return;
    const Entry en = decode_entry(p, p.fb_list[item / OT_TH]);
    const int tile = en.tile, b = en.b, lvl = en.lvl, branch = en.branch;
    const int Hl = p.h >> lvl, Wl = p.w >> lvl;
    const int n0 = ((tile / p.tiles_x) * OT_TH + item % OT_TH) * p.w + (tile % p.tiles_x) * OT_TW;
    // CUDA-core path (the r01 kernel's): per query, the plane on its own box if that is fewer dot products than four per tap
    if (threadIdx.x < kBlendQueries) box_reset(s_box[threadIdx.x]);
    __syncthreads();       // (cta_tap_coords syncs once more before any tap is consumed)
    cta_tap_coords(p, branch, lvl, b, n0, s_axis, [&](int q, int t, float ix, float iy) {
      s_ix[q * K2 + t] = ix, s_iy[q * K2 + t] = iy;
      box_add_tap(s_box[q], (int)floorf(ix), (int)floorf(iy), Wl, Hl);
    });
    __syncthreads();
    const float *f2 = p.f2[branch][lvl] + (long long)b * Hl * Wl * p.C;
    const int nvec = p.C / 128;
    for (int q = 0; q < kBlendQueries; ++q) {
      const int n = n0 + q;
      if (n >= p.N) break;
      const float *q_ix = s_ix + q * K2, *q_iy = s_iy + q * K2;
      const int x_lo = s_box[q][0], x_hi = s_box[q][1], y_lo = s_box[q][4], y_hi = s_box[q][5];      // plain columns
      const bool empty = y_lo > y_hi;
      const int bw = empty ? 0 : x_hi - x_lo + 1, bh = empty ? 0 : y_hi - y_lo + 1;
      const int area = bw * bh;
      float4 qv[kMaxVec];
      const float4 *f1v = reinterpret_cast<const float4 *>(p.f1[branch] + ((long long)b * p.N + n) * p.C);
#pragma unroll
      for (int j = 0; j < kMaxVec; ++j) qv[j] = (j < nvec) ? __ldg(f1v + j * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (area <= kBoxDots) {
        for (int pix = warp * 4; pix < area; pix += kBlendThreads / 8) {       // four pixels per warp step
          const float *ptr[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int pc = pix + c, yy = pc / bw, xx = pc - yy * bw;
            ptr[c] = pc < area ? f2 + ((long long)(y_lo + yy) * Wl + (x_lo + xx)) * p.C : nullptr;
          }
          const float d = ot_dot4(ptr, qv, nvec, lane);
          if ((lane & 7) == 0 && pix + (lane >> 3) < area) s_dots[pix + (lane >> 3)] = d * p.scale;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < K2; t += kBlendThreads) {
          const Taps tp = make_taps(q_ix[t], q_iy[t]);
          const bool xin0 = (unsigned)tp.x0 < (unsigned)Wl, xin1 = (unsigned)(tp.x0 + 1) < (unsigned)Wl;
          const bool yin0 = (unsigned)tp.y0 < (unsigned)Hl, yin1 = (unsigned)(tp.y0 + 1) < (unsigned)Hl;
          const int base = (tp.y0 - y_lo) * bw + (tp.x0 - x_lo);
          const float v_nw = (yin0 && xin0) ? s_dots[base] : 0.f;
          const float v_ne = (yin0 && xin1) ? s_dots[base + 1] : 0.f;
          const float v_sw = (yin1 && xin0) ? s_dots[base + bw] : 0.f;
          const float v_se = (yin1 && xin1) ? s_dots[base + bw + 1] : 0.f;
          float acc = __fmul_rn(v_nw, tp.nw);
          acc = __fmaf_rn(v_ne, tp.ne, acc);
          acc = __fmaf_rn(v_sw, tp.sw, acc);
          acc = __fmaf_rn(v_se, tp.se, acc);
          s_out[t][q] = acc;
        }
      } else {
        // large boxes (poles of the rotation map): a warp per tap, its four corners in one step
        for (int t = warp; t < K2; t += kBlendThreads / 32) {
          const Taps tp = make_taps(q_ix[t], q_iy[t]);
          const float *ptr[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int xx = tp.x0 + (c & 1), yy = tp.y0 + (c >> 1);
            const bool in = (unsigned)xx < (unsigned)Wl && (unsigned)yy < (unsigned)Hl;
            ptr[c] = in ? f2 + ((long long)yy * Wl + xx) * p.C : nullptr;
          }
          const float d = ot_dot4(ptr, qv, nvec, lane) * p.scale;
          const float v1 = __shfl_sync(0xffffffffu, d, 8), v2 = __shfl_sync(0xffffffffu, d, 16), v3 = __shfl_sync(0xffffffffu, d, 24);
          if (lane == 0) {
            float acc = __fmul_rn(d, tp.nw);
            acc = __fmaf_rn(v1, tp.ne, acc);
            acc = __fmaf_rn(v2, tp.sw, acc);
            acc = __fmaf_rn(v3, tp.se, acc);
            s_out[t][q] = acc;
          }
        }
      }
      __syncthreads();
    }
    write_taps(p, s_out, branch, lvl, b, n0);
  }
}

// ------------------------------------------------------------------------------------------------ operand planes
__global__ void __launch_bounds__(256) otf_absmax_kernel(const float *__restrict__ x, long long n, uint32_t *__restrict__ out) {
  uint32_t m = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = max(m, __float_as_uint(fabsf(x[i])));
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}
// row_elems > 0: every row of `row_elems` elements (one image row: Wl * C) is written twice, side by side (the seam, see top)
__global__ void __launch_bounds__(256) otf_split_kernel(const float *__restrict__ x, long long n, const uint32_t *__restrict__ amax, __half *__restrict__ hi,
                                                        __half *__restrict__ lo, long long row_elems) {
  const float s = ot_split_scale(*amax);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i] * s;
    const __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
    if (row_elems > 0) {
      const long long r = i / row_elems, o = i + r * row_elems;
      hi[o] = h, lo[o] = l;
      hi[o + row_elems] = h, lo[o + row_elems] = l;
    } else {
      hi[i] = h, lo[i] = l;
    }
  }
}

}  // namespace pf

// fp32 channels-last features -> fp16 hi/lo planes of the same layout (dup_row_elems = Wl * C for the target levels: rows doubled).  amax: one uint32 (zeroed by the caller before the FIRST
// tensor that shares it; the pooled levels of a feature map share level 0's word: |avg| <= max).
extern "C" int pf_onthefly_absmax(const float *x, long long count, void *amax, void *stream) {
  using namespace pf;
  PF_REQUIRE(x && amax && count > 0, "pf_onthefly_absmax: bad arguments");
  otf_absmax_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(x, count, reinterpret_cast<uint32_t *>(amax));
  return check_launch("pf_onthefly_absmax");
}
extern "C" int pf_onthefly_split(const float *x, long long count, const void *amax, void *hi, void *lo, long long dup_row_elems, void *stream) {
  using namespace pf;
  PF_REQUIRE(x && amax && hi && lo && count > 0 && dup_row_elems >= 0, "pf_onthefly_split: bad arguments");
  PF_REQUIRE(dup_row_elems == 0 || count % dup_row_elems == 0, "pf_onthefly_split: count must be a multiple of the row length");
  otf_split_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(x, count, reinterpret_cast<const uint32_t *>(amax), reinterpret_cast<__half *>(hi),
                                                          reinterpret_cast<__half *>(lo), dup_row_elems);
  return check_launch("pf_onthefly_split");
}

extern "C" int pf_lookup_onthefly_tc(const pf_onthefly_tc_args *t, void *stream) {
  using namespace pf;
  PF_REQUIRE(t != nullptr, "pf_lookup_onthefly_tc: null args");
  const pf_onthefly_args *a = &t->base;
  PF_REQUIRE(a->batch > 0 && a->h > 0 && a->w > 0, "pf_lookup_onthefly_tc: bad shape");
  PF_REQUIRE(a->radius == 4 && a->cyclic == 1, "pf_lookup_onthefly_tc: built for the model's radius-4 cyclic lookup");
  PF_REQUIRE(a->channels % 128 == 0 && a->channels <= 512 && a->channels % OT_BK == 0, "pf_lookup_onthefly_tc: channels must be a multiple of 128, <= 512");
  PF_REQUIRE(a->h % OT_TH == 0 && a->w % OT_TW == 0, "pf_lookup_onthefly_tc: the query grid must tile by %dx%d (got %dx%d)", OT_TH, OT_TW, a->h, a->w);
  PF_REQUIRE(a->num_levels >= 1 && a->num_levels <= PF_MAX_LEVELS, "pf_lookup_onthefly_tc: num_levels must be 1..%d", PF_MAX_LEVELS);
  PF_REQUIRE(a->coords && a->fmap1_own && a->out_own && t->worklist && t->pool && t->amax_own, "pf_lookup_onthefly_tc: null pointer");
  PF_REQUIRE(t->pool_segments >= 8 && t->pool_segments < (1ll << 30), "pf_lookup_onthefly_tc: pool_segments must be in [8, 2^30)");
  const bool dual = a->fmap1_other != nullptr;
  PF_REQUIRE(!dual || (a->grid_w2c && a->scratch && t->amax_other && t->tap_xy && (t->no_rotate || (a->grid_c2w && a->out_other))),
             "pf_lookup_onthefly_tc: dual lookup needs grids, out_other, scratch");
  cudaStream_t st = (cudaStream_t)stream;
  const int views = dual ? 2 : 1, L = a->num_levels, B = a->batch, h = a->h, w = a->w, C = a->channels, N = h * w;
  OtfTcParams p;
  p.B = B, p.N = N, p.h = h, p.w = w, p.C = C, p.L = L, p.div_mode = a->div_mode;
  p.tiles_x = w / OT_TW, p.tiles = p.tiles_x * (h / OT_TH), p.T = views * L * B * p.tiles;
  p.coords = a->coords;
  p.f1[0] = a->fmap1_own, p.f1[1] = a->fmap1_other;
  for (int l = 0; l < PF_MAX_LEVELS; ++l) {
    p.f2[0][l] = l < L ? a->fmap2_own[l] : nullptr;
    p.f2[1][l] = (dual && l < L) ? a->fmap2_other[l] : nullptr;
    p.axH[l] = make_axis((h >> l) > 0 ? (h >> l) : 1), p.axW[l] = make_axis((w >> l) > 0 ? (w >> l) : 1);
    if (l < L) {
      PF_REQUIRE((h >> l) >= 1 && (w >> l) >= 1, "pf_lookup_onthefly_tc: level %d is empty", l);
      PF_REQUIRE(p.f2[0][l] && t->f2_hi_own[l] && t->f2_lo_own[l], "pf_lookup_onthefly_tc: own level %d: null pointer", l);
      PF_REQUIRE(!dual || (p.f2[1][l] && t->f2_hi_other[l] && t->f2_lo_other[l]), "pf_lookup_onthefly_tc: other level %d: null pointer", l);
    }
  }
  p.ax_gw = make_axis(w), p.ax_gh = make_axis(h);
  p.grid_w2c = a->grid_w2c, p.grid_bs = a->grid_batch_stride;
  p.scale = 1.0f / sqrtf((float)C);
  // the work buffer: PF_OTF_WORK_INTS(T, pool_segments) ints = [16 counters | box_lo 4T | box_hi 4T | alloc T | fb_list T | items 2 * max_items]
  p.pool = t->pool, p.pool_segs = (int)t->pool_segments, p.max_items = (int)(t->pool_segments / 8) + p.T;
  static const int item_chunks = [] {
    const char *e = getenv("PF_OTF_ITEM_CHUNKS");       // tuning knob: passes per work item (balance vs per-item latency)
    const int v = e ? atoi(e) : 1;
    return v >= 1 && v <= 64 ? v : 1;
  }();
  p.item_chunks = item_chunks;
  p.ctr = t->worklist;
  p.box_lo = t->worklist + 16, p.box_hi = p.box_lo + 4 * p.T;
  p.alloc = p.box_hi + 4 * p.T, p.fb_list = p.alloc + p.T;
  p.items = reinterpret_cast<int2 *>(p.fb_list + p.T + (p.T & 1));
  p.amax[0] = reinterpret_cast<const uint32_t *>(t->amax_own), p.amax[1] = reinterpret_cast<const uint32_t *>(t->amax_other);
  p.out_own = a->out_own, p.out_raw = a->scratch, p.own_cl = t->no_rotate ? 1 : 0;
  p.tapxy = reinterpret_cast<float2 *>(t->tap_xy);
  otf_init_kernel<<<ceil_div(4ll * p.T, 256) < 296 ? ceil_div(4ll * p.T, 256) : 296, 256, 0, st>>>(p);
  if (int e = check_launch("pf_lookup_onthefly_tc(init)")) return e;
  const dim3 qgrid(ceil_div(N, kBlendQueries), L * views, B);
  otf_box_kernel<<<qgrid, kBlendThreads, 0, st>>>(p);
  if (int e = check_launch("pf_lookup_onthefly_tc(box)")) return e;
  otf_alloc_kernel<<<1, 1024, 0, st>>>(p);
  if (int e = check_launch("pf_lookup_onthefly_tc(alloc)")) return e;
  cudaFuncSetAttribute(otf_dots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OT_SMEM);
  OtfMaps maps;
  for (int v = 0; v < 2; ++v) {
    const int vv = v < views ? v : 0;     // single view: the unused half repeats view 0
    OtfViewMaps &m = maps.view[v];
    const void *f1_hi = vv ? t->f1_hi_other : t->f1_hi_own, *f1_lo = vv ? t->f1_lo_other : t->f1_lo_own;
    PF_REQUIRE(f1_hi && f1_lo, "pf_lookup_onthefly_tc: f1 planes missing");
    {
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
      cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)w * C * 2, (cuuint64_t)h * w * C * 2};
      cuuint32_t box[4] = {OT_BK, OT_TW, OT_TH, 1};
      if (int e = encode(&m.f1_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(f1_hi), dims, strides, box, "otf f1.hi")) return e;
      if (int e = encode(&m.f1_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(f1_lo), dims, strides, box, "otf f1.lo")) return e;
    }
    for (int l = 0; l < PF_MAX_LEVELS; ++l) {
      const int ll = l < L ? l : 0;
      const int Hl = h >> ll, Wl = w >> ll;
      const void *hi = vv ? t->f2_hi_other[ll] : t->f2_hi_own[ll], *lo = vv ? t->f2_lo_other[ll] : t->f2_lo_own[ll];
      for (int pi = 0; pi < 4; ++pi) {
        const int pitch = 32 << pi;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)(2 * Wl), (cuuint64_t)Hl, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)2 * Wl * C * 2, (cuuint64_t)Hl * 2 * Wl * C * 2};
        cuuint32_t box[4] = {OT_BK, (cuuint32_t)pitch, (cuuint32_t)(256 / pitch), 1};
        if (int e = encode(&m.f2_hi[pi][l], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(hi), dims, strides, box, "otf f2.hi")) return e;
        if (int e = encode(&m.f2_lo[pi][l], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(lo), dims, strides, box, "otf f2.lo")) return e;
      }
    }
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  otf_dots_kernel<<<sms, OT_THREADS, OT_SMEM, st>>>(maps, p);
  if (int e = check_launch("pf_lookup_onthefly_tc(dots)")) return e;
  otf_fallback_kernel<<<2 * sms, kBlendThreads, 0, st>>>(p);
  if (int e = check_launch("pf_lookup_onthefly_tc(fallback)")) return e;
  otf_blend_kernel<<<qgrid, kBlendThreads, 0, st>>>(p);
  if (int e = check_launch("pf_lookup_onthefly_tc(blend)")) return e;
  if (dual && !t->no_rotate)
    return rotate_forward(B, h, w, L, a->radius, a->div_mode, a->grid_c2w, a->grid_batch_stride, a->scratch, a->out_other, 0, 0, st, nullptr);
  return 0;
}
