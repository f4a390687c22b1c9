// (b) Dual-cost pyramid lookup — DCCL.__call__ (PriOr-RAFT/core/corr.py:113-144) and
// CorrBlock.__call__ (core/corr.py:30-51): one gather launch serving both views + one rotate launch.
//
// Forward, radius 4 (the model's):  lookup_rows_kernel  ->  rotate_fwd_kernel   (documented where they are defined)
// Forward, other radii; backward:   lookup_kernel<R, kBwd>  ->  rotate_fwd_kernel / rotate_kernel<kBwd>
//
// lookup_kernel (r01; generic radius and the scatter/backward instantiation) — grid = (ceil(N/32) query chunks,
//   levels x branches, batch), 256 threads.  A CTA owns 32 consecutive query pixels of one level of one branch; each
//   warp walks 4 queries, a lane owns a tap.
//   * Window coordinates are separable: the 2r+1 x-coordinates and 2r+1 y-coordinates of a window
//     go through the (remainder, normalise, unnormalise, floor, clamp) chain once each, on lanes
//     0..2k-1, into a per-warp shared-memory table of {clamped offsets, validity-folded weights};
//     a tap is then two 16-byte table reads, four adds and four multiplies away from its loads.
//   * The (k+1)^2 footprint the window touches — of the query's private plane (own view) or of the
//     two channels of the level-0 rotation grid (other view) — is staged in shared memory with
//     cooperative loads whenever consecutive window cells share a corner; taps then blend from smem.
//   * Branch 0 (own view): the blend is the result.  NCHW output goes through a shared-memory
//     transpose and is written as full 128-byte rows; channels-last output is written directly.
//   * Branch 1 (other view): the blend of the grid is the mapped point; pyr_other[l][n] is sampled
//     there (scale mixing is the reference's).  The pre-rotation map goes to
//     `scratch` CHANNELS-LAST ([B, N, L*81]) so that the rotate kernels read whole 1296-byte vectors.
// rotate_kernel<kBwd> — scalar img_rotate(., grid_c2w) of that map (core/corr.py:137-138) and its adjoint: 32 pixels x
//   one level per CTA, lanes across channels.  The forward of the model's shapes runs rotate_fwd_kernel (float4).
// Coordinates are bit-exact restatements (pf_common.cuh); values are ATen's FMA chain.
#include "pf_common.cuh"

#include <stdlib.h>

namespace pf {

constexpr int kQueriesPerCta = 32;
constexpr int kLookupThreads = 256;
constexpr int kQueriesPerWarp = kQueriesPerCta / (kLookupThreads / 32);

struct LookupParams {
  int B, N, h, w;  // query grid, N = h*w
  int radius, L, cyclic, div_mode, dual, channels_last, fuse_sum;
  int w_pow2, w2_pow2;                 // the rotation grid's width / the pyramid's level-0 width is a power of two
  int l2_hint;                         // lookup_rows_kernel: per-level L2 eviction policies on the plane reads (pf_common.cuh)
  int q_begin, q_end;                  // lookup_kernel (backward): queries [q_begin, q_end) of every batch item; the gradient pyramids
                                       // then hold (q_end - q_begin) planes per batch item (chunked, volume-free backward)
  const float *coords;
  const float *own[PF_MAX_LEVELS];
  const float *other[PF_MAX_LEVELS];
  int Hl[PF_MAX_LEVELS], Wl[PF_MAX_LEVELS];
  Axis axW[PF_MAX_LEVELS], axH[PF_MAX_LEVELS];
  Axis ax_gw, ax_gh;  // axes of the rotation grid (query resolution)
  const float *grid_w2c;
  long long grid_bs;
  float *out_own;                      // forward: [B, L*K2, N] (or channels-last [B, N, L*K2]) output.  backward: incoming gradient
  float *raw;                          // forward: [B, N, L*K2] pre-rotation map.  backward: its gradient (read only)
  float *dbg_own, *dbg_other;
  float *d_own[PF_MAX_LEVELS];         // backward: gradient pyramids (+=)
  float *d_other[PF_MAX_LEVELS];
};

// p.X[lvl] with a run-time lvl makes nvcc copy the whole parameter array to local memory (r03a: 56-byte stack frame in
// lookup_rows_kernel, ~200 prologue instructions per warp).  Four selects on constant-bank operands instead.
template <typename T>
__device__ __forceinline__ T pick_level(const T (&a)[PF_MAX_LEVELS], int i) {
  static_assert(PF_MAX_LEVELS == 4, "pick_level is written for four levels");
  return i == 0 ? a[0] : (i == 1 ? a[1] : (i == 2 ? a[2] : a[3]));
}
__device__ __forceinline__ Axis pick_axis(const Axis (&a)[PF_MAX_LEVELS], int i) {
  Axis r;
  r.size = i == 0 ? a[0].size : (i == 1 ? a[1].size : (i == 2 ? a[2].size : a[3].size));
  r.size_m1 = i == 0 ? a[0].size_m1 : (i == 1 ? a[1].size_m1 : (i == 2 ? a[2].size_m1 : a[3].size_m1));
  r.inv_m1 = i == 0 ? a[0].inv_m1 : (i == 1 ? a[1].inv_m1 : (i == 2 ? a[2].inv_m1 : a[3].inv_m1));
  return r;
}

// Adjoint of blend_zeros: scatter g * w into the four in-bounds corners.
__device__ __forceinline__ void scatter_zeros(float *__restrict__ plane, int H, int W, const Taps &t, float g) {
  const bool xin0 = (unsigned)t.x0 < (unsigned)W, xin1 = (unsigned)(t.x0 + 1) < (unsigned)W;
  const bool yin0 = (unsigned)t.y0 < (unsigned)H, yin1 = (unsigned)(t.y0 + 1) < (unsigned)H;
  float *r0 = plane + (long long)t.y0 * W + t.x0;
  float *r1 = r0 + W;
  if (yin0 && xin0) atomicAdd(r0, g * t.nw);
  if (yin0 && xin1) atomicAdd(r0 + 1, g * t.ne);
  if (yin1 && xin0) atomicAdd(r1, g * t.sw);
  if (yin1 && xin1) atomicAdd(r1 + 1, g * t.se);
}

// Per-warp shared-memory scratch of lookup_kernel (sizes for window edge k = 2r+1):
//   T   [32] AxisEntry     one entry per window column (lanes 0..k-1) and row (lanes k..2k-1)
//   pos [2][16] int        the k+1 distinct x (resp. y*stride) offsets of the footprint
//   fp  [2][(k+1)(k+2)]    the staged (k+1)^2 footprint of the plane (own view) or of both grid channels (other view)
//   dbg [32] float         sample coordinates for the debug dump
__host__ __device__ constexpr int lookup_fp_floats(int k) { return (k + 1) * (k + 2); }
__host__ __device__ constexpr int lookup_warp_floats(int k) { return 2 * (32 * 4 + 32 + 2 * lookup_fp_floats(k) + 32); }  // two sets

template <int R, bool kBwd, int kDiv, int BRANCH>
__device__ __forceinline__ void lookup_body(const LookupParams &p, const int lvl) {
  constexpr int branch = BRANCH;
  const int r = (R > 0) ? R : p.radius;
  const int k = 2 * r + 1;
  const int K2 = k * k;
  const int k1 = k + 1, pitch = k + 2, FP = k1 * k1;
  constexpr int kRounds = (R > 0) ? ((2 * R + 1) * (2 * R + 1) + 31) / 32 : 8;     // radius <= 7 -> <= 225 taps
  constexpr int kFpRounds = (R > 0) ? ((2 * R + 2) * (2 * R + 2) + 31) / 32 : 8;   // <= 256 footprint cells
  extern __shared__ float4 smem4[];
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp-uniform for the compiler
  float *wbase = reinterpret_cast<float *>(smem4) + warp * lookup_warp_floats(k);
  AxisEntry *T = reinterpret_cast<AxisEntry *>(wbase);              // [2][32]
  int *pos = reinterpret_cast<int *>(wbase + 256);                   // [2][32]
  float *fp = wbase + 320;                                           // [2][2][fp_floats]
  float *dbg_tab = fp + 4 * lookup_fp_floats(k);                     // [2][32]
  float *tile = reinterpret_cast<float *>(smem4) + (kLookupThreads / 32) * lookup_warp_floats(k);  // [K2][33], NCHW own view
  const int b = blockIdx.z;
  const int n0 = p.q_begin + blockIdx.x * kQueriesPerCta;
  const int Hl = p.Hl[lvl], Wl = p.Wl[lvl];
  const Axis axW = p.axW[lvl], axH = p.axH[lvl];
  const float inv_scale = 1.0f / (float)(1 << lvl);  // `coords / 2**i` is exact either way
  const float *vol = branch ? p.other[lvl] : p.own[lvl];
  const float *gridx = opaque(p.grid_w2c + (long long)b * p.grid_bs);
  const float *gridy = opaque(gridx + p.N);
  float *dbg = branch ? p.dbg_other : p.dbg_own;
  const bool use_tile = !branch && !p.channels_last;   // compile-time false for the other view
  float *io = p.out_own + ((long long)b * p.L + lvl) * K2 * (long long)p.N + n0;
  if constexpr (kBwd) {  // NCHW own branch: stage the incoming gradient tile [K2][32 queries], coalesced rows
    if (use_tile) {
      if (n0 + lane < p.q_end)
        for (int ch = warp; ch < K2; ch += kLookupThreads / 32) tile[ch * 33 + lane] = io[(long long)ch * p.N + lane];
      __syncthreads();
    }
  }
  // this lane's window axis: lanes [0,k) hold x offsets, lanes [k,2k) y offsets
  const bool is_x = lane < k;
  const float off = (float)((is_x ? lane : lane - k) - r);
  const Axis ax1 = branch ? (is_x ? p.ax_gw : p.ax_gh) : (is_x ? axW : axH);  // first sampler's axis
  const int size1 = branch ? (is_x ? p.w : p.h) : (is_x ? Wl : Hl);
  const int stride1 = is_x ? 1 : (branch ? p.w : Wl);
  const bool wrap1 = is_x && (branch || p.cyclic);
  const bool chain_lane = lane < 2 * k;
  const bool has_next = chain_lane && lane != k - 1 && lane != 2 * k - 1;   // the next lane holds the next cell of my axis
  // tap -> (window column a, window row b): fixed per lane and round.
  // own view: lanes walk x fastest (coalesced rows); other view: lanes walk the output channel.
  int t_a[kRounds], t_b[kRounds];
#pragma unroll
  for (int it = 0; it < kRounds; ++it) {
    const int t = it * 32 + lane;
    const int tt = t < K2 ? t : 0;
    const int hi = tt / k, lo = tt - hi * k;
    t_a[it] = branch ? hi : lo;
    t_b[it] = branch ? lo : hi;
  }
  // footprint cell -> (row, column), fixed per lane and round
  int f_rc[kFpRounds];
#pragma unroll
  for (int j = 0; j < kFpRounds; ++j) {
    const int i = j * 32 + lane;
    const int ii = i < FP ? i : 0;
    const int rr = ii / k1;
    f_rc[j] = (rr << 8) | (ii - rr * k1);
  }

  const int plane_sz = Hl * Wl;
  const int q0 = warp * kQueriesPerWarp;
  const long long nq0 = (long long)b * p.N + n0 + q0;   // first (batch, query) row of this warp
  const float *coord_ptr = opaque(p.coords + ((long long)b * 2 + (is_x ? 0 : 1)) * p.N + n0 + q0);
  const float *plane_it = kBwd ? nullptr : opaque(vol + nq0 * plane_sz);
  // gradient planes are indexed relative to the chunk: plane (b, n) lives at b * (q_end - q_begin) + (n - q_begin)
  const long long dq0 = (long long)b * (p.q_end - p.q_begin) + (n0 + q0 - p.q_begin);
  float *dplane_it = kBwd ? opaque((branch ? p.d_other[lvl] : p.d_own[lvl]) + dq0 * plane_sz) : nullptr;
  float *out_it = opaque((branch ? p.raw : p.out_own) + (nq0 * p.L + lvl) * K2);   // channels-last row of this query
  const bool has_dbg = dbg != nullptr;

  // Software pipeline over the warp's queries: while the taps of query qi run, the coordinate chain of qi+1 has
  // already been resolved into the other table set and its footprint loads are in flight (held in registers, parked
  // in shared memory at the top of the next iteration) — a warp always has DRAM requests outstanding.
  float pf0[kFpRounds], pf1[kFpRounds];   // prefetched footprint cells (second array: grid y channel, other view)
  // (1) core/corr.py:123-126 + the sampler's coordinate chain, once per window row / column, into table set `set`
  auto resolve = [&](int qi, int set) -> bool {
    const float c = __fmul_rn(__ldg(coord_ptr + qi), inv_scale);
    float pc = __fadd_rn(c, off);
    if (wrap1) pc = remainder_pos(pc, ax1.size);
    const float sc = sample_coord<kDiv>(pc, ax1);
    const AxisEntry e = make_axis_entry(sc, size1, stride1);
    T[set * 32 + lane] = e;
    if (chain_lane) {
      pos[set * 32 + (is_x ? 0 : 16) + (is_x ? lane : lane - k)] = e.o0;
      if (!has_next) pos[set * 32 + (is_x ? 0 : 16) + k] = e.o1;
    }
    if (has_dbg) dbg_tab[set * 32 + lane] = sc;
    // the footprint is a (k+1)^2 lattice iff consecutive cells share a corner (false across the seam)
    const int next_o0 = __shfl_down_sync(0xffffffffu, e.o0, 1);
    const bool lattice = !kBwd && __all_sync(0xffffffffu, !has_next || e.o1 == next_o0);
    __syncwarp();
    return lattice;
  };
  // (2a) issue the cooperative footprint loads (coalesced row segments of the plane / of the grid) into registers
  auto prefetch = [&](const float *plane, int set) {
#pragma unroll
    for (int j = 0; j < kFpRounds; ++j) {
      if (j * 32 < FP && j * 32 + lane < FP) {
        const int o = pos[set * 32 + 16 + (f_rc[j] >> 8)] + pos[set * 32 + (f_rc[j] & 255)];
        if (branch) {
          pf0[j] = __ldg(gridx + o);
          pf1[j] = __ldg(gridy + o);
        } else {
          pf0[j] = __ldg(plane + o);
        }
      }
    }
  };
  // (2b) park them in shared memory
  auto park = [&](int set) {
    float *f = fp + set * 2 * lookup_fp_floats(k);
#pragma unroll
    for (int j = 0; j < kFpRounds; ++j) {
      if (j * 32 < FP && j * 32 + lane < FP) {
        const int cell = (f_rc[j] >> 8) * pitch + (f_rc[j] & 255);
        f[cell] = pf0[j];
        if (branch) f[lookup_fp_floats(k) + cell] = pf1[j];
      }
    }
  };
  const int nq_w = min(kQueriesPerWarp, p.q_end - n0 - q0);   // live queries of this warp (may be <= 0)
  bool fast_next = false;
  if (nq_w > 0) {
    fast_next = resolve(0, 0);
    if (fast_next) prefetch(plane_it, 0);
  }

#pragma unroll 1
  for (int qi = 0; qi < nq_w; ++qi) {
    const int q = q0 + qi;
    const int n = n0 + q;
    const int set = qi & 1;
    const bool fast = fast_next;
    const float *plane = plane_it;
    float *dplane = dplane_it;
    float *outq = out_it;
    plane_it += plane_sz;
    dplane_it += plane_sz;
    out_it += p.L * K2;
    if (fast) park(set);
    __syncwarp();                       // footprint of qi visible; taps of qi-1 (other set) finished in every lane
    if (qi + 1 < nq_w) {
      fast_next = resolve(qi + 1, set ^ 1);
      if (fast_next) prefetch(plane_it, set ^ 1);
    }
    const AxisEntry *Tq = T + set * 32;
    const float *fpq = fp + set * 2 * lookup_fp_floats(k);
    const float *dbgq = dbg_tab + set * 32;
    // ---- (3) taps
#pragma unroll
    for (int it = 0; it < kRounds; ++it) {
      if (it * 32 >= K2) break;
      if (it * 32 + lane < K2) {
        const int aa = t_a[it], bb = t_b[it];
        const int ch = aa * k + bb;  // x-major channel order of the reference
        const AxisEntry ex = Tq[aa], ey = Tq[k + bb];
        float w_nw = __fmul_rn(ex.w0, ey.w0), w_ne = __fmul_rn(ex.w1, ey.w0);
        float w_sw = __fmul_rn(ex.w0, ey.w1), w_se = __fmul_rn(ex.w1, ey.w1);
        float ix = 0.f, iy = 0.f, val = 0.f;
        if constexpr (!branch) {
          if constexpr (kBwd) {
            const float g = use_tile ? tile[ch * 33 + q] : __ldg(outq + ch);
            if (w_nw != 0.f) atomicAdd(dplane + ey.o0 + ex.o0, g * w_nw);
            if (w_ne != 0.f) atomicAdd(dplane + ey.o0 + ex.o1, g * w_ne);
            if (w_sw != 0.f) atomicAdd(dplane + ey.o1 + ex.o0, g * w_sw);
            if (w_se != 0.f) atomicAdd(dplane + ey.o1 + ex.o1, g * w_se);
          } else if (fast) {
            const float *f = fpq + bb * pitch + aa;
            val = __fmul_rn(f[0], w_nw);
            val = __fmaf_rn(f[1], w_ne, val);
            val = __fmaf_rn(f[pitch], w_sw, val);
            val = __fmaf_rn(f[pitch + 1], w_se, val);
          } else {
            val = __fmul_rn(__ldg(plane + ey.o0 + ex.o0), w_nw);
            val = __fmaf_rn(__ldg(plane + ey.o0 + ex.o1), w_ne, val);
            val = __fmaf_rn(__ldg(plane + ey.o1 + ex.o0), w_sw, val);
            val = __fmaf_rn(__ldg(plane + ey.o1 + ex.o1), w_se, val);
          }
        } else {
          // core/corr.py:132-136 — map through the level-0 rotation grid, then index the level-l volume
          float sx, sy;
          if (fast) {
            const float *f = fpq + bb * pitch + aa, *g = f + lookup_fp_floats(k);
            sx = __fmul_rn(f[0], w_nw), sy = __fmul_rn(g[0], w_nw);
            sx = __fmaf_rn(f[1], w_ne, sx), sy = __fmaf_rn(g[1], w_ne, sy);
            sx = __fmaf_rn(f[pitch], w_sw, sx), sy = __fmaf_rn(g[pitch], w_sw, sy);
            sx = __fmaf_rn(f[pitch + 1], w_se, sx), sy = __fmaf_rn(g[pitch + 1], w_se, sy);
          } else {
            const int o_nw = ey.o0 + ex.o0, o_ne = ey.o0 + ex.o1, o_sw = ey.o1 + ex.o0, o_se = ey.o1 + ex.o1;
            sx = __fmul_rn(__ldg(gridx + o_nw), w_nw), sy = __fmul_rn(__ldg(gridy + o_nw), w_nw);
            sx = __fmaf_rn(__ldg(gridx + o_ne), w_ne, sx), sy = __fmaf_rn(__ldg(gridy + o_ne), w_ne, sy);
            sx = __fmaf_rn(__ldg(gridx + o_sw), w_sw, sx), sy = __fmaf_rn(__ldg(gridy + o_sw), w_sw, sy);
            sx = __fmaf_rn(__ldg(gridx + o_se), w_se, sx), sy = __fmaf_rn(__ldg(gridy + o_se), w_se, sy);
          }
          ix = sample_coord<kDiv>(remainder_pos(sx, axW.size), axW);
          iy = sample_coord<kDiv>(sy, axH);
          const float fx = floorf(ix), fy = floorf(iy);
          const int x0 = (int)fx, y0 = (int)fy;
          if (!kBwd && (unsigned)x0 < (unsigned)(Wl - 1) && (unsigned)y0 < (unsigned)(Hl - 1)) {
            // interior: all four taps valid, one base pointer, immediate offsets
            const float dxw = __fsub_rn(ix, fx), dxe = __fsub_rn(__fadd_rn(fx, 1.f), ix);
            const float dyn = __fsub_rn(iy, fy), dys = __fsub_rn(__fadd_rn(fy, 1.f), iy);
            const float *s0 = plane + (y0 * Wl + x0), *s1 = s0 + Wl;
            val = __fmul_rn(__ldg(s0), __fmul_rn(dxe, dys));
            val = __fmaf_rn(__ldg(s0 + 1), __fmul_rn(dxw, dys), val);
            val = __fmaf_rn(__ldg(s1), __fmul_rn(dxe, dyn), val);
            val = __fmaf_rn(__ldg(s1 + 1), __fmul_rn(dxw, dyn), val);
          } else if (x0 >= -1 && x0 < Wl && y0 >= -1 && y0 < Hl) {
            // straddles the border (a tap entirely outside the plane is an exact 0 and needs no DRAM round trip:
            // most of levels 2-3 here, whose level-0-unit coordinates index a 16x32 / 8x16 plane)
            const AxisEntry gx = make_axis_entry(ix, Wl, 1), gy = make_axis_entry(iy, Hl, Wl);
            const float v_nw = __fmul_rn(gx.w0, gy.w0), v_ne = __fmul_rn(gx.w1, gy.w0);
            const float v_sw = __fmul_rn(gx.w0, gy.w1), v_se = __fmul_rn(gx.w1, gy.w1);
            if constexpr (kBwd) {
              const float g = __ldg(outq + ch);
              if (v_nw != 0.f) atomicAdd(dplane + gy.o0 + gx.o0, g * v_nw);
              if (v_ne != 0.f) atomicAdd(dplane + gy.o0 + gx.o1, g * v_ne);
              if (v_sw != 0.f) atomicAdd(dplane + gy.o1 + gx.o0, g * v_sw);
              if (v_se != 0.f) atomicAdd(dplane + gy.o1 + gx.o1, g * v_se);
            } else {
              val = __fmul_rn(__ldg(plane + gy.o0 + gx.o0), v_nw);
              val = __fmaf_rn(__ldg(plane + gy.o0 + gx.o1), v_ne, val);
              val = __fmaf_rn(__ldg(plane + gy.o1 + gx.o0), v_sw, val);
              val = __fmaf_rn(__ldg(plane + gy.o1 + gx.o1), v_se, val);
            }
          }
        }
        if constexpr (!kBwd) {
          if (use_tile)
            tile[ch * 33 + q] = val;
          else
            outq[ch] = val;
          if (has_dbg) {
            if (!branch) {
              ix = dbgq[aa];
              iy = dbgq[k + bb];
            }
            float *d = dbg + ((((long long)b * p.N + n) * p.L + lvl) * K2 + ch) * 2;
            d[0] = ix;
            d[1] = iy;
          }
        }
      }
    }
  }
  if constexpr (!kBwd) {
    if (use_tile) {
      __syncthreads();
      if (n0 + lane < p.q_end)
        for (int ch = warp; ch < K2; ch += kLookupThreads / 32) io[(long long)ch * p.N + lane] = tile[ch * 33 + lane];
    }
  }
}

template <int R, bool kBwd, int kDiv>
__global__ void __launch_bounds__(kLookupThreads, 4) lookup_kernel(const LookupParams p) {
  // one launch serves both views; each CTA runs the body specialised for its branch
  if (blockIdx.y < p.L)
    lookup_body<R, kBwd, kDiv, 0>(p, blockIdx.y);
  else
    lookup_body<R, kBwd, kDiv, 1>(p, blockIdx.y - p.L);
}

// ------------------------------------------------------------------------------------------------
// lookup_rows_kernel — the radius-4 forward kernel (r02).  ncu r01e showed the tap-per-lane kernel above issue-bound
// (30 M warp instructions per call at 65 % issue-slot utilisation, L1 79 % busy, DRAM 17 %): per (query, level) it
// paid a coordinate-table build, a staged footprint copy and 3 rounds of 81/96-efficient taps that each re-read the
// tables.  Here a lane owns one COLUMN of one query's window and walks its rows:
//   * lane = 10 * qq + a: three queries per warp, a = 0..9 the 10 columns of the (k+1)^2 footprint.  The lane runs
//     the x coordinate chain of its own column (weights stay in registers for all 9 rows) and the y chain of window
//     row a, which it publishes to a 400-byte per-warp table (offsets as int4 rows, weights as float2).
//   * own view: one LDG per footprint row (three coalesced 40-byte segments per warp instruction), the right
//     neighbour by SHFL, 4 weight products + 4 FMAs per output: ~60 instructions per (query, level) instead of ~270.
//   * other view: the same walk over the two channels of the rotation grid gives the mapped point of every tap;
//     the second sampler then runs in groups of three rows so that 12 plane loads per lane are in flight at once,
//     with a warp vote that skips rows mapped entirely outside the (smaller, scale-mixed) level-l plane — most of
//     levels 1-3 — and an all-interior shortcut (one base pointer, no clamps).
//   * the lattice assumption (corner sharing between neighbouring cells) is verified per triple with one vote;
//     the rare exceptions (a coordinate within an ulp of an integer) take a direct four-load path.
//   * outputs go through one [81][Q|1] shared tile per CTA (Q = 48 queries x one level x one branch) and leave as
//     128-byte rows (NCHW) or 324-byte channel runs (channels-last scratch).
// Coordinate arithmetic is shortened with exact identities only (see sample_coord_x): results are bit-identical
// to the long chain, which the debug dump (kDbg) still proves tap by tap.
constexpr int kRowsThreads = 256;
constexpr int kRowsWarps = kRowsThreads / 32;
constexpr int kRowsK = 9, kRowsK2 = 81;
#ifndef PF_ROWS_MIN_CTAS
#define PF_ROWS_MIN_CTAS 6   // sweep r02d (profiles/): Q = 48 x 6 CTAs per SM is the fastest shape; L1 (what the carve-out leaves) matters as much as warps
#endif
#ifndef PF_OWN_STREAM
#define PF_OWN_STREAM 1
#endif
constexpr bool kOwnStream = PF_OWN_STREAM;   // own-view footprint rows are read exactly once: evict-first loads keep L1 for the grid
constexpr int kRowsWarpWords = 3 * 12 * 2 + 3 * 10 * 2 + 32;   // o0 rows, o1 rows (int, pitch 12), weights (float2, pitch 10), dbg y

template <int kDiv, bool kDbg, int kRowsQ, int BRANCH>
__device__ __forceinline__ void lookup_rows_body(const LookupParams &p, const int lvl) {
  constexpr int kRowsPitch = kRowsQ | 1;   // tile pitch (odd -> conflict-free both ways)
  extern __shared__ float4 smem4[];
  float *tile = reinterpret_cast<float *>(smem4);                                  // [81][Q|1]
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  float *wbase = tile + ((kRowsK2 * kRowsPitch + 3) & ~3) + warp * kRowsWarpWords;
  int *tyo0 = reinterpret_cast<int *>(wbase);            // [3][12] row offsets of tap y0 (entry 9 = o1 of row 8)
  int *tyo1 = tyo0 + 36;                                 // [3][12] row offsets of tap y1 (direct path only)
  float2 *tyw = reinterpret_cast<float2 *>(tyo1 + 36);   // [3][10] (w0, w1) per window row
  float *tdbg = reinterpret_cast<float *>(tyw + 30);     // [27] y sample coordinate (debug dump)
  const int b = blockIdx.z, n0 = blockIdx.x * kRowsQ;
  const int Hl = pick_level(p.Hl, lvl), Wl = pick_level(p.Wl, lvl);
  const Axis axW = pick_axis(p.axW, lvl), axH = pick_axis(p.axH, lvl);
  const float hW = __fmul_rn(0.5f, axW.size_m1), hH = __fmul_rn(0.5f, axH.size_m1);
  const float inv_scale = 1.0f / (float)(1 << lvl);
  const int l10 = lane / 10;
  const int qq = l10 < 2 ? l10 : 2;             // lanes 30, 31 shadow lanes 20, 21 (valid addresses, no stores)
  const int a = lane - 10 * l10;
  const bool act = lane < 30 && a < kRowsK;     // this lane produces outputs (column a of query qq)
  const float off = (float)(a - 4);
  const Axis ax1x = BRANCH ? p.ax_gw : axW, ax1y = BRANCH ? p.ax_gh : axH;   // first sampler's axes
  const float h1x = __fmul_rn(0.5f, ax1x.size_m1), h1y = __fmul_rn(0.5f, ax1y.size_m1);
  const int size1x = BRANCH ? p.w : Wl, size1y = BRANCH ? p.h : Hl;
  const bool wrap1 = BRANCH || p.cyclic;
  const bool pow2_1 = BRANCH ? p.w_pow2 : p.w2_pow2;
  const float *vol = opaque(BRANCH ? pick_level(p.other, lvl) : pick_level(p.own, lvl));
  const int plane_sz = Hl * Wl;
  const float *gridx = opaque(p.grid_w2c + (long long)b * p.grid_bs);
  const float *cxp = opaque(p.coords + (long long)b * 2 * p.N);
  float *dbg = BRANCH ? p.dbg_other : p.dbg_own;
  float *tcol = tile + a * kRowsK * kRowsPitch;   // + b * pitch + query
  // planes of the small levels are re-read by the next lookup call, planes of the large ones are not (see pf_common.cuh)
  const uint64_t pol = make_l2_policy(p.l2_hint ? (lvl >= 2 ? kL2Last : kL2First) : (BRANCH ? kL2Normal : kL2First));

  // the coordinates of the warp's next triple are fetched one iteration ahead (the chain below starts with them)
  float ncx, ncy;
  {
    const int n_first = min(n0 + 3 * warp + qq, p.N - 1);
    ncx = __ldg(cxp + n_first), ncy = __ldg(cxp + p.N + n_first);
  }
#pragma unroll 1
  for (int t = warp; t < kRowsQ / 3; t += kRowsWarps) {
    if (n0 + 3 * t >= p.N) break;                                  // warp-uniform
    __syncwarp();                                                  // the previous triple's table reads are done
    const int ql = 3 * t + qq;
    const int n = min(n0 + ql, p.N - 1);
    // ---- coordinate chains: column a (x) and window row a (y) of query qq   (core/corr.py:123-126 + utils.py:85-86)
    const float cx = __fmul_rn(ncx, inv_scale), cy = __fmul_rn(ncy, inv_scale);
    {
      const int n_next = min(n0 + ql + 3 * kRowsWarps, p.N - 1);
      ncx = __ldg(cxp + n_next), ncy = __ldg(cxp + p.N + n_next);
    }
    float px = __fadd_rn(cx, off);
    if (wrap1) px = remainder_sel(px, ax1x, pow2_1);
    const float scx = sample_coord_x<kDiv>(px, ax1x, h1x);
    const float scy = sample_coord_x<kDiv>(__fadd_rn(cy, off), ax1y, h1y);
    const AxisEntry ex = make_axis_entry(scx, size1x, 1);
    const AxisEntry ey = make_axis_entry(scy, size1y, size1x);
    if (act) {
      tyo0[qq * 12 + a] = ey.o0;
      tyo1[qq * 12 + a] = ey.o1;
      if (a == kRowsK - 1) tyo0[qq * 12 + kRowsK] = ey.o1;
      tyw[qq * 10 + a] = make_float2(ey.w0, ey.w1);
      if (kDbg) tdbg[qq * 9 + a] = scy;
    }
    const int o1_left = __shfl_up_sync(0xffffffffu, ex.o1, 1);
    const int col = (a == kRowsK) ? o1_left : ex.o0;               // the footprint column this lane loads
    const int col_right = __shfl_down_sync(0xffffffffu, col, 1);
    const int yo0_next = __shfl_down_sync(0xffffffffu, ey.o0, 1);
    const bool okx = !act || ex.w1 == 0.f || ex.o1 == col_right;
    const bool oky = !act || a == kRowsK - 1 || ey.w1 == 0.f || ey.o1 == yo0_next;
    const bool lattice = __all_sync(0xffffffffu, okx && oky);
    __syncwarp();
    const long long row = (long long)b * p.N + n;
    const float2 *wy = tyw + qq * 10;

    if constexpr (!BRANCH) {
      // ================= own view =================
      const float *pl = opaque(vol + row * plane_sz);
      if (lattice) {
        const float *plc = opaque(pl + col);
        const int4 ya = *reinterpret_cast<const int4 *>(tyo0 + qq * 12), yb = *reinterpret_cast<const int4 *>(tyo0 + qq * 12 + 4);
        const int2 yc = *reinterpret_cast<const int2 *>(tyo0 + qq * 12 + 8);
        const int yo[10] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w, yc.x, yc.y};
        float v[10], vr[10];
#pragma unroll
        for (int r = 0; r < 10; ++r) v[r] = ld_hint_stream(plc + yo[r], pol);
#pragma unroll
        for (int r = 0; r < 10; ++r) vr[r] = __shfl_down_sync(0xffffffffu, v[r], 1);
#pragma unroll
        for (int bb = 0; bb < kRowsK; ++bb) {
          const float2 w = wy[bb];
          float val = __fmul_rn(v[bb], __fmul_rn(ex.w0, w.x));
          val = __fmaf_rn(vr[bb], __fmul_rn(ex.w1, w.x), val);
          val = __fmaf_rn(v[bb + 1], __fmul_rn(ex.w0, w.y), val);
          val = __fmaf_rn(vr[bb + 1], __fmul_rn(ex.w1, w.y), val);
          if (act) tcol[bb * kRowsPitch + ql] = val;
        }
      } else {
#pragma unroll 1
        for (int bb = 0; bb < kRowsK; ++bb) {
          const float2 w = wy[bb];
          const float *r0 = pl + tyo0[qq * 12 + bb], *r1 = pl + tyo1[qq * 12 + bb];
          float val = __fmul_rn(__ldg(r0 + ex.o0), __fmul_rn(ex.w0, w.x));
          val = __fmaf_rn(__ldg(r0 + ex.o1), __fmul_rn(ex.w1, w.x), val);
          val = __fmaf_rn(__ldg(r1 + ex.o0), __fmul_rn(ex.w0, w.y), val);
          val = __fmaf_rn(__ldg(r1 + ex.o1), __fmul_rn(ex.w1, w.y), val);
          if (act) tcol[bb * kRowsPitch + ql] = val;
        }
      }
      if (kDbg && dbg != nullptr && act && n0 + ql < p.N) {
        for (int bb = 0; bb < kRowsK; ++bb) {
          float *d = dbg + (((row * p.L + lvl) * kRowsK2) + a * kRowsK + bb) * 2;
          d[0] = scx;
          d[1] = tdbg[qq * 9 + bb];
        }
      }
    } else {
      // ================= other view =================  core/corr.py:132-136
      const float *pl = opaque(vol + row * plane_sz);
      float *dq = (kDbg && dbg != nullptr && act && n0 + ql < p.N) ? dbg + ((row * p.L + lvl) * kRowsK2 + a * kRowsK) * 2 : nullptr;
      if (!lattice) {
        // rare: a window coordinate within an ulp of an integer broke the corner sharing — four direct loads per tap
        const float *gy0 = gridx + p.N;
#pragma unroll 1
        for (int bb = 0; bb < kRowsK; ++bb) {
          const float2 w = wy[bb];
          const int r0 = tyo0[qq * 12 + bb], r1 = tyo1[qq * 12 + bb];
          const float w_nw = __fmul_rn(ex.w0, w.x), w_ne = __fmul_rn(ex.w1, w.x);
          const float w_sw = __fmul_rn(ex.w0, w.y), w_se = __fmul_rn(ex.w1, w.y);
          float x = __fmul_rn(__ldg(gridx + r0 + ex.o0), w_nw), y = __fmul_rn(__ldg(gy0 + r0 + ex.o0), w_nw);
          x = __fmaf_rn(__ldg(gridx + r0 + ex.o1), w_ne, x), y = __fmaf_rn(__ldg(gy0 + r0 + ex.o1), w_ne, y);
          x = __fmaf_rn(__ldg(gridx + r1 + ex.o0), w_sw, x), y = __fmaf_rn(__ldg(gy0 + r1 + ex.o0), w_sw, y);
          x = __fmaf_rn(__ldg(gridx + r1 + ex.o1), w_se, x), y = __fmaf_rn(__ldg(gy0 + r1 + ex.o1), w_se, y);
          const float ix = sample_coord_x<kDiv>(remainder_sel(x, axW, p.w2_pow2), axW, hW);
          const float iy = sample_coord_x<kDiv>(y, axH, hH);
          const AxisEntry gxe = make_axis_entry(ix, Wl, 1), gye = make_axis_entry(iy, Hl, Wl);
          float v = __fmul_rn(__ldg(pl + gye.o0 + gxe.o0), __fmul_rn(gxe.w0, gye.w0));
          v = __fmaf_rn(__ldg(pl + gye.o0 + gxe.o1), __fmul_rn(gxe.w1, gye.w0), v);
          v = __fmaf_rn(__ldg(pl + gye.o1 + gxe.o0), __fmul_rn(gxe.w0, gye.w1), v);
          v = __fmaf_rn(__ldg(pl + gye.o1 + gxe.o1), __fmul_rn(gxe.w1, gye.w1), v);
          if (act) tcol[bb * kRowsPitch + ql] = v;
          if (kDbg && dq != nullptr) dq[2 * bb] = ix, dq[2 * bb + 1] = iy;
        }
      } else {
        const float *gxc = opaque(gridx + col), *gyc = opaque(gxc + p.N);
        if (!kDbg && Hl < p.h) {
          // Scale-mixed levels: the mapped y of a tap is a convex combination of the footprint's grid-y values (all four
          // corners of every first-sampler tap inside the grid => weights sum to 1 up to rounding).  If even the smallest
          // of them lies below the level-l plane by a margin far above the chain's rounding (1e-2 px vs < 1e-4 px), every
          // tap of the triple has floor(iy) >= Hl and the result is exact zeros — without touching the x channel, the
          // coordinate chains or the other view's volume.  ~55 % of the level 1-3 triples leave here.
          const int *yall = tyo0 + qq * 12;
          float gmin = __ldg(gyc + yall[0]);
#pragma unroll
          for (int r = 1; r < 10; ++r) gmin = fminf(gmin, __ldg(gyc + yall[r]));
          const bool inside = !act || (ex.o1 == ex.o0 + 1 && ey.o1 == ey.o0 + size1x);
          const bool all_inside = __all_sync(0xffffffffu, inside);
          const int gmin_bits = __reduce_min_sync(0xffffffffu, __float_as_int(gmin));   // int order == float order for values >= 0
          if (all_inside && gmin_bits > __float_as_int((float)Hl + 0.01f)) {
            if (act) {
#pragma unroll
              for (int bb = 0; bb < kRowsK; ++bb) tcol[bb * kRowsPitch + ql] = 0.f;
            }
            continue;
          }
        }
        // three window rows at a time: four rows of both grid channels (L1-resident; reloading the shared row per group
        // costs two loads and keeps the kernel at 64 registers = 4 CTAs per SM), first sampler, then the second sampler
        // with 12 plane loads per lane in flight
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const int *yrow = tyo0 + qq * 12 + 3 * g;   // rows 3g .. 3g+3 (entry 9 = o1 of row 8)
          const int yo[4] = {yrow[0], yrow[1], yrow[2], yrow[3]};
          float gx[4], gy[4], gxr[4], gyr[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) gx[r] = __ldg(gxc + yo[r]), gy[r] = __ldg(gyc + yo[r]);
#pragma unroll
          for (int r = 0; r < 4; ++r) gxr[r] = __shfl_down_sync(0xffffffffu, gx[r], 1), gyr[r] = __shfl_down_sync(0xffffffffu, gy[r], 1);
          float sx[3], iy[3], fy[3];
          int y0[3];
          bool in_y = false;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float2 w = wy[3 * g + j];
            const float w_nw = __fmul_rn(ex.w0, w.x), w_ne = __fmul_rn(ex.w1, w.x);
            const float w_sw = __fmul_rn(ex.w0, w.y), w_se = __fmul_rn(ex.w1, w.y);
            float x = __fmul_rn(gx[j], w_nw), y = __fmul_rn(gy[j], w_nw);
            x = __fmaf_rn(gxr[j], w_ne, x), y = __fmaf_rn(gyr[j], w_ne, y);
            x = __fmaf_rn(gx[j + 1], w_sw, x), y = __fmaf_rn(gy[j + 1], w_sw, y);
            x = __fmaf_rn(gxr[j + 1], w_se, x), y = __fmaf_rn(gyr[j + 1], w_se, y);
            sx[j] = x;
            iy[j] = sample_coord_x<kDiv>(y, axH, hH);
            fy[j] = floorf(iy[j]);
            y0[j] = (int)fy[j];
            in_y |= (unsigned)(y0[j] + 1) <= (unsigned)Hl;           // y0 in [-1, Hl-1]: at least one tap row inside
          }
          float val[3] = {0.f, 0.f, 0.f};
          float ix[3] = {0.f, 0.f, 0.f};
          // rows mapped entirely outside the (scale-mixed, smaller) level-l plane are exact zeros: most of levels 1-3
          if (kDbg || __any_sync(0xffffffffu, act && in_y)) {
            float fx[3];
            int x0[3];
            bool interior = true;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              ix[j] = sample_coord_x<kDiv>(remainder_sel(sx[j], axW, p.w2_pow2), axW, hW);
              fx[j] = floorf(ix[j]);
              x0[j] = (int)fx[j];
              interior &= (unsigned)x0[j] < (unsigned)(Wl - 1) && (unsigned)y0[j] < (unsigned)(Hl - 1);
            }
            float t00[3], t01[3], t10[3], t11[3];
            if (__all_sync(0xffffffffu, !act || interior)) {
              // all four taps of every lane in range: one base pointer, immediate offsets, no clamps
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const float *s0 = pl + (act ? y0[j] * Wl + x0[j] : 0);
                t00[j] = ld_hint(s0, pol), t01[j] = ld_hint(s0 + 1, pol), t10[j] = ld_hint(s0 + Wl, pol), t11[j] = ld_hint(s0 + Wl + 1, pol);
              }
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const float dxw = __fsub_rn(ix[j], fx[j]), dxe = __fsub_rn(__fadd_rn(fx[j], 1.f), ix[j]);
                const float dyn = __fsub_rn(iy[j], fy[j]), dys = __fsub_rn(__fadd_rn(fy[j], 1.f), iy[j]);
                float v = __fmul_rn(t00[j], __fmul_rn(dxe, dys));
                v = __fmaf_rn(t01[j], __fmul_rn(dxw, dys), v);
                v = __fmaf_rn(t10[j], __fmul_rn(dxe, dyn), v);
                val[j] = __fmaf_rn(t11[j], __fmul_rn(dxw, dyn), v);
              }
            } else {
              AxisEntry gxe[3], gye[3];
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                gxe[j] = make_axis_entry(ix[j], Wl, 1), gye[j] = make_axis_entry(iy[j], Hl, Wl);
                t00[j] = ld_hint(pl + gye[j].o0 + gxe[j].o0, pol), t01[j] = ld_hint(pl + gye[j].o0 + gxe[j].o1, pol);
                t10[j] = ld_hint(pl + gye[j].o1 + gxe[j].o0, pol), t11[j] = ld_hint(pl + gye[j].o1 + gxe[j].o1, pol);
              }
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                float v = __fmul_rn(t00[j], __fmul_rn(gxe[j].w0, gye[j].w0));
                v = __fmaf_rn(t01[j], __fmul_rn(gxe[j].w1, gye[j].w0), v);
                v = __fmaf_rn(t10[j], __fmul_rn(gxe[j].w0, gye[j].w1), v);
                val[j] = __fmaf_rn(t11[j], __fmul_rn(gxe[j].w1, gye[j].w1), v);
              }
            }
          }
          if (act) {
#pragma unroll
            for (int j = 0; j < 3; ++j) tcol[(3 * g + j) * kRowsPitch + ql] = val[j];
            if (kDbg && dq != nullptr) {
#pragma unroll
              for (int j = 0; j < 3; ++j) dq[2 * (3 * g + j)] = ix[j], dq[2 * (3 * g + j) + 1] = iy[j];
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- CTA write-out of the [81][96] tile
  const int nq = min(kRowsQ, p.N - n0);
  if (!BRANCH && !p.channels_last) {
    float *o = p.out_own + ((long long)b * p.L + lvl) * kRowsK2 * (long long)p.N + n0;       // NCHW rows
    for (int ch = warp; ch < kRowsK2; ch += kRowsWarps) {
      const float *trow = tile + ch * kRowsPitch;
      float *orow = o + (long long)ch * p.N;
#pragma unroll
      for (int j = 0; j < (kRowsQ + 31) / 32; ++j)
        if (j * 32 + lane < nq) orow[j * 32 + lane] = trow[j * 32 + lane];
    }
  } else {
    float *o = (BRANCH ? p.raw : p.out_own) + (((long long)b * p.N + n0) * p.L + lvl) * kRowsK2;   // channels-last rows
    for (int q = warp; q < nq; q += kRowsWarps) {
      float *orow = o + (long long)q * p.L * kRowsK2;
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (j * 32 + lane < kRowsK2) orow[j * 32 + lane] = tile[(j * 32 + lane) * kRowsPitch + q];
    }
  }
}

template <int kDiv, bool kDbg, int kRowsQ, int kMinCtas>
__global__ void __launch_bounds__(kRowsThreads, kMinCtas) lookup_rows_kernel(const LookupParams p) {
  // heaviest CTAs first: other view level 0..L-1, then own view
  const int y = blockIdx.y;
  // lets the dependent rotate kernel be scheduled as soon as the LAST CTA of this grid has started (it still waits for
  // this grid's completion before reading anything this grid writes)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (p.dual && y < p.L)
    lookup_rows_body<kDiv, kDbg, kRowsQ, 1>(p, y);
  else
    lookup_rows_body<kDiv, kDbg, kRowsQ, 0>(p, p.dual ? y - p.L : y);
}

static size_t lookup_rows_smem_bytes(int Q) {
  return (size_t)(((kRowsK2 * (Q | 1) + 3) & ~3) + kRowsWarps * kRowsWarpWords) * sizeof(float);
}

// Default shape of the kernel; PF_LOOKUP_Q / PF_LOOKUP_OCC select the other compiled shapes for A/B timing.
#ifndef PF_ROWS_Q
#define PF_ROWS_Q 48
#endif

template <int kDiv, bool kDbg>
static void launch_rows_shape(const LookupParams &p, bool dual, int q, int occ, cudaStream_t st) {
#define PF_ROWS_SHAPE(Q, OCC)                                                                       \
  if (q == Q && occ == OCC) {                                                                       \
    dim3 grid(ceil_div(p.N, Q), p.L * (dual ? 2 : 1), p.B);                                         \
    lookup_rows_kernel<kDiv, kDbg, Q, OCC><<<grid, kRowsThreads, lookup_rows_smem_bytes(Q), st>>>(p); \
    return;                                                                                         \
  }
  if constexpr (kDiv == PF_DIV_ATEN_CUDA && !kDbg) {   // the tuning shapes exist for the production flavour only
    PF_ROWS_SHAPE(96, 3) PF_ROWS_SHAPE(96, 4) PF_ROWS_SHAPE(48, 3) PF_ROWS_SHAPE(48, 4) PF_ROWS_SHAPE(48, 5) PF_ROWS_SHAPE(48, 6)
    PF_ROWS_SHAPE(72, 4) PF_ROWS_SHAPE(72, 5) PF_ROWS_SHAPE(24, 6)
  }
  q = PF_ROWS_Q, occ = PF_ROWS_MIN_CTAS;
  PF_ROWS_SHAPE(PF_ROWS_Q, PF_ROWS_MIN_CTAS)
#undef PF_ROWS_SHAPE
}

static int launch_lookup_rows(const LookupParams &p, bool dual, cudaStream_t st, const char *who) {
  static const int q = getenv("PF_LOOKUP_Q") ? atoi(getenv("PF_LOOKUP_Q")) : PF_ROWS_Q;
  static const int occ = getenv("PF_LOOKUP_OCC") ? atoi(getenv("PF_LOOKUP_OCC")) : PF_ROWS_MIN_CTAS;
  const bool recip = p.div_mode == PF_DIV_ATEN_CUDA;
  const bool dbg = p.dbg_own != nullptr || p.dbg_other != nullptr;
  if (recip) {
    if (dbg) launch_rows_shape<PF_DIV_ATEN_CUDA, true>(p, dual, q, occ, st); else launch_rows_shape<PF_DIV_ATEN_CUDA, false>(p, dual, q, occ, st);
  } else {
    if (dbg) launch_rows_shape<PF_DIV_IEEE, true>(p, dual, q, occ, st); else launch_rows_shape<PF_DIV_IEEE, false>(p, dual, q, occ, st);
  }
  return check_launch(who);
}

// ------------------------------------------------------------------------------------------------
// img_rotate of the channels-last pre-rotation map: out[b, c, p] = sum_t w_t(p) raw[b, src_t(p), c].
constexpr int kRotThreads = 256;
constexpr int kRotPixels = 32;     // backward / scalar kernel
constexpr int kRotFwdPixels = 16;  // forward float4 kernel: two pixels per warp, 512 CTAs at 64x128

struct RotateParams {
  int B, N, h, w, L, K2, div_mode, channels_last, fuse_sum;
  Axis axW, axH;
  const float *grid_c2w;
  long long grid_bs;
  const float *raw;   // fwd: [B, N, L*K2] in.   bwd: unused
  float *out;         // fwd: out_other (or out_own when fuse_sum), [B, L*K2, N] or channels-last.  bwd: incoming gradient
  float *draw;        // bwd: [B, N, L*K2] (+=)
  const float *own_cl;  // fwd, NCHW + fuse_sum: the own view staged channels-last [B, N, L*K2]; added before the single write pass
};

// Forward: one CTA = 16 output pixels x ALL channels, a warp = 2 pixels.  The channels-last map makes a pixel's L*K2
// floats one contiguous, 16-byte aligned vector, so lanes walk float4s: 4 taps x 2 pixels = 8 LDG.128 in flight per
// lane and round.  Channels-last output (and the fused `own + other` add) is written directly as float4.  NCHW output
// goes through a pixel-major shared tile [16][C + 8] — conflict-free STS.128 in, 64-byte channel rows out, two rows
// per warp instruction — instead of the r01 version's channel-major [C][9] tile (4-way conflicts on the way in, 32-byte
// rows and a division per element on the way out: ncu r01e, 3.7 M instructions for 2.65 M outputs).
__global__ void __launch_bounds__(kRotThreads) rotate_fwd_kernel(const RotateParams p) {
  extern __shared__ __align__(16) float tile[];  // [kRotFwdPixels][Cp] (NCHW only), then one Taps4 per pixel
  const int C = p.L * p.K2, C4 = C >> 2, Cp = C + 8;
  Taps4 *taps = reinterpret_cast<Taps4 *>(tile + (p.channels_last ? 0 : kRotFwdPixels * Cp));
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kRotFwdPixels;
  if (threadIdx.x < kRotFwdPixels) {
    // module-local cyclic sampler of projection_prim_ortho.py:119-135 on the [.., h, w] map
    const float *gx = p.grid_c2w + (long long)b * p.grid_bs;
    const int n = min(n0 + (int)threadIdx.x, p.N - 1);
    const float x = remainder_pos(__ldg(gx + n), p.axW.size);
    const float y = __ldg(gx + p.N + n);
    Taps4 t = clamp_taps(make_taps(to_sample_coord(x, p.axW, p.div_mode), to_sample_coord(y, p.axH, p.div_mode)), p.h, p.w);
    t.o_nw *= C4, t.o_ne *= C4, t.o_sw *= C4, t.o_se *= C4;   // float4 offsets into the channels-last map
    taps[threadIdx.x] = t;
  }
  // programmatic dependent launch: everything above reads only grid_c2w (an input), so this CTA may be resident and
  // have its taps ready while the last CTAs of lookup_rows_kernel are still running; raw / own_cl / out are touched
  // only after the primary grid has completed and flushed (no-op when launched without the attribute)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();
  const float4 *src = reinterpret_cast<const float4 *>(opaque(p.raw + (long long)b * p.N * C));
  float4 *out_cl = reinterpret_cast<float4 *>(opaque(p.out + (long long)b * p.N * C));
  constexpr int kPerWarp = kRotFwdPixels / (kRotThreads / 32);   // 2
  const int q0 = warp * kPerWarp;
  Taps4 t[kPerWarp];
#pragma unroll
  for (int i = 0; i < kPerWarp; ++i) t[i] = taps[q0 + i];
  for (int c4 = lane; c4 < C4; c4 += 32) {
    float4 r[kPerWarp];
#pragma unroll
    for (int i = 0; i < kPerWarp; ++i) {
      const float4 a = ld_f4(src + t[i].o_nw + c4), bq = ld_f4(src + t[i].o_ne + c4);   // never .nc: the producer grid may
      const float4 c = ld_f4(src + t[i].o_sw + c4), d = ld_f4(src + t[i].o_se + c4);    // still be running when this kernel starts
      r[i].x = __fmaf_rn(d.x, t[i].se, __fmaf_rn(c.x, t[i].sw, __fmaf_rn(bq.x, t[i].ne, __fmul_rn(a.x, t[i].nw))));
      r[i].y = __fmaf_rn(d.y, t[i].se, __fmaf_rn(c.y, t[i].sw, __fmaf_rn(bq.y, t[i].ne, __fmul_rn(a.y, t[i].nw))));
      r[i].z = __fmaf_rn(d.z, t[i].se, __fmaf_rn(c.z, t[i].sw, __fmaf_rn(bq.z, t[i].ne, __fmul_rn(a.z, t[i].nw))));
      r[i].w = __fmaf_rn(d.w, t[i].se, __fmaf_rn(c.w, t[i].sw, __fmaf_rn(bq.w, t[i].ne, __fmul_rn(a.w, t[i].nw))));
    }
#pragma unroll
    for (int i = 0; i < kPerWarp; ++i) {
      if (p.channels_last) {
        if (n0 + q0 + i < p.N) {
          float4 *o = out_cl + (long long)(n0 + q0 + i) * C4 + c4;
          if (p.fuse_sum) {   // corr_A + corr_B_A (core/prior_raft.py:187)
            const float4 own = *o;
            r[i].x = __fadd_rn(own.x, r[i].x), r[i].y = __fadd_rn(own.y, r[i].y);
            r[i].z = __fadd_rn(own.z, r[i].z), r[i].w = __fadd_rn(own.w, r[i].w);
          }
          *o = r[i];
        }
      } else {
        if (p.own_cl != nullptr) {   // corr_A + corr_B_A with the own view read as one coalesced float4
          const float4 own = ld_f4(reinterpret_cast<const float4 *>(p.own_cl + (long long)b * p.N * C) + (long long)min(n0 + q0 + i, p.N - 1) * C4 + c4);
          r[i].x = __fadd_rn(own.x, r[i].x), r[i].y = __fadd_rn(own.y, r[i].y);
          r[i].z = __fadd_rn(own.z, r[i].z), r[i].w = __fadd_rn(own.w, r[i].w);
        }
        *reinterpret_cast<float4 *>(tile + (q0 + i) * Cp + 4 * c4) = r[i];
      }
    }
  }
  if (!p.channels_last) {
    __syncthreads();
    // 64-byte channel rows: lanes 0-15 write row c, lanes 16-31 row c + 1
    const int px = lane & (kRotFwdPixels - 1), par = lane >> 4;
    if (n0 + px < p.N) {
      float *o = opaque(p.out + (long long)b * C * p.N + n0 + px);
      const float *tp = tile + px * Cp;
#pragma unroll 4
      for (int c = 2 * warp + par; c < C; c += 2 * (kRotThreads / 32)) {
        float *oc = o + (long long)c * p.N;
        *oc = (p.fuse_sum && p.own_cl == nullptr) ? __fadd_rn(*oc, tp[c]) : tp[c];
      }
    }
  }
}

static size_t rotate_fwd_smem_bytes(int C, int channels_last) {
  return (channels_last ? 0 : (size_t)kRotFwdPixels * (C + 8) * sizeof(float)) + kRotFwdPixels * sizeof(Taps4);
}

template <bool kBwd>
__global__ void __launch_bounds__(kRotThreads) rotate_kernel(const RotateParams p) {
  extern __shared__ float tile[];  // [K2][33] transpose tile (NCHW only), then one Taps4 per pixel
  Taps4 *taps = reinterpret_cast<Taps4 *>(tile + p.K2 * 33 + (p.K2 & 1));
  const int lvl = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kRotPixels;
  const int C = p.L * p.K2;
  float *io = p.out + ((long long)b * C + (long long)lvl * p.K2) * p.N + n0;   // NCHW rows
  if (threadIdx.x < kRotPixels && n0 + threadIdx.x < p.N) {
    // module-local cyclic sampler of projection_prim_ortho.py:119-135 on the [.., h, w] map
    const float *gx = p.grid_c2w + (long long)b * p.grid_bs;
    const int n = n0 + threadIdx.x;
    const float x = remainder_pos(__ldg(gx + n), p.axW.size);
    const float y = __ldg(gx + p.N + n);
    Taps4 t = clamp_taps(make_taps(to_sample_coord(x, p.axW, p.div_mode), to_sample_coord(y, p.axH, p.div_mode)), p.h, p.w);
    t.o_nw *= C, t.o_ne *= C, t.o_sw *= C, t.o_se *= C;   // element offsets into the channels-last map
    taps[threadIdx.x] = t;
  }
  if constexpr (kBwd) {
    if (!p.channels_last && n0 + lane < p.N)
      for (int ch = warp; ch < p.K2; ch += kRotThreads / 32) tile[ch * 33 + lane] = io[(long long)ch * p.N + lane];
  }
  __syncthreads();
  const long long base = (long long)b * p.N * C + (long long)lvl * p.K2;
  const float *src = kBwd ? nullptr : opaque(p.raw + base);
  float *dst = kBwd ? opaque(p.draw + base) : nullptr;
  float *out_cl = opaque(p.out + base);   // channels-last rows: + n * C + ch
#pragma unroll 1
  for (int qi = 0; qi < kRotPixels / (kRotThreads / 32); ++qi) {
    const int q = warp * (kRotPixels / (kRotThreads / 32)) + qi;
    if (n0 + q >= p.N) break;
    const Taps4 t = taps[q];
    float *oq = out_cl + (long long)(n0 + q) * C;
    if constexpr (kBwd) {
      for (int ch = lane; ch < p.K2; ch += 32) {
        const float g = p.channels_last ? __ldg(oq + ch) : tile[ch * 33 + q];
        if (t.nw != 0.f) atomicAdd(dst + t.o_nw + ch, g * t.nw);
        if (t.ne != 0.f) atomicAdd(dst + t.o_ne + ch, g * t.ne);
        if (t.sw != 0.f) atomicAdd(dst + t.o_sw + ch, g * t.sw);
        if (t.se != 0.f) atomicAdd(dst + t.o_se + ch, g * t.se);
      }
    } else {
#pragma unroll 3
      for (int ch = lane; ch < p.K2; ch += 32) {
        float acc = __fmul_rn(__ldg(src + t.o_nw + ch), t.nw);
        acc = __fmaf_rn(__ldg(src + t.o_ne + ch), t.ne, acc);
        acc = __fmaf_rn(__ldg(src + t.o_sw + ch), t.sw, acc);
        acc = __fmaf_rn(__ldg(src + t.o_se + ch), t.se, acc);
        if (p.channels_last)
          oq[ch] = p.fuse_sum ? __fadd_rn(oq[ch], acc) : acc;   // corr_A + corr_B_A (core/prior_raft.py:187)
        else
          tile[ch * 33 + q] = acc;
      }
    }
  }
  if constexpr (!kBwd) {
    if (!p.channels_last) {
      __syncthreads();
      if (n0 + lane < p.N)
        for (int ch = warp; ch < p.K2; ch += kRotThreads / 32) {
          float *o = io + (long long)ch * p.N + lane;
          *o = p.fuse_sum ? __fadd_rn(*o, tile[ch * 33 + lane]) : tile[ch * 33 + lane];
        }
    }
  }
}

static size_t rotate_smem_bytes(int K2) {
  return ((size_t)K2 * 33 + (K2 & 1)) * sizeof(float) + kRotPixels * sizeof(Taps4);
}

static int fill_lookup_params(const pf_lookup_args *a, LookupParams &p, bool dual, const char *who) {
  PF_REQUIRE(a->batch > 0 && a->h > 0 && a->w > 0 && a->h2 > 0 && a->w2 > 0, "%s: bad shape", who);
  PF_REQUIRE(a->num_levels >= 1 && a->num_levels <= PF_MAX_LEVELS, "%s: num_levels must be 1..%d", who, PF_MAX_LEVELS);
  PF_REQUIRE(a->radius >= 0 && a->radius <= 7, "%s: radius must be 0..7", who);
  PF_REQUIRE(a->coords != nullptr, "%s: coords is required", who);
  PF_REQUIRE((a->h2 >> (a->num_levels - 1)) >= 1 && (a->w2 >> (a->num_levels - 1)) >= 1,
             "%s: pyramid too deep for %dx%d", who, a->h2, a->w2);
  p.B = a->batch;
  p.h = a->h;
  p.w = a->w;
  p.N = a->h * a->w;
  p.radius = a->radius;
  p.L = a->num_levels;
  p.cyclic = a->cyclic;
  p.div_mode = a->div_mode;
  p.dual = dual;
  p.w_pow2 = (a->w & (a->w - 1)) == 0;
  p.w2_pow2 = (a->w2 & (a->w2 - 1)) == 0 && (a->w2 >> (a->num_levels - 1)) >= 1;
  p.q_begin = 0;
  p.q_end = p.N;
  static const bool l2_hint = getenv("PF_LOOKUP_L2HINT") != nullptr && getenv("PF_LOOKUP_L2HINT")[0] == '1';   // A/B (r03)
  p.l2_hint = l2_hint;
  p.channels_last = a->out_channels_last;
  p.fuse_sum = a->fuse_sum;
  p.coords = a->coords;
  for (int l = 0; l < PF_MAX_LEVELS; ++l) {
    p.own[l] = l < p.L ? a->own[l] : nullptr;
    p.other[l] = l < p.L ? a->other[l] : nullptr;
    p.d_own[l] = p.d_other[l] = nullptr;
    p.Hl[l] = a->h2 >> l;
    p.Wl[l] = a->w2 >> l;
    p.axH[l] = make_axis(p.Hl[l] > 0 ? p.Hl[l] : 1);
    p.axW[l] = make_axis(p.Wl[l] > 0 ? p.Wl[l] : 1);
  }
  p.ax_gw = make_axis(a->w);
  p.ax_gh = make_axis(a->h);
  p.grid_w2c = a->grid_w2c;
  p.grid_bs = a->grid_batch_stride;
  p.out_own = a->out_own;
  p.raw = a->scratch;
  p.dbg_own = a->dbg_own_xy;
  p.dbg_other = a->dbg_other_xy;
  if (dual) {
    PF_REQUIRE(a->grid_w2c && a->grid_c2w && a->scratch, "%s: dual lookup needs grid_w2c, grid_c2w and scratch", who);
    PF_REQUIRE(a->cyclic, "%s: the dual (DCCL) lookup is defined for the cyclic sampler only", who);
  }
  return 0;
}

static void fill_rotate_params(const pf_lookup_args *a, RotateParams &rp) {
  const int k = 2 * a->radius + 1;
  rp.B = a->batch;
  rp.N = a->h * a->w;
  rp.h = a->h;
  rp.w = a->w;
  rp.L = a->num_levels;
  rp.K2 = k * k;
  rp.div_mode = a->div_mode;
  rp.channels_last = a->out_channels_last;
  rp.fuse_sum = a->fuse_sum;
  rp.axW = make_axis(a->w);
  rp.axH = make_axis(a->h);
  rp.grid_c2w = a->grid_c2w;
  rp.grid_bs = a->grid_batch_stride;
  rp.raw = a->scratch;
  rp.out = a->fuse_sum ? a->out_own : a->out_other;
  rp.draw = nullptr;
  rp.own_cl = nullptr;
}

template <bool kBwd>
static int launch_lookup(const LookupParams &p, int radius, bool dual, cudaStream_t st, const char *who) {
  const int k = 2 * radius + 1, K2 = k * k;
  dim3 grid(ceil_div(p.q_end - p.q_begin, kQueriesPerCta), p.L * (dual ? 2 : 1), p.B);
  const size_t smem = ((size_t)(kLookupThreads / 32) * lookup_warp_floats(k) + (size_t)K2 * 33) * sizeof(float);
  const bool recip = p.div_mode == PF_DIV_ATEN_CUDA;
  if (radius == 4) {
    if (recip)
      lookup_kernel<4, kBwd, PF_DIV_ATEN_CUDA><<<grid, kLookupThreads, smem, st>>>(p);
    else
      lookup_kernel<4, kBwd, PF_DIV_IEEE><<<grid, kLookupThreads, smem, st>>>(p);
  } else {
    if (recip)
      lookup_kernel<0, kBwd, PF_DIV_ATEN_CUDA><<<grid, kLookupThreads, smem, st>>>(p);
    else
      lookup_kernel<0, kBwd, PF_DIV_IEEE><<<grid, kLookupThreads, smem, st>>>(p);
  }
  return check_launch(who);
}

// img_rotate of a channels-last pre-rotation map, shared with the on-the-fly path (pf_onthefly.cu).
int rotate_forward(int batch, int h, int w, int num_levels, int radius, int div_mode, const float *grid_c2w,
                   long long grid_bs, const float *raw, float *out, int channels_last, int fuse_sum, cudaStream_t st,
                   const float *own_cl, bool after_lookup_rows) {
  const int k = 2 * radius + 1;
  RotateParams rp;
  rp.B = batch;
  rp.N = h * w;
  rp.h = h;
  rp.w = w;
  rp.L = num_levels;
  rp.K2 = k * k;
  rp.div_mode = div_mode;
  rp.channels_last = channels_last;
  rp.fuse_sum = fuse_sum;
  rp.axW = make_axis(w);
  rp.axH = make_axis(h);
  rp.grid_c2w = grid_c2w;
  rp.grid_bs = grid_bs;
  rp.raw = raw;
  rp.out = out;
  rp.draw = nullptr;
  rp.own_cl = (fuse_sum && !channels_last) ? own_cl : nullptr;
  const int C = rp.L * rp.K2;
  if (C % 4 == 0 && (((uintptr_t)raw | (uintptr_t)out) & 15) == 0) {
    const size_t smem = rotate_fwd_smem_bytes(C, channels_last);
    if (smem > 48 * 1024) cudaFuncSetAttribute(rotate_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // PF_ROTATE_PDL=0 disables the programmatic dependent launch (A/B timing / old drivers)
    static const bool pdl = !(getenv("PF_ROTATE_PDL") != nullptr && getenv("PF_ROTATE_PDL")[0] == '0');
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ceil_div(rp.N, kRotFwdPixels), rp.B);
    cfg.blockDim = dim3(kRotThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && after_lookup_rows) ? 1 : 0;
    if (cudaLaunchKernelEx(&cfg, rotate_fwd_kernel, rp) != cudaSuccess) {
      (void)cudaGetLastError();   // e.g. a driver without programmatic dependent launch: plain launch
      rotate_fwd_kernel<<<cfg.gridDim, cfg.blockDim, smem, st>>>(rp);
    }
    return check_launch("rotate_fwd_kernel");
  }
  dim3 grid(ceil_div(rp.N, kRotPixels), rp.L, rp.B);   // odd channel counts: scalar per-level kernel
  rotate_kernel<false><<<grid, kRotThreads, rotate_smem_bytes(rp.K2), st>>>(rp);
  return check_launch("rotate_kernel");
}

}  // namespace pf

extern "C" int pf_lookup_dual(const pf_lookup_args *a, void *stream) {
  using namespace pf;
  PF_REQUIRE(a != nullptr, "pf_lookup_dual: null args");
  const bool dual = a->other[0] != nullptr;
  LookupParams p;
  if (int e = fill_lookup_params(a, p, dual, "pf_lookup_dual")) return e;
  PF_REQUIRE(a->out_own != nullptr, "pf_lookup_dual: out_own is required");
  PF_REQUIRE(!dual || a->fuse_sum || a->no_rotate || a->out_other != nullptr, "pf_lookup_dual: out_other is required unless fuse_sum / no_rotate");
  PF_REQUIRE(!a->no_rotate || (dual && a->out_channels_last && !a->fuse_sum), "pf_lookup_dual: no_rotate needs the dual lookup with channels-last outputs and no fuse_sum");
  for (int l = 0; l < p.L; ++l) {
    PF_REQUIRE(a->own[l] != nullptr, "pf_lookup_dual: own[%d] is null", l);
    PF_REQUIRE(!dual || a->other[l] != nullptr, "pf_lookup_dual: other[%d] is null", l);
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int C_all = p.L * (2 * a->radius + 1) * (2 * a->radius + 1);
  // radius 4 (the model's) runs the column-walk kernel; other radii the generic tap-per-lane kernel
  // (PF_LOOKUP_LEGACY=1 forces the latter, for A/B timing only)
  static const bool legacy = getenv("PF_LOOKUP_LEGACY") != nullptr && getenv("PF_LOOKUP_LEGACY")[0] == '1';
  // fused sum into an NCHW tensor: stage the own view channels-last (coalesced 324-byte runs, no transpose tile) and
  // let the rotate kernel add it while it transposes — one write pass over out_own instead of write + read-modify-write
  const bool stage_own = dual && a->fuse_sum && !a->out_channels_last && a->scratch_own != nullptr && C_all % 4 == 0 &&
                         (((uintptr_t)a->scratch_own | (uintptr_t)a->scratch | (uintptr_t)a->out_own) & 15) == 0;
  if (stage_own) {
    p.channels_last = 1;
    p.out_own = a->scratch_own;
  }
  if (a->radius == 4 && !legacy) {
    if (int e = launch_lookup_rows(p, dual, st, "pf_lookup_dual")) return e;
  } else {
    if (int e = launch_lookup<false>(p, a->radius, dual, st, "pf_lookup_dual")) return e;
  }
  if (dual && !a->no_rotate) {
    // core/corr.py:137-138 — img_rotate of the [B, L*81, h, w] map with grid_c2w.
    return rotate_forward(a->batch, a->h, a->w, a->num_levels, a->radius, a->div_mode, a->grid_c2w,
                          a->grid_batch_stride, a->scratch, a->fuse_sum ? a->out_own : a->out_other, a->out_channels_last,
                          a->fuse_sum, st, stage_own ? a->scratch_own : nullptr, a->radius == 4 && !legacy);
  }
  return 0;
}

// (e) d(lookup)/d(pyramids).  The orthogonal branch first runs the adjoint of img_rotate
// (scatter of grad_other into the zeroed channels-last scratch map), then both branches scatter
// through the forward's coordinates.  A query's plane is private to it, so the atomics only ever
// collide between taps of one window.
extern "C" int pf_lookup_dual_bwd(const pf_lookup_bwd_args *ba, void *stream) {
  using namespace pf;
  PF_REQUIRE(ba != nullptr, "pf_lookup_dual_bwd: null args");
  const pf_lookup_args *a = &ba->fwd;
  const bool dual = ba->grad_other != nullptr;
  LookupParams p;
  if (int e = fill_lookup_params(a, p, dual, "pf_lookup_dual_bwd")) return e;
  p.fuse_sum = 0;
  PF_REQUIRE(ba->grad_own != nullptr, "pf_lookup_dual_bwd: grad_own is required");
  for (int l = 0; l < p.L; ++l) {
    PF_REQUIRE(ba->dgrad_own[l] != nullptr, "pf_lookup_dual_bwd: dgrad_own[%d] is null", l);
    PF_REQUIRE(!dual || ba->dgrad_other[l] != nullptr, "pf_lookup_dual_bwd: dgrad_other[%d] is null", l);
    p.d_own[l] = ba->dgrad_own[l];
    p.d_other[l] = ba->dgrad_other[l];
  }
  p.dbg_own = p.dbg_other = nullptr;
  p.out_own = const_cast<float *>(ba->grad_own);
  if (ba->query_count > 0) {
    PF_REQUIRE(ba->query_begin >= 0 && ba->query_begin + ba->query_count <= p.N, "pf_lookup_dual_bwd: query range [%d, +%d) outside 0..%d",
               ba->query_begin, ba->query_count, p.N);
    p.q_begin = ba->query_begin;
    p.q_end = ba->query_begin + ba->query_count;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (dual && !ba->scratch_ready) {
    RotateParams rp;
    fill_rotate_params(a, rp);
    const size_t bytes = (size_t)a->batch * rp.L * rp.K2 * rp.N * sizeof(float);
    if (cudaMemsetAsync(a->scratch, 0, bytes, st) != cudaSuccess) return check_launch("pf_lookup_dual_bwd(memset)");
    rp.out = const_cast<float *>(ba->grad_other);
    rp.fuse_sum = 0;
    rp.draw = a->scratch;
    dim3 grid(ceil_div(rp.N, kRotPixels), rp.L, rp.B);
    rotate_kernel<true><<<grid, kRotThreads, rotate_smem_bytes(rp.K2), st>>>(rp);
    if (int e = check_launch("pf_lookup_dual_bwd(rotate)")) return e;
  }
  return launch_lookup<true>(p, a->radius, dual, st, "pf_lookup_dual_bwd");
}
