// (b) Dual-cost pyramid lookup — DCCL.__call__ (PriOr-RAFT/core/corr.py:113-144) and
// CorrBlock.__call__ (core/corr.py:30-51): one gather launch serving both views + one rotate launch.
//
// lookup_kernel — grid = (ceil(N/32) query chunks, levels x branches, batch), 256 threads.
//   A CTA owns 32 consecutive query pixels of one level of one branch; each warp walks 4 queries.
//   * Window coordinates are separable: the 2r+1 x-coordinates and 2r+1 y-coordinates of a window
//     go through the (remainder, normalise, unnormalise, floor) chain once each, on lanes 0..2k-1,
//     and the (2r+1)^2 taps fetch theirs with warp shuffles.  That chain — not memory — was the
//     limiter of the first version (ncu r01a: 76 % issue-slot utilisation, 13 % DRAM).
//   * Branch 0 (own view): taps are ordered y-major across lanes (x fastest), so a warp-wide load
//     touches ~4 rows of <=10 contiguous floats of the query's private plane; L1 serves the
//     four-corner overlap and every DRAM sector is fetched once.  Results are transposed through
//     shared memory and written as full 128-byte rows of the [B, L*81, h, w] output.
//   * Branch 1 (other view): the window is mapped through the LEVEL-0 rotation grid (8 L1-resident
//     loads), then pyr_other[l][n] is sampled at the mapped point (scale mixing is the
//     reference's).  The pre-rotation map goes to `scratch` CHANNELS-LAST
//     ([B, N, L*81]) so that rotate_kernel reads whole 1296-byte vectors.
// rotate_kernel — img_rotate(., grid_c2w) of that map (core/corr.py:137-138): a cross-pixel
//   gather, hence a second pass; 32 output pixels x one level per CTA, lanes across channels, four
//   coalesced vector reads per pixel (the 10.6 MB intermediate stays in L2), shared-memory
//   transpose, 128-byte output rows.
// Both kernels have a kBwd instantiation (scatter instead of gather) for training.
// Coordinates are bit-exact restatements (pf_common.cuh); values are ATen's FMA chain.
#include "pf_common.cuh"

namespace pf {

constexpr int kQueriesPerCta = 32;
constexpr int kLookupThreads = 256;
constexpr int kQueriesPerWarp = kQueriesPerCta / (kLookupThreads / 32);

struct LookupParams {
  int B, N, h, w;  // query grid, N = h*w
  int radius, L, cyclic, div_mode, dual;
  const float *coords;
  const float *own[PF_MAX_LEVELS];
  const float *other[PF_MAX_LEVELS];
  int Hl[PF_MAX_LEVELS], Wl[PF_MAX_LEVELS];
  Axis axW[PF_MAX_LEVELS], axH[PF_MAX_LEVELS];
  Axis ax_gw, ax_gh;  // axes of the rotation grid (query resolution)
  const float *grid_w2c;
  long long grid_bs;
  float *out_own;                      // forward: [B, L*K2, N] output.  backward: incoming gradient (read only)
  float *raw;                          // forward: [B, N, L*K2] pre-rotation map.  backward: its gradient (read only)
  float *dbg_own, *dbg_other;
  float *d_own[PF_MAX_LEVELS];         // backward: gradient pyramids (+=)
  float *d_other[PF_MAX_LEVELS];
};

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

template <int R, bool kBwd>
__global__ void __launch_bounds__(kLookupThreads) lookup_kernel(const LookupParams p) {
  const int r = (R > 0) ? R : p.radius;
  const int k = 2 * r + 1;
  const int K2 = k * k;
  extern __shared__ float tile[];  // [K2][33], own-view branch only
  const int lvl = blockIdx.y % p.L;
  const int branch = blockIdx.y / p.L;
  const int b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kQueriesPerCta;
  const int Hl = p.Hl[lvl], Wl = p.Wl[lvl];
  const Axis axW = p.axW[lvl], axH = p.axH[lvl];
  const float inv_scale = 1.0f / (float)(1 << lvl);  // `coords / 2**i` is exact either way
  const float *vol = branch ? p.other[lvl] : p.own[lvl];
  const float *gridx = p.grid_w2c + (long long)b * p.grid_bs;
  const float *gridy = gridx + p.N;
  float *dbg = branch ? p.dbg_other : p.dbg_own;
  float *io = p.out_own + ((long long)b * p.L + lvl) * K2 * (long long)p.N + n0;
  if constexpr (kBwd) {  // own branch: stage the incoming gradient tile [K2][32 queries], coalesced rows
    if (branch == 0) {
      if (n0 + lane < p.N)
        for (int ch = warp; ch < K2; ch += kLookupThreads / 32) tile[ch * 33 + lane] = io[(long long)ch * p.N + lane];
      __syncthreads();
    }
  }
  // this lane's window axis: lanes [0,k) hold x offsets, lanes [k,2k) y offsets
  const bool is_x = lane < k;
  const int off = (is_x ? lane : lane - k) - r;
  const Axis ax1 = branch ? (is_x ? p.ax_gw : p.ax_gh) : (is_x ? axW : axH);  // first sampler's axis
  const bool wrap1 = is_x && (branch || p.cyclic);

#pragma unroll 1
  for (int qi = 0; qi < kQueriesPerWarp; ++qi) {
    const int q = warp * kQueriesPerWarp + qi;
    const int n = n0 + q;
    if (n >= p.N) break;
    const float c = __fmul_rn(__ldg(p.coords + ((long long)b * 2 + (is_x ? 0 : 1)) * p.N + n), inv_scale);
    // core/corr.py:123-126 then the sampler's coordinate chain, once per window row / column
    float pc = __fadd_rn(c, (float)off);
    if (wrap1) pc = remainder_pos(pc, ax1.size);
    const float sc = to_sample_coord(pc, ax1, p.div_mode);
    const float fl = floorf(sc);
    const int i0 = (int)fl;
    const float w_hi = __fsub_rn(sc, fl);                  // ix - ix_nw
    const float w_lo = __fsub_rn(__fadd_rn(fl, 1.f), sc);  // ix_se - ix
    const long long plane_off = ((long long)b * p.N + n) * (long long)(Hl * Wl);
    const float *plane = kBwd ? nullptr : vol + plane_off;
    float *dplane = kBwd ? (branch ? p.d_other[lvl] : p.d_own[lvl]) + plane_off : nullptr;
    float *rawq = p.raw + (((long long)b * p.N + n) * p.L + lvl) * K2;

    auto do_tap = [&](int t) {
      const bool live = t < K2;
      const int tt = live ? t : 0;
      // own view: lanes walk x fastest (coalesced plane rows); other view: lanes walk the output channel
      const int hi = tt / k, lo = tt - hi * k;
      const int aa = branch ? hi : lo, bb = branch ? lo : hi;
      Taps tp;
      tp.x0 = __shfl_sync(0xffffffffu, i0, aa);
      tp.y0 = __shfl_sync(0xffffffffu, i0, k + bb);
      const float dxe = __shfl_sync(0xffffffffu, w_lo, aa), dxw = __shfl_sync(0xffffffffu, w_hi, aa);
      const float dys = __shfl_sync(0xffffffffu, w_lo, k + bb), dyn = __shfl_sync(0xffffffffu, w_hi, k + bb);
      float ix = 0.f, iy = 0.f;
      if (dbg != nullptr) {
        ix = __shfl_sync(0xffffffffu, sc, aa);
        iy = __shfl_sync(0xffffffffu, sc, k + bb);
      }
      if (!live) return;
      tp.nw = __fmul_rn(dxe, dys);
      tp.ne = __fmul_rn(dxw, dys);
      tp.sw = __fmul_rn(dxe, dyn);
      tp.se = __fmul_rn(dxw, dyn);
      const int ch = aa * k + bb;  // x-major channel order of the reference
      if (branch) {
        // core/corr.py:132-136 — map through the level-0 rotation grid, then index the level-l volume
        const Taps4 tg = clamp_taps(tp, p.h, p.w);
        const float sx = blend4(gridx, tg), sy = blend4(gridy, tg);
        ix = to_sample_coord(remainder_pos(sx, axW.size), axW, p.div_mode);
        iy = to_sample_coord(sy, axH, p.div_mode);
        tp = make_taps(ix, iy);
      }
      if constexpr (kBwd) {
        scatter_zeros(dplane, Hl, Wl, tp, branch ? __ldg(rawq + ch) : tile[ch * 33 + q]);
      } else {
        const float val = blend4(plane, clamp_taps(tp, Hl, Wl));
        if (branch)
          rawq[ch] = val;
        else
          tile[ch * 33 + q] = val;
        if (dbg != nullptr) {
          float *d = dbg + ((((long long)b * p.N + n) * p.L + lvl) * K2 + ch) * 2;
          d[0] = ix;
          d[1] = iy;
        }
      }
    };
    if constexpr (R > 0) {
#pragma unroll
      for (int it = 0; it < (K2 + 31) / 32; ++it) do_tap(it * 32 + lane);
    } else {
      for (int t0 = 0; t0 < K2; t0 += 32) do_tap(t0 + lane);
    }
  }
  if constexpr (!kBwd) {
    if (branch == 0) {
      __syncthreads();
      if (n0 + lane < p.N)
        for (int ch = warp; ch < K2; ch += kLookupThreads / 32) io[(long long)ch * p.N + lane] = tile[ch * 33 + lane];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// img_rotate of the channels-last pre-rotation map: out[b, c, p] = sum_t w_t(p) raw[b, src_t(p), c].
constexpr int kRotThreads = 256;
constexpr int kRotPixels = 32;

struct RotateParams {
  int B, N, h, w, L, K2, div_mode;
  Axis axW, axH;
  const float *grid_c2w;
  long long grid_bs;
  const float *raw;   // fwd: [B, N, L*K2] in.   bwd: unused
  float *out;         // fwd: [B, L*K2, N] out.  bwd: incoming gradient (read only)
  float *draw;        // bwd: [B, N, L*K2] (+=)
};

template <bool kBwd>
__global__ void __launch_bounds__(kRotThreads) rotate_kernel(const RotateParams p) {
  extern __shared__ float tile[];  // [K2][33]
  const int lvl = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kRotPixels;
  const int C = p.L * p.K2;
  float *io = p.out + ((long long)b * C + (long long)lvl * p.K2) * p.N + n0;
  if constexpr (kBwd) {
    if (n0 + lane < p.N)
      for (int ch = warp; ch < p.K2; ch += kRotThreads / 32) tile[ch * 33 + lane] = io[(long long)ch * p.N + lane];
    __syncthreads();
  }
  const float *gx = p.grid_c2w + (long long)b * p.grid_bs;
#pragma unroll 1
  for (int qi = 0; qi < kRotPixels / (kRotThreads / 32); ++qi) {
    const int q = warp * (kRotPixels / (kRotThreads / 32)) + qi;
    const int n = n0 + q;
    if (n >= p.N) break;
    // module-local cyclic sampler of projection_prim_ortho.py:119-135 on the [.., h, w] map
    const float x = remainder_pos(__ldg(gx + n), p.axW.size);
    const float y = __ldg(gx + p.N + n);
    const Taps4 t = clamp_taps(make_taps(to_sample_coord(x, p.axW, p.div_mode), to_sample_coord(y, p.axH, p.div_mode)),
                               p.h, p.w);
    const long long base = (long long)b * p.N * C + (long long)lvl * p.K2;
    if constexpr (kBwd) {
      float *d = p.draw + base;
      for (int ch = lane; ch < p.K2; ch += 32) {
        const float g = tile[ch * 33 + q];
        if (t.nw != 0.f) atomicAdd(d + (long long)t.o_nw * C + ch, g * t.nw);
        if (t.ne != 0.f) atomicAdd(d + (long long)t.o_ne * C + ch, g * t.ne);
        if (t.sw != 0.f) atomicAdd(d + (long long)t.o_sw * C + ch, g * t.sw);
        if (t.se != 0.f) atomicAdd(d + (long long)t.o_se * C + ch, g * t.se);
      }
    } else {
      const float *s = p.raw + base;
      for (int ch = lane; ch < p.K2; ch += 32) {
        float acc = __fmul_rn(__ldg(s + (long long)t.o_nw * C + ch), t.nw);
        acc = __fmaf_rn(__ldg(s + (long long)t.o_ne * C + ch), t.ne, acc);
        acc = __fmaf_rn(__ldg(s + (long long)t.o_sw * C + ch), t.sw, acc);
        acc = __fmaf_rn(__ldg(s + (long long)t.o_se * C + ch), t.se, acc);
        tile[ch * 33 + q] = acc;
      }
    }
  }
  if constexpr (!kBwd) {
    __syncthreads();
    if (n0 + lane < p.N)
      for (int ch = warp; ch < p.K2; ch += kRotThreads / 32) io[(long long)ch * p.N + lane] = tile[ch * 33 + lane];
  }
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
  rp.axW = make_axis(a->w);
  rp.axH = make_axis(a->h);
  rp.grid_c2w = a->grid_c2w;
  rp.grid_bs = a->grid_batch_stride;
  rp.raw = a->scratch;
  rp.out = a->out_other;
  rp.draw = nullptr;
}

template <bool kBwd>
static int launch_lookup(const LookupParams &p, int radius, bool dual, cudaStream_t st, const char *who) {
  const int k = 2 * radius + 1, K2 = k * k;
  dim3 grid(ceil_div(p.N, kQueriesPerCta), p.L * (dual ? 2 : 1), p.B);
  const size_t smem = (size_t)K2 * 33 * sizeof(float);
  if (radius == 4)
    lookup_kernel<4, kBwd><<<grid, kLookupThreads, smem, st>>>(p);
  else
    lookup_kernel<0, kBwd><<<grid, kLookupThreads, smem, st>>>(p);
  return check_launch(who);
}

// img_rotate of a channels-last pre-rotation map, shared with the on-the-fly path (pf_onthefly.cu).
int rotate_forward(int batch, int h, int w, int num_levels, int radius, int div_mode, const float *grid_c2w,
                   long long grid_bs, const float *raw, float *out, cudaStream_t st) {
  const int k = 2 * radius + 1;
  RotateParams rp;
  rp.B = batch;
  rp.N = h * w;
  rp.h = h;
  rp.w = w;
  rp.L = num_levels;
  rp.K2 = k * k;
  rp.div_mode = div_mode;
  rp.axW = make_axis(w);
  rp.axH = make_axis(h);
  rp.grid_c2w = grid_c2w;
  rp.grid_bs = grid_bs;
  rp.raw = raw;
  rp.out = out;
  rp.draw = nullptr;
  dim3 grid(ceil_div(rp.N, kRotPixels), rp.L, rp.B);
  rotate_kernel<false><<<grid, kRotThreads, (size_t)rp.K2 * 33 * sizeof(float), st>>>(rp);
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
  PF_REQUIRE(!dual || a->out_other != nullptr, "pf_lookup_dual: out_other is required");
  for (int l = 0; l < p.L; ++l) {
    PF_REQUIRE(a->own[l] != nullptr, "pf_lookup_dual: own[%d] is null", l);
    PF_REQUIRE(!dual || a->other[l] != nullptr, "pf_lookup_dual: other[%d] is null", l);
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (int e = launch_lookup<false>(p, a->radius, dual, st, "pf_lookup_dual")) return e;
  if (dual) {
    // core/corr.py:137-138 — img_rotate of the [B, L*81, h, w] map with grid_c2w.
    return rotate_forward(a->batch, a->h, a->w, a->num_levels, a->radius, a->div_mode, a->grid_c2w,
                          a->grid_batch_stride, a->scratch, a->out_other, st);
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
  PF_REQUIRE(ba->grad_own != nullptr, "pf_lookup_dual_bwd: grad_own is required");
  for (int l = 0; l < p.L; ++l) {
    PF_REQUIRE(ba->dgrad_own[l] != nullptr, "pf_lookup_dual_bwd: dgrad_own[%d] is null", l);
    PF_REQUIRE(!dual || ba->dgrad_other[l] != nullptr, "pf_lookup_dual_bwd: dgrad_other[%d] is null", l);
    p.d_own[l] = ba->dgrad_own[l];
    p.d_other[l] = ba->dgrad_other[l];
  }
  p.dbg_own = p.dbg_other = nullptr;
  p.out_own = const_cast<float *>(ba->grad_own);
  cudaStream_t st = (cudaStream_t)stream;
  if (dual) {
    RotateParams rp;
    fill_rotate_params(a, rp);
    const size_t bytes = (size_t)a->batch * rp.L * rp.K2 * rp.N * sizeof(float);
    if (cudaMemsetAsync(a->scratch, 0, bytes, st) != cudaSuccess) return check_launch("pf_lookup_dual_bwd(memset)");
    rp.out = const_cast<float *>(ba->grad_other);
    rp.draw = a->scratch;
    dim3 grid(ceil_div(rp.N, kRotPixels), rp.L, rp.B);
    rotate_kernel<true><<<grid, kRotThreads, (size_t)rp.K2 * 33 * sizeof(float), st>>>(rp);
    if (int e = check_launch("pf_lookup_dual_bwd(rotate)")) return e;
  }
  return launch_lookup<true>(p, a->radius, dual, st, "pf_lookup_dual_bwd");
}
