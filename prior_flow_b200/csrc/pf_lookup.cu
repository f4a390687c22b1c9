// (b) Dual-cost pyramid lookup — DCCL.__call__ (PriOr-RAFT/core/corr.py:113-144) and
// CorrBlock.__call__ (core/corr.py:30-51): one gather launch serving both views + one rotate launch.
//
// lookup_kernel — grid = (ceil(N/32) query chunks, levels x branches, batch), 256 threads.
//   A CTA owns 32 consecutive query pixels of one level of one branch; each warp walks 4 queries.
//   * Window coordinates are separable: the 2r+1 x-coordinates and 2r+1 y-coordinates of a window
//     go through the (remainder, normalise, unnormalise, floor, clamp) chain once each, on lanes
//     0..2k-1, into a per-warp shared-memory table of {clamped offsets, validity-folded weights};
//     a tap is then two 16-byte table reads, four adds and four multiplies away from its loads.
//     Instruction issue — not memory — limited the earlier versions (ncu r01a: 76 % issue-slot
//     utilisation at 13 % DRAM; r01b: 25 instructions per load, half of them integer address math).
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

// One axis of one window, resolved once per query: clamped element offsets of the two taps and their
// weights with the zero-padding validity folded in (0 * finite == 0 reproduces ATen's skipped tap, and
// nw = w0x * w0y is the same rounded product as (ix_se - ix) * (iy_se - iy) when both taps are valid).
struct AxisEntry {
  int o0, o1;      // clamp(i0) * stride, clamp(i0 + 1) * stride
  float w0, w1;    // (i0+1 - s) if i0 in range else 0 ; (s - i0) if i0+1 in range else 0
};

template <int kDiv>
__device__ __forceinline__ float sample_coord(float p, const Axis ax) {   // to_sample_coord, div mode resolved at compile time
  const float t = __fmul_rn(2.f, p);
  float g = (kDiv == PF_DIV_ATEN_CUDA) ? __fmul_rn(t, ax.inv_m1) : __fdiv_rn(t, ax.size_m1);
  g = __fsub_rn(g, 1.f);
  float v = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), ax.size_m1);
  if (!(fabsf(v) <= 2147483648.f)) v = -100.f;   // non-finite or outside the int range (safe_downgrade_to_int_range)
  return v;
}

__device__ __forceinline__ AxisEntry make_axis_entry(float s, int size, int stride) {
  const float fl = floorf(s);
  const int i0 = (int)fl;
  AxisEntry e;
  e.w1 = ((unsigned)(i0 + 1) < (unsigned)size) ? __fsub_rn(s, fl) : 0.f;
  e.w0 = ((unsigned)i0 < (unsigned)size) ? __fsub_rn(__fadd_rn(fl, 1.f), s) : 0.f;
  e.o0 = min(max(i0, 0), size - 1) * stride;
  e.o1 = min(max(i0 + 1, 0), size - 1) * stride;
  return e;
}

template <int R, bool kBwd, int kDiv>
__global__ void __launch_bounds__(kLookupThreads, 5) lookup_kernel(const LookupParams p) {
  const int r = (R > 0) ? R : p.radius;
  const int k = 2 * r + 1;
  const int K2 = k * k;
  constexpr int kRounds = (R > 0) ? ((2 * R + 1) * (2 * R + 1) + 31) / 32 : 8;  // radius <= 7 -> <= 225 taps
  extern __shared__ float4 smem4[];
  // [8 warps][2 buffers][32] axis entries, then the [K2][33] transpose tile (own-view branch only)
  AxisEntry *tab = reinterpret_cast<AxisEntry *>(smem4) + (threadIdx.x >> 5) * 64;
  float *tile = reinterpret_cast<float *>(smem4 + (kLookupThreads / 32) * 64);
  float *dbg_tab = tile + K2 * 33 + (threadIdx.x >> 5) * 32;   // sample coordinates for the debug dump
  const int lvl = blockIdx.y % p.L;
  const int branch = blockIdx.y / p.L;
  const int b = blockIdx.z;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp-uniform for the compiler
  const int n0 = blockIdx.x * kQueriesPerCta;
  const int Hl = p.Hl[lvl], Wl = p.Wl[lvl];
  const Axis axW = p.axW[lvl], axH = p.axH[lvl];
  const float inv_scale = 1.0f / (float)(1 << lvl);  // `coords / 2**i` is exact either way
  const float *vol = branch ? p.other[lvl] : p.own[lvl];
  const float *gridx = opaque(p.grid_w2c + (long long)b * p.grid_bs);
  const float *gridy = opaque(gridx + p.N);
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
  const float off = (float)((is_x ? lane : lane - k) - r);
  const Axis ax1 = branch ? (is_x ? p.ax_gw : p.ax_gh) : (is_x ? axW : axH);  // first sampler's axis
  const int size1 = branch ? (is_x ? p.w : p.h) : (is_x ? Wl : Hl);
  const int stride1 = is_x ? 1 : (branch ? p.w : Wl);
  const bool wrap1 = is_x && (branch || p.cyclic);
  // tap -> (window column a, window row b, output channel), fixed per lane and round.
  // own view: lanes walk x fastest (coalesced plane rows); other view: lanes walk the output channel.
  int t_a[kRounds], t_b[kRounds];
#pragma unroll
  for (int it = 0; it < kRounds; ++it) {
    const int t = it * 32 + lane;
    const int tt = t < K2 ? t : 0;
    const int hi = tt / k, lo = tt - hi * k;
    t_a[it] = branch ? hi : lo;
    t_b[it] = branch ? lo : hi;
  }

#pragma unroll 1
  for (int qi = 0; qi < kQueriesPerWarp; ++qi) {
    const int q = warp * kQueriesPerWarp + qi;
    const int n = n0 + q;
    if (n >= p.N) break;
    AxisEntry *T = tab + (qi & 1) * 32;
    {
      // core/corr.py:123-126 then the sampler's coordinate chain, once per window row / column
      const float c = __fmul_rn(__ldg(p.coords + ((long long)b * 2 + (is_x ? 0 : 1)) * p.N + n), inv_scale);
      float pc = __fadd_rn(c, off);
      if (wrap1) pc = remainder_pos(pc, ax1.size);
      const float sc = sample_coord<kDiv>(pc, ax1);
      T[lane] = make_axis_entry(sc, size1, stride1);
      if (dbg != nullptr) dbg_tab[lane] = sc;
    }
    __syncwarp();
    const long long plane_off = ((long long)b * p.N + n) * (long long)(Hl * Wl);
    const float *plane = kBwd ? nullptr : opaque(vol + plane_off);
    float *dplane = kBwd ? opaque((branch ? p.d_other[lvl] : p.d_own[lvl]) + plane_off) : nullptr;
    float *rawq = opaque(p.raw + (((long long)b * p.N + n) * p.L + lvl) * K2);

#pragma unroll
    for (int it = 0; it < kRounds; ++it) {
      if (it * 32 >= K2) break;
      if (it * 32 + lane < K2) {
        const int aa = t_a[it], bb = t_b[it];
        const int ch = aa * k + bb;  // x-major channel order of the reference
        const AxisEntry ex = T[aa], ey = T[k + bb];
        int o_nw = ey.o0 + ex.o0, o_ne = ey.o0 + ex.o1, o_sw = ey.o1 + ex.o0, o_se = ey.o1 + ex.o1;
        float w_nw = __fmul_rn(ex.w0, ey.w0), w_ne = __fmul_rn(ex.w1, ey.w0);
        float w_sw = __fmul_rn(ex.w0, ey.w1), w_se = __fmul_rn(ex.w1, ey.w1);
        float ix = 0.f, iy = 0.f;
        if (branch) {
          // core/corr.py:132-136 — map through the level-0 rotation grid, then index the level-l volume
          float sx = __fmul_rn(__ldg(gridx + o_nw), w_nw), sy = __fmul_rn(__ldg(gridy + o_nw), w_nw);
          sx = __fmaf_rn(__ldg(gridx + o_ne), w_ne, sx);
          sy = __fmaf_rn(__ldg(gridy + o_ne), w_ne, sy);
          sx = __fmaf_rn(__ldg(gridx + o_sw), w_sw, sx);
          sy = __fmaf_rn(__ldg(gridy + o_sw), w_sw, sy);
          sx = __fmaf_rn(__ldg(gridx + o_se), w_se, sx);
          sy = __fmaf_rn(__ldg(gridy + o_se), w_se, sy);
          ix = sample_coord<kDiv>(remainder_pos(sx, axW.size), axW);
          iy = sample_coord<kDiv>(sy, axH);
          const AxisEntry fx = make_axis_entry(ix, Wl, 1), fy = make_axis_entry(iy, Hl, Wl);
          o_nw = fy.o0 + fx.o0, o_ne = fy.o0 + fx.o1, o_sw = fy.o1 + fx.o0, o_se = fy.o1 + fx.o1;
          w_nw = __fmul_rn(fx.w0, fy.w0), w_ne = __fmul_rn(fx.w1, fy.w0);
          w_sw = __fmul_rn(fx.w0, fy.w1), w_se = __fmul_rn(fx.w1, fy.w1);
        }
        if constexpr (kBwd) {
          const float g = branch ? __ldg(rawq + ch) : tile[ch * 33 + q];
          if (w_nw != 0.f) atomicAdd(dplane + o_nw, g * w_nw);
          if (w_ne != 0.f) atomicAdd(dplane + o_ne, g * w_ne);
          if (w_sw != 0.f) atomicAdd(dplane + o_sw, g * w_sw);
          if (w_se != 0.f) atomicAdd(dplane + o_se, g * w_se);
        } else {
          // a tap that falls entirely outside the plane (most of levels 2-3 of the orthogonal branch, whose
          // level-0-unit coordinates index a 16x32 / 8x16 plane) contributes an exact 0: skip its DRAM round trip
          float val = 0.f;
          if (!branch || fmaxf(fmaxf(w_nw, w_ne), fmaxf(w_sw, w_se)) > 0.f) {
            val = __fmul_rn(__ldg(plane + o_nw), w_nw);
            val = __fmaf_rn(__ldg(plane + o_ne), w_ne, val);
            val = __fmaf_rn(__ldg(plane + o_sw), w_sw, val);
            val = __fmaf_rn(__ldg(plane + o_se), w_se, val);
          }
          if (branch)
            rawq[ch] = val;
          else
            tile[ch * 33 + q] = val;
          if (dbg != nullptr) {
            if (!branch) {
              ix = dbg_tab[aa];
              iy = dbg_tab[k + bb];
            }
            float *d = dbg + ((((long long)b * p.N + n) * p.L + lvl) * K2 + ch) * 2;
            d[0] = ix;
            d[1] = iy;
          }
        }
      }
    }
    if (dbg != nullptr) __syncwarp();   // dbg_tab is single-buffered
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
  extern __shared__ float tile[];  // [K2][33] transpose tile, then one Taps4 per pixel
  Taps4 *taps = reinterpret_cast<Taps4 *>(tile + p.K2 * 33 + (p.K2 & 1));
  const int lvl = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kRotPixels;
  const int C = p.L * p.K2;
  float *io = p.out + ((long long)b * C + (long long)lvl * p.K2) * p.N + n0;
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
    if (n0 + lane < p.N)
      for (int ch = warp; ch < p.K2; ch += kRotThreads / 32) tile[ch * 33 + lane] = io[(long long)ch * p.N + lane];
  }
  __syncthreads();
  const long long base = (long long)b * p.N * C + (long long)lvl * p.K2;
  const float *src = kBwd ? nullptr : opaque(p.raw + base);
  float *dst = kBwd ? opaque(p.draw + base) : nullptr;
#pragma unroll 1
  for (int qi = 0; qi < kRotPixels / (kRotThreads / 32); ++qi) {
    const int q = warp * (kRotPixels / (kRotThreads / 32)) + qi;
    if (n0 + q >= p.N) break;
    const Taps4 t = taps[q];
    if constexpr (kBwd) {
      for (int ch = lane; ch < p.K2; ch += 32) {
        const float g = tile[ch * 33 + q];
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
  const size_t smem = (size_t)(kLookupThreads / 32) * 64 * sizeof(float4) + ((size_t)K2 * 33 + kLookupThreads) * sizeof(float);
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
    rotate_kernel<true><<<grid, kRotThreads, rotate_smem_bytes(rp.K2), st>>>(rp);
    if (int e = check_launch("pf_lookup_dual_bwd(rotate)")) return e;
  }
  return launch_lookup<true>(p, a->radius, dual, st, "pf_lookup_dual_bwd");
}
