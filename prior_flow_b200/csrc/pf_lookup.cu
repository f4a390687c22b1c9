// (b) Dual-cost pyramid lookup — DCCL.__call__ (PriOr-RAFT/core/corr.py:113-144) and
// CorrBlock.__call__ (core/corr.py:30-51) as one bandwidth-bound launch plus one remap launch.
//
// Work decomposition (one launch serves both views):
//   grid = (ceil(N/32) query chunks, levels x branches, batch), 256 threads.
//   A CTA owns 32 consecutive query pixels of one level of one branch.  Each warp walks 4 queries;
//   its lanes are the (2r+1)^2 window taps ordered y-major (x fastest across lanes) so that one
//   warp-wide load touches ~4 rows of <=10 contiguous floats of the query's private plane: the
//   plane rows are the only HBM traffic and every sector is fetched once (L1 serves the four-tap
//   overlap).  Results are transposed through shared memory so that the [B, L*81, h, w] output is
//   written as full 128-byte rows.
//   Branch 0 (own view):   sample pyr_own[l][n] at (c/2^l + d).
//   Branch 1 (other view): map (c/2^l + d) through the LEVEL-0 rotation grid, sample
//                          pyr_other[l][n] there (scale mixing is the reference's, SURVEY.md §0
//                          fact 9), write the pre-rotation map to `scratch`; pf_remap then applies
//                          img_rotate(., grid_c2w) (a cross-pixel gather, hence the second pass —
//                          the 10.6 MB intermediate stays in the 126 MB L2).
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
  float *out_own, *out_raw;            // forward: outputs.  backward: the incoming gradients (read only)
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
  extern __shared__ float tile[];  // [K2][33]
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
  float *io = (branch ? p.out_raw : p.out_own) + ((long long)b * p.L + lvl) * K2 * (long long)p.N + n0;
  if constexpr (kBwd) {  // stage the incoming gradient tile [K2][32 queries] with coalesced row reads
    if (n0 + lane < p.N)
      for (int ch = warp; ch < K2; ch += kLookupThreads / 32) tile[ch * 33 + lane] = io[(long long)ch * p.N + lane];
    __syncthreads();
  }

#pragma unroll 1
  for (int qi = 0; qi < kQueriesPerWarp; ++qi) {
    const int q = warp * kQueriesPerWarp + qi;
    const int n = n0 + q;
    if (n >= p.N) break;
    const float cx = __fmul_rn(__ldg(p.coords + ((long long)b * 2 + 0) * p.N + n), inv_scale);
    const float cy = __fmul_rn(__ldg(p.coords + ((long long)b * 2 + 1) * p.N + n), inv_scale);
    const long long plane_off = ((long long)b * p.N + n) * (long long)(Hl * Wl);
    const float *plane = kBwd ? nullptr : vol + plane_off;
    float *dplane = kBwd ? (branch ? p.d_other[lvl] : p.d_own[lvl]) + plane_off : nullptr;
    auto do_tap = [&](int t) {
      const int bb = t / k, aa = t - bb * k;  // lanes walk x fastest
      const float px = __fadd_rn(cx, (float)(aa - r));
      const float py = __fadd_rn(cy, (float)(bb - r));
      float sx, sy;
      if (branch == 0) {
        sx = px;
        sy = py;
      } else {
        // core/corr.py:132-133 — cycle_bilinear_sampler(sample_grid_W2C_8x, coords_lvl)
        const float gx = to_sample_coord(remainder_pos(px, p.ax_gw.size), p.ax_gw, p.div_mode);
        const float gy = to_sample_coord(py, p.ax_gh, p.div_mode);
        const Taps tg = make_taps(gx, gy);
        sx = blend_zeros(gridx, p.h, p.w, tg);
        sy = blend_zeros(gridy, p.h, p.w, tg);
      }
      const float x = p.cyclic ? remainder_pos(sx, axW.size) : sx;
      const float ix = to_sample_coord(x, axW, p.div_mode);
      const float iy = to_sample_coord(sy, axH, p.div_mode);
      const int ch = aa * k + bb;  // x-major channel order of the reference
      if constexpr (kBwd) {
        scatter_zeros(dplane, Hl, Wl, make_taps(ix, iy), tile[ch * 33 + q]);
        return;
      } else {
        tile[ch * 33 + q] = blend_zeros(plane, Hl, Wl, make_taps(ix, iy));
      }
      if (dbg != nullptr) {
        float *d = dbg + ((((long long)b * p.N + n) * p.L + lvl) * K2 + ch) * 2;
        d[0] = ix;
        d[1] = iy;
      }
    };
    if constexpr (R > 0) {
#pragma unroll
      for (int it = 0; it < (K2 + 31) / 32; ++it) {
        const int t = it * 32 + lane;
        if (t < K2) do_tap(t);
      }
    } else {
      for (int t = lane; t < K2; t += 32) do_tap(t);
    }
  }
  if constexpr (!kBwd) {
    __syncthreads();
    if (n0 + lane < p.N)
      for (int ch = warp; ch < K2; ch += kLookupThreads / 32) io[(long long)ch * p.N + lane] = tile[ch * 33 + lane];
  }
}

}  // namespace pf

namespace pf {

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
  p.out_raw = a->scratch;
  p.dbg_own = a->dbg_own_xy;
  p.dbg_other = a->dbg_other_xy;
  if (dual) {
    PF_REQUIRE(a->grid_w2c && a->grid_c2w && a->scratch, "%s: dual lookup needs grid_w2c, grid_c2w and scratch", who);
    PF_REQUIRE(a->cyclic, "%s: the dual (DCCL) lookup is defined for the cyclic sampler only", who);
  }
  return 0;
}

static void fill_rotate_args(const pf_lookup_args *a, pf_remap_args &ra) {
  const int k = 2 * a->radius + 1;
  ra.batch = a->batch;
  ra.channels = a->num_levels * k * k;
  ra.H = ra.Ho = a->h;
  ra.W = ra.Wo = a->w;
  ra.cyclic = 1;
  ra.div_mode = a->div_mode;
  ra.src = a->scratch;
  ra.coords = a->grid_c2w;
  ra.coord_batch_stride = a->grid_batch_stride;
  ra.coord_pixel_stride = 1;
  ra.coord_xy_stride = (long long)a->h * a->w;
  ra.out = a->out_other;
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
  const int k = 2 * a->radius + 1, K2 = k * k;
  dim3 grid(ceil_div(p.N, kQueriesPerCta), p.L * (dual ? 2 : 1), p.B);
  size_t smem = (size_t)K2 * 33 * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->radius == 4)
    lookup_kernel<4, false><<<grid, kLookupThreads, smem, st>>>(p);
  else
    lookup_kernel<0, false><<<grid, kLookupThreads, smem, st>>>(p);
  if (int e = check_launch("pf_lookup_dual")) return e;
  if (dual) {
    // core/corr.py:137-138 — img_rotate of the [B, L*81, h, w] map with grid_c2w.
    pf_remap_args ra;
    fill_rotate_args(a, ra);
    return pf_remap(&ra, stream);
  }
  return 0;
}

// (e) d(lookup)/d(pyramids).  The orthogonal branch first runs the adjoint of img_rotate
// (scatter of grad_other into the zeroed scratch map), then both branches scatter through the
// forward's coordinates.  A query's plane is private to it, so the atomics only ever collide
// between taps of one window.
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
  const int k = 2 * a->radius + 1, K2 = k * k;
  if (dual) {
    const size_t bytes = (size_t)a->batch * p.L * K2 * p.N * sizeof(float);
    if (cudaMemsetAsync(a->scratch, 0, bytes, st) != cudaSuccess) return check_launch("pf_lookup_dual_bwd(memset)");
    pf_remap_args ra;
    fill_rotate_args(a, ra);
    ra.out = nullptr;
    if (int e = pf_remap_bwd(&ra, ba->grad_other, a->scratch, stream)) return e;
  }
  dim3 grid(ceil_div(p.N, kQueriesPerCta), p.L * (dual ? 2 : 1), p.B);
  size_t smem = (size_t)K2 * 33 * sizeof(float);
  if (a->radius == 4)
    lookup_kernel<4, true><<<grid, kLookupThreads, smem, st>>>(p);
  else
    lookup_kernel<0, true><<<grid, kLookupThreads, smem, st>>>(p);
  return check_launch("pf_lookup_dual_bwd");
}
