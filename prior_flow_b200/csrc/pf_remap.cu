// (d) ERP <-> orthogonal-view geometry kernels: sample-grid generation, bilinear remap
// (img_rotate / cycle_bilinear_sampler / bilinear_sampler), fused flo_rotate, fused feature warp +
// group-wise correlation, and the small pooling helpers.  All HBM-bound streaming kernels: one
// thread per output pixel on the coordinate path (computed once, reused across channels), fully
// coalesced stores, gathers served by L1/L2 (rotations are locally smooth, so a warp's taps fall
// into a handful of sectors).
#include <string.h>

#include "pf_common.cuh"

namespace pf {

// ------------------------------------------------------------------------------------------------
// generate_samplegrid — core/utils/projection_prim_ortho.py:432-443.  fp32 op-for-op (SURVEY.md
// §A.4): np.pi is cast to fp32 at every tensor-scalar op; cosf/sinf/asinf/atan2f are libdevice's,
// the same functions ATen's CUDA kernels call.
struct Rot {
  float m[9];
};

__device__ __forceinline__ float scalar_div(float a, float s, float inv_s, int div_mode) {
  return div_mode == PF_DIV_ATEN_CUDA ? __fmul_rn(a, inv_s) : __fdiv_rn(a, s);
}

__device__ __forceinline__ float diverge_zero(float t) {  // projection_prim_ortho.py:69-74
  const float eps = 1e-6f;
  float near = fabsf(t) < eps ? 1.f : 0.f;
  float sgn = (t > 0.f) ? 1.f : ((t < 0.f) ? -1.f : 0.f);
  return __fadd_rn(t, __fmul_rn(__fmul_rn(sgn, near), eps));
}

__global__ void samplegrid_kernel(float *__restrict__ out, int B, int H, int W, Rot R, int div_mode) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= H * W) return;
  const int n = pix / W, m = pix - n * W;
  const float PI = 3.14159274101257324f;       // float(np.pi)
  const float TWO_PI = 6.28318548202514648f;   // float(2 * np.pi)
  const float Wf = (float)W, Hf = (float)H;
  // ERP.plane2spherical (:397-411): theta = ((m + .5)/W - .5) * 2 * pi ; phi = (.5 - (n + .5)/H) * pi
  const float u = scalar_div(__fadd_rn((float)m, 0.5f), Wf, 1.0f / Wf, div_mode);
  const float theta = __fmul_rn(__fmul_rn(__fsub_rn(u, 0.5f), 2.f), PI);
  const float v = scalar_div(__fadd_rn((float)n, 0.5f), Hf, 1.0f / Hf, div_mode);
  const float phi = __fmul_rn(__fsub_rn(0.5f, v), PI);
  // Spherical2Cartesian (:77-89)
  const float cphi = cosf(phi);
  const float x = __fmul_rn(cphi, cosf(theta));
  const float y = __fmul_rn(cphi, sinf(theta));
  const float z = sinf(phi);
  // rotate_cartesian (:247-261): R @ v as a k = 0..2 FMA chain
  const float xr = __fmaf_rn(R.m[2], z, __fmaf_rn(R.m[1], y, __fmul_rn(R.m[0], x)));
  const float yr = __fmaf_rn(R.m[5], z, __fmaf_rn(R.m[4], y, __fmul_rn(R.m[3], x)));
  const float zr = __fmaf_rn(R.m[8], z, __fmaf_rn(R.m[7], y, __fmul_rn(R.m[6], x)));
  // Cartesian2Spherical (:51-66)
  const float phi2 = asinf(zr);
  const float theta2 = atan2f(diverge_zero(yr), diverge_zero(xr));
  // ERP.spherical2plane (:413-429): m' = (theta/(2pi) + .5) * W - .5 ; n' = (.5 - phi/pi) * H - .5
  const float u2 = __fadd_rn(scalar_div(theta2, TWO_PI, 1.0f / TWO_PI, div_mode), 0.5f);
  const float m2 = __fsub_rn(__fmul_rn(u2, Wf), 0.5f);
  const float v2 = __fsub_rn(0.5f, scalar_div(phi2, PI, 1.0f / PI, div_mode));
  const float n2 = __fsub_rn(__fmul_rn(v2, Hf), 0.5f);
  const long long HW = (long long)H * W;
  for (int b = 0; b < B; ++b) {  // batch-invariant (generate_plane_grid repeats over B, :16-17)
    out[(2LL * b + 0) * HW + pix] = m2;
    out[(2LL * b + 1) * HW + pix] = n2;
  }
}

// ------------------------------------------------------------------------------------------------
// Bilinear remap: out[b,c,p] = grid_sample(src[b,c], coords[b,p]) with the wrappers' semantics.
constexpr int kRemapThreads = 256;
constexpr int kRemapChannelsPerThread = 8;

struct RemapParams {
  int B, C, H, W, P;  // P = Ho*Wo
  int cyclic, div_mode;
  Axis axW, axH;
  const float *src, *coords;
  long long cbs, cps, cxs;
  float *out;
};

__global__ void __launch_bounds__(kRemapThreads) remap_kernel(const RemapParams p) {
  const int pix = blockIdx.x * kRemapThreads + threadIdx.x;
  if (pix >= p.P) return;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * kRemapChannelsPerThread;
  const float *cp = p.coords + (long long)b * p.cbs + (long long)pix * p.cps;
  float x = __ldg(cp), y = __ldg(cp + p.cxs);
  if (p.cyclic) x = remainder_pos(x, p.axW.size);
  const Taps4 t = clamp_taps(make_taps(to_sample_coord(x, p.axW, p.div_mode), to_sample_coord(y, p.axH, p.div_mode)),
                             p.H, p.W);
  const long long plane = (long long)p.H * p.W;
  const float *src = opaque(p.src + ((long long)b * p.C + c0) * plane);
  float *out = opaque(p.out + ((long long)b * p.C + c0) * p.P + pix);
  const int cn = min(kRemapChannelsPerThread, p.C - c0);
#pragma unroll 8
  for (int c = 0; c < cn; ++c, src += plane, out += p.P) *out = blend4(src, t);
}

// ------------------------------------------------------------------------------------------------
// flo_rotate — projection_prim_ortho.py:531-546, fused.  cycle_grid_sample's blend is a chain of
// separate torch ops, so every mul/add is individually rounded (no FMA) — this kernel is
// bit-exact against the reference on either device.
struct GridTaps {  // my_cycle_sample.py:31-56
  float wa, wb, wc, wd;
  int x0, x1, y0, y1;
};
__device__ __forceinline__ int pymod(int a, int m) {
  int r = a % m;
  return r < 0 ? r + m : r;
}
__device__ __forceinline__ GridTaps make_grid_taps(float gx_raw, float gy, int H, int W) {
  GridTaps t;
  const float gx = remainder_pos(gx_raw, (float)W);
  const float fx = floorf(gx), fy = floorf(gy);
  const float xw = __fsub_rn(gx, fx), yw = __fsub_rn(gy, fy);
  const float xm = __fsub_rn(1.f, xw), ym = __fsub_rn(1.f, yw);
  t.wa = __fmul_rn(xm, ym);
  t.wb = __fmul_rn(xm, yw);
  t.wc = __fmul_rn(xw, ym);
  t.wd = __fmul_rn(xw, yw);
  const int xi = (int)fx, yi = (int)fy;
  t.x0 = pymod(xi, W);
  t.x1 = pymod(xi + 1, W);
  t.y0 = min(max(yi, 0), H - 1);
  t.y1 = min(max(yi + 1, 0), H - 1);
  return t;
}
__device__ __forceinline__ float blend4(const GridTaps &t, float Ia, float Ib, float Ic, float Id) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.wa, Ia), __fmul_rn(t.wb, Ib)), __fmul_rn(t.wc, Ic)),
                   __fmul_rn(t.wd, Id));
}
// adjust_sample_m (my_cycle_sample.py:91-93): Ia + ((I - Ia) + W/2) % W - W/2
__device__ __forceinline__ float recentre(float I, float Ia, float Wf, float half) {
  return __fsub_rn(__fadd_rn(Ia, remainder_pos(__fadd_rn(__fsub_rn(I, Ia), half), Wf)), half);
}

// flow in the orthogonal ("camera") frame at integer pixel (qx, qy): lines :537-542.
__device__ __forceinline__ void flow_c_at(const float *__restrict__ flow, const float *__restrict__ gw, int H, int W,
                                          int qx, int qy, float &fcx, float &fcy) {
  const int HW = H * W, q = qy * W + qx;
  const float Wf = (float)W, half = Wf * 0.5f;
  // flow2endpoint (:200-218)
  const float ex0 = __fadd_rn((float)qx, __ldg(flow + q));
  const float ey0 = __fadd_rn((float)qy, __ldg(flow + HW + q));
  const float ex = __fsub_rn(remainder_pos(__fadd_rn(ex0, 0.5f), Wf), 0.5f);
  const float ey = fminf(fmaxf(ey0, -0.5f), (float)H - 0.5f);
  // cycle_grid_sample(grid_W2C, end_W, is_grid=True)
  const GridTaps t = make_grid_taps(ex, ey, H, W);
  const int ia = t.y0 * W + t.x0, ib = t.y1 * W + t.x0, ic = t.y0 * W + t.x1, id = t.y1 * W + t.x1;
  const float Ma = __ldg(gw + ia);
  const float Mb = recentre(__ldg(gw + ib), Ma, Wf, half);
  const float Mc = recentre(__ldg(gw + ic), Ma, Wf, half);
  const float Md = recentre(__ldg(gw + id), Ma, Wf, half);
  const float end_m = blend4(t, Ma, Mb, Mc, Md);
  const float end_n = blend4(t, __ldg(gw + HW + ia), __ldg(gw + HW + ib), __ldg(gw + HW + ic), __ldg(gw + HW + id));
  // flow_C = end_C - start_C ; u_clip on the m component (:541-542, :234-244)
  const float fm = __fsub_rn(end_m, __ldg(gw + q));
  fcx = __fsub_rn(remainder_pos(__fadd_rn(fm, half), Wf), half);
  fcy = __fsub_rn(end_n, __ldg(gw + HW + q));
}

__global__ void flo_rotate_kernel(const float *__restrict__ flow, const float *__restrict__ grid_w2c,
                                  const float *__restrict__ grid_c2w, long long grid_bs, float *__restrict__ out, int H,
                                  int W) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (pix >= HW) return;
  const int b = blockIdx.y;
  const float *fl = flow + 2LL * b * HW;
  const float *gw = grid_w2c + b * grid_bs;
  const float *gc = grid_c2w + b * grid_bs;
  // cycle_grid_sample(flow_C, grid_C2W, is_grid=False) (:545)
  const GridTaps t = make_grid_taps(__ldg(gc + pix), __ldg(gc + HW + pix), H, W);
  float ax, ay, bx, by, cx, cy, dx, dy;
  flow_c_at(fl, gw, H, W, t.x0, t.y0, ax, ay);
  flow_c_at(fl, gw, H, W, t.x0, t.y1, bx, by);
  flow_c_at(fl, gw, H, W, t.x1, t.y0, cx, cy);
  flow_c_at(fl, gw, H, W, t.x1, t.y1, dx, dy);
  out[2LL * b * HW + pix] = blend4(t, ax, bx, cx, dx);
  out[2LL * b * HW + HW + pix] = blend4(t, ay, by, cy, dy);
}

// ------------------------------------------------------------------------------------------------
// Feature warp + group-wise correlation — core/prior_raft.py:173-174 with :77-83.
// CTA = 32 consecutive pixels x one channel group; its 4 warps split the group's channels; lanes are
// pixels, so fmap1 loads and the 4 warp taps of fmap2 are coalesced rows.  The bilinear taps are
// resolved once per pixel (clamped offsets + zero weights), the channel loop is 5 loads + 5 FMAs.
// Partial sums are combined through shared memory in a fixed order.
constexpr int kWgcWarps = 4;

__global__ void __launch_bounds__(kWgcWarps * 32) warp_groupcorr_kernel(
    const float *__restrict__ f1, const float *__restrict__ f2, const float *__restrict__ coords,
    float *__restrict__ out, int C, int H, int W, int G, Axis axW, Axis axH, int div_mode) {
  __shared__ float part[kWgcWarps][33];
  const int HW = H * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pix = blockIdx.x * 32 + lane;
  const int g = blockIdx.y, b = blockIdx.z;
  const int cpg = C / G, cpw = cpg / kWgcWarps;  // host guarantees divisibility
  float acc = 0.f;
  if (pix < HW) {
    const float x = remainder_pos(__ldg(coords + 2LL * b * HW + pix), axW.size);
    const float y = __ldg(coords + 2LL * b * HW + HW + pix);
    const Taps4 t = clamp_taps(make_taps(to_sample_coord(x, axW, div_mode), to_sample_coord(y, axH, div_mode)), H, W);
    const long long c0 = (long long)b * C + g * cpg + warp * cpw;
    const float *p1 = opaque(f1 + c0 * HW + pix);
    const float *p2 = opaque(f2 + c0 * HW);
#pragma unroll 8
    for (int c = 0; c < cpw; ++c, p1 += HW, p2 += HW) acc = __fmaf_rn(__ldg(p1), blend4(p2, t), acc);
  }
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && pix < HW) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kWgcWarps; ++j) s += part[j][lane];
    out[((long long)b * G + g) * HW + pix] = s / (float)cpg;
  }
}

// ------------------------------------------------------------------------------------------------
// 2x2 average pool (F.avg_pool2d(x, 2, stride=2)): ((a + b) + c + d) / 4 in window row-major order.
__global__ void avg_pool2x2_kernel(const float *__restrict__ in, float *__restrict__ out, long long planes, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;   // odd tails are dropped, like avg_pool2d
  const long long total = planes * Ho * Wo;
  const bool vec = (W % 2 == 0) && ((reinterpret_cast<uintptr_t>(in) & 7) == 0);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xo = (int)(i % Wo);
    const long long r = i / Wo;
    const int yo = (int)(r % Ho);
    const long long pl = r / Ho;
    const float *s = in + (pl * H + 2 * yo) * W + 2 * xo;
    float a, b, c, d;
    if (vec) {
      const float2 top = *reinterpret_cast<const float2 *>(s);
      const float2 bot = *reinterpret_cast<const float2 *>(s + W);
      a = top.x, b = top.y, c = bot.x, d = bot.y;
    } else {
      a = s[0], b = s[1], c = s[W], d = s[W + 1];
    }
    out[i] = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d), 0.25f);
  }
}

}  // namespace pf

extern "C" {

int pf_samplegrid(float *out, int batch, int H, int W, const float *R_host, int div_mode, void *stream) {
  using namespace pf;
  PF_REQUIRE(out && R_host && batch > 0 && H > 0 && W > 0, "pf_samplegrid: bad arguments");
  Rot R;
  memcpy(R.m, R_host, sizeof(R.m));
  samplegrid_kernel<<<ceil_div((long long)H * W, 256), 256, 0, (cudaStream_t)stream>>>(out, batch, H, W, R, div_mode);
  return check_launch("pf_samplegrid");
}

int pf_remap(const pf_remap_args *a, void *stream) {
  using namespace pf;
  PF_REQUIRE(a && a->src && a->coords && a->out, "pf_remap: null pointer");
  PF_REQUIRE(a->batch > 0 && a->channels > 0 && a->H > 0 && a->W > 0 && a->Ho > 0 && a->Wo > 0, "pf_remap: bad shape");
  RemapParams p;
  p.B = a->batch;
  p.C = a->channels;
  p.H = a->H;
  p.W = a->W;
  p.P = a->Ho * a->Wo;
  p.cyclic = a->cyclic;
  p.div_mode = a->div_mode;
  p.axW = make_axis(a->W);
  p.axH = make_axis(a->H);
  p.src = a->src;
  p.coords = a->coords;
  p.cbs = a->coord_batch_stride;
  p.cps = a->coord_pixel_stride;
  p.cxs = a->coord_xy_stride;
  p.out = a->out;
  dim3 grid(ceil_div(p.P, kRemapThreads), ceil_div(p.C, kRemapChannelsPerThread), p.B);
  PF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "pf_remap: too many channels/batches");
  remap_kernel<<<grid, kRemapThreads, 0, (cudaStream_t)stream>>>(p);
  return check_launch("pf_remap");
}

int pf_flo_rotate(const float *flow, const float *grid_w2c, const float *grid_c2w, long long grid_batch_stride, float *out,
                  int batch, int H, int W, void *stream) {
  using namespace pf;
  PF_REQUIRE(flow && grid_w2c && grid_c2w && out && batch > 0 && H > 0 && W > 0, "pf_flo_rotate: bad arguments");
  dim3 grid(ceil_div((long long)H * W, 128), batch);
  flo_rotate_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(flow, grid_w2c, grid_c2w, grid_batch_stride, out, H, W);
  return check_launch("pf_flo_rotate");
}

int pf_warp_groupcorr(const float *fmap1, const float *fmap2, const float *coords, float *out, int batch, int channels,
                      int h, int w, int groups, int div_mode, void *stream) {
  using namespace pf;
  PF_REQUIRE(fmap1 && fmap2 && coords && out, "pf_warp_groupcorr: null pointer");
  PF_REQUIRE(groups > 0 && channels % groups == 0 && (channels / groups) % kWgcWarps == 0,
             "pf_warp_groupcorr: need groups | channels and %d | channels/groups (got C=%d, G=%d)", kWgcWarps, channels,
             groups);
  dim3 grid(ceil_div((long long)h * w, 32), groups, batch);
  warp_groupcorr_kernel<<<grid, kWgcWarps * 32, 0, (cudaStream_t)stream>>>(fmap1, fmap2, coords, out, channels, h, w,
                                                                           groups, make_axis(w), make_axis(h), div_mode);
  return check_launch("pf_warp_groupcorr");
}

int pf_avg_pool2x2(const float *in, float *out, long long planes, int H, int W, void *stream) {
  using namespace pf;
  PF_REQUIRE(in && out && planes > 0 && H >= 2 && W >= 2, "pf_avg_pool2x2: bad arguments");
  const long long total = planes * (H / 2) * (W / 2);
  unsigned blocks = (unsigned)((total + 255) / 256 < 148LL * 32 ? (total + 255) / 256 : 148LL * 32);
  avg_pool2x2_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, out, planes, H, W);
  return check_launch("pf_avg_pool2x2");
}

int pf_fmap_pyramid(const float *fmap, float *const *levels, int num_levels, long long planes, int H, int W, void *stream) {
  using namespace pf;
  PF_REQUIRE(fmap && levels && num_levels >= 1 && num_levels <= PF_MAX_LEVELS, "pf_fmap_pyramid: bad arguments");
  const float *src = fmap;
  for (int l = 1; l < num_levels; ++l) {
    PF_REQUIRE(levels[l] != nullptr, "pf_fmap_pyramid: levels[%d] is null", l);
    if (int e = pf_avg_pool2x2(src, levels[l], planes, H >> (l - 1), W >> (l - 1), stream)) return e;
    src = levels[l];
  }
  return 0;
}

}  // extern "C"
