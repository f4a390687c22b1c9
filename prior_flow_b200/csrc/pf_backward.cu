// (e) Backward kernels for training (train_flow.py step).  Coordinates and sample
// grids never need gradients (coords are detached every iteration, PriOr-RAFT/core/prior_raft.py:171,176),
// so every adjoint here is a scatter of the output gradient through the forward's bilinear taps.
// The lookup adjoint lives in pf_lookup.cu (same kernel, kBwd = true); the volume GEMM adjoints
// dF1 = dV F2^T / sqrt(C), dF2 = dV^T F1 / sqrt(C) are plain library GEMMs issued by the host layer.
#include "pf_common.cuh"

namespace pf {

// ---- adjoint of remap_kernel w.r.t. src (img_rotate / cycle_bilinear_sampler input gradient)
struct RemapBwdParams {
  int B, C, H, W, P;
  int cyclic, div_mode;
  Axis axW, axH;
  const float *coords, *dout;
  long long cbs, cps, cxs;
  float *dsrc;
};

constexpr int kRemapBwdChannels = 8;

__global__ void __launch_bounds__(256) remap_bwd_kernel(const RemapBwdParams p) {
  const int pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= p.P) return;
  const int b = blockIdx.z, c0 = blockIdx.y * kRemapBwdChannels;
  const float *cp = p.coords + (long long)b * p.cbs + (long long)pix * p.cps;
  float x = __ldg(cp), y = __ldg(cp + p.cxs);
  if (p.cyclic) x = remainder_pos(x, p.axW.size);
  const Taps t = make_taps(to_sample_coord(x, p.axW, p.div_mode), to_sample_coord(y, p.axH, p.div_mode));
  const bool xin0 = (unsigned)t.x0 < (unsigned)p.W, xin1 = (unsigned)(t.x0 + 1) < (unsigned)p.W;
  const bool yin0 = (unsigned)t.y0 < (unsigned)p.H, yin1 = (unsigned)(t.y0 + 1) < (unsigned)p.H;
  const long long plane = (long long)p.H * p.W;
  const long long o00 = (long long)t.y0 * p.W + t.x0;
  const int cn = min(kRemapBwdChannels, p.C - c0);
  for (int c = 0; c < cn; ++c) {
    const float g = __ldg(p.dout + ((long long)b * p.C + c0 + c) * p.P + pix);
    float *d = p.dsrc + ((long long)b * p.C + c0 + c) * plane + o00;
    if (yin0 && xin0) atomicAdd(d, g * t.nw);
    if (yin0 && xin1) atomicAdd(d + 1, g * t.ne);
    if (yin1 && xin0) atomicAdd(d + p.W, g * t.sw);
    if (yin1 && xin1) atomicAdd(d + p.W + 1, g * t.se);
  }
}

// ---- adjoint of the avg-pool pyramid, folded into level 0 in place
__global__ void pyramid_fold_kernel(float *__restrict__ g0, const float *__restrict__ g1, const float *__restrict__ g2,
                                    const float *__restrict__ g3, long long planes, int H, int W) {
  const long long total = planes * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long r = i / W;
    const int y = (int)(r % H);
    const long long pl = r / H;
    float v = g0[i];
    if (g1 && (y >> 1) < (H >> 1) && (x >> 1) < (W >> 1)) v += 0.25f * g1[(pl * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)];
    if (g2 && (y >> 2) < (H >> 2) && (x >> 2) < (W >> 2))
      v += 0.0625f * g2[(pl * (H >> 2) + (y >> 2)) * (W >> 2) + (x >> 2)];
    if (g3 && (y >> 3) < (H >> 3) && (x >> 3) < (W >> 3))
      v += 0.015625f * g3[(pl * (H >> 3) + (y >> 3)) * (W >> 3) + (x >> 3)];
    g0[i] = v;
  }
}

// The same for W % 4 == 0: a warp per plane row, float4 per lane (the fold is a pure stream: 340 MB read, 256 MB written per view)
__global__ void __launch_bounds__(256) pyramid_fold_rows_kernel(float *__restrict__ g0, const float *__restrict__ g1,
                                                                const float *__restrict__ g2, const float *__restrict__ g3,
                                                                long long planes, int H, int W) {
  const int lane = threadIdx.x & 31;
  const long long rows = planes * H, warps = (long long)gridDim.x * (blockDim.x >> 5);
  const int H1 = H >> 1, W1 = W >> 1, H2 = H >> 2, W2 = W >> 2, H3 = H >> 3, W3 = W >> 3;
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    const long long pl = r / H;
    const int y = (int)(r - pl * H);
    float4 *row = reinterpret_cast<float4 *>(g0 + r * W);
    const float *r1 = (g1 && (y >> 1) < H1) ? g1 + (pl * H1 + (y >> 1)) * W1 : nullptr;
    const float *r2 = (g2 && (y >> 2) < H2) ? g2 + (pl * H2 + (y >> 2)) * W2 : nullptr;
    const float *r3 = (g3 && (y >> 3) < H3) ? g3 + (pl * H3 + (y >> 3)) * W3 : nullptr;
    for (int x = lane * 4; x < W; x += 128) {
      float4 v = row[x >> 2];
      float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int xx = x + k;
        if (r1 && (xx >> 1) < W1) a[k] += 0.25f * __ldg(r1 + (xx >> 1));
        if (r2 && (xx >> 2) < W2) a[k] += 0.0625f * __ldg(r2 + (xx >> 2));
        if (r3 && (xx >> 3) < W3) a[k] += 0.015625f * __ldg(r3 + (xx >> 3));
      }
      row[x >> 2] = make_float4(a[0], a[1], a[2], a[3]);
    }
  }
}

// ---- adjoint of warp_groupcorr_kernel: dfmap1 written, dfmap2 accumulated with atomics
__global__ void __launch_bounds__(256) warp_groupcorr_bwd_kernel(const float *__restrict__ f1, const float *__restrict__ f2,
                                                                 const float *__restrict__ coords,
                                                                 const float *__restrict__ dout, float *__restrict__ df1,
                                                                 float *__restrict__ df2, int C, int H, int W, int G,
                                                                 Axis axW, Axis axH, int div_mode) {
  const int HW = H * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pix = blockIdx.x * 32 + lane;
  if (pix >= HW) return;
  const int b = blockIdx.y;
  const int cpw = C / 8, cpg = C / G;
  const float x = remainder_pos(__ldg(coords + 2LL * b * HW + pix), axW.size);
  const float y = __ldg(coords + 2LL * b * HW + HW + pix);
  const Taps t = make_taps(to_sample_coord(x, axW, div_mode), to_sample_coord(y, axH, div_mode));
  const bool xin0 = (unsigned)t.x0 < (unsigned)W, xin1 = (unsigned)(t.x0 + 1) < (unsigned)W;
  const bool yin0 = (unsigned)t.y0 < (unsigned)H, yin1 = (unsigned)(t.y0 + 1) < (unsigned)H;
  const long long o00 = (long long)t.y0 * W + t.x0;
  const float inv = 1.0f / (float)cpg;
  for (int c = warp * cpw; c < (warp + 1) * cpw; ++c) {
    const float g = __ldg(dout + ((long long)b * G + c / cpg) * HW + pix) * inv;
    const long long cb = ((long long)b * C + c) * HW;
    df1[cb + pix] = g * blend_zeros(f2 + cb, H, W, t);
    const float gf = g * __ldg(f1 + cb + pix);
    float *d = df2 + cb + o00;
    if (yin0 && xin0) atomicAdd(d, gf * t.nw);
    if (yin0 && xin1) atomicAdd(d + 1, gf * t.ne);
    if (yin1 && xin0) atomicAdd(d + W, gf * t.sw);
    if (yin1 && xin1) atomicAdd(d + W + 1, gf * t.se);
  }
}

}  // namespace pf

extern "C" {

int pf_remap_bwd(const pf_remap_args *a, const float *dout, float *dsrc, void *stream) {
  using namespace pf;
  PF_REQUIRE(a && a->coords && dout && dsrc, "pf_remap_bwd: null pointer");
  PF_REQUIRE(a->batch > 0 && a->channels > 0 && a->H > 0 && a->W > 0 && a->Ho > 0 && a->Wo > 0, "pf_remap_bwd: bad shape");
  RemapBwdParams p;
  p.B = a->batch;
  p.C = a->channels;
  p.H = a->H;
  p.W = a->W;
  p.P = a->Ho * a->Wo;
  p.cyclic = a->cyclic;
  p.div_mode = a->div_mode;
  p.axW = make_axis(a->W);
  p.axH = make_axis(a->H);
  p.coords = a->coords;
  p.dout = dout;
  p.cbs = a->coord_batch_stride;
  p.cps = a->coord_pixel_stride;
  p.cxs = a->coord_xy_stride;
  p.dsrc = dsrc;
  dim3 grid(ceil_div(p.P, 256), ceil_div(p.C, kRemapBwdChannels), p.B);
  PF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "pf_remap_bwd: too many channels/batches");
  remap_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("pf_remap_bwd");
}

int pf_pyramid_fold_bwd(float *const *glevel, int num_levels, long long planes, int H, int W, void *stream) {
  using namespace pf;
  PF_REQUIRE(glevel && num_levels >= 1 && num_levels <= PF_MAX_LEVELS && glevel[0], "pf_pyramid_fold_bwd: bad arguments");
  if (num_levels == 1) return 0;
  const float *g1 = glevel[1], *g2 = num_levels > 2 ? glevel[2] : nullptr, *g3 = num_levels > 3 ? glevel[3] : nullptr;
  const long long total = planes * H * W;
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148LL * 32 ? (total + 255) / 256 : 148LL * 32);
  if (W % 4 == 0 && (reinterpret_cast<uintptr_t>(glevel[0]) & 15) == 0) {
    const long long rows = planes * H;
    const unsigned rb = (unsigned)((rows + 7) / 8 < 148LL * 16 ? (rows + 7) / 8 : 148LL * 16);
    pyramid_fold_rows_kernel<<<rb, 256, 0, (cudaStream_t)stream>>>(glevel[0], g1, g2, g3, planes, H, W);
  } else {
    pyramid_fold_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(glevel[0], g1, g2, g3, planes, H, W);
  }
  return check_launch("pf_pyramid_fold_bwd");
}

int pf_warp_groupcorr_bwd(const float *fmap1, const float *fmap2, const float *coords, const float *dout, float *dfmap1,
                          float *dfmap2, int batch, int channels, int h, int w, int groups, int div_mode, void *stream) {
  using namespace pf;
  PF_REQUIRE(fmap1 && fmap2 && coords && dout && dfmap1 && dfmap2, "pf_warp_groupcorr_bwd: null pointer");
  PF_REQUIRE(groups > 0 && channels % 8 == 0 && channels % groups == 0,
             "pf_warp_groupcorr_bwd: need 8 | channels and groups | channels (got C=%d, G=%d)", channels, groups);
  dim3 grid(ceil_div((long long)h * w, 32), batch);
  warp_groupcorr_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(fmap1, fmap2, coords, dout, dfmap1, dfmap2, channels, h,
                                                                    w, groups, make_axis(w), make_axis(h), div_mode);
  return check_launch("pf_warp_groupcorr_bwd");
}

}  // extern "C"
