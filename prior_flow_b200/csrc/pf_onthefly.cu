// (c) On-the-fly (memory-efficient) dual lookup: the same outputs as pf_lookup_dual without ever
// materialising the O(N^2) volume — for high-resolution ERP (1024x2048: 4 GiB/view/level-0).
// Stands in for the unshipped `alt_cuda_corr.forward` behind AlternateCorrBlock
// (PriOr-RAFT/core/corr.py:64-91) and takes its operand convention: channels-last feature maps
// ([B,H,W,C], corr.py:82-83).  Level l correlates fmap1 with the avg-pooled fmap2 — equal to the
// lookup into the pooled volume because avg-pool is linear.
//
// One CTA = 8 consecutive queries of one level of one branch.  Per query: (A) all taps compute
// their bit-exact sample coordinates and the window's integer bounding box; (B) the correlation
// "plane" is evaluated only on that box — one warp per target pixel, lanes across channels,
// coalesced float4 reads of the K-major vectors, butterfly reduce — into shared memory; (C) taps
// blend from the box exactly like blend_zeros.  Boxes larger than the shared buffer (windows that
// straddle the seam or sit on a pole of the rotation map) fall back to four dot products per tap.
#include <climits>

#include "pf_common.cuh"

namespace pf {

constexpr int kOtfThreads = 128;
constexpr int kOtfQueries = 8;
constexpr int kOtfMaxBox = 576;   // 24 x 24 target pixels
constexpr int kOtfMaxTaps = 225;  // radius <= 7
constexpr int kOtfMaxVec = 4;     // C <= 512 (float4 per lane per 128 channels)

struct OtfParams {
  int B, N, h, w, C;
  int radius, L, cyclic, div_mode;
  const float *coords;
  const float *f1[2];                   // [B, N, C] own / other
  const float *f2[2][PF_MAX_LEVELS];    // [B, Hl*Wl, C]
  int Hl[PF_MAX_LEVELS], Wl[PF_MAX_LEVELS];
  Axis axW[PF_MAX_LEVELS], axH[PF_MAX_LEVELS], ax_gw, ax_gh;
  const float *grid_w2c;
  long long grid_bs;
  float scale;
  float *out_own, *out_raw;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float dot_pixel(const float4 *__restrict__ f2v, const float4 (&q)[kOtfMaxVec], int nvec,
                                           int lane) {
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < kOtfMaxVec; ++j) {
    if (j < nvec) {
      const float4 v = __ldg(f2v + j * 32 + lane);
      acc = fmaf(q[j].x, v.x, acc);
      acc = fmaf(q[j].y, v.y, acc);
      acc = fmaf(q[j].z, v.z, acc);
      acc = fmaf(q[j].w, v.w, acc);
    }
  }
  return warp_sum(acc);
}

__global__ void __launch_bounds__(kOtfThreads) onthefly_kernel(const OtfParams p) {
  __shared__ float s_ix[kOtfMaxTaps], s_iy[kOtfMaxTaps];
  __shared__ float s_dots[kOtfMaxBox];
  __shared__ float s_out[kOtfMaxTaps][kOtfQueries + 1];
  __shared__ int s_box[4];
  const int r = p.radius, k = 2 * r + 1, K2 = k * k;
  const int lvl = blockIdx.y % p.L, branch = blockIdx.y / p.L, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Hl = p.Hl[lvl], Wl = p.Wl[lvl];
  const Axis axW = p.axW[lvl], axH = p.axH[lvl];
  const float inv_scale = 1.0f / (float)(1 << lvl);
  const float *gridx = p.grid_w2c + (long long)b * p.grid_bs, *gridy = gridx + p.N;
  const float *f2 = p.f2[branch][lvl] + (long long)b * Hl * Wl * p.C;
  const int nvec = p.C / 128;
  const int n0 = blockIdx.x * kOtfQueries;

  for (int q = 0; q < kOtfQueries; ++q) {
    const int n = n0 + q;
    if (n >= p.N) break;
    if (threadIdx.x == 0) {
      s_box[0] = INT_MAX;
      s_box[1] = INT_MIN;
      s_box[2] = INT_MAX;
      s_box[3] = INT_MIN;
    }
    __syncthreads();
    const float cx = __fmul_rn(__ldg(p.coords + ((long long)b * 2 + 0) * p.N + n), inv_scale);
    const float cy = __fmul_rn(__ldg(p.coords + ((long long)b * 2 + 1) * p.N + n), inv_scale);
    // ---- (A) coordinates, identical to lookup_kernel
    for (int t = threadIdx.x; t < K2; t += kOtfThreads) {
      const int aa = t / k, bb = t - aa * k;  // t is the output channel a*k + b
      const float px = __fadd_rn(cx, (float)(aa - r)), py = __fadd_rn(cy, (float)(bb - r));
      float sx = px, sy = py;
      if (branch) {
        const float gx = to_sample_coord(remainder_pos(px, p.ax_gw.size), p.ax_gw, p.div_mode);
        const float gy = to_sample_coord(py, p.ax_gh, p.div_mode);
        const Taps tg = make_taps(gx, gy);
        sx = blend_zeros(gridx, p.h, p.w, tg);
        sy = blend_zeros(gridy, p.h, p.w, tg);
      }
      const float x = p.cyclic ? remainder_pos(sx, axW.size) : sx;
      const float ix = to_sample_coord(x, axW, p.div_mode), iy = to_sample_coord(sy, axH, p.div_mode);
      s_ix[t] = ix;
      s_iy[t] = iy;
      const int x0 = (int)floorf(ix), y0 = (int)floorf(iy);
      if (x0 + 1 >= 0 && x0 < Wl && y0 + 1 >= 0 && y0 < Hl) {  // tap touches the plane
        atomicMin(&s_box[0], max(x0, 0));
        atomicMax(&s_box[1], min(x0 + 1, Wl - 1));
        atomicMin(&s_box[2], max(y0, 0));
        atomicMax(&s_box[3], min(y0 + 1, Hl - 1));
      }
    }
    __syncthreads();
    const int x_lo = s_box[0], x_hi = s_box[1], y_lo = s_box[2], y_hi = s_box[3];
    const bool empty = x_lo > x_hi || y_lo > y_hi;
    const int bw = empty ? 0 : x_hi - x_lo + 1, bh = empty ? 0 : y_hi - y_lo + 1;
    const int area = bw * bh;
    // query vector: lane holds channels {128 j + 4 lane .. + 3}
    float4 qv[kOtfMaxVec];
    const float4 *f1v = reinterpret_cast<const float4 *>(p.f1[branch] + ((long long)b * p.N + n) * p.C);
#pragma unroll
    for (int j = 0; j < kOtfMaxVec; ++j) qv[j] = (j < nvec) ? __ldg(f1v + j * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);

    if (area <= kOtfMaxBox) {
      // ---- (B) the correlation plane restricted to the box
      for (int pix = warp; pix < area; pix += kOtfThreads / 32) {
        const int yy = pix / bw, xx = pix - yy * bw;
        const long long m = (long long)(y_lo + yy) * Wl + (x_lo + xx);
        const float d = dot_pixel(reinterpret_cast<const float4 *>(f2 + m * p.C), qv, nvec, lane);
        if (lane == 0) s_dots[pix] = d * p.scale;
      }
      __syncthreads();
      // ---- (C) blend (ATen order nw, ne, sw, se; out-of-plane taps contribute nothing)
      for (int t = threadIdx.x; t < K2; t += kOtfThreads) {
        const Taps tp = make_taps(s_ix[t], s_iy[t]);
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
      // ---- fallback: a warp per tap, four dot products each
      for (int t = warp; t < K2; t += kOtfThreads / 32) {
        const Taps tp = make_taps(s_ix[t], s_iy[t]);
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int xx = tp.x0 + (c & 1), yy = tp.y0 + (c >> 1);
          const bool in = (unsigned)xx < (unsigned)Wl && (unsigned)yy < (unsigned)Hl;  // warp-uniform
          v[c] = in ? dot_pixel(reinterpret_cast<const float4 *>(f2 + ((long long)yy * Wl + xx) * p.C), qv, nvec, lane) *
                          p.scale
                    : 0.f;
        }
        if (lane == 0) {
          float acc = __fmul_rn(v[0], tp.nw);
          acc = __fmaf_rn(v[1], tp.ne, acc);
          acc = __fmaf_rn(v[2], tp.sw, acc);
          acc = __fmaf_rn(v[3], tp.se, acc);
          s_out[t][q] = acc;
        }
      }
    }
    __syncthreads();
  }
  // ---- own view: [K2][8 queries] = 32-byte row segments of the [B, L*K2, N] output;
  //      other view: channels-last [B, N, L*K2] pre-rotation map (contiguous per query), see pf_lookup.cu
  if (branch == 0) {
    float *out = p.out_own + ((long long)b * p.L + lvl) * K2 * (long long)p.N + n0;
    for (int i = threadIdx.x; i < K2 * kOtfQueries; i += kOtfThreads) {
      const int ch = i / kOtfQueries, q = i - ch * kOtfQueries;
      if (n0 + q < p.N) out[(long long)ch * p.N + q] = s_out[ch][q];
    }
  } else {
    for (int i = threadIdx.x; i < K2 * kOtfQueries; i += kOtfThreads) {
      const int q = i / K2, ch = i - q * K2;
      if (n0 + q < p.N) p.out_raw[(((long long)b * p.N + n0 + q) * p.L + lvl) * K2 + ch] = s_out[ch][q];
    }
  }
}

}  // namespace pf

extern "C" int pf_lookup_onthefly(const pf_onthefly_args *a, void *stream) {
  using namespace pf;
  PF_REQUIRE(a != nullptr, "pf_lookup_onthefly: null args");
  PF_REQUIRE(a->batch > 0 && a->h > 0 && a->w > 0, "pf_lookup_onthefly: bad shape");
  PF_REQUIRE(a->channels % 128 == 0 && a->channels <= 128 * kOtfMaxVec,
             "pf_lookup_onthefly: channels must be a multiple of 128 and <= %d (got %d)", 128 * kOtfMaxVec, a->channels);
  PF_REQUIRE(a->num_levels >= 1 && a->num_levels <= PF_MAX_LEVELS, "pf_lookup_onthefly: num_levels must be 1..%d",
             PF_MAX_LEVELS);
  PF_REQUIRE(a->radius >= 0 && a->radius <= 7, "pf_lookup_onthefly: radius must be 0..7");
  PF_REQUIRE(a->coords && a->fmap1_own && a->out_own, "pf_lookup_onthefly: coords/fmap1_own/out_own are required");
  const bool dual = a->fmap1_other != nullptr;
  OtfParams p;
  p.B = a->batch;
  p.h = a->h;
  p.w = a->w;
  p.N = a->h * a->w;
  p.C = a->channels;
  p.radius = a->radius;
  p.L = a->num_levels;
  p.cyclic = a->cyclic;
  p.div_mode = a->div_mode;
  p.coords = a->coords;
  p.f1[0] = a->fmap1_own;
  p.f1[1] = a->fmap1_other;
  for (int l = 0; l < PF_MAX_LEVELS; ++l) {
    p.f2[0][l] = l < p.L ? a->fmap2_own[l] : nullptr;
    p.f2[1][l] = l < p.L ? a->fmap2_other[l] : nullptr;
    p.Hl[l] = a->h >> l;
    p.Wl[l] = a->w >> l;
    p.axH[l] = make_axis(p.Hl[l] > 0 ? p.Hl[l] : 1);
    p.axW[l] = make_axis(p.Wl[l] > 0 ? p.Wl[l] : 1);
    if (l < p.L) {
      PF_REQUIRE(p.Hl[l] >= 1 && p.Wl[l] >= 1, "pf_lookup_onthefly: level %d is empty", l);
      PF_REQUIRE(a->fmap2_own[l] != nullptr, "pf_lookup_onthefly: fmap2_own[%d] is null", l);
      PF_REQUIRE(!dual || a->fmap2_other[l] != nullptr, "pf_lookup_onthefly: fmap2_other[%d] is null", l);
    }
  }
  p.ax_gw = make_axis(a->w);
  p.ax_gh = make_axis(a->h);
  p.grid_w2c = a->grid_w2c;
  p.grid_bs = a->grid_batch_stride;
  p.scale = 1.0f / sqrtf((float)a->channels);
  p.out_own = a->out_own;
  p.out_raw = a->scratch;
  if (dual) {
    PF_REQUIRE(a->grid_w2c && a->grid_c2w && a->out_other && a->scratch,
               "pf_lookup_onthefly: dual lookup needs grid_w2c, grid_c2w, out_other and scratch");
    PF_REQUIRE(a->cyclic, "pf_lookup_onthefly: the dual (DCCL) lookup is defined for the cyclic sampler only");
  }
  dim3 grid(ceil_div(p.N, kOtfQueries), p.L * (dual ? 2 : 1), p.B);
  onthefly_kernel<<<grid, kOtfThreads, 0, (cudaStream_t)stream>>>(p);
  if (int e = check_launch("pf_lookup_onthefly")) return e;
  if (dual)
    return rotate_forward(a->batch, a->h, a->w, a->num_levels, a->radius, a->div_mode, a->grid_c2w,
                          a->grid_batch_stride, a->scratch, a->out_other, 0, 0, (cudaStream_t)stream, nullptr);
  return 0;
}
