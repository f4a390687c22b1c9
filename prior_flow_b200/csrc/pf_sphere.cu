// the callers either side of the hot path that reuse its geometry.
//   pf_convex_upsample   PriOr_RAFT.upsample_flow (PriOr-RAFT/core/prior_raft.py:58-67): softmax over the 9 mask logits +
//                        convex combination of the 3x3 neighbourhood of 8*flow, [B,2,h,w] -> [B,2,8h,8w], one launch instead
//                        of softmax / unfold / mul / sum / permute over a 576-channel tensor
//   pf_uniform_loss_*    one term of uniform_loss (train_flow.py:55-79): sum(valid * cos-latitude weight * |pred - gt|_1)
//                        and its gradient w.r.t. the prediction (24 terms per training step)
//   pf_great_circle      calculate_great_circle_distance, Haversine (core/utils/spherical.py:20-53, the SEPE metric of
//                        evaluate.py:354): ERP endpoints of predicted and true flow -> angle between them on the sphere
// All bandwidth-bound, fp32.  Sums of floats are re-associated (stated tolerance 1e-5 relative); nothing here feeds coordinates.
#include "pf_common.cuh"

namespace pf {

// Thread = (coarse pixel, sub-pixel d = di*8 + dj), 4 coarse pixels per CTA.  Two mappings, by the mask's memory format:
//   channels-last mask: 64 consecutive threads = the 64 sub-pixels of one pixel (their logits of one k are contiguous)
//   NCHW mask:          warp = di, lane = 8 * pixel + dj: a warp reads 4 consecutive pixels of 8 channel planes (16-byte pieces)
//                       and writes 32 consecutive floats of one output row
template <bool kCL>
__device__ __forceinline__ void upsample_thread(int &pl, int &d) {
  if (kCL) {
    pl = threadIdx.x >> 6, d = threadIdx.x & 63;
  } else {
    const int lane = threadIdx.x & 31;
    pl = lane >> 3, d = (threadIdx.x >> 5) * 8 + (lane & 7);
  }
}

template <bool kCL>
__global__ void __launch_bounds__(256) convex_upsample_kernel(const float *__restrict__ flow, const float *__restrict__ mask,
                                                              float *__restrict__ out, int h, int w, long long m_bs,
                                                              long long m_cs, long long m_ps) {
  const int b = blockIdx.y;
  int pl, d;
  upsample_thread<kCL>(pl, d);
  const int pix = blockIdx.x * 4 + pl;                       // 4 coarse pixels per CTA
  if (pix >= h * w) return;
  const int di = d >> 3, dj = d & 7;
  const int i = pix / w, j = pix - i * w;
  const float *mp = mask + (long long)b * m_bs + (long long)pix * m_ps + (long long)d * m_cs;
  float lg[9], mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    lg[k] = __ldg(mp + (long long)(k * 64) * m_cs);
    mx = fmaxf(mx, lg[k]);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    lg[k] = expf(lg[k] - mx);
    s += lg[k];
  }
  const float inv = 1.0f / s;
  const float *f = flow + (long long)b * 2 * h * w;
  float u = 0.f, v = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int y = i + k / 3 - 1, x = j + k % 3 - 1;          // F.unfold(8 * flow, [3,3], padding=1): zero padding
    if ((unsigned)y < (unsigned)h && (unsigned)x < (unsigned)w) {
      const float wk = lg[k] * inv;
      u = fmaf(wk, 8.f * __ldg(f + y * w + x), u);
      v = fmaf(wk, 8.f * __ldg(f + h * w + y * w + x), v);
    }
  }
  const long long HW = 64LL * h * w;
  float *o = out + (long long)b * 2 * HW + (long long)(8 * i + di) * (8 * w) + 8 * j + dj;
  o[0] = u;
  o[HW] = v;
}

// Adjoint of convex_upsample_kernel, same thread mappings (the softmax is recomputed, nothing is saved):
//   dmask[k*64+d] = w_k (dw_k - sum_m w_m dw_m),  dw_k = 8 (g_u flow_u[nbr_k] + g_v flow_v[nbr_k])      (0 outside the image)
//   dflow_c[nbr_k] += 8 sum_d w_k g_c          — reduced over the 64 sub-pixels of the coarse pixel first (warp shuffles, then
//   shared-memory atomics across the warps that share a pixel), then 18 global atomics per coarse pixel
template <bool kCL>
__global__ void __launch_bounds__(256) convex_upsample_bwd_kernel(const float *__restrict__ flow, const float *__restrict__ mask,
                                                                  const float *__restrict__ gout, float *__restrict__ dflow,
                                                                  float *__restrict__ dmask, int h, int w, long long m_bs,
                                                                  long long m_cs, long long m_ps) {
  __shared__ float s_df[4][18];
  const int b = blockIdx.y;
  int pl, d;
  upsample_thread<kCL>(pl, d);
  if (threadIdx.x < 72) (&s_df[0][0])[threadIdx.x] = 0.f;
  __syncthreads();
  const int pix_raw = blockIdx.x * 4 + pl;
  const bool live = pix_raw < h * w;
  const int pix = live ? pix_raw : h * w - 1;      // lanes past the end compute on the last pixel and contribute nothing (the warp
                                                   // shuffles below need every lane)
  const int di = d >> 3, dj = d & 7, lane = threadIdx.x & 31;
  const int i = pix / w, j = pix - i * w;
  float *df = dflow + (long long)b * 2 * h * w;
  const long long moff = (long long)b * m_bs + (long long)pix * m_ps + (long long)d * m_cs;
  float wk[9], mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    wk[k] = __ldg(mask + moff + (long long)(k * 64) * m_cs);
    mx = fmaxf(mx, wk[k]);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    wk[k] = expf(wk[k] - mx);
    s += wk[k];
  }
  const float inv = 1.0f / s;
  const long long HW = 64LL * h * w;
  const float *g = gout + (long long)b * 2 * HW + (long long)(8 * i + di) * (8 * w) + 8 * j + dj;
  const float gu = 8.f * __ldg(g), gv = 8.f * __ldg(g + HW);
  const float *f = flow + (long long)b * 2 * h * w;
  float dw[9], dot = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int y = i + k / 3 - 1, x = j + k % 3 - 1;
    const bool in = live && (unsigned)y < (unsigned)h && (unsigned)x < (unsigned)w;
    wk[k] *= inv;
    dw[k] = in ? fmaf(gu, __ldg(f + y * w + x), gv * __ldg(f + h * w + y * w + x)) : 0.f;
    dot = fmaf(wk[k], dw[k], dot);
    // flow gradient: reduce w_k g over the lanes of this warp that belong to the pixel (all 32, or the 8 of a dj group)
    float cu = in ? wk[k] * gu : 0.f, cv = in ? wk[k] * gv : 0.f;
#pragma unroll
    for (int o = kCL ? 16 : 4; o; o >>= 1) {
      cu += __shfl_xor_sync(0xffffffffu, cu, o);
      cv += __shfl_xor_sync(0xffffffffu, cv, o);
    }
    if ((kCL ? lane == 0 : (lane & 7) == 0) && in) {
      atomicAdd(&s_df[pl][2 * k], cu);
      atomicAdd(&s_df[pl][2 * k + 1], cv);
    }
  }
  if (live) {
#pragma unroll
    for (int k = 0; k < 9; ++k) dmask[moff + (long long)(k * 64) * m_cs] = wk[k] * (dw[k] - dot);
  }
  __syncthreads();
  if (threadIdx.x < 72) {
    const int q = threadIdx.x / 18, r = threadIdx.x - q * 18, k = r >> 1, c = r & 1;
    const int px = blockIdx.x * 4 + q;
    if (px < h * w) {
      const int y = px / w + k / 3 - 1, x = px % w + k % 3 - 1;
      if ((unsigned)y < (unsigned)h && (unsigned)x < (unsigned)w) atomicAdd(df + c * h * w + y * w + x, s_df[q][r]);
    }
  }
}

// one loss term: acc += term_weight * sum(ok * lat[y] * (|pu - gu| + |pv - gv|))
__global__ void __launch_bounds__(256) uniform_loss_fwd_kernel(const float *__restrict__ pred, const float *__restrict__ gt,
                                                               const float *__restrict__ ok, const float *__restrict__ lat, float *acc,
                                                               float term_weight, int B, int H, int W) {
  const long long HW = (long long)H * W, total = (long long)B * HW;
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    const float m = __ldg(ok + i);
    if (m != 0.f) {
      const long long o = b * 2 * HW + p;
      s += m * __ldg(lat + p / W) * (fabsf(pred[o] - __ldg(gt + o)) + fabsf(pred[o + HW] - __ldg(gt + o + HW)));
    }
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += part[k];
    atomicAdd(acc, term_weight * t);
  }
}

// d(term)/d(pred) = upstream * term_weight * ok * lat[y] * sign(pred - gt)
__global__ void __launch_bounds__(256) uniform_loss_bwd_kernel(const float *__restrict__ pred, const float *__restrict__ gt,
                                                               const float *__restrict__ ok, const float *__restrict__ lat,
                                                               const float *__restrict__ upstream, float term_weight, float *__restrict__ dpred,
                                                               int B, int H, int W) {
  const long long HW = (long long)H * W, total = (long long)B * HW;
  const float g = __ldg(upstream) * term_weight;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW, o = b * 2 * HW + p;
    const float c = g * __ldg(ok + i) * __ldg(lat + p / W);
    const float du = pred[o] - __ldg(gt + o), dv = pred[o + HW] - __ldg(gt + o + HW);
    dpred[o] = du > 0.f ? c : (du < 0.f ? -c : 0.f);          // torch's abs backward: sign(x), sign(0) = 0
    dpred[o + HW] = dv > 0.f ? c : (dv < 0.f ? -c : 0.f);
  }
}

__device__ __forceinline__ void erp_endpoint_angles(int m, int n, float fu, float fv, int H, int W, float &theta, float &phi) {
  // flow2endpoint (projection_prim_ortho.py:200-218) then ERP.plane2spherical (:397-411)
  const float PI = 3.14159274101257324f;
  const float ex = __fsub_rn(remainder_pos(__fadd_rn(__fadd_rn((float)m, fu), 0.5f), (float)W), 0.5f);
  const float ey = fminf(fmaxf(__fadd_rn((float)n, fv), -0.5f), (float)H - 0.5f);
  theta = __fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(ex, 0.5f), 1.0f / (float)W), 0.5f), 2.f), PI);
  phi = __fmul_rn(__fsub_rn(0.5f, __fmul_rn(__fadd_rn(ey, 0.5f), 1.0f / (float)H)), PI);
}

__global__ void __launch_bounds__(256) great_circle_kernel(const float *__restrict__ pred, const float *__restrict__ gt, float *__restrict__ out,
                                                           int B, int H, int W, float R) {
  const long long HW = (long long)H * W, total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW, o = b * 2 * HW + p;
    const int n = (int)(p / W), m = (int)(p - (long long)n * W);
    float tp, pp, tg, pg;
    erp_endpoint_angles(m, n, __ldg(pred + o), __ldg(pred + o + HW), H, W, tp, pp);
    erp_endpoint_angles(m, n, __ldg(gt + o), __ldg(gt + o + HW), H, W, tg, pg);
    // haversine(dphi) + cos(phi_p) cos(phi_g) haversine(dtheta); alpha = 2 asin(sqrt(.))   (spherical.py:43-49,72-84)
    const float s1 = sinf((pg - pp) * 0.5f), s2 = sinf((tg - tp) * 0.5f);
    const float hv = s1 * s1 + cosf(pp) * cosf(pg) * (s2 * s2);
    out[i] = 2.f * asinf(sqrtf(hv)) * R;
  }
}

}  // namespace pf

extern "C" int pf_convex_upsample(const float *flow, const float *mask, float *out, int batch, int h, int w, int mask_channels_last,
                                  void *stream) {
  using namespace pf;
  PF_REQUIRE(flow && mask && out && batch > 0 && h > 0 && w > 0, "pf_convex_upsample: bad arguments");
  const long long N = (long long)h * w;
  const long long m_bs = 576 * N, m_cs = mask_channels_last ? 1 : N, m_ps = mask_channels_last ? 576 : 1;
  if (mask_channels_last)
    convex_upsample_kernel<true><<<dim3(ceil_div(N, 4), batch), 256, 0, (cudaStream_t)stream>>>(flow, mask, out, h, w, m_bs, m_cs, m_ps);
  else
    convex_upsample_kernel<false><<<dim3(ceil_div(N, 4), batch), 256, 0, (cudaStream_t)stream>>>(flow, mask, out, h, w, m_bs, m_cs, m_ps);
  return check_launch("pf_convex_upsample");
}

extern "C" int pf_convex_upsample_bwd(const float *flow, const float *mask, const float *grad_out, float *dflow, float *dmask, int batch,
                                      int h, int w, int mask_channels_last, void *stream) {
  using namespace pf;
  PF_REQUIRE(flow && mask && grad_out && dflow && dmask && batch > 0 && h > 0 && w > 0, "pf_convex_upsample_bwd: bad arguments");
  const long long N = (long long)h * w;
  const long long m_bs = 576 * N, m_cs = mask_channels_last ? 1 : N, m_ps = mask_channels_last ? 576 : 1;
  if (cudaMemsetAsync(dflow, 0, sizeof(float) * 2 * N * batch, (cudaStream_t)stream) != cudaSuccess) return check_launch("pf_convex_upsample_bwd(memset)");
  if (mask_channels_last)
    convex_upsample_bwd_kernel<true><<<dim3(ceil_div(N, 4), batch), 256, 0, (cudaStream_t)stream>>>(flow, mask, grad_out, dflow, dmask, h, w, m_bs, m_cs, m_ps);
  else
    convex_upsample_bwd_kernel<false><<<dim3(ceil_div(N, 4), batch), 256, 0, (cudaStream_t)stream>>>(flow, mask, grad_out, dflow, dmask, h, w, m_bs, m_cs, m_ps);
  return check_launch("pf_convex_upsample_bwd");
}

extern "C" int pf_uniform_loss_fwd(const float *pred, const float *gt, const float *ok, const float *lat, float *acc, float term_weight,
                                   int batch, int H, int W, void *stream) {
  using namespace pf;
  PF_REQUIRE(pred && gt && ok && lat && acc && batch > 0 && H > 0 && W > 0, "pf_uniform_loss_fwd: bad arguments");
  const long long total = (long long)batch * H * W;
  const unsigned grid = (unsigned)(total / 1024 + 1 < 1184 ? total / 1024 + 1 : 1184);
  uniform_loss_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, gt, ok, lat, acc, term_weight, batch, H, W);
  return check_launch("pf_uniform_loss_fwd");
}

extern "C" int pf_uniform_loss_bwd(const float *pred, const float *gt, const float *ok, const float *lat, const float *upstream,
                                   float term_weight, float *dpred, int batch, int H, int W, void *stream) {
  using namespace pf;
  PF_REQUIRE(pred && gt && ok && lat && upstream && dpred && batch > 0 && H > 0 && W > 0, "pf_uniform_loss_bwd: bad arguments");
  const long long total = (long long)batch * H * W;
  const unsigned grid = (unsigned)(total / 1024 + 1 < 1184 ? total / 1024 + 1 : 1184);
  uniform_loss_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, gt, ok, lat, upstream, term_weight, dpred, batch, H, W);
  return check_launch("pf_uniform_loss_bwd");
}

extern "C" int pf_great_circle(const float *pred, const float *gt, float *out, int batch, int H, int W, float radius, void *stream) {
  using namespace pf;
  PF_REQUIRE(pred && gt && out && batch > 0 && H > 0 && W > 0, "pf_great_circle: bad arguments");
  const long long total = (long long)batch * H * W;
  const unsigned grid = (unsigned)(total / 1024 + 1 < 1184 ? total / 1024 + 1 : 1184);
  great_circle_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, gt, out, batch, H, W, radius);
  return check_launch("pf_great_circle");
}
