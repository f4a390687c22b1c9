// Shared device/host helpers for libpriorcorr (sm_100a).
//
// Numerical contract: the coordinate path of every sampler performs the
// same fp32 operations in the same order as the chain of eager ATen kernels the reference runs,
// each op individually rounded.  The *_rn intrinsics below are never contracted into FMAs by
// nvcc, so the file can be compiled with the default -fmad=true (the value path wants FMAs: ATen's
// grid_sampler blends its four taps with `out_acc += v * w`, an FMA chain).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/priorcorr.h"

namespace pf {

void set_error(const char *fmt, ...);
int check_launch(const char *what);

#define PF_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      ::pf::set_error(__VA_ARGS__);      \
      return 1;                          \
    }                                    \
  } while (0)

static inline unsigned ceil_div(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// torch.remainder(x, m) for m > 0: fmod, then + m when the result is negative (may return m itself
// for tiny negative x — SURVEY.md §A.2).
__device__ __forceinline__ float remainder_pos(float x, float m) {
  float r = fmodf(x, m);
  if (r != 0.f && r < 0.f) r = __fadd_rn(r, m);
  return r;
}

// One image axis of the sampler wrappers: size, (size-1) and 1/(size-1) as fp32.
struct Axis {
  float size;      // W (or H)
  float size_m1;   // W - 1
  float inv_m1;    // 1.0f / (W - 1)   (ATen CUDA `tensor / scalar` multiplies by this)
};
static inline Axis make_axis(int size) {
  Axis a;
  a.size = (float)size;
  a.size_m1 = (float)(size - 1);
  a.inv_m1 = 1.0f / a.size_m1;
  return a;
}

// Pixel coordinate -> the unnormalised coordinate ATen's grid_sampler finally uses
// (core/utils/utils.py:85-86 then grid_sampler_unnormalize(align_corners=True) and
// safe_downgrade_to_int_range).  Not the identity: low bits change on the round trip.
__device__ __forceinline__ float to_sample_coord(float p, const Axis ax, int div_mode) {
  float t = __fmul_rn(2.f, p);
  float g = (div_mode == PF_DIV_ATEN_CUDA) ? __fmul_rn(t, ax.inv_m1) : __fdiv_rn(t, ax.size_m1);
  g = __fsub_rn(g, 1.f);
  float v = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), ax.size_m1);
  if (!isfinite(v) || v > 2147483648.f || v < -2147483648.f) v = -100.f;
  return v;
}

// Bilinear weights + integer corners of ATen grid_sampler_2d (bilinear).
struct Taps {
  int x0, y0;
  float nw, ne, sw, se;
};
__device__ __forceinline__ Taps make_taps(float ix, float iy) {
  Taps t;
  float x0f = floorf(ix), y0f = floorf(iy);
  float x1f = __fadd_rn(x0f, 1.f), y1f = __fadd_rn(y0f, 1.f);
  float dxe = __fsub_rn(x1f, ix), dxw = __fsub_rn(ix, x0f);
  float dys = __fsub_rn(y1f, iy), dyn = __fsub_rn(iy, y0f);
  t.nw = __fmul_rn(dxe, dys);
  t.ne = __fmul_rn(dxw, dys);
  t.sw = __fmul_rn(dxe, dyn);
  t.se = __fmul_rn(dxw, dyn);
  t.x0 = (int)x0f;
  t.y0 = (int)y0f;
  return t;
}

// Zero-padded 4-tap blend from one dense [H, W] plane, FMA chain in ATen's order nw, ne, sw, se.
__device__ __forceinline__ float blend_zeros(const float *__restrict__ plane, int H, int W, const Taps &t) {
  const bool xin0 = (unsigned)t.x0 < (unsigned)W, xin1 = (unsigned)(t.x0 + 1) < (unsigned)W;
  const bool yin0 = (unsigned)t.y0 < (unsigned)H, yin1 = (unsigned)(t.y0 + 1) < (unsigned)H;
  const float *r0 = plane + (long long)t.y0 * W + t.x0;
  const float *r1 = r0 + W;
  float v_nw = (yin0 && xin0) ? __ldg(r0) : 0.f;
  float v_ne = (yin0 && xin1) ? __ldg(r0 + 1) : 0.f;
  float v_sw = (yin1 && xin0) ? __ldg(r1) : 0.f;
  float v_se = (yin1 && xin1) ? __ldg(r1 + 1) : 0.f;
  float acc = __fmul_rn(v_nw, t.nw);
  acc = __fmaf_rn(v_ne, t.ne, acc);
  acc = __fmaf_rn(v_sw, t.sw, acc);
  acc = __fmaf_rn(v_se, t.se, acc);
  return acc;
}

}  // namespace pf
