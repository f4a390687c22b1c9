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
// pf_lookup.cu: img_rotate of a channels-last [B, N, L*K2] map into [B, L*K2, N]
int rotate_forward(int batch, int h, int w, int num_levels, int radius, int div_mode, const float *grid_c2w,
                   long long grid_bs, const float *raw, float *out, int channels_last, int fuse_sum, cudaStream_t st,
                   const float *own_cl, bool after_lookup_rows = false);

#define PF_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      ::pf::set_error(__VA_ARGS__);      \
      return 1;                          \
    }                                    \
  } while (0)

static inline unsigned ceil_div(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// Makes a pointer opaque to the optimiser so that `base + int_offset` stays one IMAD.WIDE per access
// instead of being re-derived from the kernel parameters with 64-bit multiplies at every load
// (ncu r01b: half of the lookup kernel's instructions were such address arithmetic).
template <typename T>
__device__ __forceinline__ T *opaque(T *p) {
  asm volatile("" : "+l"(p));
  return p;
}

// Plain (coherent, L1-allocating) 16-byte load: never .nc.  For data written by a grid that may still be running when
// the reading kernel starts (programmatic dependent launch): PTX requires ld.global.nc data to be read-only for the whole
// lifetime of the kernel, so __ldg is off the table there even after griddepcontrol.wait.
__device__ __forceinline__ float4 ld_f4(const float4 *p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// L2 eviction policies for the pyramid reads of the lookup (createpolicy + ld.global.nc.L2::cache_hint).  The small levels
// of a pyramid (levels 2-3: 20 MB per view at 512x1024) are re-read by every one of the 24 lookup calls of a forward, the
// large ones (levels 0-1: 320 MB) stream through once per call: evict_last for the former keeps them L2-resident across
// calls even though ~0.5 GB of other traffic passes through the 126 MB L2 in between, evict_first for the latter keeps the
// stream from displacing them.
enum L2Policy { kL2Normal = 0, kL2First = 1, kL2Last = 2 };
__device__ __forceinline__ uint64_t make_l2_policy(int kind) {
  uint64_t pol;
  if (kind == kL2Last)
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else if (kind == kL2First)
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float ld_hint(const float *p, uint64_t pol) {            // read-only data, L1 allocate
  float v;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ld_hint_stream(const float *p, uint64_t pol) {     // read exactly once by this kernel: no L1 allocate
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}

// torch.remainder(x, m) for m > 0: fmod, then + m when the result is negative (may return m itself
// for tiny negative x — SURVEY.md §A.2).  The three fast paths are exact restatements of
// fmodf + fix-up for |x| < 2m (fmod is exact; x - m is exact by Sterbenz for m <= x < 2m) and
// cover every coordinate a sane flow produces; anything else takes the library fmodf.
__device__ __forceinline__ float remainder_pos(float x, float m) {
  if (x >= 0.f) {
    if (x < m) return x;
    if (x < __fadd_rn(m, m)) return __fsub_rn(x, m);
  } else if (x > -m) {
    return __fadd_rn(x, m);
  }
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

// Branch-free form of the zero-padded blend for channel loops: out-of-bounds taps get weight 0
// and a clamped (valid) offset, so `fma(v, 0, acc) == acc` reproduces ATen's skipped tap exactly
// for finite data.  Offsets are relative to the plane base.
struct Taps4 {
  int o_nw, o_ne, o_sw, o_se;
  float nw, ne, sw, se;
};
__device__ __forceinline__ Taps4 clamp_taps(const Taps &t, int H, int W) {
  Taps4 c;
  const bool xin0 = (unsigned)t.x0 < (unsigned)W, xin1 = (unsigned)(t.x0 + 1) < (unsigned)W;
  const bool yin0 = (unsigned)t.y0 < (unsigned)H, yin1 = (unsigned)(t.y0 + 1) < (unsigned)H;
  const int x0 = min(max(t.x0, 0), W - 1), x1 = min(max(t.x0 + 1, 0), W - 1);
  const int y0 = min(max(t.y0, 0), H - 1), y1 = min(max(t.y0 + 1, 0), H - 1);
  c.o_nw = y0 * W + x0;
  c.o_ne = y0 * W + x1;
  c.o_sw = y1 * W + x0;
  c.o_se = y1 * W + x1;
  c.nw = (yin0 && xin0) ? t.nw : 0.f;
  c.ne = (yin0 && xin1) ? t.ne : 0.f;
  c.sw = (yin1 && xin0) ? t.sw : 0.f;
  c.se = (yin1 && xin1) ? t.se : 0.f;
  return c;
}
__device__ __forceinline__ float blend4(const float *__restrict__ plane, const Taps4 &c) {
  float acc = __fmul_rn(__ldg(plane + c.o_nw), c.nw);
  acc = __fmaf_rn(__ldg(plane + c.o_ne), c.ne, acc);
  acc = __fmaf_rn(__ldg(plane + c.o_sw), c.sw, acc);
  acc = __fmaf_rn(__ldg(plane + c.o_se), c.se, acc);
  return acc;
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

// to_sample_coord with the exact scalings folded: 2p*inv == 2*(p*inv) (power-of-two scaling), fl(2q - 1) is one FMA,
// and ((g+1) * 0.5) * (W-1) == (g+1) * ((W-1)/2) because the halving is exact.  4 instructions instead of 6.
template <int kDiv>
__device__ __forceinline__ float sample_coord_x(float p, const Axis ax, const float half_m1) {
  const float q = (kDiv == PF_DIV_ATEN_CUDA) ? __fmul_rn(p, ax.inv_m1) : __fdiv_rn(p, ax.size_m1);
  const float g = __fmaf_rn(q, 2.f, -1.f);
  float v = __fmul_rn(__fadd_rn(g, 1.f), half_m1);
  if (!(fabsf(v) <= 2147483648.f)) v = -100.f;
  return v;
}

// torch.remainder(x, m), m > 0.  For a power-of-two m the quotient, its truncation, the product and the
// difference are all exact, so three instructions reproduce fmodf; otherwise the general routine.
static __device__ __noinline__ float remainder_general(float x, float m) { return remainder_pos(x, m); }   // one copy of fmodf's slow path
__device__ __forceinline__ float remainder_sel(float x, const Axis ax, const bool pow2) {
  if (pow2) {
    const float t = truncf(__fmul_rn(x, 1.0f / ax.size));
    float r = __fmaf_rn(-t, ax.size, x);
    if (r < 0.f) r = __fadd_rn(r, ax.size);
    return r;
  }
  return remainder_general(x, ax.size);
}

}  // namespace pf
