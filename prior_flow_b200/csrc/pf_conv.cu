// (f1) img_rotate + `own + other` + the motion encoder's 1x1 convolution + bias + ReLU in ONE kernel — tcgen05 / TMEM / TMA.
//
// Reference: DCCL.__call__ ends with img_rotate of the other view's [B,324,h,w] map (PriOr-RAFT/core/corr.py:137-138),
// PriOr_RAFT.forward adds it to the own view (core/prior_raft.py:187-188) and hands the sum to the motion encoder, whose
// first layer is `F.relu(self.convc1_A(corr_A))` / `F.relu(self.convc1(corr))` — Conv2d(324, 256, 1) (core/update.py:168,184
// and :85,92).  Eager, that is a 10.6 MB [B,324,h,w] tensor written and re-read per call, 24 times per pair.  Here the sum
// never exists in memory:
//     out[n, o] = relu(bias[o] + sum_c W[o, c] * X[n, c]),     X[n, c] = own[n, c] + sum_t w_t(n) raw[src_t(n), c]
// (raw / own: the channels-last maps lookup_rows_kernel leaves behind; taps = the module-local cyclic sampler of
// projection_prim_ortho.py:119-135).
//
// One CTA = 64 pixels x all 256 output channels, D^T = W X^T so that the ACCUMULATOR ROW IS THE OUTPUT CHANNEL:
//   * 16 gather warps build X for their 64 pixels exactly as rotate_fwd_kernel does (same FMA chain, same fp32 add), split
//     it into fp16 hi/lo and store it straight into shared memory in the K-major SWIZZLE_128B layout tcgen05 reads
//     (B operand, N = 64 pixels, K = 324 padded to 384);
//   * W (A operand, two M = 128 halves) is pre-split on the host into K-major fp16 hi/lo planes and streamed k-block by
//     k-block with TMA (first stages in flight before the producer grid has even finished: programmatic dependent launch);
//   * one thread issues tcgen05.mma kind::f16 (M128 N64 K16): fp32 mode = three products lo*hi + hi*lo + hi*hi into one fp32
//     TMEM accumulator (the volume kernel's scheme, fp32-class accuracy); TF32-class mode = hi*hi only — what cuDNN runs for
//     this convolution when torch.backends.cudnn.allow_tf32 is on, the reference's default on a GPU;
//   * 8 epilogue warps read TMEM (lane = output channel), add the bias, apply the ReLU and write channels-last rows (32
//     consecutive channels per warp instruction = full 128-byte lines) or NCHW.
#include <stdlib.h>

#include "pf_tc.cuh"

namespace pf {

constexpr int CV_PIX = 64;                      // pixels per CTA == UMMA N
constexpr int CV_OC = 256;                      // output channels == 2 x UMMA M
constexpr int CV_C = 324;                       // input channels (4 levels x 81 taps)
constexpr int CV_KPAD = 384;                    // K padded to 6 k-blocks of 64 fp16 (one 128-byte swizzle row each)
constexpr int CV_BK = 64, CV_NKB = CV_KPAD / CV_BK;
constexpr int CV_XKB = CV_PIX * CV_BK * 2;      // 8 KiB: X tile, one k-block, one plane
constexpr int CV_XPLANE = CV_NKB * CV_XKB;      // 48 KiB
constexpr int CV_WPLANE = CV_OC * CV_BK * 2;    // 32 KiB: W tile, one k-block, one plane (256 rows x 128 B)
constexpr int CV_GATHER_WARPS = 16;
constexpr int CV_THREADS = (CV_GATHER_WARPS + 1) * 32;
constexpr uint32_t CV_TMEM_COLS = 128;          // two accumulator halves of 64 fp32 columns
constexpr uint32_t CV_IDESC = umma_idesc_f16(128, CV_PIX);

template <bool kSplit>
struct ConvCfg {
  static constexpr int kPlanes = kSplit ? 2 : 1;
  static constexpr int kWStages = kSplit ? 2 : 4;
  static constexpr int kWStageBytes = kPlanes * CV_WPLANE;
  static constexpr int kXBytes = kPlanes * CV_XPLANE;
  static constexpr int kUsedBytes = kXBytes + kWStages * kWStageBytes + CV_PIX * (int)sizeof(Taps4) + 256 /*barriers*/;
  // split mode uses 226.25 KiB of the 227 KiB a CTA may have: the 1 KiB alignment slack does not fit entirely.  The dynamic
  // window starts 1 KiB aligned when the kernel has no static shared memory (it has none); the kernel traps if it ever does not.
  static constexpr int kSmemBytes = (kUsedBytes + 1024 <= 232448) ? kUsedBytes + 1024 : 232448;
};

struct ConvParams {
  int B, N, h, w, div_mode, out_channels_last;
  Axis axW, axH;
  const float *grid_c2w;
  long long grid_bs;
  const float *raw, *own_cl;   // [B, N, 324]
  const float *bias;           // [256]
  const float *inv_w_scale;    // device scalar: 1 / (power-of-two scale folded into the fp16 weight planes)
  float *out;                  // [B, N, 256] or [B, 256, N]
};

__device__ __forceinline__ void gather_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CV_GATHER_WARPS * 32) : "memory"); }

template <bool kSplit>
__global__ void __launch_bounds__(CV_THREADS, 1)
dccl_conv_kernel(const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo, const ConvParams p) {
  using C_ = ConvCfg<kSplit>;
  constexpr int kWStages = C_::kWStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  if (smem + C_::kUsedBytes > smem_raw + C_::kSmemBytes) __trap();   // see ConvCfg::kSmemBytes
  uint8_t *xs = smem;                                        // X planes: hi [6][64 rows][128 B], then lo
  uint8_t *ws = smem + C_::kXBytes;                          // W stages: hi [256 rows][128 B], then lo
  Taps4 *taps = reinterpret_cast<Taps4 *>(ws + kWStages * C_::kWStageBytes);
  uint64_t *bars = reinterpret_cast<uint64_t *>(taps + CV_PIX);
  const uint32_t bar_wfull = smem_u32(bars), bar_wempty = smem_u32(bars + kWStages);
  const uint32_t bar_xready = smem_u32(bars + 2 * kWStages), bar_acc = smem_u32(bars + 2 * kWStages + 1);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kWStages + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, n0 = blockIdx.x * CV_PIX;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(bar_wfull + 8 * s, 1);
      mbar_init(bar_wempty + 8 * s, 1);
    }
    mbar_init(bar_xready, CV_GATHER_WARPS);
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CV_GATHER_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(CV_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto load_w = [&](int kb, int stage) {   // one k-block of the weights: 256 rows x 64 k per plane
    const uint32_t full = bar_wfull + 8 * stage;
    mbar_arrive_expect_tx(full, (uint32_t)C_::kWStageBytes);
    const uint32_t dst = smem_u32(ws + stage * C_::kWStageBytes);
    tma_load_2d(dst, &map_w_hi, full, kb * CV_BK, 0);
    if (kSplit) tma_load_2d(dst + CV_WPLANE, &map_w_lo, full, kb * CV_BK, 0);
  };

  if (warp == CV_GATHER_WARPS) {
    // ================================================================= weights producer + MMA issuer (one thread)
    if (lane == 0) {
      // the weights are an input of the whole forward: their first stages are in flight while the producer grid still runs
      for (int kb = 0; kb < kWStages && kb < CV_NKB; ++kb) load_w(kb, kb);
      mbar_wait(bar_xready, 0);
      tc_fence_after();
      for (int kb = 0; kb < CV_NKB; ++kb) {
        const int stage = kb % kWStages;
        mbar_wait(bar_wfull + 8 * stage, (uint32_t)(kb / kWStages) & 1);
        tc_fence_after();
        const uint32_t wbase = smem_u32(ws + stage * C_::kWStageBytes), xbase = smem_u32(xs + kb * CV_XKB);
        const uint64_t x_hi = make_smem_desc(xbase), x_lo = make_smem_desc(xbase + CV_XPLANE);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t d = tmem_base + half * CV_PIX;
          const uint64_t w_hi = make_smem_desc(wbase + half * (CV_WPLANE / 2));
          const uint64_t w_lo = make_smem_desc(wbase + CV_WPLANE + half * (CV_WPLANE / 2));
          if (kSplit) {
            // small cross terms first, then the leading product
#pragma unroll
            for (int k = 0; k < CV_BK / 16; ++k) umma_f16(d, w_lo + 2 * k, x_hi + 2 * k, CV_IDESC, (kb | k) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < CV_BK / 16; ++k) umma_f16(d, w_hi + 2 * k, x_lo + 2 * k, CV_IDESC, 1u);
#pragma unroll
            for (int k = 0; k < CV_BK / 16; ++k) umma_f16(d, w_hi + 2 * k, x_hi + 2 * k, CV_IDESC, 1u);
          } else {
#pragma unroll
            for (int k = 0; k < CV_BK / 16; ++k) umma_f16(d, w_hi + 2 * k, x_hi + 2 * k, CV_IDESC, (kb | k) ? 1u : 0u);
          }
        }
        umma_commit(bar_wempty + 8 * stage);    // the stage may be refilled when these MMAs have retired
        if (kb + kWStages < CV_NKB) {           // refill: wait for exactly that, then fetch k-block kb + kWStages
          mbar_wait(bar_wempty + 8 * stage, (uint32_t)(kb / kWStages) & 1);
          load_w(kb + kWStages, stage);
        }
      }
      umma_commit(bar_acc);                     // accumulators complete -> epilogue
    }
  } else {
    // ================================================================= gather warps: X tile -> swizzled K-major fp16 planes
    // zero the last k-block (k = 320..383; 320..323 are overwritten below): padding must be exact zeros, not stale smem
    for (int i = threadIdx.x; i < C_::kPlanes * (CV_XKB / 16); i += CV_GATHER_WARPS * 32) {
      const int plane = i / (CV_XKB / 16), o = i - plane * (CV_XKB / 16);
      *reinterpret_cast<uint4 *>(xs + plane * CV_XPLANE + (CV_NKB - 1) * CV_XKB + o * 16) = make_uint4(0, 0, 0, 0);
    }
    if (threadIdx.x < CV_PIX) {
      // module-local cyclic sampler of projection_prim_ortho.py:119-135 on the [.., h, w] map (reads only grid_c2w, an input)
      const float *gx = p.grid_c2w + (long long)b * p.grid_bs;
      const int n = min(n0 + (int)threadIdx.x, p.N - 1);
      const float x = remainder_pos(__ldg(gx + n), p.axW.size);
      const float y = __ldg(gx + p.N + n);
      taps[threadIdx.x] = clamp_taps(make_taps(to_sample_coord(x, p.axW, p.div_mode), to_sample_coord(y, p.axH, p.div_mode)), p.h, p.w);
    }
    // everything above touches inputs only; raw / own_cl are read after the producer grid has completed and flushed
    asm volatile("griddepcontrol.wait;" ::: "memory");
    gather_bar_sync();
    constexpr int C4 = CV_C / 4;   // 81 float4 per pixel
    const float4 *src = reinterpret_cast<const float4 *>(p.raw + (long long)b * p.N * CV_C);
    const float4 *own = reinterpret_cast<const float4 *>(p.own_cl + (long long)b * p.N * CV_C);
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int r0 = warp * 4 + pass * 2;           // two pixels (tile rows) per pass
      Taps4 t[2];
      t[0] = taps[r0], t[1] = taps[r0 + 1];
      const int n_a = min(n0 + r0, p.N - 1), n_b = min(n0 + r0 + 1, p.N - 1);
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
        const int c4 = lane + 32 * rr;
        if (c4 < C4) {
          float4 X[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float4 A = ld_f4(src + (long long)t[i].o_nw * C4 + c4), Bq = ld_f4(src + (long long)t[i].o_ne * C4 + c4);
            const float4 Cq = ld_f4(src + (long long)t[i].o_sw * C4 + c4), D = ld_f4(src + (long long)t[i].o_se * C4 + c4);
            const float4 O = ld_f4(own + (long long)(i ? n_b : n_a) * C4 + c4);
            float4 r;   // rotate_fwd_kernel's chain, then corr_A + corr_B_A (core/prior_raft.py:187) as one fp32 add
            r.x = __fmaf_rn(D.x, t[i].se, __fmaf_rn(Cq.x, t[i].sw, __fmaf_rn(Bq.x, t[i].ne, __fmul_rn(A.x, t[i].nw))));
            r.y = __fmaf_rn(D.y, t[i].se, __fmaf_rn(Cq.y, t[i].sw, __fmaf_rn(Bq.y, t[i].ne, __fmul_rn(A.y, t[i].nw))));
            r.z = __fmaf_rn(D.z, t[i].se, __fmaf_rn(Cq.z, t[i].sw, __fmaf_rn(Bq.z, t[i].ne, __fmul_rn(A.z, t[i].nw))));
            r.w = __fmaf_rn(D.w, t[i].se, __fmaf_rn(Cq.w, t[i].sw, __fmaf_rn(Bq.w, t[i].ne, __fmul_rn(A.w, t[i].nw))));
            X[i] = make_float4(__fadd_rn(O.x, r.x), __fadd_rn(O.y, r.y), __fadd_rn(O.z, r.z), __fadd_rn(O.w, r.w));
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int r = r0 + i;
            const int k = 4 * c4, kb = k >> 6, chunk = (k & 63) >> 3;
            uint8_t *dst = xs + kb * CV_XKB + (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4) + (c4 & 1) * 8;
            const __half h0 = __float2half_rn(X[i].x), h1 = __float2half_rn(X[i].y), h2 = __float2half_rn(X[i].z), h3 = __float2half_rn(X[i].w);
            __half2 a = __halves2half2(h0, h1), c = __halves2half2(h2, h3);
            *reinterpret_cast<uint2 *>(dst) = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&c));
            if (kSplit) {
              __half2 la = __halves2half2(__float2half_rn(X[i].x - __half2float(h0)), __float2half_rn(X[i].y - __half2float(h1)));
              __half2 lc = __halves2half2(__float2half_rn(X[i].z - __half2float(h2)), __float2half_rn(X[i].w - __half2float(h3)));
              *reinterpret_cast<uint2 *>(dst + CV_XPLANE) = make_uint2(*reinterpret_cast<uint32_t *>(&la), *reinterpret_cast<uint32_t *>(&lc));
            }
          }
        }
      }
    }
    fence_async_smem();            // generic-proxy writes of the X tile -> visible to the tensor core's (async proxy) reads
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_xready);

    // ================================================================= epilogue: warps 0..7 (TMEM lane = output channel)
    if (warp < 8) {
      const int quarter = warp & 3, half = warp >> 2;
      const int o = half * 128 + quarter * 32 + lane;
      const float bias = __ldg(p.bias + o), inv_s = __ldg(p.inv_w_scale);
      mbar_wait(bar_acc, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * CV_PIX;
      uint32_t v0[32], v1[32];
      tmem_ld32(taddr, v0);
      tmem_ld32(taddr + 32, v1);
      tmem_ld_wait();
      const int live = min(CV_PIX, p.N - n0);
      if (p.out_channels_last) {
        float *ob = p.out + ((long long)b * p.N + n0) * CV_OC + o;     // + j * 256: a warp writes 32 consecutive channels
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < live) ob[(long long)j * CV_OC] = fmaxf(__fmaf_rn(__uint_as_float(v0[j]), inv_s, bias), 0.f);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (32 + j < live) ob[(long long)(32 + j) * CV_OC] = fmaxf(__fmaf_rn(__uint_as_float(v1[j]), inv_s, bias), 0.f);
      } else {
        float *ob = p.out + ((long long)b * CV_OC + o) * p.N + n0;     // 64 consecutive pixels of channel o
        if (live == CV_PIX && (p.N & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4 *>(ob + 4 * j) = make_float4(fmaxf(__fmaf_rn(__uint_as_float(v0[4 * j]), inv_s, bias), 0.f),
                                                                   fmaxf(__fmaf_rn(__uint_as_float(v0[4 * j + 1]), inv_s, bias), 0.f),
                                                                   fmaxf(__fmaf_rn(__uint_as_float(v0[4 * j + 2]), inv_s, bias), 0.f),
                                                                   fmaxf(__fmaf_rn(__uint_as_float(v0[4 * j + 3]), inv_s, bias), 0.f));
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4 *>(ob + 32 + 4 * j) = make_float4(fmaxf(__fmaf_rn(__uint_as_float(v1[4 * j]), inv_s, bias), 0.f),
                                                                        fmaxf(__fmaf_rn(__uint_as_float(v1[4 * j + 1]), inv_s, bias), 0.f),
                                                                        fmaxf(__fmaf_rn(__uint_as_float(v1[4 * j + 2]), inv_s, bias), 0.f),
                                                                        fmaxf(__fmaf_rn(__uint_as_float(v1[4 * j + 3]), inv_s, bias), 0.f));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < live) ob[j] = fmaxf(__fmaf_rn(__uint_as_float(v0[j]), inv_s, bias), 0.f);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (32 + j < live) ob[32 + j] = fmaxf(__fmaf_rn(__uint_as_float(v1[j]), inv_s, bias), 0.f);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == CV_GATHER_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(CV_TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------- weight preparation
// W [256, 324] fp32 -> K-major fp16 planes [256, 384] (zero padded), hi = fp16(W s), lo = fp16(W s - hi), s a power of two
// chosen from absmax (the volume kernel's split_scale) so that neither plane leaves fp16's normal range.
__global__ void __launch_bounds__(256) conv_weight_absmax_kernel(const float *__restrict__ w, int n, uint32_t *__restrict__ out) {
  uint32_t m = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, __float_as_uint(fabsf(w[i])));
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

__device__ __forceinline__ float conv_split_scale(uint32_t amax_bits) {
  int e = (int)((amax_bits >> 23) & 0xff) - 127;
  if ((amax_bits & 0x7fffffffu) == 0u) e = 13;
  int se = 13 - e;
  se = se < -100 ? -100 : (se > 100 ? 100 : se);
  return __uint_as_float((uint32_t)(se + 127) << 23);
}

__global__ void __launch_bounds__(256) conv_weight_split_kernel(const float *__restrict__ w, int O, int C, int Kpad, const uint32_t *__restrict__ amax,
                                                                 __half *__restrict__ hi, __half *__restrict__ lo, float *__restrict__ inv_scale) {
  const float s = conv_split_scale(*amax);
  if (blockIdx.x == 0 && threadIdx.x == 0) *inv_scale = 1.0f / s;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < O * Kpad; i += gridDim.x * blockDim.x) {
    const int o = i / Kpad, k = i - o * Kpad;
    const float v = k < C ? w[o * C + k] * s : 0.f;
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
  }
}

}  // namespace pf

extern "C" long long pf_dccl_conv_weight_bytes(void) { return 2LL * pf::CV_OC * pf::CV_KPAD * 2 + 256; }

extern "C" int pf_dccl_conv_prepare(const float *weight, int out_channels, int in_channels, void *prepared, void *stream) {
  using namespace pf;
  PF_REQUIRE(weight && prepared, "pf_dccl_conv_prepare: null pointer");
  PF_REQUIRE(out_channels == CV_OC && in_channels == CV_C, "pf_dccl_conv_prepare: the kernel is built for Conv2d(%d, %d, 1) (got %d -> %d)", CV_C,
             CV_OC, in_channels, out_channels);
  PF_REQUIRE(((uintptr_t)prepared & 1023) == 0, "pf_dccl_conv_prepare: the prepared buffer must be 1 KiB aligned");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t *base = reinterpret_cast<uint8_t *>(prepared);
  __half *hi = reinterpret_cast<__half *>(base), *lo = hi + CV_OC * CV_KPAD;
  uint32_t *amax = reinterpret_cast<uint32_t *>(base + 2LL * CV_OC * CV_KPAD * 2);
  float *inv_scale = reinterpret_cast<float *>(amax + 1);
  if (cudaMemsetAsync(amax, 0, 8, st) != cudaSuccess) return check_launch("pf_dccl_conv_prepare(memset)");
  conv_weight_absmax_kernel<<<32, 256, 0, st>>>(weight, CV_OC * CV_C, amax);
  conv_weight_split_kernel<<<96, 256, 0, st>>>(weight, CV_OC, CV_C, CV_KPAD, amax, hi, lo, inv_scale);
  return check_launch("pf_dccl_conv_prepare");
}

extern "C" int pf_dccl_conv(const pf_dccl_conv_args *a, void *stream) {
  using namespace pf;
  PF_REQUIRE(a != nullptr, "pf_dccl_conv: null args");
  PF_REQUIRE(a->raw && a->own_cl && a->grid_c2w && a->prepared_weight && a->bias && a->out, "pf_dccl_conv: null pointer");
  PF_REQUIRE(a->batch > 0 && a->h > 0 && a->w > 0, "pf_dccl_conv: bad shape");
  PF_REQUIRE(a->in_channels == CV_C && a->out_channels == CV_OC, "pf_dccl_conv: built for Conv2d(%d, %d, 1) (got %d -> %d)", CV_C, CV_OC,
             a->in_channels, a->out_channels);
  PF_REQUIRE((((uintptr_t)a->raw | (uintptr_t)a->own_cl | (uintptr_t)a->out) & 15) == 0 && ((uintptr_t)a->prepared_weight & 1023) == 0,
             "pf_dccl_conv: raw / own_cl / out must be 16-byte aligned, the prepared weights 1 KiB aligned");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t *base = reinterpret_cast<uint8_t *>(const_cast<void *>(a->prepared_weight));
  __half *hi = reinterpret_cast<__half *>(base), *lo = hi + CV_OC * CV_KPAD;
  const float *inv_scale_dev = reinterpret_cast<const float *>(base + 2LL * CV_OC * CV_KPAD * 2 + 4);
  CUtensorMap m_hi, m_lo;
  {
    cuuint64_t dims[2] = {(cuuint64_t)CV_KPAD, (cuuint64_t)CV_OC};
    cuuint64_t strides[1] = {(cuuint64_t)CV_KPAD * 2};
    cuuint32_t box[2] = {CV_BK, CV_OC};
    if (int e = encode(&m_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, hi, dims, strides, box, "W.hi")) return e;
    if (int e = encode(&m_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, lo, dims, strides, box, "W.lo")) return e;
  }
  ConvParams p;
  p.B = a->batch, p.N = a->h * a->w, p.h = a->h, p.w = a->w, p.div_mode = a->div_mode, p.out_channels_last = a->out_channels_last;
  p.axW = make_axis(a->w), p.axH = make_axis(a->h);
  p.grid_c2w = a->grid_c2w, p.grid_bs = a->grid_batch_stride;
  p.raw = a->raw, p.own_cl = a->own_cl, p.bias = a->bias, p.out = a->out;
  p.inv_w_scale = inv_scale_dev;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ceil_div(p.N, CV_PIX), p.B);
  cfg.blockDim = dim3(CV_THREADS);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = !(getenv("PF_ROTATE_PDL") != nullptr && getenv("PF_ROTATE_PDL")[0] == '0');
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && a->after_lookup) ? 1 : 0;
  auto launch = [&](auto kern, int smem) -> cudaError_t {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cfg.dynamicSmemBytes = smem;
    return cudaLaunchKernelEx(&cfg, kern, m_hi, m_lo, p);
  };
  const cudaError_t err = a->split ? launch(dccl_conv_kernel<true>, ConvCfg<true>::kSmemBytes) : launch(dccl_conv_kernel<false>, ConvCfg<false>::kSmemBytes);
  if (err != cudaSuccess) {
    set_error("pf_dccl_conv: launch failed: %s", cudaGetErrorString(err));
    return 2;
  }
  return check_launch("pf_dccl_conv");
}
