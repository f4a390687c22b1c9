// (e) Adjoints of the all-pairs contraction (autograd of PriOr-RAFT/core/prior_raft.py:73-75) on tcgen05:
//     dF1[b, c, n] = 1/sqrt(C) * sum_m dV[b, n, m] F2[b, c, m]          dF2[b, c, m] = 1/sqrt(C) * sum_n dV[b, n, m] F1[b, c, n]
// dV is the level-0 gradient volume after pf_pyramid_fold_bwd: [B, N, N] fp32, 256 MiB per view at 512x1024 — the operand that
// has to be streamed from HBM, once per gradient.  The reference leaves these two GEMMs to cuBLAS fp32 (CUDA cores).
//
// One launch computes both gradients: blockIdx.y selects the gradient, a CTA owns 128 output positions (n for dF1, m for dF2)
// x all C = 256 channels and walks the contraction axis in k-blocks of 64.
//   * 16 converter warps read the fp32 dV tile straight from global memory (coalesced either way: along m for dF1, and for
//     dF2 — where the tile is needed TRANSPOSED — lanes run along m while a thread collects 8 consecutive n), split it into
//     bf16 hi + lo and store it K-major / SWIZZLE_128B into a 2-stage shared-memory ring (A operand).  bf16 keeps fp32's
//     exponent, so gradients of any magnitude need no scaling pass; hi + lo carries 16 significant bits.
//   * one producer thread streams the feature planes (pre-split once into bf16 hi/lo [B, C, N], already K-major) with TMA
//     (B operand, N = 256 channels); one thread issues tcgen05.mma kind::f16 (bf16 inputs, fp32 accumulate in TMEM), three
//     products lo*hi + hi*lo + hi*hi per k-step: relative error ~2^-16, far inside what a gradient needs (cuBLAS under
//     allow_tf32 gives 2^-11) and checked against the fp32 GEMM in tests/test_gpu_parity.py.
//   * 4 epilogue warps read the 128 x 256 accumulator (TMEM lane = output position) and write [B, C, N] rows coalesced.
// HBM: 2 x 256 MiB of dV reads per view (one pass per gradient); tensor time 2 x 34.4 GFLOP x 3 products.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "pf_tc.cuh"

namespace pf {

constexpr int VB_M = 128;                      // output positions per CTA == TMEM lanes
constexpr int VB_C = 256;                      // channels == UMMA N
constexpr int VB_BK = 64;                      // contraction elements per k-block (one 128-byte swizzle row of bf16)
constexpr int VB_APLANE = VB_M * VB_BK * 2;    // 16 KiB
constexpr int VB_BPLANE = VB_C * VB_BK * 2;    // 32 KiB
constexpr int VB_ASTAGES = 2, VB_BSTAGES = 2;
constexpr int VB_CONV_WARPS = 16;
constexpr int VB_THREADS = (VB_CONV_WARPS + 2) * 32;   // + TMA producer warp + MMA warp
constexpr int VB_SMEM = VB_ASTAGES * 2 * VB_APLANE + VB_BSTAGES * 2 * VB_BPLANE + 256 + 1024;
// kind::f16 instruction descriptor with BF16 inputs: D = F32 (bit 4), A = BF16 (bit 7), B = BF16 (bit 10), K-major both
constexpr uint32_t VB_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(VB_C >> 3) << 17) | ((uint32_t)(VB_M >> 4) << 24);

struct VolBwdParams {
  int B, N;                // N = h*w: pitch of the feature planes and of the outputs
  int Nq, q_begin;         // dV holds the rows of queries [q_begin, q_begin + Nq) only (Nq == N, q_begin == 0: the whole volume)
  int accumulate2;         // dF2 += instead of = (chunked backward: one chunk of query rows per call)
  float scale;             // 1 / sqrt(C)
  const float *dV;         // [B, Nq, N]
  float *dF1, *dF2;        // [B, C, N] (either may be null); dF1 is written for the chunk's queries only
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}

// 8 consecutive-k fp32 values -> one 16-byte chunk of the hi plane and one of the lo plane
__device__ __forceinline__ void store_chunk(uint8_t *a_stage, int r, int chunk, const float (&x)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]), h1 = __float2bfloat16_rn(x[2 * i + 1]);
    hi[i] = pack_bf16(__bfloat162float(h0), __bfloat162float(h1));
    lo[i] = pack_bf16(x[2 * i] - __bfloat162float(h0), x[2 * i + 1] - __bfloat162float(h1));
  }
  uint8_t *dst = a_stage + (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4);
  *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4 *>(dst + VB_APLANE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

__global__ void __launch_bounds__(VB_THREADS, 1)
volume_bwd_kernel(const __grid_constant__ CUtensorMap map_f1_hi, const __grid_constant__ CUtensorMap map_f1_lo,
                  const __grid_constant__ CUtensorMap map_f2_hi, const __grid_constant__ CUtensorMap map_f2_lo, const VolBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *as = smem;                                           // A stages: [hi 16 KiB | lo 16 KiB]
  uint8_t *bs = smem + VB_ASTAGES * 2 * VB_APLANE;              // B stages: [hi 32 KiB | lo 32 KiB]
  uint64_t *bars = reinterpret_cast<uint64_t *>(bs + VB_BSTAGES * 2 * VB_BPLANE);
  const uint32_t bar_afull = smem_u32(bars), bar_aempty = smem_u32(bars + VB_ASTAGES);
  const uint32_t bar_bfull = smem_u32(bars + 2 * VB_ASTAGES), bar_bempty = smem_u32(bars + 2 * VB_ASTAGES + VB_BSTAGES);
  const uint32_t bar_acc = smem_u32(bars + 2 * VB_ASTAGES + 2 * VB_BSTAGES);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * VB_ASTAGES + 2 * VB_BSTAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int which = blockIdx.y, b = blockIdx.z;
  const int row0 = blockIdx.x * VB_M;            // first output position (query of the chunk for dF1, target m for dF2)
  const int rows = which ? p.N : p.Nq;           // output positions of this gradient
  const int kblocks = (which ? p.Nq : p.N) / VB_BK;   // contraction length: targets for dF1, the chunk's queries for dF2
  float *out = which ? p.dF2 : p.dF1;
  if (out == nullptr || row0 >= rows) return;    // uniform per CTA: that gradient is not wanted / beyond its rows

  if (threadIdx.x == 0) {
    for (int s = 0; s < VB_ASTAGES; ++s) {
      mbar_init(bar_afull + 8 * s, VB_CONV_WARPS);
      mbar_init(bar_aempty + 8 * s, 1);
    }
    for (int s = 0; s < VB_BSTAGES; ++s) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == VB_CONV_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)VB_C) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == VB_CONV_WARPS) {
    // ================================================================= TMA producer: feature planes of the OTHER map
    if (lane == 0) {
      const CUtensorMap *mh = which ? &map_f1_hi : &map_f2_hi, *ml = which ? &map_f1_lo : &map_f2_lo;
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % VB_BSTAGES;
        mbar_wait(bar_bempty + 8 * s, ((uint32_t)(kb / VB_BSTAGES) & 1) ^ 1);
        const uint32_t full = bar_bfull + 8 * s;
        mbar_arrive_expect_tx(full, 2u * VB_BPLANE);
        const uint32_t dst = smem_u32(bs + s * 2 * VB_BPLANE);
        const int k0 = (which ? p.q_begin : 0) + kb * VB_BK;     // dF2 contracts over the chunk's queries: F1 columns q_begin + ...
        tma_load_2d(dst, mh, full, k0, b * VB_C);
        tma_load_2d(dst + VB_BPLANE, ml, full, k0, b * VB_C);
      }
    }
  } else if (warp == VB_CONV_WARPS + 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const int sa = kb % VB_ASTAGES, sb = kb % VB_BSTAGES;
        mbar_wait(bar_afull + 8 * sa, (uint32_t)(kb / VB_ASTAGES) & 1);
        mbar_wait(bar_bfull + 8 * sb, (uint32_t)(kb / VB_BSTAGES) & 1);
        tc_fence_after();
        const uint32_t abase = smem_u32(as + sa * 2 * VB_APLANE), bbase = smem_u32(bs + sb * 2 * VB_BPLANE);
        const uint64_t a_hi = make_smem_desc(abase), a_lo = make_smem_desc(abase + VB_APLANE);
        const uint64_t b_hi = make_smem_desc(bbase), b_lo = make_smem_desc(bbase + VB_BPLANE);
#pragma unroll
        for (int k = 0; k < VB_BK / 16; ++k) umma_f16(tmem_base, a_lo + 2 * k, b_hi + 2 * k, VB_IDESC, (kb | k) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < VB_BK / 16; ++k) umma_f16(tmem_base, a_hi + 2 * k, b_lo + 2 * k, VB_IDESC, 1u);
#pragma unroll
        for (int k = 0; k < VB_BK / 16; ++k) umma_f16(tmem_base, a_hi + 2 * k, b_hi + 2 * k, VB_IDESC, 1u);
        umma_commit(bar_aempty + 8 * sa);
        umma_commit(bar_bempty + 8 * sb);
      }
      umma_commit(bar_acc);
    }
  } else {
    // ================================================================= converter warps: fp32 dV tile -> bf16 hi/lo A stage
    const int t = threadIdx.x;                               // 0..511
    const float *dv = p.dV + (long long)b * p.Nq * p.N;
    const long long N = p.N;                                  // row pitch of dV
    // which == 0: tile element (r, k) = dV[row0 + r][k0 + k]: chunk id = t + 512 i -> row id >> 3, chunk id & 7 (4 rows x 256 B per warp)
    // which == 1: tile element (r, k) = dV[k0 + k][row0 + r]: chunk id -> row id & 127, chunk id >> 7 (lanes along m: coalesced)
    int r_[2], c_[2];
    const float *src[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int id = t + 512 * i;
      if (which == 0) {
        r_[i] = id >> 3, c_[i] = id & 7;
        src[i] = dv + (long long)(row0 + r_[i]) * N + c_[i] * 8;          // + k0
      } else {
        r_[i] = id & 127, c_[i] = id >> 7;
        src[i] = dv + (long long)(c_[i] * 8) * N + row0 + r_[i];          // + k0 * N, element kk at + kk * N
      }
    }
    float x[2][2][8];                                        // [register set][chunk][8]
    auto load_tile = [&](int kb, float (&d)[2][8]) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (which == 0) {
          const float4 *q = reinterpret_cast<const float4 *>(src[i] + (long long)kb * VB_BK);
          const float4 u = __ldcs(q), v = __ldcs(q + 1);   // dV is read exactly once per gradient: evict-first
          d[i][0] = u.x, d[i][1] = u.y, d[i][2] = u.z, d[i][3] = u.w, d[i][4] = v.x, d[i][5] = v.y, d[i][6] = v.z, d[i][7] = v.w;
        } else {
          const float *q = src[i] + (long long)kb * VB_BK * N;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) d[i][kk] = __ldcs(q + kk * N);
        }
      }
    };
    load_tile(0, x[0]);
#pragma unroll 1
    for (int kb = 0; kb < kblocks; kb += 2) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {                          // two k-blocks per trip: the register sets alternate statically
        const int k = kb + u;
        if (k < kblocks) {
          if (k + 1 < kblocks) load_tile(k + 1, x[u ^ 1]);   // next tile's loads are in flight while this one is converted
          const int s = k % VB_ASTAGES;
          mbar_wait(bar_aempty + 8 * s, ((uint32_t)(k / VB_ASTAGES) & 1) ^ 1);
          uint8_t *stage = as + s * 2 * VB_APLANE;
          store_chunk(stage, r_[0], c_[0], x[u][0]);
          store_chunk(stage, r_[1], c_[1], x[u][1]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_afull + 8 * s);
        }
      }
    }
    // ================================================================= epilogue: warps 0..3, TMEM lane = output position
    if (warp < 4) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
      const int r = warp * 32 + lane;
      float *o = out + (long long)b * VB_C * N + (which ? 0 : p.q_begin) + row0 + r;      // + c * N: a warp writes 32 consecutive positions of channel c
      const bool acc2 = which && p.accumulate2;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
      for (int cb = 0; cb < VB_C / 32; ++cb) {
        uint32_t v[32];
        tmem_ld32(taddr + cb * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float *dst = o + (long long)(cb * 32 + j) * N;
          const float val = __uint_as_float(v[j]) * p.scale;
          *dst = acc2 ? *dst + val : val;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == VB_CONV_WARPS + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)VB_C) : "memory");
  }
}

// fmap fp32 [B, C, N] -> bf16 hi / lo planes of the same layout (K-major for both adjoint GEMMs)
__global__ void __launch_bounds__(256) split_bf16_kernel(const float *__restrict__ x0, const float *__restrict__ x1, __nv_bfloat16 *__restrict__ hi0,
                                                         __nv_bfloat16 *__restrict__ lo0, __nv_bfloat16 *__restrict__ hi1, __nv_bfloat16 *__restrict__ lo1,
                                                         long long n) {
  const float *x = blockIdx.y ? x1 : x0;
  __nv_bfloat16 *hi = blockIdx.y ? hi1 : hi0, *lo = blockIdx.y ? lo1 : lo0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

}  // namespace pf

extern "C" long long pf_volume_bwd_workspace_bytes(int batch, int channels, int h, int w) {
  const long long plane = ((long long)batch * channels * h * w * 2 + 1023) / 1024 * 1024;
  return 4 * plane;
}

extern "C" int pf_volume_bwd(const pf_volume_bwd_args *a, void *stream) {
  using namespace pf;
  PF_REQUIRE(a != nullptr && a->fmap1 && a->fmap2 && a->dvolume && a->workspace, "pf_volume_bwd: null pointer");
  PF_REQUIRE(a->dfmap1 || a->dfmap2, "pf_volume_bwd: nothing to compute");
  const int B = a->batch, C = a->channels, N = a->h * a->w;
  PF_REQUIRE(B > 0 && C == VB_C && N % VB_M == 0, "pf_volume_bwd(tcgen05): built for C = %d and h*w %% %d == 0 (got C = %d, h*w = %d)", VB_C, VB_M, C, N);
  const int Nq = a->query_count > 0 ? a->query_count : N, q_begin = a->query_count > 0 ? a->query_begin : 0;
  PF_REQUIRE(Nq % VB_M == 0 && q_begin % VB_BK == 0 && q_begin >= 0 && q_begin + Nq <= N,
             "pf_volume_bwd: the query chunk must be a multiple of %d rows starting at a multiple of %d inside the map (got [%d, +%d) of %d)", VB_M,
             VB_BK, q_begin, Nq, N);
  PF_REQUIRE(a->workspace_bytes >= pf_volume_bwd_workspace_bytes(B, C, a->h, a->w) && ((uintptr_t)a->workspace & 1023) == 0,
             "pf_volume_bwd: workspace too small or not 1 KiB aligned");
  PF_REQUIRE((((uintptr_t)a->dvolume | (uintptr_t)a->dfmap1 | (uintptr_t)a->dfmap2) & 15) == 0, "pf_volume_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * C * N;
  const long long plane = (n * 2 + 1023) / 1024 * 1024;
  uint8_t *ws = reinterpret_cast<uint8_t *>(a->workspace);
  __nv_bfloat16 *f1_hi = reinterpret_cast<__nv_bfloat16 *>(ws), *f1_lo = reinterpret_cast<__nv_bfloat16 *>(ws + plane);
  __nv_bfloat16 *f2_hi = reinterpret_cast<__nv_bfloat16 *>(ws + 2 * plane), *f2_lo = reinterpret_cast<__nv_bfloat16 *>(ws + 3 * plane);
  if (!a->planes_ready) {   // the bf16 planes of the feature maps: once per backward pass, reused by every chunk
    split_bf16_kernel<<<dim3(592, 2), 256, 0, st>>>(a->fmap1, a->fmap2, f1_hi, f1_lo, f2_hi, f2_lo, n);
    if (int e = check_launch("pf_volume_bwd(split)")) return e;
  }
  CUtensorMap m[4];
  __nv_bfloat16 *planes[4] = {f1_hi, f1_lo, f2_hi, f2_lo};
  for (int i = 0; i < 4; ++i) {
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)B * C};
    cuuint64_t strides[1] = {(cuuint64_t)N * 2};
    cuuint32_t box[2] = {VB_BK, VB_C};
    if (int e = encode(&m[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, planes[i], dims, strides, box, "fmap plane")) return e;
  }
  VolBwdParams p;
  p.B = B, p.N = N, p.Nq = Nq, p.q_begin = q_begin, p.accumulate2 = a->accumulate_dfmap2, p.scale = 1.0f / sqrtf((float)C);
  p.dV = a->dvolume, p.dF1 = a->dfmap1, p.dF2 = a->dfmap2;
  cudaFuncSetAttribute(volume_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_SMEM);   // per device: cheap, every call
  volume_bwd_kernel<<<dim3(N / VB_M, 2, B), VB_THREADS, VB_SMEM, st>>>(m[0], m[1], m[2], m[3], p);
  return check_launch("pf_volume_bwd");
}
