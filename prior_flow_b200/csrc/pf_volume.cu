// (a) All-pairs correlation volume + pyramid: dispatch, and the CUDA-core (exact fp32 FFMA)
// variant used as the on-device checker of the tcgen05 kernel in pf_volume_tc.cu.
// Replaces PriOr_RAFT.corr (core/prior_raft.py:69-75) and DCCL.build_pyramid (core/corr.py:99-111).
#include "pf_common.cuh"

namespace pf {

int volume_build_tc(const pf_volume_args *a, cudaStream_t st);            // pf_volume_tc.cu
long long volume_tc_workspace_bytes(int batch, int channels, int h, int w, int mode);

// V[b, n, m] = scale * sum_c A[b, c, n] * B[b, c, m].  Both operands are "MN-contiguous", the
// natural layout for a register-blocked outer product: 128x128 tile, BK = 8, 8x8 per thread.
constexpr int kSimtTile = 128, kSimtBK = 8;

template <bool kVec>   // kVec: N % 4 == 0, float4 global accesses; otherwise element-wise with bounds checks
__global__ void __launch_bounds__(256) volume_simt_kernel(const float *__restrict__ A, const float *__restrict__ Bm,
                                                          float *__restrict__ V, int C, int N, float scale) {
  __shared__ __align__(16) float As[kSimtBK][kSimtTile];
  __shared__ __align__(16) float Bs[kSimtBK][kSimtTile];
  const int b = blockIdx.z;
  const int n0 = blockIdx.y * kSimtTile, m0 = blockIdx.x * kSimtTile;
  const float *Ab = A + (long long)b * C * N;
  const float *Bb = Bm + (long long)b * C * N;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads; thread owns rows ty*4+{0..3}, 64+ty*4+{0..3}
  const int lr = tid >> 5, lc = (tid & 31) * 4;  // loader: row lr of the k-slab, 4 consecutive columns
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < C; k0 += kSimtBK) {
    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
    if (k0 + lr < C) {
      const float *pa = Ab + (long long)(k0 + lr) * N + n0 + lc, *pb = Bb + (long long)(k0 + lr) * N + m0 + lc;
      if (kVec) {
        if (n0 + lc < N) va = *reinterpret_cast<const float4 *>(pa);
        if (m0 + lc < N) vb = *reinterpret_cast<const float4 *>(pb);
      } else {
        float ta[4] = {0.f, 0.f, 0.f, 0.f}, tb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n0 + lc + j < N) ta[j] = pa[j];
          if (m0 + lc + j < N) tb[j] = pb[j];
        }
        va = make_float4(ta[0], ta[1], ta[2], ta[3]);
        vb = make_float4(tb[0], tb[1], tb[2], tb[3]);
      }
    }
    *reinterpret_cast<float4 *>(&As[lr][lc]) = va;
    *reinterpret_cast<float4 *>(&Bs[lr][lc]) = vb;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kSimtBK; ++kk) {
      float a[8], bv[8];
      *reinterpret_cast<float4 *>(&a[0]) = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
      *reinterpret_cast<float4 *>(&a[4]) = *reinterpret_cast<const float4 *>(&As[kk][64 + ty * 4]);
      *reinterpret_cast<float4 *>(&bv[0]) = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
      *reinterpret_cast<float4 *>(&bv[4]) = *reinterpret_cast<const float4 *>(&Bs[kk][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float *Vb = V + (long long)b * N * N;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = n0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (n >= N) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int m = m0 + jh * 64 + tx * 4;
      if (kVec) {
        if (m < N) {
          float4 o = make_float4(acc[i][jh * 4 + 0] * scale, acc[i][jh * 4 + 1] * scale, acc[i][jh * 4 + 2] * scale,
                                 acc[i][jh * 4 + 3] * scale);
          *reinterpret_cast<float4 *>(Vb + (long long)n * N + m) = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (m + j < N) Vb[(long long)n * N + m + j] = acc[i][jh * 4 + j] * scale;
      }
    }
  }
}

static int volume_build_simt(const pf_volume_args *a, cudaStream_t st) {
  const int N = a->h * a->w;
  dim3 grid(ceil_div(N, kSimtTile), ceil_div(N, kSimtTile), a->batch);
  const float scale = 1.0f / sqrtf((float)a->channels);
  const bool vec = N % 4 == 0 && (((uintptr_t)a->fmap1 | (uintptr_t)a->fmap2 | (uintptr_t)a->level[0]) & 15) == 0;
  if (vec)
    volume_simt_kernel<true><<<grid, 256, 0, st>>>(a->fmap1, a->fmap2, a->level[0], a->channels, N, scale);
  else
    volume_simt_kernel<false><<<grid, 256, 0, st>>>(a->fmap1, a->fmap2, a->level[0], a->channels, N, scale);
  if (int e = check_launch("pf_volume_build(simt)")) return e;
  for (int l = 1; l < a->num_levels; ++l) {
    if (int e = pf_avg_pool2x2(a->level[l - 1], a->level[l], (long long)a->batch * N, a->h >> (l - 1), a->w >> (l - 1),
                               (void *)st))
      return e;
  }
  return 0;
}

}  // namespace pf

extern "C" {

long long pf_volume_workspace_bytes(int batch, int channels, int h, int w, int mode) {
  if (mode == PF_VOL_FP32_SIMT) return 0;
  return pf::volume_tc_workspace_bytes(batch, channels, h, w, mode);
}

int pf_volume_build(const pf_volume_args *a, void *stream) {
  using namespace pf;
  PF_REQUIRE(a && a->fmap1 && a->fmap2, "pf_volume_build: null pointer");
  PF_REQUIRE(a->batch > 0 && a->channels > 0 && a->h > 0 && a->w > 0, "pf_volume_build: bad shape");
  PF_REQUIRE(a->num_levels >= 1 && a->num_levels <= PF_MAX_LEVELS, "pf_volume_build: num_levels must be 1..%d",
             PF_MAX_LEVELS);
  for (int l = 0; l < a->num_levels; ++l) {
    PF_REQUIRE(a->level[l] != nullptr, "pf_volume_build: level[%d] is null", l);
    PF_REQUIRE((a->h >> l) >= 1 && (a->w >> l) >= 1, "pf_volume_build: level %d is empty", l);
  }
  cudaStream_t st = (cudaStream_t)stream;
  switch (a->mode) {
    case PF_VOL_FP32_SIMT:
      return volume_build_simt(a, st);
    case PF_VOL_FP32_3XF16:
    case PF_VOL_F16:
      return volume_build_tc(a, st);
    default:
      set_error("pf_volume_build: unknown mode %d", a->mode);
      return 1;
  }
}

}  // extern "C"
