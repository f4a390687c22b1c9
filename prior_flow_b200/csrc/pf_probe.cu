// Measurement aids (bench.py, scripts/): what this GPU sustains for the lookup's ACCESS PATTERN, measured beside the kernel.
// A DCCL lookup reads, per query pixel and level, a 10x10 footprint of that query's PRIVATE plane ([N][h2][w2] fp32, the
// reference's pyramid layout, core/corr.py:99-111): ten 40-byte row segments w2*4 bytes apart, each in a different DRAM page
// from its neighbours'.  `pf_probe_gather` issues exactly those loads (three queries per warp, one coalesced 40-byte segment
// per query and row, evict-first) and nothing else — no coordinate chains, no blends, no stores — so its time is the floor
// of any kernel that has to fetch these footprints from HBM; `pf_probe_stream_read` is the read-only streaming ceiling.
#include "pf_common.cuh"

namespace pf {

__global__ void __launch_bounds__(256) probe_gather_kernel(const float *__restrict__ vol, const int *__restrict__ pos_xy, float *sink,
                                                           long long planes, int H, int W) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int l10 = lane / 10, qq = l10 < 2 ? l10 : 2, a = lane - 10 * l10;
  float acc = 0.f;
  for (long long t = warp; 3 * t < planes; t += nwarps) {
    const long long n = min(3 * t + qq, planes - 1);
    const int x = __ldg(pos_xy + 2 * n), y = __ldg(pos_xy + 2 * n + 1);
    const float *pl = vol + n * H * W + (long long)y * W + x + min(a, 9);
    float v[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) v[r] = __ldcs(pl + r * W);
#pragma unroll
    for (int r = 0; r < 10; ++r) acc += v[r];
  }
  if (acc == 123.456f) sink[0] = acc;   // never true for the probe's data: keeps the loads alive
}

__global__ void __launch_bounds__(256) probe_stream_kernel(const float4 *__restrict__ p, float *sink, long long n4) {
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(p + i);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 123.456f) sink[0] = acc;
}

}  // namespace pf

extern "C" int pf_probe_gather(const float *vol, long long planes, int H, int W, const int *pos_xy, float *sink, void *stream) {
  using namespace pf;
  PF_REQUIRE(vol && pos_xy && sink && planes > 0 && H >= 10 && W >= 10, "pf_probe_gather: bad arguments");
  probe_gather_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(vol, pos_xy, sink, planes, H, W);
  return check_launch("pf_probe_gather");
}

extern "C" int pf_probe_stream_read(const float *src, long long count, float *sink, void *stream) {
  using namespace pf;
  PF_REQUIRE(src && sink && count > 0 && count % 4 == 0 && ((uintptr_t)src & 15) == 0, "pf_probe_stream_read: bad arguments");
  probe_stream_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(src), sink, count / 4);
  return check_launch("pf_probe_stream_read");
}
