// Probe (GPU box): what HBM sustains for the volume build's WRITE pattern.  The level-0 volume is [N planes][64][128] fp32 (32 KB per
// plane); a tile of the GEMM writes, for 128 planes, eight 128-byte lines 512 B apart — the probe writes the same 256 MiB
//   mode 0: streaming (consecutive 128-byte lines)
//   mode 1: line i goes to plane (i % N), line (i / N): consecutive warps hit consecutive PLANES (32 KB apart)  [the GEMM's order in time]
//   mode 2: like 1 but 512-byte runs (four lines of one plane row together)
//   mode 3: like 1 but 4 KB runs (a whole 8-row x 128-float strip of one plane)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_write(float4 *out, long long lines, int N, int run_lines, float v, int pad_lines) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long lines_per_plane = lines / N;
  // all quantities are powers of two (N = 8192, run_lines in {1, 4, 32}): shifts and masks, no 64-bit division in the loop
  const int run_shift = run_lines > 0 ? 31 - __clz(run_lines) : 0, n_shift = 31 - __clz(N);
  for (long long i = warp; i < lines; i += nwarps) {
    long long dst = i;
    if (run_lines > 0) {
      const unsigned run = (unsigned)(i >> run_shift), within = (unsigned)i & (run_lines - 1);
      const unsigned plane = run & (N - 1), slot = run >> n_shift;        // consecutive runs -> consecutive planes
      dst = (long long)plane * (lines_per_plane + pad_lines) + (slot << run_shift) + within;   // pad_lines: skew the 32 KB plane pitch
    }
    if (lane < 8) out[dst * 8 + lane] = make_float4(v, v, v, v);        // 8 x 16 B = one 128-byte line
  }
}
int main() {
  const int N = 8192;
  const long long lines = (long long)N * 64 * 128 * 4 / 128;            // 2 Mi lines = 256 MiB
  float4 *buf[4];
  for (int s = 0; s < 4; ++s) cudaMalloc(&buf[s], (lines + 64LL * N) * 128);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const int runs[4] = {0, 1, 4, 32};
  const int pads[6] = {0, 1, 2, 3, 8, 9};
  for (int pi = 0; pi < 6; ++pi)
  for (int mode = 0; mode < 4; ++mode) {
    const int pad = pads[pi];
    if (mode == 0 && pi > 0) continue;
    float best = 1e9f;
    for (int it = 0; it < 12; ++it) {
      cudaEventRecord(e0);
      k_write<<<148 * 8, 256>>>(buf[it % 4], lines, N, runs[mode], (float)it, pad);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 2 && ms < best) best = ms;
    }
    printf("pad %d lines | mode %d (run = %d lines): %.1f us for 256 MiB = %.2f TB/s  err=%d\n", pad, mode, runs[mode], best * 1e3, lines * 128.0 / (best * 1e-3) / 1e12,
           (int)cudaGetLastError());
  }
  return 0;
}
