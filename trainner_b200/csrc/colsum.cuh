// Column (per-channel) sums of an NHWC bf16 slice -> fp32 bias gradients; HBM-bound, so 16-byte loads,
// four of them in flight per thread.  Shared by conv_wgrad.cu (one conv per launch) and wgrad_rdb.cu
// (one launch for all dense-block convs).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace b200 {

// blockDim.x = 256; `red` is shared memory for 256 * 8 floats.  Requires c % 8 == 0, c <= 2048,
// pitch % 8 == 0, coff % 8 == 0 (16-byte aligned vectors).  Adds scale * sum into dst[0..c).
// Deterministic: every block of the gridDim.x blocks that share `dst` parks its c partial sums in
// part[c][gridDim.x] (+ c * ceil(gridDim.x / 16) floats of group sums behind it); two-level ordered sum
// (common.cuh: det_reduce; `counter` = 1 + ceil(gridDim.x / 16) words).
__device__ __forceinline__ void colsum_vec(const __nv_bfloat16* __restrict__ src, long long npix, int pitch,
                                           int coff, int c, float scale, float* __restrict__ dst,
                                           float* __restrict__ red, float* __restrict__ part,
                                           unsigned* counter) {
  const int vl = c >> 3;            // 8-channel vector lanes per pixel
  const int ppb = 256 / vl;         // pixel lanes per block
  const int tid = threadIdx.x;
  const int v = tid % vl, pl = tid / vl;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (pl < ppb) {
    const long long stride = (long long)gridDim.x * ppb;
    const __nv_bfloat16* base = src + coff + v * 8;
    long long pix = (long long)blockIdx.x * ppb + pl;
    for (; pix + 3 * stride < npix; pix += 4 * stride) {
      uint4 u[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) u[q] = *reinterpret_cast<const uint4*>(base + (pix + q * stride) * pitch);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u[q]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __bfloat1622float2(h[j]);
          s[2 * j] += t.x;
          s[2 * j + 1] += t.y;
        }
      }
    }
    for (; pix < npix; pix += stride) {
      const uint4 u = *reinterpret_cast<const uint4*>(base + pix * pitch);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __bfloat1622float2(h[j]);
        s[2 * j] += t.x;
        s[2 * j + 1] += t.y;
      }
    }
  }
  __syncthreads();   // `red` may still be read by the previous entry's reduction
#pragma unroll
  for (int j = 0; j < 8; ++j) red[j * 256 + tid] = s[j];
  __syncthreads();
  for (int ch = tid; ch < c; ch += 256) {
    const int lane = ch >> 3, j = ch & 7;
    float tot = 0.f;
    for (int p = 0; p < ppb; ++p) tot += red[j * 256 + p * vl + lane];
    part[(size_t)ch * gridDim.x + blockIdx.x] = tot;
  }
  det_reduce(part, part + (size_t)c * gridDim.x, counter, gridDim.x, c, [&](int ch, float tot) { dst[ch] += scale * tot; });
}

}  // namespace b200
