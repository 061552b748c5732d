// 3x3 stride-1 pad-1 convolutions with a thin (<= 4 channel, image) side: CUDA-core direct kernels.
// K = 27 (or N = 3) is a degenerate GEMM for tcgen05, and these layers are HBM/LSU bound:
//   RRDBNet fea_conv 3->64 and HR_conv1 64->3 (RRDBNet_arch.py:23,40), Discriminator_VGG conv0
//   (discriminators.py:21), VGG19 conv1_1 with the ImageNet input normalisation folded in
//   (perceptual.py:207,210) -- forward, input gradient and weight gradient.
// The thin side is NCHW fp32 (the layout of images at the nn.Module boundary), the wide side is
// NHWC bf16 (the internal activation layout).
#include "common.cuh"

namespace b200 {
namespace {

constexpr int kMaxThin = 4;

__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// ---------------------------------------------------------------- thin -> wide
// block (32 x 8 pixels); dynamic smem: weights [cs*9][cw] fp32
template <int CS>
__global__ void __launch_bounds__(256)
thin_to_wide_kernel(const float* __restrict__ x, const float* __restrict__ wgt,
                    const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int n, int h,
                    int w, int cw, int cy, int y_coff, int transpose_w,
                    const float* __restrict__ mean, const float* __restrict__ stdv, int act,
                    float slope, const __nv_bfloat16* __restrict__ mask, int mask_c, int mask_coff,
                    float mask_slope) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float wsm[];  // [CS*9][cw]
  __shared__ float patch[CS][10][34];
  const int tid = threadIdx.x;
  for (int i = tid; i < CS * 9 * cw; i += 256) {
    const int j = i % cw;        // wide channel
    const int it = i / cw;       // cs_i * 9 + t'
    const int ci = it / 9, t = it % 9;
    wsm[i] = transpose_w ? wgt[((size_t)ci * cw + j) * 9 + (8 - t)] : wgt[((size_t)j * CS + ci) * 9 + t];
  }
  const int tiles_x = (w + 31) / 32, tiles_y = (h + 7) / 8;
  const int b = blockIdx.x / (tiles_x * tiles_y);
  const int tr = blockIdx.x % (tiles_x * tiles_y);
  const int x0 = (tr % tiles_x) * 32, y0 = (tr / tiles_x) * 8;
  for (int i = tid; i < CS * 10 * 34; i += 256) {
    const int px = i % 34, py = (i / 34) % 10, ci = i / 340;
    const int gx = x0 + px - 1, gy = y0 + py - 1;
    float v = 0.f;
    if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
      v = x[(((size_t)b * CS + ci) * h + gy) * w + gx];
      if (mean) v = (v - mean[ci]) / stdv[ci];
    }
    patch[ci][py][px] = v;
  }
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5;
  const int gx = x0 + tx, gy = y0 + ty;
  if (gx >= w || gy >= h) return;
  float pv[CS * 9];
#pragma unroll
  for (int ci = 0; ci < CS; ++ci)
#pragma unroll
    for (int t = 0; t < 9; ++t) pv[ci * 9 + t] = patch[ci][ty + t / 3][tx + t % 3];
  __nv_bfloat16* dst = y + (((size_t)b * h + gy) * w + gx) * cy + y_coff;
  for (int g = 0; g < cw; g += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias ? bias[g + j] : 0.f;
#pragma unroll
    for (int k = 0; k < CS * 9; ++k) {
      const float4 w0 = *reinterpret_cast<const float4*>(&wsm[k * cw + g]);
      const float4 w1 = *reinterpret_cast<const float4*>(&wsm[k * cw + g + 4]);
      const float v = pv[k];
      acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]);
      acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
      acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
      acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
    }
    if (act) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = acc[j] > 0.f ? acc[j] : acc[j] * slope;
    }
    if (mask) {
      float mv[8];
      unpack8(*reinterpret_cast<const uint4*>(mask + (((size_t)b * h + gy) * w + gx) * mask_c +
                                              mask_coff + g),
              mv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = mv[j] > 0.f ? acc[j] : acc[j] * mask_slope;
    }
    *reinterpret_cast<uint4*>(dst + g) = pack8(acc);
  }
}

// ---------------------------------------------------------------- wide -> thin
// block 16x16 pixels; dynamic smem: tile [18*18][cw] bf16 (16-byte chunks XOR-swizzled by pixel) +
// weights [9][cw][4] fp32
__global__ void __launch_bounds__(256)
wide_to_thin_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wgt,
                    const float* __restrict__ bias, float* __restrict__ y, int n, int h, int w, int cw,
                    int cx, int x_coff, int cs, int transpose_w, const float* __restrict__ inv_std,
                    float out_scale) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t dsm[];
  const int chunks = cw / 8;
  uint4* tile = reinterpret_cast<uint4*>(dsm);                       // [324][chunks]
  float4* wsm = reinterpret_cast<float4*>(dsm + (size_t)324 * chunks * 16);  // [9][cw]
  const int tid = threadIdx.x;
  for (int i = tid; i < 9 * cw; i += 256) {
    const int j = i % cw, t = i / cw;
    float wv[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ci = 0; ci < cs; ++ci)
      wv[ci] = transpose_w ? wgt[((size_t)j * cs + ci) * 9 + (8 - t)] : wgt[((size_t)ci * cw + j) * 9 + t];
    wsm[i] = make_float4(wv[0], wv[1], wv[2], wv[3]);
  }
  const int tiles_x = (w + 15) / 16, tiles_y = (h + 15) / 16;
  const int b = blockIdx.x / (tiles_x * tiles_y);
  const int tr = blockIdx.x % (tiles_x * tiles_y);
  const int x0 = (tr % tiles_x) * 16, y0 = (tr / tiles_x) * 16;
  const int swz_mask = chunks >= 8 ? 7 : (chunks - 1);
  for (int i = tid; i < 324 * chunks; i += 256) {
    const int ch = i % chunks, pix = i / chunks;
    const int px = pix % 18, py = pix / 18;
    const int gx = x0 + px - 1, gy = y0 + py - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (gx >= 0 && gx < w && gy >= 0 && gy < h)
      v = *reinterpret_cast<const uint4*>(x + (((size_t)b * h + gy) * w + gx) * cx + x_coff + ch * 8);
    tile[pix * chunks + (ch ^ (px & swz_mask))] = v;
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  const int gx = x0 + tx, gy = y0 + ty;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int t = 0; t < 9; ++t) {
    const int px = tx + t % 3, py = ty + t / 3;
    const int pix = py * 18 + px;
    for (int ch = 0; ch < chunks; ++ch) {
      float f[8];
      unpack8(tile[pix * chunks + (ch ^ (px & swz_mask))], f);
      const float4* wp = wsm + t * cw + ch * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 wv = wp[k];
        acc[0] = fmaf(f[k], wv.x, acc[0]);
        acc[1] = fmaf(f[k], wv.y, acc[1]);
        acc[2] = fmaf(f[k], wv.z, acc[2]);
        acc[3] = fmaf(f[k], wv.w, acc[3]);
      }
    }
  }
  if (gx >= w || gy >= h) return;
  for (int co = 0; co < cs; ++co) {
    float v = acc[co] + (bias ? bias[co] : 0.f);
    if (inv_std) v *= inv_std[co];
    y[(((size_t)b * cs + co) * h + gy) * w + gx] = v * out_scale;
  }
}

// ---------------------------------------------------------------- weight gradient
// acc[j (wide)][i (thin)][t] = sum_p wide[p, j] * thin[p + sign * off_t, i]
// block = 64 wide-channel lanes x 4 pixel subsets; one image row per iteration.
template <int CS>
__global__ void __launch_bounds__(256)
thin_wgrad_kernel(const float* __restrict__ thin, const __nv_bfloat16* __restrict__ wide,
                  float* __restrict__ dw, float* __restrict__ dbias_wide, int n, int h, int w,
                  int cw, int cwide_buf, int wide_coff, int wide_is_out, float* __restrict__ part,
                  unsigned* __restrict__ counters) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float rows[];  // [CS][3][w + 2]
  __shared__ float red[4][64];
  const int tid = threadIdx.x;
  const int lane_c = tid & 63, sub = tid >> 6;
  const int j = blockIdx.y * 64 + lane_c;
  const int sign = wide_is_out ? 1 : -1;
  const int wp = w + 2;
  float acc[CS * 9];
#pragma unroll
  for (int k = 0; k < CS * 9; ++k) acc[k] = 0.f;
  float bsum = 0.f;
  const int total_rows = n * h;
  for (int row = blockIdx.x; row < total_rows; row += gridDim.x) {
    const int b = row / h, yy = row % h;
    __syncthreads();
    for (int i = tid; i < CS * 3 * wp; i += 256) {
      const int px = i % wp, r = (i / wp) % 3, ci = i / (3 * wp);
      const int gy = yy + (r - 1), gx = px - 1;
      float v = 0.f;
      if (gx >= 0 && gx < w && gy >= 0 && gy < h) v = thin[(((size_t)b * CS + ci) * h + gy) * w + gx];
      rows[i] = v;
    }
    __syncthreads();
    if (j < cw) {
      const __nv_bfloat16* wrow = wide + ((size_t)b * h + yy) * w * cwide_buf + wide_coff + j;
      // each pixel subset walks a contiguous quarter of the row with a 3-wide sliding window of the thin
      // values per (channel, kernel row): 9 shared-memory loads (warp broadcast) feed 27 FMAs per pixel
      const int seg = (w + 3) / 4;
      const int xa = sub * seg, xb = min(xa + seg, w);
      if (xa < xb) {
        float win[CS * 3][3];
#pragma unroll
        for (int q = 0; q < CS * 3; ++q) {
          const int ci = q / 3, ky = q % 3;
          const int r = 1 + sign * (ky - 1);
          const float* rp = rows + (ci * 3 + r) * wp;
          win[q][0] = (xa >= 1) ? rp[xa - 1 + 1] : rp[0];   // column xa-1 (index +1 for the left halo)
          win[q][1] = rp[xa + 1];
          win[q][2] = 0.f;
        }
        for (int xx = xa; xx < xb; ++xx) {
          const float v = __bfloat162float(wrow[(size_t)xx * cwide_buf]);
          bsum += v;
#pragma unroll
          for (int q = 0; q < CS * 3; ++q) {
            const int ci = q / 3, ky = q % 3;
            const int r = 1 + sign * (ky - 1);
            win[q][2] = rows[(ci * 3 + r) * wp + xx + 2];   // column xx+1
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const float tv = (sign > 0) ? win[q][kx] : win[q][2 - kx];   // column xx + sign*(kx-1)
              acc[ci * 9 + ky * 3 + kx] = fmaf(v, tv, acc[ci * 9 + ky * 3 + kx]);
            }
            win[q][0] = win[q][1];
            win[q][1] = win[q][2];
          }
        }
      }
    }
  }
  // reduce the 4 pixel subsets into this block's partial row [64 lanes][CS*9 + 1]; the last block of the
  // channel group (blockIdx.y) to arrive adds the rows of all blocks in block order (deterministic)
  constexpr int NC1 = CS * 9 + 1;
  float* mine = part + (size_t)blockIdx.y * (gridDim.x + det_groups(gridDim.x)) * 64 * NC1;   // output-major [64 * NC1][gridDim.x] + group sums
#pragma unroll
  for (int k = 0; k < NC1; ++k) {
    __syncthreads();
    red[sub][lane_c] = (k < CS * 9) ? acc[k] : bsum;
    __syncthreads();
    if (sub == 0)
      mine[(size_t)(lane_c * NC1 + k) * gridDim.x + blockIdx.x] =
          red[0][lane_c] + red[1][lane_c] + red[2][lane_c] + red[3][lane_c];
  }
  det_reduce(mine, mine + (size_t)64 * NC1 * gridDim.x, counters + blockIdx.y * (1 + det_groups(gridDim.x)), gridDim.x,
             64 * NC1, [&](int i, float sv) {
    const int m = i / NC1, k = i % NC1;
    const int jj = blockIdx.y * 64 + m;
    if (jj >= cw) return;
    if (k < CS * 9) {
      const int ci = k / 9, t = k % 9;
      const size_t idx = wide_is_out ? ((size_t)jj * CS + ci) * 9 + t : ((size_t)ci * cw + jj) * 9 + t;
      dw[idx] += sv;
    } else if (dbias_wide) {
      dbias_wide[jj] += sv;
    }
  });
}

// ================================================================= tensor-core (mma.sync) versions
// The thin convolutions are HBM-bound layers with a degenerate GEMM shape (K = 27 or N = 3): the
// CUDA-core kernels above are FMA/LDS-bound at 6-20x their memory roofline.  The kernels below do the
// same arithmetic as warp-level m16n8k16 bf16 MMAs with fp32 accumulation (the thin operand is rounded
// to bf16, which is what the reference's autocast path does to these conv inputs too); tcgen05 buys
// nothing here because the tensor work is < 5 % of the memory time.
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- thin -> wide, GEMM [px] x [27->16*KS] x [64]
// block = 32 x 8 pixels, warp = one row of 32 pixels = two m16 tiles.  Output channel of (n-tile j, column c)
// is 16*(c>>1) + 2*j + (c&1): the accumulator fragment of lane (g, t) is then 16 consecutive channels of
// pixel rows g and g+8, stored as two 16-byte vectors each.
template <int CS>
__global__ void __launch_bounds__(256)
thin_to_wide_mma_kernel(const float* __restrict__ x, const float* __restrict__ wgt,
                        const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int n, int h,
                        int w, int cw, int cy, int y_coff, int transpose_w,
                        const float* __restrict__ mean, const float* __restrict__ stdv, int act,
                        float slope, const __nv_bfloat16* __restrict__ mask, int mask_c, int mask_coff,
                        float mask_slope) {
  pdl_trigger();
  pdl_wait();
  constexpr int KS = (CS * 9 + 15) / 16;
  extern __shared__ uint2 bfrag[];  // [cw/64][KS][8][32]
  __shared__ float patch[CS * 10 * 34];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int groups = cw / 64;
  for (int i = tid; i < groups * KS * 256; i += 256) {
    const int l = i & 31, j = (i >> 5) & 7, ks = (i >> 8) % KS, grp = i / (256 * KS);
    const int gg = l >> 2, tt = l & 3;
    const int ch = grp * 64 + 16 * (gg >> 1) + 2 * j + (gg & 1);
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = 16 * ks + 2 * tt + (q & 1) + (q >> 1) * 8;
      float r = 0.f;
      if (k < CS * 9) {
        const int ci = k / 9, tp = k % 9;
        r = transpose_w ? wgt[((size_t)ci * cw + ch) * 9 + (8 - tp)] : wgt[((size_t)ch * CS + ci) * 9 + tp];
      }
      v[q] = r;
    }
    bfrag[i] = make_uint2(pack2(v[0], v[1]), pack2(v[2], v[3]));
  }
  const int tiles_x = (w + 31) / 32, tiles_y = (h + 7) / 8;
  const int b = blockIdx.x / (tiles_x * tiles_y);
  const int tr = blockIdx.x % (tiles_x * tiles_y);
  const int x0 = (tr % tiles_x) * 32, y0 = (tr / tiles_x) * 8;
  for (int i = tid; i < CS * 10 * 34; i += 256) {
    const int px = i % 34, py = (i / 34) % 10, ci = i / 340;
    const int gx = x0 + px - 1, gy = y0 + py - 1;
    float v = 0.f;
    if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
      v = x[(((size_t)b * CS + ci) * h + gy) * w + gx];
      if (mean) v = (v - mean[ci]) / stdv[ci];
    }
    patch[i] = v;
  }
  __syncthreads();
  int off[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = 16 * ks + 2 * t + (q & 1) + (q >> 1) * 8;
      off[ks][q] = k < CS * 9 ? (k / 9) * 340 + ((k % 9) / 3) * 34 + (k % 9) % 3 : -1;
    }
  const int gy = y0 + warp;
#pragma unroll 1
  for (int mt = 0; mt < 2; ++mt) {
    const int xb = mt * 16;
    const int base = warp * 34 + xb + g;
    uint32_t a[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      float v[4][2];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[q][0] = off[ks][q] >= 0 ? patch[base + off[ks][q]] : 0.f;
        v[q][1] = off[ks][q] >= 0 ? patch[base + 8 + off[ks][q]] : 0.f;
      }
      a[ks][0] = pack2(v[0][0], v[1][0]);
      a[ks][1] = pack2(v[0][1], v[1][1]);
      a[ks][2] = pack2(v[2][0], v[3][0]);
      a[ks][3] = pack2(v[2][1], v[3][1]);
    }
    for (int grp = 0; grp < groups; ++grp) {
      float acc[8][4];
      const int cbase = grp * 64 + 16 * t;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float b0 = bias ? bias[cbase + 2 * j] : 0.f, b1 = bias ? bias[cbase + 2 * j + 1] : 0.f;
        acc[j][0] = b0; acc[j][1] = b1; acc[j][2] = b0; acc[j][3] = b1;
      }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint2 bb = bfrag[((grp * KS + ks) * 8 + j) * 32 + lane];
          mma_16816(acc[j], a[ks], bb.x, bb.y);
        }
#pragma unroll
      for (int rs = 0; rs < 2; ++rs) {
        const int gx = x0 + xb + g + rs * 8;
        if (gx >= w || gy >= h) continue;
        const size_t pix = ((size_t)b * h + gy) * w + gx;
        float v[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[2 * j] = acc[j][rs * 2];
          v[2 * j + 1] = acc[j][rs * 2 + 1];
        }
        if (act) {
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = v[q] > 0.f ? v[q] : v[q] * slope;
        }
        if (mask) {
          float mv[16];
          const uint4* mp = reinterpret_cast<const uint4*>(mask + pix * mask_c + mask_coff + cbase);
          unpack8(mp[0], mv);
          unpack8(mp[1], mv + 8);
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = mv[q] > 0.f ? v[q] : v[q] * mask_slope;
        }
        uint4* dst = reinterpret_cast<uint4*>(y + pix * cy + y_coff + cbase);
        dst[0] = pack8(v);
        dst[1] = pack8(v + 8);
      }
    }
  }
}

// ---------------------------------------------------------------- wide -> thin, GEMM [px] x [9*cw] x [8]
// block = 16 x 16 pixels, warp = two rows = two m16 tiles sharing every weight fragment.
__global__ void __launch_bounds__(256)
wide_to_thin_mma_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wgt,
                        const float* __restrict__ bias, float* __restrict__ y, int n, int h, int w, int cw,
                        int cx, int x_coff, int cs, int transpose_w, const float* __restrict__ inv_std,
                        float out_scale) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t dsm[];
  const int chunks = cw / 8, kcs = cw / 16;
  uint4* tile = reinterpret_cast<uint4*>(dsm);                                   // [324][chunks], swizzled
  uint2* bfrag = reinterpret_cast<uint2*>(dsm + (size_t)324 * chunks * 16);      // [9][kcs][32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int i = tid; i < 9 * kcs * 32; i += 256) {
    const int l = i & 31, kc = (i >> 5) % kcs, tap = i / (32 * kcs);
    const int gg = l >> 2, tt = l & 3;
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = kc * 16 + 2 * tt + (q & 1) + (q >> 1) * 8;   // wide channel
      float r = 0.f;
      if (gg < cs)
        r = transpose_w ? wgt[((size_t)j * cs + gg) * 9 + (8 - tap)] : wgt[((size_t)gg * cw + j) * 9 + tap];
      v[q] = r;
    }
    bfrag[i] = make_uint2(pack2(v[0], v[1]), pack2(v[2], v[3]));
  }
  // Persistent over 16 x 16 pixel tiles (weight fragments built once per CTA); the haloed input tile arrives through
  // cp.async (zero fill outside the image): all of a thread's ~10 16-byte loads are in flight together instead of one
  // dependent global load per loop iteration, which is what kept this HBM-bound kernel at 20 % of the HBM peak.
  const int tiles_x = (w + 15) / 16, tiles_y = (h + 15) / 16;
  const int total_tiles = n * tiles_x * tiles_y;
  const int swz_mask = chunks >= 8 ? 7 : (chunks - 1);
  const uint32_t tile_s = smem_addr(tile);
  for (int tile_i = blockIdx.x; tile_i < total_tiles; tile_i += gridDim.x) {
  const int b = tile_i / (tiles_x * tiles_y);
  const int tr = tile_i % (tiles_x * tiles_y);
  const int x0 = (tr % tiles_x) * 16, y0 = (tr / tiles_x) * 16;
  __syncthreads();   // previous tile fully consumed (first pass: bfrag complete)
  for (int i = tid; i < 324 * chunks; i += 256) {
    const int ch = i % chunks, pix = i / chunks;
    const int px = pix % 18, py = pix / 18;
    const int gx = x0 + px - 1, gy = y0 + py - 1;
    const bool in = gx >= 0 && gx < w && gy >= 0 && gy < h;
    const __nv_bfloat16* src = in ? x + (((size_t)b * h + gy) * w + gx) * cx + x_coff + ch * 8 : x;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tile_s + (uint32_t)((pix * chunks + (ch ^ (px & swz_mask))) * 16)),
                 "l"(src), "r"(in ? 16 : 0)
                 : "memory");
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  float acc[2][4];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
  const int mi = lane >> 3, lr = lane & 7;
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3, dx = tap % 3;
    const int px = lr + (mi & 1) * 8 + dx;
    const int sw = px & swz_mask;
    const int pix0 = (2 * warp + dy) * 18 + px;
    for (int kc = 0; kc < kcs; ++kc) {
      const int chunk = (kc * 2 + (mi >> 1)) ^ sw;
      uint32_t a0[4], a1[4];
      ldsm_x4(a0, tile_s + (uint32_t)((pix0 * chunks + chunk) * 16));
      ldsm_x4(a1, tile_s + (uint32_t)(((pix0 + 18) * chunks + chunk) * 16));
      const uint2 bb = bfrag[(tap * kcs + kc) * 32 + lane];
      mma_16816(acc[0], a0, bb.x, bb.y);
      mma_16816(acc[1], a1, bb.x, bb.y);
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int gy = y0 + 2 * warp + r;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int co = 2 * t + (q & 1);
      const int gx = x0 + g + (q >> 1) * 8;
      if (co >= cs || gx >= w || gy >= h) continue;
      float v = acc[r][q] + (bias ? bias[co] : 0.f);
      if (inv_std) v *= inv_std[co];
      y[(((size_t)b * cs + co) * h + gy) * w + gx] = v * out_scale;
    }
  }
  }   // tile loop
}

// ---------------------------------------------------------------- weight gradient, GEMM [64] x [px] x [CS*9 + 1]
// A = wide^T (ldmatrix.trans from the [pixel][channel] row tile), B = shifted thin values (+ a column of ones
// that yields the wide-side bias gradient), K = the pixels of one image row per block iteration.
template <int CS>
__global__ void __launch_bounds__(256)
thin_wgrad_mma_kernel(const float* __restrict__ thin, const __nv_bfloat16* __restrict__ wide,
                      float* __restrict__ dw, float* __restrict__ dbias_wide, int n, int h, int w, int w16,
                      int cw, int cwide_buf, int wide_coff, int wide_is_out, float* __restrict__ part,
                      unsigned* __restrict__ counters) {
  pdl_trigger();
  pdl_wait();
  constexpr int NC = CS * 9 + 1;          // used columns (the last one is the ones column)
  constexpr int NT = (NC + 7) / 8;
  extern __shared__ __align__(16) uint8_t dsm[];
  const int wp = w16 + 2;
  uint4* tile = reinterpret_cast<uint4*>(dsm);                               // [w16][8], swizzled
  float* rows = reinterpret_cast<float*>(dsm + (size_t)w16 * 128);          // [CS][3][wp]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int sign = wide_is_out ? 1 : -1;
  const int jbase = blockIdx.y * 64;
  int off[NT];   // >= 0: offset into rows ; -1: ones ; -2: zero
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int nn = j * 8 + g;
    if (nn < CS * 9) {
      const int ci = nn / 9, ky = (nn % 9) / 3, kx = nn % 3;
      off[j] = (ci * 3 + 1 + sign * (ky - 1)) * wp + sign * (kx - 1) + 1;
    } else {
      off[j] = nn == CS * 9 ? -1 : -2;
    }
  }
  float acc[4][NT][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[m][j][q] = 0.f;
  const uint32_t tile_s = smem_addr(tile);
  const int mi = lane >> 3, lr = lane & 7;
  const int total_rows = n * h;
  for (int row = blockIdx.x; row < total_rows; row += gridDim.x) {
    const int b = row / h, yy = row % h;
    __syncthreads();
    for (int i = tid; i < CS * 3 * wp; i += 256) {
      const int px = i % wp, r = (i / wp) % 3, ci = i / (3 * wp);
      const int gy = yy + (r - 1), gx = px - 1;
      float v = 0.f;
      if (gx >= 0 && gx < w && gy >= 0 && gy < h) v = thin[(((size_t)b * CS + ci) * h + gy) * w + gx];
      rows[i] = v;
    }
    const __nv_bfloat16* wrow = wide + ((size_t)b * h + yy) * w * cwide_buf + wide_coff + jbase;
    for (int i = tid; i < w16 * 8; i += 256) {
      const int ch = i & 7, px = i >> 3;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (px < w) v = *reinterpret_cast<const uint4*>(wrow + (size_t)px * cwide_buf + ch * 8);
      tile[px * 8 + (ch ^ (px & 7))] = v;
    }
    __syncthreads();
    for (int p0 = warp * 16; p0 < w16; p0 += 128) {
      uint32_t bq[NT][2];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        if (off[j] >= 0) {
          const float* rp = rows + off[j] + p0 + 2 * t;
          bq[j][0] = pack2(rp[0], rp[1]);
          bq[j][1] = pack2(rp[8], rp[9]);
        } else {
          bq[j][0] = bq[j][1] = off[j] == -1 ? 0x3F803F80u : 0u;
        }
      }
      const int px = p0 + lr + (mi >> 1) * 8;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        uint32_t a[4];
        ldsm_x4_trans(a, tile_s + (uint32_t)((px * 8 + ((m * 2 + (mi & 1)) ^ (px & 7))) * 16));
#pragma unroll
        for (int j = 0; j < NT; ++j) mma_16816(acc[m][j], a, bq[j][0], bq[j][1]);
      }
    }
  }
  // cross-warp reduction in shared memory (the row tile is dead now), then one atomic per output
  __syncthreads();
  float* red = reinterpret_cast<float*>(dsm);   // [64][NT*8 + 1]
  constexpr int RS = NT * 8 + 1;
  for (int i = tid; i < 64 * RS; i += 256) red[i] = 0.f;
  __syncthreads();
  for (int wi = 0; wi < 8; ++wi) {
    if (warp == wi) {
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            red[(m * 16 + g + (q >> 1) * 8) * RS + j * 8 + 2 * t + (q & 1)] += acc[m][j][q];
    }
    __syncthreads();
  }
  // this block's partial [64][NC] -> scratch; the last block of the channel group adds all blocks in block order
  float* mine = part + (size_t)blockIdx.y * (gridDim.x + det_groups(gridDim.x)) * 64 * NC;   // output-major [64 * NC][gridDim.x] + group sums
  for (int i = tid; i < 64 * NC; i += 256) mine[(size_t)i * gridDim.x + blockIdx.x] = red[(i / NC) * RS + (i % NC)];
  det_reduce(mine, mine + (size_t)64 * NC * gridDim.x, counters + blockIdx.y * (1 + det_groups(gridDim.x)), gridDim.x,
             64 * NC, [&](int i, float sv) {
    const int m = i / NC, nn = i % NC;
    const int j = jbase + m;
    if (j >= cw) return;
    if (nn < CS * 9) {
      const int ci = nn / 9, tp = nn % 9;
      const size_t idx = wide_is_out ? ((size_t)j * CS + ci) * 9 + tp : ((size_t)ci * cw + j) * 9 + tp;
      dw[idx] += sv;
    } else if (dbias_wide) {
      dbias_wide[j] += sv;
    }
  });
}

__global__ void plane_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int n, int c,
                                 long long hw, float* __restrict__ part, unsigned* __restrict__ counters) {
  pdl_trigger();
  pdl_wait();
  const int ch = blockIdx.y;
  float s = 0.f;
  for (int b = 0; b < n; ++b) {
    const float* p = x + ((size_t)b * c + ch) * hw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hw;
         i += (long long)gridDim.x * blockDim.x)
      s += p[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    part[ch * gridDim.x + blockIdx.x] = t;
  }
  if (det_arrive_last(counters + ch, gridDim.x) && threadIdx.x == 0) {
    float t = 0.f;
    for (unsigned b = 0; b < gridDim.x; ++b) t += part[ch * gridDim.x + b];
    out[ch] += t;
  }
}

}  // namespace
}  // namespace b200

using namespace b200;
typedef __nv_bfloat16 bf16;

extern "C" {

int b200_conv3x3_thin_to_wide(const float* x, const float* w_oihw, const float* bias, void* y,
                              int32_t n, int32_t h, int32_t w, int32_t cs, int32_t cw, int32_t cy,
                              int32_t y_coff, int32_t transpose_w, const float* mean,
                              const float* stdv, int32_t act, float slope, const void* mask,
                              int32_t mask_c, int32_t mask_coff, float mask_slope,
                              b200_stream_t stream) {
  B200_REQUIRE(cs >= 1 && cs <= kMaxThin, "thin_to_wide: thin channels %d not in 1..4", cs);
  B200_REQUIRE(cw % 8 == 0 && cy % 8 == 0 && y_coff % 8 == 0, "thin_to_wide: wide channels must be multiples of 8");
  const int blocks = n * ((w + 31) / 32) * ((h + 7) / 8);
  if (cw % 64 == 0 && cw <= 256) {
    const size_t fsm = (size_t)(cw / 64) * ((cs * 9 + 15) / 16) * 256 * sizeof(uint2);
#define LAUNCH_TWM(CS)                                                                           \
  ::b200::launch_kernel(thin_to_wide_mma_kernel<CS>, blocks, 256, fsm, as_stream(stream),                           \
      x, w_oihw, bias, (bf16*)y, n, h, w, cw, cy, y_coff, transpose_w, mean, stdv, act, slope, \
      (const bf16*)mask, mask_c, mask_coff, mask_slope)
    switch (cs) {
      case 1: LAUNCH_TWM(1); break;
      case 2: LAUNCH_TWM(2); break;
      case 3: LAUNCH_TWM(3); break;
      default: LAUNCH_TWM(4); break;
    }
#undef LAUNCH_TWM
    B200_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = (size_t)cs * 9 * cw * sizeof(float);
  B200_REQUIRE(smem <= 40 * 1024, "thin_to_wide: cw too large");
#define LAUNCH_TW(CS)                                                                            \
  ::b200::launch_kernel(thin_to_wide_kernel<CS>, blocks, 256, smem, as_stream(stream),                              \
      x, w_oihw, bias, (bf16*)y, n, h, w, cw, cy, y_coff, transpose_w, mean, stdv, act, slope, \
      (const bf16*)mask, mask_c, mask_coff, mask_slope)
  switch (cs) {
    case 1: LAUNCH_TW(1); break;
    case 2: LAUNCH_TW(2); break;
    case 3: LAUNCH_TW(3); break;
    default: LAUNCH_TW(4); break;
  }
#undef LAUNCH_TW
  B200_LAUNCH_CHECK();
  return 0;
}

int b200_conv3x3_wide_to_thin(const void* x, const float* w_oihw, const float* bias, float* y,
                              int32_t n, int32_t h, int32_t w, int32_t cw, int32_t cx, int32_t x_coff,
                              int32_t cs, int32_t transpose_w, const float* inv_std, float out_scale,
                              b200_stream_t stream) {
  B200_REQUIRE(cs >= 1 && cs <= kMaxThin, "wide_to_thin: thin channels %d not in 1..4", cs);
  B200_REQUIRE(cw % 8 == 0 && cx % 8 == 0 && x_coff % 8 == 0 && cw <= 128,
               "wide_to_thin: wide channels must be multiples of 8 and <= 128");
  const int chunks = cw / 8;
  B200_REQUIRE((chunks & (chunks - 1)) == 0, "wide_to_thin: cw/8 must be a power of two");
  const int blocks_m = n * ((w + 15) / 16) * ((h + 15) / 16);
  if (cw % 16 == 0) {
    const size_t msm = (size_t)324 * chunks * 16 + (size_t)9 * (cw / 16) * 32 * sizeof(uint2);
    static size_t msm_set = 0;
    if (msm > 48 * 1024 && msm > msm_set) {
      B200_CHECK_CUDA(cudaFuncSetAttribute(wide_to_thin_mma_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
      msm_set = msm;
    }
    ::b200::launch_kernel(wide_to_thin_mma_kernel, (blocks_m < 4 * sm_count() ? blocks_m : 4 * sm_count()), 256, msm, as_stream(stream), 
        (const bf16*)x, w_oihw, bias, y, n, h, w, cw, cx, x_coff, cs, transpose_w, inv_std, out_scale);
    B200_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = (size_t)324 * chunks * 16 + (size_t)9 * cw * 16;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(wide_to_thin_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const int blocks = n * ((w + 15) / 16) * ((h + 15) / 16);
  ::b200::launch_kernel(wide_to_thin_kernel, blocks, 256, smem, as_stream(stream), 
      (const bf16*)x, w_oihw, bias, y, n, h, w, cw, cx, x_coff, cs, transpose_w, inv_std, out_scale);
  B200_LAUNCH_CHECK();
  return 0;
}

int b200_conv3x3_thin_wgrad(const float* thin, const void* wide, float* dw, float* dbias_wide,
                            float* dbias_thin, int32_t n, int32_t h, int32_t w, int32_t cs,
                            int32_t cw, int32_t cwide_buf, int32_t wide_coff, int32_t wide_is_out,
                            const float* mean, const float* stdv, b200_stream_t stream) {
  (void)mean;
  (void)stdv;
  B200_REQUIRE(cs >= 1 && cs <= kMaxThin, "thin_wgrad: thin channels %d not in 1..4", cs);
  int gx = n * h < 2 * sm_count() ? n * h : 2 * sm_count();
  dim3 grid(gx, (cw + 63) / 64);
  // deterministic cross-block reduction: per-block partials [grid.y][grid.x][64][cs*9+1] in the library scratch;
  // the plane sums (launched after, stream-ordered) reuse the front of the same arena
  DetScratch ds;
  if (det_scratch(&ds, (size_t)(grid.x + det_groups(grid.x)) * grid.y * 64 * (cs * 9 + 1) + 32 * 4,
                  (int)grid.y * (1 + det_groups(grid.x)) + 8)) return 1;
  const int w16 = (w + 15) / 16 * 16;
  const size_t msm = (size_t)w16 * 128 + (size_t)cs * 3 * (w16 + 2) * sizeof(float);
  if (cw % 64 == 0 && cwide_buf % 8 == 0 && wide_coff % 8 == 0 && msm <= 48 * 1024 && msm >= 64 * 41 * 4) {
#define LAUNCH_WGM(CS)                                                              \
  ::b200::launch_kernel(thin_wgrad_mma_kernel<CS>, grid, 256, msm, as_stream(stream),                  \
      thin, (const bf16*)wide, dw, dbias_wide, n, h, w, w16, cw, cwide_buf, wide_coff, wide_is_out, ds.part, ds.counters)
    switch (cs) {
      case 1: LAUNCH_WGM(1); break;
      case 2: LAUNCH_WGM(2); break;
      case 3: LAUNCH_WGM(3); break;
      default: LAUNCH_WGM(4); break;
    }
#undef LAUNCH_WGM
    B200_LAUNCH_CHECK();
    if (dbias_thin) {
      dim3 g2(32, cs);
      ::b200::launch_kernel(plane_sum_kernel, g2, 256, 0, as_stream(stream), thin, dbias_thin, n, cs, (long long)h * w, ds.part, ds.counters);
      B200_LAUNCH_CHECK();
    }
    return 0;
  }
  const size_t smem = (size_t)cs * 3 * (w + 2) * sizeof(float);
  B200_REQUIRE(smem <= 48 * 1024, "thin_wgrad: row too wide");
#define LAUNCH_WG(CS)                                                               \
  ::b200::launch_kernel(thin_wgrad_kernel<CS>, grid, 256, smem, as_stream(stream),                     \
      thin, (const bf16*)wide, dw, dbias_wide, n, h, w, cw, cwide_buf, wide_coff, wide_is_out, ds.part, ds.counters)
  switch (cs) {
    case 1: LAUNCH_WG(1); break;
    case 2: LAUNCH_WG(2); break;
    case 3: LAUNCH_WG(3); break;
    default: LAUNCH_WG(4); break;
  }
#undef LAUNCH_WG
  B200_LAUNCH_CHECK();
  if (dbias_thin) {
    dim3 g2(32, cs);
    ::b200::launch_kernel(plane_sum_kernel, g2, 256, 0, as_stream(stream), thin, dbias_thin, n, cs, (long long)h * w, ds.part, ds.counters);
    B200_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
