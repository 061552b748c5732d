// Convolution weight gradient on tcgen05 tensor cores (sm_100a), generic (any stride / kernel).
//
//   dW[t][ci][co] = sum over pixels p of  X[p @ tap t, ci] * dY[p, co]
//
// GEMM view: D[M = 128 input channels, N = BN output channels], K = pixels.  Both operands are
// the NHWC activation tiles exactly as TMA lands them ([128 pixel rows][64 channels], SWIZZLE_128B),
// consumed MN-major (a_major = b_major = 1): no transposes, no im2col.
// Work item = (tap, ci block of 128, co block of BN, pixel split); fp32 accumulation in TMEM over
// the item's pixel tiles, then fp32 red.global.add into dW laid out OIHW (nn.Conv2d.weight.grad).
//
// Reference call site: autograd backward of nn.Conv2d, block.py:238 (aten::convolution_backward).
#include "common.cuh"
#include "colsum.cuh"
#include "sm100_ptx.cuh"

namespace b200 {
namespace {

constexpr int kThreads = 192;
constexpr int kMaxStages = 4;
constexpr uint32_t kAtomBytes = 128 * 128;  // [128 px][64 ch] bf16

struct WgradParams {
  CUtensorMap x_map;
  CUtensorMap dy_map;
  int tw, th, tn, tiles_x, tiles_y, tiles_n, pixel_tiles;
  int taps, kw, stride, pad;
  int m_blocks, n_blocks, splits, tiles_per_split, total_items;
  int x_coff, dy_coff, cin, cout;
  int BN, n_atoms;
  int stages;
  uint32_t stage_bytes;
  float scale;
  float* dw;
  int pack2;   // cin <= 64: the two 64-channel atoms of the M = 128 operand hold TWO TAPS of the same 64 input channels
  float* ws;   // split-K partials [split][tap][co][ci] (nullptr when splits == 1: single owner, direct +=)
};

__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem =
      reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tfull_bar[0], 1);
    mbar_init(&tfull_bar[1], 1);
    mbar_init(&tempty_bar[0], 4);
    mbar_init(&tempty_bar[1], 4);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // Dependents may be scheduled from here on: this CTA already owns its TMEM columns, so a co-resident
  // CTA of the next kernel can never make it wait for an allocation (which would deadlock, because that
  // CTA in turn waits for this grid to complete).
  pdl_trigger();
  pdl_wait();   // everything above overlapped the previous kernel's tail
  const int acc_cols = (p.BN + 31) & ~31;

  // item -> (t, mb, nb, ks)
  auto decode = [&](int item, int& t, int& mb, int& nb, int& ks) {
    ks = item % p.splits;
    int r = item / p.splits;
    nb = r % p.n_blocks;
    r /= p.n_blocks;
    mb = r % p.m_blocks;
    t = r / p.m_blocks;
  };

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int t, mb, nb, ks;
      decode(item, t, mb, nb, ks);
      const int ta = p.pack2 ? 2 * t : t;   // first (or only) tap of the item
      const int ky = ta / p.kw, kx = ta % p.kw;
      const int pt0 = ks * p.tiles_per_split;
      const int pt1 = min(pt0 + p.tiles_per_split, p.pixel_tiles);
      for (int pt = pt0; pt < pt1; ++pt) {
        const int x0 = (pt % p.tiles_x) * p.tw;
        const int y0 = ((pt / p.tiles_x) % p.tiles_y) * p.th;
        const int n0 = (pt / (p.tiles_x * p.tiles_y)) * p.tn;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sx = smem + (size_t)stage * p.stage_bytes;
          uint8_t* sy = sx + 2 * kAtomBytes;
          mbar_expect_tx(&full_bar[stage], (2 + p.n_atoms) * kAtomBytes);
          const int cx = x0 * p.stride + kx - p.pad, cy = y0 * p.stride + ky - p.pad;
          tma_load_4d(sx, &p.x_map, &full_bar[stage], p.x_coff + mb * 128, cx, cy, n0);
          if (p.pack2) {   // atom 1 = the same 64 channels seen through the item's second tap
            const int t2 = (2 * t + 1 < p.taps) ? 2 * t + 1 : 2 * t;
            const int cx2 = x0 * p.stride + (t2 % p.kw) - p.pad, cy2 = y0 * p.stride + (t2 / p.kw) - p.pad;
            tma_load_4d(sx + kAtomBytes, &p.x_map, &full_bar[stage], p.x_coff, cx2, cy2, n0);
          } else {
            tma_load_4d(sx + kAtomBytes, &p.x_map, &full_bar[stage], p.x_coff + mb * 128 + 64, cx, cy, n0);
          }
          for (int j = 0; j < p.n_atoms; ++j)
            tma_load_4d(sy + j * kAtomBytes, &p.dy_map, &full_bar[stage],
                        p.dy_coff + nb * p.BN + j * 64, x0, y0, n0);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, p.BN, 1, 1);
    // MN-major SW128: LBO = byte distance between 64-channel atoms, SBO = 8 pixel rows = 1024 B
    const uint64_t desc_hi = make_smem_desc(0, kAtomBytes, 1024, LAYOUT_SW128, 0);
    const uint32_t smem_base = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int t, mb, nb, ks;
      decode(item, t, mb, nb, ks);
      const int pt0 = ks * p.tiles_per_split;
      const int pt1 = min(pt0 + p.tiles_per_split, p.pixel_tiles);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem + acc * acc_cols;
      for (int pt = pt0; pt < pt1; ++pt) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_base + stage * p.stage_bytes;
        const uint32_t b_addr = a_addr + 2 * kAtomBytes;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t ad = desc_hi | (uint64_t)(((a_addr + k * 2048) >> 4) & 0x3FFF);
            const uint64_t bd = desc_hi | (uint64_t)(((b_addr + k * 2048) >> 4) & 0x3FFF);
            umma_f16(d_tmem, ad, bd, idesc, (pt > pt0) || (k > 0));
          }
          umma_commit(&empty_bar[stage]);
          if (pt == pt1 - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int t, mb, nb, ks;
      decode(item, t, mb, nb, ks);
      int ci = mb * 128 + m;
      if (p.pack2) {   // rows 0..63 -> tap 2t, rows 64..127 -> tap 2t + 1 (absent for the last item of an odd tap count)
        ci = m & 63;
        t = 2 * t + (m >> 6);
        if (t >= p.taps) ci = p.cin;   // nothing to store
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem + ((uint32_t)(quad * 32) << 16) + acc * acc_cols;
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_row + c0, r);
        tmem_ld_wait();
        if (ci < p.cin) {
          // no atomics: an element of dW is owned by exactly one item per pixel split
          if (p.ws) {
            float* slab = p.ws + ((size_t)ks * p.taps + t) * p.cout * p.cin;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int co = nb * p.BN + c0 + j;
              if (co < p.cout) slab[(size_t)co * p.cin + ci] = __uint_as_float(r[j]);   // lanes = consecutive ci
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int co = nb * p.BN + c0 + j;
              if (co < p.cout) p.dw[((size_t)co * p.cin + ci) * p.taps + t] += p.scale * __uint_as_float(r[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// Per-channel column sum of an NHWC bf16 slice: db[c] += scale * sum_p dy[p, coff + c]
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ db,
                                                     long long npix, int cdy, int coff, int c, float scale,
                                                     float* __restrict__ part, unsigned* __restrict__ counter) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[256 * 8];
  colsum_vec(dy, npix, cdy, coff, c, scale, db, red, part, counter);
}

// Second pass of the split-K weight gradient: dW (OIHW) += scale * sum over the pixel splits, in split order.
// ws is tap-major [split][tap][co][ci] (the layout the tcgen05 epilogue writes with full 128-byte lines); one block owns
// (co, 32 input channels): 32 x 8 threads read the taps' values coalesced along ci, transpose through shared memory and
// update the 32 * taps CONTIGUOUS floats of dW[co][ci0 .. ci0+32][taps] with coalesced read-modify-writes.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw,
                                                           int splits, int taps, int cout, int cin, float scale) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32 * 17];   // [ci][tap], taps <= 16, padded
  const int co = blockIdx.y, ci0 = blockIdx.x * 32, tid = threadIdx.x;
  const int cl = tid & 31, tl = tid >> 5;   // 32 input channels x 8 tap lanes
  const int nci = cin - ci0 < 32 ? cin - ci0 : 32;
  const size_t per = (size_t)taps * cout * cin;
  if (cl < nci) {
    for (int t = tl; t < taps; t += 8) {
      const float* src = ws + ((size_t)t * cout + co) * cin + ci0 + cl;
      float s = 0.f;
#pragma unroll 4
      for (int k = 0; k < splits; ++k) s += src[(size_t)k * per];
      tile[cl * 17 + t] = s;
    }
  }
  __syncthreads();
  float* dst = dw + ((size_t)co * cin + ci0) * taps;
  for (int i = tid; i < nci * taps; i += 256) dst[i] += scale * tile[(i / taps) * 17 + (i % taps)];
}

// One thread converts all taps of one (row, col) weight: the fp32 source is read as `taps` consecutive
// floats (coalesced across the warp along the source's fastest dimension) instead of one strided gather
// per tap, and every tap plane of the destination is written with consecutive columns.
__global__ void pack_weights_kernel(const b200_pack_entry* __restrict__ table, int count) {
  pdl_trigger();
  pdl_wait();
  const b200_pack_entry e = table[blockIdx.y];
  const long long total = (long long)e.rows_pad * e.cols_pad;
  const long long plane = total;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e.dst);
  const int mul = e.co_mul ? e.co_mul : 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % e.cols_pad);
    const int r = (int)(i / e.cols_pad);
    const float* src = nullptr;
    if (e.mode == 0) {  // rows = co, cols = ci
      if (r < e.cout && c < e.cin) src = e.src + ((size_t)(r * mul + e.co_off) * e.cin + c) * e.taps;
    } else {            // rows = ci, cols = co
      if (r < e.cin && c < e.cout) src = e.src + ((size_t)(c * mul + e.co_off) * e.cin + r) * e.taps;
    }
    for (int t = 0; t < e.taps; ++t) dst[t * plane + i] = __float2bfloat16(src ? src[t] : 0.f);
  }
}

__global__ void pack_cat_kernel(const b200_packcat_entry* __restrict__ table) {
  pdl_trigger();
  pdl_wait();
  const b200_packcat_entry e = table[blockIdx.y];
  const long long total = (long long)e.n_rows * e.cout;
  const size_t plane = (size_t)e.rows_pad * e.cols_pad;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e.dst);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // the destination's fastest dimension is the fastest thread dimension
    const int r = e.mode ? (int)(i / e.cout) : (int)(i % e.n_rows);    // input-channel index within the block
    const int co = e.mode ? (int)(i % e.cout) : (int)(i / e.n_rows);
    const float* src = e.src + ((size_t)co * e.cin + e.ci_off + r) * e.taps;
    const size_t row = e.mode ? (size_t)(e.row_off + r) : (size_t)(e.row_off + co);
    const size_t col = e.mode ? (size_t)(e.col_off + co) : (size_t)(e.col_off + r);
    for (int t = 0; t < e.taps; ++t)
      dst[(size_t)t * plane + row * e.cols_pad + col] = __float2bfloat16(e.scale * src[t]);
  }
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" int b200_pack_cat(const b200_packcat_entry* table_dev, int32_t count, int32_t max_elems,
                             b200_stream_t stream) {
  if (count <= 0) return 0;
  B200_REQUIRE(table_dev, "b200_pack_cat: null table");
  int bx = (max_elems + 256 * 8 - 1) / (256 * 8);
  if (bx < 1) bx = 1;
  if (bx > 32) bx = 32;
  dim3 grid(bx, count);
  ::b200::launch_kernel(pack_cat_kernel, grid, 256, 0, as_stream(stream), table_dev);
  B200_LAUNCH_CHECK();
  return 0;
}

extern "C" int b200_conv_wgrad(const b200_wgrad_desc* d, const void* x, const void* dy, float* dw,
                               float* dbias, b200_stream_t stream) {
  B200_REQUIRE(d && x && dy, "b200_conv_wgrad: null argument");
  B200_REQUIRE(d->cx % 8 == 0 && d->cdy % 8 == 0 && d->x_coff % 8 == 0 && d->dy_coff % 8 == 0,
               "b200_conv_wgrad: channel pitches/offsets must be multiples of 8");
  const long long npix = (long long)d->n * d->h_out * d->w_out;
  if (dbias) {
    B200_REQUIRE(d->cout % 8 == 0 && d->cout <= 2048, "b200_conv_wgrad: bias gradient needs cout %% 8 == 0, <= 2048");
    // one block covers (256 / (cout/8)) pixel lanes x 4 pixels in flight; at most 4 waves of blocks
    const long long per_block = (long long)(256 / (d->cout / 8)) * 16;
    long long gx = (npix + per_block - 1) / per_block;
    if (gx > 148 * 2) gx = 148 * 2;   // the last block adds the per-block partials: keep them few
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, 1);
    DetScratch ds;
    if (det_scratch(&ds, (size_t)(gx + det_groups((int)gx)) * d->cout, 1 + det_groups((int)gx))) return 1;
    ::b200::launch_kernel(colsum_kernel, grid, 256, 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(dy),
                                                      dbias, npix, d->cdy, d->dy_coff, d->cout,
                                                      d->scale, ds.part, ds.counters);
    B200_LAUNCH_CHECK();
  }
  if (!dw) return 0;

  const int kSmemBytes = 200 * 1024;
  B200_ENSURE_SMEM(conv_wgrad_kernel, kSmemBytes);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.tw = d->w_out > 8 ? 16 : (d->w_out > 4 ? 8 : 4);
  int th_max = 128 / p.tw, hp = 1;
  while (hp < d->h_out) hp <<= 1;
  p.th = hp < th_max ? hp : th_max;
  p.tn = 128 / (p.tw * p.th);
  p.tiles_x = (d->w_out + p.tw - 1) / p.tw;
  p.tiles_y = (d->h_out + p.th - 1) / p.th;
  p.tiles_n = (d->n + p.tn - 1) / p.tn;
  p.pixel_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  p.taps = d->kh * d->kw;
  p.kw = d->kw;
  p.stride = d->stride;
  p.pad = d->pad;
  p.cin = d->cin;
  p.cout = d->cout;
  p.x_coff = d->x_coff;
  p.dy_coff = d->dy_coff;
  p.pack2 = (d->cin <= 64) ? 1 : 0;
  p.m_blocks = (d->cin + 127) / 128;
  int BN = d->cout <= 128 ? ((d->cout + 15) / 16) * 16 : 128;
  p.BN = BN;
  p.n_atoms = (BN + 63) / 64;
  p.n_blocks = (d->cout + BN - 1) / BN;
  const int tap_items = p.pack2 ? (p.taps + 1) / 2 : p.taps;
  const int base_items = tap_items * p.m_blocks * p.n_blocks;
  const int sms = sm_count();
  int splits = (2 * sms + base_items - 1) / base_items;  // ~2 items per SM
  if (splits > p.pixel_tiles) splits = p.pixel_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.pixel_tiles + splits - 1) / splits;
  p.splits = (p.pixel_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.total_items = base_items * p.splits;
  p.stage_bytes = (2 + p.n_atoms) * kAtomBytes;
  p.stages = (kSmemBytes - 2048) / (int)p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  B200_REQUIRE(p.stages >= 2, "b200_conv_wgrad: smem");
  p.scale = d->scale;
  p.dw = dw;
  p.ws = nullptr;
  if (p.splits > 1) {
    DetScratch ds;
    if (det_scratch(&ds, (size_t)p.splits * p.taps * d->cout * d->cin, 0)) return 1;
    p.ws = ds.part;
  }
  {
    uint64_t dims[4] = {(uint64_t)(d->x_coff + d->cin), (uint64_t)d->w_in, (uint64_t)d->h_in,
                        (uint64_t)d->n};
    uint64_t strides[3] = {(uint64_t)d->cx * 2, (uint64_t)d->w_in * d->cx * 2,
                           (uint64_t)d->h_in * d->w_in * d->cx * 2};
    uint32_t box[4] = {64, (uint32_t)(p.tw * d->stride), (uint32_t)(p.th * d->stride), (uint32_t)p.tn};
    uint32_t es[4] = {1, (uint32_t)d->stride, (uint32_t)d->stride, 1};
    if (make_tensor_map(&p.x_map, x, 4, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    uint64_t dims[4] = {(uint64_t)(d->dy_coff + d->cout), (uint64_t)d->w_out, (uint64_t)d->h_out,
                        (uint64_t)d->n};
    uint64_t strides[3] = {(uint64_t)d->cdy * 2, (uint64_t)d->w_out * d->cdy * 2,
                           (uint64_t)d->h_out * d->w_out * d->cdy * 2};
    uint32_t box[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
    if (make_tensor_map(&p.dy_map, dy, 4, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B))
      return 1;
  }
  const int grid = p.total_items < sms ? p.total_items : sms;
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  ::b200::launch_kernel(conv_wgrad_kernel, grid, kThreads, smem, as_stream(stream), p);
  B200_LAUNCH_CHECK();
  if (p.ws) {
    B200_REQUIRE(p.taps <= 16, "b200_conv_wgrad: at most 16 taps");
    dim3 rgrid((d->cin + 31) / 32, d->cout);
    ::b200::launch_kernel(wgrad_reduce_kernel, rgrid, 256, 0, as_stream(stream), (const float*)p.ws, dw, p.splits,
                          p.taps, (int)d->cout, (int)d->cin, d->scale);
    B200_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int b200_pack_weights(const b200_pack_entry* table_dev, int32_t count, int32_t max_elems,
                                 b200_stream_t stream) {
  if (count <= 0) return 0;
  B200_REQUIRE(table_dev, "b200_pack_weights: null table");
  int bx = (max_elems + 256 * 8 - 1) / (256 * 8);
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  dim3 grid(bx, count);
  ::b200::launch_kernel(pack_weights_kernel, grid, 256, 0, as_stream(stream), table_dev, count);
  B200_LAUNCH_CHECK();
  return 0;
}
