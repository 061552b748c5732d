// TMEM-persistent, stage-merged residual dense block (forward AND gather-form input gradient).
//
// A dense block is 5 stacked 3x3 convs where conv_k consumes [x, x1..x_{k-1}].  Computed conv by
// conv, four of them have only 32 output channels and the tcgen05 MMA is shared-memory-operand
// bound at 40 % of the tensor pipe (N = 32; profiles/r01_umma_issue_rate_probe.log).  Here the
// block is computed INPUT SLICE by input slice instead:
//   stage j consumes the newest 32/64-channel slice S_j (x for j = 0, then x1..x4) and adds its
//   contribution to ALL convs that still need it:  D[:, 32j:192] += S_j (*) W_stage_j   (N = 192-32j)
// The fp32 partial sums of one 256-row tile (2 x 128 rows x 192 columns) stay resident in TMEM
// for the whole block.  After stage j the 32 (last stage: 64) columns that just became complete
// are read back, finished (bias / LeakyReLU / mask / residuals) and written to HBM as bf16 -- they
// are the next stage's input slice.  The halo rows of that slice belong to the neighbouring tiles,
// so CTAs exchange per-tile stage counters in global memory (release/acquire + proxy fence before
// the TMA loads); all CTAs of the launch are co-resident (cooperative launch, <= 1 tile per SM).
// N is 192/160/128/96/64 instead of 32/32/32/32/64: ~2.3x fewer MMA cycles per block.
//
// The input-gradient of the block in gather form has exactly the same shape with the slices taken
// in reverse order (dO, dY4, dY3, dY2, dY1 -> d(x4), d(x3), d(x2), d(x1), d(x)).
//
// Reference: ResidualDenseBlock_5C.forward + RRDB residuals (RRDBNet_arch.py:89-96,150-163) and
// their autograd input gradients.
#include <stdlib.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace b200 {
namespace {

constexpr int kThreads = 384;   // A producer, 2 MMA issuers, 8 epilogue warps, B producer
constexpr int kTileM = 256;
constexpr int kAStages = 2;
constexpr int kBStages = 3;
constexpr int kNTotal = 192;
constexpr uint32_t kBStageBytes = 32 * 1024;   // one 64-channel tap tile (24 KB) or three 32-channel tap tiles (<= 30 KB)

struct StageEpi {
  __nv_bfloat16* out;       // flat [P, out_c]
  const float* bias;        // indexed by column within the stage's completing slice
  const __nv_bfloat16* mask;
  const __nv_bfloat16* res1;
  const __nv_bfloat16* res2;
  int out_c, out_coff, mask_c, mask_coff, res1_c, res1_coff, res2_c, res2_coff;
  float alpha, beta1, beta2, slope, mask_slope;
  int act;
};

struct RdbParams {
  CUtensorMap in_map[5];
  CUtensorMap w_map[5];
  StageEpi epi[5];
  int in_ch[5];        // channel coordinate of the stage's input slice
  int nk[5];           // K = 16 * nk channels (4 or 2)
  int P, Hp, Wp, HpWp, h, w;
  int total_tiles, halo, nbox, box_rows;
  uint32_t a_bytes, a_stage_bytes;
  int tap_sign;        // +1 forward taps, -1 input-gradient taps
  int* flags;          // per tile: flag_base + number of finished stages
  long long* dbg;      // optional timeline (clock64), 32 slots per CTA
  int flag_base;
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Operands of the stage epilogue that do not depend on the accumulator (bias, mask, residuals) are
// requested BEFORE the wait on the MMAs so that their latency is off the inter-stage critical path.
struct EpiPre {
  uint4 lm[4], l1[4], l2[4];
  float4 b[8];
};

__device__ __forceinline__ void prefetch32(const StageEpi& e, EpiPre& q, int c0, long long m) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c = c0 + g * 8;
    if (e.mask) q.lm[g] = __ldg(reinterpret_cast<const uint4*>(e.mask + m * e.mask_c + e.mask_coff + c));
    if (e.res1) q.l1[g] = __ldg(reinterpret_cast<const uint4*>(e.res1 + m * e.res1_c + e.res1_coff + c));
    if (e.res2) q.l2[g] = __ldg(reinterpret_cast<const uint4*>(e.res2 + m * e.res2_c + e.res2_coff + c));
    if (e.bias) {
      q.b[2 * g] = __ldg(reinterpret_cast<const float4*>(e.bias + c));
      q.b[2 * g + 1] = __ldg(reinterpret_cast<const float4*>(e.bias + c + 4));
    }
  }
}

// finish 32 accumulator columns of one row
__device__ __forceinline__ void finish32(const StageEpi& e, const uint32_t* acc, const EpiPre& q, int c0,
                                         long long m) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int c = c0 + g * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[g * 8 + j]);
    if (e.bias) {
      const float4 b0 = q.b[2 * g], b1 = q.b[2 * g + 1];
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= e.alpha;
    if (e.res1) {
      float r[8];
      unpack8(q.l1[g], r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(e.beta1, r[j], v[j]);
    }
    if (e.res2) {
      float r[8];
      unpack8(q.l2[g], r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(e.beta2, r[j], v[j]);
    }
    if (e.act) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * e.slope;
    }
    if (e.mask) {
      float r[8];
      unpack8(q.lm[g], r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = r[j] > 0.f ? v[j] : v[j] * e.mask_slope;
    }
    *reinterpret_cast<uint4*>(e.out + m * e.out_c + e.out_coff + c) = pack8(v);
  }
}

#define RDBG(slot) do { if (p.dbg && lane == 0) p.dbg[blockIdx.x * 32 + (slot)] = clock64(); } while (0)

__global__ void __launch_bounds__(kThreads, 1)
rdb_persist_kernel(const __grid_constant__ RdbParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem =
      reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t a_full[kAStages], a_empty[kAStages], b_full[kBStages], b_empty[kBStages], stage_done;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  if (warp == 0) RDBG(0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kAStages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 2);
    }
    for (int s = 0; s < kBStages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 2);
    }
    mbar_init(&stage_done, 2);
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
  const uint32_t b_ring_off = kAStages * p.a_stage_bytes;

  if (warp == 0) {
    // ------------------------------------------------------------ A producer (input slices + halo)
    const int row0 = tile * kTileM - p.halo;
    for (int j = 0; j < 5; ++j) {
      if (j > 0) {
        // the slice produced by stage j-1: ours and both neighbours' (halo rows)
        const int want = p.flag_base + j;
        const int lo = tile > 0 ? tile - 1 : tile;
        const int hi = tile + 1 < p.total_tiles ? tile + 1 : tile;
        if (lane <= hi - lo) {
          const int* f = p.flags + lo + lane;
          while (ld_acquire(f) < want) {
          }
        }
        __syncwarp();
        asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes -> TMA (async proxy) reads
      }
      RDBG(1 + j * 6);  // flags satisfied
      const int as = j & 1;
      const uint32_t aph = (j >> 1) & 1;
      mbar_wait(&a_empty[as], aph ^ 1);
      if (elect_one()) {
        uint8_t* sa = smem + (size_t)as * p.a_stage_bytes;
        mbar_expect_tx(&a_full[as], p.a_bytes);
        for (int bx = 0; bx < p.nbox; ++bx)
          tma_load_2d(sa + (size_t)bx * p.box_rows * 128, &p.in_map[j], &a_full[as], p.in_ch[j],
                      row0 + bx * p.box_rows);
      }
      __syncwarp();
    }
  } else if (warp == 11) {
    // ------------------------------------------------------------ B producer (stage weights, free running)
    int bs = 0;
    uint32_t bph = 0;
    for (int j = 0; j < 5; ++j) {
      const int tpb = (p.nk[j] == 4) ? 1 : 3;   // taps per weight slot: fewer barrier round trips for the K = 32 stages
      const uint32_t tap_bytes = (uint32_t)(kNTotal - 32 * j) * (p.nk[j] == 4 ? 128 : 64);
      for (int t0 = 0; t0 < 9; t0 += tpb) {
        mbar_wait(&b_empty[bs], bph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&b_full[bs], tap_bytes * tpb);
          for (int q = 0; q < tpb; ++q)
            tma_load_3d(smem + b_ring_off + (size_t)bs * kBStageBytes + (size_t)q * tap_bytes, &p.w_map[j],
                        &b_full[bs], 0, 0, t0 + q);
        }
        __syncwarp();
        if (++bs == kBStages) {
          bs = 0;
          bph ^= 1;
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------ MMA issuers (one 128-row half each)
    const int half = warp - 1;
    const uint64_t desc_hi = make_smem_desc(0, 16, 1024, LAYOUT_SW128, 0);
    const uint64_t desc_b64 = make_smem_desc(0, 16, 512, LAYOUT_SW64, 0);   // 32-channel weight tiles: 64-byte rows
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t d_half = tmem + half * kNTotal;
    int bs = 0;
    uint32_t bph = 0;
    for (int j = 0; j < 5; ++j) {
      const int as = j & 1;
      const uint32_t aph = (j >> 1) & 1;
      const int N = kNTotal - 32 * j;
      const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
      const uint32_t d_tmem = d_half + 32 * j;
      const int nk = p.nk[j];
      mbar_wait(&a_full[as], aph);
      tc_fence_after();
      if (half == 0) RDBG(2 + j * 6);  // A landed
      const uint32_t a_base = smem_base + as * p.a_stage_bytes + (uint32_t)(p.halo + half * 128) * 128;
      const int tpb = (nk == 4) ? 1 : 3;
      const uint32_t tap_bytes = (uint32_t)N * (nk == 4 ? 128 : 64);
      for (int t0 = 0; t0 < 9; t0 += tpb) {
        mbar_wait(&b_full[bs], bph);
        tc_fence_after();
        const uint32_t b_slot = smem_base + b_ring_off + bs * kBStageBytes;
        if (elect_one()) {
          for (int q = 0; q < tpb; ++q) {
            const int t = t0 + q;
            const int shift = p.tap_sign * ((t / 3 - 1) * p.Wp + (t % 3 - 1));
            const uint32_t a_addr = a_base + (uint32_t)(shift * 128);
            const uint32_t b_addr = b_slot + q * tap_bytes;
            for (int k = 0; k < nk; ++k) {
              const uint64_t ad = desc_hi | (uint64_t)(((a_addr + k * 32) >> 4) & 0x3FFF);
              const uint64_t bd = (nk == 4 ? desc_hi : desc_b64) | (uint64_t)(((b_addr + k * 32) >> 4) & 0x3FFF);
              umma_f16(d_tmem, ad, bd, idesc, (j | t | k) != 0);
            }
          }
          umma_commit(&b_empty[bs]);
          if (t0 + tpb >= 9) {
            umma_commit(&a_empty[as]);
            umma_commit(&stage_done);
          }
        }
        __syncwarp();
        if (++bs == kBStages) {
          bs = 0;
          bph ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 3..10)
    const int quad = warp & 3;
    const int half = (warp - 3) >> 2;
    const long long m = (long long)tile * kTileM + half * 128 + quad * 32 + lane;
    const int nimg = (int)(m / p.HpWp);
    const int rem = (int)(m - (long long)nimg * p.HpWp);
    const int yp = rem / p.Wp, xp = rem - yp * p.Wp;
    const bool valid = (m < p.P) && yp >= 1 && yp <= p.h && xp >= 1 && xp <= p.w;
    const uint32_t t_row = tmem + ((uint32_t)(quad * 32) << 16) + half * kNTotal;
    for (int j = 0; j < 5; ++j) {
      EpiPre q;
      if (valid) prefetch32(p.epi[j], q, 0, m);
      mbar_wait(&stage_done, j & 1);
      tc_fence_after();
      if (warp == 3) RDBG(3 + j * 6);  // MMAs of the stage complete
      const int ncols = (j == 4) ? 64 : 32;
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + 32 * j + c0, r);
        if (c0 > 0 && valid) prefetch32(p.epi[j], q, c0, m);
        tmem_ld_wait();
        if (valid) finish32(p.epi[j], r, q, c0, m);
      }
      // publish: every epilogue thread's stores -> device scope, then one flag update per tile
      if (warp == 3) RDBG(4 + j * 6);  // stores issued
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 3 && lane == 0) {
        __threadfence();   // cumulative over the other epilogue threads' stores (ordered by the barrier)
        st_release(p.flags + tile, p.flag_base + j + 1);
      }
      if (warp == 3) RDBG(5 + j * 6);  // flag published
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace
}  // namespace b200

using namespace b200;

// Host entry: one launch = one dense block (forward or gather-form input gradient) over the
// positions [pos0, pos0 + n_img*(h+2)*(w+2)) of the flat tensors (n_img images per launch so that the
// number of 256-row tiles does not exceed the SM count).
extern "C" int b200_rdb_persist(const b200_rdb_desc* d, int* flags, int32_t flag_base,
                                b200_stream_t stream) {
  B200_REQUIRE(d && flags, "b200_rdb_persist: null argument");
  const int kSmemBytes = 202 * 1024;
  B200_ENSURE_SMEM(rdb_persist_kernel, kSmemBytes);
  RdbParams p;
  memset(&p, 0, sizeof(p));
  p.h = d->h; p.w = d->w;
  p.Hp = d->h + 2; p.Wp = d->w + 2; p.HpWp = p.Hp * p.Wp;
  const long long P = (long long)d->n * p.HpWp;
  p.P = (int)P;
  p.total_tiles = (int)((P + kTileM - 1) / kTileM);
  const int sms = sm_count();
  B200_REQUIRE(p.total_tiles <= sms, "b200_rdb_persist: %d tiles exceed the %d SMs (split the batch)", p.total_tiles, sms);
  p.halo = p.Wp + 1;
  const int region = kTileM + 2 * p.halo;
  p.nbox = (region + 255) / 256;
  p.box_rows = (((region + p.nbox - 1) / p.nbox) + 7) & ~7;
  if (p.box_rows > 256) { p.nbox += 1; p.box_rows = (((region + p.nbox - 1) / p.nbox) + 7) & ~7; }
  p.a_bytes = (uint32_t)p.nbox * p.box_rows * 128;
  p.a_stage_bytes = (p.a_bytes + 1023) & ~1023u;
  B200_REQUIRE(kAStages * p.a_stage_bytes + kBStages * kBStageBytes + 2048 <= (uint32_t)kSmemBytes,
               "b200_rdb_persist: image too wide for the shared-memory A region (w=%d)", d->w);
  for (int j = 0; j < 5; ++j) {
    const b200_rdb_stage* s = &d->stage[j];
    B200_REQUIRE(s->x && s->w_packed && s->out, "b200_rdb_persist: null stage pointer");
    B200_REQUIRE(s->cin == 64 || s->cin == 32, "b200_rdb_persist: stage input must be 64 or 32 channels");
    p.in_ch[j] = s->cin_off;
    p.nk[j] = s->cin / 16;
    {
      uint64_t dims[2] = {(uint64_t)(s->cin_off + s->cin), (uint64_t)P};
      uint64_t strides[1] = {(uint64_t)s->cx * 2};
      uint32_t box[2] = {64, (uint32_t)p.box_rows};
      if (make_tensor_map(&p.in_map[j], s->x, 2, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    }
    {
      const int N = kNTotal - 32 * j;
      const uint64_t kc = (uint64_t)s->cin;   // packed weights are [9][N][cin] (cin = 64 or 32)
      uint64_t dims[3] = {kc, (uint64_t)N, 9};
      uint64_t strides[2] = {kc * 2, (uint64_t)N * kc * 2};
      uint32_t box[3] = {(uint32_t)kc, (uint32_t)N, 1};
      if (make_tensor_map(&p.w_map[j], s->w_packed, 3, dims, strides, box, nullptr,
                          kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B))
        return 1;
    }
    StageEpi& e = p.epi[j];
    e.out = (__nv_bfloat16*)s->out; e.out_c = s->out_c; e.out_coff = s->out_coff;
    e.bias = s->bias;
    e.mask = (const __nv_bfloat16*)s->mask; e.mask_c = s->mask_c; e.mask_coff = s->mask_coff;
    e.res1 = (const __nv_bfloat16*)s->res1; e.res1_c = s->res1_c; e.res1_coff = s->res1_coff;
    e.res2 = (const __nv_bfloat16*)s->res2; e.res2_c = s->res2_c; e.res2_coff = s->res2_coff;
    e.alpha = s->alpha; e.beta1 = s->beta1; e.beta2 = s->beta2; e.slope = s->slope; e.mask_slope = s->mask_slope;
    e.act = s->act;
  }
  p.tap_sign = d->flip_taps ? -1 : 1;
  {
    const char* e = getenv("B200_RDB_DBG_PTR");
    p.dbg = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr;
  }
  p.flags = flags;
  p.flag_base = flag_base;
  // stage counters start from zero for every launch (stream-ordered after the previous block finished)
  B200_CHECK_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * p.total_tiles, as_stream(stream)));

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.total_tiles);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kAStages * p.a_stage_bytes + kBStages * kBStageBytes + 1024;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the neighbour flags cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B200_CHECK_CUDA(cudaLaunchKernelEx(&cfg, rdb_persist_kernel, p));
  g_launches.fetch_add(1);
  return 0;
}
