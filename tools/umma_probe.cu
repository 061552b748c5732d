// Standalone probe for tcgen05/TMA descriptor semantics on sm_100a (run once on a B200 via gpurun).
// Answers, against a CPU reference:
//   K-major SW128 operands via TMA                          (mode 0, shift = 0)
//   A-operand start address shifted by s rows of 128 B      (mode 0, shift > 0; base_offset 0 / computed)
//   MN-major SW128 operands (the wgrad formulation)         (mode 1), shifted along K too
//   MN-major B with N = 32 inside a 64-wide SW128 atom      (mode 1, N = 32) and SW64 (mode 2)
//   MMA issue rate vs N with smem-resident operands         (mode 3)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu -I../trainner_b200/csrc
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "sm100_ptx.cuh"

using namespace b200;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

typedef CUresult (*EncodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled_t g_encode = nullptr;

static void make_map_2d(CUtensorMap* m, void* ptr, uint64_t inner, uint64_t outer,
                        uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer,
                        CUtensorMapSwizzle sw) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed %d\n", (int)r);
    exit(3);
  }
}

struct Params {
  int mode;      // 0 K-major, 1 MN-major (B SW128), 2 MN-major (B SW64, N=32), 3 rate
  int N;         // MMA N
  int K;         // reduction length (elements)
  int shift;     // rows of shift applied to A's start address
  int base_off;  // 0: base_offset field = 0; 1: (addr >> 7) & 7
  int a_rows;    // rows of the A tile in smem (>= 128 + shift for mode 0, K + shift for mode 1)
  int reps;      // mode 3: MMAs to issue
  int a_mn;      // mode 3: A MN-major
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ bool wait_bounded(uint64_t* bar, uint32_t parity) {
  for (long i = 0; i < 20000000L; ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
             Params p, float* __restrict__ D, long long* __restrict__ cycles, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_load, bar_mma, bar_mma2[4];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_mma2[i], 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  uint8_t* sA = smem;
  uint8_t* sB = nullptr;
  bool ok = true;

  if (p.mode == 0) {
    const int kchunks = p.K / 64;
    const uint32_t a_chunk_bytes = p.a_rows * 128;
    sB = sA + ((kchunks * a_chunk_bytes + 1023) & ~1023u);
    const uint32_t b_chunk_bytes = p.N * 128;
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar_load, kchunks * (a_chunk_bytes + b_chunk_bytes));
      for (int c = 0; c < kchunks; ++c) {
        tma_load_2d(sA + c * a_chunk_bytes, &mapA, &bar_load, c * 64, 0);
        tma_load_2d(sB + c * b_chunk_bytes, &mapB, &bar_load, c * 64, 0);
      }
      ok = wait_bounded(&bar_load, 0);
      tc_fence_after();
      if (ok) {
        const uint32_t idesc = make_idesc_bf16(128, p.N, 0, 0);
        for (int c = 0; c < kchunks; ++c)
          for (int k = 0; k < 4; ++k) {
            uint32_t a_addr = smem_u32(sA + c * a_chunk_bytes) + p.shift * 128 + k * 32;
            uint32_t b_addr = smem_u32(sB + c * b_chunk_bytes) + k * 32;
            uint32_t bo = p.base_off ? ((a_addr >> 7) & 7) : 0;
            // reps (unused in this mode) = rows between consecutive 8-row groups of A (0: the canonical 8)
            uint64_t ad = make_smem_desc(a_addr, 16, p.reps ? p.reps * 128 : 1024, LAYOUT_SW128, bo);
            uint64_t bd = make_smem_desc(b_addr, 16, 1024, LAYOUT_SW128, 0);
            umma_f16(tmem, ad, bd, idesc, (c | k) != 0);
          }
        umma_commit(&bar_mma);
        ok = wait_bounded(&bar_mma, 0);
      }
      if (!ok) *err = 1;
    }
  } else if (p.mode == 1 || p.mode == 2) {
    // A: X tile [a_rows px][128 ch] as two 64-channel SW128 boxes; MN-major, M = 128.
    const uint32_t a_atom_bytes = p.a_rows * 128;
    sB = sA + ((2 * a_atom_bytes + 1023) & ~1023u);
    const uint32_t b_row_bytes = (p.mode == 1) ? 128 : 64;
    const uint32_t b_bytes = p.K * b_row_bytes;
    const int b_atoms = (p.mode == 1) ? (p.N + 63) / 64 : 1;
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar_load, 2 * a_atom_bytes + b_atoms * b_bytes);
      tma_load_2d(sA, &mapA, &bar_load, 0, 0);
      tma_load_2d(sA + a_atom_bytes, &mapA, &bar_load, 64, 0);
      for (int j = 0; j < b_atoms; ++j) tma_load_2d(sB + j * b_bytes, &mapB, &bar_load, j * 64, 0);
      ok = wait_bounded(&bar_load, 0);
      tc_fence_after();
      if (ok) {
        const uint32_t idesc = make_idesc_bf16(128, p.N, 1, 1);
        for (int k = 0; k < p.K / 16; ++k) {
          uint32_t a_addr = smem_u32(sA) + (p.shift + k * 16) * 128;
          uint32_t b_addr = smem_u32(sB) + k * 16 * b_row_bytes;
          uint32_t bo = p.base_off ? ((a_addr >> 7) & 7) : 0;
          uint64_t ad = make_smem_desc(a_addr, a_atom_bytes, 1024, LAYOUT_SW128, bo);
          uint64_t bd = (p.mode == 1) ? make_smem_desc(b_addr, b_bytes, 1024, LAYOUT_SW128, 0)
                                      : make_smem_desc(b_addr, b_bytes, 512, LAYOUT_SW64, 0);
          umma_f16(tmem, ad, bd, idesc, k != 0);
        }
        umma_commit(&bar_mma);
        ok = wait_bounded(&bar_mma, 0);
      }
      if (!ok) *err = 1;
    }
  } else {
    // rate probe: zero operands resident in smem, 4 rotating stage addresses.
    const uint32_t a_stage = 176 * 128;      // 128 rows x 64 bf16 (+ slack for shifted starts / strided groups)
    const uint32_t b_stage = 128 * 128;      // 16 KB (N <= 128 here)
    for (uint32_t i = threadIdx.x; i < (4 * (a_stage + b_stage)) / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    sB = sA + 4 * a_stage;
    const int nissue = p.a_rows > 0 ? p.a_rows : 1;
    if (warp < nissue) {
      // warp-uniform issue loop: operands live in uniform registers, one elected lane issues.
      // a_mn bit 0: A MN-major, bit 1: B MN-major (the wgrad formulation: both operands pixel-major)
      const uint32_t idesc = make_idesc_bf16(128, p.N, p.a_mn & 1, (p.a_mn >> 1) & 1);
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      const uint32_t a_kstep = (p.a_mn & 1) ? 16 * 128 : 32;
      const uint32_t b_kstep = (p.a_mn & 2) ? 16 * 128 : 32;
      const uint64_t adesc_hi = (p.a_mn & 1) ? (make_smem_desc(0, 8192, 1024, LAYOUT_SW128, 0))
                                       : (make_smem_desc(0, 16, p.K ? p.K * 128 : 1024, LAYOUT_SW128, 0));
      const uint64_t bdesc_hi = (p.a_mn & 2) ? make_smem_desc(0, 8192, 1024, LAYOUT_SW128, 0)
                                             : make_smem_desc(0, 16, 1024, LAYOUT_SW128, 0);
      long long t0 = clock64();
      for (int r = 0; r < p.reps; r += 4) {
        const int st = (r >> 2) & 3;
        const uint32_t a_addr = a0 + st * a_stage + p.shift * 128, b_addr = b0 + st * b_stage;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint64_t ad = adesc_hi | (uint64_t)(((a_addr + k * a_kstep) >> 4) & 0x3FFF);
            uint64_t bd = bdesc_hi | (uint64_t)(((b_addr + k * b_kstep) >> 4) & 0x3FFF);
            umma_f16(tmem + warp * 128 + ((p.base_off && (k & 1)) ? 64 : 0), ad, bd, idesc, (r | k) > 1);
          }
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar_mma2[warp]);
      __syncwarp();
      ok = wait_bounded(&bar_mma2[warp], 0);
      long long t1 = clock64();
      if (lane_id() == 0) {
        atomicMax((unsigned long long*)&cycles[blockIdx.x], (unsigned long long)(t1 - t0));
        if (!ok) *err = 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (p.mode != 3 && blockIdx.x == 0) {
    // epilogue: lane = threadIdx.x (row m), columns 0..N-1
    for (int c0 = 0; c0 < p.N; c0 += 16) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) D[threadIdx.x * p.N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// Pipeline emulation (mode "pipe"): the conv_flat issue structure without any TMA traffic.
// warps 0..nissue-1 issue G MMAs per stage (own accumulator), commit the stage's "empty" barrier
// and move on; warp 3 is the producer: waits "empty", arrives "full".  Measures what the
// barrier handshakes and commits cost on top of the raw MMA rate (mode 3).
struct PipeParams {
  int N, G, S, groups, nissue;
  int commit_only;   // 1: no full/empty handshake, just a commit every G MMAs
  int sep;           // bit 0: issuers read different A tiles, bit 1: different B tiles
  int dstride;       // TMEM column distance between the issuers' accumulators
  int data;          // 0: all-zero operands, 1: random bf16 operands
};

__global__ void __launch_bounds__(352, 1)
pipe_kernel(PipeParams p, long long* __restrict__ cycles, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8], empty[8], done[4];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], p.nissue);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&done[i], 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  const uint32_t a_stage = 144 * 128, b_stage = 256 * 128;
  for (uint32_t i = threadIdx.x; i < (4 * (a_stage + b_stage)) / 4; i += blockDim.x) {
    uint32_t r = (i + 1) * 2654435761u;
    r ^= r >> 13;
    // two bf16 in (0.0078 .. 2), random sign: realistic switching activity instead of all-zero operands
    uint32_t v = (0x3C003C00u | (r & 0x03FF03FFu) | ((r << 3) & 0x80008000u));
    reinterpret_cast<uint32_t*>(smem)[i] = p.data ? v : 0u;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  bool ok = true;
  if (warp < p.nissue) {
    const uint32_t idesc = make_idesc_bf16(128, p.N, 0, 0);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 4 * a_stage;
    const uint64_t desc_hi = make_smem_desc(0, 16, 1024, LAYOUT_SW128, 0);
    int st = 0;
    uint32_t ph = 0;
    long long t0 = clock64();
    for (int g = 0; g < p.groups; ++g) {
      if (!p.commit_only) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
      }
      const uint32_t a_addr = a0 + ((g + ((p.sep & 1) ? 2 * warp : 0)) & 3) * a_stage + (g % 7) * 128;
      const uint32_t b_addr = b0 + ((g + ((p.sep & 2) ? 2 * warp : 0)) & 3) * b_stage;
      if (elect_one()) {
        for (int j = 0; j < p.G; j += 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint64_t ad = desc_hi | (uint64_t)(((a_addr + k * 32) >> 4) & 0x3FFF);
            uint64_t bd = desc_hi | (uint64_t)(((b_addr + ((j >> 2) & 3) * 8192 + k * 32) >> 4) & 0x3FFF);
            umma_f16(tmem + warp * p.dstride, ad, bd, idesc, (g | j | k) != 0);
          }
        }
        umma_commit(&empty[st]);
      }
      __syncwarp();
      if (++st == p.S) {
        st = 0;
        ph ^= 1;
      }
    }
    if (elect_one()) umma_commit(&done[warp]);
    __syncwarp();
    ok = wait_bounded(&done[warp], 0);
    long long t1 = clock64();
    if (lane_id() == 0) {
      atomicMax((unsigned long long*)&cycles[blockIdx.x], (unsigned long long)(t1 - t0));
      if (!ok) *err = 1;
    }
  } else if (warp >= 4) {
    // spinners: emulate epilogue warps polling an mbarrier that completes only when issuer 0 is done
    mbar_wait(&done[0], 0);
  } else if (warp == 3 && !p.commit_only) {
    int st = 0;
    uint32_t ph = 0;
    for (int g = 0; g < p.groups; ++g) {
      mbar_wait(&empty[st], ph ^ 1);
      if (elect_one()) mbar_arrive(&full[st]);
      __syncwarp();
      if (++st == p.S) {
        st = 0;
        ph ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

struct Case {
  const char* name;
  Params p;
};

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    g_encode = (EncodeTiled_t)fn;
    if (!g_encode) {
      printf("no cuTensorMapEncodeTiled\n");
      return 3;
    }
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  const int SMEM = 200 * 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));

  float* dD;
  long long* dcyc;
  int* derr;
  CK(cudaMalloc(&dD, 128 * 256 * sizeof(float)));
  CK(cudaMalloc(&dcyc, 1024 * sizeof(long long)));
  CK(cudaMalloc(&derr, sizeof(int)));

  srand(1234);
  std::vector<Case> cases;
  // ---- K-major ----
  for (int N : {32, 64, 128, 256})
    cases.push_back({"kmajor", {0, N, 128, 0, 0, 128, 0, 0}});
  for (int s : {1, 3, 8, 9, 17, 66, 67, 127})
    for (int bo : {0, 1}) cases.push_back({"kmajor_shift", {0, 64, 128, s, bo, 256, 0, 0}});
  // 8-row groups of A every `sbo` rows (SBO = sbo * 128 B): a (h+2) x (w+2) halo patch read in place, tap = row shift
  for (int sbo : {10, 12, 16, 9})
    for (int s : {0, 1, 11, 22}) cases.push_back({"kmajor_sbo", {0, 64, 128, s, 0, 256, sbo, 0}});
  // ---- MN-major ----
  for (int N : {32, 64, 128}) cases.push_back({"mnmajor_sw128", {1, N, 128, 0, 0, 128, 0, 0}});
  for (int s : {1, 3, 8, 9, 17, 67})
    for (int bo : {0, 1}) cases.push_back({"mnmajor_shift", {1, 64, 128, s, bo, 128 + 72, 0, 0}});
  cases.push_back({"mnmajor_B_sw64", {2, 32, 128, 0, 0, 128, 0, 0}});
  cases.push_back({"mnmajor_B_sw64_shift", {2, 32, 128, 5, 0, 136, 0, 0}});

  if (argc > 1 && (!strcmp(argv[1], "rate") || !strcmp(argv[1], "pipe"))) cases.clear();
  for (auto& cs : cases) {
    Params p = cs.p;
    // host data
    std::vector<float> hA, hB;
    std::vector<__nv_bfloat16> bA, bB;
    int a_rows = p.a_rows;
    CUtensorMap mA, mB;
    void *dA = nullptr, *dB = nullptr;
    std::vector<float> ref(128 * p.N, 0.f);
    if (p.mode == 0) {
      hA.resize((size_t)a_rows * p.K);
      hB.resize((size_t)p.N * p.K);
      for (auto& v : hA) v = bf((rand() % 2001 - 1000) / 1000.f);
      for (auto& v : hB) v = bf((rand() % 2001 - 1000) / 1000.f);
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < p.N; ++n) {
          double acc = 0;
          for (int k = 0; k < p.K; ++k)
            acc += (double)hA[(size_t)((p.reps ? (m / 8) * p.reps + (m % 8) : m) + p.shift) * p.K + k] * hB[(size_t)n * p.K + k];
          ref[m * p.N + n] = (float)acc;
        }
    } else {
      // X [a_rows][128], Y [K][Ncols] ; Ncols storage = 64-multiple for mode 1, 32 for mode 2
      int ncols = (p.mode == 1) ? ((p.N + 63) / 64) * 64 : 32;
      hA.resize((size_t)a_rows * 128);
      hB.resize((size_t)p.K * ncols);
      for (auto& v : hA) v = bf((rand() % 2001 - 1000) / 1000.f);
      for (auto& v : hB) v = bf((rand() % 2001 - 1000) / 1000.f);
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < p.N; ++n) {
          double acc = 0;
          for (int k = 0; k < p.K; ++k)
            acc += (double)hA[(size_t)(k + p.shift) * 128 + m] * hB[(size_t)k * ncols + n];
          ref[m * p.N + n] = (float)acc;
        }
    }
    bA.resize(hA.size());
    bB.resize(hB.size());
    for (size_t i = 0; i < hA.size(); ++i) bA[i] = __float2bfloat16(hA[i]);
    for (size_t i = 0; i < hB.size(); ++i) bB[i] = __float2bfloat16(hB[i]);
    CK(cudaMalloc(&dA, bA.size() * 2));
    CK(cudaMalloc(&dB, bB.size() * 2));
    CK(cudaMemcpy(dA, bA.data(), bA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, bB.data(), bB.size() * 2, cudaMemcpyHostToDevice));
    if (p.mode == 0) {
      // boxes limited to 256 rows; a_rows <= 264 -> clamp the box (rows beyond are never read when
      // shift + 128 <= box rows; the probe keeps shift + 128 <= a_rows <= 256+8, so clamp to 256)
      int box_rows = a_rows > 256 ? 256 : a_rows;
      if (p.shift + (p.reps ? 15 * p.reps + 8 : 128) > box_rows) {
        printf("%s: bad config\n", cs.name);
        continue;
      }
      p.a_rows = box_rows;
      make_map_2d(&mA, dA, p.K, a_rows, (uint64_t)p.K * 2, 64, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
      make_map_2d(&mB, dB, p.K, p.N, (uint64_t)p.K * 2, 64, p.N, CU_TENSOR_MAP_SWIZZLE_128B);
    } else if (p.mode == 1) {
      int ncols = ((p.N + 63) / 64) * 64;
      make_map_2d(&mA, dA, 128, a_rows, 256, 64, a_rows, CU_TENSOR_MAP_SWIZZLE_128B);
      make_map_2d(&mB, dB, ncols, p.K, (uint64_t)ncols * 2, 64, p.K, CU_TENSOR_MAP_SWIZZLE_128B);
    } else {
      make_map_2d(&mA, dA, 128, a_rows, 256, 64, a_rows, CU_TENSOR_MAP_SWIZZLE_128B);
      make_map_2d(&mB, dB, 32, p.K, 64, 32, p.K, CU_TENSOR_MAP_SWIZZLE_64B);
    }
    CK(cudaMemset(dD, 0, 128 * 256 * sizeof(float)));
    CK(cudaMemset(derr, 0, sizeof(int)));
    probe_kernel<<<1, 128, SMEM>>>(mA, mB, p, dD, dcyc, derr);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-22s N=%3d shift=%3d bo=%d : CUDA ERROR %s\n", cs.name, p.N, p.shift, p.base_off,
             cudaGetErrorString(e));
      return 4;
    }
    int herr = 0;
    CK(cudaMemcpy(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<float> out(128 * p.N);
    CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (size_t i = 0; i < out.size(); ++i) {
      maxerr = fmax(maxerr, fabs((double)out[i] - ref[i]));
      maxref = fmax(maxref, fabs((double)ref[i]));
    }
    if (p.mode == 0 && p.reps) printf("[sbo rows %d] ", p.reps);
    printf("%-22s N=%3d shift=%3d bo=%d : %s maxerr=%.4g (maxref %.3g) timeout=%d\n", cs.name, p.N,
           p.shift, p.base_off, (maxerr < 1e-2 * fmax(1.0, maxref) / 10 && !herr) ? "PASS" : "FAIL",
           maxerr, maxref, herr);
    cudaFree(dA);
    cudaFree(dB);
  }

  if (argc > 1 && !strcmp(argv[1], "pipe")) {
    CK(cudaFuncSetAttribute(pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    for (int threads : {352})
    for (int data : {1})
    for (int dstride : {128})
    for (int sep : {3})
    for (int commit_only : {1, 0})
      for (int N : {32, 64})
        for (int G : {4, 12, 36})
          for (int S : {2, 3, 4, 8}) {
            if (S != 3) continue;
            if (dstride < N) continue;
            PipeParams pp = {N, G, S, 1440 / G, 2, commit_only, sep, dstride, data};
            for (int it = 0; it < 2; ++it) {
              CK(cudaMemset(derr, 0, sizeof(int)));
              CK(cudaMemset(dcyc, 0, 1024 * sizeof(long long)));
              pipe_kernel<<<148, threads, SMEM>>>(pp, dcyc, derr);
              CK(cudaDeviceSynchronize());
            }
            std::vector<long long> cyc(148);
            CK(cudaMemcpy(cyc.data(), dcyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (auto c : cyc) mx = c > mx ? c : mx;
            int herr = 0;
            CK(cudaMemcpy(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost));
            printf("pipe data=%d dstride=%d threads=%d %s sep=%d N=%2d G=%2d stages=%d : %.1f cyc/MMA aggregate (2 issuers)%s\n",
                   data, dstride, threads, commit_only ? "commit-only" : "full/empty ", sep, N, G, S, (double)mx / (1440.0 * 2), herr ? " TIMEOUT" : "");
          }
    printf("done\n");
    return 0;
  }
  // ---- rate ----
  {
    CUtensorMap dummy;
    memset(&dummy, 0, sizeof(dummy));
    const bool shift_sweep = argc > 2 && !strcmp(argv[2], "shift");
    for (int shift : {0, 1, 2, 3, 5, 7})
    for (int nissue : {1, 2, 4})
      for (int N : {16, 32, 64, 128}) {
          if (!shift_sweep && shift) continue;
          if (shift_sweep && (nissue != 2 || N == 16)) continue;
          Params p = {3, N, (argc > 3 ? atoi(argv[3]) : 0), shift, 0, nissue, 4096, (argc > 4 ? atoi(argv[4]) : 0)};
          CK(cudaMemset(derr, 0, sizeof(int)));
          CK(cudaMemset(dcyc, 0, 1024 * sizeof(long long)));
          probe_kernel<<<148, 128, SMEM>>>(dummy, dummy, p, dD, dcyc, derr);
          CK(cudaDeviceSynchronize());
          CK(cudaMemset(dcyc, 0, 1024 * sizeof(long long)));
          probe_kernel<<<148, 128, SMEM>>>(dummy, dummy, p, dD, dcyc, derr);
          CK(cudaDeviceSynchronize());
          std::vector<long long> cyc(148);
          CK(cudaMemcpy(cyc.data(), dcyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
          long long mx = 0;
          for (auto c : cyc) mx = c > mx ? c : mx;
          double per = (double)mx / (p.reps * nissue);
          printf("rate issuers=%d N=%3d A-shift=%d rows : %.1f cyc/MMA aggregate (ideal %.1f)\n", nissue, N, shift, per, N / 2.0);
      }
  }
  printf("done\n");
  return 0;
}
