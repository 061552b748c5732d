// Host-side plumbing shared by all translation units of libtrainner_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/trainner_b200.h"

namespace b200 {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define B200_CHECK_CUDA(expr)                                                               \
  do {                                                                                      \
    cudaError_t e_ = (expr);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return ::b200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                               __LINE__);                                                   \
  } while (0)

#define B200_REQUIRE(cond, ...)                      \
  do {                                               \
    if (!(cond)) return ::b200::set_error(__VA_ARGS__); \
  } while (0)

// after a kernel launch
#define B200_LAUNCH_CHECK()                \
  do {                                     \
    ::b200::g_launches.fetch_add(1);       \
    B200_CHECK_CUDA(cudaGetLastError());   \
  } while (0)

typedef CUresult (*EncodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiled_t get_encode_tiled();
int sm_count();   // of the CURRENT device (cached per device)

// cudaFuncAttributeMaxDynamicSharedMemorySize (and optionally the max-shared carveout) is a per-device
// attribute of a kernel: set once per (kernel, device) under a mutex, so that a process that drives
// several GPUs (nn.DataParallel / gpu_ids lists of the reference) never launches with a stale limit.
int ensure_max_smem(const void* kernel, int bytes, bool prefer_max_carveout = false);
#define B200_ENSURE_SMEM(kernel, bytes) \
  do { if (::b200::ensure_max_smem(reinterpret_cast<const void*>(kernel), (bytes))) return 1; } while (0)

// bf16 tensor map, rank <= 5, dims/strides innermost first (strides in bytes, rank-1 entries).
int make_tensor_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                    CUtensorMapSwizzle swizzle);

// Deterministic cross-block reductions ("last block sums the partials in block order"): a small library-owned
// arena per device (64 MB of fp32 partials + 8192 arrival counters, allocated on first use -- never during
// stream capture) shared by the reduction kernels of the library, which run one after another on a stream
// (every kernel waits for its predecessor; concurrent use from several streams is not supported).
// No float atomics anywhere: two runs on the same inputs give bit-identical results.
struct DetScratch {
  float* part;
  unsigned* counters;
};
constexpr size_t kDetFloats = 16u << 20;
constexpr int kDetCounters = 8192;
int det_scratch(DetScratch* out, size_t floats_needed, int counters_needed);

inline cudaStream_t as_stream(b200_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch: every kernel of the library is launched with the
// programmatic-stream-serialization attribute, calls pdl_trigger() on entry and pdl_wait() before its
// first access to global memory, so that the launch latency and prologue (barrier init, TMEM
// allocation, descriptor prefetch) of kernel N+1 overlap the tail of kernel N.  griddepcontrol.wait
// returns only when the preceding grid has completed and flushed, so ordering is unchanged; because
// every kernel waits, completion stays transitive along the stream.  B200_PDL=0 turns the attribute off.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Block-uniform: true in the LAST of `nblk` blocks to arrive on `counter`.  Partials written (by any thread of
// the calling block) before the call are visible to the last block after it; the counter resets itself.
__device__ __forceinline__ bool det_arrive_last(unsigned* counter, unsigned nblk) {
  __shared__ unsigned s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
    const unsigned t = atomicAdd(counter, 1u);
    s_last = (t == nblk - 1) ? 1u : 0u;
    if (s_last) *counter = 0u;
  }
  __syncthreads();
  const bool last = s_last != 0u;
  if (last) __threadfence();
  return last;
}

// Deterministic cross-block sum in two levels (a single last block adding nblk x nout partials would be a
// multi-microsecond serial tail): the blocks form groups of kDetGroup; the last block of a group to arrive adds the
// group's partials (contiguous in memory: part is OUTPUT-major, part[k * nblk + b]) in block order into
// gpart[k * ngrp + g]; the last GROUP to finish adds the group sums in group order and calls out(k, total);
// returns true (block-uniform) in that one block.
// counters: 1 + ngrp words (self-resetting).  Every block of the grid must call it (blockDim.x = 256).
constexpr int kDetGroup = 16;
__host__ __device__ __forceinline__ int det_groups(int nblk) { return (nblk + kDetGroup - 1) / kDetGroup; }

template <typename F>
__device__ __forceinline__ bool det_reduce(const float* __restrict__ part, float* __restrict__ gpart,
                                           unsigned* counters, unsigned nblk, int nout, F&& out) {
  const unsigned grp = blockIdx.x / kDetGroup, ngrp = det_groups((int)nblk);
  const unsigned g0 = grp * kDetGroup;
  const unsigned members = (nblk - g0) < (unsigned)kDetGroup ? (nblk - g0) : (unsigned)kDetGroup;
  if (!det_arrive_last(counters + 1 + grp, members)) return false;
  for (int k = threadIdx.x; k < nout; k += blockDim.x) {
    const float* row = part + (size_t)k * nblk + g0;
    float t = 0.f;
#pragma unroll 4
    for (unsigned i = 0; i < members; ++i) t += row[i];
    gpart[(size_t)k * ngrp + grp] = t;
  }
  if (!det_arrive_last(counters, ngrp)) return false;
  for (int k = threadIdx.x; k < nout; k += blockDim.x) {
    const float* row = gpart + (size_t)k * ngrp;
    float t = 0.f;
#pragma unroll 4
    for (unsigned g = 0; g < ngrp; ++g) t += row[g];
    out(k, t);
  }
  return true;   // block-uniform: this block produced the totals
}
#endif

}  // namespace b200
