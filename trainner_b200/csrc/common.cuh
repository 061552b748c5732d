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
int sm_count();

// bf16 tensor map, rank <= 5, dims/strides innermost first (strides in bytes, rank-1 entries).
int make_tensor_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                    CUtensorMapSwizzle swizzle);

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
#endif

}  // namespace b200
