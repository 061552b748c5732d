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

}  // namespace b200
