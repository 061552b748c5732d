#include "common.cuh"

#include <map>
#include <mutex>
#include <utility>
#include <stdlib.h>

namespace b200 {

thread_local char g_err[512] = {0};
std::atomic<long long> g_launches{0};

EncodeTiled_t get_encode_tiled() {
  static EncodeTiled_t fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) ==
            cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiled_t>(p);
  });
  return fn;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("B200_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

int sm_count() {
  static std::mutex mu;
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::lock_guard<std::mutex> g(mu);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cache[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

int ensure_max_smem(const void* kernel, int bytes, bool prefer_max_carveout) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
  std::lock_guard<std::mutex> g(mu);
  auto key = std::make_pair(kernel, dev);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return 0;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && prefer_max_carveout)
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess)
    return set_error("cudaFuncSetAttribute(max dynamic smem %d) failed on device %d: %s", bytes, dev,
                     cudaGetErrorString(e));
  done[key] = bytes;
  return 0;
}

int det_scratch(DetScratch* out, size_t floats_needed, int counters_needed) {
  static std::mutex mu;
  static DetScratch per_dev[64] = {};
  if (floats_needed > kDetFloats || counters_needed > kDetCounters)
    return set_error("deterministic-reduction scratch too small: need %zu floats / %d counters", floats_needed,
                     counters_needed);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) return set_error("det_scratch: cudaGetDevice failed");
  std::lock_guard<std::mutex> g(mu);
  if (!per_dev[dev].part) {
    void* p = nullptr;
    e = cudaMalloc(&p, kDetFloats * sizeof(float) + kDetCounters * sizeof(unsigned));
    if (e != cudaSuccess)
      return set_error("det_scratch: cudaMalloc failed (%s) -- the first call of a reduction kernel must not happen "
                       "inside a stream capture", cudaGetErrorString(e));
    e = cudaMemset(static_cast<float*>(p) + kDetFloats, 0, kDetCounters * sizeof(unsigned));
    if (e != cudaSuccess) return set_error("det_scratch: cudaMemset failed (%s)", cudaGetErrorString(e));
    per_dev[dev].part = static_cast<float*>(p);
    per_dev[dev].counters = reinterpret_cast<unsigned*>(static_cast<float*>(p) + kDetFloats);
  }
  *out = per_dev[dev];
  return 0;
}

int make_tensor_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box,
                    const uint32_t* elem_strides, CUtensorMapSwizzle swizzle) {
  EncodeTiled_t enc = get_encode_tiled();
  if (!enc) return set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    e[i] = elem_strides ? elem_strides[i] : 1;
    if (i < rank - 1) s[i] = strides_bytes[i];
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), d, s, b, e,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(
        "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u] "
        "stride0 %llu ptr %p",
        (int)r, rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0),
        (unsigned long long)(rank > 2 ? d[2] : 0), (unsigned long long)(rank > 3 ? d[3] : 0), b[0],
        rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0, rank > 3 ? b[3] : 0,
        (unsigned long long)(rank > 1 ? s[0] : 0), ptr);
  }
  return 0;
}

}  // namespace b200

extern "C" {

const char* b200_last_error(void) { return b200::g_err; }
int b200_version(void) { return 100; }
int64_t b200_launch_count(void) { return b200::g_launches.load(); }
int b200_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}
}
