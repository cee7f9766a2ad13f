// Process-wide support for the C ABI: latched error state, device property cache and
// the driver entry point for tensor-map encoding (resolved through the CUDA runtime so
// the library has no link-time dependency on libcuda).
#include <atomic>
#include <mutex>
#include <string>

#include "../../include/b200_vllm.h"
#include "common.cuh"

namespace b200 {

static std::mutex g_err_mu;
static int g_err_code = 0;
static std::string g_err_msg;

void set_error(int code, const char* msg) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  if (g_err_code == 0) {   // keep the first failure
    g_err_code = code;
    g_err_msg = msg ? msg : "";
  }
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;   // B200
  }
  return cached;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

CUresult encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, uint32_t rank, void* gaddr,
                           const uint64_t* dims, const uint64_t* strides, const uint32_t* box,
                           const uint32_t* estrides, CUtensorMapSwizzle swz,
                           CUtensorMapL2promotion l2) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
            cudaSuccess ||
        p == nullptr)
      return CUDA_ERROR_NOT_SUPPORTED;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn(map, dt, rank, gaddr, reinterpret_cast<const cuuint64_t*>(dims),
            reinterpret_cast<const cuuint64_t*>(strides), box, estrides,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace b200

extern "C" {
int b200_last_error(void) {
  std::lock_guard<std::mutex> lk(b200::g_err_mu);
  return b200::g_err_code;
}
const char* b200_last_error_string(void) {
  std::lock_guard<std::mutex> lk(b200::g_err_mu);
  static thread_local std::string copy;
  copy = b200::g_err_msg;
  return copy.c_str();
}
void b200_clear_error(void) {
  std::lock_guard<std::mutex> lk(b200::g_err_mu);
  b200::g_err_code = 0;
  b200::g_err_msg.clear();
}
int b200_abi_version(void) { return 1; }
int64_t b200_launch_count(void) { return b200::launch_count(); }
}
