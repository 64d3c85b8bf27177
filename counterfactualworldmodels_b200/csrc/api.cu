// api.cu -- library-level entry points, error string, TMA descriptor factory.
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "common.cuh"

namespace cwm {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch() { ++g_launches; }
int launches() { return g_launches; }
void reset_launches() { g_launches = 0; }

// cuTensorMapEncodeTiled resolved through the runtime so the library has no link-time libcuda dependency.
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
  });
  return fn;
}

int make_tmap_2d_f16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                     uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(CWM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0)
    return fail(CWM_ERR_INVALID, "TMA operand must be 16-byte aligned (base %p, ld %llu)", base,
                (unsigned long long)ld);
  if (box_cols * 2 > 128 || box_rows > 256)
    return fail(CWM_ERR_INVALID, "TMA box [%u, %u] too large for 128B swizzle", box_rows, box_cols);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(CWM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=[%u,%u]", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
  return CWM_OK;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace cwm

extern "C" {

int cwm_abi_version(void) { return CWM_B200_ABI_VERSION; }

const char* cwm_last_error(void) { return cwm::g_err; }

int cwm_device_check(void) {
  int dev = 0;
  CWM_CUDA_CHECK(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  CWM_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CWM_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) return cwm::fail(CWM_ERR_ARCH, "device %d is sm_%d%d; libcwm_b200 is built for sm_100a only", dev, major, minor);
  return CWM_OK;
}

int cwm_last_forward_launches(void) { return cwm::g_launches; }

}  // extern "C"
