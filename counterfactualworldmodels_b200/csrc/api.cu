// api.cu -- library-level entry points, error string, TMA descriptor factory.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <cstdlib>

#include "common.cuh"

namespace cwm {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;
static thread_local long long g_total_launches = 0;

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

void count_launch() { ++g_launches; ++g_total_launches; }
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

int make_tmap_2d(CUtensorMap* map, const void* base, int dtype, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols, int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(CWM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const uint64_t elt = (dtype == CWM_TMAP_F32) ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * elt) % 16 != 0)
    return fail(CWM_ERR_INVALID, "TMA operand must be 16-byte aligned (base %p, ld %llu)", base,
                (unsigned long long)ld);
  if (box_cols * elt > static_cast<uint64_t>(swizzle_bytes) || box_rows > 256 || (swizzle_bytes != 128 && swizzle_bytes != 64))
    return fail(CWM_ERR_INVALID, "TMA box [%u, %u] too large for %dB swizzle", box_rows, box_cols, swizzle_bytes);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * elt};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dtype == CWM_TMAP_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(CWM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=[%u,%u]", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
  return CWM_OK;
}

int make_tmap_3d_f32(CUtensorMap* map, const void* base, uint64_t S, uint64_t R, uint64_t C, uint32_t box_rows,
                     uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(CWM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (C * 4) % 16 != 0 || box_cols * 4 > 128 || box_rows > 256)
    return fail(CWM_ERR_INVALID, "TMA operand (3-D fp32) misaligned or box too large (base %p, C %llu, box [%u, %u])", base,
                (unsigned long long)C, box_rows, box_cols);
  cuuint64_t gdim[3] = {C, R, S};
  cuuint64_t gstride[2] = {C * 4, R * C * 4};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(CWM_ERR_CUDA, "cuTensorMapEncodeTiled (3-D fp32) failed (%d) S=%llu R=%llu C=%llu", (int)r,
                (unsigned long long)S, (unsigned long long)R, (unsigned long long)C);
  return CWM_OK;
}

int make_tmap_nhwc(CUtensorMap* map, const void* base, uint64_t S, uint64_t H, uint64_t W, uint64_t C, uint64_t ld,
                   uint32_t box_h, uint32_t box_w, uint32_t box_c, uint32_t pixel_stride) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(CWM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0 || ld < C)
    return fail(CWM_ERR_INVALID, "TMA operand must be 16-byte aligned with ld >= C (base %p, ld %llu, C %llu)", base,
                (unsigned long long)ld, (unsigned long long)C);
  if (box_c * 2 > 128 || box_w > 256 || box_h > 256)
    return fail(CWM_ERR_INVALID, "TMA box [%u, %u, %u] too large", box_h, box_w, box_c);
  cuuint64_t gdim[4] = {C, W, H, S};
  cuuint64_t gstride[3] = {ld * 2, W * ld * 2, H * W * ld * 2};
  cuuint32_t box[4] = {box_c, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, pixel_stride, pixel_stride, 1};  // traversal stride: a box of b pixels loads ceil(b / stride)
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(CWM_ERR_CUDA, "cuTensorMapEncodeTiled (NHWC) failed (%d) S=%llu H=%llu W=%llu C=%llu ld=%llu box=[%u,%u,%u]",
                (int)r, (unsigned long long)S, (unsigned long long)H, (unsigned long long)W, (unsigned long long)C,
                (unsigned long long)ld, box_h, box_w, box_c);
  return CWM_OK;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* v = getenv("CWM_PDL");
    on = (v == nullptr) ? 1 : atoi(v);
  }
  return on != 0;
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

// ---- profiling ---------------------------------------------------------------------------------------
struct ProfRec {
  const char* name;
  double flops, bytes;
  cudaEvent_t e0, e1;
};
static bool g_prof_on = false;
static std::vector<ProfRec*> g_prof;

ProfileScope::ProfileScope(cudaStream_t s, const char* name, double flops, double bytes) : stream(s) {
  if (!g_prof_on) return;
  ProfRec* r = new ProfRec{name, flops, bytes, nullptr, nullptr};
  cudaEventCreate(&r->e0);
  cudaEventCreate(&r->e1);
  cudaEventRecord(r->e0, s);
  rec = r;
}
ProfileScope::~ProfileScope() {
  if (!rec) return;
  ProfRec* r = static_cast<ProfRec*>(rec);
  cudaEventRecord(r->e1, stream);
  g_prof.push_back(r);
}

}  // namespace cwm

extern "C" {

int cwm_profile_begin(void) {
  for (auto* r : cwm::g_prof) { cudaEventDestroy(r->e0); cudaEventDestroy(r->e1); delete r; }
  cwm::g_prof.clear();
  cwm::g_prof_on = true;
  return CWM_OK;
}

int cwm_profile_end(cwm_profile_entry* out, int max_entries, int* n_entries) {
  cwm::g_prof_on = false;
  CWM_REQUIRE(out && n_entries && max_entries > 0, "cwm_profile_end: null pointer");
  std::map<std::string, cwm_profile_entry> agg;
  std::vector<std::string> order;
  for (auto* r : cwm::g_prof) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(r->e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r->e0, r->e1);
    if (e != cudaSuccess) return cwm::fail(CWM_ERR_CUDA, "cwm_profile_end: %s", cudaGetErrorString(e));
    auto it = agg.find(r->name);
    if (it == agg.end()) {
      cwm_profile_entry en;
      memset(&en, 0, sizeof(en));
      strncpy(en.name, r->name, sizeof(en.name) - 1);
      it = agg.emplace(r->name, en).first;
      order.push_back(r->name);
    }
    it->second.launches += 1;
    it->second.ms += ms;
    it->second.flops += r->flops;
    it->second.bytes += r->bytes;
    cudaEventDestroy(r->e0);
    cudaEventDestroy(r->e1);
    delete r;
  }
  cwm::g_prof.clear();
  int n = 0;
  for (auto& k : order) {
    if (n >= max_entries) break;
    out[n++] = agg[k];
  }
  *n_entries = n;
  return CWM_OK;
}

int cwm_abi_version(void) { return CWM_B200_ABI_VERSION; }
int cwm_act_dtype(void) { return CWM_ACT_IS_BF16; }

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

long long cwm_total_launches(void) { return cwm::g_total_launches; }

int cwm_launch_count_reset(void) {
  cwm::g_launches = 0;
  return CWM_OK;
}

}  // extern "C"
