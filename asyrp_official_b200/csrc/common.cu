// Error string, driver entry points and device queries shared by libasyrp_b200.so.
#include "common.h"
#include <cstdlib>
#include <mutex>

namespace asyrp {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled resolve_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int encode_tensor_map(CUtensorMap* out, CUtensorMapDataType dt, uint32_t rank, const void* gaddr,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMapSwizzle swz) {
  PFN_encodeTiled fn = resolve_encode();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return ASYRP_ERR_NO_DEVICE;
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (uint32_t i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gs[i] = strides_bytes[i];
  }
  CUresult r = fn(out, dt, rank, const_cast<void*>(gaddr), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %u dims %llu,%llu,%llu box %u,%u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
              box[0], box[1], rank > 2 ? box[2] : 0);
    return ASYRP_ERR_CUDA;
  }
  return ASYRP_OK;
}

void set_pdl(int on);

int sm_count() {
  static int n = -1;
  if (n < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    n = v;
  }
  return n;
}

static int g_pdl = -1;
int pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("ASYRP_PDL");
    g_pdl = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return g_pdl;
}
void set_pdl(int on) { g_pdl = on ? 1 : 0; }

}  // namespace asyrp

extern "C" ASYRP_API const char* asyrp_last_error(void) { return asyrp::get_error(); }
extern "C" ASYRP_API int asyrp_set_pdl(int enabled) {
  asyrp::set_pdl(enabled);
  return asyrp::ASYRP_OK;
}
extern "C" ASYRP_API int asyrp_get_pdl(void) { return asyrp::pdl_enabled(); }
