// Host-side helpers shared by the .cu translation units of libasyrp_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#define ASYRP_API __attribute__((visibility("default")))

namespace asyrp {

// Error codes returned through the C ABI (0 = success). Mirrors include/asyrp_b200.h.
enum : int {
  ASYRP_OK = 0,
  ASYRP_ERR_INVALID = -1,   // bad argument / unsupported shape
  ASYRP_ERR_CUDA = -2,      // CUDA runtime / driver error
  ASYRP_ERR_NO_DEVICE = -3, // no sm_100 device available
};

void set_error(const char* fmt, ...);
const char* get_error();

#define ASYRP_CHECK_CUDA(expr)                                                                   \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      ::asyrp::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return ::asyrp::ASYRP_ERR_CUDA;                                                            \
    }                                                                                            \
  } while (0)

#define ASYRP_REQUIRE(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::asyrp::set_error(__VA_ARGS__);      \
      return ::asyrp::ASYRP_ERR_INVALID;    \
    }                                       \
  } while (0)

// cuTensorMapEncodeTiled resolved at run time (no link-time dependency on libcuda, so the library loads
// on a machine without a driver and simply fails loudly when an op is created).
int encode_tensor_map(CUtensorMap* out, CUtensorMapDataType dt, uint32_t rank, const void* gaddr,
                      const uint64_t* dims, const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box,
                      CUtensorMapSwizzle swz);

int sm_count();

}  // namespace asyrp
