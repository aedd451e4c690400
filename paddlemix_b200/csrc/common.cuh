// Host-side helpers shared by the C-ABI translation units: error channel, driver entry points, tensor maps.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/b200mix.h"

namespace b200 {

void set_error(const char* fmt, ...);
int ensure_device();  // 0 if an sm_100 device is current, else negative status (with message)
int num_sms();

#define B200_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::b200::set_error(__VA_ARGS__);             \
      return B200MIX_ERR_INVALID;                 \
    }                                             \
  } while (0)

#define B200_CUDA(call)                                                                              \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      ::b200::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B200MIX_ERR_CUDA;                                                                       \
    }                                                                                                \
  } while (0)

#define B200_LAUNCH_CHECK()                                                                       \
  do {                                                                                            \
    cudaError_t _e = cudaGetLastError();                                                          \
    if (_e != cudaSuccess) {                                                                      \
      ::b200::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B200MIX_ERR_CUDA;                                                                    \
    }                                                                                             \
  } while (0)

// Encode a bf16 tiled tensor map with 128-byte swizzle and zero OOB fill.
// dims/box have `rank` entries (innermost first); strides_bytes has rank-1 entries (for dims 1..rank-1).
int encode_tmap_bf16_sw128(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box);

}  // namespace b200
