// Host-side helpers shared by the C-ABI translation units: thread-local error
// string, CUDA error checks and TMA tensor-map encoding (driver entry point is
// resolved at run time so the library links without libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

namespace smx {

char* err_buf();
int set_error(const char* fmt, ...);

#define SMX_CHECK_CUDA(expr)                                                                 \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ::smx::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define SMX_REQUIRE(cond, ...)                         \
  do {                                                 \
    if (!(cond)) return ::smx::set_error(__VA_ARGS__); \
  } while (0)

// Encode a bf16 tensor map with up to 4 dims.  dims[0] is the contiguous dim;
// strides_elems[i] is the stride of dims[i+1] in elements.  box[] in elements.
// swizzle128: 128-byte swizzle (box[0] must be <= 64 elements) else none.
int encode_tmap_bf16(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims,
                     const uint64_t* strides_elems, const uint32_t* box, bool swizzle128);
int encode_tmap_f32(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims,
                    const uint64_t* strides_elems, const uint32_t* box);

int num_sms();

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace smx
