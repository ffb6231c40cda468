#include "host_common.h"

#include <stdlib.h>

#include <string.h>

#include "../../include/speechmix_sm100.h"

namespace smx {

static thread_local char g_err[1024] = {0};

char* err_buf() { return g_err; }

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// A tensor map is a pure function of (base, type, rank, dims, strides, box, swizzle): nothing in it depends on what
// the memory holds.  The caching allocator hands the same addresses back step after step, so an eager training step
// re-encodes the same ~1500 descriptors every time (up to five per GEMM launch, ~1 us each through the driver).
// Direct-mapped, thread-local cache keyed on the full argument tuple; a hit is one hash + one 128-byte compare.
struct TmapKey {
  const void* ptr;
  uint64_t dims[5], strides[4];
  uint32_t box[5];
  int32_t dt, rank, swz;
};
struct TmapSlot {
  TmapKey key;
  CUtensorMap map;
  bool used;
};
constexpr int TMAP_SLOTS = 8192;
static thread_local TmapSlot* g_tmap_cache = nullptr;

static int encode_uncached(CUtensorMap* out, CUtensorMapDataType dt, int esize, const void* ptr, int rank,
                           const uint64_t* dims, const uint64_t* strides_elems, const uint32_t* box, bool swizzle128);

static int encode_generic(CUtensorMap* out, CUtensorMapDataType dt, int esize, const void* ptr, int rank,
                          const uint64_t* dims, const uint64_t* strides_elems, const uint32_t* box,
                          bool swizzle128) {
  if (rank < 1 || rank > 5) return set_error("tensor map rank %d unsupported", rank);
  TmapKey k;
  memset(&k, 0, sizeof(k));
  k.ptr = ptr, k.dt = (int32_t)dt, k.rank = rank, k.swz = swizzle128 ? 1 : 0;
  for (int i = 0; i < rank; ++i) {
    k.dims[i] = dims[i], k.box[i] = box[i];
    if (i > 0) k.strides[i - 1] = strides_elems[i - 1];
  }
  uint64_t h = 1469598103934665603ull;
  const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
  for (size_t i = 0; i < sizeof(k) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
  if (!g_tmap_cache) g_tmap_cache = static_cast<TmapSlot*>(calloc(TMAP_SLOTS, sizeof(TmapSlot)));
  TmapSlot* slot = g_tmap_cache ? &g_tmap_cache[(h >> 17) & (TMAP_SLOTS - 1)] : nullptr;
  if (slot && slot->used && memcmp(&slot->key, &k, sizeof(k)) == 0) {
    memcpy(out, &slot->map, sizeof(CUtensorMap));
    return 0;
  }
  const int rc = encode_uncached(out, dt, esize, ptr, rank, dims, strides_elems, box, swizzle128);
  if (rc == 0 && slot) {
    slot->key = k;
    memcpy(&slot->map, out, sizeof(CUtensorMap));
    slot->used = true;
  }
  return rc;
}

static int encode_uncached(CUtensorMap* out, CUtensorMapDataType dt, int esize, const void* ptr, int rank,
                           const uint64_t* dims, const uint64_t* strides_elems, const uint32_t* box,
                           bool swizzle128) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstrides[i - 1] = strides_elems[i - 1] * (uint64_t)esize;
      if (gstrides[i - 1] % 16 != 0)
        return set_error("tensor map stride %d (%llu bytes) not a multiple of 16", i,
                         (unsigned long long)gstrides[i - 1]);
    }
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return set_error("tensor map base not 16-byte aligned");
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(ptr), gdims, gstrides, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error("cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]",
                     (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                     box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  }
  return 0;
}

int encode_tmap_bf16(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims,
                     const uint64_t* strides_elems, const uint32_t* box, bool swizzle128) {
  return encode_generic(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, rank, dims, strides_elems, box, swizzle128);
}
int encode_tmap_f32(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims,
                    const uint64_t* strides_elems, const uint32_t* box) {
  return encode_generic(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ptr, rank, dims, strides_elems, box, false);
}

int num_sms() {
  static int n = 0;
  if (n) return n;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  return n;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SMX_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

}  // namespace smx

extern "C" {

const char* smx_last_error(void) { return smx::err_buf(); }
int smx_abi_version(void) { return SMX_ABI_VERSION; }
int smx_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return smx::set_error("no CUDA device");
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return smx::set_error("cannot query device");
  return major == 10 ? 1 : 0;
}
}
