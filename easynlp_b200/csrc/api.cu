// Library-level state of the clipk C ABI: error string, launch counter, TMA descriptor encoding.
#include <stdarg.h>
#include <atomic>
#include "common.cuh"
#include "../../include/clipk.h"

namespace clipk {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return CLIPK_ERR_CUDA;
}
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    // resolved through the runtime so that the library does not link libcuda (absent on build hosts)
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor [outer, inner] (inner contiguous, row stride ld_elems), SWIZZLE_128B boxes, zero OOB fill.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return CLIPK_ERR_CUDA; }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld_elems % 8)) {
    set_error("TMA operand must be 16-byte aligned with ld %% 8 == 0 (base=%p ld=%llu)", base, (unsigned long long)ld_elems);
    return CLIPK_ERR_ARG;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: %d (inner=%llu outer=%llu ld=%llu box=%ux%u)", (int)r, (unsigned long long)inner,
              (unsigned long long)outer, (unsigned long long)ld_elems, box_inner, box_outer);
    return CLIPK_ERR_CUDA;
  }
  return 0;
}

DropArg make_drop_arg(const void* ptr) {
  DropArg a{};
  const clipk_dropout_t* d = reinterpret_cast<const clipk_dropout_t*>(ptr);
  if (!d || !(d->p > 0.f)) return a;
  float p = d->p > 0.999f ? 0.999f : d->p;
  a.on = 1;
  a.k0 = (uint32_t)(d->seed & 0xffffffffull); a.k1 = (uint32_t)(d->seed >> 32);
  a.site = d->site;
  a.thresh = (uint32_t)((double)p * 4294967296.0);
  a.scale = 1.0f / (1.0f - p);
  a.dev_offset = d->dev_offset;
  return a;
}

}  // namespace clipk

extern "C" const char* clipk_last_error(void) { return clipk::g_err; }
extern "C" int clipk_version(void) { return 100; }
extern "C" int64_t clipk_launch_count(void) { return clipk::g_launches.load(); }
