// Shared device helpers for the clipk kernels (sm_100a only): mbarrier, TMA, tcgen05/TMEM PTX wrappers,
// UMMA shared-memory / instruction descriptors, small math helpers.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#ifndef CLIPK_HANG_TRAP
#define CLIPK_HANG_TRAP 1   // bounded mbarrier spins: a pipeline deadlock traps instead of hanging the GPU
#endif

namespace clipk {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------ host error state
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define CLIPK_CUDA(call)                                         \
  do {                                                           \
    cudaError_t _e = (call);                                     \
    if (_e != cudaSuccess) return clipk::cuda_fail(_e, #call);   \
  } while (0)

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer, int swizzle_bytes = 128);
int sm_count();
void note_launch(int n = 1);
struct DropArg;
DropArg make_drop_arg(const void* clipk_dropout);   // NULL / p <= 0 -> off

// ------------------------------------------------------------------------------------------ misc device
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t.reg .b32 R;\n\t"
      "elect.sync R|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if CLIPK_HANG_TRAP
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("clipk: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}

// ------------------------------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// 2D tiled load global -> shared, completion on an mbarrier (bytes). c0 = inner coordinate, c1 = outer.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 2D tiled store shared -> global (bulk async group of the issuing thread); out-of-bounds parts of the box are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until all but the newest N bulk groups of this thread have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// generic-proxy writes to smem -> visible to the async proxy (UMMA / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// warp-collective: lane i reads 32-bit columns [col, col+N) of TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// ------------------------------------------------------------------------------------------ CTA pairs (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-SM TMA load: data lands in THIS CTA's smem, completion bytes are signalled on the mbarrier at `bar_cluster_addr` (the leader's)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_holder, uint32_t ncols) {  // one full warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, 256 rows over the CTA pair] (+)= A * B ; issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once per CTA in `cta_mask`) on the mbarrier at this smem offset when all prior MMAs of the issuing thread are done
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ------------------------------------------------------------------------------------------ UMMA descriptors
// Instruction descriptor, kind::f16, bf16 x bf16 -> fp32 (cute/arch/mma_sm100_desc.hpp InstrDescriptor bit layout):
//   [4,6) c_format=1 (F32)  [7,10) a_format=1 (BF16)  [10,13) b_format=1  [15] a_major  [16] b_major (0 = K-major, 1 = MN-major)
//   [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Shared-memory matrix descriptor, SWIZZLE_128B (layout_type 2 at [61,64)), version 1 at [46,48).
//   start address >> 4 at [0,14), leading byte offset >> 4 at [16,30), stride byte offset >> 4 at [32,46).
// K-major operand  (tile = rows x 64 bf16, 128 B per row, 8-row swizzle atoms): SBO = 1024 B (next 8-row group), LBO unused.
// MN-major operand (tile = 64-wide MN blocks; block = k rows x 128 B): SBO = 1024 B (next 8 k-rows), LBO = block stride.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// The same descriptor as two 32-bit halves.  The MMA-issuing thread is a single thread whose instruction stream is mostly descriptor
// arithmetic: rebuilding the 64-bit descriptor (shift / mask / or) for every tcgen05.mma costs ~20 dependent integer instructions per
// MMA -- more than the 32 cycles an M128 x N64 x K16 MMA takes (attention timelines: 23 MMAs issued in ~2300 cycles).  With the halves
// kept apart, the next k-step's descriptor is ONE add on the low word: lo + (byte offset >> 4) (start addresses stay below 2^14 * 16 B).
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__host__ __device__ constexpr uint32_t umma_desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ void umma_bf16_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  constexpr uint32_t hi = umma_desc_hi(1024);
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(hi)
      : "memory");
}
// A operand from tensor memory
__device__ __forceinline__ void umma_bf16_ts_lh(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  constexpr uint32_t hi = umma_desc_hi(1024);
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(hi)
      : "memory");
}

// byte offset of element (row r, 16-byte chunk c) inside a [rows x 128 B] SWIZZLE_128B tile whose base is 1024-B aligned
__device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t chunk16) { return r * 128u + ((chunk16 ^ (r & 7u)) << 4); }

// ------------------------------------------------------------------------------------------ dropout (Philox4x32-10)
// Counter-based RNG: the keep/drop decision of an element depends only on (seed, per-step device offset, site, element index), so
// the backward kernels regenerate the forward masks instead of storing them.  One call yields 4 independent 32-bit draws.
// (__host__ too: tests/test_philox_host.py runs the Random123 known-answer vectors through this very function on the CPU)
__host__ __device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    const uint32_t hi0 = (uint32_t)(((uint64_t)0xD2511F53u * c0) >> 32), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = (uint32_t)(((uint64_t)0xCD9E8D57u * c2) >> 32), lo1 = 0xCD9E8D57u * c2;
#endif
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
struct DropCtx {       // resolved on the device from clipk_dropout_t
  uint32_t k0, k1, site, thresh;
  float scale;         // 1 / (1 - p)
  bool on;
};
struct DropArg {       // built on the host (make_drop_arg) and passed by value to kernels
  uint32_t k0, k1, site, thresh;
  float scale;
  const unsigned int* dev_offset;
  int on;
};
__device__ __forceinline__ DropCtx drop_ctx(const DropArg& a) {
  DropCtx d;
  d.k0 = a.k0 ^ (a.dev_offset ? *a.dev_offset : 0u); d.k1 = a.k1; d.site = a.site; d.thresh = a.thresh; d.scale = a.scale; d.on = a.on != 0;
  return d;
}
// multipliers (0 or 1/(1-p)) for the 4 consecutive elements starting at element index 4*quad of row `row`
__device__ __forceinline__ float4 drop_mult4(const DropCtx& d, uint32_t row, uint32_t quad) {
  const uint4 r = philox4x32(row, quad, d.site, 0x2545F491u, d.k0, d.k1);
  return make_float4(r.x >= d.thresh ? d.scale : 0.f, r.y >= d.thresh ? d.scale : 0.f, r.z >= d.thresh ? d.scale : 0.f,
                     r.w >= d.thresh ? d.scale : 0.f);
}

// ------------------------------------------------------------------------------------------ math
__device__ __forceinline__ float quick_gelu_f(float x) { return x / (1.0f + __expf(-1.702f * x)); }
__device__ __forceinline__ float quick_gelu_grad_f(float x) {
  float s = 1.0f / (1.0f + __expf(-1.702f * x));
  return s * (1.0f + 1.702f * x * (1.0f - s));
}
__device__ __forceinline__ float erf_gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float erf_gelu_grad_f(float x) {
  float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// activation and its derivative from one pass over z (the forward GEMM epilogue saves both).  The K = 768 GEMM leaves ~6 k issue
// cycles per 128 x 256 tile, so these are written for instruction count: one MUFU (+ one for erf) and a handful of FMAs per element.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// QuickGELU x * sigmoid(1.702 x) (modeling_chineseclip.py:179-181): sigmoid(a) = 0.5 + 0.5 tanh(a / 2)
__device__ __forceinline__ void quick_gelu_both(float x, float& act, float& grad) {
  const float s = fmaf(tanh_approx(0.851f * x), 0.5f, 0.5f);
  act = x * s;
  grad = fmaf(s, 1.702f * (x - act), s);          // s * (1 + 1.702 x (1 - s))
}
// exact (erf) GELU (modelzoo/activations.py:45-48).  erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7), whose exp(-u^2) with
// u = x / sqrt(2) is at the same time the Gaussian density needed by the derivative.
__device__ __forceinline__ void erf_gelu_both(float x, float& act, float& grad) {
  const float u = 0.70710678118654752f * x;
  const float au = fabsf(u);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, au, 1.0f));
  const float ex = __expf(-u * u);
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, ex, 1.0f);
  const float cdf = fmaf(copysignf(erf_abs, u), 0.5f, 0.5f);
  act = x * cdf;
  grad = fmaf(x * 0.39894228040143268f, ex, cdf);
}
// column sums over the 32 rows (lanes) of a chunk held as 32 column values per lane: butterfly transpose-reduce, lane j ends with column j
__device__ __forceinline__ float colsum32(const float* v, int lane) {
  float w[16];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int k = 0; k < 16; ++k) { const float send = up ? v[k] : v[k + 16]; const float keep = up ? v[k + 16] : v[k]; w[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
  }
#pragma unroll
  for (int h = 8; h >= 1; h >>= 1) {
    const bool up = lane & h;
#pragma unroll
    for (int k = 0; k < h; ++k) { const float send = up ? w[k] : w[k + h]; const float keep = up ? w[k + h] : w[k]; w[k] = keep + __shfl_xor_sync(0xffffffffu, send, h); }
  }
  return w[0];
}

// the same for 16 column values per lane: lanes 2j and 2j+1 both end with column j
__device__ __forceinline__ float colsum16(const float* v, int lane) {
  float w[8];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float send = up ? v[k] : v[k + 8]; const float keep = up ? v[k + 8] : v[k]; w[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
  }
#pragma unroll
  for (int h = 4; h >= 1; h >>= 1) {
    const bool up = lane & (2 * h);
#pragma unroll
    for (int k = 0; k < h; ++k) { const float send = up ? w[k] : w[k + h]; const float keep = up ? w[k + h] : w[k]; w[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * h); }
  }
  return w[0] + __shfl_xor_sync(0xffffffffu, w[0], 1);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace clipk
