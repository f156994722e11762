// Shared declarations of the attention kernels (attention.cu: first-generation one-shot kernels; attention_fwd2.cu / attention_bwd2.cu:
// persistent warp-specialised pipelines).
#pragma once
#include "common.cuh"

namespace clipk {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct AttnParams {
  int B, L, H, d;
  int lk_pad;          // keys padded to a multiple of 16 (<= 256)
  int q_tiles;         // ceil(L / 128)
  float scale;         // 1/sqrt(64)
  const float* mask;   // [B, L] additive key mask or null
  bf16* ctx;           // fwd out [B*L, d]
  float* lse;          // [B, H, L] natural-log LSE of the scaled+masked scores
  const bf16* ctx_in;  // bwd in
  const bf16* dctx;    // bwd in  [B*L, d]
  bf16* dqkv;          // bwd out [B*L, 3d]
  float* dqkv_colsum;  // bwd out (optional) [3d] += column sums of dqkv = gradient of the QKV projection bias
  int causal;          // 1: key j attends only to queries i >= j (additive -inf above the diagonal; OPEN_CLIP.build_attention_mask, modeling_openclip.py:346-352)
  int flags;           // experiment switches (CLIPK_ATTN_FLAGS): bit 0 = issue the gradient MMA chains one after the other instead of interleaved
  long long* dbg;      // optional (diagnostics, CLIPK_ATTN_DBG_PTR): per-role clock64 timestamps of CTA 0, [64 tiles][16 events]
  DropArg drop;        // dropout on the attention probabilities (modeling_bert.py:238); element (b,h,q,j): row = (b*H+h)*L+q, quad = j/4
};

__device__ __forceinline__ uint64_t desc_k(uint32_t addr) { return umma_smem_desc(addr, 16, 1024); }                // K-major
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo) { return umma_smem_desc(addr, lbo, 1024); }  // MN-major


// ---- tcgen05.ld / st of NREG consecutive 32-bit columns of the thread's TMEM lane (32x32b shape) --------------------------------
template <int N> struct TmemIO;
template <> struct TmemIO<2> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a) : "memory");
  }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a), "r"(r[0]), "r"(r[1]) : "memory");
  }
};
template <> struct TmemIO<4> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
  }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
  }
};
template <> struct TmemIO<8> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
  }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
  }
};
template <> struct TmemIO<16> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t* r) { tmem_ld_x16(a, r); }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(a),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
  }
};
template <> struct TmemIO<32> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t* r) { tmem_ld_x32(a, r); }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(a),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
  }
};
// N columns as a sum of power-of-two pieces (largest first), fully unrolled at compile time
template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t a, uint32_t* r) {
  if constexpr (N >= 32) { TmemIO<32>::ld(a, r); tmem_ld_n<N - 32>(a + 32, r + 32); }
  else if constexpr (N >= 16) { TmemIO<16>::ld(a, r); tmem_ld_n<N - 16>(a + 16, r + 16); }
  else if constexpr (N >= 8) { TmemIO<8>::ld(a, r); tmem_ld_n<N - 8>(a + 8, r + 8); }
  else if constexpr (N >= 4) { TmemIO<4>::ld(a, r); tmem_ld_n<N - 4>(a + 4, r + 4); }
  else if constexpr (N >= 2) { TmemIO<2>::ld(a, r); tmem_ld_n<N - 2>(a + 2, r + 2); }
  else static_assert(N == 0, "column count must be even");
}
template <int N>
__device__ __forceinline__ void tmem_st_n(uint32_t a, const uint32_t* r) {
  if constexpr (N >= 32) { TmemIO<32>::st(a, r); tmem_st_n<N - 32>(a + 32, r + 32); }
  else if constexpr (N >= 16) { TmemIO<16>::st(a, r); tmem_st_n<N - 16>(a + 16, r + 16); }
  else if constexpr (N >= 8) { TmemIO<8>::st(a, r); tmem_st_n<N - 8>(a + 8, r + 8); }
  else if constexpr (N >= 4) { TmemIO<4>::st(a, r); tmem_st_n<N - 4>(a + 4, r + 4); }
  else if constexpr (N >= 2) { TmemIO<2>::st(a, r); tmem_st_n<N - 2>(a + 2, r + 2); }
  else static_assert(N == 0, "column count must be even");
}

// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (here: the bf16 probabilities, two per 32-bit column) is read from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// register re-partitioning between warpgroups (all 4 warps of a warpgroup execute the same instruction)
template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// second-generation kernels (host entry points; return 0 or a CLIPK_ERR_*)
int attention_fwd2(const void* qkv, const AttnParams& p, cudaStream_t stream);
int attention_bwd2(const void* qkv, const AttnParams& p, cudaStream_t stream);

}  // namespace clipk
