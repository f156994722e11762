// Persistent warp-specialised bf16 GEMM on tcgen05 tensor cores (sm_100a).
//
//   D[M,N] = epilogue( A (*) B )      fp32 accumulation in TMEM
//
// What it replaces in the reference (all of them ATen/cuBLAS calls plus separate elementwise kernels there):
//   * nn.Linear / F.linear of the ViT blocks -- in_proj + out_proj inside nn.MultiheadAttention and mlp.c_fc / c_proj
//     (easynlp/modelzoo/models/clip/modeling_chineseclip.py:188-204), the patch nn.Conv2d as a GEMM over im2col rows (:224,237),
//     `x @ self.proj` (:251) and `x[:, 0, :] @ self.text_projection` (:350);
//   * BERT's query/key/value/dense Linears (easynlp/modelzoo/models/bert/modeling_bert.py:145-147,264-268,329-346);
//   * the activations fused as epilogues: QuickGELU x*sigmoid(1.702x) (modeling_chineseclip.py:179-181), erf-GELU
//     (easynlp/modelzoo/activations.py:45-48), bias adds, and in backward the multiply by the saved derivative + bias-gradient sums;
//   * autograd's dgrad / wgrad GEMMs of all of the above (no transposes materialised: see operand storage below);
//   * the contrastive logits `logit_scale * text @ image.t()` (easynlp/appzoo/clip/model.py:148-149) through bf16 hi/lo splits, and
//     CLIPEvaluator's N x N similarity + per-row sort (easynlp/appzoo/clip/evaluator.py:47-61) through the RANK_COUNT epilogue.
//
// Operand storage (both bf16, leading dimension in elements, multiple of 8):
//   A K-major  : A[M, K] row-major (the activation in y = x W^T)           a_mn_major = 0
//   A MN-major : A[K, M] row-major (dY in dW = dY^T X, contraction on rows) a_mn_major = 1
//   B K-major  : B[N, K] row-major (an nn.Linear weight [out, in])          b_mn_major = 0
//   B MN-major : B[K, N] row-major (W in dX = dY W, X in dW = dY^T X)       b_mn_major = 1
//
// Structure: grid = #SMs, DYNAMIC tile scheduler (global atomic counter -> 2-deep smem queue), 128 x BN output tile, BK = 64,
// 4-stage TMA->smem ring (SWIZZLE_128B), one MMA-issuing thread (tcgen05.mma cta_group::1, M=128, N=BN, K=16), accumulators
// double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//   warp 0: TMA producer + scheduler   warp 1: MMA issuer + TMEM owner   warps 2-9 (2-17 with CLIPK_GEMM_EPI16): epilogue
//   epilogue variants: see the EPI template parameter below and profiles/r01_gemm_epilogue_tma.md
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "gemm_tiles.h"
#include "../../include/clipk.h"

namespace clipk {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int EPI_WARPS = 8;                       // two warps per TMEM lane quarter, each owning half of the tile's columns
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;  // warp 0 TMA, warp 1 MMA, warps 2.. epilogue

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, splits, k_per_split;  // k_per_split multiple of BK
  int* tile_counter;                          // {next tile, CTAs done}: dynamic tile scheduler of the 1-CTA kernel (self re-arming)
  int group_m;                                // tile rasterisation: consecutive tile ids walk down group_m row tiles before the next column tile
  clipk_epilogue_t epi;
};

// tile id -> (row tile, column tile): gemm_tiles.h (also compiled for the host by tests/test_tile_coords_host.py)
__device__ __forceinline__ void tile_coords(const GemmParams& p, int mn, int& mi, int& ni) {
  tile_coords_raw(p.m_tiles, p.n_tiles, p.group_m, mn, mi, ni);
}

template <int BN>
struct GemmSmem {
  static constexpr int NSTAGE = STAGES;
  static constexpr int EPI_WARP_BYTES = 32 * 32 * 4;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_OFFSET = NSTAGE * STAGE_BYTES;          // per epilogue warp: staging (fp32 transpose slab / bf16 store tiles)
  static constexpr int EPI_BYTES = EPI_WARPS * EPI_WARP_BYTES;
  static constexpr int BAR_OFFSET = EPI_OFFSET + EPI_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // + barriers + alignment slack
};

// Epilogue of one staged [32 rows x 32 cols] chunk.  8 lanes cover 32 consecutive columns of a row and a warp covers 4 rows per
// step, so every global access is a coalesced 128-B (fp32) / 64-B (bf16) segment.  All global LOADS of the chunk (residual, GELU
// pre-activation) are issued before any store: `out` may alias `residual`, so the compiler cannot hoist them itself and a
// load -> store -> load chain would expose one DRAM latency per row group.
// bf16 multiplier tile (MUL_AUX) of one chunk: 8 row groups x 4 columns per lane.  Loaded one chunk AHEAD of its use (and the first chunk
// before the accumulator is even complete) so the DRAM/L2 latency overlaps the previous chunk / the mainloop.
__device__ __forceinline__ void epi_load_aux(const GemmParams& p, int row0, int col, int sub_r, uint2* zz) {
  const clipk_epilogue_t& e = p.epi;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = row0 + 4 * i + sub_r;
    zz[i] = (row < p.M && col < p.N) ? *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(e.aux) + (size_t)row * e.ldaux + col) : make_uint2(0u, 0u);
  }
}

__device__ __forceinline__ void epi_chunk(const GemmParams& p, const float* slab, int row0, int col, int sub_r, int sub_c, const uint2* zz) {
  const clipk_epilogue_t& e = p.epi;
  const bool col_ok = col < p.N;
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = 4 * i + sub_r;
    v[i] = *reinterpret_cast<const float4*>(slab + rr * 32 + (((sub_c >> 2) ^ (rr & 7)) << 2));
  }
  if (!col_ok) return;      // warp-uniform whenever N % 32 == 0 (required for the colsum shuffles, checked on the host)
  const float al = e.alpha;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e.bias) b = __ldg(reinterpret_cast<const float4*>(e.bias + col));
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i].x = v[i].x * al + b.x; v[i].y = v[i].y * al + b.y; v[i].z = v[i].z * al + b.z; v[i].w = v[i].w * al + b.w; }

  if (e.mode == CLIPK_EPI_ATOMIC_ADD) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row0 + 4 * i + sub_r;
      if (row < p.M) atomicAdd(reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + (size_t)row * e.ldo + col), v[i]);
    }
    return;
  }
  if (e.mode == CLIPK_EPI_QUICK_GELU || e.mode == CLIPK_EPI_ERF_GELU) {
    // out2 = act(z) (bf16, the next GEMM's operand); out = act'(z) (bf16), saved INSTEAD of z: the backward epilogue then is a
    // single multiply (no MUFU), which keeps the K = 768 dgrad GEMM off the epilogue roofline.  Both are evaluated on the fp32 z.
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row0 + 4 * i + sub_r;
      float a0, a1, a2, a3, g0, g1, g2, g3;
      if (e.mode == CLIPK_EPI_QUICK_GELU) {
        quick_gelu_both(v[i].x, a0, g0); quick_gelu_both(v[i].y, a1, g1); quick_gelu_both(v[i].z, a2, g2); quick_gelu_both(v[i].w, a3, g3);
      } else {
        erf_gelu_both(v[i].x, a0, g0); erf_gelu_both(v[i].y, a1, g1); erf_gelu_both(v[i].z, a2, g2); erf_gelu_both(v[i].w, a3, g3);
      }
      if (row < p.M) {
        uint2 a, g;
        a.x = pack_bf16x2(a0, a1); a.y = pack_bf16x2(a2, a3);
        g.x = pack_bf16x2(g0, g1); g.y = pack_bf16x2(g2, g3);
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(e.out) + (size_t)row * e.ldo + col) = g;
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(e.out2) + (size_t)row * e.ldo2 + col) = a;
      }
    }
    return;
  }
  if (e.mode == CLIPK_EPI_MUL_AUX) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint2 z = zz[i];
      const float2 z0 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&z.x));
      const float2 z1 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&z.y));
      v[i].x *= z0.x; v[i].y *= z0.y; v[i].z *= z1.x; v[i].w *= z1.y;
    }
  }
  if (e.residual) {
    float4 r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row0 + 4 * i + sub_r;
      r[i] = (row < p.M) ? *reinterpret_cast<const float4*>(e.residual + (size_t)row * e.ldr + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i].x += r[i].x; v[i].y += r[i].y; v[i].z += r[i].z; v[i].w += r[i].w; }
  }
  if (e.colsum) {
    // column sums of this [32 rows x 32 cols] chunk: 8 rows in registers, then the 4 row groups of the warp (lane bits 3,4)
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (row0 + 4 * i + sub_r < p.M) { cs.x += v[i].x; cs.y += v[i].y; cs.z += v[i].z; cs.w += v[i].w; }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
      cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
    }
    if (sub_r == 0) atomicAdd(reinterpret_cast<float4*>(e.colsum + col), cs);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = row0 + 4 * i + sub_r;
    if (row >= p.M) continue;
    uint2 o; o.x = pack_bf16x2(v[i].x, v[i].y); o.y = pack_bf16x2(v[i].z, v[i].w);
    if (e.out_dtype == CLIPK_F32) *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + (size_t)row * e.ldo + col) = v[i];
    else *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(e.out) + (size_t)row * e.ldo + col) = o;
    if (e.out2) *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(e.out2) + (size_t)row * e.ldo2 + col) = o;   // bf16 shadow of an fp32 result
  }
}

// L2 prefetch of the epilogue's global INPUTS (fp32 residual / bf16 GELU pre-activation) for one warp's [32 rows x BN/2 cols]
// slice, issued one tile ahead so the epilogue's loads hit L2 instead of exposing a DRAM latency per chunk.
template <int BN>
__device__ __forceinline__ void epi_prefetch(const GemmParams& p, int tile, int tiles_mn, int q, int half, int lane) {
  const clipk_epilogue_t& e = p.epi;
  if (!e.residual && !e.aux) return;
  const int mn = tile % tiles_mn;
  int mi, ni;
  tile_coords(p, mn, mi, ni);
  const int row = mi * BM + q * 32 + lane;
  const int col = ni * BN + half * (BN / 2);
  if (row >= p.M || col >= p.N) return;
  if (e.residual) {
    const char* ptr = reinterpret_cast<const char*>(e.residual + (size_t)row * e.ldr + col);
#pragma unroll
    for (int i = 0; i < (BN / 2) * 4 / 128; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + i * 128));
  }
  if (e.aux) {
    const char* ptr = reinterpret_cast<const char*>(reinterpret_cast<const bf16*>(e.aux) + (size_t)row * e.ldaux + col);
#pragma unroll
    for (int i = 0; i < (BN / 2) * 2 / 128; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + i * 128));
  }
}


// bf16 multiplier row segment (MUL_AUX) for one lane = one accumulator row: 32 consecutive columns = 64 B
__device__ __forceinline__ void epi_load_aux_row(const GemmParams& p, int row, int col, uint4* z) {
  const clipk_epilogue_t& e = p.epi;
  if (row < p.M && col < p.N) {
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(e.aux) + (size_t)row * e.ldaux + col);
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j] = src[j];
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// Epilogue of one warp's [32 rows x BN/2 cols] slice for bf16 outputs WITHOUT a residual input: the math runs directly on the
// tcgen05.ld register layout (lane = row, 32 consecutive columns per chunk), the packed bf16 chunk goes to a 64B-swizzled
// 32 x 32 staging tile (conflict-free STS.128) and ONE thread hands it to the TMA store unit, which also clips at the matrix edge.
// Compared with the transpose-through-smem path this removes the LDS pass, the per-row address arithmetic and all global store
// instructions from the warp.  MODE is a template parameter and the chunk loop is NOT unrolled: the fully unrolled runtime-mode
// version was 7.4 k SASS instructions and stalled on instruction fetch (profiles/r01_gemm_epilogue_tma.md).
// stage: NBUF x 2 KB per warp; seq: running count of committed bulk groups of this warp (GELU commits a pair of tiles per group).
template <int BN, int NBUF, int MODE>
__device__ __forceinline__ void epi_tile_tma(const GemmParams& p, const CUtensorMap* tmO, const CUtensorMap* tmO2, uint8_t* stage, uint32_t& seq,
                                             uint64_t* full_bar, uint32_t full_phase, uint32_t t_row, int row0, int col0, int lane, bool has_k) {
  constexpr int CHUNKS = BN / 64;
  constexpr bool GELU = MODE == CLIPK_EPI_QUICK_GELU || MODE == CLIPK_EPI_ERF_GELU;
  constexpr bool AUX = MODE == CLIPK_EPI_MUL_AUX;
  const clipk_epilogue_t& e = p.epi;
  const int row = row0 + lane;
  const float al = e.alpha;
  uint4 zc[4], zn[4];
  uint32_t r[32];
  if (AUX) epi_load_aux_row(p, row, col0, zc);              // in flight while the MMAs of this tile finish
  mbar_wait(full_bar, full_phase);
  tc_fence_after();
  tmem_ld_x32(t_row, r);
  const uint32_t sw = (uint32_t)((lane >> 1) & 3);
#pragma unroll 1
  for (int c = 0; c < CHUNKS; ++c) {
    const int col = col0 + c * 32;
    const bool live = has_k && col < p.N;                   // warp-uniform (N % 32 == 0 on this path)
    float4 b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = (e.bias && live) ? __ldg(reinterpret_cast<const float4*>(e.bias + col) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    tmem_wait_ld();
    float v[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[4 * j] = fmaf(__uint_as_float(r[4 * j]), al, b[j].x); v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), al, b[j].y);
      v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), al, b[j].z); v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), al, b[j].w);
    }
    if (c + 1 < CHUNKS) {
      tmem_ld_x32(t_row + (c + 1) * 32, r);                 // next chunk streams in while this one is processed
      if (AUX) epi_load_aux_row(p, row, col + 32, zn);
    }
    if (live) {
      uint32_t o[16], o2[16];
      if (GELU) {
        // out2 = act(z) (the next GEMM's operand); out = act'(z), saved INSTEAD of z so that the backward epilogue is a plain multiply
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a0, a1, g0, g1;
          if (MODE == CLIPK_EPI_QUICK_GELU) { quick_gelu_both(v[2 * j], a0, g0); quick_gelu_both(v[2 * j + 1], a1, g1); }
          else                              { erf_gelu_both(v[2 * j], a0, g0);   erf_gelu_both(v[2 * j + 1], a1, g1); }
          o[j] = pack_bf16x2(g0, g1); o2[j] = pack_bf16x2(a0, a1);
        }
      } else {
        if (AUX) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t zw[4] = {zc[j].x, zc[j].y, zc[j].z, zc[j].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              v[8 * j + 2 * t] *= __uint_as_float(zw[t] << 16);             // bf16 -> fp32 is a shift / mask
              v[8 * j + 2 * t + 1] *= __uint_as_float(zw[t] & 0xffff0000u);
            }
          }
        }
        if (e.colsum) {
          if (row >= p.M) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          const float cs = colsum32(v, lane);
          atomicAdd(e.colsum + col + lane, cs);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      }
      uint8_t* bufO = stage + (GELU ? (seq % (NBUF / 2)) * 4096u : (seq % NBUF) * 2048u);
      uint8_t* bufA = bufO + 2048;
      if (lane == 0) { if (GELU) tma_store_wait_read<NBUF / 2 - 1>(); else tma_store_wait_read<NBUF - 1>(); }   // the tile(s) written now are released
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t off = (uint32_t)lane * 64u + ((((uint32_t)j) ^ sw) << 4);
        *reinterpret_cast<uint4*>(bufO + off) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        if (GELU) *reinterpret_cast<uint4*>(bufA + off) = make_uint4(o2[4 * j], o2[4 * j + 1], o2[4 * j + 2], o2[4 * j + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(tmO, bufO, col, row0);
        if (GELU) tma_store_2d(tmO2, bufA, col, row0);
        tma_store_commit();
      }
      ++seq;
    }
    if (AUX) {
#pragma unroll
      for (int j = 0; j < 4; ++j) zc[j] = zn[j];
    }
  }
}

// 16-epilogue-warp form of epi_tile_tma (EPI 6..9, opt-in: CLIPK_GEMM_EPI16=1): four warps per scheduler cover the TMEM / L2 / bulk-store
// latencies by thread-level parallelism instead of software pipelining.  A warp owns [32 rows x BN/4 cols]; 576 threads leave 112
// registers per thread, so there is no accumulator / multiplier prefetch, and ONE 2 KB staging tile per warp (GELU stores its two
// tiles one after the other).
template <int BN, int MODE>
__device__ __forceinline__ void epi_tile_tma16(const GemmParams& p, const CUtensorMap* tmO, const CUtensorMap* tmO2, uint8_t* tile,
                                               uint64_t* full_bar, uint32_t full_phase, uint32_t t_row, int row0, int col0, int lane, bool has_k) {
  constexpr int CHUNKS = BN / 128;
  constexpr bool GELU = MODE == CLIPK_EPI_QUICK_GELU || MODE == CLIPK_EPI_ERF_GELU;
  constexpr bool AUX = MODE == CLIPK_EPI_MUL_AUX;
  const clipk_epilogue_t& e = p.epi;
  const int row = row0 + lane;
  const float al = e.alpha;
  const uint32_t sw = (uint32_t)((lane >> 1) & 3);
  uint4 zc[4];
  if (AUX) epi_load_aux_row(p, row, col0, zc);
  mbar_wait(full_bar, full_phase);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < CHUNKS; ++c) {
    const int col = col0 + c * 32;
    const bool live = has_k && col < p.N;                   // warp-uniform (N % 32 == 0 on this path)
    uint32_t r[32];
    tmem_ld_x32(t_row + c * 32, r);
    if (AUX && c > 0) epi_load_aux_row(p, row, col, zc);
    tmem_wait_ld();
    if (!live) continue;
    uint32_t o[16], o2[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = e.bias ? __ldg(reinterpret_cast<const float4*>(e.bias + col) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      float v0 = fmaf(__uint_as_float(r[4 * j]), al, b.x), v1 = fmaf(__uint_as_float(r[4 * j + 1]), al, b.y);
      float v2 = fmaf(__uint_as_float(r[4 * j + 2]), al, b.z), v3 = fmaf(__uint_as_float(r[4 * j + 3]), al, b.w);
      if (GELU) {
        float a0, a1, a2, a3, g0, g1, g2, g3;
        if (MODE == CLIPK_EPI_QUICK_GELU) { quick_gelu_both(v0, a0, g0); quick_gelu_both(v1, a1, g1); quick_gelu_both(v2, a2, g2); quick_gelu_both(v3, a3, g3); }
        else                              { erf_gelu_both(v0, a0, g0);   erf_gelu_both(v1, a1, g1);   erf_gelu_both(v2, a2, g2);   erf_gelu_both(v3, a3, g3); }
        o[2 * j] = pack_bf16x2(g0, g1); o[2 * j + 1] = pack_bf16x2(g2, g3);
        o2[2 * j] = pack_bf16x2(a0, a1); o2[2 * j + 1] = pack_bf16x2(a2, a3);
      } else {
        if (AUX) {
          const uint4 z = zc[j >> 1];
          const uint32_t w0 = (j & 1) ? z.z : z.x, w1 = (j & 1) ? z.w : z.y;
          v0 *= __uint_as_float(w0 << 16); v1 *= __uint_as_float(w0 & 0xffff0000u);
          v2 *= __uint_as_float(w1 << 16); v3 *= __uint_as_float(w1 & 0xffff0000u);
        }
        r[4 * j] = __float_as_uint(v0); r[4 * j + 1] = __float_as_uint(v1); r[4 * j + 2] = __float_as_uint(v2); r[4 * j + 3] = __float_as_uint(v3);
        o[2 * j] = pack_bf16x2(v0, v1); o[2 * j + 1] = pack_bf16x2(v2, v3);
      }
    }
    if (!GELU && e.colsum) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (row < p.M) ? __uint_as_float(r[j]) : 0.f;
      const float cs = colsum32(v, lane);
      atomicAdd(e.colsum + col + lane, cs);
    }
#pragma unroll
    for (int pass = 0; pass < (GELU ? 2 : 1); ++pass) {
      if (lane == 0) tma_store_wait_read<0>();               // the tile is no longer being read by the previous bulk store
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t off = (uint32_t)lane * 64u + ((((uint32_t)j) ^ sw) << 4);
        *reinterpret_cast<uint4*>(tile + off) = pass ? make_uint4(o2[4 * j], o2[4 * j + 1], o2[4 * j + 2], o2[4 * j + 3])
                                                     : make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { tma_store_2d(pass ? tmO2 : tmO, tile, col, row0); tma_store_commit(); }
    }
  }
}

// Epilogue of CLIPK_EPI_RANK_COUNT: nothing is stored -- every accumulator of the warp's [32 rows x BN/2 cols] slice is compared with the
// row's threshold (the score of the query's own match) and the number of larger ones goes to rank[row] with one atomic per row and tile.
template <int BN>
__device__ __forceinline__ void epi_tile_rank(const GemmParams& p, uint64_t* full_bar, uint32_t full_phase, uint32_t t_row, int row0, int col0,
                                              int lane, bool has_k) {
  constexpr int CHUNKS = BN / 64;
  const clipk_epilogue_t& e = p.epi;
  const int row = row0 + lane;
  const float thr = (row < p.M) ? reinterpret_cast<const float*>(e.aux)[row] : INFINITY;
  const int lab = e.label_offset + row;
  const float al = e.alpha;
  uint32_t r[32];
  int cnt = 0;
  mbar_wait(full_bar, full_phase);
  tc_fence_after();
  tmem_ld_x32(t_row, r);
#pragma unroll 1
  for (int c = 0; c < CHUNKS; ++c) {
    const int col = col0 + c * 32;
    tmem_wait_ld();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * al;
    if (c + 1 < CHUNKS) tmem_ld_x32(t_row + (c + 1) * 32, r);
    if (has_k && col < p.N) {
#pragma unroll
      for (int j = 0; j < 32; ++j) cnt += (col + j < p.N && col + j != lab && v[j] > thr) ? 1 : 0;
    }
  }
  if (row < p.M && cnt) atomicAdd(reinterpret_cast<int*>(e.out) + row, cnt);
}

// EPI: 0 = transpose-through-smem epilogue (any output type / mode, chosen at run time);
//      1 + mode = TMA-store epilogue of that mode (LINEAR / QUICK_GELU / ERF_GELU / MUL_AUX), 2 staging tiles per warp
//                 (a 3-stage ring + 4 staging tiles per warp was measured too: no gain at K = 768, -7 % at K = 3072);
//      5 = CLIPK_EPI_RANK_COUNT (compare + count, no output matrix).
template <int BN, int A_MN, int B_MN, int EPI>
__global__ void __launch_bounds__(EPI >= 6 ? 64 + 32 * 16 : GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
                 const __grid_constant__ CUtensorMap tmO2, const GemmParams p) {
  constexpr bool TMA_OUT = (EPI >= 1 && EPI <= 4) || EPI >= 6;
  constexpr int EPW = EPI >= 6 ? 16 : EPI_WARPS;       // epilogue warps
  constexpr int NSTAGE = STAGES;
  constexpr int NBUF = 2;                         // 2 KB bf16 store tiles per epilogue warp (EPI == 0: one 4 KB fp32 slab)
  constexpr int MODE = EPI >= 6 ? EPI - 6 : (TMA_OUT ? EPI - 1 : 0);
  using L = GemmSmem<BN>;
  // 1024-B aligned dynamic smem (SWIZZLE_128B atoms); indexing the __shared__ array directly keeps the address space known to
  // the compiler (LDS/STS instead of generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tmem_full = empty_bar + NSTAGE;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* sched_full = tmem_empty + 2;          // dynamic tile scheduler: 2-deep queue of tile ids, producer -> MMA + epilogue warps
  uint64_t* sched_empty = sched_full + 2;
  int* sched_tile = reinterpret_cast<int*>(sched_empty + 2);
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(sched_tile + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int num_tiles = tiles_mn * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EPW); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sched_full[s], 1); mbar_init(&sched_empty[s], 1 + EPW); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int slot = 0; uint32_t sphase = 0;
      while (true) {
        // tiles come from a global counter, not from blockIdx: a CTA that starts late (its SM was busy with a concurrent kernel, e.g.
        // an NCCL all-reduce overlapping the backward pass) finds the counter exhausted instead of holding the whole GEMM back
        int tile = atomicAdd(p.tile_counter, 1);
        if (tile >= num_tiles) tile = -1;
        mbar_wait(&sched_empty[slot], sphase ^ 1);
        sched_tile[slot] = tile;
        mbar_arrive(&sched_full[slot]);
        if (++slot == 2) { slot = 0; sphase ^= 1; }
        if (tile < 0) break;
        const int ks = tile / tiles_mn;
        const int mn = tile - ks * tiles_mn;
        int mi, ni;
        tile_coords(p, mn, mi, ni);
        const int m0 = mi * BM;
        const int n0 = ni * BN;
        const int k_begin = ks * p.k_per_split;
        const int k_end = min(p.K, k_begin + p.k_per_split);
        for (int k0 = k_begin; k0 < k_end; k0 += BK) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * L::STAGE_BYTES;
          uint8_t* sB = sA + L::A_BYTES;
          mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sA + j * (BK * 128), &tmA, &full_bar[stage], m0 + 64 * j, k0);
          } else {
            tma_load_2d(sA, &tmA, &full_bar[stage], k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + j * (BK * 128), &tmB, &full_bar[stage], n0 + 64 * j, k0);
          } else {
            tma_load_2d(sB, &tmB, &full_bar[stage], k0, n0);
          }
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
      // the last CTA to run dry re-arms the counters for the next launch that uses this slot (every CTA's final fetch precedes its
      // arrival here, so nobody reads the counter afterwards)
      if (atomicAdd(p.tile_counter + 1, 1) == (int)gridDim.x - 1) { p.tile_counter[0] = 0; p.tile_counter[1] = 0; __threadfence(); }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (one thread; elect.sync lets ptxas keep the issue loop in
    // the uniform datapath without a per-MMA "for each active lane" loop)
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      int slot = 0; uint32_t sphase = 0;
      while (true) {
        mbar_wait(&sched_full[slot], sphase);
        const int tile = sched_tile[slot];
        mbar_arrive(&sched_empty[slot]);
        if (++slot == 2) { slot = 0; sphase ^= 1; }
        if (tile < 0) break;
        const int ks = tile / tiles_mn;
        const int k_begin = ks * p.k_per_split;
        const int k_end = min(p.K, k_begin + p.k_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t accumulate = 0;
        for (int k0 = k_begin; k0 < k_end; k0 += BK) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sB = sA + L::A_BYTES;
          // descriptors as (lo, hi) halves: the next k-step is one add on the low word (K-major: 16 elements = 32 B -> +2; MN-major: 16 rows
          // of 128 B = 2048 B -> +128) instead of rebuilding 64-bit descriptors (common.cuh: umma_bf16_lh)
          const uint32_t da = A_MN ? umma_desc_lo(sA, BK * 128) : umma_desc_lo(sA, 16);
          const uint32_t db = B_MN ? umma_desc_lo(sB, BK * 128) : umma_desc_lo(sB, 16);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_bf16_lh(d_tmem, da + (A_MN ? 128 : 2) * k, db + (B_MN ? 128 : 2) * k, idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);      // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================================================================ epilogue warps: TMEM lane quarter = warp % 4, column half = (warp-2)/4
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;          // column half (8 epilogue warps) / column quarter (16)
    constexpr int CHUNKS = BN / 32 / 2;        // 32-column chunks per warp (8-warp forms)
    constexpr int WCOLS = BN / (EPW / 4);      // columns per epilogue warp
    int acc = 0; uint32_t acc_phase = 0;
    uint32_t store_seq = 0;
    float* slab = reinterpret_cast<float*>(smem + L::EPI_OFFSET) + (warp - 2) * 32 * 32;
    const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
    int slot = 0; uint32_t sphase = 0;
    while (true) {
      mbar_wait(&sched_full[slot], sphase);
      const int tile = sched_tile[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&sched_empty[slot]);
      if (++slot == 2) { slot = 0; sphase ^= 1; }
      if (tile < 0) break;
      if constexpr (EPI < 6) epi_prefetch<BN>(p, tile, tiles_mn, q, half, lane);    // the id arrives >= 1 tile ahead of the accumulator: L2 prefetch of the epilogue inputs
      const int ks = tile / tiles_mn;
      const int mn = tile - ks * tiles_mn;
      int mi, ni;
      tile_coords(p, mn, mi, ni);
      const int m0 = mi * BM;
      const int n0 = ni * BN + half * WCOLS;
      const int k_begin = ks * p.k_per_split;
      const bool has_k = k_begin < p.K;   // an empty split contributes nothing (host never creates one, but be safe)
      if constexpr (EPI >= 6) {
        epi_tile_tma16<BN, MODE>(p, &tmO, &tmO2, smem + L::EPI_OFFSET + (warp - 2) * 2048, &tmem_full[acc], acc_phase,
                                 tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + half * WCOLS, m0 + q * 32, n0, lane, has_k);
      } else if constexpr (EPI == 5) {
        epi_tile_rank<BN>(p, &tmem_full[acc], acc_phase, tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + half * (BN / 2), m0 + q * 32, n0,
                          lane, has_k);
      } else if constexpr (TMA_OUT) {
        epi_tile_tma<BN, NBUF, MODE>(p, &tmO, &tmO2, smem + L::EPI_OFFSET + (warp - 2) * (NBUF * 2048), store_seq, &tmem_full[acc], acc_phase,
                         tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + half * (BN / 2), m0 + q * 32, n0, lane, has_k);
      } else {
      const bool aux_mode = p.epi.mode == CLIPK_EPI_MUL_AUX;
      uint2 zz[8], zn[8];
      if (aux_mode) epi_load_aux(p, m0 + q * 32, n0 + sub_c, sub_r, zz);     // in flight while the MMAs of this tile finish
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + half * (BN / 2);
      uint32_t r[32];
      tmem_ld_x32(t_row, r);
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        tmem_wait_ld();
        // phase 1: thread = accumulator row -> staging slab (row-major, 16-B units XOR-swizzled by row)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(slab + lane * 32 + ((j ^ (lane & 7)) << 2)) =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        if (c + 1 < CHUNKS) {
          tmem_ld_x32(t_row + (c + 1) * 32, r);   // next chunk streams in while this one is written out
          if (aux_mode) epi_load_aux(p, m0 + q * 32, n0 + (c + 1) * 32 + sub_c, sub_r, zn);
        }
        __syncwarp();
        // phase 2: 8 lanes per row, 4 rows per step -> coalesced global traffic
        if (has_k) epi_chunk(p, slab, m0 + q * 32, n0 + c * 32 + sub_c, sub_r, sub_c, zz);
#pragma unroll
        for (int i = 0; i < 8; ++i) zz[i] = zn[i];
        __syncwarp();
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (TMA_OUT && lane == 0) tma_store_wait_read<0>();   // staging tiles must outlive the bulk stores reading them
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int BN, int A_MN, int B_MN, int EPI>
static int launch_gemm(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tO, const CUtensorMap& tO2, const GemmParams& p,
                       cudaStream_t stream) {
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, EPI>;
  static bool configured = false;
  const int smem = GemmSmem<BN>::TOTAL;
  if (!configured) {
    CLIPK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.splits;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, EPI >= 6 ? 64 + 32 * 16 : GEMM_THREADS, smem, stream>>>(tA, tB, tO, tO2, p);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}


// =====================================================================================================================
// CTA-pair variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile.  CTA r holds rows
// [128 r, 128 r + 128) of A and of the accumulator (its own TMEM) and columns [128 r, 128 r + 128) of B; the leader's single
// MMA thread issues tcgen05.mma.cta_group::2 (M = 256) which reads both CTAs' shared memory, so each SM stages only HALF of
// B per k-block (32 KB / stage instead of 48 KB): operand smem traffic per SM drops from 12 to 8 KB per MMA and the ring is
// 6 stages deep.  Barriers: full[s] (leader; 2 producer arrivals + both CTAs' TMA bytes), empty[s] / tmem_full[a] (per CTA,
// multicast tcgen05.commit), tmem_empty[a] (leader; all epilogue warps of both CTAs).
constexpr int STAGES2 = 6;
constexpr int BN2 = 256;

struct Gemm2Smem {
  static constexpr int A_BYTES = 128 * BK * 2;
  static constexpr int B_BYTES = 128 * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_OFFSET = STAGES2 * STAGE_BYTES;
  static constexpr int EPI_BYTES = EPI_WARPS * 32 * 32 * 4;
  static constexpr int BAR_OFFSET = EPI_OFFSET + EPI_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

template <int A_MN, int B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using L = Gemm2Smem;
  // 1024-B aligned dynamic smem (SWIZZLE_128B atoms); indexing the __shared__ array directly keeps the address space known to
  // the compiler (LDS/STS instead of generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES2;
  uint64_t* tmem_full = empty_bar + STAGES2;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int tiles_mn = p.m_tiles * p.n_tiles;            // m_tiles counts 256-row pair tiles here
  const int num_tiles = tiles_mn * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES2; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 2 * EPI_WARPS); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_holder, 2 * BN2);
  tc_fence_before();
  cluster_sync_all();          // barriers of BOTH CTAs are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ================================================================ TMA producer (both CTAs, each loads its halves)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int ks = tile / tiles_mn;
        const int mn = tile - ks * tiles_mn;
        const int m0 = (mn / p.n_tiles) * 256 + rank * 128;
        const int n0 = (mn % p.n_tiles) * BN2 + rank * 128;
        const int k_begin = ks * p.k_per_split;
        const int k_end = min(p.K, k_begin + p.k_per_split);
        for (int k0 = k_begin; k0 < k_end; k0 += BK) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * L::STAGE_BYTES;
          uint8_t* sB = sA + L::A_BYTES;
          const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_2sm(sA + j * (BK * 128), &tmA, lead_bar, m0 + 64 * j, k0);
          } else {
            tma_load_2d_2sm(sA, &tmA, lead_bar, k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_2sm(sB + j * (BK * 128), &tmB, lead_bar, n0 + 64 * j, k0);
          } else {
            tma_load_2d_2sm(sB, &tmB, lead_bar, k0, n0);
          }
          if (!leader) mbar_arrive_cluster(lead_bar);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (one thread of the leader CTA)
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN2, A_MN, B_MN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int ks = tile / tiles_mn;
        const int k_begin = ks * p.k_per_split;
        const int k_end = min(p.K, k_begin + p.k_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN2;
        uint32_t accumulate = 0;
        for (int k0 = k_begin; k0 < k_end; k0 += BK) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sB = sA + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? umma_smem_desc(sA + k * 2048, BK * 128, 1024) : umma_smem_desc(sA + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_smem_desc(sB + k * 2048, BK * 128, 1024) : umma_smem_desc(sB + k * 32, 16, 1024);
            umma_bf16_2sm(d_tmem, da, db, idesc, accumulate);
            accumulate = 1;
          }
          umma_commit_2sm(&empty_bar[stage], 3);   // both CTAs' smem slots
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tmem_full[acc], 3);       // both CTAs' epilogues
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================================================================ epilogue warps (each CTA drains its own 128 accumulator rows)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int CHUNKS = BN2 / 32 / 2;
    int acc = 0; uint32_t acc_phase = 0;
    float* slab = reinterpret_cast<float*>(smem + L::EPI_OFFSET) + (warp - 2) * 32 * 32;
    const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
    for (int tile = pair; tile < num_tiles; tile += npairs) {
      const int ks = tile / tiles_mn;
      const int mn = tile - ks * tiles_mn;
      const int m0 = (mn / p.n_tiles) * 256 + rank * 128;
      const int n0 = (mn % p.n_tiles) * BN2 + half * (BN2 / 2);
      const int k_begin = ks * p.k_per_split;
      const bool has_k = k_begin < p.K;
      const bool aux_mode = p.epi.mode == CLIPK_EPI_MUL_AUX;
      uint2 zz[8], zn[8];
      if (aux_mode) epi_load_aux(p, m0 + q * 32, n0 + sub_c, sub_r, zz);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN2 + half * (BN2 / 2);
      uint32_t r[32];
      tmem_ld_x32(t_row, r);
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(slab + lane * 32 + ((j ^ (lane & 7)) << 2)) =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        if (c + 1 < CHUNKS) {
          tmem_ld_x32(t_row + (c + 1) * 32, r);
          if (aux_mode) epi_load_aux(p, m0 + q * 32, n0 + (c + 1) * 32 + sub_c, sub_r, zn);
        }
        __syncwarp();
        if (has_k) epi_chunk(p, slab, m0 + q * 32, n0 + c * 32 + sub_c, sub_r, sub_c, zz);
#pragma unroll
        for (int i = 0; i < 8; ++i) zz[i] = zn[i];
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));   // the leader's MMA thread owns the accumulator ring
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();          // nobody exits (or frees TMEM) while the peer may still touch this CTA's smem / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BN2);
  }
}

template <int A_MN, int B_MN>
static int launch_gemm2(const CUtensorMap& tA, const CUtensorMap& tB, const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm2_bf16_kernel<A_MN, B_MN>;
  static bool configured = false;
  const int smem = Gemm2Smem::TOTAL;
  if (!configured) {
    CLIPK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles * p.splits;
  int pairs = sm_count() / 2;
  if (tiles < pairs) pairs = tiles;
  kern<<<2 * pairs, GEMM_THREADS, smem, stream>>>(tA, tB, p);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace clipk

using namespace clipk;

extern "C" int clipk_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, int M, int N,
                               int K, const clipk_epilogue_t* epi, int splits, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) { set_error("clipk_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K); return CLIPK_ERR_ARG; }
  const bool rank_mode = epi && epi->mode == CLIPK_EPI_RANK_COUNT;
  if ((lda % 8) || (ldb % 8) || ((N % 8) && !rank_mode)) { set_error("clipk_gemm_bf16: lda/ldb/N must be multiples of 8 (lda=%d ldb=%d N=%d)", lda, ldb, N); return CLIPK_ERR_ARG; }
  if (!epi || !epi->out) { set_error("clipk_gemm_bf16: epilogue output missing"); return CLIPK_ERR_ARG; }
  if (splits < 1) splits = 1;
  if (splits > 1 && epi->mode != CLIPK_EPI_ATOMIC_ADD) { set_error("clipk_gemm_bf16: split-K requires CLIPK_EPI_ATOMIC_ADD"); return CLIPK_ERR_ARG; }
  if (epi->mode == CLIPK_EPI_ATOMIC_ADD && epi->out_dtype != CLIPK_F32) { set_error("clipk_gemm_bf16: atomic epilogue needs fp32 output"); return CLIPK_ERR_ARG; }
  if ((epi->mode == CLIPK_EPI_QUICK_GELU || epi->mode == CLIPK_EPI_ERF_GELU) && !epi->out2) { set_error("clipk_gemm_bf16: GELU epilogue needs out2"); return CLIPK_ERR_ARG; }
  if (epi->colsum && (N % 32)) { set_error("clipk_gemm_bf16: the fused column sum needs N %% 32 == 0"); return CLIPK_ERR_ARG; }
  if (epi->mode == CLIPK_EPI_MUL_AUX && !epi->aux) { set_error("clipk_gemm_bf16: MUL_AUX epilogue needs aux"); return CLIPK_ERR_ARG; }
  if (rank_mode && (!epi->aux || a_mn_major || splits > 1)) { set_error("clipk_gemm_bf16: RANK_COUNT needs aux (f32 thresholds), a K-major A and no split-K"); return CLIPK_ERR_ARG; }
  if (epi->mode < 0 || epi->mode > CLIPK_EPI_ATOMIC_ADD) { set_error("clipk_gemm_bf16: unknown epilogue mode %d", epi->mode); return CLIPK_ERR_ARG; }

  const int BN = (N % 256 == 0 || N > 512) ? 256 : 128;
  static int use_pair = -1;
  // experimental: correct (tests pass) but measured at half the 1-CTA rate on B200 (profiles/r01_gemm_2cta_note.md) -> opt-in only
  if (use_pair < 0) { const char* ev = getenv("CLIPK_GEMM_2CTA"); use_pair = (ev && ev[0] == '1') ? 1 : 0; }
  const bool pair = use_pair && BN == 256 && M >= 256 && !rank_mode;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = pair ? (M + 255) / 256 : (M + BM - 1) / BM;
  p.n_tiles = (N + BN - 1) / BN;
  int kblocks = (K + BK - 1) / BK;
  if (splits > kblocks) splits = kblocks;
  int kb_per = (kblocks + splits - 1) / splits;
  splits = (kblocks + kb_per - 1) / kb_per;   // no empty split
  p.splits = splits;
  p.k_per_split = kb_per * BK;
  p.epi = *epi;
  if (p.epi.alpha == 0.0f) p.epi.alpha = 1.0f;
  // many column tiles AND a B operand that cannot stay in the 126 MB L2 (the retrieval gallery): band rasterisation (tile_coords)
  p.group_m = (!pair && p.n_tiles > 64 && (size_t)N * (size_t)K * 2 > ((size_t)48 << 20)) ? 16 : 1;
  {
    // per-device pool of self-re-arming scheduler counters; consecutive launches rotate through it so that back-to-back GEMMs never share one
    constexpr int POOL = 64, MAX_DEV = 64;
    static int* pools[MAX_DEV] = {nullptr};
    static unsigned next = 0;
    int dev = 0;
    CLIPK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEV) { set_error("clipk_gemm_bf16: device ordinal %d out of range", dev); return CLIPK_ERR_ARG; }
    if (!pools[dev]) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(stream, &cs);
      if (cs != cudaStreamCaptureStatusNone) { set_error("clipk_gemm_bf16: the first call on a device (allocates the scheduler counters) must not happen during stream capture"); return CLIPK_ERR_CUDA; }
      CLIPK_CUDA(cudaMalloc(&pools[dev], POOL * 2 * sizeof(int)));
      CLIPK_CUDA(cudaMemset(pools[dev], 0, POOL * 2 * sizeof(int)));
    }
    p.tile_counter = pools[dev] + 2 * (next++ % POOL);
  }

  CUtensorMap tA, tB;
  int rc;
  if (a_mn_major) rc = make_tmap_2d_bf16(&tA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK);
  else            rc = make_tmap_2d_bf16(&tA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
  if (rc) return rc;
  if (b_mn_major) rc = make_tmap_2d_bf16(&tB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
  else            rc = make_tmap_2d_bf16(&tB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, pair ? 128 : BN);
  if (rc) return rc;
  if (pair) {
    if (a_mn_major) {
      if (b_mn_major) return launch_gemm2<1, 1>(tA, tB, p, stream);
      return launch_gemm2<1, 0>(tA, tB, p, stream);
    }
    if (b_mn_major) return launch_gemm2<0, 1>(tA, tB, p, stream);
    return launch_gemm2<0, 0>(tA, tB, p, stream);
  }

  // bf16 results without a residual input leave through the TMA store unit (epi_tile_tma); fp32 / atomic / residual epilogues keep the
  // transpose-through-smem path (CLIPK_GEMM_NO_TMA_OUT=1 forces it everywhere, for A/B measurements)
  const clipk_epilogue_t& ee = p.epi;
  const bool gelu = ee.mode == CLIPK_EPI_QUICK_GELU || ee.mode == CLIPK_EPI_ERF_GELU;
  static int no_tma_out = -1;
  if (no_tma_out < 0) { const char* ev = getenv("CLIPK_GEMM_NO_TMA_OUT"); no_tma_out = (ev && ev[0] == '1') ? 1 : 0; }
  const bool tma_out = !no_tma_out && ee.out_dtype == CLIPK_BF16 && !ee.residual && ee.mode != CLIPK_EPI_ATOMIC_ADD && ee.mode != CLIPK_EPI_RANK_COUNT &&
                       (gelu ? ee.out2 != nullptr : ee.out2 == nullptr) && (N % 32 == 0) && (ee.ldo % 8 == 0) &&
                       !(reinterpret_cast<uintptr_t>(ee.out) & 15) && (!gelu || (ee.ldo2 % 8 == 0 && !(reinterpret_cast<uintptr_t>(ee.out2) & 15))) &&
                       (ee.mode != CLIPK_EPI_MUL_AUX || (ee.ldaux % 8 == 0 && !(reinterpret_cast<uintptr_t>(ee.aux) & 15))) &&
                       (!ee.bias || !(reinterpret_cast<uintptr_t>(ee.bias) & 15));
  const int variant = (tma_out && !a_mn_major) ? 1 : 0;   // 0 = transpose-through-smem epilogue, 1 = TMA-store epilogue
  CUtensorMap tO = tA, tO2 = tA;
  if (variant) {
    rc = make_tmap_2d_bf16(&tO, ee.out, (uint64_t)N, (uint64_t)M, (uint64_t)ee.ldo, 32, 32, 64);
    if (rc) return rc;
    if (gelu) { rc = make_tmap_2d_bf16(&tO2, ee.out2, (uint64_t)N, (uint64_t)M, (uint64_t)ee.ldo2, 32, 32, 64); if (rc) return rc; }
  }
  if (rank_mode) {
    if (BN == 256) { if (b_mn_major) return launch_gemm<256, 0, 1, 5>(tA, tB, tO, tO2, p, stream); return launch_gemm<256, 0, 0, 5>(tA, tB, tO, tO2, p, stream); }
    if (b_mn_major) return launch_gemm<128, 0, 1, 5>(tA, tB, tO, tO2, p, stream);
    return launch_gemm<128, 0, 0, 5>(tA, tB, tO, tO2, p, stream);
  }
  if (variant == 0) {
#define CLIPK_DISPATCH(BN_)                                                          \
    if (a_mn_major) {                                                                \
      if (b_mn_major) return launch_gemm<BN_, 1, 1, 0>(tA, tB, tO, tO2, p, stream);  \
      return launch_gemm<BN_, 1, 0, 0>(tA, tB, tO, tO2, p, stream);                  \
    } else {                                                                         \
      if (b_mn_major) return launch_gemm<BN_, 0, 1, 0>(tA, tB, tO, tO2, p, stream);  \
      return launch_gemm<BN_, 0, 0, 0>(tA, tB, tO, tO2, p, stream);                  \
    }
    if (BN == 256) { CLIPK_DISPATCH(256) } else { CLIPK_DISPATCH(128) }
#undef CLIPK_DISPATCH
  }
  static int epi16 = -1;   // opt-in: 16 epilogue warps (CLIPK_GEMM_EPI16=1)
  if (epi16 < 0) { const char* ev = getenv("CLIPK_GEMM_EPI16"); epi16 = (ev && ev[0] == '1') ? 1 : 0; }
  const int epi_id = (epi16 ? 6 : 1) + ee.mode;   // ee.mode in 0..3 on this path
#define CLIPK_DISPATCH_E(BN_, E_)                                                    \
  case E_:                                                                           \
    if (b_mn_major) return launch_gemm<BN_, 0, 1, E_>(tA, tB, tO, tO2, p, stream);   \
    return launch_gemm<BN_, 0, 0, E_>(tA, tB, tO, tO2, p, stream);
#define CLIPK_DISPATCH_BN(BN_)                                                       \
  switch (epi_id) {                                                                  \
    CLIPK_DISPATCH_E(BN_, 1) CLIPK_DISPATCH_E(BN_, 2) CLIPK_DISPATCH_E(BN_, 3) CLIPK_DISPATCH_E(BN_, 4)      \
    CLIPK_DISPATCH_E(BN_, 6) CLIPK_DISPATCH_E(BN_, 7) CLIPK_DISPATCH_E(BN_, 8) CLIPK_DISPATCH_E(BN_, 9)      \
    default: break;                                                                  \
  }
  if (BN == 256) { CLIPK_DISPATCH_BN(256) } else { CLIPK_DISPATCH_BN(128) }
  set_error("clipk_gemm_bf16: internal dispatch error (epilogue id %d)", epi_id);
  return CLIPK_ERR_ARG;
#undef CLIPK_DISPATCH_E
#undef CLIPK_DISPATCH_BN
}
