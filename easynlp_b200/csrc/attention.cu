// Multi-head self-attention (head dim 64) forward and backward on tcgen05 tensor cores, one-shot over the whole key
// range (sequences on this path are 197 / 77 / <= 256 tokens, so K and V of one head live in shared memory).
//
//   forward : S = Q K^T (TMEM) -> row softmax in registers (one thread per query row, straight from TMEM)
//             -> P (bf16, SWIZZLE_128B in smem) -> O = P V (TMEM) -> ctx, log-sum-exp
//   backward: per 128-query tile: S -> P ; dP = dO V^T ; dV += P^T dO ; dS = scale * P o (dP - D) ;
//             dQ = dS K ; dK += dS^T Q      (dV / dK stay resident in TMEM across query tiles)
//
// Replaces nn.MultiheadAttention's SDPA (modeling_chineseclip.py:188,198-200) and BertSelfAttention's unfused
// QK^T / +mask / softmax / PV chain (modeling_bert.py:210-244, additive key mask (1-m)*-10000 from
// modeling_utils.py:438-439), which materialises [B,12,L,L] scores in HBM.
//
// Input layout: packed projections qkv[B*L, 3*d] bf16 (Q | K | V column blocks, head h at columns h*64 of each block),
// i.e. exactly the output of the in_proj / fused query-key-value GEMM; output ctx[B*L, d].
#include <stdlib.h>
#include "common.cuh"
#include "attention_common.cuh"
#include "../../include/clipk.h"

namespace clipk {

__device__ __forceinline__ void store_row8_sw128(uint8_t* tile_base, int row, int col, const float* v) {
  // 8 consecutive bf16 (cols col..col+7, col % 8 == 0) of row `row` in a [128-row x 64-col]-blocked SWIZZLE_128B tile
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(tile_base + (col >> 6) * 16384 + sw128_offset(row, (col & 63) >> 3)) = o;
}

// ====================================================================================================== forward
// 256 threads: warp w owns TMEM lane quarter (w & 3) -- one thread per query row -- and column half (w >> 2).
constexpr int ATT_THREADS = 256;

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnParams p) {
  // 1024-B aligned dynamic smem (SWIZZLE_128B atoms); indexing the __shared__ array directly keeps the address space known to
  // the compiler (LDS/STS instead of generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q4 = warp & 3, half = warp >> 2;
  const int k_bytes = p.lk_pad * 128;
  const int p_bytes = ((p.lk_pad + 63) >> 6) * 16384;
  const int v_off = max(p_bytes, 16384 + k_bytes);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;
  uint8_t* sP = smem;             // overlays Q and K once S = Q K^T has completed
  uint8_t* sV = smem + v_off;
  float* smask = reinterpret_cast<float*>(sV + k_bytes);          // [256]
  float* sred = smask + 256;                                      // [2][128] row max per column half
  float* ssum = sred + 256;                                       // [2][128] row sum per column half
  uint64_t* bars = reinterpret_cast<uint64_t*>(ssum + 256);       // load, s, o
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 3);

  const uint32_t ncols = p.lk_pad > 128 ? 256 : (p.lk_pad > 64 ? 128 : 64);
  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV);
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_holder, ncols);
  for (int j = tid; j < 256; j += ATT_THREADS)
    smask[j] = (j < p.L) ? (p.mask ? p.mask[(long long)b * p.L + j] * LOG2E : 0.f) : -INFINITY;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;

  if (tid == 0) {
    mbar_expect_tx(&bars[0], 16384 + 2 * k_bytes);
    tma_load_2d(sQ, &tmQ, &bars[0], h * 64, b * p.L + qt * 128);
    tma_load_2d(sK, &tmKV, &bars[0], p.d + h * 64, b * p.L);
    tma_load_2d(sV, &tmKV, &bars[0], 2 * p.d + h * 64, b * p.L);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, p.lk_pad, 0, 0);
    const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, desc_k(aQ + k * 32), desc_k(aK + k * 32), idesc, k > 0);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();

  const int row = q4 * 32 + lane;
  const int q = qt * 128 + row;
  const uint32_t t_row = tmem + ((uint32_t)(q4 * 32) << 16);
  const float sc = p.scale * LOG2E;
  const int split = ((p.lk_pad >> 1) + 15) & ~15;
  const int c_begin = half ? split : 0, c_end = half ? p.lk_pad : split;
  const DropCtx dc = drop_ctx(p.drop);
  const uint32_t drow = (uint32_t)((b * p.H + h) * p.L + q);
  float mx = -INFINITY;
  {
    uint32_t r[16], rn[16];
    if (c_begin < c_end) tmem_ld_x16(t_row + c_begin, r);
    for (int c0 = c_begin; c0 < c_end; c0 += 16) {
      tmem_wait_ld();
      if (c0 + 16 < c_end) tmem_ld_x16(t_row + c0 + 16, rn);     // next chunk streams in behind the math
      float m[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(m + 4 * j) = *reinterpret_cast<const float4*>(smask + c0 + 4 * j);
#pragma unroll
      for (int j = 0; j < 16; ++j) mx = fmaxf(mx, fmaf(__uint_as_float(r[j]), sc, m[j]));
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = rn[j];
    }
  }
  sred[half * 128 + row] = mx;
  __syncthreads();
  mx = fmaxf(sred[row], sred[128 + row]);
  float sum = 0.f;
  {
    uint32_t r[16], rn[16];
    if (c_begin < c_end) tmem_ld_x16(t_row + c_begin, r);
    for (int c0 = c_begin; c0 < c_end; c0 += 16) {
      tmem_wait_ld();
      if (c0 + 16 < c_end) tmem_ld_x16(t_row + c0 + 16, rn);
      float m[16], pv[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(m + 4 * j) = *reinterpret_cast<const float4*>(smask + c0 + 4 * j);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        pv[j] = exp2f(fmaf(__uint_as_float(r[j]), sc, m[j] - mx));
        sum += pv[j];
      }
      if (dc.on) {       // the row sum (softmax denominator) is taken before dropout, the mask only multiplies P
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 mm = drop_mult4(dc, drow, (c0 >> 2) + j4);
          pv[4 * j4] *= mm.x; pv[4 * j4 + 1] *= mm.y; pv[4 * j4 + 2] *= mm.z; pv[4 * j4 + 3] *= mm.w;
        }
      }
      store_row8_sw128(sP, row, c0, pv);
      store_row8_sw128(sP, row, c0 + 8, pv + 8);
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = rn[j];
    }
  }
  ssum[half * 128 + row] = sum;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
    const uint32_t aP = smem_u32(sP), aV = smem_u32(sV);
    const int ksteps = p.lk_pad >> 4;
    for (int t = 0; t < ksteps; ++t)
      umma_bf16(tmem, desc_k(aP + (t >> 2) * 16384 + (t & 3) * 32), desc_mn(aV + t * 2048, 16384), idesc, t > 0);
    umma_commit(&bars[2]);
  }
  sum = ssum[row] + ssum[128 + row];
  mbar_wait(&bars[2], 0);
  tc_fence_after();
  {
    const float inv = 1.0f / sum;
    bf16* dst = p.ctx + (long long)(b * p.L + q) * p.d + h * 64 + half * 32;
    uint32_t r[32];
    tmem_ld_x32(t_row + half * 32, r);
    tmem_wait_ld();
    if (q < p.L) {
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(r[s4 * 8 + 0]) * inv, __uint_as_float(r[s4 * 8 + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(r[s4 * 8 + 2]) * inv, __uint_as_float(r[s4 * 8 + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(r[s4 * 8 + 4]) * inv, __uint_as_float(r[s4 * 8 + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(r[s4 * 8 + 6]) * inv, __uint_as_float(r[s4 * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + s4 * 8) = o;
      }
      if (half == 0 && p.lse) p.lse[((long long)b * p.H + h) * p.L + q] = (mx + log2f(sum)) * LN2;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, ncols); }
}

// ====================================================================================================== backward
// TMEM map (512 columns): [0,256) S -> dP -> dQ(64) ; [256,384) dV key-tiles 0,1 ; [384,512) dK key-tiles 0,1
constexpr int ATT_BWD_THREADS = 512;   // 16 warps: TMEM lane quarter (w & 3) x column quarter (w >> 2)

__global__ void __launch_bounds__(ATT_BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmDO,
                const AttnParams p) {
  // 1024-B aligned dynamic smem (SWIZZLE_128B atoms); indexing the __shared__ array directly keeps the address space known to
  // the compiler (LDS/STS instead of generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q4 = warp & 3, part = warp >> 2;     // TMEM lane quarter (one thread per row) / column quarter
  const int k_bytes = p.lk_pad * 128;
  const int n_kt = p.lk_pad > 128 ? 2 : 1;       // 128-key tiles of dK / dV
  const int kt_pad = n_kt * 128;
  const int ps_bytes = n_kt * 2 * 16384;         // P / dS tiles: 64-key blocks of [128 x 128 B]
  uint8_t* sQ0 = smem;                            // Q / dO tiles are double-buffered: tile qt+1 streams in while tile qt is processed
  uint8_t* sDO0 = smem + 2 * 16384;
  uint8_t* sK = smem + 4 * 16384;
  uint8_t* sV = sK + k_bytes;
  uint8_t* sP = sV + k_bytes;                     // P, overwritten in place by dS once dV += P^T dO has completed
  uint8_t* sDS = sP;
  const DropCtx dc = drop_ctx(p.drop);
  uint8_t* sPd = sP + (dc.on ? ps_bytes : 0);     // dropout: the dV MMA reads the MASKED probabilities from a second tile
  float* smask = reinterpret_cast<float*>(sPd + ps_bytes);        // [256]
  float* sDp = smask + 256;                                       // [4][128] partial rowsum(dO o O) per column quarter
  float* scol = sDp + 512;                                        // [3][64] column sums of this head's dQ | dK | dV
  uint64_t* bars = reinterpret_cast<uint64_t*>(scol + 192);       // kv, qdo[0], qdo[1], s, dp, dq, dk
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 7);

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV); tma_prefetch_desc(&tmDO);
    for (int i = 0; i < 7; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_holder, 512);
  for (int j = tid; j < 256; j += ATT_BWD_THREADS)
    smask[j] = (j < p.L) ? (p.mask ? p.mask[(long long)b * p.L + j] * LOG2E : 0.f) : -INFINITY;
  if (tid < 192) scol[tid] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;
  const uint32_t t_row = tmem + ((uint32_t)(q4 * 32) << 16);
  // key columns are split between the 4 column quarters at multiples of 16, balancing the VALID keys; the last quarter also
  // zero-fills [lk_pad, kt_pad)
  const int c_begin = ((p.lk_pad * part / 4) + 15) & ~15;
  const int c_end = part == 3 ? kt_pad : min(kt_pad, ((p.lk_pad * (part + 1) / 4) + 15) & ~15);
  const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sPd), aDS = smem_u32(sDS);
  const float sc = p.scale * LOG2E;
  const int ksteps = p.lk_pad >> 4;

  if (tid == 0) {
    mbar_expect_tx(&bars[0], 2 * k_bytes);
    tma_load_2d(sK, &tmKV, &bars[0], p.d + h * 64, b * p.L);
    tma_load_2d(sV, &tmKV, &bars[0], 2 * p.d + h * 64, b * p.L);
    mbar_expect_tx(&bars[1], 2 * 16384);
    tma_load_2d(sQ0, &tmQ, &bars[1], h * 64, b * p.L);
    tma_load_2d(sDO0, &tmDO, &bars[1], h * 64, b * p.L);
  }

  for (int qt = 0; qt < p.q_tiles; ++qt) {
    const uint32_t ph = qt & 1;
    const int buf = qt & 1;
    const uint32_t aQ = smem_u32(sQ0 + buf * 16384), aDO = smem_u32(sDO0 + buf * 16384);
    const int row = q4 * 32 + lane;
    const int q = qt * 128 + row;
    const bool qvalid = q < p.L;
    const uint32_t drow = (uint32_t)((b * p.H + h) * p.L + q);
    // ---- loads + S = Q K^T
    if (tid == 0) {
      if (qt + 1 < p.q_tiles) {   // prefetch the next query tile into the other buffer (its last reader was tile qt-1's dK MMA)
        if (qt >= 1) mbar_wait(&bars[6], (qt - 1) & 1);
        mbar_expect_tx(&bars[1 + (buf ^ 1)], 2 * 16384);
        tma_load_2d(sQ0 + (buf ^ 1) * 16384, &tmQ, &bars[1 + (buf ^ 1)], h * 64, b * p.L + (qt + 1) * 128);
        tma_load_2d(sDO0 + (buf ^ 1) * 16384, &tmDO, &bars[1 + (buf ^ 1)], h * 64, b * p.L + (qt + 1) * 128);
      }
      if (qt == 0) mbar_wait(&bars[0], 0);
      mbar_wait(&bars[1 + buf], (qt >> 1) & 1);
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, p.lk_pad, 0, 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, desc_k(aQ + k * 32), desc_k(aK + k * 32), idesc, k > 0);
      umma_commit(&bars[3]);
    }
    // D = rowsum(dO o O): each column quarter takes 16 of the 64 head columns (no redundant global loads), partials meet in shared
    // memory at the barrier between the two passes; the row's LSE comes straight from global memory while the MMA runs
    float lse2 = 0.f;
    {
      float dpart = 0.f;
      if (qvalid) {
        const uint4* po = reinterpret_cast<const uint4*>(p.ctx_in + (long long)(b * p.L + q) * p.d + h * 64 + part * 16);
        const uint4* pd = reinterpret_cast<const uint4*>(p.dctx + (long long)(b * p.L + q) * p.d + h * 64 + part * 16);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          uint4 o = po[i], g = pd[i];
          const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&o);
          const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 a = __bfloat1622float2(o2[j]), c = __bfloat1622float2(g2[j]);
            dpart += a.x * c.x + a.y * c.y;
          }
        }
        lse2 = p.lse[((long long)b * p.H + h) * p.L + q] * LOG2E;
      }
      sDp[part * 128 + row] = dpart;
    }
    mbar_wait(&bars[3], ph);
    tc_fence_after();
    // ---- pass 1: P = exp(S - lse)  (rows beyond L and keys beyond L are exactly zero)
    for (int c0 = c_begin; c0 < c_end; c0 += 16) {
      float pv[16];
      if (c0 < p.lk_pad) {
        uint32_t r[16];
        tmem_ld_x16(t_row + c0, r);
        float m[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(m + 4 * j) = *reinterpret_cast<const float4*>(smask + c0 + 4 * j);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) pv[j] = qvalid ? exp2f(fmaf(__uint_as_float(r[j]), sc, m[j] - lse2)) : 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) pv[j] = 0.f;
      }
      store_row8_sw128(sP, row, c0, pv);
      store_row8_sw128(sP, row, c0 + 8, pv + 8);
      if (dc.on) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 mm = drop_mult4(dc, drow, (c0 >> 2) + j4);
          pv[4 * j4] *= mm.x; pv[4 * j4 + 1] *= mm.y; pv[4 * j4 + 2] *= mm.z; pv[4 * j4 + 3] *= mm.w;
        }
        store_row8_sw128(sPd, row, c0, pv);
        store_row8_sw128(sPd, row, c0 + 8, pv + 8);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- dP = dO V^T (overwrites S) ; dV += P^T dO
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc_dp = umma_idesc_bf16(128, p.lk_pad, 0, 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, desc_k(aDO + k * 32), desc_k(aV + k * 32), idesc_dp, k > 0);
      const uint32_t idesc_t = umma_idesc_bf16(128, 64, 1, 1);
      for (int mt = 0; mt < n_kt; ++mt)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_bf16(tmem + 256 + mt * 64, desc_mn(aP + mt * 32768 + ks * 2048, 16384), desc_mn(aDO + ks * 2048, 16384), idesc_t,
                    (qt > 0 || ks > 0) ? 1u : 0u);
      umma_commit(&bars[4]);
    }
    mbar_wait(&bars[4], ph);
    tc_fence_after();
    // ---- pass 2: dS = scale * P o (dP - D)
    const float Dq = sDp[row] + sDp[128 + row] + sDp[256 + row] + sDp[384 + row];
    for (int c0 = c_begin; c0 < c_end; c0 += 16) {
      float dv[16];
      if (c0 < p.lk_pad) {
        uint32_t r[16];
        tmem_ld_x16(t_row + c0, r);
        uint4 pa = *reinterpret_cast<const uint4*>(sP + (c0 >> 6) * 16384 + sw128_offset(row, (c0 & 63) >> 3));
        uint4 pb = *reinterpret_cast<const uint4*>(sP + (c0 >> 6) * 16384 + sw128_offset(row, ((c0 + 8) & 63) >> 3));
        tmem_wait_ld();
        const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&pa);
        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&pb);
        if (dc.on) {     // dP = dP~ o mask (P~ = P o mask is what fed the value matmul)
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 mm = drop_mult4(dc, drow, (c0 >> 2) + j4);
            r[4 * j4] = __float_as_uint(__uint_as_float(r[4 * j4]) * mm.x); r[4 * j4 + 1] = __float_as_uint(__uint_as_float(r[4 * j4 + 1]) * mm.y);
            r[4 * j4 + 2] = __float_as_uint(__uint_as_float(r[4 * j4 + 2]) * mm.z); r[4 * j4 + 3] = __float_as_uint(__uint_as_float(r[4 * j4 + 3]) * mm.w);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 x = __bfloat1622float2(a2[j]), y = __bfloat1622float2(b2[j]);
          dv[2 * j] = p.scale * x.x * (__uint_as_float(r[2 * j]) - Dq);
          dv[2 * j + 1] = p.scale * x.y * (__uint_as_float(r[2 * j + 1]) - Dq);
          dv[8 + 2 * j] = p.scale * y.x * (__uint_as_float(r[8 + 2 * j]) - Dq);
          dv[8 + 2 * j + 1] = p.scale * y.y * (__uint_as_float(r[8 + 2 * j + 1]) - Dq);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) dv[j] = 0.f;
      }
      store_row8_sw128(sDS, row, c0, dv);
      store_row8_sw128(sDS, row, c0 + 8, dv + 8);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- dQ = dS K (overwrites dP) ; dK += dS^T Q
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc_dq = umma_idesc_bf16(128, 64, 0, 1);
      for (int t = 0; t < ksteps; ++t)
        umma_bf16(tmem, desc_k(aDS + (t >> 2) * 16384 + (t & 3) * 32), desc_mn(aK + t * 2048, 16384), idesc_dq, t > 0);
      umma_commit(&bars[5]);     // dQ only: the threads read it out while the dK MMAs below still run
      const uint32_t idesc_t = umma_idesc_bf16(128, 64, 1, 1);
      for (int mt = 0; mt < n_kt; ++mt)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_bf16(tmem + 384 + mt * 64, desc_mn(aDS + mt * 32768 + ks * 2048, 16384), desc_mn(aQ + ks * 2048, 16384), idesc_t,
                    (qt > 0 || ks > 0) ? 1u : 0u);
      // dK reads the dS and Q tiles: the next query tile overwrites them only after ITS S = Q K^T has completed (pass 1), and MMAs of
      // one thread complete in issue order, so that wait covers this one too; the epilogue waits on bars[6] explicitly
      umma_commit(&bars[6]);
    }
    mbar_wait(&bars[5], ph);
    tc_fence_after();
    {
      bf16* dst = p.dqkv + (long long)(b * p.L + q) * (3 * p.d) + h * 64 + part * 16;
      uint32_t r[16];
      tmem_ld_x16(t_row + part * 16, r);
      tmem_wait_ld();
      if (p.dqkv_colsum) {    // rows of invalid queries are exactly zero (their dS rows are)
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        const float cs = colsum16(v, lane);
        if (!(lane & 1)) atomicAdd(&scol[part * 16 + (lane >> 1)], cs);
      }
      if (qvalid) {
#pragma unroll
        for (int s4 = 0; s4 < 2; ++s4) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(r[s4 * 8 + 0]), __uint_as_float(r[s4 * 8 + 1]));
          o.y = pack_bf16x2(__uint_as_float(r[s4 * 8 + 2]), __uint_as_float(r[s4 * 8 + 3]));
          o.z = pack_bf16x2(__uint_as_float(r[s4 * 8 + 4]), __uint_as_float(r[s4 * 8 + 5]));
          o.w = pack_bf16x2(__uint_as_float(r[s4 * 8 + 6]), __uint_as_float(r[s4 * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + s4 * 8) = o;
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // TMEM [0,256) and the Q/dO/P/dS tiles are free for the next query tile
    tc_fence_after();
  }

  // ---- epilogue: dK, dV rows (one thread per key; quarters 0,1 write the two 32-column halves of dV, quarters 2,3 of dK)
  mbar_wait(&bars[6], (p.q_tiles - 1) & 1);
  tc_fence_after();
  for (int mt = 0; mt < n_kt; ++mt) {
    const int key = mt * 128 + q4 * 32 + lane;
    const bool kvalid = key < p.L;
    const int which = part >> 1, ch = part & 1;
    bf16* dst = p.dqkv + (long long)(b * p.L + key) * (3 * p.d) + (which == 0 ? 2 : 1) * p.d + h * 64 + ch * 32;
    const uint32_t tcol = (which == 0 ? 256 : 384) + mt * 64 + ch * 32;
    uint32_t r[32];
    tmem_ld_x32(t_row + tcol, r);
    tmem_wait_ld();
    if (p.dqkv_colsum) {      // rows of invalid / padded keys are exactly zero (their P and dS columns are)
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      const float cs = colsum32(v, lane);
      atomicAdd(&scol[(which == 0 ? 128 : 64) + ch * 32 + lane], cs);
    }
    if (kvalid) {
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(r[s4 * 8 + 0]), __uint_as_float(r[s4 * 8 + 1]));
        o.y = pack_bf16x2(__uint_as_float(r[s4 * 8 + 2]), __uint_as_float(r[s4 * 8 + 3]));
        o.z = pack_bf16x2(__uint_as_float(r[s4 * 8 + 4]), __uint_as_float(r[s4 * 8 + 5]));
        o.w = pack_bf16x2(__uint_as_float(r[s4 * 8 + 6]), __uint_as_float(r[s4 * 8 + 7]));
        *reinterpret_cast<uint4*>(dst + s4 * 8) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.dqkv_colsum && tid < 192) atomicAdd(p.dqkv_colsum + (tid >> 6) * p.d + h * 64 + (tid & 63), scol[tid]);
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------------ backward, L <= 128
// One 128-query tile and one 128-key tile per (batch, head): the whole chain is latency bound, so this variant (a) needs only 256
// TMEM columns and ~110 KB of shared memory so that TWO CTAs share an SM and overlap each other's waits, and (b) computes S and dP
// side by side so that P and dS come out of ONE pass over the accumulators (no second TMEM sweep, no P re-read, 3 instead of 6
// block-wide syncs).  TMEM map (256 columns): [0,128) S -> dQ(64) | dK(64) ; [128,256) dP -> dV(64).
// 256 threads: TMEM lane quarter (w & 3) = one thread per query / key row, column half (w >> 2).
constexpr int ATT_BWD_SMALL_THREADS = 256;

__global__ void __launch_bounds__(ATT_BWD_SMALL_THREADS, 2)
attn_bwd_small_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmDO,
                      const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q4 = warp & 3, half = warp >> 2;
  const int k_bytes = p.lk_pad * 128;
  uint8_t* sQ = smem;
  uint8_t* sDO = smem + 16384;
  uint8_t* sK = smem + 32768;
  uint8_t* sV = sK + k_bytes;
  uint8_t* sP = sV;                               // (masked) probabilities, over V once dP = dO V^T has completed
  uint8_t* sDS = sP + 32768;
  float* smask = reinterpret_cast<float*>(sDS + 32768);            // [128]
  float* sDp = smask + 128;                                        // [2][128] partial rowsum(dO o O) per column half
  float* scol = sDp + 256;                                         // [3][64] column sums of this head's dQ | dK | dV
  uint64_t* bars = reinterpret_cast<uint64_t*>(scol + 192);        // loads, s/dp, grads
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 3);
  const DropCtx dc = drop_ctx(p.drop);

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV); tma_prefetch_desc(&tmDO);
    for (int i = 0; i < 3; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_holder, 256);
  if (tid < 128) smask[tid] = (tid < p.L) ? (p.mask ? p.mask[(long long)b * p.L + tid] * LOG2E : 0.f) : -INFINITY;
  if (tid < 192) scol[tid] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;
  const uint32_t t_row = tmem + ((uint32_t)(q4 * 32) << 16);
  const uint32_t aQ = smem_u32(sQ), aDO = smem_u32(sDO), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP), aDS = smem_u32(sDS);
  const float sc = p.scale * LOG2E;
  const int ksteps = p.lk_pad >> 4;
  const int row = q4 * 32 + lane;
  const bool qvalid = row < p.L;
  const uint32_t drow = (uint32_t)((b * p.H + h) * p.L + row);

  if (tid == 0) {
    mbar_expect_tx(&bars[0], 2 * k_bytes + 2 * 16384);
    tma_load_2d(sK, &tmKV, &bars[0], p.d + h * 64, b * p.L);
    tma_load_2d(sV, &tmKV, &bars[0], 2 * p.d + h * 64, b * p.L);
    tma_load_2d(sQ, &tmQ, &bars[0], h * 64, b * p.L);
    tma_load_2d(sDO, &tmDO, &bars[0], h * 64, b * p.L);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, p.lk_pad, 0, 0);
    // S = Q K^T and dP = dO V^T are independent accumulators: alternating their k-steps hides the accumulate latency of each chain
    // (back-to-back MMAs into one accumulator were measured at ~90 cycles each, whatever N; attention_bwd2.cu)
    {
      const uint32_t dQ_ = umma_desc_lo(aQ, 16), dK_ = umma_desc_lo(aK, 16), dDO_ = umma_desc_lo(aDO, 16), dV_ = umma_desc_lo(aV, 16);
#pragma unroll
      for (int k = 0; k < 4; ++k) {       // descriptors advance by one add on the low word (common.cuh: umma_bf16_lh)
        umma_bf16_lh(tmem, dQ_ + 2 * k, dK_ + 2 * k, idesc, k > 0);
        umma_bf16_lh(tmem + 128, dDO_ + 2 * k, dV_ + 2 * k, idesc, k > 0);
      }
    }
    umma_commit(&bars[1]);
  }
  // D = rowsum(dO o O): each column half takes 32 of the 64 head columns; the loads overlap the TMA + MMA latency
  float lse2 = 0.f;
  {
    float dpart = 0.f;
    if (qvalid) {
      const uint4* po = reinterpret_cast<const uint4*>(p.ctx_in + (long long)(b * p.L + row) * p.d + h * 64 + half * 32);
      const uint4* pd = reinterpret_cast<const uint4*>(p.dctx + (long long)(b * p.L + row) * p.d + h * 64 + half * 32);
      uint4 o[4], g[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { o[i] = po[i]; g[i] = pd[i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&o[i]);
        const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&g[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 a = __bfloat1622float2(o2[j]), c = __bfloat1622float2(g2[j]);
          dpart += a.x * c.x + a.y * c.y;
        }
      }
      lse2 = p.lse[((long long)b * p.H + h) * p.L + row] * LOG2E;
    }
    sDp[half * 128 + row] = dpart;
  }
  __syncthreads();
  const float Dq = sDp[row] + sDp[128 + row];
  // key columns split between the two halves at a multiple of 16; the upper half also zero-fills [lk_pad, 128)
  const int split = min(p.lk_pad, ((p.lk_pad >> 1) + 15) & ~15);
  const int c_begin = half ? split : 0;
  const int c_end = half ? 128 : split;
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  // ---- single pass: P = exp(S - lse) [o mask], dS = scale * P o (dP [o mask] - D)
  for (int c0 = c_begin; c0 < c_end; c0 += 16) {
    float pv[16], dv[16];
    if (c0 < p.lk_pad) {
      uint32_t rs[16], rp[16];
      tmem_ld_x16(t_row + c0, rs);
      tmem_ld_x16(t_row + 128 + c0, rp);
      float m[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(m + 4 * j) = *reinterpret_cast<const float4*>(smask + c0 + 4 * j);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) pv[j] = (qvalid && !(p.causal && c0 + j > row)) ? exp2f(fmaf(__uint_as_float(rs[j]), sc, m[j] - lse2)) : 0.f;
      if (dc.on) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 mm = drop_mult4(dc, drow, (c0 >> 2) + j4);
          const float mk[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int j = 4 * j4 + t;
            dv[j] = p.scale * pv[j] * (__uint_as_float(rp[j]) * mk[t] - Dq);     // dP = dP~ o mask
            pv[j] *= mk[t];                                                       // P~ = P o mask feeds dV = P~^T dO
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) dv[j] = p.scale * pv[j] * (__uint_as_float(rp[j]) - Dq);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) { pv[j] = 0.f; dv[j] = 0.f; }
    }
    store_row8_sw128(sP, row, c0, pv);
    store_row8_sw128(sP, row, c0 + 8, pv + 8);
    store_row8_sw128(sDS, row, c0, dv);
    store_row8_sw128(sDS, row, c0 + 8, dv + 8);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  // ---- dQ = dS K -> [0,64) ; dK = dS^T Q -> [64,128) ; dV = P~^T dO -> [128,192)
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc_dq = umma_idesc_bf16(128, 64, 0, 1);
    const uint32_t idesc_t = umma_idesc_bf16(128, 64, 1, 1);
    const uint32_t mP = umma_desc_lo(aP, 16384), mDS = umma_desc_lo(aDS, 16384), mDO = umma_desc_lo(aDO, 16384), mQ = umma_desc_lo(aQ, 16384);
    const uint32_t mK = umma_desc_lo(aK, 16384), kDS = umma_desc_lo(aDS, 16);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {       // three independent accumulators, k-steps interleaved
      umma_bf16_lh(tmem + 64, mDS + 128 * ks, mQ + 128 * ks, idesc_t, ks > 0);
      umma_bf16_lh(tmem + 128, mP + 128 * ks, mDO + 128 * ks, idesc_t, ks > 0);
      if (ks < ksteps) umma_bf16_lh(tmem, kDS + (ks >> 2) * 1024 + (ks & 3) * 2, mK + 128 * ks, idesc_dq, ks > 0);
    }
    umma_commit(&bars[2]);
  }
  mbar_wait(&bars[2], 0);
  tc_fence_after();
  {   // rows are queries for dQ and keys for dK / dV; both ranges are [0, L).  tcgen05.ld is warp-collective: no divergence around it
    bf16* dst = p.dqkv + (long long)(b * p.L + row) * (3 * p.d) + h * 64 + half * 32;
#pragma unroll 1
    for (int w = 0; w < 3; ++w) {
      uint32_t r[32];
      tmem_ld_x32(t_row + w * 64 + half * 32, r);
      tmem_wait_ld();
      if (p.dqkv_colsum) {    // invalid query / key rows are exactly zero
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const float cs = colsum32(v, lane);
        atomicAdd(&scol[w * 64 + half * 32 + lane], cs);
      }
      if (qvalid) {
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(r[s4 * 8 + 0]), __uint_as_float(r[s4 * 8 + 1]));
          o.y = pack_bf16x2(__uint_as_float(r[s4 * 8 + 2]), __uint_as_float(r[s4 * 8 + 3]));
          o.z = pack_bf16x2(__uint_as_float(r[s4 * 8 + 4]), __uint_as_float(r[s4 * 8 + 5]));
          o.w = pack_bf16x2(__uint_as_float(r[s4 * 8 + 6]), __uint_as_float(r[s4 * 8 + 7]));
          *reinterpret_cast<uint4*>(dst + w * p.d + s4 * 8) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.dqkv_colsum && tid < 192) atomicAdd(p.dqkv_colsum + (tid >> 6) * p.d + h * 64 + (tid & 63), scol[tid]);
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

static int check_shapes(const char* who, int B, int L, int H, int d, int lmax = 256) {
  if (B <= 0 || L <= 0 || H <= 0 || d != H * 64) { set_error("%s: need d == 64*H (B=%d L=%d H=%d d=%d)", who, B, L, H, d); return CLIPK_ERR_ARG; }
  if (L > lmax) { set_error("%s: sequence length %d > 256 is not supported by the one-shot kernel yet", who, L); return CLIPK_ERR_UNSUPPORTED; }
  return 0;
}

}  // namespace clipk

using namespace clipk;

static int attention_fwd_impl(const void* qkv, const float* key_mask, void* ctx, float* lse, int B, int L, int H, int d, const clipk_dropout_t* drop,
                              int causal, cudaStream_t stream);
static int attention_bwd_impl(const void* qkv, const float* key_mask, const void* ctx, const float* lse, const void* dctx, void* dqkv, float* dqkv_colsum,
                              int B, int L, int H, int d, const clipk_dropout_t* drop, int causal, cudaStream_t stream);

extern "C" int clipk_attention_fwd(const void* qkv, const float* key_mask, void* ctx, float* lse, int B, int L, int H, int d,
                                   const clipk_dropout_t* drop, cudaStream_t stream) {
  return attention_fwd_impl(qkv, key_mask, ctx, lse, B, L, H, d, drop, 0, stream);
}
extern "C" int clipk_attention_causal_fwd(const void* qkv, void* ctx, float* lse, int B, int L, int H, int d, cudaStream_t stream) {
  return attention_fwd_impl(qkv, nullptr, ctx, lse, B, L, H, d, nullptr, 1, stream);
}
extern "C" int clipk_attention_bwd(const void* qkv, const float* key_mask, const void* ctx, const float* lse, const void* dctx, void* dqkv,
                                   float* dqkv_colsum, int B, int L, int H, int d, const clipk_dropout_t* drop, cudaStream_t stream) {
  return attention_bwd_impl(qkv, key_mask, ctx, lse, dctx, dqkv, dqkv_colsum, B, L, H, d, drop, 0, stream);
}
extern "C" int clipk_attention_causal_bwd(const void* qkv, const void* ctx, const float* lse, const void* dctx, void* dqkv, float* dqkv_colsum, int B,
                                          int L, int H, int d, cudaStream_t stream) {
  return attention_bwd_impl(qkv, nullptr, ctx, lse, dctx, dqkv, dqkv_colsum, B, L, H, d, nullptr, 1, stream);
}

static int attention_fwd_impl(const void* qkv, const float* key_mask, void* ctx, float* lse, int B, int L, int H, int d, const clipk_dropout_t* drop,
                              int causal, cudaStream_t stream) {
  int rc = check_shapes("attention_fwd", B, L, H, d, 272);
  if (rc) return rc;
  AttnParams p{};
  p.causal = causal;
  p.B = B; p.L = L; p.H = H; p.d = d;
  p.lk_pad = (L + 15) & ~15;
  p.q_tiles = (L + 127) / 128;
  p.scale = 0.125f;
  p.mask = key_mask; p.ctx = (bf16*)ctx; p.lse = lse;
  p.drop = make_drop_arg(drop);
  // default: the persistent warp-specialised pipeline (attention_fwd2.cu); CLIPK_ATTN_V1=1 selects the first-generation kernel (A/B runs)
  { const char* ev = getenv("CLIPK_ATTN_DBG_PTR"); if (ev) p.dbg = reinterpret_cast<long long*>(strtoull(ev, nullptr, 0)); }
  { const char* ev = getenv("CLIPK_ATTN_V1"); if (!(ev && ev[0] == '1')) return attention_fwd2(qkv, p, stream); }
  if (L > 256 || causal) { set_error("attention_fwd: the first-generation kernel handles L <= 256 without a causal mask (L=%d)", L); return CLIPK_ERR_UNSUPPORTED; }
  CUtensorMap tQ, tKV;
  if ((rc = make_tmap_2d_bf16(&tQ, qkv, 3ull * d, (uint64_t)B * L, 3ull * d, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tKV, qkv, 3ull * d, (uint64_t)B * L, 3ull * d, 64, p.lk_pad))) return rc;
  const int k_bytes = p.lk_pad * 128;
  const int p_bytes = ((p.lk_pad + 63) >> 6) * 16384;
  const int v_off = p_bytes > 16384 + k_bytes ? p_bytes : 16384 + k_bytes;
  const int smem = v_off + k_bytes + 1024 + 2048 + 64 + 1024;
  static int configured = 0;
  if (configured < smem) {
    CLIPK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  attn_fwd_kernel<<<dim3(p.q_tiles, H, B), ATT_THREADS, smem, stream>>>(tQ, tKV, p);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

static int attention_bwd_impl(const void* qkv, const float* key_mask, const void* ctx, const float* lse, const void* dctx, void* dqkv, float* dqkv_colsum,
                              int B, int L, int H, int d, const clipk_dropout_t* drop, int causal, cudaStream_t stream) {
  int rc = check_shapes("attention_bwd", B, L, H, d);
  if (rc) return rc;
  if (causal && L > 128) { set_error("attention_bwd: the causal variant covers L <= 128 (text towers)"); return CLIPK_ERR_UNSUPPORTED; }
  AttnParams p{};
  p.causal = causal;
  p.B = B; p.L = L; p.H = H; p.d = d;
  p.lk_pad = (L + 15) & ~15;
  p.q_tiles = (L + 127) / 128;
  p.scale = 0.125f;
  p.mask = key_mask; p.lse = const_cast<float*>(lse);
  p.ctx_in = (const bf16*)ctx; p.dctx = (const bf16*)dctx; p.dqkv = (bf16*)dqkv; p.dqkv_colsum = dqkv_colsum;
  p.drop = make_drop_arg(drop);
  if (L <= 128) {
    CUtensorMap tQ, tKV, tDO;
    if ((rc = make_tmap_2d_bf16(&tQ, qkv, 3ull * d, (uint64_t)B * L, 3ull * d, 64, 128))) return rc;
    if ((rc = make_tmap_2d_bf16(&tKV, qkv, 3ull * d, (uint64_t)B * L, 3ull * d, 64, p.lk_pad))) return rc;
    if ((rc = make_tmap_2d_bf16(&tDO, dctx, (uint64_t)d, (uint64_t)B * L, (uint64_t)d, 64, 128))) return rc;
    const int smem = 32768 + p.lk_pad * 128 + 65536 + 512 + 1024 + 768 + 64;
    static int configured_small = 0;
    if (configured_small < smem) {
      CLIPK_CUDA(cudaFuncSetAttribute(attn_bwd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured_small = smem;
    }
    attn_bwd_small_kernel<<<dim3(H, B), ATT_BWD_SMALL_THREADS, smem, stream>>>(tQ, tKV, tDO, p);
    note_launch();
    CLIPK_CUDA(cudaGetLastError());
    return 0;
  }
  if (p.drop.on && L > 128) { set_error("attention_bwd: attention dropout is implemented for L <= 128 (the text tower)"); return CLIPK_ERR_UNSUPPORTED; }
  // default for 128 < L <= 256 without a key mask (the ViT tower): the persistent warp-specialised pipeline (attention_bwd2.cu);
  // CLIPK_ATTN_V1=1 selects the first-generation kernel (A/B runs)
  { const char* ev = getenv("CLIPK_ATTN_DBG_PTR"); if (ev) p.dbg = reinterpret_cast<long long*>(strtoull(ev, nullptr, 0)); }
  { const char* ev = getenv("CLIPK_ATTN_FLAGS"); if (ev) p.flags = atoi(ev); }
  { const char* ev = getenv("CLIPK_ATTN_V1"); if (!(ev && ev[0] == '1') && !key_mask) return attention_bwd2(qkv, p, stream); }
  CUtensorMap tQ, tKV, tDO;
  if ((rc = make_tmap_2d_bf16(&tQ, qkv, 3ull * d, (uint64_t)B * L, 3ull * d, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tKV, qkv, 3ull * d, (uint64_t)B * L, 3ull * d, 64, p.lk_pad))) return rc;
  if ((rc = make_tmap_2d_bf16(&tDO, dctx, (uint64_t)d, (uint64_t)B * L, (uint64_t)d, 64, 128))) return rc;
  const int k_bytes = p.lk_pad * 128;
  const int n_kt = p.lk_pad > 128 ? 2 : 1;
  const int smem = 4 * 16384 + 2 * k_bytes + (n_kt * 2 * 16384) * (p.drop.on ? 2 : 1) + 1024 + 2048 + 768 + 64 + 1024;
  static int configured = 0;
  if (configured < smem) {
    CLIPK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  attn_bwd_kernel<<<dim3(H, B), ATT_BWD_THREADS, smem, stream>>>(tQ, tKV, tDO, p);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}
