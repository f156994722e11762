// Attention backward, second generation (sequences of 129..256 tokens: the ViT tower): ONE persistent warp-specialised CTA per SM.
//
// A (sample, head) item is walked in STEPS (key tile kt, query tile t), kt outer.  Per step
//     S  = Q_t K_kt^T ,  dP = dO_t V_kt^T            (tcgen05, side by side in TMEM)
//     P  = exp2(S * scale - lse) ,  dS = scale * P o (dP - D)      ONE pass over both accumulators, 16 sweep warps
//     dV_kt += P^T dO_t ,  dK_kt += dS^T Q_t ,  dQ_t += dS K_kt     (P / dS as bf16 operands through shared memory)
// so every (query, key) pair is visited once, the exp is evaluated once, and no second TMEM sweep or P re-read is needed (the
// first-generation attn_bwd_kernel made two sweeps per query tile with six block-wide barriers, everything serialised in one CTA).
//
//   warp 0      TMA producer: Q/dO tile pairs through a 3-deep ring, K/V key tiles through a 2-deep ring -- the next item's tiles land
//               while the current item is still being processed
//   warp 1      MMA issuer  : S/dP of step s+1 are issued BEFORE the gradient MMAs of step s, so they run during sweep s
//   warps 4-19  sweep       : 4 TMEM lane quarters x 4 column quarters; the math of sweep s+1 overlaps the gradient MMAs of step s, only
//               the short store burst into the single P/dS buffer waits for them
//   warps 20-23 read-out    : dV_kt / dK_kt after the last query tile of a key tile, dQ_t after the last key tile -> dqkv (bf16)
//
// Fused QKV-bias gradient (no key mask, no dropout on this path): the QUERY block is the column sum of the fp32 dQ accumulators
// (read-out warps); the KEY block is identically zero -- a key bias shifts every score of a row by the same amount and softmax is shift
// invariant -- so nothing is added; the VALUE block is sum_q (sum_k P[q,k]) dO[q,:] = the column sum of dO (rows of P sum to 1), taken by
// the sweep warps from the dO values they load for D anyway (a 16-column butterfly per warp and query tile).
//
// TMEM (512 columns): S [0,128) | dP [128,256) | dV [256,320) | dK [320,384) | dQ_0 [384,448) | dQ_1 [448,512).
// Replaces autograd through nn.MultiheadAttention's SDPA (modeling_chineseclip.py:188,198-200).
#include "common.cuh"
#include "attention_common.cuh"
#include "../../include/clipk.h"

namespace clipk {

constexpr int B2_THREADS = 768;
constexpr int B2_SWEEP_WARP0 = 4;
constexpr int B2_OUT_WARP0 = 20;

struct Bwd2Smem {
  int nq;                 // Q/dO ring depth
  int qdo_off;            // nq x 32 KB: [Q tile 16 KB | dO tile 16 KB]
  int kv_slot;            // bytes of one K (or V) tile slot
  int kv_off;             // 2 x [K tile | V tile]
  int p_off, ds_off;      // 32 KB each
  int dl_off;             // D [2][128], lse2 [2][128], Dpart [4][128]
  int col_off;            // [3][64] column sums
  int bar_off, total;
};
__host__ __device__ inline Bwd2Smem bwd2_layout(int kn0) {
  Bwd2Smem s;
  s.kv_slot = kn0 * 128;
  s.nq = (3 * 32768 + 4 * s.kv_slot + 65536 + 8192 <= 232448 - 1024) ? 3 : 2;
  s.qdo_off = 0;
  s.kv_off = s.nq * 32768;
  s.p_off = s.kv_off + 4 * s.kv_slot;
  s.ds_off = s.p_off + 32768;
  s.dl_off = s.ds_off + 32768;
  s.col_off = s.dl_off + 8 * 128 * 4;
  s.bar_off = s.col_off + 192 * 4;
  s.total = s.bar_off + 256;
  return s;
}

// timeline probe (diagnostics): event e of step g of CTA 0
#define B2_DBG(g, e) do { if (p.dbg && blockIdx.x == 0 && (g) < 64) p.dbg[(g) * 16 + (e)] = clock64(); } while (0)

// 8 consecutive bf16 (cols col..col+7, col % 8 == 0) of row `row` in a [128-row x 64-col]-blocked SWIZZLE_128B tile
__device__ __forceinline__ void st_row8_packed(uint8_t* tile_base, int row, int col, const uint32_t* pk) {
  *reinterpret_cast<uint4*>(tile_base + (col >> 6) * 16384 + sw128_offset(row, (col & 63) >> 3)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// KN0 / KN1: keys of key tile 0 / 1 (multiples of 16, <= 128; KN1 == 0: one key tile).  T: query tiles (1 or 2).
template <int KN0, int KN1, int T>
__global__ void __launch_bounds__(B2_THREADS, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmDO,
                 const AttnParams p) {
  constexpr int KT = KN1 > 0 ? 2 : 1;
  constexpr int STEPS = KT * T;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const Bwd2Smem lay = bwd2_layout(KN0);
  const int NQ = lay.nq;
  uint8_t* sP = smem + lay.p_off;
  uint8_t* sDS = smem + lay.ds_off;
  float* sD = reinterpret_cast<float*>(smem + lay.dl_off);        // [2][128] rowsum(dO o O) per query tile
  float* sLse = sD + 256;                                           // [2][128] lse in log2 units
  float* sDpart = sLse + 256;                                       // [4][128] partial D per column quarter
  float* scol = reinterpret_cast<float*>(smem + lay.col_off);       // [3][64] column sums of this item's dQ | dK | dV
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
  uint64_t* qdo_full = bars;          // [3]
  uint64_t* qdo_empty = bars + 3;     // [3]
  uint64_t* kv_full = bars + 6;       // [2]
  uint64_t* kv_empty = bars + 8;      // [2]
  uint64_t* sdp_full = bars + 10;     // S and dP of a step are complete
  uint64_t* sdp_free = bars + 11;     // every sweep thread holds its S / dP slice in registers (16 warps)
  uint64_t* pds_full = bars + 12;     // P and dS of a step are in shared memory (16 warps)
  uint64_t* pds_empty = bars + 13;    // the gradient MMAs of a step have read P / dS
  uint64_t* kvacc_full = bars + 14;   // dV_kt / dK_kt complete (last query tile of the key tile)
  uint64_t* kvacc_free = bars + 15;   // ... and read out (4 warps)
  uint64_t* dq_full = bars + 16;      // [2] dQ_t complete (last key tile)
  uint64_t* dq_free = bars + 18;      // [2] ... and read out (4 warps)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 20);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_items = p.B * p.H;
  const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV); tma_prefetch_desc(&tmDO);
    for (int s = 0; s < 3; ++s) { mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); mbar_init(&dq_full[s], 1); mbar_init(&dq_free[s], 4); }
    mbar_init(sdp_full, 1); mbar_init(sdp_free, 16); mbar_init(pds_full, 16); mbar_init(pds_empty, 1);
    mbar_init(kvacc_full, 1); mbar_init(kvacc_free, 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, 512);
  if (tid < 192) scol[tid] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;

  if (warp < 4) {
    reg_dec<40>();
    if (warp == 0) {
      // ========================================================================================== TMA producer
      // Tile loads in the order the steps first need them: per item Q/dO pair 0, K/V tile 0, Q/dO pair 1, K/V tile 1.
      if (lane == 0) {
        int nq_i = 0, nkv_i = 0;            // running counts of Q/dO pairs and K/V tiles loaded
        for (int jj = 0; jj < my_items; ++jj) {
          const int item = blockIdx.x + jj * gridDim.x;
          const int b = item / p.H, h = item - b * p.H;
          for (int u = 0; u < (T > KT ? T : KT); ++u) {
            if (u < T) {
              const int slot = nq_i % NQ; const uint32_t ph = (nq_i / NQ) & 1;
              mbar_wait(&qdo_empty[slot], ph ^ 1);
              uint8_t* dst = smem + lay.qdo_off + slot * 32768;
              mbar_expect_tx(&qdo_full[slot], 32768);
              tma_load_2d(dst, &tmQ, &qdo_full[slot], h * 64, b * p.L + u * 128);
              tma_load_2d(dst + 16384, &tmDO, &qdo_full[slot], h * 64, b * p.L + u * 128);
              ++nq_i;
            }
            if (u < KT) {
              const int slot = nkv_i & 1; const uint32_t ph = (nkv_i >> 1) & 1;
              mbar_wait(&kv_empty[slot], ph ^ 1);
              uint8_t* dst = smem + lay.kv_off + slot * 2 * lay.kv_slot;
              mbar_expect_tx(&kv_full[slot], 2 * lay.kv_slot);
              tma_load_2d(dst, &tmKV, &kv_full[slot], p.d + h * 64, b * p.L + u * KN0);
              tma_load_2d(dst + lay.kv_slot, &tmKV, &kv_full[slot], 2 * p.d + h * 64, b * p.L + u * KN0);
              ++nkv_i;
            }
          }
        }
      }
    } else if (warp == 1) {
      // ========================================================================================== MMA issuer (one thread)
      if (elect_one()) {
        const int n_steps = my_items * STEPS;
        // step g -> (item ordinal jj, kt, t); Q/dO pair ordinal = jj*T + t; K/V tile ordinal = jj*KT + kt
        auto issue_sdp = [&](int g) {
          const int jj = g / STEPS, r = g - jj * STEPS, kt = r / T, t = r - kt * T;
          const int qi = jj * T + t, ki = jj * KT + kt;
          const int qs = qi % NQ, ks = ki & 1;
          B2_DBG(g, 0);
          if (kt == 0) mbar_wait(&qdo_full[qs], (qi / NQ) & 1);       // first use of this Q/dO pair
          if (t == 0) mbar_wait(&kv_full[ks], (ki >> 1) & 1);          // first use of this K/V tile
          B2_DBG(g, 1);
          mbar_wait(sdp_free, (g & 1) ^ 1);                            // sweep g-1 has pulled S/dP into registers
          tc_fence_after();
          B2_DBG(g, 2);
          const uint32_t aQ = smem_u32(smem + lay.qdo_off + qs * 32768), aDO = aQ + 16384;
          const uint32_t aK = smem_u32(smem + lay.kv_off + ks * 2 * lay.kv_slot), aV = aK + lay.kv_slot;
          const uint32_t idesc = kt == 0 ? umma_idesc_bf16(128, KN0, 0, 0) : umma_idesc_bf16(128, KN1 > 0 ? KN1 : 16, 0, 0);
          // K-major operands: k-step = 16 head dims = 32 B -> +2 on the descriptor's low word
          const uint32_t dQ_ = umma_desc_lo(aQ, 16), dK_ = umma_desc_lo(aK, 16), dDO_ = umma_desc_lo(aDO, 16), dV_ = umma_desc_lo(aV, 16);
#pragma unroll
          for (int k = 0; k < 4; ++k) {       // S and dP: independent accumulators, k-steps alternate
            umma_bf16_lh(tmem, dQ_ + 2 * k, dK_ + 2 * k, idesc, k > 0);
            umma_bf16_lh(tmem + 128, dDO_ + 2 * k, dV_ + 2 * k, idesc, k > 0);
          }
          umma_commit(sdp_full);
        };
        auto issue_grads = [&](int g) {
          const int jj = g / STEPS, r = g - jj * STEPS, kt = r / T, t = r - kt * T;
          const int qi = jj * T + t, ki = jj * KT + kt;
          const int qs = qi % NQ, ks = ki & 1;
          B2_DBG(g, 3);
          mbar_wait(pds_full, g & 1);
          B2_DBG(g, 4);
          // accumulators must have been read out: dV/dK at the first query tile of a key tile, dQ_t at the first key tile
          if (t == 0) mbar_wait(kvacc_free, (ki & 1) ^ 1);
          if (kt == 0) mbar_wait(&dq_free[t], (jj & 1) ^ 1);
          tc_fence_after();
          B2_DBG(g, 5);
          const uint32_t aQ = smem_u32(smem + lay.qdo_off + qs * 32768), aDO = aQ + 16384;
          const uint32_t aK = smem_u32(smem + lay.kv_off + ks * 2 * lay.kv_slot);
          const uint32_t aP = smem_u32(sP), aDS = smem_u32(sDS);
          const int kn16 = (kt == 0 ? KN0 : KN1) >> 4;
          constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64, 1, 1);
          constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64, 0, 1);
          // dV_kt += P^T dO_t ; dK_kt += dS^T Q_t ; dQ_t += dS K_kt : three independent accumulators, k-steps interleaved.
          // MN-major operands (P^T, dS^T, dO, Q, K): k-step = 16 rows of 128 B = 2048 B -> +128; block stride (LBO) 16 KB.
          // dS as the K-major A operand of dQ: k-step = 32 B inside a 64-key block (+2), next block +16 KB (+1024).
          const uint32_t mP = umma_desc_lo(aP, 16384), mDS = umma_desc_lo(aDS, 16384), mDO = umma_desc_lo(aDO, 16384), mQ = umma_desc_lo(aQ, 16384);
          const uint32_t mK = umma_desc_lo(aK, 16384), kDS = umma_desc_lo(aDS, 16);
          const uint32_t acc_kv = t > 0 ? 1u : 0u, acc_q = kt > 0 ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            umma_bf16_lh(tmem + 256, mP + 128 * k, mDO + 128 * k, idesc_t, k > 0 ? 1u : acc_kv);
            umma_bf16_lh(tmem + 320, mDS + 128 * k, mQ + 128 * k, idesc_t, k > 0 ? 1u : acc_kv);
            if (k < kn16) umma_bf16_lh(tmem + 384 + t * 64, kDS + (k >> 2) * 1024 + (k & 3) * 2, mK + 128 * k, idesc_q, k > 0 ? 1u : acc_q);
          }
          if (t == T - 1) umma_commit(kvacc_full);
          umma_commit(pds_empty);
          if (t == T - 1) umma_commit(&kv_empty[ks]);        // K / V tile: last read by this step's dQ MMAs
          if (kt == KT - 1) { umma_commit(&dq_full[t]); umma_commit(&qdo_empty[qs]); }
          B2_DBG(g, 6);
        };
        // program order: S/dP(0) | S/dP(1) grads(0) | S/dP(2) grads(1) | ...
        if (n_steps > 0) issue_sdp(0);
        for (int g = 0; g < n_steps; ++g) {
          if (g + 1 < n_steps) issue_sdp(g + 1);
          issue_grads(g);
        }
      }
    }
  } else if (warp < B2_OUT_WARP0) {
    // ============================================================================================ sweep warps
    reg_inc<96>();
    const int q4 = warp & 3;
    const int cq = (warp - B2_SWEEP_WARP0) >> 2;
    const int row = q4 * 32 + lane;
    const uint32_t t_row = tmem + ((uint32_t)(q4 * 32) << 16);
    const float sc = p.scale * LOG2E;
    int g = 0;
    for (int jj = 0; jj < my_items; ++jj) {
      const int item = blockIdx.x + jj * gridDim.x;
      const int b = item / p.H, h = item - b * p.H;
#pragma unroll 1
      for (int r = 0; r < STEPS; ++r, ++g) {
        const int kt = r / T, t = r - kt * T;
        const int q = t * 128 + row;
        const bool qvalid = q < p.L;
        // 8-column groups of this key tile split over the 4 column quarters (balanced; a warp owns up to 4 groups)
        const int G = (kt == 0 ? KN0 : KN1) >> 3;
        const int gbase = G >> 2, grem = G & 3;
        const int g0 = cq * gbase + (cq < grem ? cq : grem);
        const int ng = gbase + (cq < grem ? 1 : 0);
        float dpart = 0.f, lse2 = 0.f;
        if (kt == 0) {
          float dov[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) dov[j] = 0.f;
          // first visit of this query tile: D = rowsum(dO o O) (16 of the 64 head columns per column quarter) and the row's LSE;
          // the global loads are in flight while S / dP are still being computed
          if (qvalid) {
            const uint4* po = reinterpret_cast<const uint4*>(p.ctx_in + (long long)(b * p.L + q) * p.d + h * 64 + cq * 16);
            const uint4* pd = reinterpret_cast<const uint4*>(p.dctx + (long long)(b * p.L + q) * p.d + h * 64 + cq * 16);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              uint4 o = po[i], gg = pd[i];
              const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&o);
              const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&gg);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 a = __bfloat1622float2(o2[j]), c = __bfloat1622float2(g2[j]);
                dpart += a.x * c.x + a.y * c.y;
                dov[i * 8 + 2 * j] = c.x; dov[i * 8 + 2 * j + 1] = c.y;
              }
            }
            lse2 = p.lse[((long long)b * p.H + h) * p.L + q] * LOG2E;
          }
          if (p.dqkv_colsum) {
            // value-bias gradient: sum_q (sum_k P[q,k]) dO[q,:] = column sums of dO (rows of P sum to 1; no mask / dropout on this path).
            // The dO values are in registers anyway: one 16-column butterfly per warp and query tile, 16 global reductions.
            const float cs = colsum16(dov, lane);
            if (!(lane & 1)) atomicAdd(p.dqkv_colsum + 2 * p.d + h * 64 + cq * 16 + (lane >> 1), cs);
          }
          sDpart[cq * 128 + row] = dpart;
          named_bar_sync(1 + q4, 128);
          const float Dq = (sDpart[row] + sDpart[128 + row]) + (sDpart[256 + row] + sDpart[384 + row]);
          if (cq == 0) { sD[t * 128 + row] = Dq; sLse[t * 128 + row] = lse2; }
          named_bar_sync(1 + q4, 128);         // everyone has read the partials before the next query tile overwrites them
          dpart = Dq;
        } else {
          dpart = sD[t * 128 + row]; lse2 = sLse[t * 128 + row];
        }
        const float Dq = dpart;
        const float nl = qvalid ? -lse2 : -INFINITY;         // rows beyond L: P = exp2(-inf) = 0 exactly
        if (warp == B2_SWEEP_WARP0 && lane == 0) B2_DBG(g, 8);
        mbar_wait(sdp_full, g & 1);
        tc_fence_after();
        if (warp == B2_SWEEP_WARP0 && lane == 0) B2_DBG(g, 9);
        uint32_t ppk[16], dpk[16];                             // packed bf16 pairs of P and dS: 4 groups x 4 words
        const int key0 = kt * KN0;                             // first key of this key tile
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t rs[16], rp[16];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int gi = half * 2 + u;
            if (gi < ng) { TmemIO<8>::ld(t_row + (g0 + gi) * 8, rs + 8 * u); TmemIO<8>::ld(t_row + 128 + (g0 + gi) * 8, rp + 8 * u); }
          }
          tmem_wait_ld();
          if (half == 1) {          // the whole slice is in registers: S / dP may be overwritten by the next step's MMAs
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sdp_free);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int gi = half * 2 + u;
            if (gi < ng) {
              const int col = (g0 + gi) * 8;
#pragma unroll
              for (int c = 0; c < 8; c += 2) {
                // keys beyond L (padding rows of the K / V boxes) get P = 0
                const float m0 = (key0 + col + c < p.L) ? nl : -INFINITY, m1 = (key0 + col + c + 1 < p.L) ? nl : -INFINITY;
                const float p0 = exp2f(fmaf(__uint_as_float(rs[8 * u + c]), sc, m0)), p1 = exp2f(fmaf(__uint_as_float(rs[8 * u + c + 1]), sc, m1));
                const float d0 = p.scale * p0 * (__uint_as_float(rp[8 * u + c]) - Dq), d1 = p.scale * p1 * (__uint_as_float(rp[8 * u + c + 1]) - Dq);
                ppk[gi * 4 + (c >> 1)] = pack_bf16x2(p0, p1);
                dpk[gi * 4 + (c >> 1)] = pack_bf16x2(d0, d1);
              }
            }
          }
        }
        // the gradient MMAs of the previous step must have finished reading the P / dS tiles
        if (warp == B2_SWEEP_WARP0 && lane == 0) B2_DBG(g, 10);
        mbar_wait(pds_empty, (g & 1) ^ 1);
        if (warp == B2_SWEEP_WARP0 && lane == 0) B2_DBG(g, 11);
#pragma unroll
        for (int gi = 0; gi < 4; ++gi)
          if (gi < ng) { st_row8_packed(sP, row, (g0 + gi) * 8, ppk + gi * 4); st_row8_packed(sDS, row, (g0 + gi) * 8, dpk + gi * 4); }
        fence_proxy_async_smem();
        __syncwarp();
        if (warp == B2_SWEEP_WARP0 && lane == 0) B2_DBG(g, 12);
        if (p.dbg && lane == 0 && blockIdx.x == 0 && g < 64) atomicMax(reinterpret_cast<unsigned long long*>(p.dbg) + g * 16 + 13, (unsigned long long)clock64());
        if (lane == 0) mbar_arrive(pds_full);
      }
    }
  } else {
    // ============================================================================================ read-out warps
    reg_dec<56>();
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    const uint32_t t_row = tmem + ((uint32_t)(q4 * 32) << 16);
    auto emit = [&](uint32_t tcol, bf16* dst, bool valid, float* cs) {     // 64 accumulator columns of this thread's row -> bf16 (+ column sums)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r[32];
        tmem_ld_x32(t_row + tcol + hh * 32, r);
        tmem_wait_ld();
        if (cs) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = valid ? __uint_as_float(r[j]) : 0.f;
          const float s = colsum32(v, lane);
          atomicAdd(&cs[hh * 32 + lane], s);
        }
        if (valid) {
#pragma unroll
          for (int s4 = 0; s4 < 4; ++s4) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(r[s4 * 8 + 0]), __uint_as_float(r[s4 * 8 + 1]));
            o.y = pack_bf16x2(__uint_as_float(r[s4 * 8 + 2]), __uint_as_float(r[s4 * 8 + 3]));
            o.z = pack_bf16x2(__uint_as_float(r[s4 * 8 + 4]), __uint_as_float(r[s4 * 8 + 5]));
            o.w = pack_bf16x2(__uint_as_float(r[s4 * 8 + 6]), __uint_as_float(r[s4 * 8 + 7]));
            *reinterpret_cast<uint4*>(dst + hh * 32 + s4 * 8) = o;
          }
        }
      }
    };
    for (int jj = 0; jj < my_items; ++jj) {
      const int item = blockIdx.x + jj * gridDim.x;
      const int b = item / p.H, h = item - b * p.H;
      for (int kt = 0; kt < KT; ++kt) {
        // ---- dV_kt, dK_kt (rows = keys of this key tile)
        const int ki = jj * KT + kt;
        mbar_wait(kvacc_full, ki & 1);
        tc_fence_after();
        const int key = kt * KN0 + row;
        const bool kvalid = row < (kt == 0 ? KN0 : KN1) && key < p.L;
        bf16* dstk = p.dqkv + (long long)(b * p.L + key) * (3 * p.d) + h * 64;
        emit(256, dstk + 2 * p.d, kvalid, nullptr);
        emit(320, dstk + p.d, kvalid, nullptr);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(kvacc_free);
        if (kt == KT - 1) {
          // ---- dQ_t of both query tiles (complete after the last key tile)
          for (int t = 0; t < T; ++t) {
            mbar_wait(&dq_full[t], jj & 1);
            tc_fence_after();
            const int q = t * 128 + row;
            emit(384 + t * 64, p.dqkv + (long long)(b * p.L + q) * (3 * p.d) + h * 64, q < p.L, p.dqkv_colsum ? scol : nullptr);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dq_free[t]);
          }
        }
      }
      // fused QKV-bias gradient of this item: the 4 read-out warps meet, then 192 threads flush and clear the column sums
      if (p.dqkv_colsum) {
        named_bar_sync(5, 128);
        const int i = (warp - B2_OUT_WARP0) * 32 + lane;
        if (i < 64) {      // query block only (key block: identically zero; value block: added by the sweep warps)
          atomicAdd(p.dqkv_colsum + h * 64 + i, scol[i]);
          scol[i] = 0.f;
        }
        named_bar_sync(5, 128);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int KN0, int KN1, int T>
static int launch_bwd2(const void* qkv, const AttnParams& p, cudaStream_t stream) {
  CUtensorMap tQ, tKV, tDO;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tQ, qkv, 3ull * p.d, (uint64_t)p.B * p.L, 3ull * p.d, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tKV, qkv, 3ull * p.d, (uint64_t)p.B * p.L, 3ull * p.d, 64, KN0))) return rc;
  if ((rc = make_tmap_2d_bf16(&tDO, p.dctx, (uint64_t)p.d, (uint64_t)p.B * p.L, (uint64_t)p.d, 64, 128))) return rc;
  const Bwd2Smem lay = bwd2_layout(KN0);
  const int smem = lay.total + 1024;
  static bool configured = false;
  if (!configured) {
    CLIPK_CUDA(cudaFuncSetAttribute(attn_bwd2_kernel<KN0, KN1, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int items = p.B * p.H;
  const int grid = items < sm_count() ? items : sm_count();
  attn_bwd2_kernel<KN0, KN1, T><<<grid, B2_THREADS, smem, stream>>>(tQ, tKV, tDO, p);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

// 128 < L <= 256.  Key tiles: L <= 208 -> 112 + 96 keys (the ViT-B/16 sequence of 197 tokens: 11 padding keys instead of 59);
// otherwise 128 + 128.
int attention_bwd2(const void* qkv, const AttnParams& p, cudaStream_t stream) {
  if (p.L <= 128 || p.L > 256) { set_error("attention_bwd2: 128 < L <= 256 (L=%d)", p.L); return CLIPK_ERR_UNSUPPORTED; }
  if (p.drop.on) { set_error("attention_bwd2: attention dropout is implemented for L <= 128 (the text tower)"); return CLIPK_ERR_UNSUPPORTED; }
  if (p.L <= 208) return launch_bwd2<112, 96, 2>(qkv, p, stream);
  return launch_bwd2<128, 128, 2>(qkv, p, stream);
}

}  // namespace clipk
