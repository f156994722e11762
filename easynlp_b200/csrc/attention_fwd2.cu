// Attention forward, second generation: ONE persistent warp-specialised CTA per SM that pipelines (sample, head) work items.
//
//   warp 0      TMA producer: K, V and the query tiles of item j+1 stream into the second smem stage while item j is computed
//   warp 1      MMA issuer : S = Q K^T  (tcgen05, A/B from smem)  and  O = P V  (A = P read straight from TENSOR MEMORY)
//   warps 4-19  softmax    : 16 warps = 4 TMEM lane quarters x 4 column quarters.  A thread owns one query row and a quarter of the key
//               range; its scores are read from TMEM ONCE and stay in registers between the max and the exp pass; the bf16
//               probabilities go back into the SAME TMEM columns with tcgen05.st (no shared-memory round trip, no proxy fence)
//   warps 20-23 epilogue   : O / rowsum -> ctx (bf16), log-sum-exp; runs concurrently with the softmax of the next tile
//
// TMEM (512 columns): two tile buffers (S [0, keys) -> P packed bf16 [0, keys/2)) and, when they leave 64 columns free (keys <= 224),
// ONE separate O accumulator at [448, 512): the next-but-one S = Q K^T is then queued right behind O = P V in the MMA thread's program
// order (tcgen05.mma of one thread execute in order) and the epilogue that drains O is off the critical path.  keys = 256: O lives
// inside its tile buffer at [128, 192) and S waits for the epilogue.
// While the softmax warps work on tile i (exp on the MUFU pipe), the tensor pipe runs O(i-1) = P V and S(i+1) = Q K^T, the TMA
// unit fetches the next item and the epilogue warps drain O(i-1): the serial chain of the first-generation kernel
// (attention.cu: load -> S -> max -> exp -> smem P -> PV -> store, 2 CTAs / SM as the only overlap) becomes a pipeline.
//
// Replaces nn.MultiheadAttention's SDPA (modeling_chineseclip.py:188,198-200) and BertSelfAttention's QK^T / +mask / softmax /
// dropout / PV chain (modeling_bert.py:210-244; additive key mask (1-m)*-10000, modeling_utils.py:438-439).
#include "common.cuh"
#include "attention_common.cuh"
#include "../../include/clipk.h"

namespace clipk {

constexpr int F2_THREADS = 768;          // 24 warps: WG0 = {TMA, MMA, -, -}, WG1-4 = softmax, WG5 = epilogue
constexpr int F2_SOFTMAX_WARP0 = 4;
constexpr int F2_EPI_WARP0 = 20;

// timeline probe (diagnostics): event e of tile i of CTA 0
#define F2_DBG(i, e) do { if (p.dbg && blockIdx.x == 0 && (i) < 64) p.dbg[(i) * 16 + (e)] = clock64(); } while (0)

struct Fwd2Smem {            // byte offsets inside the dynamic shared memory (base 1024-B aligned)
  int kv_bytes;              // one of K / V: keys_pad * 128
  int stage_bytes;           // K + V + q_tiles * 16 KB
  int mask_off, red_off, bar_off, total;
};
constexpr int F2_MASK = 288;     // floats of key mask per stage (keys <= 272)
// one item stage = K + V + all query tiles; two stages (the next item is prefetched) when they fit, else one
__host__ __device__ inline int fwd2_stages(int n16, int q_tiles) { return 2 * (2 * n16 * 2048 + q_tiles * 16384) + 16384 <= 232448 - 1024 ? 2 : 1; }
__host__ __device__ inline Fwd2Smem fwd2_layout(int n16, int q_tiles) {
  Fwd2Smem s;
  s.kv_bytes = n16 * 16 * 128;
  s.stage_bytes = 2 * s.kv_bytes + q_tiles * 16384;
  s.mask_off = fwd2_stages(n16, q_tiles) * s.stage_bytes;   // [2 stages][F2_MASK] floats
  s.red_off = s.mask_off + 2 * F2_MASK * 4;       // max [2 buf][4][128], sum [3][4][128], rowmax [3][128]
  s.bar_off = s.red_off + (2 * 4 * 128 + 3 * 4 * 128 + 3 * 128) * 4;
  s.total = s.bar_off + 256;
  return s;
}

// N16 = padded key count / 16 (compile time: the per-thread score slice of 4 * N16 columns lives in registers)
template <int N16>
__global__ void __launch_bounds__(F2_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnParams p) {
  constexpr int KEYS = N16 * 16;          // key columns of S (keys beyond L are masked to -inf)
  constexpr int CPT = N16 * 4;            // columns per softmax thread
  constexpr int BUF1 = (KEYS + 31) / 32 * 32;             // column of the second tile buffer
  constexpr int NBUF = KEYS <= 256 ? 2 : 1;               // keys > 256 (ViT-L/14: 257 tokens -> 272): one tile buffer of 272 columns
  constexpr bool O_SEP = NBUF == 1 || BUF1 + KEYS <= 448; // room for a separate O accumulator at [448, 512)
  constexpr int BUF_STRIDE = NBUF == 1 ? 0 : (O_SEP ? BUF1 : 256);
  constexpr int KBOX = KEYS <= 256 ? 1 : 2;               // TMA boxes per K / V tile (a box holds <= 256 rows)
  const int NST = fwd2_stages(N16, p.q_tiles);
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const Fwd2Smem lay = fwd2_layout(N16, p.q_tiles);
  float* smask = reinterpret_cast<float*>(smem + lay.mask_off);
  float* sred_max = reinterpret_cast<float*>(smem + lay.red_off);         // [buf][cq][row]
  // row sums / row maxima travel from the softmax warps of tile i to its epilogue in slot i % 3: the slot is rewritten by tile i+3, whose
  // S is issued after PV(i+1), which waited for the epilogue of tile i (o_free / buf_free)
  float* sred_sum = sred_max + 2 * 4 * 128;                               // [i % 3][cq][row]
  float* srow_max = sred_sum + 3 * 4 * 128;                               // [i % 3][row]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
  uint64_t* kq_full = bars;            // [2] K + Q tiles + mask of a stage have landed
  uint64_t* v_full = bars + 2;         // [2]
  uint64_t* stage_empty = bars + 4;    // [2] every MMA reading the stage has completed
  uint64_t* s_full = bars + 6;         // [2] S of a tile buffer is complete
  uint64_t* p_full = bars + 8;         // [2] P has been written to TMEM by all 16 softmax warps
  uint64_t* o_full = bars + 10;        // [2] O = P V is complete
  uint64_t* buf_free = bars + 12;      // [2] (!O_SEP) the epilogue has drained O: the tile buffer can take the next S
  uint64_t* o_free = bars + 14;        // (O_SEP) the epilogue has drained the single O accumulator
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 15);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_items = p.B * p.H;
  const int T = p.q_tiles;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kq_full[s], 2); mbar_init(&v_full[s], 1); mbar_init(&stage_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 16); mbar_init(&o_full[s], 1); mbar_init(&buf_free[s], 4);
    }
    mbar_init(o_free, 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_holder;

  // 768 threads start with 80 registers each; the two single-thread roles and the epilogue hand registers to the 512 softmax threads,
  // whose score slices (up to 64 values) then stay in registers: the CTA owns 768 x 80 = 61440 registers: 128 x 40 + 128 x 56 + 512 x 96
  // (each setmaxnreg sits at the top of its role's branch: ptxas budgets registers per control-flow region)
  if (warp < 4) {
  reg_dec<40>();
  if (warp == 0) {
    // ============================================================================================ TMA producer (+ key mask)
    int j = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++j) {
      const int b = item / p.H, h = item - b * p.H;
      const int stage = j % NST;
      mbar_wait(&stage_empty[stage], ((j / NST) & 1) ^ 1);
      uint8_t* sK = smem + stage * lay.stage_bytes;
      uint8_t* sV = sK + lay.kv_bytes;
      uint8_t* sQ = sV + lay.kv_bytes;
      if (lane == 0) {
        mbar_expect_tx(&kq_full[stage], lay.kv_bytes + T * 16384);
#pragma unroll
        for (int bx = 0; bx < KBOX; ++bx) tma_load_2d(sK + bx * (lay.kv_bytes / KBOX), &tmKV, &kq_full[stage], p.d + h * 64, b * p.L + bx * (KEYS / KBOX));
        for (int t = 0; t < T; ++t) tma_load_2d(sQ + t * 16384, &tmQ, &kq_full[stage], h * 64, b * p.L + t * 128);
        mbar_expect_tx(&v_full[stage], lay.kv_bytes);
#pragma unroll
        for (int bx = 0; bx < KBOX; ++bx) tma_load_2d(sV + bx * (lay.kv_bytes / KBOX), &tmKV, &v_full[stage], 2 * p.d + h * 64, b * p.L + bx * (KEYS / KBOX));
      }
      // additive key mask of this sample in log2 units; keys beyond L (padding rows of the box = the next sample's tokens) -> -inf
      float* m = smask + stage * F2_MASK;
      for (int c = lane; c < KEYS; c += 32) m[c] = (c < p.L) ? (p.mask ? p.mask[(long long)b * p.L + c] * LOG2E : 0.f) : -INFINITY;
      __syncwarp();
      if (lane == 0) mbar_arrive(&kq_full[stage]);
    }
  } else if (warp == 1) {
    // ============================================================================================ MMA issuer (one thread)
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, KEYS <= 256 ? KEYS : 256, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
      const int n_tiles = ((n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * T;
      auto issue_s = [&](int k) {           // S(k) = Q_t K^T of tile k -> tile buffer k & 1
        const int j = k / T, t = k - j * T;
        const int stage = j % NST, buf = k % NBUF;
        if (t == 0) mbar_wait(&kq_full[stage], (j / NST) & 1);
        F2_DBG(k, 0);
        if (!O_SEP) mbar_wait(&buf_free[buf], ((k / NBUF) & 1) ^ 1);
        tc_fence_after();
        F2_DBG(k, 1);
        const uint32_t aK = smem_u32(smem + stage * lay.stage_bytes);
        const uint32_t aQ = aK + 2 * lay.kv_bytes + t * 16384;
        const uint32_t dQ_ = umma_desc_lo(aQ, 16), dK_ = umma_desc_lo(aK, 16);       // K-major: k-step of 16 head dims = 32 B -> +2
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if constexpr (KEYS <= 256) {
            umma_bf16_lh(tmem + buf * BUF_STRIDE, dQ_ + 2 * kk, dK_ + 2 * kk, idesc_s, kk > 0);
          } else {      // N = 256 + (KEYS - 256): an MMA instruction takes at most 256 columns
            umma_bf16_lh(tmem, dQ_ + 2 * kk, dK_ + 2 * kk, umma_idesc_bf16(128, 256, 0, 0), kk > 0);
            umma_bf16_lh(tmem + 256, dQ_ + 2 * kk, dK_ + (256 * 128 >> 4) + 2 * kk, umma_idesc_bf16(128, KEYS - 256, 0, 0), kk > 0);
          }
        }
        umma_commit(&s_full[buf]);
      };
      auto issue_pv = [&](int k) {          // O(k) = P(k) V
        const int j = k / T, t = k - j * T;
        const int stage = j % NST, buf = k % NBUF;
        mbar_wait(&p_full[buf], (k / NBUF) & 1);
        if (t == 0) mbar_wait(&v_full[stage], (j / NST) & 1);
        if (O_SEP) mbar_wait(o_free, (k & 1) ^ 1);           // the epilogue of tile k-1 has drained O
        tc_fence_after();
        F2_DBG(k, 2);
        const uint32_t aV = smem_u32(smem + stage * lay.stage_bytes + lay.kv_bytes);
        const uint32_t tbuf = tmem + buf * BUF_STRIDE;
        const uint32_t tO = O_SEP ? tmem + 448 : tbuf + 128;
        const uint32_t dV_ = umma_desc_lo(aV, 16384);          // V as MN-major B: k-step of 16 keys = 2048 B -> +128
#pragma unroll
        for (int kk = 0; kk < N16; ++kk) umma_bf16_ts_lh(tO, tbuf + kk * 8, dV_ + 128 * kk, idesc_o, kk > 0);
        umma_commit(O_SEP ? &o_full[0] : &o_full[buf]);
        if (t == T - 1) umma_commit(&stage_empty[stage]);
        F2_DBG(k, 3);
      };
      // program order S(0) S(1) | PV(0) S(2) | PV(1) S(3) | ... : S(k+NBUF) overwrites the buffer whose P was just consumed by PV(k)
      for (int k = 0; k < NBUF && k < n_tiles; ++k) issue_s(k);
      for (int k = 0; k < n_tiles; ++k) {
        issue_pv(k);
        if (k + NBUF < n_tiles) issue_s(k + NBUF);
      }
    }
  }
  } else if (warp < F2_EPI_WARP0) {
    // ============================================================================================ softmax warps
    reg_inc<96>();
    const int q4 = warp & 3;                          // TMEM lane quarter this warp may access
    const int cq = (warp - F2_SOFTMAX_WARP0) >> 2;    // column quarter
    const int row = q4 * 32 + lane;
    const float sc = p.scale * LOG2E;
    const DropCtx dc = drop_ctx(p.drop);
    // plain: no key mask and every column of the slice is a real key.  tail: no key mask and the padding keys all sit in the last 16
    // columns of the slice (the usual case: L rounded up to a multiple of 16) -> they are set to -inf in registers, then as plain.
    const int n_valid = p.L - cq * CPT;               // valid columns of this thread's slice (may be <= 0 or >= CPT)
    const bool tail = (p.mask == nullptr) && !p.causal && n_valid < CPT && n_valid >= CPT - 16 && n_valid > 0;
    const bool plain = (p.mask == nullptr) && !p.causal && (n_valid >= CPT || tail);
    int i = 0, j = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++j) {
      const int stage = j % NST;
      const float* m = smask + stage * F2_MASK + cq * CPT;
      for (int t = 0; t < T; ++t, ++i) {
        const int buf = i % NBUF;
        const uint32_t t_row = tmem + ((uint32_t)(q4 * 32) << 16) + buf * BUF_STRIDE;
        if (warp == F2_SOFTMAX_WARP0 && lane == 0) F2_DBG(i, 4);
        mbar_wait(&s_full[buf], (i / NBUF) & 1);
        tc_fence_after();
        if (warp == F2_SOFTMAX_WARP0 && lane == 0) F2_DBG(i, 5);
        uint32_t s[CPT];
        tmem_ld_n<CPT>(t_row + cq * CPT, s);
        tmem_wait_ld();
        // Row max over this thread's slice.  All of this thread's TMEM reads are complete before the barrier below, after which the
        // threads of the same rows overwrite these columns with probabilities.  `plain` (warp-uniform): no key mask and every column of
        // the slice is a real key -> the max is taken on the raw dots (1 FMNMX3 per 2 scores), the scale is folded into the exponent.
        float mx = -INFINITY, k1, k2;
        if (tail) {
#pragma unroll
          for (int c = (CPT >= 16 ? CPT - 16 : 0); c < CPT; ++c)
            if (c >= n_valid) s[c] = 0xff800000u;       // -inf
        }
        if (plain) {
#pragma unroll
          for (int c = 0; c < CPT; c += 4) {
            mx = fmax3(mx, __uint_as_float(s[c]), __uint_as_float(s[c + 1])); mx = fmax3(mx, __uint_as_float(s[c + 2]), __uint_as_float(s[c + 3]));
          }
          mx *= sc; k1 = sc;
        } else {
#pragma unroll
          for (int c = 0; c < CPT; c += 4) {
            const float4 mm = *reinterpret_cast<const float4*>(m + c);
            const float a0 = fmaf(__uint_as_float(s[c]), sc, mm.x), a1 = fmaf(__uint_as_float(s[c + 1]), sc, mm.y);
            const float a2 = fmaf(__uint_as_float(s[c + 2]), sc, mm.z), a3 = fmaf(__uint_as_float(s[c + 3]), sc, mm.w);
            s[c] = __float_as_uint(a0); s[c + 1] = __float_as_uint(a1); s[c + 2] = __float_as_uint(a2); s[c + 3] = __float_as_uint(a3);
            mx = fmax3(mx, a0, a1); mx = fmax3(mx, a2, a3);
          }
          if (p.causal) {        // keys after the query position are masked out (the diagonal itself is kept: a row is never empty)
            const int qpos = t * 128 + row;
            mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              if (cq * CPT + c > qpos) s[c] = 0xff800000u;
              mx = fmaxf(mx, __uint_as_float(s[c]));
            }
          }
          k1 = 1.0f;
        }
        sred_max[(buf * 4 + cq) * 128 + row] = mx;
        tc_fence_before();
        if (warp == F2_SOFTMAX_WARP0 && lane == 0) F2_DBG(i, 6);
        named_bar_sync(1 + q4, 128);       // the 4 warps that share these 32 rows (they also share an SM sub-partition)
        tc_fence_after();
        if (warp == F2_SOFTMAX_WARP0 && lane == 0) F2_DBG(i, 7);
        {
          const float* r = sred_max + buf * 4 * 128 + row;
          mx = fmaxf(fmaxf(r[0], r[128]), fmaxf(r[256], r[384]));
        }
        k2 = -mx;
        // exp -> bf16 pairs -> TMEM, 16 columns at a time (the packed probabilities of a chunk leave the registers immediately).
        // The bf16 pair (keys 2c, 2c+1) goes to 32-bit column c of the buffer: the layout tcgen05.mma reads an A operand from.
        float sum = 0.f;
        uint32_t drow = 0;
        if (dc.on) {      // dropout on the probabilities (modeling_bert.py:238): the row sum is taken before the mask
          const int b = item / p.H, h = item - b * p.H;
          drow = (uint32_t)((b * p.H + h) * p.L + t * 128 + row);
        }
#pragma unroll
        for (int c0 = 0; c0 < CPT; c0 += 16) {
          constexpr int FULL = 16;
          const int n = (CPT - c0) < FULL ? (CPT - c0) : FULL;       // compile-time after unrolling: 16, or the 4 / 8-column tail
          uint32_t pk[8];
#pragma unroll
          for (int c = 0; c < FULL; c += 4) {
            if (c < n) {
              const float p0 = exp2f(fmaf(__uint_as_float(s[c0 + c]), k1, k2)), p1 = exp2f(fmaf(__uint_as_float(s[c0 + c + 1]), k1, k2));
              const float p2 = exp2f(fmaf(__uint_as_float(s[c0 + c + 2]), k1, k2)), p3 = exp2f(fmaf(__uint_as_float(s[c0 + c + 3]), k1, k2));
              sum += (p0 + p1) + (p2 + p3);
              if (dc.on) {
                const float4 dm = drop_mult4(dc, drow, (uint32_t)((cq * CPT + c0 + c) >> 2));
                pk[c >> 1] = pack_bf16x2(p0 * dm.x, p1 * dm.y); pk[(c >> 1) + 1] = pack_bf16x2(p2 * dm.z, p3 * dm.w);
              } else {
                pk[c >> 1] = pack_bf16x2(p0, p1); pk[(c >> 1) + 1] = pack_bf16x2(p2, p3);
              }
            }
          }
          const uint32_t dstc = t_row + cq * (CPT / 2) + (c0 >> 1);
          if (n == 16) TmemIO<8>::st(dstc, pk);
          else if (n == 8) TmemIO<4>::st(dstc, pk);
          else TmemIO<2>::st(dstc, pk);                               // n == 4
        }
        sred_sum[((i % 3) * 4 + cq) * 128 + row] = sum;
        if (cq == 0) srow_max[(i % 3) * 128 + row] = mx;
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (warp == F2_SOFTMAX_WARP0 && lane == 0) F2_DBG(i, 8);
        if (p.dbg && lane == 0 && blockIdx.x == 0 && i < 64) atomicMax(reinterpret_cast<unsigned long long*>(p.dbg) + i * 16 + 9, (unsigned long long)clock64());
        if (lane == 0) mbar_arrive(&p_full[buf]);
      }
    }
  } else {
    // ============================================================================================ epilogue warps
    reg_dec<56>();
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    int i = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / p.H, h = item - b * p.H;
      for (int t = 0; t < T; ++t, ++i) {
        const int buf = i % NBUF;
        const uint32_t t_row = tmem + ((uint32_t)(q4 * 32) << 16) + (O_SEP ? 448 : buf * BUF_STRIDE + 128);
        const int q = t * 128 + row;
        if (O_SEP) mbar_wait(&o_full[0], i & 1); else mbar_wait(&o_full[buf], (i / NBUF) & 1);
        tc_fence_after();
        if (warp == F2_EPI_WARP0 && lane == 0) F2_DBG(i, 10);
        const float* rs = sred_sum + (i % 3) * 4 * 128 + row;
        const float sum = (rs[0] + rs[128]) + (rs[256] + rs[384]);
        const float mx = srow_max[(i % 3) * 128 + row];
        const float inv = 1.0f / sum;
        bf16* dst = p.ctx + (long long)(b * p.L + q) * p.d + h * 64;
        // The tile buffer is handed back to the MMA warp as soon as the second half of O sits in registers (the next-but-one S = Q K^T
        // waits for it); scaling and the global stores of that half happen afterwards.
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[32];
          tmem_ld_x32(t_row + hh * 32, r);
          tmem_wait_ld();
          if (hh == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(O_SEP ? o_free : &buf_free[buf]);
            if (warp == F2_EPI_WARP0 && lane == 0) F2_DBG(i, 11);
          }
          if (q < p.L) {
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
              uint4 o;
              o.x = pack_bf16x2(__uint_as_float(r[s4 * 8 + 0]) * inv, __uint_as_float(r[s4 * 8 + 1]) * inv);
              o.y = pack_bf16x2(__uint_as_float(r[s4 * 8 + 2]) * inv, __uint_as_float(r[s4 * 8 + 3]) * inv);
              o.z = pack_bf16x2(__uint_as_float(r[s4 * 8 + 4]) * inv, __uint_as_float(r[s4 * 8 + 5]) * inv);
              o.w = pack_bf16x2(__uint_as_float(r[s4 * 8 + 6]) * inv, __uint_as_float(r[s4 * 8 + 7]) * inv);
              *reinterpret_cast<uint4*>(dst + hh * 32 + s4 * 8) = o;
            }
          }
        }
        if (q < p.L && p.lse) p.lse[((long long)b * p.H + h) * p.L + q] = (mx + log2f(sum)) * LN2;
        if (warp == F2_EPI_WARP0 && lane == 0) F2_DBG(i, 12);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N16>
static int launch_fwd2(const CUtensorMap& tQ, const CUtensorMap& tKV, const AttnParams& p, cudaStream_t stream) {
  const Fwd2Smem lay = fwd2_layout(N16, p.q_tiles);
  const int smem = lay.total + 1024;
  static int configured = 0;
  if (configured < smem) {
    CLIPK_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<N16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  const int items = p.B * p.H;
  const int grid = items < sm_count() ? items : sm_count();
  attn_fwd2_kernel<N16><<<grid, F2_THREADS, smem, stream>>>(tQ, tKV, p);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

// supported padded key counts (x16): a sequence is rounded up to the next one, the extra key columns are masked
static int fwd2_round_n16(int n16) {
  const int sizes[] = {2, 4, 5, 8, 13, 16, 17};
  for (int s : sizes) if (n16 <= s) return s;
  return 0;
}

int attention_fwd2(const void* qkv, const AttnParams& p_in, cudaStream_t stream) {
  AttnParams p = p_in;
  const int n16 = fwd2_round_n16((p.L + 15) / 16);
  if (!n16) { set_error("attention_fwd2: L=%d > 272", p.L); return CLIPK_ERR_UNSUPPORTED; }
  p.lk_pad = n16 * 16;
  CUtensorMap tQ, tKV;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tQ, qkv, 3ull * p.d, (uint64_t)p.B * p.L, 3ull * p.d, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tKV, qkv, 3ull * p.d, (uint64_t)p.B * p.L, 3ull * p.d, 64, p.lk_pad <= 256 ? p.lk_pad : p.lk_pad / 2))) return rc;
  switch (n16) {
    case 2: return launch_fwd2<2>(tQ, tKV, p, stream);
    case 4: return launch_fwd2<4>(tQ, tKV, p, stream);
    case 5: return launch_fwd2<5>(tQ, tKV, p, stream);
    case 8: return launch_fwd2<8>(tQ, tKV, p, stream);
    case 13: return launch_fwd2<13>(tQ, tKV, p, stream);
    case 16: return launch_fwd2<16>(tQ, tKV, p, stream);
    default: return launch_fwd2<17>(tQ, tKV, p, stream);
  }
}

}  // namespace clipk
