// Fused contrastive-logit strip kernels: logits = scale * Q K^T, row softmax cross-entropy against the diagonal
// (label = label_offset + row), streamed over the gallery with an online log-sum-exp -- the [B_local, B_global]
// matrix is only written when the caller asks for it (CLIPApp.forward returns logits_per_text).
//   reference: appzoo/clip/model.py:148-149 (logits), :154-160 (clip_loss = (CE(S) + CE(S^T)) / 2)
// fp32 CUDA-core math on purpose: the logits/loss parity budget (rtol 1e-3) does not survive bf16 embeddings,
// and the op is 2*Bl*Bg*E flops (0.07 GF at B=256; 4.3 GF for a 512 x 4096 strip).
// A CTA owns 32 rows (8 warps x 4 rows, rows cached in registers); the other operand is streamed through shared
// memory in 32-row tiles.  A lane owns the float4 column groups {lane + 32 t}.
#include "common.cuh"
#include "../../include/clipk.h"

namespace clipk {

constexpr int CE_ROWS_PER_WARP = 4;
constexpr int CE_WARPS = 8;
constexpr int CE_ROWS = CE_ROWS_PER_WARP * CE_WARPS;  // 32
constexpr int CE_TILE = 32;

__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

template <int NV>
__global__ void __launch_bounds__(256) ce_strip_fwd_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ logit_scale_log,
                                                           int label_offset, float* __restrict__ S_out, long long lds, int transpose_out,
                                                           float* __restrict__ lse_out, float* __restrict__ loss_rows, int nq, int nk, int E) {
  extern __shared__ float ktile[];  // [CE_TILE][E]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float scale = __expf(*logit_scale_log);
  const int row0 = (blockIdx.x * (blockDim.x >> 5) + warp) * CE_ROWS_PER_WARP;
  float4 q[CE_ROWS_PER_WARP][NV];
#pragma unroll
  for (int r = 0; r < CE_ROWS_PER_WARP; ++r)
#pragma unroll
    for (int t = 0; t < NV; ++t)
      q[r][t] = (row0 + r < nq) ? *reinterpret_cast<const float4*>(Q + (long long)(row0 + r) * E + (lane + 32 * t) * 4) : make_float4(0, 0, 0, 0);
  float mx[CE_ROWS_PER_WARP], sm[CE_ROWS_PER_WARP], lab[CE_ROWS_PER_WARP];
#pragma unroll
  for (int r = 0; r < CE_ROWS_PER_WARP; ++r) { mx[r] = -INFINITY; sm[r] = 0.f; lab[r] = 0.f; }

  for (int k0 = 0; k0 < nk; k0 += CE_TILE) {
    __syncthreads();
    const int nrows = min(CE_TILE, nk - k0);
    for (int idx = threadIdx.x; idx < nrows * (E / 4); idx += blockDim.x)
      reinterpret_cast<float4*>(ktile)[idx] = reinterpret_cast<const float4*>(K + (long long)k0 * E)[idx];
    __syncthreads();
    for (int j = 0; j < nrows; ++j) {
      float p[CE_ROWS_PER_WARP];
#pragma unroll
      for (int r = 0; r < CE_ROWS_PER_WARP; ++r) p[r] = 0.f;
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        const float4 kv = *reinterpret_cast<const float4*>(ktile + j * E + (lane + 32 * t) * 4);
#pragma unroll
        for (int r = 0; r < CE_ROWS_PER_WARP; ++r) p[r] += dot4(q[r][t], kv);
      }
#pragma unroll
      for (int r = 0; r < CE_ROWS_PER_WARP; ++r) {
        const float s = warp_sum(p[r]) * scale;
        const int col = k0 + j;
        if (col == label_offset + row0 + r) lab[r] = s;
        if (s > mx[r]) { sm[r] = sm[r] * __expf(mx[r] - s) + 1.f; mx[r] = s; }
        else sm[r] += __expf(s - mx[r]);
        if (S_out && lane == 0 && row0 + r < nq) {
          if (transpose_out) S_out[(long long)col * lds + row0 + r] = s;
          else S_out[(long long)(row0 + r) * lds + col] = s;
        }
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < CE_ROWS_PER_WARP; ++r)
      if (row0 + r < nq) {
        const float lse = mx[r] + __logf(sm[r]);
        lse_out[row0 + r] = lse;
        loss_rows[row0 + r] = lse - lab[r];
      }
  }
}

// Gradient of  coef * sum_i CE_i  w.r.t. the OWNED operand rows.
//   own_is_query = 1: own = Q rows i (lse/label indexed by own row),   out_i += scale * sum_j dS_ij K_j
//   own_is_query = 0: own = K rows j (lse/label indexed by streamed i), out_j += scale * sum_i dS_ij Q_i
//   dS_ij = coef * (exp(s_ij - lse_i) - [j == label_offset + i]);  dscale_log += sum dS_ij * s_ij (query mode only)
template <int NV>
__global__ void __launch_bounds__(256) ce_strip_bwd_kernel(const float* __restrict__ OWN, const float* __restrict__ STR,
                                                           const float* __restrict__ logit_scale_log, const float* __restrict__ lse,
                                                           int label_offset, float coef, int own_is_query, float* __restrict__ out,
                                                           int accumulate, float* __restrict__ dscale_log, int n_own, int n_str, int E) {
  extern __shared__ float stile[];  // [CE_TILE][E]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float scale = __expf(*logit_scale_log);
  const int row0 = (blockIdx.x * (blockDim.x >> 5) + warp) * CE_ROWS_PER_WARP;
  float4 o[CE_ROWS_PER_WARP][NV], acc[CE_ROWS_PER_WARP][NV];
  float own_lse[CE_ROWS_PER_WARP];
#pragma unroll
  for (int r = 0; r < CE_ROWS_PER_WARP; ++r) {
    own_lse[r] = (own_is_query && row0 + r < n_own) ? lse[row0 + r] : 0.f;
#pragma unroll
    for (int t = 0; t < NV; ++t) {
      o[r][t] = (row0 + r < n_own) ? *reinterpret_cast<const float4*>(OWN + (long long)(row0 + r) * E + (lane + 32 * t) * 4) : make_float4(0, 0, 0, 0);
      acc[r][t] = make_float4(0, 0, 0, 0);
    }
  }
  float dsc = 0.f;
  for (int k0 = 0; k0 < n_str; k0 += CE_TILE) {
    __syncthreads();
    const int nrows = min(CE_TILE, n_str - k0);
    for (int idx = threadIdx.x; idx < nrows * (E / 4); idx += blockDim.x)
      reinterpret_cast<float4*>(stile)[idx] = reinterpret_cast<const float4*>(STR + (long long)k0 * E)[idx];
    __syncthreads();
    for (int j = 0; j < nrows; ++j) {
      float p[CE_ROWS_PER_WARP];
      float4 sv[NV];
#pragma unroll
      for (int r = 0; r < CE_ROWS_PER_WARP; ++r) p[r] = 0.f;
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        sv[t] = *reinterpret_cast<const float4*>(stile + j * E + (lane + 32 * t) * 4);
#pragma unroll
        for (int r = 0; r < CE_ROWS_PER_WARP; ++r) p[r] += dot4(o[r][t], sv[t]);
      }
      const int sidx = k0 + j;
      const float str_lse = own_is_query ? 0.f : lse[sidx];
#pragma unroll
      for (int r = 0; r < CE_ROWS_PER_WARP; ++r) {
        const float s = warp_sum(p[r]) * scale;
        const int own_idx = row0 + r;
        const int qi = own_is_query ? own_idx : sidx;
        const int kj = own_is_query ? sidx : own_idx;
        const float l = own_is_query ? own_lse[r] : str_lse;
        float ds = coef * (__expf(s - l) - ((kj == label_offset + qi) ? 1.f : 0.f));
        if (own_idx >= n_own) ds = 0.f;
        dsc += ds * s;
        const float w = ds * scale;
#pragma unroll
        for (int t = 0; t < NV; ++t) {
          acc[r][t].x += w * sv[t].x; acc[r][t].y += w * sv[t].y; acc[r][t].z += w * sv[t].z; acc[r][t].w += w * sv[t].w;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < CE_ROWS_PER_WARP; ++r)
    if (row0 + r < n_own)
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        float4* dst = reinterpret_cast<float4*>(out + (long long)(row0 + r) * E + (lane + 32 * t) * 4);
        float4 v = acc[r][t];
        if (accumulate) { float4 c = *dst; v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w; }
        *dst = v;
      }
  if (own_is_query && dscale_log && lane == 0) atomicAdd(dscale_log, dsc);  // dsc identical on all lanes
}

// rank_out[i] = #{ j : s_ij > s_{i,label(i)} }  -- text->image retrieval rank of the matching gallery item
// (appzoo/clip/evaluator.py:47-61: hit@K <=> the match is among the first K of torch.sort(descending) <=> rank < K,
// exact unless two gallery scores tie bit-for-bit).  Both the diagonal score and the streamed scores go through the
// same summation order, so the comparison is self-consistent.
template <int NV>
__global__ void __launch_bounds__(256) retrieval_rank_kernel(const float* __restrict__ Q, const float* __restrict__ K, int label_offset,
                                                             int* __restrict__ rank_out, int nq, int nk, int E) {
  extern __shared__ float ktile[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row0 = (blockIdx.x * (blockDim.x >> 5) + warp) * CE_ROWS_PER_WARP;
  float4 q[CE_ROWS_PER_WARP][NV];
  float diag[CE_ROWS_PER_WARP];
  int cnt[CE_ROWS_PER_WARP];
#pragma unroll
  for (int r = 0; r < CE_ROWS_PER_WARP; ++r) {
    cnt[r] = 0;
    float p = 0.f;
    const int lab = label_offset + row0 + r;
#pragma unroll
    for (int t = 0; t < NV; ++t) {
      q[r][t] = (row0 + r < nq) ? *reinterpret_cast<const float4*>(Q + (long long)(row0 + r) * E + (lane + 32 * t) * 4) : make_float4(0, 0, 0, 0);
      const float4 kv = (row0 + r < nq && lab < nk) ? *reinterpret_cast<const float4*>(K + (long long)lab * E + (lane + 32 * t) * 4) : make_float4(0, 0, 0, 0);
      p += dot4(q[r][t], kv);
    }
    diag[r] = warp_sum(p);
  }
  for (int k0 = 0; k0 < nk; k0 += CE_TILE) {
    __syncthreads();
    const int nrows = min(CE_TILE, nk - k0);
    for (int idx = threadIdx.x; idx < nrows * (E / 4); idx += blockDim.x)
      reinterpret_cast<float4*>(ktile)[idx] = reinterpret_cast<const float4*>(K + (long long)k0 * E)[idx];
    __syncthreads();
    for (int j = 0; j < nrows; ++j) {
      float p[CE_ROWS_PER_WARP];
#pragma unroll
      for (int r = 0; r < CE_ROWS_PER_WARP; ++r) p[r] = 0.f;
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        const float4 kv = *reinterpret_cast<const float4*>(ktile + j * E + (lane + 32 * t) * 4);
#pragma unroll
        for (int r = 0; r < CE_ROWS_PER_WARP; ++r) p[r] += dot4(q[r][t], kv);
      }
#pragma unroll
      for (int r = 0; r < CE_ROWS_PER_WARP; ++r) cnt[r] += (warp_sum(p[r]) > diag[r]) ? 1 : 0;
    }
  }
  if (lane == 0)
#pragma unroll
    for (int r = 0; r < CE_ROWS_PER_WARP; ++r)
      if (row0 + r < nq) rank_out[row0 + r] = cnt[r];
}

// out[0] (+)= scale * sum(x[0:n])   single CTA, deterministic order
__global__ void __launch_bounds__(256) reduce_sum_kernel(const float* __restrict__ x, int n, float scale, float* __restrict__ out, int accumulate) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[0] = (accumulate ? out[0] : 0.f) + t * scale;
  }
}

// warps (x 4 rows) per CTA.  Measured on B200 at B = 256: 8 warps / CTA (8 CTAs) 0.39 ms fwd, 1 warp / CTA (64 CTAs) 0.83 ms -- the
// gallery tile copy into shared memory needs the full CTA's load bandwidth, so small strips keep the big CTA.
static inline int ce_warps_for(int n_rows) {
  (void)n_rows;
  return 8;
}


// ====================================================================================================== tensor-core path
// The strip kernels above stream the gallery through CUDA cores with 32 query rows per CTA: at B_local = 256 that is 8 CTAs on 148
// SMs, and their time grows with the GLOBAL batch.  The training step therefore takes the logits from the tcgen05 GEMM instead,
// without giving up fp32-level logits: each fp32 embedding is split x = hi + lo (two bf16) and
//     <q, k>  ~=  <q_hi, k_hi> + <q_hi, k_lo> + <q_lo, k_hi>          (error ~2^-17 relative, the lo*lo term)
// is ONE GEMM with K = 3E over the concatenations  Qs = [q_hi | q_hi | q_lo],  Ks = [k_hi | k_lo | k_hi].  The row kernels below
// turn the raw dots into scaled logits + log-sum-exp + per-row loss, and into the bf16 dS operand of the two gradient GEMMs
// (dOwn = dS Ks_hi, dGallery = dS^T Qs_hi; bf16 like every other backward operand).
__global__ void __launch_bounds__(256) split_bf16x3_kernel(const float* __restrict__ x, bf16* __restrict__ out, int rows, int cols, int pattern,
                                                           long long ld_out) {
  const long long n4 = (long long)rows * (cols / 4);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(t / (cols / 4)), c = (int)(t % (cols / 4)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + (long long)r * cols + c);
    const float f[4] = {v.x, v.y, v.z, v.w};
    float hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { hi[j] = __bfloat162float(__float2bfloat16_rn(f[j])); lo[j] = f[j] - hi[j]; }
    uint2 h, l;
    h.x = pack_bf16x2(hi[0], hi[1]); h.y = pack_bf16x2(hi[2], hi[3]);
    l.x = pack_bf16x2(lo[0], lo[1]); l.y = pack_bf16x2(lo[2], lo[3]);
    bf16* o = out + (long long)r * ld_out + c;
    *reinterpret_cast<uint2*>(o) = h;                                   // pattern 0: [hi | hi | lo]   pattern 1: [hi | lo | hi]
    *reinterpret_cast<uint2*>(o + cols) = pattern ? l : h;
    *reinterpret_cast<uint2*>(o + 2 * cols) = pattern ? h : l;
  }
}

// one warp per row: S (raw dots) -> scaled logits in place, lse, loss_rows = lse - logit[label]
__global__ void __launch_bounds__(256) ce_rows_fwd_kernel(float* __restrict__ S, long long lds, const float* __restrict__ logit_scale_log,
                                                          int label_offset, float* __restrict__ lse_out, float* __restrict__ loss_rows, int nq,
                                                          int nk) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= nq) return;
  const float scale = __expf(*logit_scale_log);
  float* s = S + (long long)row * lds;
  float mx = -INFINITY;
  for (int j = lane; j < nk; j += 32) { const float v = s[j] * scale; s[j] = v; mx = fmaxf(mx, v); }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < nk; j += 32) sum += __expf(s[j] - mx);     // own writes: same lane, same addresses
  sum = warp_sum(sum);
  if (lane == 0) {
    const float lse = mx + __logf(sum);
    const int lab = label_offset + row;
    lse_out[row] = lse;
    loss_rows[row] = lse - ((lab >= 0 && lab < nk) ? s[lab] : 0.f);
  }
}

// one warp per row: dS' = scale * coef * (exp(s - lse) - [j == label]) as bf16 (columns [nk, ldds) zero-filled);
// dscale_log += sum_j coef * (p - onehot) * s
__global__ void __launch_bounds__(256) ce_rows_bwd_kernel(const float* __restrict__ S, long long lds, const float* __restrict__ logit_scale_log,
                                                          const float* __restrict__ lse, int label_offset, float coef, bf16* __restrict__ dS,
                                                          long long ldds, float* __restrict__ dscale_log, int nq, int nk) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= nq) return;
  const float scale = __expf(*logit_scale_log);
  const float* s = S + (long long)row * lds;
  bf16* o = dS + (long long)row * ldds;
  const float l = lse[row];
  const int lab = label_offset + row;
  float acc = 0.f;
  for (int j = lane; j < (int)ldds; j += 32) {
    float g = 0.f;
    if (j < nk) {
      const float v = s[j];
      g = coef * (__expf(v - l) - (j == lab ? 1.f : 0.f));
      acc += g * v;
    }
    o[j] = __float2bfloat16_rn(g * scale);
  }
  if (dscale_log) {
    acc = warp_sum(acc);
    if (lane == 0) atomicAdd(dscale_log, acc);
  }
}


// thr[i] = <Q_i, K_{label_offset + i}> (fp32): the score of query i's own match, the threshold of the rank-count GEMM epilogue
__global__ void __launch_bounds__(256) match_score_kernel(const float* __restrict__ Q, const float* __restrict__ K, int label_offset,
                                                          float* __restrict__ thr, int nq, int nk, int E) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= nq) return;
  const int lab = label_offset + row;
  float acc = 0.f;
  if (lab >= 0 && lab < nk)
    for (int c = lane * 4; c < E; c += 128) acc += dot4(*reinterpret_cast<const float4*>(Q + (long long)row * E + c), *reinterpret_cast<const float4*>(K + (long long)lab * E + c));
  acc = warp_sum(acc);
  if (lane == 0) thr[row] = (lab >= 0 && lab < nk) ? acc : INFINITY;     // no match in the gallery: nothing ranks above it
}

}  // namespace clipk

using namespace clipk;

#define CE_DISPATCH(NV, ...)                                        \
  switch (NV) {                                                     \
    case 1: __VA_ARGS__(1); break; case 2: __VA_ARGS__(2); break;   \
    case 4: __VA_ARGS__(4); break; case 6: __VA_ARGS__(6); break;   \
    case 8: __VA_ARGS__(8); break;                                  \
    default: set_error("clip_ce: embed dim %d unsupported (128,256,512,768,1024)", E); return CLIPK_ERR_UNSUPPORTED; }

extern "C" int clipk_ce_strip_fwd(const float* Q, const float* K, const float* logit_scale_log, int label_offset, float* S_out,
                                  long long lds, int transpose_out, float* lse, float* loss_rows, int nq, int nk, int E,
                                  cudaStream_t stream) {
  if (nq <= 0 || nk <= 0) return 0;
  if (E % 128) { set_error("clip_ce: E %% 128 != 0"); return CLIPK_ERR_UNSUPPORTED; }
  const int nv = E / 128;
  const int smem = CE_TILE * E * 4;
  const int nw = ce_warps_for(nq);
  dim3 grid((nq + nw * CE_ROWS_PER_WARP - 1) / (nw * CE_ROWS_PER_WARP));
#define LAUNCH(NV)                                                                                                         \
  {                                                                                                                        \
    static bool cfg = false;                                                                                               \
    if (!cfg) { CLIPK_CUDA(cudaFuncSetAttribute(ce_strip_fwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072)); cfg = true; } \
    ce_strip_fwd_kernel<NV><<<grid, 32 * nw, smem, stream>>>(Q, K, logit_scale_log, label_offset, S_out, lds, transpose_out, lse, loss_rows, nq, nk, E); \
  }
  CE_DISPATCH(nv, LAUNCH)
#undef LAUNCH
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_ce_strip_bwd(const float* own, const float* streamed, const float* logit_scale_log, const float* lse, int label_offset,
                                  float coef, int own_is_query, float* out, int accumulate, float* dscale_log, int n_own, int n_streamed,
                                  int E, cudaStream_t stream) {
  if (n_own <= 0 || n_streamed <= 0) return 0;
  if (E % 128) { set_error("clip_ce: E %% 128 != 0"); return CLIPK_ERR_UNSUPPORTED; }
  const int nv = E / 128;
  const int smem = CE_TILE * E * 4;
  const int nw = ce_warps_for(n_own);
  dim3 grid((n_own + nw * CE_ROWS_PER_WARP - 1) / (nw * CE_ROWS_PER_WARP));
#define LAUNCH(NV)                                                                                                         \
  {                                                                                                                        \
    static bool cfg = false;                                                                                               \
    if (!cfg) { CLIPK_CUDA(cudaFuncSetAttribute(ce_strip_bwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072)); cfg = true; } \
    ce_strip_bwd_kernel<NV><<<grid, 32 * nw, smem, stream>>>(own, streamed, logit_scale_log, lse, label_offset, coef, own_is_query, out, accumulate, dscale_log, n_own, n_streamed, E); \
  }
  CE_DISPATCH(nv, LAUNCH)
#undef LAUNCH
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_reduce_sum(const float* x, int n, float scale, float* out, int accumulate, cudaStream_t stream) {
  reduce_sum_kernel<<<1, 256, 0, stream>>>(x, n, scale, out, accumulate);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_retrieval_rank(const float* Q, const float* K, int label_offset, int* rank_out, int nq, int nk, int E, cudaStream_t stream) {
  if (nq <= 0 || nk <= 0) return 0;
  if (E % 128) { set_error("retrieval_rank: E %% 128 != 0"); return CLIPK_ERR_UNSUPPORTED; }
  const int nv = E / 128;
  const int smem = CE_TILE * E * 4;
  const int nw = ce_warps_for(nq);
  dim3 grid((nq + nw * CE_ROWS_PER_WARP - 1) / (nw * CE_ROWS_PER_WARP));
#define LAUNCH(NV)                                                                                                         \
  {                                                                                                                        \
    static bool cfg = false;                                                                                               \
    if (!cfg) { CLIPK_CUDA(cudaFuncSetAttribute(retrieval_rank_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072)); cfg = true; } \
    retrieval_rank_kernel<NV><<<grid, 32 * nw, smem, stream>>>(Q, K, label_offset, rank_out, nq, nk, E);                       \
  }
  CE_DISPATCH(nv, LAUNCH)
#undef LAUNCH
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_split_bf16x3(const float* x, void* out_bf16, int rows, int cols, int pattern, long long ld_out, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return 0;
  if ((cols % 4) || (ld_out % 4) || ld_out < 3LL * cols) { set_error("split_bf16x3: cols=%d ld_out=%lld unsupported", cols, ld_out); return CLIPK_ERR_ARG; }
  const long long n4 = (long long)rows * (cols / 4);
  long long g = (n4 + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  split_bf16x3_kernel<<<(int)g, 256, 0, stream>>>(x, (bf16*)out_bf16, rows, cols, pattern ? 1 : 0, ld_out);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_ce_rows_fwd(float* S, long long lds, const float* logit_scale_log, int label_offset, float* lse, float* loss_rows, int nq,
                                 int nk, cudaStream_t stream) {
  if (nq <= 0 || nk <= 0) return 0;
  ce_rows_fwd_kernel<<<(nq + 7) / 8, 256, 0, stream>>>(S, lds, logit_scale_log, label_offset, lse, loss_rows, nq, nk);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_ce_rows_bwd(const float* S, long long lds, const float* logit_scale_log, const float* lse, int label_offset, float coef,
                                 void* dS_bf16, long long ldds, float* dscale_log, int nq, int nk, cudaStream_t stream) {
  if (nq <= 0 || nk <= 0) return 0;
  if (ldds < nk) { set_error("ce_rows_bwd: ldds=%lld < nk=%d", ldds, nk); return CLIPK_ERR_ARG; }
  ce_rows_bwd_kernel<<<(nq + 7) / 8, 256, 0, stream>>>(S, lds, logit_scale_log, lse, label_offset, coef, (bf16*)dS_bf16, ldds, dscale_log, nq, nk);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

static size_t rank_tc_align(size_t x) { return (x + 255) / 256 * 256; }

extern "C" size_t clipk_retrieval_rank_tc_workspace(int nq, int nk, int E) {
  return rank_tc_align((size_t)nq * 3 * E * 2) + rank_tc_align((size_t)nk * 3 * E * 2) + rank_tc_align((size_t)nq * 4);
}

extern "C" int clipk_retrieval_rank_tc(const float* Q, const float* K, int label_offset, int* rank_out, int nq, int nk, int E, void* workspace,
                                       size_t workspace_bytes, cudaStream_t stream) {
  if (nq <= 0 || nk <= 0) return 0;
  if (E % 8) { set_error("retrieval_rank_tc: E %% 8 != 0"); return CLIPK_ERR_UNSUPPORTED; }
  if (!workspace || workspace_bytes < clipk_retrieval_rank_tc_workspace(nq, nk, E) || (reinterpret_cast<uintptr_t>(workspace) & 15)) {
    set_error("retrieval_rank_tc: workspace of %zu bytes (16-byte aligned) required", clipk_retrieval_rank_tc_workspace(nq, nk, E));
    return CLIPK_ERR_ARG;
  }
  char* w = reinterpret_cast<char*>(workspace);
  void* Qs = w; w += rank_tc_align((size_t)nq * 3 * E * 2);
  void* Ks = w; w += rank_tc_align((size_t)nk * 3 * E * 2);
  float* thr = reinterpret_cast<float*>(w);
  int rc;
  if ((rc = clipk_split_bf16x3(Q, Qs, nq, E, 0, 3LL * E, stream))) return rc;
  if ((rc = clipk_split_bf16x3(K, Ks, nk, E, 1, 3LL * E, stream))) return rc;
  match_score_kernel<<<(nq + 7) / 8, 256, 0, stream>>>(Q, K, label_offset, thr, nq, nk, E);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  CLIPK_CUDA(cudaMemsetAsync(rank_out, 0, (size_t)nq * sizeof(int), stream));
  clipk_epilogue_t e{};
  e.mode = CLIPK_EPI_RANK_COUNT; e.out_dtype = CLIPK_F32; e.out = rank_out; e.aux = thr; e.alpha = 1.0f; e.label_offset = label_offset;
  return clipk_gemm_bf16(Qs, 3 * E, 0, Ks, 3 * E, 0, nq, nk, 3 * E, &e, 1, stream);
}
