// HBM-bound kernels of the CLIP towers: LayerNorm fwd/bwd, patch im2col, ViT token assembly, BERT embedding
// gather/scatter, column sums (bias grads), L2 normalisation.  One warp owns one row; a lane owns the float4
// column groups {lane + 32 i}, so per-column reductions over rows (dgamma, dbeta, dbias) stay in registers.
#include <limits.h>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/clipk.h"

namespace clipk {

constexpr int LN_MAXV = 8;  // float4 per lane -> d <= 1024, d % 128 == 0

__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_bf4(bf16* p, float4 v) {
  uint2 o; o.x = pack_bf16x2(v.x, v.y); o.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = o;
}
__device__ __forceinline__ float4 ld_bf4(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.x));
  float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// ------------------------------------------------------------------------------------------------ LayerNorm fwd
// reference: modeling_chineseclip.py:170-176 (fp32 LN, eps 1e-5) and nn.LayerNorm(eps=1e-12) in modeling_bert.py:84,266,344
// optional fused residual add: xs = x + add (add = the bf16 output of the preceding projection GEMM), xs is written to x_out
// (fp32, kept for backward / the next residual) and normalised -- the GEMM then has a plain bf16 epilogue and the fp32 residual
// stream is only touched by this HBM-streaming kernel.
template <int NV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, long long ldx, const bf16* __restrict__ add,
                                                            long long ldadd, float* __restrict__ x_out, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, bf16* __restrict__ y_bf16,
                                                            float* __restrict__ y_f32, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out, int rows, int d, const DropArg da,
                                                            const int drop_mode) {
  // drop_mode 1: dropout on `add` (BertSelfOutput / BertOutput: LN(dropout(dense) + input), modeling_bert.py:266-268,344-346)
  // drop_mode 2: dropout on the normalised output (BertEmbeddings, modeling_bert.py:127-128)
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const DropCtx dc = drop_ctx(da);
  const float* xr = x + (long long)row * ldx;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = ld_f4(xr + (lane + 32 * i) * 4);
  if (add) {
    const bf16* ar = add + (long long)row * ldadd;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 a = ld_bf4(ar + (lane + 32 * i) * 4);
      if (dc.on && drop_mode == 1) { const float4 m = drop_mult4(dc, row, lane + 32 * i); a.x *= m.x; a.y *= m.y; a.z *= m.z; a.w *= m.w; }
      v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w;
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    s += v[i].x + v[i].y + v[i].z + v[i].w;
    if (x_out) st_f4(x_out + (long long)row * d + (lane + 32 * i) * 4, v[i]);
  }
  const float mean = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
    q += a * a + b * b + c * c + e * e;
  }
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 4;
    float4 g = ld_f4(gamma + c), b = ld_f4(beta + c), o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (dc.on && drop_mode == 2) { const float4 m = drop_mult4(dc, row, lane + 32 * i); o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w; }
    if (y_f32) st_f4(y_f32 + (long long)row * d + c, o);
    if (y_bf16) st_bf4(y_bf16 + (long long)row * d + c, o);
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm bwd
//   g   = (dy [+ dy_add]) * gamma
//   dx  = rstd * (g - mean_d(g) - xhat * mean_d(g * xhat))  [+ dx_add]
//   dgamma += sum_rows (dy+dy_add) * xhat ; dbeta += sum_rows (dy+dy_add) ; dbias += sum_rows dx   (fp32 atomics)
template <int NV>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_kernel(const void* __restrict__ dy, int dy_is_f32, const float* __restrict__ dy_add,
                                                               const float* __restrict__ x, long long ldx,
                                                               const float* __restrict__ gamma, const float* __restrict__ mean_in,
                                                               const float* __restrict__ rstd_in, const float* __restrict__ dx_add,
                                                               float* __restrict__ dx_f32, long long lddx, bf16* __restrict__ dx_bf16,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                               float* __restrict__ dbias, int rows, int d, const DropArg da,
                                                               const int drop_mode) {
  // drop_mode 1: forward was LN(x + dropout(t)): the gradient handed to t's producers (dx_bf16, dbias) carries the mask, the
  //              residual path (dx_f32) does not.   drop_mode 2: forward was dropout(LN(x)): the incoming gradient is masked first.
  // The per-column partial sums (dgamma, dbeta, dbias) live in per-warp shared-memory rows (a lane only ever touches its own
  // columns, so plain read-modify-write is race-free): registers stay <= 128 and two CTAs fit per SM for latency hiding.
  extern __shared__ float acc_smem[];            // [nwarp][3][d]
  const DropCtx dc = drop_ctx(da);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarp = blockDim.x >> 5;
  float* acc = acc_smem + (size_t)warp * 3 * d;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 4;
    st_f4(acc + c, make_float4(0.f, 0.f, 0.f, 0.f)); st_f4(acc + d + c, make_float4(0.f, 0.f, 0.f, 0.f)); st_f4(acc + 2 * d + c, make_float4(0.f, 0.f, 0.f, 0.f));
  }
  for (int row = blockIdx.x * nwarp + warp; row < rows; row += gridDim.x * nwarp) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    const float* xr = x + (long long)row * ldx;
    float4 xh[NV], dvv[NV];
    float s1 = 0.f, s2 = 0.f;
    // the residual gradient added to dx is only needed after the row reduction; fetching it into L1 now keeps its DRAM round trip
    // off the critical path without holding registers (the kernel is latency bound: ncu long_scoreboard.  Measured +6 % GB/s;
    // loading it into registers here instead spills at 128 registers and gains nothing)
    if (dx_add) {
#pragma unroll
      for (int i = 0; i < NV; ++i) asm volatile("prefetch.global.L1 [%0];" ::"l"(dx_add + (long long)row * d + (lane + 32 * i) * 4));
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 4;
      float4 xv = ld_f4(xr + c);
      float4 dv = dy_is_f32 ? ld_f4(reinterpret_cast<const float*>(dy) + (long long)row * d + c)
                            : ld_bf4(reinterpret_cast<const bf16*>(dy) + (long long)row * d + c);
      if (dy_add) { float4 a = ld_f4(dy_add + (long long)row * d + c); dv.x += a.x; dv.y += a.y; dv.z += a.z; dv.w += a.w; }
      if (dc.on && drop_mode == 2) { const float4 m = drop_mult4(dc, row, lane + 32 * i); dv.x *= m.x; dv.y *= m.y; dv.z *= m.z; dv.w *= m.w; }
      xh[i].x = (xv.x - mean) * rstd; xh[i].y = (xv.y - mean) * rstd; xh[i].z = (xv.z - mean) * rstd; xh[i].w = (xv.w - mean) * rstd;
      dvv[i] = dv;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float gx = dv.x * gm.x, gy = dv.y * gm.y, gz = dv.z * gm.z, gw = dv.w * gm.w;
      s1 += gx + gy + gz + gw;
      s2 += gx * xh[i].x + gy * xh[i].y + gz * xh[i].z + gw * xh[i].w;
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 4;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 dv = dvv[i];
      float4 o;
      o.x = rstd * (dv.x * gm.x - s1 - xh[i].x * s2);
      o.y = rstd * (dv.y * gm.y - s1 - xh[i].y * s2);
      o.z = rstd * (dv.z * gm.z - s1 - xh[i].z * s2);
      o.w = rstd * (dv.w * gm.w - s1 - xh[i].w * s2);
      if (dx_add) { float4 a = ld_f4(dx_add + (long long)row * d + c); o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
      if (dx_f32) st_f4(dx_f32 + (long long)row * lddx + c, o);
      if (dc.on && drop_mode == 1) { const float4 m = drop_mult4(dc, row, lane + 32 * i); o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w; }
      if (dx_bf16) st_bf4(dx_bf16 + (long long)row * d + c, o);
      float4 a0 = ld_f4(acc + c), a1 = ld_f4(acc + d + c), a2 = ld_f4(acc + 2 * d + c);
      a0.x += dv.x * xh[i].x; a0.y += dv.y * xh[i].y; a0.z += dv.z * xh[i].z; a0.w += dv.w * xh[i].w;
      a1.x += dv.x; a1.y += dv.y; a1.z += dv.z; a1.w += dv.w;
      a2.x += o.x; a2.y += o.y; a2.z += o.z; a2.w += o.w;
      st_f4(acc + c, a0); st_f4(acc + d + c, a1); st_f4(acc + 2 * d + c, a2);
    }
  }
  // cross-warp reduction of the per-column partials, then one atomic per column per CTA
  __syncthreads();
  for (int t = threadIdx.x; t < 3 * d; t += blockDim.x) {
    float sum = 0.f;
    for (int w = 0; w < nwarp; ++w) sum += acc_smem[(size_t)w * 3 * d + t];
    const int which = t / d, col = t - which * d;
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dbias);
    if (dst) atomicAdd(dst + col, sum);
  }
}

// ------------------------------------------------------------------------------------------------ column sum
// out[n] += sum_rows x[rows, n]  (bias gradients; also d(pos-embed), d(type-embed)); n % 4 == 0
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ x, int is_f32, long long ldx, float* __restrict__ out,
                                                     int rows, int n, int rows_per_cta) {
  const int c4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c4 >= n) return;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(rows, r0 + rows_per_cta);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  int r = r0;
  // 8 independent row loads in flight per thread: the kernel is pure HBM streaming
  for (; r + 8 <= r1; r += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      v[u] = is_f32 ? ld_f4(reinterpret_cast<const float*>(x) + (long long)(r + u) * ldx + c4)
                    : ld_bf4(reinterpret_cast<const bf16*>(x) + (long long)(r + u) * ldx + c4);
#pragma unroll
    for (int u = 0; u < 8; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
  }
  for (; r < r1; ++r) {
    float4 v = is_f32 ? ld_f4(reinterpret_cast<const float*>(x) + (long long)r * ldx + c4)
                      : ld_bf4(reinterpret_cast<const bf16*>(x) + (long long)r * ldx + c4);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  atomicAdd(out + c4, a.x); atomicAdd(out + c4 + 1, a.y); atomicAdd(out + c4 + 2, a.z); atomicAdd(out + c4 + 3, a.w);
}

// ------------------------------------------------------------------------------------------------ patch im2col
// pixels f32 [B,3,R,R] -> patches bf16 [B*g*g, 3*P*P], column = c*P*P + i*P + j  (== conv1.weight.view(W, -1) order,
// modeling_chineseclip.py:224,237).  P % 2 == 0.
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ pix, bf16* __restrict__ out, int B, int R, int P, int g, int ld) {
  const int kdim = 3 * P * P;
  const long long total2 = (long long)B * g * g * kdim / 2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total2; t += (long long)gridDim.x * blockDim.x) {
    const long long e = t * 2;
    const int col = (int)(e % kdim);
    const long long rowi = e / kdim;
    const int px = (int)(rowi % g), py = (int)((rowi / g) % g), b = (int)(rowi / ((long long)g * g));
    const int c = col / (P * P), i = (col / P) % P, j = col % P;
    const float2 v = *reinterpret_cast<const float2*>(pix + (((long long)b * 3 + c) * R + (py * P + i)) * R + px * P + j);
    *reinterpret_cast<uint32_t*>(out + rowi * ld + col) = pack_bf16x2(v.x, v.y);
  }
}

// x0[b, l, :] = (l == 0 ? class_embedding : patch[b, l-1, :]) + positional_embedding[l]   (modeling_chineseclip.py:238-241)
__global__ void __launch_bounds__(256) vit_assemble_kernel(const float* __restrict__ patch, const float* __restrict__ cls,
                                                           const float* __restrict__ pos, float* __restrict__ x0, int B, int L, int W) {
  const long long total4 = (long long)B * L * W / 4;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total4; t += (long long)gridDim.x * blockDim.x) {
    const long long e = t * 4;
    const int c = (int)(e % W);
    const long long rl = e / W;
    const int l = (int)(rl % L);
    const long long b = rl / L;
    float4 v = l == 0 ? ld_f4(cls + c) : ld_f4(patch + ((b * (L - 1) + (l - 1)) * W + c));
    float4 p = ld_f4(pos + (long long)l * W + c);
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    st_f4(x0 + e, v);
  }
}
// backward of the assembly: dpatch bf16 [B*(L-1), W] = dx0[b, 1+p, :]  (dpos / dcls are column sums taken by the host wrapper)
__global__ void __launch_bounds__(256) vit_assemble_bwd_kernel(const float* __restrict__ dx0, bf16* __restrict__ dpatch, int B, int L, int W) {
  const long long total4 = (long long)B * (L - 1) * W / 4;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total4; t += (long long)gridDim.x * blockDim.x) {
    const long long e = t * 4;
    const int c = (int)(e % W);
    const long long rl = e / W;
    const int p = (int)(rl % (L - 1));
    const long long b = rl / (L - 1);
    st_bf4(dpatch + e, ld_f4(dx0 + ((b * L + p + 1) * W + c)));
  }
}

// ------------------------------------------------------------------------------------------------ BERT embeddings
// e[b,l,:] = word[ids[b,l]] + type[0] + pos[l]       (modeling_bert.py:95-129; LN applied by layernorm_fwd)
__global__ void __launch_bounds__(256) bert_embed_kernel(const long long* __restrict__ ids, const float* __restrict__ word,
                                                         const float* __restrict__ pos, const float* __restrict__ type0,
                                                         float* __restrict__ e, float* __restrict__ key_mask, int rows, int L, int H,
                                                         int vocab) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  long long id = ids[row];
  if (id < 0 || id >= vocab) id = 0;  // host validates; never index out of the table
  // attention_mask = ids.ne(0) (modeling_chineseclip.py:347-348) as the additive (1-m)*-10000 of modeling_utils.py:438-439
  if (key_mask && lane == 0) key_mask[row] = (id == 0) ? -10000.0f : 0.0f;
  const int l = row % L;
  for (int c = lane * 4; c < H; c += 128) {
    float4 w = ld_f4(word + id * H + c), p = ld_f4(pos + (long long)l * H + c), t = ld_f4(type0 + c);
    st_f4(e + (long long)row * H + c, make_float4(w.x + p.x + t.x, w.y + p.y + t.y, w.z + p.z + t.z, w.w + p.w + t.w));
  }
}
// dword[ids] += de (row 0 = padding_idx gets no gradient, modeling_bert.py:77)
__global__ void __launch_bounds__(256) bert_embed_bwd_kernel(const long long* __restrict__ ids, const float* __restrict__ de,
                                                             float* __restrict__ dword, int rows, int H, int vocab) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long id = ids[row];
  if (id <= 0 || id >= vocab) return;
  for (int c = lane * 4; c < H; c += 128) {
    float4 v = ld_f4(de + (long long)row * H + c);
    atomicAdd(reinterpret_cast<float4*>(dword + id * H + c), v);
  }
}

// ------------------------------------------------------------------------------------------------ RoBERTa-style embeddings
// position ids of create_position_ids_from_input_ids (modeling_roberta.py:1497-1510): pos = cumsum(ids != pad) * (ids != pad) + pad.
// One warp per sequence: 32 tokens per step, warp-inclusive scan of the 0/1 flags plus the running count.
__global__ void __launch_bounds__(256) position_ids_kernel(const long long* __restrict__ ids, int* __restrict__ pos, int B, int L, int pad) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  int run = 0;
  for (int l0 = 0; l0 < L; l0 += 32) {
    const int l = l0 + lane;
    const int m = (l < L && ids[(long long)b * L + l] != pad) ? 1 : 0;
    int x = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (l < L) pos[(long long)b * L + l] = (run + x) * m + pad;
    run += __shfl_sync(0xffffffffu, x, 31);
  }
}
// e[row,:] = word[ids[row]] + pos_table[pos_ids[row]] + type_table[type_ids[row]]  (RobertaEmbeddings.forward, modeling_roberta.py:100-130)
// key_mask[row] = (1 - attention_mask[row]) * -10000 (modeling_utils.py:438-439), or from ids != pad when no mask is given
__global__ void __launch_bounds__(256) embed_gather_kernel(const long long* __restrict__ ids, const int* __restrict__ pos_ids,
                                                           const long long* __restrict__ type_ids, const long long* __restrict__ attn_mask,
                                                           const float* __restrict__ word, const float* __restrict__ pos,
                                                           const float* __restrict__ type, float* __restrict__ e, float* __restrict__ key_mask,
                                                           int rows, int H, int vocab, int npos, int ntype, int pad) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  long long id = ids[row];
  if (id < 0 || id >= vocab) id = pad >= 0 ? pad : 0;
  int pi = pos_ids[row]; if (pi < 0 || pi >= npos) pi = 0;
  long long ti = type_ids ? type_ids[row] : 0; if (ti < 0 || ti >= ntype) ti = 0;
  if (key_mask && lane == 0) key_mask[row] = attn_mask ? (1.0f - (float)attn_mask[row]) * -10000.0f : (id == pad ? -10000.0f : 0.0f);
  for (int c = lane * 4; c < H; c += 128) {
    float4 w = ld_f4(word + id * H + c), p = ld_f4(pos + (long long)pi * H + c), t = ld_f4(type + ti * H + c);
    st_f4(e + (long long)row * H + c, make_float4(w.x + p.x + t.x, w.y + p.y + t.y, w.z + p.z + t.z, w.w + p.w + t.w));
  }
}
// scatter of de into the three tables; the padding rows of nn.Embedding(padding_idx=pad) (word and position tables) get no gradient
__global__ void __launch_bounds__(256) embed_gather_bwd_kernel(const long long* __restrict__ ids, const int* __restrict__ pos_ids,
                                                               const long long* __restrict__ type_ids, const float* __restrict__ de,
                                                               float* __restrict__ dword, float* __restrict__ dpos, float* __restrict__ dtype,
                                                               int rows, int H, int vocab, int npos, int ntype, int pad) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long id = ids[row];
  const int pi = pos_ids[row];
  long long ti = type_ids ? type_ids[row] : 0; if (ti < 0 || ti >= ntype) ti = 0;
  const bool wok = id >= 0 && id < vocab && id != pad, pok = pi >= 0 && pi < npos && pi != pad;
  for (int c = lane * 4; c < H; c += 128) {
    const float4 v = ld_f4(de + (long long)row * H + c);
    if (wok) atomicAdd(reinterpret_cast<float4*>(dword + id * H + c), v);
    if (pok) atomicAdd(reinterpret_cast<float4*>(dpos + (long long)pi * H + c), v);
    atomicAdd(reinterpret_cast<float4*>(dtype + ti * H + c), v);
  }
}
// idx[b] = argmax_l ids[b, l] (first maximum): OPEN_CLIP.encode_text pools the features at the EOT token, the highest id of each
// sequence (modeling_openclip.py:367-369).  One warp per sequence.
__global__ void __launch_bounds__(256) argmax_rows_kernel(const long long* __restrict__ ids, int* __restrict__ idx, int B, int L) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  long long best = LLONG_MIN; int bi = 0;
  for (int l = lane; l < L; l += 32) { const long long v = ids[(long long)b * L + l]; if (v > best) { best = v; bi = l; } }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) idx[b] = bi;
}
// idx[b] = first l with ids[b, l] == token (0 when there is none), count[b] (optional) = number of matches: Wukong's TextTransformer pools
// the [SEP] position, `x[(ids == 102).nonzero()]` (modeling_wukong.py:349,359), which needs exactly one per sequence.  One warp per sequence.
__global__ void __launch_bounds__(256) find_token_rows_kernel(const long long* __restrict__ ids, long long token, int* __restrict__ idx,
                                                              int* __restrict__ count, int B, int L) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  int first = INT_MAX, n = 0;
  for (int l = lane; l < L; l += 32)
    if (ids[(long long)b * L + l] == token) { first = min(first, l); ++n; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { first = min(first, __shfl_xor_sync(0xffffffffu, first, o)); n += __shfl_xor_sync(0xffffffffu, n, o); }
  if (lane == 0) { idx[b] = first == INT_MAX ? 0 : first; if (count) count[b] = n; }
}
// out[b, :] = x[b * L + idx[b], :] (bf16 rows) ; dst[b * L + idx[b], :] = src[b, :] (f32 rows; dst zero-filled by the caller)
__global__ void __launch_bounds__(256) gather_rows_bf16_kernel(const bf16* __restrict__ x, const int* __restrict__ idx, bf16* __restrict__ out, int B, int L, int W) {
  const int b = blockIdx.x;
  const bf16* src = x + ((long long)b * L + idx[b]) * W;
  for (int c = threadIdx.x * 8; c < W; c += blockDim.x * 8) *reinterpret_cast<uint4*>(out + (long long)b * W + c) = *reinterpret_cast<const uint4*>(src + c);
}
__global__ void __launch_bounds__(256) scatter_rows_f32_kernel(const float* __restrict__ src, const int* __restrict__ idx, float* __restrict__ dst, int B, int L, int W) {
  const int b = blockIdx.x;
  float* d = dst + ((long long)b * L + idx[b]) * W;
  for (int c = threadIdx.x * 4; c < W; c += blockDim.x * 4) st_f4(d + c, ld_f4(src + (long long)b * W + c));
}

// masked mean over the T frames of a video (Text2VideoRetrieval._mean_pooling_for_similarity_visual, appzoo/text2video_retrieval/model.py:98-104):
// out[b, :] = sum_t mask[b,t] * x[b,t,:] / max(sum_t mask[b,t], 1) ; backward: dx[b,t,:] = mask[b,t] * dout[b,:] / count
__global__ void __launch_bounds__(256) frame_pool_fwd_kernel(const float* __restrict__ x, const long long* __restrict__ mask, float* __restrict__ out,
                                                             int B, int T, int E) {
  const int b = blockIdx.x;
  float cnt = 0.f;
  for (int t = 0; t < T; ++t) cnt += (float)mask[(long long)b * T + t];
  const float inv = 1.0f / (cnt == 0.f ? 1.f : cnt);
  for (int c = threadIdx.x * 4; c < E; c += blockDim.x * 4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; ++t) {
      const float m = (float)mask[(long long)b * T + t];
      const float4 v = ld_f4(x + ((long long)b * T + t) * E + c);
      a.x += m * v.x; a.y += m * v.y; a.z += m * v.z; a.w += m * v.w;
    }
    st_f4(out + (long long)b * E + c, make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv));
  }
}
__global__ void __launch_bounds__(256) frame_pool_bwd_kernel(const float* __restrict__ dout, const long long* __restrict__ mask, float* __restrict__ dx,
                                                             int B, int T, int E) {
  const int b = blockIdx.x;
  float cnt = 0.f;
  for (int t = 0; t < T; ++t) cnt += (float)mask[(long long)b * T + t];
  const float inv = 1.0f / (cnt == 0.f ? 1.f : cnt);
  for (int c = threadIdx.x * 4; c < E; c += blockDim.x * 4) {
    const float4 g = ld_f4(dout + (long long)b * E + c);
    for (int t = 0; t < T; ++t) {
      const float m = (float)mask[(long long)b * T + t] * inv;
      st_f4(dx + ((long long)b * T + t) * E + c, make_float4(m * g.x, m * g.y, m * g.z, m * g.w));
    }
  }
}

// y = tanh(x) (BertPooler / RobertaPooler activation, modeling_bert.py:529-541); dx = dy * (1 - y^2)
__global__ void __launch_bounds__(256) tanh_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, bf16* __restrict__ y_bf16, long long n4) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
    const float4 a = ld_f4(x + t * 4);
    const float4 o = make_float4(tanhf(a.x), tanhf(a.y), tanhf(a.z), tanhf(a.w));
    st_f4(y + t * 4, o);
    if (y_bf16) st_bf4(y_bf16 + t * 4, o);
  }
}
__global__ void __launch_bounds__(256) tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                                                       bf16* __restrict__ dx_bf16, long long n4) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
    const float4 g = ld_f4(dy + t * 4), v = ld_f4(y + t * 4);
    const float4 o = make_float4(g.x * (1.f - v.x * v.x), g.y * (1.f - v.y * v.y), g.z * (1.f - v.z * v.z), g.w * (1.f - v.w * v.w));
    if (dx) st_f4(dx + t * 4, o);
    if (dx_bf16) st_bf4(dx_bf16 + t * 4, o);
  }
}

// ------------------------------------------------------------------------------------------------ L2 normalise
// y = x / ||x||  (modeling_chineseclip.py:360,363); bwd: dx = (dy - y * <dy, y>) / ||x||
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ norm_out,
                                                         int rows, int d) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane * 4; c < d; c += 128) { float4 v = ld_f4(x + (long long)row * d + c); s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
  const float nrm = sqrtf(warp_sum(s));
  const float inv = 1.0f / nrm;
  if (lane == 0 && norm_out) norm_out[row] = nrm;
  for (int c = lane * 4; c < d; c += 128) {
    float4 v = ld_f4(x + (long long)row * d + c);
    st_f4(y + (long long)row * d + c, make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv));
  }
}
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ norm,
                                                         float* __restrict__ dx_f32, bf16* __restrict__ dx_bf16, int rows, int d) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane * 4; c < d; c += 128) {
    float4 a = ld_f4(dy + (long long)row * d + c), b = ld_f4(y + (long long)row * d + c);
    s += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  }
  s = warp_sum(s);
  const float inv = 1.0f / norm[row];
  for (int c = lane * 4; c < d; c += 128) {
    float4 a = ld_f4(dy + (long long)row * d + c), b = ld_f4(y + (long long)row * d + c);
    float4 o = make_float4((a.x - b.x * s) * inv, (a.y - b.y * s) * inv, (a.z - b.z * s) * inv, (a.w - b.w * s) * inv);
    if (dx_f32) st_f4(dx_f32 + (long long)row * d + c, o);
    if (dx_bf16) st_bf4(dx_bf16 + (long long)row * d + c, o);
  }
}

// f32 -> bf16 cast of a contiguous buffer (n % 4 == 0)
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n4) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x)
    st_bf4(y + t * 4, ld_f4(x + t * 4));
}

// y += alpha * x   (n % 4 == 0)
__global__ void __launch_bounds__(256) axpy_kernel(const float* __restrict__ x, float* __restrict__ y, float alpha, long long n4) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
    float4 a = ld_f4(x + t * 4), b = ld_f4(y + t * 4);
    st_f4(y + t * 4, make_float4(b.x + alpha * a.x, b.y + alpha * a.y, b.z + alpha * a.z, b.w + alpha * a.w));
  }
}

// multipliers of a [rows, cols] dropout site, exactly as drop_mult4 yields them inside the fused kernels
__global__ void __launch_bounds__(256) dropout_mask_kernel(float* __restrict__ out, int rows, int cols, const DropArg da) {
  const DropCtx dc = drop_ctx(da);
  const long long total4 = (long long)rows * cols / 4;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total4; t += (long long)gridDim.x * blockDim.x) {
    const uint32_t row = (uint32_t)(t / (cols / 4)), quad = (uint32_t)(t % (cols / 4));
    st_f4(out + t * 4, dc.on ? drop_mult4(dc, row, quad) : make_float4(1.f, 1.f, 1.f, 1.f));
  }
}

static inline int grid_for(long long work_items, int per_cta) {
  long long g = (work_items + per_cta - 1) / per_cta;
  long long cap = (long long)sm_count() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace clipk

using namespace clipk;

#define LN_DISPATCH(NV, ...)                                       \
  switch (NV) {                                                    \
    case 1: __VA_ARGS__(1); break; case 2: __VA_ARGS__(2); break;  \
    case 3: __VA_ARGS__(3); break; case 4: __VA_ARGS__(4); break;  \
    case 5: __VA_ARGS__(5); break; case 6: __VA_ARGS__(6); break;  \
    case 7: __VA_ARGS__(7); break; case 8: __VA_ARGS__(8); break;  \
    default: set_error("layernorm: d=%d unsupported (d %% 128 == 0, d <= 1024)", d); return CLIPK_ERR_UNSUPPORTED; }

extern "C" int clipk_layernorm_fwd(const float* x, long long ldx, const void* add_bf16, long long ldadd, float* x_out, const float* gamma,
                                   const float* beta, float eps, void* y_bf16, float* y_f32, float* mean, float* rstd, int rows, int d,
                                   const clipk_dropout_t* drop, int drop_mode, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (d % 128 || d > 128 * LN_MAXV || (ldx % 4) || (ldadd % 4)) { set_error("layernorm_fwd: d=%d ldx=%lld unsupported", d, ldx); return CLIPK_ERR_UNSUPPORTED; }
  const int nv = d / 128;
  dim3 grid((rows + 7) / 8), block(256);
  const DropArg da = make_drop_arg(drop);
#define LAUNCH(NV) layernorm_fwd_kernel<NV><<<grid, block, 0, stream>>>(x, ldx, (const bf16*)add_bf16, ldadd, x_out, gamma, beta, eps, (bf16*)y_bf16, y_f32, mean, rstd, rows, d, da, drop_mode)
  LN_DISPATCH(nv, LAUNCH)
#undef LAUNCH
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_layernorm_bwd(const void* dy, int dy_is_f32, const float* dy_add, const float* x, long long ldx, const float* gamma,
                                   const float* mean, const float* rstd, const float* dx_add, float* dx_f32, long long lddx,
                                   void* dx_bf16, float* dgamma, float* dbeta, float* dbias, int rows, int d, const clipk_dropout_t* drop,
                                   int drop_mode, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (d % 128 || d > 128 * LN_MAXV || (ldx % 4) || (lddx % 4)) { set_error("layernorm_bwd: d=%d unsupported", d); return CLIPK_ERR_UNSUPPORTED; }
  const int nv = d / 128;
  int g = (rows + 7) / 8;
  // 2 CTAs / SM resident (<= 128 registers).  One wave by default: every CTA pays a fixed cost (zeroing / reducing its 72 KB of
  // column partials, 3d atomics), so fewer, longer-lived CTAs win over finer load balancing (CLIPK_LN_BWD_WAVES for A/B runs)
  static int waves = -1;
  if (waves < 0) { const char* ev = getenv("CLIPK_LN_BWD_WAVES"); waves = ev ? atoi(ev) : 1; if (waves < 1) waves = 1; }
  const int cap = sm_count() * 2 * waves;
  if (g > cap) g = cap;
  dim3 grid(g), block(256);
  const DropArg da = make_drop_arg(drop);
  const int smem = 8 * 3 * d * (int)sizeof(float);
#define LAUNCH(NV)                                                                                                               \
  {                                                                                                                              \
    static bool cfg = false;                                                                                                     \
    if (!cfg) { CLIPK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 3 * 1024 * 4)); cfg = true; } \
    layernorm_bwd_kernel<NV><<<grid, block, smem, stream>>>(dy, dy_is_f32, dy_add, x, ldx, gamma, mean, rstd, dx_add, dx_f32, lddx, (bf16*)dx_bf16, dgamma, dbeta, dbias, rows, d, da, drop_mode); \
  }
  LN_DISPATCH(nv, LAUNCH)
#undef LAUNCH
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_colsum(const void* x, int is_f32, long long ldx, float* out, int rows, int n, cudaStream_t stream) {
  if (rows <= 0 || n <= 0) return 0;
  if (n % 4 || ldx % 4) { set_error("colsum: n=%d ldx=%lld must be multiples of 4", n, ldx); return CLIPK_ERR_ARG; }
  const int bx = (n / 4 + 255) / 256;
  int by = (sm_count() * 16 + bx - 1) / bx;
  int rows_per = (rows + by - 1) / by;
  if (rows_per < 32) rows_per = 32;
  by = (rows + rows_per - 1) / rows_per;
  colsum_kernel<<<dim3(bx, by), 256, 0, stream>>>(x, is_f32, ldx, out, rows, n, rows_per);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_im2col_patches(const float* pixels, void* patches_bf16, int B, int R, int P, int ld_out, cudaStream_t stream) {
  if (R % P || P % 2) { set_error("im2col: R=%d P=%d unsupported", R, P); return CLIPK_ERR_UNSUPPORTED; }
  const int g = R / P;
  if (ld_out <= 0) ld_out = 3 * P * P;
  if (ld_out < 3 * P * P || (ld_out % 2)) { set_error("im2col: ld_out=%d < 3*P*P or odd", ld_out); return CLIPK_ERR_ARG; }
  const long long total2 = (long long)B * g * g * 3 * P * P / 2;
  im2col_kernel<<<grid_for(total2, 256 * 4), 256, 0, stream>>>(pixels, (bf16*)patches_bf16, B, R, P, g, ld_out);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_vit_assemble(const float* patch, const float* cls, const float* pos, float* x0, int B, int L, int W, cudaStream_t stream) {
  if (W % 4) { set_error("vit_assemble: W %% 4 != 0"); return CLIPK_ERR_ARG; }
  vit_assemble_kernel<<<grid_for((long long)B * L * W / 4, 256 * 4), 256, 0, stream>>>(patch, cls, pos, x0, B, L, W);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_vit_assemble_bwd(const float* dx0, void* dpatch_bf16, int B, int L, int W, cudaStream_t stream) {
  if (W % 4) { set_error("vit_assemble_bwd: W %% 4 != 0"); return CLIPK_ERR_ARG; }
  vit_assemble_bwd_kernel<<<grid_for((long long)B * (L - 1) * W / 4, 256 * 4), 256, 0, stream>>>(dx0, (bf16*)dpatch_bf16, B, L, W);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_bert_embed(const long long* ids, const float* word, const float* pos, const float* type0, float* e, float* key_mask,
                                int rows, int L, int H, int vocab, cudaStream_t stream) {
  if (H % 4) { set_error("bert_embed: H %% 4 != 0"); return CLIPK_ERR_ARG; }
  bert_embed_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(ids, word, pos, type0, e, key_mask, rows, L, H, vocab);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_bert_embed_bwd(const long long* ids, const float* de, float* dword, int rows, int H, int vocab, cudaStream_t stream) {
  if (H % 4) { set_error("bert_embed_bwd: H %% 4 != 0"); return CLIPK_ERR_ARG; }
  bert_embed_bwd_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(ids, de, dword, rows, H, vocab);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_position_ids(const long long* ids, int* pos_ids, int B, int L, int pad_id, cudaStream_t stream) {
  if (B <= 0 || L <= 0) return 0;
  position_ids_kernel<<<(B + 7) / 8, 256, 0, stream>>>(ids, pos_ids, B, L, pad_id);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_embed_gather(const long long* ids, const int* pos_ids, const long long* type_ids, const long long* attn_mask,
                                  const float* word, const float* pos, const float* type, float* e, float* key_mask, int rows, int H,
                                  int vocab, int npos, int ntype, int pad_id, cudaStream_t stream) {
  if (H % 4) { set_error("embed_gather: H %% 4 != 0"); return CLIPK_ERR_ARG; }
  embed_gather_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(ids, pos_ids, type_ids, attn_mask, word, pos, type, e, key_mask, rows, H, vocab, npos, ntype, pad_id);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_embed_gather_bwd(const long long* ids, const int* pos_ids, const long long* type_ids, const float* de, float* dword,
                                      float* dpos, float* dtype, int rows, int H, int vocab, int npos, int ntype, int pad_id,
                                      cudaStream_t stream) {
  if (H % 4) { set_error("embed_gather_bwd: H %% 4 != 0"); return CLIPK_ERR_ARG; }
  embed_gather_bwd_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(ids, pos_ids, type_ids, de, dword, dpos, dtype, rows, H, vocab, npos, ntype, pad_id);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_argmax_rows(const long long* ids, int* idx, int B, int L, cudaStream_t stream) {
  if (B <= 0 || L <= 0) return 0;
  argmax_rows_kernel<<<(B + 7) / 8, 256, 0, stream>>>(ids, idx, B, L);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int clipk_find_token_rows(const long long* ids, long long token, int* idx, int* count, int B, int L, cudaStream_t stream) {
  if (B <= 0 || L <= 0) return 0;
  find_token_rows_kernel<<<(B + 7) / 8, 256, 0, stream>>>(ids, token, idx, count, B, L);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int clipk_gather_rows_bf16(const void* x, const int* idx, void* out, int B, int L, int W, cudaStream_t stream) {
  if (W % 8) { set_error("gather_rows: W %% 8 != 0"); return CLIPK_ERR_ARG; }
  if (B <= 0) return 0;
  gather_rows_bf16_kernel<<<B, 128, 0, stream>>>((const bf16*)x, idx, (bf16*)out, B, L, W);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int clipk_scatter_rows_f32(const float* src, const int* idx, float* dst, int B, int L, int W, cudaStream_t stream) {
  if (W % 4) { set_error("scatter_rows: W %% 4 != 0"); return CLIPK_ERR_ARG; }
  if (B <= 0) return 0;
  scatter_rows_f32_kernel<<<B, 128, 0, stream>>>(src, idx, dst, B, L, W);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_frame_pool_fwd(const float* x, const long long* mask, float* out, int B, int T, int E, cudaStream_t stream) {
  if (E % 4) { set_error("frame_pool: E %% 4 != 0"); return CLIPK_ERR_ARG; }
  if (B <= 0) return 0;
  frame_pool_fwd_kernel<<<B, 128, 0, stream>>>(x, mask, out, B, T, E);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int clipk_frame_pool_bwd(const float* dout, const long long* mask, float* dx, int B, int T, int E, cudaStream_t stream) {
  if (E % 4) { set_error("frame_pool: E %% 4 != 0"); return CLIPK_ERR_ARG; }
  if (B <= 0) return 0;
  frame_pool_bwd_kernel<<<B, 128, 0, stream>>>(dout, mask, dx, B, T, E);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_tanh_fwd(const float* x, float* y, void* y_bf16, long long n, cudaStream_t stream) {
  if (n % 4) { set_error("tanh_fwd: n %% 4 != 0"); return CLIPK_ERR_ARG; }
  if (n == 0) return 0;
  tanh_fwd_kernel<<<grid_for(n / 4, 256 * 4), 256, 0, stream>>>(x, y, (bf16*)y_bf16, n / 4);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_tanh_bwd(const float* dy, const float* y, float* dx, void* dx_bf16, long long n, cudaStream_t stream) {
  if (n % 4) { set_error("tanh_bwd: n %% 4 != 0"); return CLIPK_ERR_ARG; }
  if (n == 0) return 0;
  tanh_bwd_kernel<<<grid_for(n / 4, 256 * 4), 256, 0, stream>>>(dy, y, dx, (bf16*)dx_bf16, n / 4);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_l2norm_fwd(const float* x, float* y, float* norm, int rows, int d, cudaStream_t stream) {
  if (d % 4) { set_error("l2norm: d %% 4 != 0"); return CLIPK_ERR_ARG; }
  l2norm_fwd_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(x, y, norm, rows, d);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_l2norm_bwd(const float* dy, const float* y, const float* norm, float* dx_f32, void* dx_bf16, int rows, int d,
                                cudaStream_t stream) {
  if (d % 4) { set_error("l2norm: d %% 4 != 0"); return CLIPK_ERR_ARG; }
  l2norm_bwd_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(dy, y, norm, dx_f32, (bf16*)dx_bf16, rows, d);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_cast_bf16(const float* x, void* y, long long n, cudaStream_t stream) {
  if (n % 4) { set_error("cast_bf16: n %% 4 != 0"); return CLIPK_ERR_ARG; }
  if (n == 0) return 0;
  cast_bf16_kernel<<<grid_for(n / 4, 256 * 4), 256, 0, stream>>>(x, (bf16*)y, n / 4);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_axpy(const float* x, float* y, float alpha, long long n, cudaStream_t stream) {
  if (n % 4) { set_error("axpy: n %% 4 != 0"); return CLIPK_ERR_ARG; }
  if (n == 0) return 0;
  axpy_kernel<<<grid_for(n / 4, 256 * 4), 256, 0, stream>>>(x, y, alpha, n / 4);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_dropout_mask(float* out, int rows, int cols, const clipk_dropout_t* drop, cudaStream_t stream) {
  if (cols % 4) { set_error("dropout_mask: cols %% 4 != 0"); return CLIPK_ERR_ARG; }
  dropout_mask_kernel<<<grid_for((long long)rows * cols / 4, 256 * 4), 256, 0, stream>>>(out, rows, cols, make_drop_arg(drop));
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}
