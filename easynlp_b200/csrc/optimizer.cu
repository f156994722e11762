// Multi-tensor optimizer step over ONE flat fp32 parameter buffer: global-norm clipping + the reference's AdamW in a
// single pass that also refreshes the bf16 weight copy the GEMMs read.
//   reference: core/trainer.py:315-325 (clip_grad_norm_(max_grad_norm) when a scheduler exists),
//              core/optimizers.py:437-462 (AdamW: eps added to sqrt(v) before bias correction, decoupled decay
//              `p -= lr * wd * p` AFTER the Adam update), one Python-loop launch chain per tensor (~2k tiny kernels).
#include "common.cuh"
#include "../../include/clipk.h"

namespace clipk {

__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long long n4, double* __restrict__ partials) {
  __shared__ double red[8];
  double s = 0.0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[t];
    s += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}

// One thread: step += 1; hyper = {lr_k, lr_k * sqrt(1 - b2^k) / (1 - b1^k)} with lr_k = base_lr * lambda(k - 1), lambda = the
// reference's EasyNLPWarmupLinearSchedule (core/optimizers.py:191-204; scheduler.step() runs AFTER optimizer.step(), so step k uses
// lambda(k-1)).  t_total <= 0 means a constant learning rate.
__global__ void adam_schedule_kernel(int* __restrict__ step, float* __restrict__ hyper, float base_lr, int warmup_steps, int t_total,
                                     float beta1, float beta2) {
  const int k = step[0] + 1;
  step[0] = k;
  float lam = 1.f;
  if (t_total > 0) {
    const int s = k - 1;
    if (s < warmup_steps) lam = (float)s / (float)max(1, warmup_steps);
    else lam = fmaxf(0.f, (float)(t_total - s) / fmaxf(1.f, (float)(t_total - warmup_steps)));
  }
  const float lr = base_lr * lam;
  const double bc1 = 1.0 - pow((double)beta1, (double)k), bc2 = 1.0 - pow((double)beta2, (double)k);
  hyper[0] = lr;
  hyper[1] = (float)((double)lr * sqrt(bc2) / bc1);
}

// norm_out[0] = sqrt(sum partials) ; norm_out[1] = clip coefficient min(1, max_norm / (norm + 1e-6))
__global__ void __launch_bounds__(256) gradnorm_final_kernel(const double* __restrict__ partials, int n, float max_norm, float* __restrict__ norm_out) {
  __shared__ double red[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    const float nrm = (float)sqrt(t);
    norm_out[0] = nrm;
    const float c = max_norm / (nrm + 1e-6f);
    norm_out[1] = (max_norm > 0.f && c < 1.f) ? c : 1.f;
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, bf16* __restrict__ w_bf16, long long n4, float lr, float beta1,
                                                    float beta2, float eps, float weight_decay, float step_size,
                                                    const float* __restrict__ clip_coef, const float* __restrict__ dev_hyper) {
  const float cc = clip_coef ? clip_coef[0] : 1.f;
  if (dev_hyper) { lr = dev_hyper[0]; step_size = dev_hyper[1]; }   // device-resident schedule: lets a CUDA graph replay the step
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[t];
    float4 gv = reinterpret_cast<const float4*>(g)[t];
    float4 mv = reinterpret_cast<float4*>(m)[t];
    float4 vv = reinterpret_cast<float4*>(v)[t];
    float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = gp[k] * cc;
      mp[k] = mp[k] * beta1 + (1.f - beta1) * gr;
      vp[k] = vp[k] * beta2 + (1.f - beta2) * gr * gr;
      float x = pp[k] - step_size * (mp[k] / (sqrtf(vp[k]) + eps));
      x -= lr * weight_decay * x;
      pp[k] = x;
    }
    reinterpret_cast<float4*>(p)[t] = pv;
    reinterpret_cast<float4*>(m)[t] = mv;
    reinterpret_cast<float4*>(v)[t] = vv;
    if (w_bf16) {
      uint2 o; o.x = pack_bf16x2(pv.x, pv.y); o.y = pack_bf16x2(pv.z, pv.w);
      reinterpret_cast<uint2*>(w_bf16)[t] = o;
    }
  }
}

__global__ void counter_add_kernel(int* c, int v) { c[0] += v; }

}  // namespace clipk

using namespace clipk;

extern "C" int clipk_counter_add(int* counter_dev, int value, cudaStream_t stream) {
  if (!counter_dev) { set_error("counter_add: null counter"); return CLIPK_ERR_ARG; }
  counter_add_kernel<<<1, 1, 0, stream>>>(counter_dev, value);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_grad_norm(const float* g, long long n, float max_norm, double* workspace, int workspace_len, float* norm_and_coef,
                               cudaStream_t stream) {
  if (n % 4) { set_error("grad_norm: n %% 4 != 0"); return CLIPK_ERR_ARG; }
  int grid = sm_count() * 4;
  if (grid > workspace_len) grid = workspace_len;
  if (grid < 1) { set_error("grad_norm: workspace too small"); return CLIPK_ERR_ARG; }
  sumsq_partial_kernel<<<grid, 256, 0, stream>>>(g, n / 4, workspace);
  gradnorm_final_kernel<<<1, 256, 0, stream>>>(workspace, grid, max_norm, norm_and_coef);
  note_launch(2);
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_adamw_step(float* p, const float* g, float* m, float* v, void* w_bf16, long long n, float lr, float beta1, float beta2,
                                float eps, float weight_decay, int step, const float* clip_coef, const float* dev_hyper,
                                cudaStream_t stream) {
  if (n == 0) return 0;
  if (n % 4 || (step < 1 && !dev_hyper)) { set_error("adamw: n %% 4 != 0 or step < 1"); return CLIPK_ERR_ARG; }
  if (step < 1) step = 1;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr * sqrt(bc2) / bc1);
  long long n4 = n / 4;
  long long nb = (n4 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (nb > cap) nb = cap;
  adamw_kernel<<<(int)nb, 256, 0, stream>>>(p, g, m, v, (bf16*)w_bf16, n4, lr, beta1, beta2, eps, weight_decay, step_size, clip_coef, dev_hyper);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int clipk_adam_schedule(int* step_dev, float* hyper_dev, float base_lr, int warmup_steps, int t_total, float beta1, float beta2,
                                   cudaStream_t stream) {
  adam_schedule_kernel<<<1, 1, 0, stream>>>(step_dev, hyper_dev, base_lr, warmup_steps, t_total, beta1, beta2);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}
