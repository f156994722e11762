// Peer-memory collectives of the contrastive head (one process per GPU, NVLink 5 / NVSwitch): the all-gather of the two [local_B, E]
// embedding shards and the reduce-scatter of the gallery gradients written as plain loads / stores on CUDA-IPC mapped peer buffers,
// fused with their producer / consumer kernels -- north_star: "a hand-written NCCL-over-NVLink all-gather of the two [local_B, D]
// embedding shards forms the full contrastive matrix".
//
//   forward : l2norm_allgather_kernel normalises a tower's features (x / ||x||, modeling_chineseclip.py:360,363) and stores the result into
//             row block `rank` of EVERY rank's gallery (peer stores over NVLink) -- producer and all-gather are one kernel;
//   sync    : peer_signal_kernel publishes flag[channel][rank] = epoch in every peer's flag array (st.release.sys after a system fence),
//             peer_wait_kernel spins (ld.acquire.sys) until all `world` flags of a channel reached the epoch.  Producers never wait, so
//             the scheme cannot deadlock; a bounded spin traps instead of hanging the GPU;
//   backward: every rank evaluates d(loss)/d(gallery) for all global rows of ITS strips into its own buffer; peer_reduce_rows_kernel then
//             PULLS this rank's rows from all peers and sums them onto the local embedding gradients (reduce-scatter by peer loads).
// The reference has no counterpart (it never gathers); under torch.distributed these are all_gather_into_tensor /
// reduce_scatter_tensor (easynlp_b200/distributed.py keeps that path for backends without peer access).
#include "common.cuh"
#include "../../include/clipk.h"

namespace clipk {

__global__ void __launch_bounds__(256) l2norm_allgather_kernel(const float* __restrict__ x, float* __restrict__ y_local, float* __restrict__ norm_out,
                                                               float* const* __restrict__ gallery, int world, int rank, int rows, int d) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane * 4; c < d; c += 128) { float4 v = *reinterpret_cast<const float4*>(x + (long long)row * d + c); s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
  const float nrm = sqrtf(warp_sum(s));
  const float inv = 1.0f / nrm;
  if (lane == 0 && norm_out) norm_out[row] = nrm;
  const long long grow = (long long)rank * rows + row;
  for (int c = lane * 4; c < d; c += 128) {
    float4 v = *reinterpret_cast<const float4*>(x + (long long)row * d + c);
    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    *reinterpret_cast<float4*>(y_local + (long long)row * d + c) = v;
    for (int p = 0; p < world; ++p) *reinterpret_cast<float4*>(gallery[p] + grow * d + c) = v;      // peer stores (p == rank: the local copy)
  }
}

__global__ void peer_signal_kernel(unsigned int* const* __restrict__ flags, int world, int rank, int channel, unsigned int epoch) {
  const int p = threadIdx.x;
  if (p >= world) return;
  __threadfence_system();
  unsigned int* f = flags[p] + channel * world + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
}

__global__ void peer_wait_kernel(const unsigned int* __restrict__ my_flags, int world, int channel, unsigned int epoch) {
  const int p = threadIdx.x;
  if (p >= world) return;
  const unsigned int* f = my_flags + channel * world + p;
  long long t0 = clock64();
  while (true) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if ((int)(v - epoch) >= 0) break;
    if (clock64() - t0 > 40000000000LL) { printf("clipk: peer_wait timeout: channel %d peer %d flag %u epoch %u\n", channel, p, v, epoch); __trap(); }
    __nanosleep(200);
  }
}

// out[r, :] (+)= sum_p src[p][(rank * rows + r), :]
__global__ void __launch_bounds__(256) peer_reduce_rows_kernel(float* const* __restrict__ src, int world, int rank, float* __restrict__ out, int rows,
                                                               int d, int accumulate) {
  const long long n4 = (long long)rows * d / 4;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
    float4 a = accumulate ? *reinterpret_cast<const float4*>(out + t * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < world; ++p) {
      const float4 v = *reinterpret_cast<const float4*>(src[p] + (long long)rank * rows * d + t * 4);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    *reinterpret_cast<float4*>(out + t * 4) = a;
  }
}

}  // namespace clipk

using namespace clipk;

extern "C" int clipk_peer_alloc(void** dev_ptr, size_t bytes) {
  if (!dev_ptr || !bytes) { set_error("peer_alloc: bad arguments"); return CLIPK_ERR_ARG; }
  CLIPK_CUDA(cudaMalloc(dev_ptr, bytes));
  CLIPK_CUDA(cudaMemset(*dev_ptr, 0, bytes));
  return 0;
}
extern "C" int clipk_peer_free(void* dev_ptr) { CLIPK_CUDA(cudaFree(dev_ptr)); return 0; }
extern "C" int clipk_peer_export(const void* dev_ptr, unsigned char* handle64) {
  cudaIpcMemHandle_t h;
  CLIPK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return 0;
}
extern "C" int clipk_peer_open(const unsigned char* handle64, void** dev_ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CLIPK_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int clipk_peer_close(void* dev_ptr) { CLIPK_CUDA(cudaIpcCloseMemHandle(dev_ptr)); return 0; }

extern "C" int clipk_l2norm_allgather(const float* x, float* y_local, float* norm, float* const* gallery_ptrs, int world, int rank, int rows, int d,
                                      cudaStream_t stream) {
  if (d % 4 || world < 1 || rank < 0 || rank >= world) { set_error("l2norm_allgather: d %% 4 != 0 or bad world/rank"); return CLIPK_ERR_ARG; }
  if (rows <= 0) return 0;
  l2norm_allgather_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(x, y_local, norm, gallery_ptrs, world, rank, rows, d);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int clipk_peer_signal(unsigned int* const* flag_ptrs, int world, int rank, int channel, unsigned int epoch, cudaStream_t stream) {
  if (world < 1 || world > 32) { set_error("peer_signal: world %d", world); return CLIPK_ERR_ARG; }
  peer_signal_kernel<<<1, 32, 0, stream>>>(flag_ptrs, world, rank, channel, epoch);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int clipk_peer_wait(const unsigned int* my_flags, int world, int channel, unsigned int epoch, cudaStream_t stream) {
  if (world < 1 || world > 32) { set_error("peer_wait: world %d", world); return CLIPK_ERR_ARG; }
  peer_wait_kernel<<<1, 32, 0, stream>>>(my_flags, world, channel, epoch);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int clipk_peer_reduce_rows(float* const* src_ptrs, int world, int rank, float* out, int rows, int d, int accumulate, cudaStream_t stream) {
  if (((long long)rows * d) % 4) { set_error("peer_reduce_rows: rows * d %% 4 != 0"); return CLIPK_ERR_ARG; }
  if (rows <= 0) return 0;
  long long n4 = (long long)rows * d / 4;
  int grid = (int)((n4 + 255) / 256); if (grid > sm_count() * 4) grid = sm_count() * 4;
  peer_reduce_rows_kernel<<<grid, 256, 0, stream>>>(src_ptrs, world, rank, out, rows, d, accumulate);
  note_launch();
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}
