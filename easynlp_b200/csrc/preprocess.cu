// Image side of the input pipeline on the GPU: decoded RGB frames -> the normalised [3, S, S] tensors the image tower takes.
//
// Replaces, for a batch of decoded 8-bit RGB images, the per-image host chain of CLIPDataset.convert_single_row_to_example /
// CLIPPredictor.preprocess (easynlp/appzoo/clip/data.py:29-135,263-272; predictor.py:100-110):
//     _resize(image, 224, Image.BICUBIC)  ->  _center_crop(224)  ->  /255  ->  (x - mean) / std
// `_resize` is Pillow's ImagingResample (src/libImaging/Resample.c): a separable two-pass convolution whose intermediate image is 8-bit.
// The kernels reproduce it BIT-EXACTLY (integer work: the parity bar is equality):
//   * weights in double precision with the library's operation order (explicit round-to-nearest intrinsics, no FMA contraction):
//     centre = (i + 0.5) * scale, support = 2 * max(scale, 1), taps = round(centre -/+ support) clipped to the image, Keys' cubic
//     (a = -0.5), normalised to sum 1, rounded half away from zero to 22 fractional bits;
//   * each pass accumulates in int32 from 2^21, shifts right by 22 and clamps to [0, 255];
//   * only what the centre crop keeps is computed: S output columns per source row in the horizontal pass, S x S outputs in the vertical.
// Three launches per batch: tap tables -> horizontal pass (source rows -> uint8 [h, S, 3] scratch) -> vertical pass fused with the
// float normalisation ((v / 255 - mean) / std in IEEE single precision, as numpy evaluates it).  HBM-bound byte work: algorithmic traffic
// per image = w*h*3 (source) + 2 * rows*S*3 (scratch write + read) + S*S*3*4 (output) bytes.
#include "common.cuh"
#include "../../include/clipk.h"

namespace clipk {
namespace {

constexpr int PREC = 22;                    // 32 - 8 - 2 fractional bits (Resample.c: PRECISION_BITS)

struct ImageDesc { long long src; int w, h; long long tmp; };      // = clipk_image_desc (byte offsets into the pixel blob / the scratch area)

struct Geometry { int new_w, new_h, left, top; };

__device__ __forceinline__ Geometry geometry(int w, int h, int S) {
  // data.py:54-74 (_resize with an int size) and :44-52 (_center_crop)
  Geometry g;
  const int shrt = w <= h ? w : h, lng = w <= h ? h : w;
  int new_long = lng;
  if (shrt != S) new_long = __double2int_rz(__ddiv_rn((double)((long long)S * lng), (double)shrt));
  const int new_short = shrt != S ? S : shrt;
  g.new_w = w <= h ? new_short : new_long;
  g.new_h = w <= h ? new_long : new_short;
  g.left = __double2int_rz(__dmul_rn((double)(g.new_w - S + 1), 0.5));
  g.top = __double2int_rz(__dmul_rn((double)(g.new_h - S + 1), 0.5));
  return g;
}

__device__ __forceinline__ double keys_cubic(double x) {
  // ((a + 2) x - (a + 3)) x x + 1  resp.  (((x - 5) x + 8) x - 4) a   with a = -0.5, evaluated left to right without contraction
  if (x < 0.0) x = -x;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dadd_rn(__dmul_rn(1.5, x), -2.5), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(x, -5.0), x), 8.0), x), -4.0), -0.5);
  return 0.0;
}

// tap tables of the S kept outputs of one axis of one image: bounds[(img*2+axis)*S + i] = {first tap, taps}, k[...][kmax] fixed point
__global__ void __launch_bounds__(256) preprocess_taps_kernel(const ImageDesc* __restrict__ desc, int2* __restrict__ bounds, int* __restrict__ taps,
                                                              int S, int kmax, int* __restrict__ status) {
  const int img = blockIdx.x, axis = blockIdx.y, i = threadIdx.x;      // axis 0 = horizontal
  if (i >= S) return;
  const ImageDesc d = desc[img];
  const Geometry g = geometry(d.w, d.h, S);
  const int in_size = axis == 0 ? d.w : d.h, out_size = axis == 0 ? g.new_w : g.new_h, first = axis == 0 ? g.left : g.top;
  const long long row = ((long long)img * 2 + axis) * S + i;
  int* k = taps + row * kmax;
  const int xx = first + i;
  if (in_size == out_size) {                 // Pillow skips the pass: identity tap
    bounds[row] = make_int2(xx, 1);
    k[0] = 1 << PREC;
    return;
  }
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, filterscale);
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
  int xmin = __double2int_rz(__dadd_rn(__dadd_rn(center, -support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  if (xmax > kmax) { atomicExch(status, 1); xmax = kmax; }            // the caller's kmax was too small: reported, never silent
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x)
    ww = __dadd_rn(ww, keys_cubic(__dmul_rn(__dadd_rn(__dadd_rn((double)(x + xmin), -center), 0.5), ss)));
  for (int x = 0; x < xmax; ++x) {
    double w = keys_cubic(__dmul_rn(__dadd_rn(__dadd_rn((double)(x + xmin), -center), 0.5), ss));
    if (ww != 0.0) w = __ddiv_rn(w, ww);
    const double f = __dmul_rn(w, (double)(1 << PREC));
    k[x] = __double2int_rz(w < 0.0 ? __dadd_rn(-0.5, f) : __dadd_rn(0.5, f));
  }
  bounds[row] = make_int2(xmin, xmax);
}

__device__ __forceinline__ int clip8(int acc) {
  const int v = acc >> PREC;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: one block per (source row, image); thread x produces the 3 channels of kept column x.  Rows the vertical pass never
// reads are skipped.
__global__ void __launch_bounds__(256) preprocess_horizontal_kernel(const unsigned char* __restrict__ pixels, const ImageDesc* __restrict__ desc,
                                                                    const int2* __restrict__ bounds, const int* __restrict__ taps,
                                                                    unsigned char* __restrict__ scratch, int S, int kmax) {
  const int img = blockIdx.y, r = blockIdx.x, x = threadIdx.x;
  const ImageDesc d = desc[img];
  if (r >= d.h) return;
  const int2 vb0 = bounds[((long long)img * 2 + 1) * S], vb1 = bounds[((long long)img * 2 + 1) * S + S - 1];
  if (r < vb0.x || r >= vb1.x + vb1.y || x >= S) return;
  const long long row = ((long long)img * 2 + 0) * S + x;
  const int2 b = bounds[row];
  const int* __restrict__ k = taps + row * kmax;
  const unsigned char* __restrict__ src = pixels + d.src + ((long long)r * d.w + b.x) * 3;
  int a0 = 1 << (PREC - 1), a1 = a0, a2 = a0;
  for (int t = 0; t < b.y; ++t) {
    const int kw = k[t];
    a0 += (int)src[t * 3 + 0] * kw; a1 += (int)src[t * 3 + 1] * kw; a2 += (int)src[t * 3 + 2] * kw;
  }
  unsigned char* dst = scratch + d.tmp + ((long long)r * S + x) * 3;
  dst[0] = (unsigned char)clip8(a0); dst[1] = (unsigned char)clip8(a1); dst[2] = (unsigned char)clip8(a2);
}

struct Norm { float mean[3], std[3]; };

// vertical pass + normalisation: one block per (kept row, image); out[img][c][y][x] = ((v / 255) - mean[c]) / std[c]
__global__ void __launch_bounds__(256) preprocess_vertical_kernel(const ImageDesc* __restrict__ desc, const int2* __restrict__ bounds,
                                                                  const int* __restrict__ taps, const unsigned char* __restrict__ scratch,
                                                                  float* __restrict__ out, int S, int kmax, Norm nm) {
  const int img = blockIdx.y, y = blockIdx.x, x = threadIdx.x;
  if (x >= S) return;
  const ImageDesc d = desc[img];
  const long long row = ((long long)img * 2 + 1) * S + y;
  const int2 b = bounds[row];
  const int* __restrict__ k = taps + row * kmax;
  const unsigned char* __restrict__ src = scratch + d.tmp + ((long long)b.x * S + x) * 3;
  int a0 = 1 << (PREC - 1), a1 = a0, a2 = a0;
  for (int t = 0; t < b.y; ++t) {
    const int kw = k[t];
    const unsigned char* p = src + (long long)t * S * 3;
    a0 += (int)p[0] * kw; a1 += (int)p[1] * kw; a2 += (int)p[2] * kw;
  }
  const int v[3] = {clip8(a0), clip8(a1), clip8(a2)};
  float* o = out + (long long)img * 3 * S * S + (long long)y * S + x;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    o[(long long)c * S * S] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v[c], 255.0f), nm.mean[c]), nm.std[c]);
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace
}  // namespace clipk

using namespace clipk;

// layout of the workspace: [bounds: n*2*S int2][taps: n*2*S*kmax int][status int (+pad)][scratch: caller-assigned desc.tmp offsets]
extern "C" size_t clipk_preprocess_workspace(int n, int size, int kmax, long long scratch_bytes) {
  if (n <= 0 || size <= 0 || kmax <= 0 || scratch_bytes < 0) return 0;
  size_t b = align_up((size_t)n * 2 * size * sizeof(int2), 256);
  b += align_up((size_t)n * 2 * size * kmax * sizeof(int), 256);
  b += 256;
  return b + (size_t)scratch_bytes;
}

extern "C" int clipk_preprocess_kmax(int w, int h, int size) {
  // taps per output of the longer resampling filter of one image: (int)ceil(2 * max(scale, 1)) * 2 + 1 (Resample.c: ksize)
  if (w <= 0 || h <= 0 || size <= 0) return 0;
  const int shrt = w <= h ? w : h, lng = w <= h ? h : w;
  const int new_long = shrt == size ? lng : (int)((double)((long long)size * lng) / (double)shrt);
  const int new_short = shrt == size ? shrt : size;
  int kmax = 1;
  const int ins[2] = {shrt, lng}, outs[2] = {new_short, new_long};
  for (int a = 0; a < 2; ++a) {
    if (ins[a] == outs[a] || outs[a] <= 0) continue;
    double scale = (double)ins[a] / outs[a];
    if (scale < 1.0) scale = 1.0;
    const int ks = (int)ceil(2.0 * scale) * 2 + 1;
    if (ks > kmax) kmax = ks;
  }
  return kmax;
}

extern "C" int clipk_preprocess_images(const unsigned char* pixels, const clipk_image_desc* desc, int n, int max_h, int size, int kmax,
                                       const float* mean3, const float* std3, float* out, void* workspace, size_t workspace_bytes,
                                       long long scratch_bytes, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (!pixels || !desc || !out || !workspace || !mean3 || !std3) { set_error("clipk_preprocess_images: null argument"); return CLIPK_ERR_ARG; }
  if (size <= 0 || size > 256 || kmax <= 0 || max_h <= 0) { set_error("clipk_preprocess_images: size must be in 1..256 (got %d), kmax %d, max_h %d", size, kmax, max_h); return CLIPK_ERR_ARG; }
  if (workspace_bytes < clipk_preprocess_workspace(n, size, kmax, scratch_bytes)) { set_error("clipk_preprocess_images: workspace too small"); return CLIPK_ERR_ARG; }
  static_assert(sizeof(ImageDesc) == sizeof(clipk_image_desc), "descriptor layout");
  char* ws = static_cast<char*>(workspace);
  int2* bounds = reinterpret_cast<int2*>(ws);
  ws += align_up((size_t)n * 2 * size * sizeof(int2), 256);
  int* taps = reinterpret_cast<int*>(ws);
  ws += align_up((size_t)n * 2 * size * kmax * sizeof(int), 256);
  int* status = reinterpret_cast<int*>(ws);
  ws += 256;
  unsigned char* scratch = reinterpret_cast<unsigned char*>(ws);
  Norm nm;
  for (int c = 0; c < 3; ++c) { nm.mean[c] = mean3[c]; nm.std[c] = std3[c]; }
  const ImageDesc* d = reinterpret_cast<const ImageDesc*>(desc);
  CLIPK_CUDA(cudaMemsetAsync(status, 0, sizeof(int), stream));
  preprocess_taps_kernel<<<dim3(n, 2), 256, 0, stream>>>(d, bounds, taps, size, kmax, status);
  preprocess_horizontal_kernel<<<dim3(max_h, n), 256, 0, stream>>>(pixels, d, bounds, taps, scratch, size, kmax);
  preprocess_vertical_kernel<<<dim3(size, n), 256, 0, stream>>>(d, bounds, taps, scratch, out, size, kmax, nm);
  note_launch(3);
  CLIPK_CUDA(cudaGetLastError());
  return 0;
}

// 1 when a tap table of the last call on this workspace overflowed `kmax` (results invalid); reads the flag back (synchronises the stream)
extern "C" int clipk_preprocess_status(const void* workspace, int n, int size, int kmax, cudaStream_t stream) {
  const char* ws = static_cast<const char*>(workspace);
  ws += align_up((size_t)n * 2 * size * sizeof(int2), 256) + align_up((size_t)n * 2 * size * kmax * sizeof(int), 256);
  int v = 0;
  CLIPK_CUDA(cudaMemcpyAsync(&v, ws, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CLIPK_CUDA(cudaStreamSynchronize(stream));
  return v;
}
