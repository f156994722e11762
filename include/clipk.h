/* clipk -- C ABI of the B200 (sm_100a) kernels behind EasyNLP's CLIP contrastive path.
 *
 * The reference (alibaba/EasyNLP) is pure Python/PyTorch and has NO FFI on this path:
 * every entry point below replaces a sequence of ATen library calls made by the cited reference lines.
 * The binding a maintainer adds on the reference side is a ctypes stub; the host-side
 * mirror of the reference's plugin interface lives in easynlp_b200/ (Python, like the reference).
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless named host_*.
 *   - the caller owns every buffer (workspace included); no hidden allocation, no global state except
 *     a per-process error string and cached device attributes.
 *   - every function is asynchronous on `stream` and returns 0 on success or a negative CLIPK_ERR_*;
 *     clipk_last_error() then describes the failure.
 *   - bf16 = IEEE bfloat16 storage; "ld*" = leading dimension in ELEMENTS.
 */
#ifndef CLIPK_H_
#define CLIPK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_API_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define CLIPK_ERR_ARG (-1)
#define CLIPK_ERR_CUDA (-2)
#define CLIPK_ERR_UNSUPPORTED (-3)

#define CLIPK_BF16 0
#define CLIPK_F32 1

/* -------------------------------------------------------------------------------------------- library */
const char* clipk_last_error(void);
int clipk_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t clipk_launch_count(void);

/* -------------------------------------------------------------------------------------------- GEMM
 * Replaces nn.Linear / nn.Conv2d(patch) / MultiheadAttention in/out projections and their autograd
 * (reference: modeling_chineseclip.py:188-204,224-251; modeling_bert.py:145-147,264-268,329-346).   */
#define CLIPK_EPI_LINEAR 0       /* out = alpha*acc + bias + residual                    (out bf16|f32, optional bf16 out2) */
#define CLIPK_EPI_QUICK_GELU 1   /* out = z = acc + bias (bf16); out2 = z*sigmoid(1.702 z) (modeling_chineseclip.py:179-181) */
#define CLIPK_EPI_ERF_GELU 2     /* out = z (bf16); out2 = gelu_erf(z)                    (modelzoo/activations.py:45-48)   */
#define CLIPK_EPI_DQUICK_GELU 3  /* out = acc * d/dz quick_gelu(aux)                      (backward of mode 1)              */
#define CLIPK_EPI_DERF_GELU 4    /* out = acc * d/dz gelu_erf(aux)                        (backward of mode 2)              */
#define CLIPK_EPI_ATOMIC_ADD 5   /* out(f32) += acc  via red.add -- split-K weight gradients                                */

typedef struct {
  int mode;              /* CLIPK_EPI_* */
  int out_dtype;         /* CLIPK_BF16 | CLIPK_F32 */
  void* out;             /* [M, ldo] */
  int ldo;
  void* out2;            /* optional bf16 [M, ldo2] */
  int ldo2;
  const float* bias;     /* optional [N] */
  const float* residual; /* optional f32 [M, ldr] */
  int ldr;
  const void* aux;       /* bf16 [M, ldaux] (pre-activation for the dGELU modes) */
  int ldaux;
  float alpha;           /* 0 is read as 1 */
} clipk_epilogue_t;

/* D[M,N] = epilogue(op(A) x op(B)), bf16 operands, fp32 accumulation in TMEM (tcgen05).
 *   a_mn_major = 0: A is [M,K] row-major;  1: A is [K,M] row-major (A^T is the logical operand)
 *   b_mn_major = 0: B is [N,K] row-major (nn.Linear weight);  1: B is [K,N] row-major
 *   splits > 1: split-K, requires CLIPK_EPI_ATOMIC_ADD into a pre-zeroed/accumulating fp32 output.
 * lda, ldb, N multiples of 8; base pointers 16-byte aligned.                                          */
int clipk_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, int M, int N, int K,
                    const clipk_epilogue_t* epi, int splits, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIPK_H_ */
