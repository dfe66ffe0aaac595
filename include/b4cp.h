/* libb4cp — C ABI of the B200-native clickstream-transformer hot path.
 *
 * The reference (MiladShahidi/BERT4ClickPath) has no FFI layer: its hot path is a chain of
 * TensorFlow 2.3.1 ops called from Python classes.  Each export below replaces the TF op chain
 * at the cited reference call site (paths relative to the reference repository) with one or a
 * few hand-written sm_100a kernels.  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name says `h_` (host);
 *   - `stream` is a cudaStream_t passed as void*; calls are stream-ordered, never synchronise,
 *     never allocate (callers pass workspaces sized by the *_workspace_bytes queries);
 *   - return value: 0 = ok, <0 = bad argument, >0 = cudaError_t; b4cp_last_error() gives the
 *     thread-local message;
 *   - matrices are row-major; Dense kernels use the Keras (in, out) layout;
 *   - "bf16" buffers are passed as void* (raw __nv_bfloat16 bits).
 */
#ifndef B4CP_H_
#define B4CP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B4CP_VERSION 1
#define B4CP_MAX_FEATURES 8

const char* b4cp_last_error(void);
int b4cp_version(void);
/* fails (<0) unless the current device is compute capability 10.x */
int b4cp_device_check(void);

/* ------------------------------------------------------------------ dense layers (tcgen05)
 * Replaces tf.keras.layers.Dense MatMul+BiasAdd(+ReLU) and its autodiff transposes:
 *   clickstream_transformer/transformer.py:112-116,139-141,158 (MHA projections)
 *   clickstream_transformer/transformer.py:163-167 (feed-forward)
 *   clickstream_transformer/head.py:10-11,16-19,35-45,56-62 (head MLPs)
 *
 *   C[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) ),  bf16 operands, fp32 accumulation.
 * a_mn = 0: A is stored [M][K] (K contiguous);  a_mn = 1: A is stored [K][M] (M contiguous).
 * b_mn = 0: B is stored [N][K];                  b_mn = 1: B is stored [K][N] (Keras kernel).
 * Leading dimensions in elements, multiples of 8.  splits > 1 writes `splits` raw fp32 partial
 * products at out_f32 + z*split_stride (reduce them with b4cp_reduce_splits).
 */
typedef struct {
  float alpha;         /* scale on the accumulator (1.0f for a plain product) */
  const float* bias;   /* [N] added to every row, or NULL */
  int relu;            /* max(x, 0) after bias */
  const void* gate;    /* bf16 [M][ld_gate]: result zeroed where gate <= 0 (ReLU backward), or NULL */
  long ld_gate;
  const float* addend; /* fp32 [M][ld_addend] added last (residual / gradient accumulate), or NULL */
  long ld_addend;
  float* out_f32;      /* fp32 [M][ld_f32] or NULL */
  long ld_f32;
  long split_stride;   /* elements between split-K partials in out_f32 */
  void* out_bf16;      /* bf16 [M][ld_bf16] or NULL */
  long ld_bf16;
} b4cp_gemm_epilogue;

int b4cp_gemm_bf16(const void* A, int a_mn, long lda, const void* B, int b_mn, long ldb, int M,
                   int N, int K, int splits, const b4cp_gemm_epilogue* ep, void* stream);
/* number of K splits that fills the GPU for an M x N output */
int b4cp_gemm_splits_for(int M, int N, int K);

#ifdef __cplusplus
}
#endif
#endif /* B4CP_H_ */
