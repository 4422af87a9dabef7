/* b200mm C-ABI — the drop-in boundary below the Python host.
 *
 * Every entry point takes raw DEVICE pointers, int64 sizes, scalar hyper-parameters and a CUDA stream
 * (passed as void* so the header needs no CUDA include) and returns 0 on success or a negative
 * B200MM_ERR_* code; b200mm_last_error() returns the thread-local message of the last failure.
 * Ownership: the caller (torch) owns every buffer; kernels never allocate, free or synchronise.
 * All calls are stateless, re-entrant and stream-ordered, so they are safe from autograd worker threads.
 * There is NO CPU fallback: host pointers are rejected by the CUDA runtime, not silently handled.
 *
 * Each function cites the reference code (relative to the AntMMF tree) whose arithmetic it replaces.
 * bf16 = device buffers of __nv_bfloat16; f32 = float.
 */
#ifndef B200MM_H_
#define B200MM_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MM_OK 0
#define B200MM_ERR_SHAPE (-1)  /* unsupported / inconsistent dimensions */
#define B200MM_ERR_ALIGN (-2)  /* pointer or pitch not aligned as required */
#define B200MM_ERR_ARCH (-3)   /* device is not sm_100 */
#define B200MM_ERR_LAUNCH (-4) /* CUDA launch / driver failure (message holds cudaGetErrorString) */

/* activation selectors (GEMM epilogues, elementwise kernels) */
#define B200MM_ACT_NONE 0
#define B200MM_ACT_QUICKGELU 1 /* x*sigmoid(1.702x): antmmf/modules/vision/backbone/clip/model.py:222-224 */
#define B200MM_ACT_GELU_ERF 2  /* x*0.5*(1+erf(x/sqrt2)): antmmf/modules/vision/backbone/clip/modeling_bert.py:31-37 */

const char* b200mm_last_error(void);
int b200mm_version(void);
/* 0 if the current device is compute capability 10.x, else B200MM_ERR_ARCH */
int b200mm_check_device(void);

/* ---------------------------------------------------------------------------------------------
 * GEMM (tcgen05 + TMA):  D[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )
 *
 * Replaces every nn.Linear / F.linear / torch.matmul on the path:
 *   ViT  in_proj / out_proj / c_fc / c_proj   antmmf/modules/vision/backbone/clip/model.py:231,236-238,251
 *   BERT query/key/value/dense                antmmf/modules/vision/backbone/clip/modeling_bert.py:120-122,135-137,182,221,234
 *   projections  x @ proj                     clip/model.py:332-333, clip/cn_model.py:210
 * and their autograd counterparts (dgrad: A=dY, B=W^T-as-MN-major; wgrad: both operands MN-major).
 *
 * Operand storage (bf16, 16-byte aligned base, pitches multiples of 8 elements):
 *   a_mn == 0: A is row-major [M, K] with row pitch lda      (K-major)
 *   a_mn == 1: A is row-major [K, M] with row pitch lda      (MN-major: the reduction index is the slow one)
 *   b_mn == 0: B is row-major [N, K] with row pitch ldb      (nn.Linear.weight layout)
 *   b_mn == 1: B is row-major [K, N] with row pitch ldb
 * Epilogue, per element, in this order (fp32):
 *   v = alpha*acc; v += bias[n] (bf16, optional); if aux_out: aux_out[m,n] = v (bf16, pre-activation);
 *   v = act(v); if dact_in: v *= act'(dact_in[m,n]) (act is then applied as derivative only, not to v);
 *   v += residual[m,n] (bf16, optional); D[m,n] = v (bf16 if d_f32 == 0 else f32)
 * splits > 1 partitions K over `splits` CTAs per tile; partial sums go to `workspace`
 * (f32, >= b200mm_gemm_workspace_bytes) and a second kernel reduces them and applies the epilogue.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* A;
  int64_t lda;
  int32_t a_mn;
  const void* B;
  int64_t ldb;
  int32_t b_mn;
  void* D;
  int64_t ldd;
  int32_t d_f32;
  int64_t M, N, K;
  float alpha;
  const void* bias;     /* bf16 [N] or NULL */
  int32_t act;          /* B200MM_ACT_* */
  void* aux_out;        /* bf16 [M,N] pitch ldd, pre-activation copy, or NULL */
  const void* dact_in;  /* bf16 [M,N] pitch ld_dact: multiply by act'(dact_in) instead of applying act */
  int64_t ld_dact;
  const void* residual; /* bf16 [M,N] pitch ldr or NULL */
  int64_t ldr;
  int32_t splits;       /* >= 1 */
  void* workspace;      /* f32, needed iff splits > 1 */
  int64_t workspace_bytes;
} b200mm_gemm_args;

int64_t b200mm_gemm_workspace_bytes(int64_t M, int64_t N, int32_t splits);
int b200mm_gemm_bf16(const b200mm_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200MM_H_ */
