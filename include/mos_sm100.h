/* mos_sm100.h — C ABI of libmos_sm100.so, the B200 (sm_100a) ED-LoRA diffusion hot path.
 *
 * Conventions: every entry point returns int (0 = ok, negative = MOS_E*); the message of the
 * last failure on the calling thread is available from mos_last_error(). All pointers are raw device pointers
 * owned by the caller (PyTorch allocates everything); the library never allocates per call, never retains a
 * pointer after return and never synchronises: work is enqueued on the cudaStream_t passed as `stream`.
 * Activations are bf16, NHWC / token-major; weights are pre-packed bf16 K-major (see DESIGN.md "data layout").
 *
 * Each entry point cites the reference call site it replaces (paths relative to TencentARC/Mix-of-Show).
 */
#ifndef MOS_SM100_H
#define MOS_SM100_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOS_OK 0
#define MOS_EINVAL (-1)   /* bad argument (shape / alignment / null pointer) */
#define MOS_ECUDA (-2)    /* CUDA runtime / driver error */
#define MOS_EUNSUPPORTED (-3)

int mos_version(void);
const char* mos_last_error(void);

/* ------------------------------------------------------------------------------------------------------------
 * Fused GEMM (+ implicit-GEMM 3x3 convolution) with LoRA / bias / temb / GEGLU / residual epilogue.
 *   out[m, n] = epi( sum_k A[m, k] * W[n, k]  +  sum_r (sum_k A[m,k] * lora_down[r,k]) * lora_up[n, r] )
 * Replaces, in one tcgen05 kernel:
 *   - LoRALinearLayer.forward                     mixofshow/models/edlora.py:244-246
 *   - attn.to_q / to_k / to_v / to_out[0]         mixofshow/models/edlora.py:143-145,161 (and :69-71,88)
 *   - region to_k / to_v                          mixofshow/pipelines/pipeline_regionally_t2iadapter.py:122-126
 *   - diffusers ResnetBlock2D conv1/conv2, Transformer2DModel proj_in/proj_out, FeedForward GEGLU
 *     (called through unet(...) at mixofshow/pipelines/pipeline_edlora.py:277)
 * ---------------------------------------------------------------------------------------------------------- */
enum { MOS_OUT_BF16 = 0, MOS_OUT_HEADS = 1, MOS_OUT_F32 = 2 };
enum { MOS_SEG_ROWS = 0 /* [b,h,row,dpad] (Q, K) */, MOS_SEG_TRANSPOSED = 1 /* [b,h,d,row] (V^T) */ };

typedef struct mos_gemm_args {
  const void* A;          /* bf16 [M, lda]; conv: NHWC activation [B, H, Wd, C] */
  const void* W;          /* bf16 [N, Kw], K contiguous; conv: [N, 9*C] with k = (kh*3+kw)*C + c */
  int64_t M, N, K;        /* K = C for conv (reduction per tap) */
  int64_t lda;            /* row pitch of A in elements (conv: pixel pitch, >= C) */
  int32_t conv;           /* 0 = plain GEMM, 1 = 3x3 / stride 1 / pad 1 convolution */
  int32_t B, H, Wd, C;    /* conv geometry */
  int32_t splits;         /* split-K factor, >= 1; > 1 requires `partial` and forbids lora / geglu / heads */
  int32_t stages;         /* smem pipeline depth, 0 = default */
  float* partial;         /* fp32 workspace [splits, M, N] */
  const float* bias;      /* [N] or NULL */
  const float* bias_batch;/* [nbatch, N] or NULL; row m uses batch m / rows_per_batch (resnet temb add) */
  int64_t rows_per_batch;
  int64_t bias_batch_ld;  /* row pitch of bias_batch in elements (0 = N) */
  const void* residual;   /* bf16 [M, ldr] or NULL, added last */
  int64_t ldr;
  int32_t geglu;          /* 1: tile columns are [80 x a | 80 x gate]; writes a*gelu(gate), N_out = N/2 */
  const void* lora_down;  /* bf16 [16, K], rows >= rank zero; NULL = no LoRA */
  const float* lora_up;   /* fp32 [N, 4], pre-multiplied by alpha */
  int64_t lora_seg;       /* columns per LoRA segment (N, or C for fused q|k|v: segment s uses down rows 4s..4s+3) */
  int32_t out_mode;       /* MOS_OUT_* */
  void* out;              /* bf16 / fp32 [M, ldc] for MOS_OUT_BF16 / MOS_OUT_F32 */
  int64_t ldc;
  /* MOS_OUT_HEADS: columns are `nseg` segments of seg_len = heads*head_dim; segment s goes to seg_ptr[s] */
  void* seg_ptr[3];
  int32_t seg_kind[3];
  int64_t seg_rows_pad[3];   /* padded row count of the destination (tokens or keys) */
  int32_t heads, head_dim, dpad, dv_pad;
  int64_t tokens_per_batch;
} mos_gemm_args;

int mos_gemm_bf16(const mos_gemm_args* args, void* stream);

/* Sum split-K partials and apply bias / bias_batch / residual -> bf16 [M, ldc]. */
int mos_splitk_finalize(const float* partial, int32_t splits, int64_t M, int64_t N, const float* bias,
                        const float* bias_batch, int64_t rows_per_batch, const void* residual, int64_t ldr,
                        void* out, int64_t ldc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOS_SM100_H */
