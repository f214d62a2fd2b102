/*
 * vrr.h - C ABI of libvrr_b200.so: the B200 (sm_100a) ViT attention hot path of
 * zhengyk19/vit-rpe-rope.
 *
 * The reference is pure Python/PyTorch and has NO FFI / plugin boundary of its own
 * (SURVEY.md section 8, row B1): its boundary is the Python module API of models/.  This header
 * is therefore the boundary a maintainer would bind from that API (INTEGRATION.md shows the
 * ctypes stub).  Each entry point names the reference code (file:line under /root/reference) it
 * replaces.
 *
 * Conventions (all entry points)
 *   - plain pointers and sizes only: no torch / pybind / C++ types cross this boundary;
 *   - every data pointer is a DEVICE pointer on the current CUDA device (sm_100); the library
 *     never allocates, never synchronises, never throws; all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - return value: VRR_OK (0) or a negative vrr_status; vrr_last_error() gives the reason;
 *   - tensors are dense row-major with the shapes documented per argument; `dtype` selects the
 *     element type of activations/weights (VRR_F32 or VRR_BF16).  Position tables (cos/sin, bias
 *     table, polynomial coefficients), softmax statistics and all parameter-gradient outputs that
 *     are reductions (d_table, d_coef, d_cos, d_sin) are ALWAYS fp32;
 *   - there is no CPU path: on a machine without an sm_100 device every compute entry point
 *     returns VRR_ERR_NO_DEVICE.
 *
 * Token / head geometry: B images, N tokens per image (token 0 is the cls token, tokens 1..N-1 are
 * the g*g patches in raster order), E = H * Dh channels, H heads of Dh channels.  Dh must be one of
 * 16, 32, 64 (the tcgen05 bf16 kernels need Dh == 64).
 *
 * "qkv planes": one buffer [3][B][H][N][Dh] holding q (rotated when RoPE is on), k (likewise), v.
 */
#ifndef VRR_H_
#define VRR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRR_ABI_VERSION 1

typedef enum {
  VRR_OK = 0,
  VRR_ERR_INVALID_ARG = -1,   /* null pointer, bad enum, inconsistent sizes            */
  VRR_ERR_UNSUPPORTED = -2,   /* shape/dtype outside what the kernels implement         */
  VRR_ERR_NO_DEVICE = -3,     /* no CUDA device, or the current device is not sm_100    */
  VRR_ERR_CUDA = -4,          /* a CUDA runtime/driver call failed (see vrr_last_error) */
  VRR_ERR_WORKSPACE = -5      /* workspace too small (see vrr_*_workspace_bytes)        */
} vrr_status;

typedef enum { VRR_F32 = 0, VRR_BF16 = 1 } vrr_dtype;

/* RoPE variants of models/vit.py:51-68.  AXIAL: cos/sin are [N-1][Dh/2] shared by all heads
 * (positional_encoding.py:216-245).  MIXED: cos/sin are [H][N-1][Dh/2]
 * (positional_encoding.py:313-351).  Token 0 (cls) is never rotated (vit.py:56-57). */
typedef enum { VRR_ROPE_NONE = 0, VRR_ROPE_AXIAL = 1, VRR_ROPE_MIXED = 2 } vrr_rope_mode;

/* Additive logit bias of models/vit.py:73-81.
 * TABLE: relative-position table [H][2N-1]; bias[h][i][j] = table[h][i - j + N - 1]
 *        (positional_encoding.py:58-75,93).
 * POLY : polynomial in the L1 patch distance, coef [Hc][degree+1] with Hc = 1 (shared) or H;
 *        bias[h][i][j] = sum_k coef[k] * d^k for i,j >= 1, 0 on the cls row/column, with
 *        d = |(i-1)%g - (j-1)%g| + |(i-1)/g - (j-1)/g| (positional_encoding.py:127-171). */
typedef enum { VRR_BIAS_NONE = 0, VRR_BIAS_TABLE = 1, VRR_BIAS_POLY = 2 } vrr_bias_mode;

/* Kernel family selection for the bf16 attention / QKV kernels.  AUTO picks the tcgen05 (tensor
 * core, TMEM) kernels whenever the shape allows and the SIMT kernels otherwise; the other two
 * force one family (tests use them to cross-check).  fp32 always runs the SIMT FFMA kernels:
 * tcgen05 has no fp32 MMA and the fp32 parity bar is 1e-5. */
typedef enum { VRR_IMPL_AUTO = 0, VRR_IMPL_SIMT = 1, VRR_IMPL_TCGEN05 = 2 } vrr_impl;

/* Epilogues of vrr_gemm_ex (Linear layers, models/vit.py:91,118,285). */
typedef enum {
  VRR_EPI_NONE = 0,
  VRR_EPI_BIAS = 1,            /* c = acc + bias                                                              */
  VRR_EPI_BIAS_GELU = 2,       /* c = h = acc + bias, c2 = gelu(h)            (exact erf GELU, timm Mlp)        */
  VRR_EPI_BIAS_GELU_GRAD = 3,  /* c = gelu(h), c2 = d gelu / dh (h)           (what the Mlp backward needs)     */
  VRR_EPI_MUL = 4,             /* c = acc * c2[m][n]; c2 is an INPUT [M][N]   (d_act * gelu'(h) in fc2's dX)    */
  VRR_EPI_BIAS_GELU_ACT = 5    /* c = gelu(acc + bias), c2 unused             (the Mlp under torch.no_grad)     */
} vrr_epilogue;

typedef struct {
  int32_t mode;        /* vrr_bias_mode                                                  */
  int32_t heads;       /* rows of `param`: H for TABLE; 1 or H for POLY                  */
  int32_t len;         /* columns of `param`: 2N-1 for TABLE; degree+1 for POLY          */
  int32_t grid;        /* POLY: patch grid side g (N-1 == g*g); ignored otherwise        */
  const float* param;  /* fp32 [heads][len]; NULL when mode == NONE                      */
} vrr_bias_desc;

/* ---- library ------------------------------------------------------------------------------- */
int vrr_abi_version(void);
const char* vrr_last_error(void);          /* thread-local, valid until the next call           */
int vrr_device_ok(void);                   /* 1 if the current device is sm_100, else 0         */
int vrr_set_impl(int impl);                /* vrr_impl; process-wide; returns previous value    */
/* Number of kernel launches issued by this library since load (for bench.py's gpu_launches). */
uint64_t vrr_launch_count(void);
/* Dispatch counters per kernel family (VRR_IMPL_SIMT or VRR_IMPL_TCGEN05): how many of the entry
 * points that choose between the two families (patch_embed_fwd, qkv_rope_fwd, attn_fwd, attn_bwd, gemm)
 * took that family since load.  Lets a caller (and the parity tests) assert which family ran under
 * VRR_IMPL_AUTO instead of trusting the selection silently. */
uint64_t vrr_family_count(int family);
/* Tuning / experiment switches (process-wide): "attn_fwd_table_bulk" (0/1),
 * "attn_fwd_rescale_threshold_x100", "gemm_variant" (2 = CTA-pair kernels, default; 1 = 1-CTA kernels for the
 * QKV / patch-embed GEMMs, kept for A/B measurements), "attn_fwd_variant" (4, default: without bias the four-CTA-per-SM
 * kernel at every sequence length, with bias as 3; 3 = whole-sequence persistent kernel for N <= 256; 2 = one CTA per
 * 128-row tile always), "attn_bwd_variant" (3 = persistent kernel, default; 2 = two kernels), "attn_fwd_streams" (4 / 2
 * CTAs per SM in the tiled forward), "attn_fwd_poly_exp" (0..4 of every 8 exponential pairs by polynomial),
 * "ln_reg", "ln_bwd_minb", "simt_gemm_tile" - all kept for A/B measurements. */
int vrr_set_option(const char* name, int value);
/* Debug: CTA 0 of the next attention launches records clock64() phase stamps of its producer, issuer and
 * softmax roles into `device_buf` (1024 x int64, zero it first); NULL switches it off. */
int vrr_debug_timestamps(void* device_buf);

/* ---- (c) patch embedding: models/vit.py:164,248-258 ---------------------------------------- */
/* tokens[b][0][:]   = cls_token
 * tokens[b][1+p][:] = round_dtype(conv_weight[E][C*P*P] . unfold(images)[b][p][:] + conv_bias) (+ pos_embed[p])
 * images [B][C][Hi][Wi]                                            : element type `img_dtype`
 * weight [E][C][P][P], bias [E]                                    : element type `dtype` (the conv arithmetic)
 * cls_token [E], pos_embed, tokens [B][Np+1][E]                    : element type `tok_dtype`
 * pos_embed: NULL, or the absolute table [>= Np][E] (positional_encoding.py:37-40: added to patch
 * rows only).  (img fp32, dtype bf16, tok fp32) is the reference's autocast semantics: the image is
 * rounded to bf16 on load, the conv result is bf16, and the concatenation with the fp32 cls token
 * promotes the stream to fp32.  The bf16 tensor-core path (P % 8 == 0, C*P*P % 64 == 0, E % 64 == 0)
 * unfolds the patches into `workspace` (vrr_patch_embed_workspace_bytes, 256-byte aligned) and runs
 * a tcgen05 GEMM; otherwise workspace may be NULL. */
size_t vrr_patch_embed_workspace_bytes(int B, int C, int Hi, int Wi, int P, int E, int dtype);
int vrr_patch_embed_fwd(const void* images, const void* weight, const void* bias,
                        const void* cls_token, const void* pos_embed, void* tokens, void* workspace,
                        size_t workspace_bytes, int B, int C, int Hi, int Wi, int P, int E,
                        int img_dtype, int dtype, int tok_dtype, void* stream);
/* out[b*Np + p][c*P*P + i*P + j] = bf16(images[b][c][py*P+i][px*P+j]): the im2col matrix of the
 * stride-P convolution (needs P % 8 == 0).  Used by the forward above and by callers that take the
 * weight gradient with a plain library GEMM. */
int vrr_patch_unfold(const void* images, void* out, int B, int C, int Hi, int Wi, int P,
                     int img_dtype, void* stream);
/* d_tokens [B][Np+1][E] (`tok_dtype`).  Parameter gradients are reductions and ALWAYS fp32:
 * d_weight [E][C*P*P] (or NULL: skip it), d_bias [E], d_cls [E], d_pos [Np][E] or NULL; written (not
 * accumulated).  Images need no gradient (train.py:109-115). */
int vrr_patch_embed_bwd(const void* images, const void* d_tokens, void* d_weight, void* d_bias,
                        void* d_cls, void* d_pos, int B, int C, int Hi, int Wi, int P, int E,
                        int img_dtype, int dtype, int tok_dtype, void* stream);

/* ---- (a) QKV projection with RoPE epilogue: models/vit.py:47-68, rope_utils.py:3-37 -------- */
/* planes = split_heads(x . w_qkv^T) with q,k rows 1.. rotated by (cos,sin) in the epilogue.
 * x [B*N][E], w_qkv [3E][E] (row order (3,H,Dh), vit.py:47), cos/sin fp32 (see vrr_rope_mode),
 * planes [3][B][H][N][Dh]. */
int vrr_qkv_rope_fwd(const void* x, const void* w_qkv, const float* cos_tab, const float* sin_tab,
                     void* planes, int B, int N, int E, int H, int rope_mode, int dtype,
                     void* stream);
/* Optional fast path for the tables: `packed` = vrr_rope_pack_tables(cos_tab, sin_tab) of the SAME tables,
 * [heads][half_dim/2][rows] float4 {cos[2q], sin[2q], cos[2q+1], sin[2q+1]} (heads = H for rope-mixed, 1 for
 * rope-axial; rows = N-1; half_dim = Dh/2).  In the tcgen05 epilogue a lane is a token row: with the row-major tables
 * one load instruction touches 32 cache lines, with the packed table 4 (ViT-B rope-mixed: 189 -> ~140 us).  Results
 * are bit-identical.  `packed` may be NULL (= vrr_qkv_rope_fwd); cos_tab / sin_tab are still required (the SIMT
 * family reads them). */
int vrr_rope_pack_tables(const float* cos_tab, const float* sin_tab, float* packed, int heads, int rows,
                         int half_dim, void* stream);
int vrr_qkv_rope_fwd_packed(const void* x, const void* w_qkv, const float* cos_tab, const float* sin_tab,
                            const float* packed, void* planes, int B, int N, int E, int H, int rope_mode,
                            int dtype, void* stream);
/* Backward of the epilogue: un-rotates d_planes into token layout and reduces the table grads.
 * d_planes [3][B][H][N][Dh] (grads w.r.t. rotated q,k and v), planes = forward output,
 * d_qkv [B*N][3E] (gradient of x . w_qkv^T, ready for the two plain GEMMs dX = d_qkv . W and
 * dW = d_qkv^T . x), d_cos/d_sin fp32 with the shape of cos/sin or NULL (skipped), written. */
int vrr_qkv_rope_bwd(const void* d_planes, const void* planes, const float* cos_tab,
                     const float* sin_tab, void* d_qkv, float* d_cos, float* d_sin, int B, int N,
                     int E, int H, int rope_mode, int dtype, void* stream);

/* Stand-alone rotate-half of rope_utils.py:3-37 (the public apply_rotary_emb): q,k [B][H][Nr][Dh]
 * dense, cos/sin fp32 [Nr][Dh/2] (AXIAL) or [H][Nr][Dh/2] (MIXED); every row is rotated.
 * inverse != 0 applies the transpose rotation (the gradient w.r.t. q,k). */
int vrr_rope_apply(const void* q_in, const void* k_in, const float* cos_tab, const float* sin_tab,
                   void* q_out, void* k_out, int B, int H, int Nr, int Dh, int rope_mode,
                   int inverse, int dtype, void* stream);

/* Gradient of vrr_rope_apply (inverse == 0) w.r.t. its tables: d_cos, d_sin fp32 with the shape of
 * cos/sin, written (not accumulated); q_in, k_in = the un-rotated inputs, dq, dk = gradients w.r.t. the
 * rotated outputs.  Deterministic (no atomics). */
int vrr_rope_table_grad(const void* q_in, const void* k_in, const void* dq, const void* dk, float* d_cos,
                        float* d_sin, int B, int H, int Nr, int Dh, int rope_mode, int dtype, void* stream);

/* Plain GEMM used by the backward of the projection: C = op(A) . op(B), row-major operands,
 * op(A) is [M][K] (trans_a: A is stored [K][M]), op(B) is [K][N] (trans_b: B is stored [N][K]).
 * `dtype` is the element type of A and B, `c_dtype` that of C (fp32 C is allowed for bf16 inputs:
 * weight gradients).  fp32 accumulation always. */
int vrr_gemm(const void* a, const void* b, void* c, int M, int N, int K, int trans_a, int trans_b,
             int dtype, int c_dtype, void* stream);

/* GEMM of the Linear layers of the path and of their backward, with fused epilogues:
 * C[M][N] = op(A).op(B) with the vrr_gemm operand conventions; `dtype` = element type of A and B.
 *   forward  y = x.W^T (+b)        : trans_a 0, trans_b 1   (models/vit.py:47,91,118,285)
 *   backward dX = dY.W             : trans_a 0, trans_b 0
 *   backward dW = dY^T.X           : trans_a 1, trans_b 0, c_dtype VRR_F32
 * epilogue (vrr_epilogue): BIAS adds the fp32 `bias[N]` (cast to `dtype` first, like F.linear on `dtype`
 * tensors) before the single rounding to c_dtype; BIAS_GELU additionally writes C2 = gelu(C), exact erf GELU
 * (timm Mlp, act_layer=nn.GELU); BIAS_GELU_GRAD writes C = gelu(h) and C2 = gelu'(h) for h = round(acc + bias)
 * (h itself is not stored: the Mlp backward only needs gelu'); MUL multiplies by the INPUT matrix C2[M][N]
 * before rounding (fc2's dX fused with the GELU backward: dh = (dy . W2) * gelu'(h)).  accumulate != 0 (fp32 C, tcgen05 kernel only): C += result.
 * bf16 operands with N % 8 == 0, K % 8 == 0 (and M % 8 == 0 when trans_a) and 16-byte aligned pointers run
 * the tcgen05 CTA-pair kernel (cta_group::2, TMA in, TMA store / reduce-add out; fp32 C is split over K and
 * combined by TMA reduce-add, so its summation order is not deterministic); everything else (fp32 operands,
 * odd shapes) runs the FFMA kernel.  vrr_set_impl() forces a family; vrr_family_count() reports it. */
int vrr_gemm_ex(const void* a, const void* b, void* c, void* c2, const float* bias, int M, int N, int K,
                int trans_a, int trans_b, int dtype, int c_dtype, int epilogue, int accumulate, void* stream);

/* ---- (b) fused attention: models/vit.py:71-88 ---------------------------------------------- */
/* out = merge_heads(softmax(q k^T * scale + bias) v), lse = log-sum-exp of the logits.
 * planes [3][B][H][N][Dh], out [B][N][H*Dh], lse fp32 [B][H][N].  scale = Dh^-0.5 (vit.py:32),
 * applied after q k^T and before the bias (vit.py:71,75). */
int vrr_attn_fwd(const void* planes, const vrr_bias_desc* bias, void* out, float* lse, int B,
                 int H, int N, int Dh, float scale, int dtype, void* stream);
size_t vrr_attn_bwd_workspace_bytes(int B, int H, int N, int Dh, const vrr_bias_desc* bias);
/* d_planes [3][B][H][N][Dh] (dq, dk, dv w.r.t. the planes), written.
 * d_bias_param fp32 [bias->heads][bias->len] or NULL: gradient of the table / coefficients summed
 * over the batch, written (not accumulated).  workspace: device scratch of at least
 * vrr_attn_bwd_workspace_bytes() bytes. */
int vrr_attn_bwd(const void* planes, const vrr_bias_desc* bias, const void* out, const void* d_out,
                 const float* lse, void* d_planes, float* d_bias_param, void* workspace,
                 size_t workspace_bytes, int B, int H, int N, int Dh, float scale, int dtype,
                 void* stream);

/* ---- (f) N1, first "next" row: LayerNorm of the token stream, models/vit.py:113,117,122,124,210 - */
/* y = (x - mean) * rstd * gamma + beta over the last dimension (nn.LayerNorm, biased variance, eps
 * inside the sqrt).  x [M][E] (`x_dtype`), y [M][E] (`y_dtype`: the consumer GEMM's dtype, so that
 * under autocast the fp32 -> bf16 cast of the reference is fused into the store), gamma/beta fp32 [E],
 * mean/rstd fp32 [M] (saved for the backward). */
int vrr_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                      float* rstd, int M, int E, float eps, int x_dtype, int y_dtype, void* stream);
/* dx [M][E] (`x_dtype`), dgamma/dbeta fp32 [E] (written, not accumulated), from dy [M][E] (`y_dtype`). */
int vrr_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                      const float* rstd, void* dx, float* dgamma, float* dbeta, int M, int E,
                      int x_dtype, int y_dtype, void* stream);

/* Fused residual add + LayerNorm of the pre-LN block (models/vit.py:122-124: x = x + branch; the next
 * sub-block starts with norm(x)):  x_new = x + branch (fp32, written);  y = LayerNorm(x_new).
 * x, x_new fp32 [M][E]; branch `branch_dtype`; y `y_dtype`. */
int vrr_add_layernorm_fwd(const void* x, const void* branch, void* x_new, const float* gamma,
                          const float* beta, void* y, float* mean, float* rstd, int M, int E, float eps,
                          int branch_dtype, int y_dtype, void* stream);
/* g = d_xnew + LayerNorm'(dy): written as fp32 `dx` (gradient of x) and in `branch_dtype` as `d_branch`
 * (gradient of branch - the same values); d_xnew may be NULL (no other consumer of x_new). */
int vrr_add_layernorm_bwd(const void* dy, const void* d_xnew, const void* x_new, const float* gamma,
                          const float* mean, const float* rstd, void* dx, void* d_branch, float* dgamma,
                          float* dbeta, int M, int E, int branch_dtype, int y_dtype, void* stream);

/* fc2's input gradient with the GELU backward and fc1's bias gradient fused (autograd of timm Mlp, vit.py:118):
 *   c[m][n] = (op(a) . op(b))[m][n] * mul[m][n]   (`dtype`, as VRR_EPI_MUL)
 *   col_sums[n] = sum_m c[m][n]                   (fp32 [N], written; sums the values AS STORED, like autograd's
 *                                                  sum over the bf16 tensor)
 * bf16 on the tcgen05 kernel: the epilogue adds the column sums of each staged tile (fp32 atomics - summation
 * order not deterministic); otherwise vrr_gemm_ex(VRR_EPI_MUL) followed by vrr_colsum. */
int vrr_gemm_mul_colsum(const void* a, const void* b, void* c, const void* mul, float* col_sums, int M, int N,
                        int K, int trans_a, int trans_b, int dtype, void* stream);

/* Bias gradient of a Linear layer (out-proj vit.py:91, Mlp fc1/fc2 vit.py:118, head vit.py:285):
 * out[c] = sum_m x[m][c], fp32 [C] (written); x [M][C] `dtype`, C % 4 == 0. */
int vrr_colsum(const void* x, float* out, int M, int C, int dtype, void* stream);
/* Backward of the exact (erf) GELU between fc1 and fc2 of the Mlp (vit.py:118, act_layer=nn.GELU):
 * dh = dy * gelu'(h), written in `dtype`, and db[c] = sum_m dh[m][c] (fp32, written) - the fc1 bias
 * gradient - in the same pass.  dy, h, dh [M][C]. */
int vrr_gelu_bwd(const void* dy, const void* h, void* dh, float* db, int M, int C, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VRR_H_ */
