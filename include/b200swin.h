/* b200swin C-ABI: B200 (sm_100a) kernels for the Swin-V2 shifted-window attention block and
 * the SiLog depth loss of junnyfilm/multi-modal-monodepth-estimation.
 *
 * The reference is pure PyTorch and has no FFI of its own: its boundary is the
 * nn.Module surface of models/swin_transformer_v2.py and utils/criterion.py.  This header is the
 * thin C layer the drop-in modules (package b200swin) call through ctypes; every entry point
 * cites the reference lines whose computation it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - every function returns 0, or a negative code: -1 bad argument / unsupported shape,
 *    -2 CUDA error; b200swin_last_error() returns the text (thread-local);
 *  - all data pointers are DEVICE pointers owned by the caller, never retained or freed;
 *  - scratch memory is passed in by the caller (sizes from the *_workspace_bytes queries);
 *  - every call takes the cudaStream_t to launch on (as void*) and is asynchronous;
 *  - dtype codes: 0 = float32, 1 = bfloat16.  Reductions, softmax and LayerNorm statistics
 *    are always computed in float32 whatever the storage type;
 *  - no global state beyond lazily-resolved driver entry points; safe to call concurrently
 *    from several host threads on different streams/devices.
 */
#ifndef B200SWIN_H_
#define B200SWIN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SWIN_DTYPE_F32 0
#define B200SWIN_DTYPE_BF16 1

/* GEMM epilogues (b200swin_linear) */
#define B200SWIN_EPI_NONE 0     /* out = acc (+ bias)                                   */
#define B200SWIN_EPI_GELU 1     /* out = gelu_erf(acc + bias); aux (optional) = gelu'(acc+bias) */
#define B200SWIN_EPI_QKV 2      /* Swin-V2 qkv: + (q_bias,0,v_bias), L2-normalise q,k per head */
#define B200SWIN_EPI_DGELU 3    /* out = acc * aux_in   (aux_in = the gelu' saved by EPI_GELU) */
#define B200SWIN_EPI_ADD 4      /* out = acc + aux_in   (dgrad + the gradient of the residual branch) */
#define B200SWIN_EPI_RELU 5     /* out = max(acc + bias, 0); aux (optional) = 1 where positive, else 0 (ffn1 of
                                   Transformer_Encoder, models/cnn_transformer.py:193-194; backward through EPI_DGELU) */

int b200swin_version(void);
const char* b200swin_last_error(void);

/* ------------------------------------------------------------------------------------------
 * SiLog loss.  Replaces SiLogLoss.forward, utils/criterion.py:15-21, and its autograd.
 *   valid = target > 0;  d = log(target) - log(pred);  loss = sqrt(mean(d^2) - lambd*mean(d)^2)
 * pred: [n] (dtype pred_dtype), target: [n] float32.  stats[4] = {mean(d), n_valid, loss, mean(d^2)}
 * (float32, device) is written by fwd and read by bwd.  grad_out: device pointer to the 0-dim
 * upstream gradient.  grad_pred has pred's dtype.  No valid pixel -> NaN like the reference.
 * workspace: 3 doubles per CTA of the forward grid for n elements on the CURRENT device (ask on the device you launch on).
 * ------------------------------------------------------------------------------------------ */
size_t b200swin_silog_workspace_bytes(int64_t n);
int b200swin_silog_fwd(const void* pred, int pred_dtype, const float* target, int64_t n, float lambd,
                       float* loss, float* stats, void* workspace, size_t workspace_bytes, void* stream);
int b200swin_silog_bwd(const void* pred, int pred_dtype, const float* target, int64_t n, float lambd,
                       const float* stats, const float* grad_out, void* grad_pred, void* stream);

/* ------------------------------------------------------------------------------------------
 * Window plumbing as pure index maps (bit-exact).  Replace window_partition / window_reverse
 * (models/swin_transformer_v2.py:120-147), the cyclic torch.roll (:438, :458), the F.pad / crop
 * (:429-434, :462-463) and BasicLayer's shift mask (:874-892).
 *  window_gather : x[B,H,W,C] -> out[B*nW, ws*ws, C]   (zero pad to multiples of ws, roll by
 *                  -shift, partition) in ONE pass; shift=0 and H,W multiples of ws = window_partition.
 *  window_scatter: win[B*nW, ws*ws, C] -> out[B,H,W,C] (reverse, roll by +shift, crop); the exact
 *                  adjoint / inverse of window_gather.
 *  shift_mask    : out[nW, N, N] float32 in {0,-100} for an HxW token grid.
 * elem_bytes is the size of one element (2 or 4); C*elem_bytes must be a multiple of 4.
 * ------------------------------------------------------------------------------------------ */
int b200swin_window_gather(const void* x, void* out, int B, int H, int W, int C, int ws, int shift,
                           int elem_bytes, void* stream);
int b200swin_window_scatter(const void* win, void* out, int B, int H, int W, int C, int ws, int shift,
                            int elem_bytes, void* stream);
int b200swin_shift_mask(float* out, int H, int W, int ws, int shift, void* stream);
/* 2x2 patch merging as an index map.  Replaces the pad + four strided slices + cat of PatchMerging.forward
 * (models/swin_transformer_v2.py:660-672):
 *   backward = 0:  in = x[B,H,W,C]  ->  out = merged[B, ceil(H/2)*ceil(W/2), 4C],
 *                  merged[b, i2*W2+j2, k*C+c] = x[b, 2*i2+(k&1), 2*j2+(k>>1), c]  (zero beyond H, W);
 *   backward = 1:  in = d merged  ->  out = dx[B,H,W,C]  (the adjoint; pad positions dropped).
 * C*elem_bytes must be a multiple of 16. */
int b200swin_patch_merge(const void* in, void* out, int B, int H, int W, int C, int elem_bytes, int backward,
                         void* stream);
/* Patchify for the stride = kernel patch-embedding conv (PatchEmbed.forward, models/swin_transformer_v2.py:941-957):
 *   cols[(b,i,j), c*ph*pw + kh*pw + kw] = x[b, c, i*ph+kh, j*pw+kw]   (zero beyond H, W; x is NCHW)
 * so that conv(x, weight[E,Cin,ph,pw]) = b200swin_gemm_bf16(cols, weight viewed as [E, Cin*ph*pw]) in token layout. */
int b200swin_patchify(const void* x, int x_dtype, void* cols, int cols_dtype, int B, int Cin, int H, int W, int ph,
                      int pw, void* stream);

/* Depthwise 3x3 convolution (stride 1, zero padding 1, no bias) in the token layout: the conv_proj of ConvMlp
 * (models/swin_transformer_v2.py:98-111, mlp_type='conv' / 'conv_ln'), without the NHWC <-> NCHW permutes around cuDNN.
 *   x, y [B,H,W,C] (dtype), weight [C,1,3,3] float32 (nn.Conv2d layout):  y[b,i,j,c] = sum_uv w[c,u,v] x[b,i+u-1,j+v-1,c]
 * transpose = 1 applies the adjoint (taps flipped): dx from dy.  wgrad: dweight [C,1,3,3] float32, deterministic
 * (per-CTA partials in `workspace`, fixed-order reduce).  C % 4 == 0. */
int b200swin_dwconv3x3(const void* x, const float* weight, void* y, int B, int H, int W, int C, int dtype, int transpose,
                       void* stream);
size_t b200swin_dwconv3x3_wgrad_workspace_bytes(int B, int H, int W, int C);
int b200swin_dwconv3x3_wgrad(const void* x, const void* dy, float* dweight, int B, int H, int W, int C, int dtype,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm (+ DropPath scale + residual).  Replaces LayerNormFP32.forward
 * (models/swin_transformer_v2.py:41-47) and the post-norm residual adds (:472-474, :482-483):
 *   y[r,:] = residual[r,:] + row_scale[r / rows_per_scale] * (LN(x[r,:]) * gamma + beta)
 * residual and row_scale may be NULL.  x,residual,y,dy,dx have `dtype`; gamma,beta,mean,rstd and
 * the gradients of gamma/beta are float32.  C must be a multiple of 4.
 * bwd: dx, plus dgamma/dbeta [C] reduced deterministically through `workspace`; dcolsum (optional, NULL to
 * skip; bf16 tensors with C % 8 == 0 and C <= 1536 only) receives the fp32 column sums of dx, i.e. the bias
 * gradient of the Linear whose output was normalised (proj / fc2), saving a separate pass over dx.
 * The gradient of `residual` is dy itself (the caller aliases it).
 * ------------------------------------------------------------------------------------------ */
int b200swin_ln_fwd(const void* x, const void* residual, const float* gamma, const float* beta,
                    const float* row_scale, int64_t rows_per_scale, void* y, float* mean, float* rstd,
                    int64_t rows, int C, float eps, int dtype, void* stream);
/* bf16 activations with an fp32 RESIDUAL STREAM (what torch.autocast does: LayerNorm outputs and residual adds stay
 * fp32): x is the bf16 output of the producing GEMM, residual32 (nullable) the fp32 stream; the result is written twice,
 * y32 = residual32 + row_scale * LN(x) in fp32 (the stream handed to the next block) and y = bf16(y32) (the operand of the
 * next GEMM).  Backward is b200swin_ln_bwd on the bf16 tensors (the gradient stream stays bf16).  C % 8 == 0, C <= 1024. */
int b200swin_ln_fwd_stream32(const void* x, const float* residual32, const float* gamma, const float* beta,
                             const float* row_scale, int64_t rows_per_scale, void* y, float* y32, float* mean,
                             float* rstd, int64_t rows, int C, float eps, void* stream);
/* Post-norm transformer layer (Transformer_Encoder.forward, models/cnn_transformer.py:202-203, :208-209):
 *   y = LN(x + xadd) * gamma + beta        x: the fp32 stream, xadd: the branch output (xadd_dtype: float32 or bfloat16)
 * in ONE pass instead of an elementwise add, a LayerNorm and a cast: y float32, y16 (nullable) = bf16(y) for the next
 * GEMM, xsum (nullable) = x + xadd in float32 -- the tensor b200swin_ln_bwd normalises (dtype 0); its dx is the gradient
 * of x and of xadd alike.  C % 4 == 0, C <= 1024. */
int b200swin_ln_fwd_sum(const float* x, const void* xadd, int xadd_dtype, const float* gamma, const float* beta, float* y,
                        void* y16, float* xsum, float* mean, float* rstd, int64_t rows, int C, float eps, void* stream);
size_t b200swin_ln_bwd_workspace_bytes(int64_t rows, int C);
int b200swin_ln_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                    const float* row_scale, int64_t rows_per_scale, void* dx, float* dgamma, float* dbeta,
                    float* dcolsum, int64_t rows, int C, int dtype, void* workspace, size_t workspace_bytes,
                    void* stream);

/* ------------------------------------------------------------------------------------------
 * Continuous position bias table (cpb_mlp / rpe_mlp of Swin-V2).  Replaces the bias-table branch of
 * WindowAttention.forward (models/swin_transformer_v2.py:304-313, rpe_output_type='sigmoid'):
 *   table[T,nH] = 16 * sigmoid( relu(coords[T,2] @ W0[HID,2]^T + b0[HID]) @ W2[nH,HID]^T ),  all float32.
 * bwd: given dtable (and the saved table) writes dW0, db0, dW2 (deterministic, no atomics).  nH <= 64.
 * Optionally (logit_scale != NULL) the same launches also compute the per-head temperature of the cosine attention,
 *   scale[h] = exp(min(logit_scale[h], ln 100))                                        (:294)
 * and its gradient d logit_scale[h] = d scale[h] * scale[h] where the clamp is inactive.
 * ------------------------------------------------------------------------------------------ */
int b200swin_cpb_fwd(const float* coords, const float* w0, const float* b0, const float* w2, float* table,
                     const float* logit_scale, float* scale, int T, int HID, int nH, void* stream);
int b200swin_cpb_bwd(const float* coords, const float* w0, const float* b0, const float* w2, const float* table,
                     const float* dtable, float* dw0, float* db0, float* dw2, const float* logit_scale,
                     const float* dscale, float* dlogit_scale, int T, int HID, int nH, void* stream);

/* ------------------------------------------------------------------------------------------
 * Attention core.  Replaces, for attn_type='cosine_mh', the body of WindowAttention.forward between
 * the qkv projection and the output projection (models/swin_transformer_v2.py:292-328) TOGETHER with
 * the block's pad / roll / window_partition / window_reverse / roll-back / crop (:429-463) and
 * BasicLayer's shift mask (:874-892), all folded into the kernel's load/store addressing:
 *
 *   qkv  [B,H,W,3C]  natural token order, columns [q | k | v], head h = columns h*32..h*32+31 of each
 *                    part; q and k ALREADY L2-normalised per head (EPI_QKV epilogue of b200swin_linear),
 *   out  [B,H,W,C]   softmax(scale_h * q.k^T + table16[rel_idx(i,j),h] + mask) @ v, natural order,
 *   lse  [B*nW,nH,N] float32 log-sum-exp per row (saved for backward), N = ws*ws.
 *
 * Pad tokens (zero rows added by F.pad) are not masked by the reference: their k is 0, their v is
 * v_bias and their q is q_bias; the kernel synthesises them from qpad[C] (= normalised q_bias) and
 * vpad[C] (= v_bias), both float32 and required only when H or W is not a multiple of ws.
 * table16 [(2ws-1)^2, nH] float32 is the bias table AFTER 16*sigmoid (:304-313), scale [nH] float32 is
 * exp(min(logit_scale, ln 100)) (:294).  The shift mask ({0,-100}) is computed on the fly from token
 * coordinates when shift > 0; alternatively an explicit mask [nWm,N,N] float32 (window b uses
 * mask[b % nWm], :319-322) can be given - this is the standalone WindowAttention.forward(x, mask)
 * call, made with B = B_, H = W = ws, shift = 0.
 *
 * Backward: dout [B,H,W,C] -> dqkv [B,H,W,3C] holding dq, dk (gradients w.r.t. the UN-normalised q, k:
 * the F.normalize backward is applied with inv_norm [B*H*W, 2, nH] float32 = 1/max(|q|,1e-12),
 * 1/max(|k|,1e-12)) and dv; plus float32 accumulators that the caller zero-initialises:
 * dtable16 [(2ws-1)^2, nH], dscale [nH] (= sum dS*cos), dvpad [C] (gradient reaching v_bias through
 * pad tokens).  impl: 0 = fp32 CUDA-core kernel (any dtype, reference precision),
 * 1 = tensor-core kernels, chosen by window size (bf16 storage only; any window up to 32x32: register-resident
 * warp-level MMA kernels for windows 4/6/7/8/12 -- a window is 1..9 tiles of 16 rows, no 128-row tile to fill --
 * KV-blocked tcgen05 kernels otherwise), 2 = the KV-blocked tcgen05 kernels whatever the window, 3 = the single-tile
 * tcgen05 kernels (windows 4/6/7/8/12 forward, 4/6/7/12 backward), 4 = the warp-level MMA kernels without warp
 * specialisation (2-4 exist for A/B timing and cross-checks).  head_dim must be 32 (every Swin-V2 variant).
 * lse of a query row that is a pad token may be +inf (its output row does not exist).
 * out_lo (nullable; tensor-core implementations only): bf16 tensor of out's shape that receives the rounding residual
 * O - bf16(O).  Given back to the backward it makes D = <dO, O> accurate to ~2^-17; the bias-table and temperature
 * gradients are sums of dS = P (dP - D) that cancel row by row, and the 2^-9 rounding of O alone would put an error
 * of the size of the signal into them.
 * ------------------------------------------------------------------------------------------ */
int b200swin_attn_fwd(const void* qkv, void* out, void* out_lo, float* lse, const float* table16, const float* scale,
                      const float* qpad, const float* vpad, const float* mask, int nWm, int B, int H, int W,
                      int C, int nH, int ws, int shift, int dtype, int impl, void* stream);
/* attn_type='normal' (models/swin_transformer_v2.py:296-298): q and k are NOT normalised (the caller projects with
 * B200SWIN_EPI_NONE), scale[h] = qk_scale or head_dim^-0.5 for every head, qpad = q_bias as it is, and the backward is
 * called with inv_norm = NULL: dq, dk are then the plain gradients (no F.normalize backward).  Served by impl 0 and 2
 * (the warp-MMA forward drops the row maximum under the cosine bound |q.k| <= 1, which does not hold here). */
/* bytes of caller-owned scratch the backward needs for this shape / implementation (0: none) */
size_t b200swin_attn_bwd_workspace_bytes(int B, int H, int W, int nH, int ws, int dtype, int impl);
int b200swin_attn_bwd(const void* qkv, const void* out, const void* out_lo, const void* dout, const float* lse,
                      const float* inv_norm,
                      const float* table16, const float* scale, const float* qpad, const float* vpad,
                      const float* mask, int nWm, void* dqkv, float* dtable16, float* dscale, float* dvpad,
                      float* dqkv_colsum, int B, int H, int W, int C, int nH, int ws, int shift, int dtype, int impl,
                      void* workspace, size_t workspace_bytes, void* stream);
/* 1 when b200swin_attn_bwd for this window / dtype / impl can also add the column sums of dq and dv -- the gradients of
 * q_bias and v_bias that flow through real tokens -- to dqkv_colsum [3C] float32 (zero-initialised by the caller; the
 * middle C entries, k has no bias, are not touched).  The rows are in registers in the backward's epilogue; without it the
 * caller makes one more pass over dqkv (b200swin_colsum).  dqkv_colsum must be NULL where this returns 0. */
int b200swin_attn_bwd_colsum_supported(int ws, int dtype, int impl);

/* ------------------------------------------------------------------------------------------
 * Global multi-head attention core (no windows, no bias, no mask):  out = softmax(scale * q k^T) v  per (batch, head).
 * Replaces the scaled-dot-product inside nn.MultiheadAttention as Transformer_Encoder.forward uses it
 * (models/cnn_transformer.py:192-216: hidden 512 = 8 heads x 64 over the 30x40 = 1200 tokens of a 480x640 frame; 4 x 64
 * for hidden 256).  q [B,Nq,ldq], k [B,Nk,ldk], v [B,Nk,ldv], out [B,Nq,ldo]: rows with a stride given in ELEMENTS, head h
 * in columns h*head_dim ..., so the operands may be slices of one packed projection buffer; the projected q is NOT
 * pre-scaled (scale, normally head_dim^-0.5, is applied to the logits).  lse [B,nH,Nq] float32 is saved for the backward.
 * head_dim 32 or 64.  bf16: warp-level MMA kernels (forward; backward = D prep + dK/dV pass + dQ pass, deterministic);
 * float32: CUDA-core kernels at reference precision.  Backward workspace: B*Nq*nH floats (D = <dout, out>).
 * mha_avg_weights: weights[B,Nq,Nk] (tensor dtype) = mean over heads of the attention probabilities, the second return
 * value of nn.MultiheadAttention.forward(need_weights=True) that the reference receives (and drops) at :201.
 * ------------------------------------------------------------------------------------------ */
int b200swin_mha_fwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, void* out,
                     int64_t ldo, float* lse, int B, int Nq, int Nk, int nH, int head_dim, float scale, int dtype,
                     void* stream);
size_t b200swin_mha_bwd_workspace_bytes(int B, int Nq, int nH);
int b200swin_mha_bwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, const void* out,
                     int64_t ldo, const void* dout, int64_t lddo, const float* lse, void* dq, void* dk, void* dv,
                     int64_t lddq, int64_t lddk, int64_t lddv, int B, int Nq, int Nk, int nH, int head_dim, float scale,
                     int dtype, void* workspace, size_t workspace_bytes, void* stream);
int b200swin_mha_avg_weights(const void* q, const void* k, int64_t ldq, int64_t ldk, const float* lse, void* weights, int B,
                             int Nq, int Nk, int nH, int head_dim, float scale, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dense contraction on tcgen05 tensor cores:  out[M,N] = epilogue( A[M,K] . B[N,K]^T ).
 * Replaces F.linear / nn.Linear at models/swin_transformer_v2.py:286 (qkv), :334 (proj), :77 and :87
 * (Mlp.fc1 / fc2) and their autograd (dgrad, wgrad).
 *  - a_hi / b_hi: bf16 operands.  *_mn_major = 0: stored [M][K] / [N][K] (K contiguous, i.e. x and
 *    nn.Linear.weight as they are); 1: stored [K][M] / [K][N].  dgrad (dX = dY.W) passes W with
 *    b_mn_major = 1, wgrad (dW = dY^T.X) passes dY and X with both = 1: no transposed copy is made.
 *  - a_lo / b_lo (both or neither): low bf16 halves from b200swin_split_bf16; the kernel then
 *    accumulates hi.hi + hi.lo + lo.hi, which reproduces an fp32 GEMM to ~1e-5 relative.
 *  - epilogue (B200SWIN_EPI_*): NONE: + bias[N] (nullable).  GELU: + bias, out = erf-GELU, aux_out (nullable)
 *    receives gelu'(pre-activation).  RELU: the same with max(., 0) and its 0/1 derivative.  DGELU: out = acc * aux_in
 *    (that saved derivative).  QKV: N = 3C; adds bias
 *    (= q_bias[C]) to the q columns, nothing to k, bias2 (= v_bias[C]) to v; L2-normalises every
 *    32-wide head slice of q and k in fp32 (F.normalize, eps 1e-12, :292-293) and writes
 *    inv_norm[M,2,nH] = 1/max(|q|,eps), 1/max(|k|,eps) (nullable).
 *  - out/aux dtype = out_dtype; accumulation is fp32.  splits > 1: split-K over blockIdx.z with fp32
 *    partials in workspace and a fixed-order reduction (NONE epilogue only; used by wgrad).
 * Shape rules: N % 4 == 0; the contiguous dimension of each operand % 8 == 0; base pointers 16-B aligned.
 * ------------------------------------------------------------------------------------------ */
int b200swin_gemm_splits(int64_t M, int64_t N, int64_t K);
size_t b200swin_gemm_workspace_bytes(int64_t M, int64_t N, int splits);
int b200swin_gemm_bf16(const void* a_hi, const void* a_lo, int a_mn_major, const void* b_hi, const void* b_lo,
                       int b_mn_major, int64_t M, int64_t N, int64_t K, int epilogue, const float* bias,
                       const float* bias2, const void* aux_in, void* aux_out, float* inv_norm, int nH, void* out,
                       int out_dtype, int splits, void* workspace, size_t workspace_bytes, void* stream);

/* Helpers around the GEMM (HBM-bound): fp32 -> bf16 hi (+ lo residual, nullable) split, and column sums
 * out[n] = extra[n] + sum_m x[m, col0+n] over an [M, ld] matrix (bias gradients: q_bias/v_bias at
 * :283-285, proj/fc biases). */
int b200swin_split_bf16(const float* src, void* hi, void* lo, int64_t n, void* stream);
size_t b200swin_colsum_workspace_bytes(int64_t M, int ncols);
int b200swin_colsum(const void* x, int dtype, int64_t M, int64_t ld, int64_t col0, int ncols, const float* extra,
                    float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused multi-tensor AdamW with per-tensor learning-rate scale and weight decay.  Replaces the optimizer step the
 * reference builds with SwinLayerDecayOptimizerConstructor (models/optimizer.py:36-104: ~60 parameter groups, layer
 * decay x {decay, no_decay}) and runs with torch.optim.AdamW after rewriting every group's lr (train.py:195-203).
 * params / grads / exp_avg / exp_avg_sq are flat float32 buffers of nchunks * b200swin_adamw_chunk() elements in which
 * every tensor occupies a whole number of chunks; chunk_tensor[c] is the tensor index of chunk c (-1: padding, skipped);
 * lr_scale[t] and weight_decay[t] are per tensor.  lr and step point to DEVICE scalars (float32; step = 1, 2, ... is the
 * number of this update), so a captured launch follows a schedule.  grads are multiplied by grad_scale first (1/world
 * size after a SUM all-reduce).  params_bf16 (nullable) receives the bf16 copy of the updated parameters that the GEMMs
 * read.  Math = torch.optim.AdamW (decoupled decay, bias correction), float32.
 * ------------------------------------------------------------------------------------------ */
int b200swin_adamw_chunk(void);
int b200swin_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* params_bf16,
                        const int* chunk_tensor, const float* lr_scale, const float* weight_decay, const float* lr,
                        const float* step, double beta1, double beta2, float eps, float grad_scale, int64_t nchunks,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SWIN_H_ */
