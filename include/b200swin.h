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
#define B200SWIN_EPI_GELU 1     /* out = gelu_erf(acc + bias); aux (optional) = acc+bias */
#define B200SWIN_EPI_QKV 2      /* Swin-V2 qkv: + (q_bias,0,v_bias), L2-normalise q,k per head */
#define B200SWIN_EPI_DGELU 3    /* out = (acc) * gelu'(aux)                              */

int b200swin_version(void);
const char* b200swin_last_error(void);

/* ------------------------------------------------------------------------------------------
 * SiLog loss.  Replaces SiLogLoss.forward, utils/criterion.py:15-21, and its autograd.
 *   valid = target > 0;  d = log(target) - log(pred);  loss = sqrt(mean(d^2) - lambd*mean(d)^2)
 * pred: [n] (dtype pred_dtype), target: [n] float32.  stats[4] = {mean(d), n_valid, loss, mean(d^2)}
 * (float32, device) is written by fwd and read by bwd.  grad_out: device pointer to the 0-dim
 * upstream gradient.  grad_pred has pred's dtype.  No valid pixel -> NaN like the reference.
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

/* ------------------------------------------------------------------------------------------
 * LayerNorm (+ DropPath scale + residual).  Replaces LayerNormFP32.forward
 * (models/swin_transformer_v2.py:41-47) and the post-norm residual adds (:472-474, :482-483):
 *   y[r,:] = residual[r,:] + row_scale[r / rows_per_scale] * (LN(x[r,:]) * gamma + beta)
 * residual and row_scale may be NULL.  x,residual,y,dy,dx have `dtype`; gamma,beta,mean,rstd and
 * the gradients of gamma/beta are float32.  C must be a multiple of 4.
 * bwd: dx, plus dgamma/dbeta [C] reduced deterministically through `workspace`.
 * The gradient of `residual` is dy itself (the caller aliases it).
 * ------------------------------------------------------------------------------------------ */
int b200swin_ln_fwd(const void* x, const void* residual, const float* gamma, const float* beta,
                    const float* row_scale, int64_t rows_per_scale, void* y, float* mean, float* rstd,
                    int64_t rows, int C, float eps, int dtype, void* stream);
size_t b200swin_ln_bwd_workspace_bytes(int64_t rows, int C);
int b200swin_ln_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                    const float* row_scale, int64_t rows_per_scale, void* dx, float* dgamma, float* dbeta,
                    int64_t rows, int C, int dtype, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SWIN_H_ */
