// Dispatch of the tensor-core (impl >= 1, bf16 storage) attention core between the kernel families:
//   attn_mma.cu                       register-resident warp-level MMA, one warp per 16 query rows (12x12 windows: 144
//                                     rows are nine 16-row tiles; the SimMIM geometry [12,12,12,6] of config 2);
//   attn_fwd_ws.cu / attn_bwd_ws.cu   tcgen05, one (window, head) per work item, the whole window in ONE tile of TMEM
//                                     (windows 4, 6, 7, 8 and 12);
//   attn_flash.cu                     tcgen05, KV-blocked, any window up to 32x32 (16 / 24 / 30 and the odd sizes).
// All replace models/swin_transformer_v2.py:295-328 + :429-463 + :874-892.  impl = 1 picks by window size alone (no
// run-time switch); impl = 2 / 3 / 4 ask for the KV-blocked / single-tile / warp-MMA kernels (A/B timing, cross-checks).
#include "common.cuh"
#include "../../include/b200swin.h"

namespace b200swin {

bool attn_fwd_ws_supported(int ws);
int attn_fwd_ws(const void* qkv, void* out, void* out_lo, float* lse, const float* table16, const float* scale, const float* qpad,
                const float* vpad, int B, int H, int W, int C, int nH, int ws, int shift, cudaStream_t st);
bool attn_fwd_mma_supported(int ws);
int attn_fwd_mma(const void* qkv, void* out, void* out_lo, float* lse, const float* table16, const float* scale, const float* qpad,
                 const float* vpad, int B, int H, int W, int C, int nH, int ws, int shift, cudaStream_t st);
bool attn_bwd_mma_supported(int ws);
size_t attn_bwd_mma_workspace_bytes(int B, int H, int W, int nH);
int attn_bwd_mma(const void* qkv, const void* out, const void* out_lo, const void* dout, const float* lse, const float* inv_norm,
                 const float* table16, const float* scale, const float* qpad, const float* vpad, void* dqkv,
                 float* dtable16, float* dscale, float* dvpad, float* dcol, void* workspace, int B, int H, int W, int C,
                 int nH, int ws, int shift, bool spec, cudaStream_t st);
bool attn_bwd_ws_supported(int ws);
size_t attn_bwd_ws_workspace_bytes(int B, int H, int W, int nH);
int attn_bwd_ws(const void* qkv, const void* out, const void* out_lo, const void* dout, const float* lse, const float* inv_norm,
                const float* table16, const float* scale, const float* qpad, const float* vpad, void* dqkv,
                float* dtable16, float* dscale, float* dvpad, void* workspace, int B, int H, int W, int C, int nH, int ws,
                int shift, cudaStream_t st);
int attn_fwd_flash(const void* qkv, void* out, void* out_lo, float* lse, const float* table16, const float* scale, const float* qpad,
                   const float* vpad, int B, int H, int W, int C, int nH, int ws, int shift, cudaStream_t st);
size_t attn_bwd_flash_workspace_bytes(int B, int H, int W, int nH);
int attn_bwd_flash(const void* qkv, const void* out, const void* out_lo, const void* dout, const float* lse, const float* inv_norm,
                   const float* table16, const float* scale, const float* qpad, const float* vpad, void* dqkv,
                   float* dtable16, float* dscale, float* dvpad, void* workspace, int B, int H, int W, int C, int nH,
                   int ws, int shift, cudaStream_t st);

// kernel family asked for by the caller: impl 1 -> auto, 2 -> KV-blocked, 3 -> single-tile tcgen05, 4 -> warp-level MMA
enum { kFamilyAuto = 0, kFamilyFlash = 1, kFamilyWs = 2, kFamilyMma = 3 };

static int check_tc(const char* what, const void* qkv, const void* io, const void* mask, int B, int H, int W, int C,
                    int nH, int ws, int shift) {
  BSW_REQUIRE(mask == nullptr, "%s(tc): the tensor-core kernels derive the shift mask on the fly (no explicit mask tensor)", what);
  BSW_REQUIRE(B > 0 && H > 0 && W > 0 && nH > 0 && C == nH * 32, "%s(tc): head_dim must be 32 (C=%d, nH=%d)", what, C, nH);
  BSW_REQUIRE(ws >= 1 && ws <= 32, "%s(tc): window %d outside [1, 32]", what, ws);
  BSW_REQUIRE(shift >= 0 && shift < ws, "%s(tc): bad shift", what);
  BSW_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(io) & 15) == 0 && C % 8 == 0,
              "%s(tc): tensors must be 16-byte aligned", what);
  BSW_REQUIRE((int64_t)B * H * W < (1ll << 29), "%s(tc): too many tokens", what);
  return B200SWIN_OK;
}

int attn_fwd_tc(const void* qkv, void* out, void* out_lo, float* lse, const float* table16, const float* scale, const float* qpad,
                const float* vpad, const float* mask, int nWm, int B, int H, int W, int C, int nH, int ws, int shift,
                int family, cudaStream_t st) {
  (void)nWm;
  int rc = check_tc("attn_fwd", qkv, out, mask, B, H, W, C, nH, ws, shift);
  if (rc) return rc;
  BSW_REQUIRE(family != kFamilyWs || attn_fwd_ws_supported(ws), "attn_fwd: no single-tile kernel for window %d", ws);
  BSW_REQUIRE(family != kFamilyMma || attn_fwd_mma_supported(ws), "attn_fwd: no warp-MMA kernel for window %d", ws);
  if ((family == kFamilyAuto || family == kFamilyMma) && attn_fwd_mma_supported(ws))
    return attn_fwd_mma(qkv, out, out_lo, lse, table16, scale, qpad, vpad, B, H, W, C, nH, ws, shift, st);
  if ((family == kFamilyAuto || family == kFamilyWs) && attn_fwd_ws_supported(ws))
    return attn_fwd_ws(qkv, out, out_lo, lse, table16, scale, qpad, vpad, B, H, W, C, nH, ws, shift, st);
  return attn_fwd_flash(qkv, out, out_lo, lse, table16, scale, qpad, vpad, B, H, W, C, nH, ws, shift, st);
}

// whether the backward for this window adds the column sums of dq / dv (q_bias / v_bias gradients) to `dqkv_colsum`
bool attn_bwd_tc_colsum_supported(int ws, int family) { return family == kFamilyAuto && ws == 12 && attn_bwd_mma_supported(ws); }

size_t attn_bwd_tc_workspace_bytes(int B, int H, int W, int nH, int ws) {
  // all families need the same scratch: D = <dO, O> per (token, head)
  if (attn_bwd_mma_supported(ws)) return attn_bwd_mma_workspace_bytes(B, H, W, nH);
  return attn_bwd_ws_supported(ws) ? attn_bwd_ws_workspace_bytes(B, H, W, nH) : attn_bwd_flash_workspace_bytes(B, H, W, nH);
}

int attn_bwd_tc(const void* qkv, const void* out, const void* out_lo, const void* dout, const float* lse, const float* inv_norm,
                const float* table16, const float* scale, const float* qpad, const float* vpad, const float* mask,
                int nWm, void* dqkv, float* dtable16, float* dscale, float* dvpad, float* dcol, void* workspace,
                size_t workspace_bytes, int B, int H, int W, int C, int nH, int ws, int shift, int family,
                cudaStream_t st) {
  (void)nWm;
  int rc = check_tc("attn_bwd", qkv, dqkv, mask, B, H, W, C, nH, ws, shift);
  if (rc) return rc;
  BSW_REQUIRE(workspace && workspace_bytes >= attn_bwd_tc_workspace_bytes(B, H, W, nH, ws),
              "attn_bwd(tc): workspace too small (see b200swin_attn_bwd_workspace_bytes)");
  BSW_REQUIRE(family != kFamilyWs || attn_bwd_ws_supported(ws), "attn_bwd: no single-tile kernel for window %d", ws);
  BSW_REQUIRE(family != kFamilyMma || attn_bwd_mma_supported(ws), "attn_bwd: no warp-MMA kernel for window %d", ws);
  if ((family == kFamilyAuto || family == kFamilyMma) && attn_bwd_mma_supported(ws))
    return attn_bwd_mma(qkv, out, out_lo, dout, lse, inv_norm, table16, scale, qpad, vpad, dqkv, dtable16, dscale, dvpad, dcol,
                        workspace, B, H, W, C, nH, ws, shift, /*spec=*/family == kFamilyAuto, st);
  BSW_REQUIRE(!dcol, "attn_bwd: this kernel family does not produce the column sums (see b200swin_attn_bwd_colsum_supported)");
  if ((family == kFamilyAuto || family == kFamilyWs) && attn_bwd_ws_supported(ws))
    return attn_bwd_ws(qkv, out, out_lo, dout, lse, inv_norm, table16, scale, qpad, vpad, dqkv, dtable16, dscale, dvpad, workspace,
                       B, H, W, C, nH, ws, shift, st);
  return attn_bwd_flash(qkv, out, out_lo, dout, lse, inv_norm, table16, scale, qpad, vpad, dqkv, dtable16, dscale, dvpad,
                        workspace, B, H, W, C, nH, ws, shift, st);
}

}  // namespace b200swin
