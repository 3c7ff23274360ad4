// Attention core on tcgen05 tensor cores (impl = 1).  Placeholder until the kernel lands: reports
// "unsupported" so callers fail loudly instead of silently taking another path.
#include "common.cuh"
#include "../../include/b200swin.h"

namespace b200swin {
int attn_fwd_tc(const void*, void*, float*, const float*, const float*, const float*, const float*, const float*, int,
                int, int, int, int, int, int, int, cudaStream_t) {
  set_error("attn_fwd: tensor-core implementation not available in this build");
  return B200SWIN_EINVAL;
}
int attn_bwd_tc(const void*, const void*, const void*, const float*, const float*, const float*, const float*,
                const float*, const float*, const float*, int, void*, float*, float*, float*, int, int, int, int, int,
                int, int, cudaStream_t) {
  set_error("attn_bwd: tensor-core implementation not available in this build");
  return B200SWIN_EINVAL;
}
}  // namespace b200swin
