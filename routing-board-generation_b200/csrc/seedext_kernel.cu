// seedext_kernel.cu -- SeedExtension generator (placeholder until the kernel lands).
#include "rbg_host.h"

namespace rbg {
int launch_seedext(SeedExtParams p, int64_t max_boards, cudaStream_t stream) {
  (void)p; (void)max_boards; (void)stream;
  return set_error(RBG_EINVAL, "SeedExtension kernel not built into this library yet");
}
}  // namespace rbg
