// cib200_binned.cu - K1 "binned" strategy (queries binned by coefficient brick, bricks staged by TMA).
// Placeholder until the direct strategy is measured on hardware.
#include "cib200_internal.cuh"

namespace cib {
bool binned_supported(const cib200_itp *) { return false; }
long long binned_scratch_bytes(const cib200_itp *, long long) { return 0; }
int launch_eval_binned(cib200_itp *, const Coords &, void *, long long, cudaStream_t) {
  set_error("BINNED strategy not built");
  return CIB200_EUNSUPPORTED;
}
void binned_destroy(cib200_itp *) {}
}  // namespace cib
