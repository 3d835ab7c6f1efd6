// K6: ToMe bipartite K/V merge -- placeholder until the kernels land
#include "common.cuh"
#include "ops.h"

namespace gyre {
int tome_workspace_bytes(int B, int N, int C, size_t* bytes) {
  *bytes = 0;
  return 0;
}
int tome_merge_kv(const __half*, const __half*, int, int, int, int, int, __half*, __half*, void*, size_t, cudaStream_t) {
  set_last_error("tome_merge_kv: not implemented yet");
  return -100;
}
}  // namespace gyre
