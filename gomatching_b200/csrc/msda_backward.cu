// Backward of the core operator -- placeholder until the "next" row lands.
#include "msda_launch.h"
#include "../../include/msda_b200.h"
namespace msda {
int launch_backward_f32(const float*, const int64_t*, const int64_t*, const float*, const float*, const float*, int, int,
                        int, int, int, int, int, float*, float*, float*, cudaStream_t) {
  return MSDA_E_UNSUPPORTED;
}
}  // namespace msda
