// sl_pop_bg_tc: tcgen05 background MLP -- placeholder until the tensor-core kernel lands.
#include "common.cuh"
extern "C" int sl_pop_bg_tc(const uint16_t*, int, int, int, const uint16_t*, const uint16_t*, const uint16_t*,
                            const uint16_t*, const float*, float*, int, int, void*) {
  return SL_EINVAL;
}
