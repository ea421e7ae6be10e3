// placeholder — replaced by the tcgen05 bf16 path
#include "hmz_net.cuh"
namespace hmz {
int64_t tc_packed_bytes(int) { return -1; }
void tc_pack(const float* const*, int, void*) {}
int tc_net_recurrent(const void*, const void*, int64_t, const uint16_t*, const uint8_t*, void*, int64_t, int64_t, int,
                     float*, float*, float*, int64_t, cudaStream_t) {
  return fail(HMZ_ERR_UNSUPPORTED, "bf16 tensor-core path not built yet");
}
}  // namespace hmz
