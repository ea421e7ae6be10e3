// placeholder — replaced by the network kernels
#include "hmz_common.cuh"
using namespace hmz;
extern "C" {
int64_t hmz_weights_packed_bytes(int, int) { return -1; }
int hmz_weights_pack(const float* const*, int, int, void*) { return fail(HMZ_ERR_UNSUPPORTED, "not built yet"); }
int hmz_net_initial(const void*, int, int, const uint32_t*, const float*, void*, int64_t, int, float*, float*, int64_t, void*) { return fail(HMZ_ERR_UNSUPPORTED, "not built yet"); }
int hmz_net_recurrent(const void*, int, const void*, int64_t, const uint16_t*, const uint8_t*, void*, int64_t, int64_t, int, float*, float*, float*, int64_t, void*) { return fail(HMZ_ERR_UNSUPPORTED, "not built yet"); }
}
