// placeholder — replaced by the tree-store kernels
#include "hmz_common.cuh"
using namespace hmz;
extern "C" {
int hmz_search_minmax_reset(double*, int64_t, void*) { return fail(HMZ_ERR_UNSUPPORTED, "not built yet"); }
int hmz_search_begin(const hmz_search_t*, const double*, void*) { return fail(HMZ_ERR_UNSUPPORTED, "not built yet"); }
int hmz_search_select(const hmz_search_t*, int, const double*, double, uint16_t*, uint8_t*, uint16_t*, uint8_t*, int, void*) { return fail(HMZ_ERR_UNSUPPORTED, "not built yet"); }
int hmz_search_expand_backup(const hmz_search_t*, int, double, const uint16_t*, const uint8_t*, const float*, const float*, const float*, void*) { return fail(HMZ_ERR_UNSUPPORTED, "not built yet"); }
int hmz_search_root_policy(const hmz_search_t*, int, double, int, const double*, int32_t*, double*, double*, int32_t*, void*) { return fail(HMZ_ERR_UNSUPPORTED, "not built yet"); }
int hmz_search_run(const hmz_search_t*, const void*, int, int, const double*, double, void*) { return fail(HMZ_ERR_UNSUPPORTED, "not built yet"); }
}
