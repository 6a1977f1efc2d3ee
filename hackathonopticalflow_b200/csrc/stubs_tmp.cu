// temporary: replaced by pyrlk.cu / gftt.cu / pathfinder.cu
#include "common.cuh"
namespace b2of {
size_t pyrlk_workspace_bytes(int, int, const b2of_lk_params*, int) { set_error("pyrlk not built"); return 0; }
int pyrlk_dev(const uint8_t*, const uint8_t*, size_t, size_t, int, int, int, const float*, size_t, int, float*,
              uint8_t*, float*, const b2of_lk_params*, void*, size_t, cudaStream_t) { return fail(B2OF_E_UNSUPPORTED, "pyrlk not built"); }
size_t gftt_workspace_bytes(int, int, const b2of_gftt_params*, int) { set_error("gftt not built"); return 0; }
int gftt_dev(const uint8_t*, const uint8_t*, size_t, size_t, int, int, int, const b2of_gftt_params*, float*, int, int*,
             void*, size_t, cudaStream_t) { return fail(B2OF_E_UNSUPPORTED, "gftt not built"); }
int pathfinder_filter_dev(const float*, size_t, const float*, int, int, int, int, int32_t*, int32_t*, uint8_t*,
                          uint8_t*, int32_t*, float*, cudaStream_t) { return fail(B2OF_E_UNSUPPORTED, "pathfinder not built"); }
int flow_stats_dev(const float*, int, int, int, float*, cudaStream_t) { return fail(B2OF_E_UNSUPPORTED, "stats not built"); }
}
