// kernel='simple' -- tcgen05 / TMEM path for sm_100a (placeholder until the kernels land).
#include "common.cuh"

namespace dif {

bool simple_tc_supported(int64_t, int, int, int, int) { return false; }
int64_t simple_tc_workspace_bytes(int64_t, int, int, int, int) { return 0; }
int simple_reduce_tc(const float*, const float*, const float*, int64_t, int, int, int, int, float*, void*, int64_t, cudaStream_t) {
    return set_error(DIF_EUNSUPPORTED, "tcgen05 path not built");
}
int simple_apply_tc(const float*, const float*, double, int64_t, int, int, int, int, float*, const dif_epilogue_t*, cudaStream_t) {
    return set_error(DIF_EUNSUPPORTED, "tcgen05 path not built");
}

}  // namespace dif
