// C ABI glue: error reporting, capability query and the 'simple' forward dispatch
// (tcgen05 path when the shape qualifies, generic FFMA path otherwise).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace dif {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_error(cudaError_t e, const char* file, int line) {
    const char* base = strrchr(file, '/');
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d", (int)e, cudaGetErrorString(e), base ? base + 1 : file, line);
    (void)cudaGetLastError();   // clear the sticky-free error so later calls can proceed
    return DIF_ECUDA;
}

}  // namespace dif

using namespace dif;

extern "C" int dif_version(void) { return 100; }   // 0.1.0

extern "C" const char* dif_last_error(void) { return g_err; }

extern "C" int dif_device_supported(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return major == 10 ? 1 : 0;
}

extern "C" int64_t dif_simple_partials_len(int H, int Hv, int M, int D) { return SimpleLayout{H, Hv, M, D}.len(); }

extern "C" int64_t dif_simple_workspace_bytes(int64_t N, int H, int Hv, int M, int D) {
    if (N < 1 || H < 1 || M < 1 || D < 1) return -1;
    int64_t a = simple_generic_workspace_bytes(N, H, Hv, M, D);
    int64_t b = simple_tc_supported(N, H, Hv, M, D) ? simple_tc_workspace_bytes(N, H, Hv, M, D) : 0;
    return a > b ? a : b;
}

static int pick_impl(int impl, int64_t N, int H, int Hv, int M, int D, bool* use_tc) {
    const bool ok = simple_tc_supported(N, H, Hv, M, D);
    if (impl == DIF_IMPL_AUTO) { *use_tc = ok; return DIF_OK; }
    if (impl == DIF_IMPL_GENERIC) { *use_tc = false; return DIF_OK; }
    if (impl == DIF_IMPL_TCGEN05) {
        DIF_REQUIRE(ok, DIF_EUNSUPPORTED, "simple: tcgen05 path needs M == D == 64, Hv == H, H in {1, 2, 4} (got H=%d Hv=%d M=%d D=%d)", H, Hv, M, D);
        *use_tc = true;
        return DIF_OK;
    }
    return set_error(DIF_EARG, "simple: unknown impl %d", impl);
}

extern "C" int64_t dif_simple_prepared_bytes(int H, int Hv, int M, int D) { return simple_tc_prepared_bytes(H, Hv, M, D); }

extern "C" int dif_simple_reduce(const float* q, const float* k, const float* v, int64_t N, int H, int Hv, int M, int D,
                                 float* partials, void* prepared, float* vbar, void* workspace, int64_t workspace_bytes, int impl, void* stream) {
    DIF_REQUIRE(q && k && v && partials && workspace, DIF_EARG, "simple_reduce: null pointer");
    bool tc = false;
    int rc = pick_impl(impl, N, H, Hv, M, D, &tc);
    if (rc) return rc;
    if (tc)
        return simple_reduce_tc(q, k, v, N, H, Hv, M, D, partials, prepared, workspace, workspace_bytes, (cudaStream_t)stream,
                                nullptr, 0, 1, 0, vbar);
    rc = simple_reduce_generic(q, k, v, N, H, Hv, M, D, partials, workspace, workspace_bytes, (cudaStream_t)stream);
    if (rc == DIF_OK && vbar != nullptr) rc = dif_head_mean(v, N, Hv, D, vbar, stream);      // generic path: separate kernel
    return rc;
}

// pass 1 + the cross-GPU all-reduce of the partials in ONE kernel (tcgen05 shapes only): the tail of the
// reduce kernel exchanges its column slices with the peers over NVLink (peer-mapped dif_comm_* buffers).
extern "C" int dif_simple_reduce_allreduce(const float* q, const float* k, const float* v, int64_t N, int H, int Hv, int M, int D,
                                           float* partials, void* prepared, void* workspace, int64_t workspace_bytes,
                                           void* const* peer_bufs, int rank, int world, unsigned long long seq, void* stream) {
    DIF_REQUIRE(q && k && v && partials && workspace && peer_bufs, DIF_EARG, "simple_reduce_allreduce: null pointer");
    DIF_REQUIRE(simple_tc_supported(N, H, Hv, M, D), DIF_EUNSUPPORTED,
                "simple_reduce_allreduce: fused exchange needs the tcgen05 shapes (M == D == 64, Hv == H == 4); all-reduce separately otherwise");
    return simple_reduce_tc(q, k, v, N, H, Hv, M, D, partials, prepared, workspace, workspace_bytes, (cudaStream_t)stream,
                            peer_bufs, rank, world, seq);
}

// The whole forward in ONE cooperative kernel (tcgen05 shapes): pass 1 -> grid-wide (and, with peer buffers, cross-GPU) sum
// -> pass 2.  `partials` receives the reduced partials (saved for the backward).
extern "C" int64_t dif_simple_forward_workspace_bytes(int64_t N, int H, int Hv, int M, int D) {
    if (N < 1 || H < 1 || M < 1 || D < 1) return -1;
    return simple_fused_workspace_bytes(N, H, Hv, M, D);
}

extern "C" int dif_simple_forward(const void* q, const void* k, const void* v, int dtype, int64_t N, int H, int Hv, int M, int D, double n_total,
                                  float* partials, void* out, void* workspace, int64_t workspace_bytes,
                                  void* const* peer_bufs, int rank, int world, unsigned long long seq, void* stream) {
    DIF_REQUIRE(q && k && v && partials && out && workspace, DIF_EARG, "simple_forward: null pointer");
    DIF_REQUIRE(n_total > 0, DIF_EARG, "simple_forward: n_total must be positive");
    DIF_REQUIRE(simple_fused_workspace_bytes(N, H, Hv, M, D) > 0, DIF_EUNSUPPORTED,
                "simple_forward: the one-kernel forward needs M == D == 64, Hv == H, H in {1, 2, 4}, or one head of M == D == 128 (fp32); "
                "use dif_simple_reduce + dif_simple_apply");
    if (dtype == DIF_DTYPE_BF16 || dtype == DIF_DTYPE_F16)
        return simple_forward_lp(q, k, v, dtype, N, H, Hv, M, D, n_total, partials, out, workspace, workspace_bytes, (cudaStream_t)stream,
                                 peer_bufs, rank, world, seq);
    DIF_REQUIRE(dtype == DIF_DTYPE_F32, DIF_EARG, "simple_forward: unknown dtype %d", dtype);
    return simple_forward_tc((const float*)q, (const float*)k, (const float*)v, N, H, Hv, M, D, n_total, partials, (float*)out, workspace,
                             workspace_bytes, (cudaStream_t)stream, peer_bufs, rank, world, seq);
}

extern "C" int dif_simple_apply(const float* q, const float* partials, const void* prepared, double n_total, int64_t N, int H, int Hv, int M, int D,
                                float* out, const dif_epilogue_t* epilogue, int impl, void* stream) {
    DIF_REQUIRE(q && partials && out, DIF_EARG, "simple_apply: null pointer");
    DIF_REQUIRE(n_total > 0, DIF_EARG, "simple_apply: n_total must be positive");
    if (epilogue) {
        DIF_REQUIRE(epilogue->n_add >= 0 && epilogue->n_add <= 3, DIF_EARG, "simple_apply: n_add=%d", epilogue->n_add);
        for (int j = 0; j < epilogue->n_add; ++j) DIF_REQUIRE(epilogue->add[j], DIF_EARG, "simple_apply: addend %d is null", j);
    }
    bool tc = false;
    int rc = pick_impl(impl, N, H, Hv, M, D, &tc);
    if (rc) return rc;
    return tc ? simple_apply_tc(q, partials, prepared, n_total, N, H, Hv, M, D, out, epilogue, (cudaStream_t)stream)
              : simple_apply_generic(q, partials, n_total, N, H, Hv, M, D, out, epilogue, (cudaStream_t)stream);
}

// Pass 2 with the Wq projection folded into its operands (node classification/difformer.py:115-118): the A operand is
// the layer input x [N,64] itself (row stride ldx floats), shared by the H heads, and `vpartials` holds, in the partials layout of
// (H, Hv = H, M = 64, D = 64), the projected operands  S'_h = Wq_h^T S_h,  z'_h = Wq_h^T z_h,  u'_h = u_h + c bq_h^T S_h  and the two
// squared norms (sum q^2, sum k^2); n_total_vec[h] = N + c bq_h . z_h (device, H floats).  Then
//     out[n,h,:] = (c x_n S'_h + u'_h) / (c x_n . z'_h + n_total_vec[h])  =  full_attention_conv(x Wq^T + bq, K, V, 'simple')[n,h,:]
// without Q (nor K, V: their reductions follow from the Gram matrix X^T X, see difformer_b200/projected.py) ever being written.
extern "C" int dif_simple_apply_projected(const float* x, int64_t ldx, const float* vpartials, const float* n_total_vec, int64_t N, int H,
                                          float* out, const dif_epilogue_t* epilogue, void* stream) {
    DIF_REQUIRE(x && vpartials && n_total_vec && out, DIF_EARG, "simple_apply_projected: null pointer");
    DIF_REQUIRE(simple_tc_supported(N, H, H, 64, 64), DIF_EUNSUPPORTED, "simple_apply_projected: needs hidden = 64 and H in {1, 2, 4}");
    DIF_REQUIRE(ldx >= 64 && (ldx % 8) == 0, DIF_EARG, "simple_apply_projected: ldx must be >= 64 and a multiple of 8 floats");
    if (epilogue) {
        DIF_REQUIRE(epilogue->n_add >= 0 && epilogue->n_add <= 3, DIF_EARG, "simple_apply_projected: n_add=%d", epilogue->n_add);
        for (int j = 0; j < epilogue->n_add; ++j) DIF_REQUIRE(epilogue->add[j], DIF_EARG, "simple_apply_projected: addend %d is null", j);
    }
    return simple_apply_tc(x, vpartials, nullptr, 1.0, N, H, H, 64, 64, out, epilogue, (cudaStream_t)stream, ldx, 0, n_total_vec);
}

// The fp64 algebra between pass 1 on x (Gram matrix) and dif_simple_apply_projected (project.cu).
extern "C" int64_t dif_simple_project_workspace_bytes(int H) { return (H == 1 || H == 2 || H == 4) ? simple_project_workspace_bytes(H) : 0; }

extern "C" int dif_simple_project(const float* gram_partials, const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wv,
                                  const float* bv, double n_total, int H, float* vpartials, float* n_total_vec, float* vbar_partials,
                                  void* workspace, int64_t workspace_bytes, void* stream) {
    DIF_REQUIRE(gram_partials && Wq && bq && Wk && bk && vpartials && n_total_vec && workspace, DIF_EARG,
                "simple_project: null pointer");
    DIF_REQUIRE((Wv == nullptr) == (bv == nullptr), DIF_EARG, "simple_project: Wv and bv must both be given or both be null");
    DIF_REQUIRE(H == 1 || H == 2 || H == 4, DIF_EUNSUPPORTED, "simple_project: H=%d (needs H in {1, 2, 4}, hidden = 64)", H);
    DIF_REQUIRE(n_total > 0, DIF_EARG, "simple_project: n_total must be positive");
    DIF_REQUIRE(workspace_bytes >= simple_project_workspace_bytes(H) && (reinterpret_cast<uintptr_t>(workspace) & 7) == 0, DIF_EARG,
                "simple_project: workspace too small or not 8-byte aligned");
    return simple_project(gram_partials, Wq, bq, Wk, bk, Wv, bv, n_total, H, vpartials, n_total_vec, vbar_partials, workspace, (cudaStream_t)stream);
}

// Batched-graph 'simple' forward on the tensor cores (segmented_sm100.cu): plan + kernel.
extern "C" int64_t dif_segmented_plan_bytes(int64_t N, int max_nodes) { return segmented_plan_bytes(N, max_nodes); }

extern "C" int dif_segmented_plan_build(const int32_t* seg_ptr, int32_t B, int64_t N, int max_nodes, void* plan, int64_t plan_bytes, void* stream) {
    DIF_REQUIRE(seg_ptr && plan && B >= 1, DIF_EARG, "segmented_plan_build: bad argument");
    const int64_t need = segmented_plan_bytes(N, max_nodes);
    DIF_REQUIRE(need > 0, DIF_EUNSUPPORTED, "segmented_plan_build: needs 1 <= max_nodes <= 128 and 1 <= N < 2^31");
    DIF_REQUIRE(plan_bytes >= need && (reinterpret_cast<uintptr_t>(plan) & 15) == 0, DIF_EARG, "segmented_plan_build: plan buffer too small or misaligned");
    return segmented_plan_build(seg_ptr, B, N, max_nodes, plan, (cudaStream_t)stream);
}

extern "C" int dif_segmented_simple_fwd_tc(const float* q, const float* k, const float* v, const void* plan, int64_t plan_bytes, const float* norms,
                                           int64_t N, int max_nodes, float* out, void* stream) {
    DIF_REQUIRE(q && k && v && plan && norms && out, DIF_EARG, "segmented_fwd(tcgen05): null pointer");
    const int64_t need = segmented_plan_bytes(N, max_nodes);
    DIF_REQUIRE(need > 0, DIF_EUNSUPPORTED, "segmented_fwd(tcgen05): needs 1 <= max_nodes <= 128 and 1 <= N < 2^31");
    DIF_REQUIRE(plan_bytes >= need, DIF_EARG, "segmented_fwd(tcgen05): plan buffer too small");
    return segmented_fwd_tc(q, k, v, plan, N, max_nodes, norms, out, (cudaStream_t)stream);
}

extern "C" int dif_segmented_simple_bwd_tc(const float* q, const float* k, const float* v, const float* g, const float* out, const void* plan,
                                           int64_t plan_bytes, const float* norms, int64_t N, int max_nodes, int32_t B, float* dq, float* dk,
                                           float* dv, void* workspace, int64_t workspace_bytes, int phase, void* stream) {
    DIF_REQUIRE(q && k && v && g && out && plan && norms && dq && dk && dv && workspace, DIF_EARG, "segmented_bwd(tcgen05): null pointer");
    DIF_REQUIRE(phase >= 0 && phase <= 2 && B >= 1, DIF_EARG, "segmented_bwd(tcgen05): phase %d, B %d", phase, (int)B);
    const int64_t need = segmented_plan_bytes(N, max_nodes);
    DIF_REQUIRE(need > 0, DIF_EUNSUPPORTED, "segmented_bwd(tcgen05): needs 1 <= max_nodes <= 128 and 1 <= N < 2^31");
    DIF_REQUIRE(plan_bytes >= need, DIF_EARG, "segmented_bwd(tcgen05): plan buffer too small");
    DIF_REQUIRE(workspace_bytes >= dif_segmented_workspace_bytes(B) && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, DIF_EARG,
                "segmented_bwd(tcgen05): workspace too small (dif_segmented_workspace_bytes) or misaligned");
    float* scal = (float*)workspace + 2 * (int64_t)B;          // same place as dif_segmented_simple_bwd_phase: the caller all-reduces it between phases
    return segmented_bwd_tc(q, k, v, g, out, plan, N, max_nodes, norms, dq, dk, dv, scal + 2, scal, phase, (cudaStream_t)stream);
}

// mean_h V as a one-head pass-2 problem, from the weights alone (project.cu): the value branch of a folded layer does not wait for G.
extern "C" int dif_simple_project_values(const float* Wv, const float* bv, int H, float* vbar_partials, float* one, void* stream) {
    DIF_REQUIRE(vbar_partials && one, DIF_EARG, "simple_project_values: null pointer");
    DIF_REQUIRE((Wv == nullptr) == (bv == nullptr), DIF_EARG, "simple_project_values: Wv and bv must both be given or both be null");
    DIF_REQUIRE(H == 1 || H == 2 || H == 4, DIF_EUNSUPPORTED, "simple_project_values: H=%d (needs H in {1, 2, 4}, hidden = 64)", H);
    return simple_project_values(Wv, bv, H, vbar_partials, one, (cudaStream_t)stream);
}
