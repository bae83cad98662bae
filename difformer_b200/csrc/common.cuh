// Shared helpers for the difformer_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/difformer_b200.h"

namespace dif {

int set_error(int code, const char* fmt, ...);
int cuda_error(cudaError_t e, const char* file, int line);

#define DIF_CUDA_OK(expr)                                                  \
    do {                                                                   \
        cudaError_t _e = (expr);                                           \
        if (_e != cudaSuccess) return dif::cuda_error(_e, __FILE__, __LINE__); \
    } while (0)

#define DIF_LAUNCH_OK() DIF_CUDA_OK(cudaGetLastError())

#define DIF_REQUIRE(cond, code, ...)                       \
    do {                                                   \
        if (!(cond)) return dif::set_error(code, __VA_ARGS__); \
    } while (0)

// ---- partial-buffer layouts (floats) --------------------------------------------------------
struct SimpleLayout {
    int H, Hv, M, D;
    __host__ __device__ int64_t offS() const { return 0; }
    __host__ __device__ int64_t offZ() const { return (int64_t)H * M * D; }
    __host__ __device__ int64_t offU() const { return offZ() + (int64_t)H * M; }
    __host__ __device__ int64_t offSq() const { return offU() + (int64_t)Hv * D; }
    __host__ __device__ int64_t offSk() const { return offSq() + 1; }
    __host__ __device__ int64_t len() const { return offSk() + 1; }
    // per-chunk workspace record of the generic reduce: sq/sk get one slot per head
    __host__ __device__ int64_t wsLen() const { return (offSq() + 2 * (int64_t)H + 3) & ~(int64_t)3; }   // float4-aligned records
};

struct BwdLayout {
    int H, M, D;
    __host__ __device__ int64_t offS() const { return 0; }
    __host__ __device__ int64_t offZ() const { return (int64_t)H * M * D; }
    __host__ __device__ int64_t offU() const { return offZ() + (int64_t)H * M; }
    __host__ __device__ int64_t offTq() const { return offU() + (int64_t)H * D; }
    __host__ __device__ int64_t offTk() const { return offTq() + 1; }
    __host__ __device__ int64_t len() const { return offTk() + 1; }
    __host__ __device__ int64_t wsLen() const { return (offTq() + (int64_t)H + 3) & ~(int64_t)3; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum, result valid in every thread; `red` is >= 33 floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nw ? red[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// ---- implemented in the per-feature translation units ----------------------------------------
int simple_reduce_generic(const float* q, const float* k, const float* v, int64_t N, int H, int Hv, int M, int D,
                          float* partials, void* ws, int64_t ws_bytes, cudaStream_t st);
int simple_apply_generic(const float* q, const float* partials, double n_total, int64_t N, int H, int Hv, int M, int D,
                         float* out, const dif_epilogue_t* ep, cudaStream_t st);
int64_t simple_generic_workspace_bytes(int64_t N, int H, int Hv, int M, int D);

bool simple_tc_supported(int64_t N, int H, int Hv, int M, int D);
// ---- peer-mapped exchange buffers (comm.cu, simple_sm100.cu) ----------------------------------------------------
// Layout of one rank's buffer: [header 128 B: status u64 | device pointer of this rank's pinned host flag u64]
//                              [LL region: u64 [2 slots][kCommMaxRanks source ranks][lenpad]]
// LL ("low latency") protocol: a sender writes 64-bit words {tag = call number (32 bits) | fp32 payload} straight
// into the receiver's LL region with one scalar store per word (single-copy atomic), the receiver polls the word
// until the tag matches.  Data and flag travel together: no fence, no flag round trip, one NVLink traversal.
// Slots alternate by call number; a slot is rewritten two calls later, which a peer can only reach after this
// rank has sent (i.e. finished reading for) the call in between.
constexpr int kCommMaxRanks = 16;
constexpr int kCommHeaderBytes = 128;
constexpr unsigned long long kCommTimeoutNs = 30000000000ull;         // default: 30 s without a peer's word => give up
__host__ __device__ __forceinline__ int64_t comm_lenpad(int64_t len) { return (len + 63) & ~(int64_t)63; }
__host__ __device__ __forceinline__ unsigned long long* comm_status_ptr(void* base) { return reinterpret_cast<unsigned long long*>(base); }
__host__ __device__ __forceinline__ unsigned long long* comm_ll_ptr(void* base, int64_t lenpad, int slot, int src_rank) {
    return reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(base) + kCommHeaderBytes) +
           ((size_t)slot * kCommMaxRanks + (size_t)src_rank) * (size_t)lenpad;
}
struct CommPeers {            // kernel argument: every rank's buffer as mapped in this process
    void* bufs[kCommMaxRanks];
    int rank, world;
    unsigned long long seq;   // call number 1, 2, 3, ... identical on all ranks
    int64_t lenpad;
    unsigned long long timeout_ns;   // watchdog bound (comm_timeout_ns(): DIF_COMM_TIMEOUT_MS, default 30 s)
};
inline unsigned long long comm_timeout_ns() {
    static unsigned long long v = 0;
    if (v == 0) {
        const char* e = getenv("DIF_COMM_TIMEOUT_MS");
        const long long ms = e ? atoll(e) : 0;
        v = ms > 0 ? (unsigned long long)ms * 1000000ull : kCommTimeoutNs;
    }
    return v;
}
// A wait gave up (or found the status word already set): mark EVERY rank's status word (a peer that is not waiting right
// now sees it at its next call) and raise this rank's pinned host flag (read by the host without a device sync).
static __device__ __noinline__ void comm_fail(const CommPeers& c) {
    for (int r = 0; r < c.world; ++r)
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(comm_status_ptr(c.bufs[r])), "l"(1ull) : "memory");
    unsigned long long* host = *reinterpret_cast<unsigned long long* volatile*>(reinterpret_cast<char*>(c.bufs[c.rank]) + 8);
    if (host) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(host), "l"(1ull) : "memory");
}
// one thread per call: a peer reported a failure since the last call -> surface it on this rank's host flag too
__device__ __forceinline__ void comm_check_status(const CommPeers& c) {
    unsigned long long s;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(s) : "l"(comm_status_ptr(c.bufs[c.rank])) : "memory");
    if (s != 0) comm_fail(c);
}
__device__ __forceinline__ void comm_ll_send(unsigned long long* dst, float x, uint32_t tag) {
    const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(x);
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(dst), "l"(w) : "memory");
}
// poll one LL word until its tag matches; bounded (watchdog): on timeout the result is meaningless and comm_fail() has run
__device__ __forceinline__ float comm_ll_recv(const unsigned long long* src, uint32_t tag, const CommPeers& c) {
    unsigned long long w, t0 = 0;
    unsigned int spins = 0;
    for (;;) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
        if ((uint32_t)(w >> 32) == tag) return __uint_as_float((uint32_t)w);
        if ((++spins & 1023u) == 0) {
            unsigned long long now, s;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(s) : "l"(comm_status_ptr(c.bufs[c.rank])) : "memory");
            if (now - t0 > c.timeout_ns || s != 0) { comm_fail(c); return 0.f; }
        }
    }
}
// Receive element j of every peer's contribution (LL words of call `tag` in this rank's own buffer) and return the sum over ALL ranks
// in rank order, `mine` being this rank's term.  The pending words are loaded back to back each polling round (one L2 round trip
// per round, not one per peer); bounded like comm_ll_recv.
static __device__ __noinline__ float comm_ll_sum(const CommPeers& c, int slot, int64_t j, uint32_t tag, float mine) {
    float v[kCommMaxRanks];
    unsigned int pending = 0;
    for (int r = 0; r < c.world; ++r)
        if (r != c.rank) pending |= 1u << r;
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    while (pending) {
        unsigned long long w[kCommMaxRanks];
#pragma unroll
        for (int r = 0; r < kCommMaxRanks; ++r)
            if (pending & (1u << r))
                asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w[r]) : "l"(comm_ll_ptr(c.bufs[c.rank], c.lenpad, slot, r) + j) : "memory");
#pragma unroll
        for (int r = 0; r < kCommMaxRanks; ++r)
            if ((pending & (1u << r)) && (uint32_t)(w[r] >> 32) == tag) { v[r] = __uint_as_float((uint32_t)w[r]); pending &= ~(1u << r); }
        if (pending && (++spins & 1023u) == 0) {
            unsigned long long now, s;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(s) : "l"(comm_status_ptr(c.bufs[c.rank])) : "memory");
            if (now - t0 > c.timeout_ns || s != 0) { comm_fail(c); return 0.f; }
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < kCommMaxRanks; ++r)
        if (r < c.world) sum += r == c.rank ? mine : v[r];
    return sum;
}

// sigmoid_sm100.cu
bool sigmoid_tc_supported(int64_t N, int64_t L, int H, int Hv, int M, int D);
int sigmoid_tc_ksplit(int64_t N, int64_t L, int H);
int64_t sigmoid_tc_image_bytes(int64_t L, int H, int Hv);
int sigmoid_fwd_tc(const float* q, const float* k, const float* v, int64_t N, int64_t L, int H, int Hv, float* out, float* rowsum,
                   float* pout, float* prs, int ksplit, void* images, cudaStream_t st);
// sigmoid_bwd_sm100.cu
bool sigmoid_bwd_tc_supported(int64_t N, int64_t L, int H, int Hv, int M, int D);
int sigmoid_bwd_tc_split(int64_t own_rows, int64_t streamed_rows, int H);
int64_t sigmoid_bwd_tc_image_bytes(int64_t N, int64_t L, int H);
int sigmoid_bwd_tc(const float* q, const float* k, const float* v, const float* g, const float* drow, const float* rowsum,
                   int64_t N, int64_t L, int H, float* dq, float* dk, float* dv, void* images, int split_k, float* part_dq,
                   int split_q, float* part_dk, float* part_dv, cudaStream_t st);
int64_t simple_tc_workspace_bytes(int64_t N, int H, int Hv, int M, int D);
int64_t simple_tc_prepared_bytes(int H, int Hv, int M, int D);
int simple_reduce_tc(const float* q, const float* k, const float* v, int64_t N, int H, int Hv, int M, int D,
                     float* partials, void* prepared, void* ws, int64_t ws_bytes, cudaStream_t st,
                     void* const* peer_bufs = nullptr, int rank = 0, int world = 1, unsigned long long seq = 0, float* vbar = nullptr);
int64_t segmented_plan_bytes(int64_t N, int max_nodes);
int segmented_plan_build(const int32_t* seg_ptr, int B, int64_t N, int max_nodes, void* plan, cudaStream_t st);
int segmented_fwd_tc(const float* q, const float* k, const float* v, const void* plan, int64_t N, int max_nodes, const float* norms, float* out,
                     cudaStream_t st);
int segmented_bwd_tc(const float* q, const float* k, const float* v, const float* g, const float* out, const void* plan, int64_t N, int max_nodes,
                     const float* norms, float* dq, float* dk, float* dv, float* part, float* scal, int phase, cudaStream_t st);
int simple_project_values(const float* Wv, const float* bv, int H, float* vbar_partials, float* one, cudaStream_t st);
int64_t simple_project_workspace_bytes(int H);
int simple_project(const float* gram, const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wv, const float* bv,
                   double n_total, int H, float* vpartials, float* nvec, float* vbar_partials, void* ws, cudaStream_t st);
int simple_apply_tc(const float* q, const float* partials, const void* prepared, double n_total, int64_t N, int H, int Hv, int M, int D,
                    float* out, const dif_epilogue_t* ep, cudaStream_t st, int64_t q_ld = 0, int q_hs = 0, const float* nvec = nullptr);

bool simple_wide_supported(int64_t N, int H, int Hv, int M, int D);
int64_t simple_fused_workspace_bytes(int64_t N, int H, int Hv, int M, int D);
int simple_forward_tc(const float* q, const float* k, const float* v, int64_t N, int H, int Hv, int M, int D, double n_total,
                      float* partials, float* out, void* ws, int64_t ws_bytes, cudaStream_t st,
                      void* const* peer_bufs = nullptr, int rank = 0, int world = 1, unsigned long long seq = 0);

int simple_forward_lp(const void* q, const void* k, const void* v, int dtype, int64_t N, int H, int Hv, int M, int D, double n_total,
                      float* partials, void* out, void* ws, int64_t ws_bytes, cudaStream_t st,
                      void* const* peer_bufs = nullptr, int rank = 0, int world = 1, unsigned long long seq = 0);

int64_t simple_tc_rowscal_floats(int64_t N, int H);
int simple_bwd_reduce_tc(const float* q, const float* g, const float* out, const float* partials, double n_total,
                         int64_t N, int H, float* bwd_partials, float* rowscal, void* ws, int64_t ws_bytes, cudaStream_t st);
int simple_bwd_apply_tc(const float* q, const float* k, const float* v, const float* g, const float* partials, const float* bwd_partials,
                        const float* rowscal, int64_t N, int H, float* dq, float* dk, float* dv, cudaStream_t st);
void simple_bwd_scalars(const float* partials, float* bwd_partials, int H, int Hv, int M, int D, cudaStream_t st);

}  // namespace dif
