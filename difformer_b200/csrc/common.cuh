// Shared helpers for the difformer_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

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
// ---- peer-mapped exchange buffers (comm.cu, simple_sm100.cu): [2 data slots | flags [2][16][256] u64 | status u64]
constexpr int kCommMaxRanks = 16;
constexpr unsigned long long kCommTimeoutNs = 30000000000ull;         // 30 s without a peer's flag => give up
__host__ __device__ __forceinline__ unsigned long long* comm_status_ptr(float* base, int64_t slot_floats) {
    return reinterpret_cast<unsigned long long*>(base + 2 * slot_floats) + (size_t)2 * kCommMaxRanks * 256;
}
// spin until *flag == seq (system scope); bounded: on timeout (or if an earlier wait already timed out) set *status
__device__ __forceinline__ void comm_wait_flag(const unsigned long long* flag, unsigned long long seq, unsigned long long* status) {
    unsigned long long got, t0 = 0;
    unsigned int spins = 0;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(flag) : "memory");
        if (got == seq) return;
        if ((++spins & 1023u) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            if (now - t0 > kCommTimeoutNs || *reinterpret_cast<volatile unsigned long long*>(status) != 0) {
                atomicMax(status, 1ull);
                return;
            }
        }
    }
}
// sigmoid_sm100.cu
bool sigmoid_tc_supported(int64_t N, int64_t L, int H, int Hv, int M, int D);
int sigmoid_tc_ksplit(int64_t N, int64_t L, int H);
int64_t sigmoid_tc_image_bytes(int64_t L, int H, int Hv);
int sigmoid_fwd_tc(const float* q, const float* k, const float* v, int64_t N, int64_t L, int H, int Hv, float* out, float* rowsum,
                   float* pout, float* prs, int ksplit, void* images, cudaStream_t st);
int64_t simple_tc_workspace_bytes(int64_t N, int H, int Hv, int M, int D);
int64_t simple_tc_prepared_bytes(int H, int Hv, int M, int D);
int simple_reduce_tc(const float* q, const float* k, const float* v, int64_t N, int H, int Hv, int M, int D,
                     float* partials, void* prepared, void* ws, int64_t ws_bytes, cudaStream_t st,
                     void* const* peer_bufs = nullptr, int rank = 0, int world = 1, unsigned long long seq = 0, float* vbar = nullptr);
int simple_apply_tc(const float* q, const float* partials, const void* prepared, double n_total, int64_t N, int H, int Hv, int M, int D,
                    float* out, const dif_epilogue_t* ep, cudaStream_t st);

int64_t simple_tc_rowscal_floats(int64_t N, int H);
int simple_bwd_reduce_tc(const float* q, const float* g, const float* out, const float* partials, double n_total,
                         int64_t N, int H, float* bwd_partials, float* rowscal, void* ws, int64_t ws_bytes, cudaStream_t st);
int simple_bwd_apply_tc(const float* q, const float* k, const float* v, const float* g, const float* partials, const float* bwd_partials,
                        const float* rowscal, int64_t N, int H, float* dq, float* dk, float* dv, cudaStream_t st);
void simple_bwd_scalars(const float* partials, float* bwd_partials, int H, int Hv, int M, int D, cudaStream_t st);

}  // namespace dif
