// One-shot NVLink all-reduce of the pass-1 partials: the only collective of the path.
//
// Payload: dif_simple_partials_len() floats (67.6 KB at H=4, D=64), independent of N => latency bound.
// Every rank owns a peer-mapped buffer [2 data slots | flag slots]; pass 1 writes its partials straight
// into the local data slot.  The kernel (a) publishes "slot ready" flags into every peer's buffer with
// system-scope release stores over NVLink, (b) waits for all peers' flags, (c) sums the G slots in rank
// order reading peer memory directly (NVSwitch: every peer at full bandwidth, 7 x 68 KB per rank).  Same
// summation order on every rank => bit-identical results, no NCCL launch, no host sync.  Slots alternate
// by sequence number; a slot is only rewritten two calls later, after every peer has signalled the call
// in between (which it sends after finishing its reads).
#include <string.h>

#include "common.cuh"

namespace dif {
namespace {

constexpr int kMaxRanks = kCommMaxRanks;

struct CommArgs {
    float* bufs[kMaxRanks];   // peer-mapped base pointers, index = rank
    int rank, world;
    int64_t len;              // floats
    int64_t slot_floats;      // data slot stride (floats)
    unsigned long long seq;
    float* out;
};

// flags: [2 slots][kMaxRanks][256 slices] u64 after the two data slots, then one u64 status word.  The fused pass-1
// tail (simple_sm100.cu) uses one flag per (slot, rank, column slice); this stand-alone kernel uses slice 255.
// Watchdog: a wait that sees no flag for kCommTimeoutNs gives up, sets the status word of the LOCAL buffer and lets
// the kernel finish with a meaningless sum, so that a peer that died or never launched cannot hang the GPU;
// dif_comm_status() reports it to the host.
__device__ __forceinline__ unsigned long long* flag_ptr(float* base, int64_t slot_floats, int idx) {
    return reinterpret_cast<unsigned long long*>(base + 2 * slot_floats) + (size_t)idx * 256 + 255;
}

__global__ void __launch_bounds__(256) allreduce_kernel(CommArgs a) {
    const int slot = (int)(a.seq & 1);
    // (a) one CTA signals: flag[slot][my_rank] := seq in every peer's buffer (and my own)
    if (blockIdx.x == 0 && threadIdx.x < a.world) {
        __threadfence_system();
        unsigned long long* f = flag_ptr(a.bufs[threadIdx.x], a.slot_floats, slot * kMaxRanks + a.rank);
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(f), "l"(a.seq) : "memory");
    }
    // (b) every CTA waits for all ranks' flags in the local buffer
    if (threadIdx.x < a.world) {
        const unsigned long long* f = flag_ptr(a.bufs[a.rank], a.slot_floats, slot * kMaxRanks + threadIdx.x);
        unsigned long long* status = comm_status_ptr(a.bufs[a.rank], a.slot_floats);
        comm_wait_flag(f, a.seq, status);
    }
    __syncthreads();
    // (c) sum the slots in rank order (fixed order: every rank computes the same bits)
    const int64_t n4 = a.len >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < a.world; ++r) {
            float4 x;   // system-scope load: peer memory over NVLink, never served from a stale line
            asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                         : "l"(a.bufs[r] + slot * a.slot_floats + 4 * i) : "memory");
            s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
        }
        *reinterpret_cast<float4*>(a.out + 4 * i) = s;
    }
    if (blockIdx.x == 0)
        for (int64_t i = 4 * n4 + threadIdx.x; i < a.len; i += blockDim.x) {
            float s = 0.f;
            for (int r = 0; r < a.world; ++r) {
                float x;
                asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(x) : "l"(a.bufs[r] + slot * a.slot_floats + i) : "memory");
                s += x;
            }
            a.out[i] = s;
        }
}

}  // namespace
}  // namespace dif

using namespace dif;

extern "C" int64_t dif_comm_buffer_bytes(int64_t len) {
    const int64_t slot = (len + 63) & ~(int64_t)63;
    return 2 * slot * (int64_t)sizeof(float) + 2 * kMaxRanks * 256 * (int64_t)sizeof(unsigned long long) + 64;   // + status word
}

extern "C" int dif_comm_status(const void* own_buf, int64_t len, int* timed_out) {
    DIF_REQUIRE(own_buf && timed_out && len > 0, DIF_EARG, "comm_status: bad argument");
    const int64_t slot = (len + 63) & ~(int64_t)63;
    unsigned long long v = 0;
    DIF_CUDA_OK(cudaMemcpy(&v, comm_status_ptr((float*)own_buf, slot), sizeof(v), cudaMemcpyDeviceToHost));
    *timed_out = v != 0;
    return DIF_OK;
}

extern "C" int64_t dif_comm_slot_offset_bytes(int64_t len, unsigned long long seq) {
    const int64_t slot = (len + 63) & ~(int64_t)63;
    return (int64_t)(seq & 1) * slot * (int64_t)sizeof(float);
}

extern "C" int dif_comm_alloc(void** ptr, int64_t bytes) {
    DIF_REQUIRE(ptr && bytes > 0, DIF_EARG, "comm_alloc: bad argument");
    DIF_CUDA_OK(cudaMalloc(ptr, (size_t)bytes));
    DIF_CUDA_OK(cudaMemset(*ptr, 0, (size_t)bytes));
    DIF_CUDA_OK(cudaDeviceSynchronize());
    return DIF_OK;
}

extern "C" int dif_comm_free(void* ptr) {
    if (ptr) DIF_CUDA_OK(cudaFree(ptr));
    return DIF_OK;
}

extern "C" int dif_comm_export(void* ptr, void* handle64) {
    DIF_REQUIRE(ptr && handle64, DIF_EARG, "comm_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DIF_CUDA_OK(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr));
    return DIF_OK;
}

extern "C" int dif_comm_open(const void* handle64, void** peer_ptr) {
    DIF_REQUIRE(handle64 && peer_ptr, DIF_EARG, "comm_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    DIF_CUDA_OK(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DIF_OK;
}

extern "C" int dif_comm_close(void* peer_ptr) {
    if (peer_ptr) DIF_CUDA_OK(cudaIpcCloseMemHandle(peer_ptr));
    return DIF_OK;
}

extern "C" int dif_comm_allreduce(void* const* bufs, int rank, int world, int64_t len, unsigned long long seq, float* out, void* stream) {
    DIF_REQUIRE(bufs && out && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world && len > 0 && seq > 0, DIF_EARG,
                "comm_allreduce: bad argument (world <= %d, seq >= 1)", kMaxRanks);
    CommArgs a{};
    for (int r = 0; r < world; ++r) { DIF_REQUIRE(bufs[r], DIF_EARG, "comm_allreduce: null peer buffer %d", r); a.bufs[r] = (float*)bufs[r]; }
    a.rank = rank; a.world = world; a.len = len; a.slot_floats = (len + 63) & ~(int64_t)63; a.seq = seq; a.out = out;
    const int64_t n4 = len / 4;
    int grid = (int)((n4 + 255) / 256);
    if (grid < 1) grid = 1;
    if (grid > 32) grid = 32;
    allreduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    DIF_LAUNCH_OK();
    return DIF_OK;
}
