// One-shot NVLink all-reduce of the pass-1 partials: the only collective of the path.
//
// Payload: dif_simple_partials_len() floats (67.6 KB at H=4, D=64), independent of N => latency bound.
// Every rank owns a peer-mapped buffer [header | LL region] (layout and protocol: common.cuh).  A rank PUSHES each
// element of its contribution into every peer's region as one 64-bit word {call number | fp32} and polls the words the
// peers pushed into its own region: one NVLink traversal, no fences, no flag round trip, no NCCL launch, no host sync.
// Ranks are summed in rank order => bit-identical results everywhere.  The same protocol runs inside the tail of the
// pass-1 kernel (simple_sm100.cu: compute + collective in one kernel); this file holds the stand-alone kernel, the
// buffer management and the watchdog plumbing (a wait that sees nothing for 30 s gives up, marks every rank's status
// word and raises a pinned host flag that dif_comm_status() reads without synchronising).
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace dif {
namespace {

struct CommArgs {
    CommPeers peers;
    int64_t len;              // floats
    const float* src;         // this rank's contribution (any device memory)
    float* out;               // sum over ranks (local)
};

// Stand-alone one-shot all-reduce (backward partials, shapes on the generic path): same LL push protocol as the fused
// pass-1 tail.  Thread i owns element i: push it to every peer, poll the peers' words in the local region, add in rank
// order (every rank computes the same bits).
__global__ void __launch_bounds__(256) allreduce_kernel(const __grid_constant__ CommArgs a) {
    const CommPeers& c = a.peers;
    const int slot = (int)(c.seq & 1);
    const uint32_t tag = (uint32_t)c.seq;
    if (blockIdx.x == 0 && threadIdx.x == 0) comm_check_status(c);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.len; i += (int64_t)gridDim.x * blockDim.x) {
        const float mine = a.src[i];
        for (int p = 1; p < c.world; ++p) {
            int r = c.rank + p;
            if (r >= c.world) r -= c.world;
            comm_ll_send(comm_ll_ptr(c.bufs[r], c.lenpad, slot, c.rank) + i, mine, tag);
        }
        a.out[i] = comm_ll_sum(c, slot, i, tag, mine);
    }
}

// device buffer -> pinned, mapped host flag (written by comm_fail, read by dif_comm_status without a device sync)
std::mutex g_mu;
std::unordered_map<void*, unsigned long long*> g_host_flags;

}  // namespace
}  // namespace dif

using namespace dif;

extern "C" int64_t dif_comm_buffer_bytes(int64_t len) {
    return (int64_t)kCommHeaderBytes + 2 * (int64_t)kCommMaxRanks * comm_lenpad(len) * (int64_t)sizeof(unsigned long long);
}

extern "C" int dif_comm_status(const void* own_buf, int* timed_out) {
    DIF_REQUIRE(own_buf && timed_out, DIF_EARG, "comm_status: bad argument");
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_host_flags.find(const_cast<void*>(own_buf));
    DIF_REQUIRE(it != g_host_flags.end(), DIF_EARG, "comm_status: not a dif_comm_alloc buffer of this process");
    *timed_out = *reinterpret_cast<volatile unsigned long long*>(it->second) != 0;
    return DIF_OK;
}

extern "C" int dif_comm_reset(void* own_buf) {
    DIF_REQUIRE(own_buf, DIF_EARG, "comm_reset: null pointer");
    unsigned long long* host = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_host_flags.find(own_buf);
        DIF_REQUIRE(it != g_host_flags.end(), DIF_EARG, "comm_reset: not a dif_comm_alloc buffer of this process");
        host = it->second;
    }
    DIF_CUDA_OK(cudaDeviceSynchronize());
    DIF_CUDA_OK(cudaMemset(own_buf, 0, 8));          // status word only: the LL words keep their (stale) tags
    DIF_CUDA_OK(cudaDeviceSynchronize());
    *reinterpret_cast<volatile unsigned long long*>(host) = 0;
    return DIF_OK;
}

extern "C" int dif_comm_alloc(void** ptr, int64_t bytes) {
    DIF_REQUIRE(ptr && bytes >= kCommHeaderBytes, DIF_EARG, "comm_alloc: bad argument");
    DIF_CUDA_OK(cudaMalloc(ptr, (size_t)bytes));
    DIF_CUDA_OK(cudaMemset(*ptr, 0, (size_t)bytes));
    unsigned long long *host = nullptr, *host_dev = nullptr;
    DIF_CUDA_OK(cudaHostAlloc((void**)&host, 64, cudaHostAllocMapped));
    *host = 0;
    DIF_CUDA_OK(cudaHostGetDevicePointer((void**)&host_dev, host, 0));
    DIF_CUDA_OK(cudaMemcpy(reinterpret_cast<char*>(*ptr) + 8, &host_dev, sizeof(host_dev), cudaMemcpyHostToDevice));
    DIF_CUDA_OK(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_mu);
    g_host_flags[*ptr] = host;
    return DIF_OK;
}

extern "C" int dif_comm_free(void* ptr) {
    if (!ptr) return DIF_OK;
    unsigned long long* host = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_host_flags.find(ptr);
        if (it != g_host_flags.end()) { host = it->second; g_host_flags.erase(it); }
    }
    DIF_CUDA_OK(cudaFree(ptr));
    if (host) DIF_CUDA_OK(cudaFreeHost(host));
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

extern "C" int dif_comm_allreduce(void* const* bufs, int rank, int world, int64_t len, unsigned long long seq,
                                  const float* src, float* out, void* stream) {
    DIF_REQUIRE(bufs && src && out && world >= 1 && world <= kCommMaxRanks && rank >= 0 && rank < world && len > 0 && seq > 0, DIF_EARG,
                "comm_allreduce: bad argument (world <= %d, seq >= 1)", kCommMaxRanks);
    CommArgs a{};
    for (int r = 0; r < world; ++r) { DIF_REQUIRE(bufs[r], DIF_EARG, "comm_allreduce: null peer buffer %d", r); a.peers.bufs[r] = bufs[r]; }
    a.peers.rank = rank; a.peers.world = world; a.peers.seq = seq; a.peers.lenpad = comm_lenpad(len); a.peers.timeout_ns = comm_timeout_ns();
    a.len = len; a.src = src; a.out = out;
    int grid = (int)((len + 255) / 256);
    if (grid < 1) grid = 1;
    if (grid > 74) grid = 74;
    allreduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    DIF_LAUNCH_OK();
    return DIF_OK;
}
