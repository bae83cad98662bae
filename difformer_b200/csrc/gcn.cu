// gcn_conv -- node classification/difformer.py:63-79 (identical copy: physical particle/difformer-v2.py:30-46).
//
//   d     = in-degree histogram of col (= torch_geometric.utils.degree(col, N), :66)
//   val_e = w_e * sqrt(1/d[col_e]) * sqrt(1/d[row_e])           (:67-73, both factors from the same d)
//   val_e = 0 where not finite                                   (:74)
//   out[c,h,:] = sum_{e: col_e = c} val_e * x[row_e,h,:]         (:75-78, duplicates summed)
//
// The reference rebuilds degree, norm and a sorted SparseTensor on every forward and runs one
// SpMM per head.  Here the CSR (and its transpose for the backward) is built once per edge_index
// (dif_csr_build; the Python side caches it) and one warp-segmented kernel handles all heads.
// Sums run in CSR slot order, which is the stable edge order => bit-reproducible.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace dif {
namespace {

__global__ void degree_kernel(const int64_t* __restrict__ ei, int64_t E, int64_t N, int32_t* __restrict__ deg_in,
                              int32_t* __restrict__ deg_src, int32_t* __restrict__ key_col, int32_t* __restrict__ key_row,
                              int32_t* __restrict__ eid, int* __restrict__ bad) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int64_t r = ei[e], c = ei[E + e];
    if (r < 0 || r >= N || c < 0 || c >= N) { atomicExch(bad, 1); key_col[e] = 0; key_row[e] = 0; eid[e] = (int32_t)e; return; }
    atomicAdd(deg_in + c, 1);     // in-degree of the target (integer => order independent)
    atomicAdd(deg_src + r, 1);    // out-degree: row pointer of the transposed CSR only
    key_col[e] = (int32_t)c;
    key_row[e] = (int32_t)r;
    eid[e] = (int32_t)e;
}

// CSR slot s holds original edge perm[s]; val uses the in-degree histogram for BOTH endpoints.
__global__ void fill_kernel(const int64_t* __restrict__ ei, const float* __restrict__ w, int64_t E, int64_t N,
                            const int32_t* __restrict__ deg_in, const int32_t* __restrict__ perm, int transpose,
                            int32_t* __restrict__ idx, float* __restrict__ val) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= E) return;
    const int32_t e = perm[s];
    const int64_t r = ei[e], c = ei[E + e];
    if (r < 0 || r >= N || c < 0 || c >= N) { idx[s] = 0; val[s] = 0.f; return; }   // invalid id: the host raises (rowptr[N] != E)
    // (1./d[col]).sqrt() and (1./d[row]).sqrt() in fp32, IEEE division / square root
    const float d_in = __fsqrt_rn(__fdiv_rn(1.f, (float)deg_in[c]));
    const float d_out = __fsqrt_rn(__fdiv_rn(1.f, (float)deg_in[r]));
    float x = (w ? w[e] : 1.f);
    x = __fmul_rn(__fmul_rn(x, d_in), d_out);
    if (!isfinite(x)) x = 0.f;     // nan_to_num(nan=0, posinf=0, neginf=0)
    idx[s] = (int32_t)(transpose ? c : r);
    val[s] = x;
}

// One (sub-)warp per target row.  LPR lanes cooperate on a row; each lane owns float4 columns
// f4 = lane, lane+LPR, ... (NV of them).  HEADMEAN folds the Hx heads before the store.
template <int LPR, int NV, bool HEADMEAN>
__global__ void __launch_bounds__(256) spmm_kernel(const float* __restrict__ x, const int32_t* __restrict__ rowptr,
                                                   const int32_t* __restrict__ idx, const float* __restrict__ val,
                                                   int64_t N, int F, int Hx, int D, float* __restrict__ out) {
    const int sub = threadIdx.x / LPR, lane = threadIdx.x % LPR;
    const int64_t row = (int64_t)blockIdx.x * (256 / LPR) + sub;
    if (row >= N) return;
    const int f4n = F >> 2;
    float4 acc[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int beg = rowptr[row], end = rowptr[row + 1];
    for (int s = beg; s < end; ++s) {
        const int32_t src = __ldg(idx + s);
        const float w = __ldg(val + s);
        const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)src * F);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int f = lane + j * LPR;
            if (f < f4n) {
                const float4 v = __ldg(xr + f);
                acc[j].x = fmaf(w, v.x, acc[j].x); acc[j].y = fmaf(w, v.y, acc[j].y);
                acc[j].z = fmaf(w, v.z, acc[j].z); acc[j].w = fmaf(w, v.w, acc[j].w);
            }
        }
    }
    if (!HEADMEAN) {
        float4* o = reinterpret_cast<float4*>(out + row * F);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int f = lane + j * LPR;
            if (f < f4n) o[f] = acc[j];
        }
    } else {
        // heads are D/4 float4 apart; with LPR == D/4 the j-th register of a lane is head j
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NV; ++j)
            if (j < Hx) { m.x += acc[j].x; m.y += acc[j].y; m.z += acc[j].z; m.w += acc[j].w; }
        const float inv = 1.f / (float)Hx;
        m.x *= inv; m.y *= inv; m.z *= inv; m.w *= inv;
        reinterpret_cast<float4*>(out + row * D)[lane] = m;
    }
}

__global__ void head_mean_kernel(const float* __restrict__ x, int64_t N, int Hx, int D, float* __restrict__ out) {
    const int d4 = D >> 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * d4) return;
    const int64_t n = i / d4;
    const int c = (int)(i - n * d4);
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int h = 0; h < Hx; ++h) {
        const float4 v = ldg4(x + (n * Hx + h) * D + 4 * c);
        m.x += v.x; m.y += v.y; m.z += v.z; m.w += v.w;
    }
    const float inv = 1.f / (float)Hx;
    *reinterpret_cast<float4*>(out + n * D + 4 * c) = make_float4(m.x * inv, m.y * inv, m.z * inv, m.w * inv);
}

struct CsrScratch {
    int32_t *deg_in, *deg_src, *key_col, *key_row, *eid, *key_out, *perm_t;
    int* bad;
    void* cub;
    size_t cub_bytes;
    int64_t total;
};

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int carve(CsrScratch& s, void* base, int64_t N, int64_t E) {
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)E);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(N + 1));
    s.cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { void* r = p ? p + off : nullptr; off += align256(bytes); return r; };
    s.deg_in = (int32_t*)take((N + 1) * 4);
    s.deg_src = (int32_t*)take((N + 1) * 4);
    s.key_col = (int32_t*)take(E * 4);
    s.key_row = (int32_t*)take(E * 4);
    s.eid = (int32_t*)take(E * 4);
    s.key_out = (int32_t*)take(E * 4);
    s.perm_t = (int32_t*)take(E * 4);
    s.bad = (int*)take(4);
    s.cub = take(s.cub_bytes);
    s.total = (int64_t)off;
    return 0;
}

int bits_for(int64_t n) {
    int b = 1;
    while (((int64_t)1 << b) < n && b < 31) ++b;
    return b;
}

}  // namespace
}  // namespace dif

using namespace dif;

extern "C" int64_t dif_csr_workspace_bytes(int64_t N, int64_t E) {
    if (N < 1 || E < 0 || N >= (1ll << 31) || E >= (1ll << 31)) return -1;
    CsrScratch s;
    carve(s, nullptr, N, E > 0 ? E : 1);
    return s.total;
}

extern "C" int dif_csr_build(const int64_t* edge_index, const float* edge_weight, int64_t N, int64_t E,
                             int32_t* rowptr, int32_t* src, float* val, int32_t* perm,
                             int32_t* rowptr_t, int32_t* dst_t, float* val_t,
                             void* workspace, int64_t workspace_bytes, void* stream) {
    DIF_REQUIRE(N >= 1 && E >= 0 && N < (1ll << 31) && E < (1ll << 31), DIF_EARG, "csr_build: N=%lld E=%lld out of int32 range", (long long)N, (long long)E);
    DIF_REQUIRE(rowptr && rowptr_t && workspace, DIF_EARG, "csr_build: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    CsrScratch s;
    carve(s, workspace, N, E > 0 ? E : 1);
    DIF_REQUIRE(workspace_bytes >= s.total, DIF_EARG, "csr_build: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)s.total);
    DIF_CUDA_OK(cudaMemsetAsync(s.deg_in, 0, (N + 1) * 4, st));
    DIF_CUDA_OK(cudaMemsetAsync(s.deg_src, 0, (N + 1) * 4, st));
    DIF_CUDA_OK(cudaMemsetAsync(s.bad, 0, 4, st));
    if (E > 0) {
        DIF_REQUIRE(edge_index && src && val && perm && dst_t && val_t, DIF_EARG, "csr_build: null pointer");
        const int blocks = (int)((E + 255) / 256);
        degree_kernel<<<blocks, 256, 0, st>>>(edge_index, E, N, s.deg_in, s.deg_src, s.key_col, s.key_row, s.eid, s.bad);
        DIF_LAUNCH_OK();
    }
    size_t cb = s.cub_bytes;
    DIF_CUDA_OK(cub::DeviceScan::ExclusiveSum(s.cub, cb, s.deg_in, rowptr, (int)(N + 1), st));
    cb = s.cub_bytes;
    DIF_CUDA_OK(cub::DeviceScan::ExclusiveSum(s.cub, cb, s.deg_src, rowptr_t, (int)(N + 1), st));
    if (E > 0) {
        const int nbits = bits_for(N);
        const int blocks = (int)((E + 255) / 256);
        cb = s.cub_bytes;   // stable LSD radix sort: slots of one target keep edge order
        DIF_CUDA_OK(cub::DeviceRadixSort::SortPairs(s.cub, cb, s.key_col, s.key_out, s.eid, perm, (int)E, 0, nbits, st));
        fill_kernel<<<blocks, 256, 0, st>>>(edge_index, edge_weight, E, N, s.deg_in, perm, 0, src, val);
        DIF_LAUNCH_OK();
        cb = s.cub_bytes;
        DIF_CUDA_OK(cub::DeviceRadixSort::SortPairs(s.cub, cb, s.key_row, s.key_out, s.eid, s.perm_t, (int)E, 0, nbits, st));
        fill_kernel<<<blocks, 256, 0, st>>>(edge_index, edge_weight, E, N, s.deg_in, s.perm_t, 1, dst_t, val_t);
        DIF_LAUNCH_OK();
    }
    // out-of-range node ids are reported through rowptr[N] != E on the host side (bad flag => -1 sentinel)
    return DIF_OK;
}

extern "C" int dif_gcn_spmm(const float* x, const int32_t* rowptr, const int32_t* idx, const float* val,
                            int64_t N, int Hx, int D, int head_mean, float* out, void* stream) {
    DIF_REQUIRE(x && rowptr && out && N >= 1 && Hx >= 1 && D >= 4 && (D % 4) == 0, DIF_EARG, "gcn_spmm: bad argument (D must be a multiple of 4)");
    const int64_t F64 = (int64_t)Hx * D;
    DIF_REQUIRE(F64 <= 4096, DIF_EUNSUPPORTED, "gcn_spmm: Hx*D=%lld > 4096", (long long)F64);
    const int F = (int)F64, f4n = F / 4;
    cudaStream_t st = (cudaStream_t)stream;
#define DIF_SPMM(LPR, NV, HM)                                                                         \
    do {                                                                                              \
        const int rows_per_block = 256 / (LPR);                                                       \
        const int blocks = (int)((N + rows_per_block - 1) / rows_per_block);                          \
        spmm_kernel<LPR, NV, HM><<<blocks, 256, 0, st>>>(x, rowptr, idx, val, N, F, Hx, D, out);      \
    } while (0)
    if (head_mean) {
        // lanes-per-row == D/4 so that a lane's j-th float4 is head j
        DIF_REQUIRE((D == 32 || D == 64 || D == 128) && Hx <= 8, DIF_EUNSUPPORTED, "gcn_spmm(head_mean): D in {32,64,128}, Hx<=8 (D=%d Hx=%d)", D, Hx);
        if (D == 32) DIF_SPMM(8, 8, true);
        else if (D == 64) DIF_SPMM(16, 8, true);
        else DIF_SPMM(32, 8, true);
    } else if (f4n <= 8) DIF_SPMM(8, 1, false);
    else if (f4n <= 16) DIF_SPMM(16, 1, false);
    else if (f4n <= 32) DIF_SPMM(32, 1, false);
    else if (f4n <= 64) DIF_SPMM(32, 2, false);
    else if (f4n <= 128) DIF_SPMM(32, 4, false);
    else if (f4n <= 256) DIF_SPMM(32, 8, false);
    else if (f4n <= 512) DIF_SPMM(32, 16, false);
    else DIF_SPMM(32, 32, false);
#undef DIF_SPMM
    DIF_LAUNCH_OK();
    return DIF_OK;
}

extern "C" int dif_head_mean(const float* x, int64_t N, int Hx, int D, float* out, void* stream) {
    DIF_REQUIRE(x && out && N >= 1 && Hx >= 1 && D >= 4 && (D % 4) == 0, DIF_EARG, "head_mean: bad argument");
    const int64_t n = N * (D / 4);
    head_mean_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, N, Hx, D, out);
    DIF_LAUNCH_OK();
    return DIF_OK;
}
