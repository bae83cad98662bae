// kernel='sigmoid' -- full_attention_conv(..., 'sigmoid'), node classification/difformer.py:45-56.
//
//   P = sigmoid(Q K^T) per head, no 1/sqrt(d) scale and no q/k normalisation (:47)
//   r = row sums of P (:50-51);  out = (P / r) V (:55-56)
//
// The reference materialises three [N,L,H] tensors.  Here the score tile lives in shared memory
// only (flash style).  P is in (0,1), so no running maximum is needed: accumulate P V and r, divide
// once.  Backward recomputes P from Q,K and uses the saved row sums:
//   D_n = g_n . out_n ; dA = g V^T ; dSc = (dA - D)/r * P (1-P) ; dQ = dSc K ; dK = dSc^T Q ; dV = (P/r)^T g
// FFMA kernels (4x4 register tiles over 64x64 score tiles); Cora-sized problems (N=2708) are
// latency-sized, a tcgen05 version is the next step for N >> 10^4.
#include "common.cuh"
#include "tile.cuh"

namespace dif {
namespace {

constexpr int kT = 64;        // query rows / key rows per tile
constexpr int kLdp = kT + 4;  // leading dimension of the score tile

__device__ __forceinline__ float sigmoidf_(float s) { return 1.f / (1.f + __expf(-s)); }

struct SigArgs {
    const float *q, *k, *v, *g, *out, *rowsum, *drow;
    int64_t N, L;
    int H, Hv, M, D;
    float *o, *rs, *dq, *dk, *dv;
    int ksplit;        // fwd: key tiles are split over gridDim.z CTAs (small-N parallelism); > 1 => partial results
    float *po, *prs;   // fwd partials: [ksplit][N,H,D] un-normalised sums and [ksplit][N,H] row sums
};

// tile loader: dst[r][0..W) = src[(row0+r), head, :] (zero beyond nrows)
__device__ __forceinline__ void load_tile(float* __restrict__ dst, int ld, const float* __restrict__ src, int64_t row0,
                                          int64_t nrows, int heads, int head, int W) {
    const int w4 = W >> 2;
    for (int idx = threadIdx.x; idx < kT * w4; idx += kThreads) {
        const int r = idx / w4, c4 = idx - r * w4;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < nrows) x = ldg4(src + ((row0 + r) * heads + head) * W + 4 * c4);
        *reinterpret_cast<float4*>(dst + r * ld + 4 * c4) = x;
    }
}

template <int TPT>
__global__ void __launch_bounds__(kThreads) sigmoid_fwd_kernel(SigArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int M = p.M, D = p.D, H = p.H;
    const int ldm = M + 4, ldd = D + 4;
    float* Qs = smem;
    float* Ks = Qs + kT * ldm;
    float* Vs = Ks + kT * ldm;
    float* Ps = Vs + kT * ldd;
    float* srs = Ps + kT * kLdp;   // [kT]
    const int tid = threadIdx.x;
    const int h = blockIdx.y, hv = (p.Hv == H) ? h : 0;
    const int64_t n0 = (int64_t)blockIdx.x * kT;
    const int sri = tid >> 4, ski = tid & 15;             // score tile owner: rows 4*sri.., keys 4*ski..
    const int tilesD = D >> 2, ntile = (kT >> 2) * tilesD;

    float acc[TPT][4][4];
#pragma unroll
    for (int t = 0; t < TPT; ++t)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[t][a][b] = 0.f;
    float rs[4] = {0.f, 0.f, 0.f, 0.f};

    load_tile(Qs, ldm, p.q, n0, p.N, H, h, M);
    // key range of this CTA (blockIdx.z of p.ksplit): whole 64-key tiles
    const int64_t ltiles = (p.L + kT - 1) / kT;
    const int64_t per = (ltiles + p.ksplit - 1) / p.ksplit;
    const int64_t l_begin = (int64_t)blockIdx.z * per * kT, l_end = min(p.L, l_begin + per * kT);
    for (int64_t l0 = l_begin; l0 < l_end; l0 += kT) {
        load_tile(Ks, ldm, p.k, l0, p.L, H, h, M);
        load_tile(Vs, ldd, p.v, l0, p.L, p.Hv, hv, D);
        __syncthreads();
        float s[4][4];
        tile_abt<16>(Qs, ldm, Ks, ldm, M, sri, ski, s);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {      // key column of (ski, b) is ski + 16 b (see tile_abt)
                const float pp = (l0 + ski + 16 * b < p.L) ? sigmoidf_(s[a][b]) : 0.f;
                Ps[(4 * sri + a) * kLdp + ski + 16 * b] = pp;
                rs[a] += pp;
            }
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < TPT; ++t) {
            const int tile = tid + t * kThreads;
            if (tile < ntile) tile_mm_acc(Ps, kLdp, Vs, ldd, kT, tile / tilesD, tile % tilesD, acc[t]);
        }
        __syncthreads();
    }
    // row sums: the 16 lanes that share sri hold partial sums
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        float v = rs[a];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (ski == 0) srs[4 * sri + a] = v;
    }
    __syncthreads();
    const bool partial = p.ksplit > 1;
    float* rs_dst = partial ? p.prs + (int64_t)blockIdx.z * p.N * H : p.rs;
    float* o_dst = partial ? p.po + (int64_t)blockIdx.z * p.N * H * D : p.o;
    if (tid < kT && n0 + tid < p.N) rs_dst[(n0 + tid) * H + h] = srs[tid];
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
        const int tile = tid + t * kThreads;
        if (tile < ntile) {
            const int ri = tile / tilesD, di = tile % tilesD;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int64_t row = n0 + 4 * ri + a;
                if (row >= p.N) continue;
                const float r = partial ? 1.f : srs[4 * ri + a];
                *reinterpret_cast<float4*>(o_dst + (row * H + h) * D + 4 * di) =
                    make_float4(acc[t][a][0] / r, acc[t][a][1] / r, acc[t][a][2] / r, acc[t][a][3] / r);
            }
        }
    }
}

// key-split combine: out = (sum_s po[s]) / (sum_s prs[s]), fixed order => deterministic
__global__ void sigmoid_combine_kernel(const float* __restrict__ po, const float* __restrict__ prs, int ksplit, int64_t rows, int D,
                                       float* __restrict__ out, float* __restrict__ rowsum) {
    const int d4 = D >> 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * d4) return;
    const int64_t row = i / d4;
    const int c = (int)(i - row * d4);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    float r = 0.f;
    for (int s = 0; s < ksplit; ++s) {
        const float4 x = ldg4(po + ((int64_t)s * rows + row) * D + 4 * c);
        a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
        r += prs[(int64_t)s * rows + row];
    }
    *reinterpret_cast<float4*>(out + row * D + 4 * c) = make_float4(a.x / r, a.y / r, a.z / r, a.w / r);
    if (c == 0) rowsum[row] = r;
}

// dst = sum_s src[s]  (fixed order => deterministic); count4 = float4 elements per partial
__global__ void sum_partials_kernel(const float* __restrict__ src, int nsplit, int64_t count4, float* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count4) return;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < nsplit; ++s) {
        const float4 x = ldg4(src + ((int64_t)s * count4 + i) * 4);
        a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
    }
    *reinterpret_cast<float4*>(dst + 4 * i) = a;
}

// D_n = g_n . out_n  (one warp per (n,h) row)
__global__ void sigmoid_drow_kernel(const float* __restrict__ g, const float* __restrict__ out, int64_t rows, int D,
                                    float* __restrict__ drow) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s = fmaf(g[row * D + i], out[row * D + i], s);
    s = warp_sum(s);
    if (lane == 0) drow[row] = s;
}

// dQ: one CTA per (query tile, head), loops over key tiles
template <int TPT>
__global__ void __launch_bounds__(kThreads) sigmoid_dq_kernel(SigArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int M = p.M, D = p.D, H = p.H;
    const int ldm = M + 4, ldd = D + 4;
    float* Qs = smem;
    float* Gs = Qs + kT * ldm;
    float* Ks = Gs + kT * ldd;
    float* Vs = Ks + kT * ldm;
    float* Ps = Vs + kT * ldd;
    float* srs = Ps + kT * kLdp;    // 1/r
    float* sdr = srs + kT;          // D_n
    const int tid = threadIdx.x;
    const int h = blockIdx.y, hv = (p.Hv == H) ? h : 0;
    const int64_t n0 = (int64_t)blockIdx.x * kT;
    const int sri = tid >> 4, ski = tid & 15;
    const int tilesM = M >> 2, ntile = (kT >> 2) * tilesM;

    float acc[TPT][4][4];
#pragma unroll
    for (int t = 0; t < TPT; ++t)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[t][a][b] = 0.f;

    load_tile(Qs, ldm, p.q, n0, p.N, H, h, M);
    load_tile(Gs, ldd, p.g, n0, p.N, H, h, D);
    if (tid < kT) {
        const bool ok = n0 + tid < p.N;
        srs[tid] = ok ? 1.f / p.rowsum[(n0 + tid) * H + h] : 0.f;
        sdr[tid] = ok ? p.drow[(n0 + tid) * H + h] : 0.f;
    }
    const int64_t ltiles = (p.L + kT - 1) / kT, lper = (ltiles + p.ksplit - 1) / p.ksplit;
    const int64_t l_begin = (int64_t)blockIdx.z * lper * kT, l_end = min(p.L, l_begin + lper * kT);
    for (int64_t l0 = l_begin; l0 < l_end; l0 += kT) {
        load_tile(Ks, ldm, p.k, l0, p.L, H, h, M);
        load_tile(Vs, ldd, p.v, l0, p.L, p.Hv, hv, D);
        __syncthreads();
        float s[4][4], da[4][4];
        tile_abt<16>(Qs, ldm, Ks, ldm, M, sri, ski, s);
        tile_abt<16>(Gs, ldd, Vs, ldd, D, sri, ski, da);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float ir = srs[4 * sri + a], dr = sdr[4 * sri + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float pp = sigmoidf_(s[a][b]);
                Ps[(4 * sri + a) * kLdp + ski + 16 * b] =
                    (l0 + ski + 16 * b < p.L) ? (da[a][b] - dr) * ir * pp * (1.f - pp) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < TPT; ++t) {
            const int tile = tid + t * kThreads;
            if (tile < ntile) tile_mm_acc(Ps, kLdp, Ks, ldm, kT, tile / tilesM, tile % tilesM, acc[t]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
        const int tile = tid + t * kThreads;
        if (tile < ntile) {
            const int ri = tile / tilesM, mi = tile % tilesM;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int64_t row = n0 + 4 * ri + a;
                float* dq_dst = p.ksplit > 1 ? p.po + (int64_t)blockIdx.z * p.N * H * M : p.dq;
                if (row < p.N)
                    *reinterpret_cast<float4*>(dq_dst + (row * H + h) * M + 4 * mi) =
                        make_float4(acc[t][a][0], acc[t][a][1], acc[t][a][2], acc[t][a][3]);
            }
        }
    }
}

// dK, dV: one CTA per key tile (and per head unless V is broadcast), loops over query tiles
template <int TPT_K, int TPT_V>
__global__ void __launch_bounds__(kThreads) sigmoid_dkv_kernel(SigArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int M = p.M, D = p.D, H = p.H;
    const int ldm = M + 4, ldd = D + 4;
    float* Qs = smem;
    float* Gs = Qs + kT * ldm;
    float* Ks = Gs + kT * ldd;
    float* Vs = Ks + kT * ldm;
    float* Ps = Vs + kT * ldd;      // dSc[n][l]
    float* As = Ps + kT * kLdp;     // P/r [n][l]
    float* srs = As + kT * kLdp;
    float* sdr = srs + kT;
    const int tid = threadIdx.x;
    const bool bcast = (p.Hv != H);
    const int64_t l0 = (int64_t)blockIdx.x * kT;
    const int sri = tid >> 4, ski = tid & 15;
    const int tilesM = M >> 2, ntileK = (kT >> 2) * tilesM;
    const int tilesD = D >> 2, ntileV = (kT >> 2) * tilesD;

    float accv[TPT_V][4][4];
#pragma unroll
    for (int t = 0; t < TPT_V; ++t)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) accv[t][a][b] = 0.f;

    const int h_begin = bcast ? 0 : blockIdx.y, h_end = bcast ? H : blockIdx.y + 1;
    for (int h = h_begin; h < h_end; ++h) {
        const int hv = bcast ? 0 : h;
        float acck[TPT_K][4][4];
#pragma unroll
        for (int t = 0; t < TPT_K; ++t)
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acck[t][a][b] = 0.f;
        __syncthreads();
        load_tile(Ks, ldm, p.k, l0, p.L, H, h, M);
        if (h == h_begin) load_tile(Vs, ldd, p.v, l0, p.L, p.Hv, hv, D);
        const int64_t ntiles_q = (p.N + kT - 1) / kT, nper = (ntiles_q + p.ksplit - 1) / p.ksplit;
        const int64_t n_begin = (int64_t)blockIdx.z * nper * kT, n_end = min(p.N, n_begin + nper * kT);
        for (int64_t n0 = n_begin; n0 < n_end; n0 += kT) {
            load_tile(Qs, ldm, p.q, n0, p.N, H, h, M);
            load_tile(Gs, ldd, p.g, n0, p.N, H, h, D);
            if (tid < kT) {
                const bool ok = n0 + tid < p.N;
                srs[tid] = ok ? 1.f / p.rowsum[(n0 + tid) * H + h] : 0.f;
                sdr[tid] = ok ? p.drow[(n0 + tid) * H + h] : 0.f;
            }
            __syncthreads();
            float s[4][4], da[4][4];
            tile_abt<16>(Qs, ldm, Ks, ldm, M, sri, ski, s);
            tile_abt<16>(Gs, ldd, Vs, ldd, D, sri, ski, da);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float ir = srs[4 * sri + a], dr = sdr[4 * sri + a];   // ir == 0 for padded query rows
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const float pp = sigmoidf_(s[a][b]);
                    As[(4 * sri + a) * kLdp + ski + 16 * b] = pp * ir;
                    Ps[(4 * sri + a) * kLdp + ski + 16 * b] = (da[a][b] - dr) * ir * pp * (1.f - pp);
                }
            }
            __syncthreads();
#pragma unroll
            for (int t = 0; t < TPT_K; ++t) {
                const int tile = tid + t * kThreads;
                if (tile < ntileK) tile_atb_acc(Ps, kLdp, Qs, ldm, kT, tile / tilesM, tile % tilesM, acck[t]);
            }
#pragma unroll
            for (int t = 0; t < TPT_V; ++t) {
                const int tile = tid + t * kThreads;
                if (tile < ntileV) tile_atb_acc(As, kLdp, Gs, ldd, kT, tile / tilesD, tile % tilesD, accv[t]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int t = 0; t < TPT_K; ++t) {
            const int tile = tid + t * kThreads;
            if (tile < ntileK) {
                const int li = tile / tilesM, mi = tile % tilesM;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int64_t row = l0 + 4 * li + a;
                    float* dk_dst = p.ksplit > 1 ? p.po + (int64_t)blockIdx.z * p.L * H * M : p.dk;
                    if (row < p.L)
                        *reinterpret_cast<float4*>(dk_dst + (row * H + h) * M + 4 * mi) =
                            make_float4(acck[t][a][0], acck[t][a][1], acck[t][a][2], acck[t][a][3]);
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < TPT_V; ++t) {
        const int tile = tid + t * kThreads;
        if (tile < ntileV) {
            const int li = tile / tilesD, di = tile % tilesD;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int64_t row = l0 + 4 * li + a;
                float* dv_dst = p.ksplit > 1 ? p.prs + (int64_t)blockIdx.z * p.L * p.Hv * D : p.dv;
                if (row < p.L)
                    *reinterpret_cast<float4*>(dv_dst + (row * p.Hv + (bcast ? 0 : blockIdx.y)) * D + 4 * di) =
                        make_float4(accv[t][a][0], accv[t][a][1], accv[t][a][2], accv[t][a][3]);
            }
        }
    }
}

int sig_check(int64_t N, int64_t L, int H, int Hv, int M, int D) {
    DIF_REQUIRE(N >= 1 && L >= 1 && H >= 1 && H <= 65535, DIF_EARG, "sigmoid: bad N/L/H");
    DIF_REQUIRE(Hv == H || Hv == 1, DIF_EARG, "sigmoid: Hv=%d must equal H=%d or 1", Hv, H);
    DIF_REQUIRE(M >= 4 && D >= 4 && (M % 4) == 0 && (D % 4) == 0 && M <= 128 && D <= 128, DIF_EUNSUPPORTED,
                "sigmoid: need M,D multiples of 4 in [4,128], got M=%d D=%d", M, D);
    return DIF_OK;
}

template <typename K>
int set_smem_(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) DIF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return DIF_OK;
}

}  // namespace
}  // namespace dif

using namespace dif;

static int sigmoid_ksplit(int64_t N, int64_t L, int H) {
    const int64_t ctas = ((N + kT - 1) / kT) * H, ltiles = (L + kT - 1) / kT;
    int64_t s = (2 * (int64_t)sm_count() + ctas - 1) / ctas;          // aim at ~2 CTAs per SM
    if (s > ltiles) s = ltiles;
    if (s > 32) s = 32;
    return s < 1 ? 1 : (int)s;
}

static int g_sigmoid_impl = DIF_IMPL_AUTO;

extern "C" int dif_sigmoid_set_impl(int impl) {
    DIF_REQUIRE(impl == DIF_IMPL_AUTO || impl == DIF_IMPL_GENERIC || impl == DIF_IMPL_TCGEN05, DIF_EARG, "sigmoid: unknown impl %d", impl);
    g_sigmoid_impl = impl;
    return DIF_OK;
}

extern "C" int64_t dif_sigmoid_fwd_workspace_bytes(int64_t N, int64_t L, int H, int Hv, int M, int D) {
    const int s = sigmoid_ksplit(N, L, H);
    const bool tc = sigmoid_tc_supported(N, L, H, Hv, M, D);
    const int t = tc ? sigmoid_tc_ksplit(N, L, H) : 1;
    const int m = s > t ? s : t;                                      // either implementation fits
    return (m > 1 ? (int64_t)m * N * H * (D + 1) * (int64_t)sizeof(float) : 0) + (tc ? sigmoid_tc_image_bytes(L, H, Hv) : 0);
}

extern "C" int dif_sigmoid_fwd(const float* q, const float* k, const float* v, int64_t N, int64_t L, int H, int Hv, int M, int D,
                               float* out, float* rowsum, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = sig_check(N, L, H, Hv, M, D);
    if (rc) return rc;
    DIF_REQUIRE(q && k && v && out && rowsum, DIF_EARG, "sigmoid_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const bool tc_ok = sigmoid_tc_supported(N, L, H, Hv, M, D);
    DIF_REQUIRE(g_sigmoid_impl != DIF_IMPL_TCGEN05 || tc_ok, DIF_EUNSUPPORTED, "sigmoid: tcgen05 path needs M == D == 64 (got M=%d D=%d)", M, D);
    if (tc_ok && g_sigmoid_impl != DIF_IMPL_GENERIC) {
        const int ks = sigmoid_tc_ksplit(N, L, H);
        float *po = nullptr, *prs = nullptr;
        const int64_t pbytes = ks > 1 ? (int64_t)ks * N * H * (D + 1) * 4 : 0;      // [partials | K, V operand images]
        DIF_REQUIRE(workspace && workspace_bytes >= pbytes + sigmoid_tc_image_bytes(L, H, Hv), DIF_EARG, "sigmoid_fwd: workspace too small");
        if (ks > 1) {
            po = (float*)workspace;
            prs = po + (int64_t)ks * N * H * D;
        }
        if ((rc = sigmoid_fwd_tc(q, k, v, N, L, H, Hv, out, rowsum, po, prs, ks, (char*)workspace + pbytes, st))) return rc;
        if (ks > 1) {
            const int64_t n = N * H * (D / 4);
            sigmoid_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(po, prs, ks, N * H, D, out, rowsum);
            DIF_LAUNCH_OK();
        }
        return DIF_OK;
    }
    SigArgs a{};
    a.q = q; a.k = k; a.v = v; a.N = N; a.L = L; a.H = H; a.Hv = Hv; a.M = M; a.D = D; a.o = out; a.rs = rowsum;
    a.ksplit = sigmoid_ksplit(N, L, H);
    if (a.ksplit > 1) {
        DIF_REQUIRE(workspace && workspace_bytes >= (int64_t)a.ksplit * N * H * (D + 1) * 4, DIF_EARG, "sigmoid_fwd: workspace too small");
        a.po = (float*)workspace;
        a.prs = a.po + (int64_t)a.ksplit * N * H * D;
    }
    const size_t smem = ((size_t)2 * kT * (M + 4) + (size_t)kT * (D + 4) + (size_t)kT * kLdp + kT) * sizeof(float);
    dim3 grid((unsigned)((N + kT - 1) / kT), H, a.ksplit);
    if ((kT / 4) * (D / 4) <= kThreads) {
        if ((rc = set_smem_(sigmoid_fwd_kernel<1>, smem))) return rc;
        sigmoid_fwd_kernel<1><<<grid, kThreads, smem, st>>>(a);
    } else {
        if ((rc = set_smem_(sigmoid_fwd_kernel<2>, smem))) return rc;
        sigmoid_fwd_kernel<2><<<grid, kThreads, smem, st>>>(a);
    }
    DIF_LAUNCH_OK();
    if (a.ksplit > 1) {
        const int64_t n = N * H * (D / 4);
        sigmoid_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a.po, a.prs, a.ksplit, N * H, D, out, rowsum);
        DIF_LAUNCH_OK();
    }
    return DIF_OK;
}

static int sigmoid_bwd_split(int64_t rows, int64_t other, int heads) {
    const int64_t ctas = ((rows + kT - 1) / kT) * heads, otiles = (other + kT - 1) / kT;
    int64_t s = (2 * (int64_t)sm_count() + ctas - 1) / ctas;
    if (s > otiles) s = otiles;
    if (s > 32) s = 32;
    return s < 1 ? 1 : (int)s;
}

static int64_t drow_floats(int64_t N, int H) { return (N * H + 15) & ~(int64_t)15; }    // 64-byte aligned

extern "C" int64_t dif_sigmoid_bwd_workspace_bytes(int64_t N, int64_t L, int H, int Hv, int M, int D) {
    const int sq = sigmoid_bwd_split(N, L, H), skv = sigmoid_bwd_split(L, N, Hv == H ? H : 1);
    int64_t fl = N * H;                                                   // D_n = g.out
    int64_t a = sq > 1 ? (int64_t)sq * N * H * M : 0;                      // dq partials
    int64_t b = skv > 1 ? (int64_t)skv * (L * H * M + L * Hv * D) : 0;     // dk, dv partials
    int64_t generic = (fl + (a > b ? a : b) + 64) * (int64_t)sizeof(float);
    if (!sigmoid_bwd_tc_supported(N, L, H, Hv, M, D)) return generic;
    // tcgen05 backward: D_n | bf16 hi/lo operand images of Qs, G, K, V + (D, 1/r) scalars | split partials
    const int tq = sigmoid_bwd_tc_split(N, L, H), tkv = sigmoid_bwd_tc_split(L, N, H);
    const int64_t pa = tq > 1 ? (int64_t)tq * N * H * M : 0, pb = tkv > 1 ? (int64_t)tkv * 2 * L * H * M : 0;
    const int64_t tc = drow_floats(N, H) * 4 + sigmoid_bwd_tc_image_bytes(N, L, H) + (pa + pb) * 4 + 256;
    return tc > generic ? tc : generic;
}

extern "C" int dif_sigmoid_bwd(const float* q, const float* k, const float* v, const float* g, const float* out,
                               const float* rowsum, int64_t N, int64_t L, int H, int Hv, int M, int D,
                               float* dq, float* dk, float* dv, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = sig_check(N, L, H, Hv, M, D);
    if (rc) return rc;
    DIF_REQUIRE(q && k && v && g && out && rowsum && dq && dk && dv && workspace, DIF_EARG, "sigmoid_bwd: null pointer");
    DIF_REQUIRE(workspace_bytes >= dif_sigmoid_bwd_workspace_bytes(N, L, H, Hv, M, D), DIF_EARG, "sigmoid_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* drow = (float*)workspace;
    float* pbuf = drow + drow_floats(N, H);                               // 64-byte aligned partial buffers
    {
        const int64_t rows = N * H;
        const int64_t threads = rows * 32;
        sigmoid_drow_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(g, out, rows, D, drow);
        DIF_LAUNCH_OK();
    }
    if (g_sigmoid_impl != DIF_IMPL_GENERIC && sigmoid_bwd_tc_supported(N, L, H, Hv, M, D)) {
        // ---- tensor-core backward (sigmoid_bwd_sm100.cu): operand images, then the dq kernel and the dk/dv kernel
        uint8_t* images = reinterpret_cast<uint8_t*>(pbuf);
        float* part_dq = reinterpret_cast<float*>(images + sigmoid_bwd_tc_image_bytes(N, L, H));
        const int tq = sigmoid_bwd_tc_split(N, L, H), tkv = sigmoid_bwd_tc_split(L, N, H);
        float* part_dk = part_dq + (tq > 1 ? (int64_t)tq * N * H * M : 0);
        float* part_dv = part_dk + (int64_t)tkv * L * H * M;
        if ((rc = sigmoid_bwd_tc(q, k, v, g, drow, rowsum, N, L, H, dq, dk, dv, images, tq, part_dq, tkv, part_dk, part_dv, st))) return rc;
        if (tq > 1) {
            const int64_t c4 = N * H * M / 4;
            sum_partials_kernel<<<(unsigned)((c4 + 255) / 256), 256, 0, st>>>(part_dq, tq, c4, dq);
            DIF_LAUNCH_OK();
        }
        if (tkv > 1) {
            const int64_t c4 = L * H * M / 4;
            sum_partials_kernel<<<(unsigned)((c4 + 255) / 256), 256, 0, st>>>(part_dk, tkv, c4, dk);
            sum_partials_kernel<<<(unsigned)((c4 + 255) / 256), 256, 0, st>>>(part_dv, tkv, c4, dv);
            DIF_LAUNCH_OK();
        }
        return DIF_OK;
    }
    SigArgs a{};
    a.q = q; a.k = k; a.v = v; a.g = g; a.out = out; a.rowsum = rowsum; a.drow = drow;
    a.N = N; a.L = L; a.H = H; a.Hv = Hv; a.M = M; a.D = D; a.dq = dq; a.dk = dk; a.dv = dv;
    {
        const size_t smem = ((size_t)2 * kT * (M + 4) + (size_t)2 * kT * (D + 4) + (size_t)kT * kLdp + 2 * kT) * sizeof(float);
        a.ksplit = sigmoid_bwd_split(N, L, H);        // small N: split the key loop over gridDim.z, sum the partials
        a.po = pbuf;
        dim3 grid((unsigned)((N + kT - 1) / kT), H, a.ksplit);
        if ((kT / 4) * (M / 4) <= kThreads) {
            if ((rc = set_smem_(sigmoid_dq_kernel<1>, smem))) return rc;
            sigmoid_dq_kernel<1><<<grid, kThreads, smem, st>>>(a);
        } else {
            if ((rc = set_smem_(sigmoid_dq_kernel<2>, smem))) return rc;
            sigmoid_dq_kernel<2><<<grid, kThreads, smem, st>>>(a);
        }
        DIF_LAUNCH_OK();
        if (a.ksplit > 1) {
            const int64_t c4 = N * H * M / 4;
            sum_partials_kernel<<<(unsigned)((c4 + 255) / 256), 256, 0, st>>>(pbuf, a.ksplit, c4, dq);
            DIF_LAUNCH_OK();
        }
    }
    {
        const size_t smem = ((size_t)2 * kT * (M + 4) + (size_t)2 * kT * (D + 4) + (size_t)2 * kT * kLdp + 2 * kT) * sizeof(float);
        a.ksplit = sigmoid_bwd_split(L, N, Hv == H ? H : 1);
        a.po = pbuf;                                          // dk partials
        a.prs = pbuf + (int64_t)a.ksplit * L * H * M;         // dv partials
        dim3 grid((unsigned)((L + kT - 1) / kT), Hv == H ? H : 1, a.ksplit);
        const int tk = (kT / 4) * (M / 4) <= kThreads ? 1 : 2;
        const int tv = (kT / 4) * (D / 4) <= kThreads ? 1 : 2;
#define DIF_SKV(A, B_)                                                            \
    do {                                                                          \
        if ((rc = set_smem_(sigmoid_dkv_kernel<A, B_>, smem))) return rc;         \
        sigmoid_dkv_kernel<A, B_><<<grid, kThreads, smem, st>>>(a);               \
    } while (0)
        if (tk == 1 && tv == 1) DIF_SKV(1, 1);
        else if (tk == 1) DIF_SKV(1, 2);
        else if (tv == 1) DIF_SKV(2, 1);
        else DIF_SKV(2, 2);
#undef DIF_SKV
        DIF_LAUNCH_OK();
        if (a.ksplit > 1) {
            const int64_t ck = L * H * M / 4, cv = L * Hv * D / 4;
            sum_partials_kernel<<<(unsigned)((ck + 255) / 256), 256, 0, st>>>(a.po, a.ksplit, ck, dk);
            sum_partials_kernel<<<(unsigned)((cv + 255) / 256), 256, 0, st>>>(a.prs, a.ksplit, cv, dv);
            DIF_LAUNCH_OK();
        }
    }
    return DIF_OK;
}
