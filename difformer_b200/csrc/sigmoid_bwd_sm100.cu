// Backward of kernel='sigmoid' (derived from node classification/difformer.py:45-56; the reference uses autograd) on
// tcgen05, in the structure of the forward (sigmoid_sm100.cu).  Dispatched by dif_sigmoid_bwd (sigmoid.cu) for
// M == D == 64, Hv == H unless the implementation is pinned to DIF_IMPL_GENERIC.
//
// Math, with P = sigmoid(S), S = Q K^T, r_n = sum_l P_nl, out = (P V) / r, D_n = g_n . out_n:
//     dV_l = sum_n (P_nl / r_n) g_n                    dP_nl = (g_n . v_l - D_n) / r_n
//     dS_nl = dP_nl P_nl (1 - P_nl)                    dQ_n = sum_l dS_nl k_l          dK_l = sum_n dS_nl q_n
//
// One kernel template, two instantiations.  "own" rows sit on the 128 TMEM lanes (A operands, shared memory, loaded
// once), "streamed" 128-row tiles are the B operands (double-buffered TMA bulk copies of pre-converted bf16 hi|lo
// images, as in the forward):
//     KV = false (dq):     own = (Qs, G) rows of a query tile, streamed = (K, V) key tiles
//                          S = Qs K^T, T = G V^T  ->  dS  ->  dQ += dS K
//     KV = true  (dk, dv): own = (K, V) rows of a key tile, streamed = (Qs, G) query tiles + their (D_n, 1/r_n)
//                          S' = K Qs^T, T' = V G^T  ->  W = P / r, dS  ->  dV += W G,  dK += dS Qs
// Qs = -log2(e) Q (so that P = 1 / (1 + 2^S) needs no multiply); dK is therefore accumulated against Qs and scaled
// by -ln 2 at the end.  P (1 - P) is evaluated as e P^2 with e = 2^x, x clamped at 62: accurate in both tails, no
// cancellation, no inf * 0.
// Rows beyond N / L need no masks: their image rows are zero (and their (D, 1/r) are zero), so every product they
// enter is multiplied by a zero operand row or never stored.
//
// TMEM (512 columns):  S | T | acc0 | acc1 (128 each).  dS (and W) are written as bf16 hi|lo over the score columns
// they came from ([16 words hi | 16 words lo] per 32-column chunk) and read by the MMA as A operands from tensor
// memory; the streamed tile is the MN-major B operand, its hi and lo blocks (16 KB apart) forming one N = 128 operand,
// so acc = [ . x hi | . x lo ] and the epilogue adds the two 64-column halves.
// S/T are single-buffered (no room for more): per tile  [S,T MMAs] -> [elementwise] -> [accumulate MMAs].
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dif {
namespace {

constexpr int kT = 128;
constexpr int kOpT = kT * 128;           // one bf16 [128 rows][64] operand tile: 16 KB
constexpr int kImg = 2 * kOpT;           // hi | lo
constexpr int kScal = 2 * kT * 4;        // (D_n, 1/r_n) of one query tile: 1 KB
constexpr int kBwdWarps = 11, kBwdThreads = kBwdWarps * 32;
constexpr int kSmemBwd = 2 * kImg + 2 * (2 * kImg + kScal) + 1024;      // own (2 images) + 2 stages x (2 images + scalars)

struct SigBwdArgs {
    const uint8_t *own_a, *own_b;        // images of the own rows: [heads][tiles] x 32 KB  (dq: Qs, G; dkv: K, V)
    const uint8_t *str_a, *str_b;        // images of the streamed rows                      (dq: K, V; dkv: Qs, G)
    const float* str_scal;               // dkv: [H][qtiles][2][128] (D_n, 1/r_n), zero-padded; dq: unused
    const float *drow, *rowsum;          // dq: D_n and r_n of the own rows, [N, H]
    int64_t n_own, n_str;                // rows on the own / streamed side
    int H, split;
    float *out0, *out1;                  // dq: out0 = dq; dkv: out0 = dv, out1 = dk   (split == 1: final [rows,H,64])
    float *part0, *part1;                // split > 1: [split][rows,H,64] partial sums
};

// e = 2^min(x, 62), P = 1 / (1 + e), pp = e P^2 = P (1 - P)
__device__ __forceinline__ void sig_pair(float x, float& P, float& pp) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(x, 62.f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(P) : "f"(1.f + e));
    pp = e * P * P;
}
// 8 values -> 4 words bf16 hi, 4 words bf16 lo
__device__ __forceinline__ void split_words(const float (&v)[8], uint32_t* h, uint32_t* l) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        h[k] = bf2_bits(v[2 * k], v[2 * k + 1]);
        l[k] = bf2_bits(v[2 * k] - __uint_as_float(h[k] << 16), v[2 * k + 1] - __uint_as_float(h[k] & 0xffff0000u));
    }
}

template <bool KV>
__global__ void __launch_bounds__(kBwdThreads, 1) sigmoid_bwd_tc_kernel(const __grid_constant__ SigBwdArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* OwnA = base;                               // hi | lo
    uint8_t* OwnB = OwnA + kImg;
    uint8_t* Stg = OwnB + kImg;                         // 2 x (A image | B image | scalars)
    constexpr int kStage = 2 * kImg + kScal;
    __shared__ uint64_t ownfull, sfull[2], sempty[2], stready, dsready, done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.H, h = blockIdx.y;
    const int64_t own_tiles = (p.n_own + kT - 1) / kT, str_tiles = (p.n_str + kT - 1) / kT;
    const int64_t per = (str_tiles + p.split - 1) / p.split;
    const int64_t t0 = (int64_t)blockIdx.z * per, t1 = min(str_tiles, t0 + per);
    const int T = (int)max((int64_t)0, t1 - t0);
    const int64_t row0 = (int64_t)blockIdx.x * kT;

    if (tid == 0) {
        mbar_init(&ownfull, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&sfull[s], 1); mbar_init(&sempty[s], 1); }
        mbar_init(&stready, 1); mbar_init(&dsready, 8); mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 10) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t kColS = 0, kColT = kT, kAcc0 = 2 * kT, kAcc1 = 3 * kT;

    if (warp < 8) {
        // ===== elementwise: thread = (own row r, half of the streamed rows)
        const int quad = warp & 3, half = warp >> 2, r = quad * 32 + lane;
        const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
        const int64_t row = row0 + r;
        float Dn = 0.f, In = 0.f;                        // dq: scalars of this thread's own (query) row
        if (!KV && row < p.n_own) { Dn = p.drow[row * H + h]; In = 1.f / p.rowsum[row * H + h]; }
        for (int i = 0; i < T; ++i) {
            if (KV) mbar_wait(&sfull[i & 1], (i >> 1) & 1);      // the (D, 1/r) scalars of this tile came in with the stage
            mbar_wait(&stready, i & 1);
            tc_fence_after();
            const float* sc = reinterpret_cast<const float*>(Stg + (i & 1) * kStage + 2 * kImg);      // [D: 128 | 1/r: 128]
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int col = half * 64 + 32 * c;
                uint32_t xs[32], xt[32];
                tmem_ld32(tlane + kColS + col, xs);
                tmem_ld32(tlane + kColT + col, xt);
                tmem_ld_wait32(xs);
                tmem_ld_wait32(xt);
                uint32_t ws[32], wt[32];                 // over S: dS (dq) / W (dkv); over T: dS (dkv)
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    float a[8], b[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float P, pp;
                        sig_pair(__uint_as_float(xs[j + e]), P, pp);
                        const float t = __uint_as_float(xt[j + e]);
                        if (KV) {
                            const float d = sc[col + j + e], inv = sc[kT + col + j + e];      // warp-uniform addresses: broadcast
                            a[e] = P * inv;
                            b[e] = (t - d) * inv * pp;
                        } else {
                            a[e] = (t - Dn) * In * pp;
                        }
                    }
                    split_words(a, &ws[j >> 1], &ws[16 + (j >> 1)]);
                    if (KV) split_words(b, &wt[j >> 1], &wt[16 + (j >> 1)]);
                }
                tmem_st32(tlane + kColS + col, ws);
                if (KV) tmem_st32(tlane + kColT + col, wt);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dsready);
        }
        // ---- epilogue: acc = [ . x hi | . x lo ]; this thread owns output columns [half * 32, +32) of its row
        mbar_wait(&done, 0);
        tc_fence_after();
        auto store_acc = [&](uint32_t colbase, float* dst, float* part, float scale) {
            uint32_t o[32], o2[32];
            if (T > 0) {
                tmem_ld32(tlane + colbase + half * 32, o);
                tmem_ld32(tlane + colbase + kDim + half * 32, o2);
                tmem_ld_wait32(o);
                tmem_ld_wait32(o2);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) { o[j] = 0u; o2[j] = 0u; }
            }
            float* d = p.split > 1 ? part + (int64_t)blockIdx.z * p.n_own * H * kDim : dst;
            if (row < p.n_own) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(d + (row * H + h) * kDim + half * 32 + j) =
                        make_float4((__uint_as_float(o[j]) + __uint_as_float(o2[j])) * scale, (__uint_as_float(o[j + 1]) + __uint_as_float(o2[j + 1])) * scale,
                                    (__uint_as_float(o[j + 2]) + __uint_as_float(o2[j + 2])) * scale, (__uint_as_float(o[j + 3]) + __uint_as_float(o2[j + 3])) * scale);
            }
        };
        store_acc(kAcc0, p.out0, p.part0, 1.f);                                  // dq  /  dv
        if (KV) store_acc(kAcc1, p.out1, p.part1, -0.6931471805599453f);         // dk = (dS Qs) / (-log2 e)
    } else if (warp == 8 || warp == 9) {
        // ===== loader (warp 8, lane 0): the own images once, then one stage (two images + scalars) per streamed tile; warp 9 idles
        if (lane == 0) {
            if (warp == 8) {
                mbar_expect_tx(&ownfull, 2 * kImg);
                tma_load_1d(smem_u32(OwnA), p.own_a + ((int64_t)h * own_tiles + blockIdx.x) * kImg, kImg, &ownfull);
                tma_load_1d(smem_u32(OwnB), p.own_b + ((int64_t)h * own_tiles + blockIdx.x) * kImg, kImg, &ownfull);
            }
            for (int i = 0; i < T && warp == 8; ++i) {
                const int s = i & 1;
                if (i >= 2) mbar_wait(&sempty[s], ((i >> 1) - 1) & 1);
                mbar_expect_tx(&sfull[s], 2 * kImg + (KV ? kScal : 0));
                tma_load_1d(smem_u32(Stg) + s * kStage, p.str_a + ((int64_t)h * str_tiles + t0 + i) * kImg, kImg, &sfull[s]);
                tma_load_1d(smem_u32(Stg) + s * kStage + kImg, p.str_b + ((int64_t)h * str_tiles + t0 + i) * kImg, kImg, &sfull[s]);
                if (KV) tma_load_1d(smem_u32(Stg) + s * kStage + 2 * kImg, p.str_scal + ((int64_t)h * str_tiles + t0 + i) * (2 * kT), kScal, &sfull[s]);
            }
        }
    } else {
        // ===== MMA issuer: whole warp converged, one elected lane issues
        const bool leader = elect_one();
        const uint32_t idSS = make_idesc(kT, kT, 0, 0);          // S, T: both operands K-major (shared memory)
        const uint32_t idTS = make_idesc(kT, 2 * kDim, 0, 1);    // acc += dS [hi | lo]: A in TMEM, B MN-major, N = 128
        mbar_wait(&ownfull, 0);
        for (int i = 0; i < T; ++i) {
            const int s = i & 1;
            mbar_wait(&sfull[s], (i >> 1) & 1);
            tc_fence_after();
            const uint32_t sa = smem_u32(Stg) + s * kStage, sb = sa + kImg, oa = smem_u32(OwnA), ob = smem_u32(OwnB);
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {                 // S = ownA strA^T : hi hi + lo hi + hi lo
                    const uint64_t ah = make_desc(oa + ks * 32, kKmajLBO, kKmajSBO), al = make_desc(oa + kOpT + ks * 32, kKmajLBO, kKmajSBO);
                    const uint64_t bh = make_desc(sa + ks * 32, kKmajLBO, kKmajSBO), bl = make_desc(sa + kOpT + ks * 32, kKmajLBO, kKmajSBO);
                    umma(tmem + kColS, ah, bh, idSS, ks > 0 ? 1u : 0u);
                    umma(tmem + kColS, al, bh, idSS, 1u);
                    umma(tmem + kColS, ah, bl, idSS, 1u);
                }
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {                 // T = ownB strB^T
                    const uint64_t ah = make_desc(ob + ks * 32, kKmajLBO, kKmajSBO), al = make_desc(ob + kOpT + ks * 32, kKmajLBO, kKmajSBO);
                    const uint64_t bh = make_desc(sb + ks * 32, kKmajLBO, kKmajSBO), bl = make_desc(sb + kOpT + ks * 32, kKmajLBO, kKmajSBO);
                    umma(tmem + kColT, ah, bh, idSS, ks > 0 ? 1u : 0u);
                    umma(tmem + kColT, al, bh, idSS, 1u);
                    umma(tmem + kColT, ah, bl, idSS, 1u);
                }
                umma_commit(&stready);
            }
            mbar_wait(&dsready, i & 1);
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {                 // 16 streamed rows per step = 8 packed TMEM columns
                    const uint32_t ch = (ks >> 1) * 32 + (ks & 1) * 8;       // 32-row chunk: [16 words hi | 16 words lo]
                    if (!KV) {                                   // dQ += dS K      (K = streamed A image, MN-major [hi | lo])
                        const uint64_t kb = make_desc(sa + ks * 2048, kOpT, 1024);
                        umma_ts(tmem + kAcc0, tmem + kColS + ch, kb, idTS, (i > 0 || ks > 0) ? 1u : 0u);
                        umma_ts(tmem + kAcc0, tmem + kColS + ch + 16, kb, idTS, 1u);
                    } else {                                     // dV += W G (streamed B image), dK += dS Qs (streamed A image)
                        const uint64_t gb = make_desc(sb + ks * 2048, kOpT, 1024), qb = make_desc(sa + ks * 2048, kOpT, 1024);
                        umma_ts(tmem + kAcc0, tmem + kColS + ch, gb, idTS, (i > 0 || ks > 0) ? 1u : 0u);
                        umma_ts(tmem + kAcc0, tmem + kColS + ch + 16, gb, idTS, 1u);
                        umma_ts(tmem + kAcc1, tmem + kColT + ch, qb, idTS, (i > 0 || ks > 0) ? 1u : 0u);
                        umma_ts(tmem + kAcc1, tmem + kColT + ch + 16, qb, idTS, 1u);
                    }
                }
                umma_commit(&sempty[s]);
            }
            __syncwarp();
        }
        if (leader) { if (T > 0) umma_commit(&done); else mbar_arrive(&done); }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 10) tmem_dealloc(tmem, 512);
}

// [128 rows][64 floats] of one head -> bf16 hi/lo image (rows of 128 B, 8-row swizzle atoms): serves as K-major
// operand (rows = M or N index) and as MN-major B operand (rows = K index).  256 threads.
__device__ __forceinline__ void convert_tile_g(const float* src, int heads, int head, int64_t row0, int64_t nrows, int tid, uint8_t* g_hi,
                                               float scale) {
    float x[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int t = tid + 256 * j;
        const int64_t row = row0 + (t >> 3);
        if (row < nrows) ldg256_keep(src + (row * heads + head) * kDim + (t & 7) * 8, x[j]);
        else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[j][i] = 0.f;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int t = tid + 256 * j;
#pragma unroll
        for (int i = 0; i < 8; ++i) x[j][i] *= scale;
        uint4 hi, lo;
        split8(x[j], hi, lo);
        const uint32_t off = sw128(t >> 3, t & 7);
        *reinterpret_cast<uint4*>(g_hi + off) = hi;
        *reinterpret_cast<uint4*>(g_hi + kOpT + off) = lo;
    }
}

// grid (max(qtiles, ktiles), 4 H): y / H selects the tensor: 0 Qs = -log2(e) Q (+ the (D, 1/r) scalars), 1 G, 2 K, 3 V
__global__ void __launch_bounds__(256) sigmoid_bwd_prepare_kernel(const float* __restrict__ q, const float* __restrict__ g, const float* __restrict__ k,
                                                                  const float* __restrict__ v, const float* __restrict__ drow,
                                                                  const float* __restrict__ rowsum, int64_t N, int64_t L, int H, uint8_t* qimg,
                                                                  uint8_t* gimg, uint8_t* kimg, uint8_t* vimg, float* scal) {
    const int which = blockIdx.y / H, h = blockIdx.y % H;
    const int64_t t = blockIdx.x, qtiles = (N + kT - 1) / kT, ktiles = (L + kT - 1) / kT;
    if (which < 2) {
        if (t >= qtiles) return;
        convert_tile_g(which == 0 ? q : g, H, h, t * kT, N, threadIdx.x, (which == 0 ? qimg : gimg) + ((int64_t)h * qtiles + t) * kImg,
                       which == 0 ? -1.4426950408889634f : 1.f);
        if (which == 0 && threadIdx.x < kT) {
            const int64_t row = t * kT + threadIdx.x;
            float* s = scal + ((int64_t)h * qtiles + t) * (2 * kT);
            s[threadIdx.x] = row < N ? drow[row * H + h] : 0.f;
            s[kT + threadIdx.x] = row < N ? 1.f / rowsum[row * H + h] : 0.f;
        }
    } else {
        if (t >= ktiles) return;
        convert_tile_g(which == 2 ? k : v, H, h, t * kT, L, threadIdx.x, (which == 2 ? kimg : vimg) + ((int64_t)h * ktiles + t) * kImg, 1.f);
    }
}

}  // namespace

// Workspace (1 KB aligned): Qs, G images [H][qtiles] x 32 KB, K, V images [H][ktiles] x 32 KB, scalars [H][qtiles] x 1 KB,
// then the split partials.  Hv == H only (V shared by the heads accumulates dv over heads: keep the FFMA path for it).
bool sigmoid_bwd_tc_supported(int64_t N, int64_t L, int H, int Hv, int M, int D) {
    return N >= 1 && L >= 1 && Hv == H && M == kDim && D == kDim;
}

// split of the streamed side over gridDim.z (small problems: enough CTAs to fill the SMs)
int sigmoid_bwd_tc_split(int64_t own_rows, int64_t streamed_rows, int H) {
    const int64_t ctas = ((own_rows + kT - 1) / kT) * H, otiles = (streamed_rows + kT - 1) / kT;
    int64_t s = ((int64_t)sm_count() + ctas - 1) / ctas;
    if (s > otiles) s = otiles;
    if (s > 32) s = 32;
    return s < 1 ? 1 : (int)s;
}

int64_t sigmoid_bwd_tc_image_bytes(int64_t N, int64_t L, int H) {
    const int64_t qt = (N + kT - 1) / kT, kt = (L + kT - 1) / kT;
    return (2 * qt + 2 * kt) * (int64_t)H * kImg + qt * (int64_t)H * kScal + 1024;
}

// drow = g . out per (n, h) must already be in `drow` (sigmoid_drow_kernel, sigmoid.cu).  split_q / split_k as in the FFMA path;
// part_* are only used when the corresponding split > 1 (sum them with sum_partials_kernel afterwards).
int sigmoid_bwd_tc(const float* q, const float* k, const float* v, const float* g, const float* drow, const float* rowsum,
                   int64_t N, int64_t L, int H, float* dq, float* dk, float* dv, void* images, int split_k, float* part_dq,
                   int split_q, float* part_dk, float* part_dv, cudaStream_t st) {
    const int64_t qt = (N + kT - 1) / kT, kt = (L + kT - 1) / kT;
    uint8_t* qimg = (uint8_t*)(((uintptr_t)images + 1023) & ~(uintptr_t)1023);
    uint8_t* gimg = qimg + qt * H * kImg;
    uint8_t* kimg = gimg + qt * H * kImg;
    uint8_t* vimg = kimg + kt * H * kImg;
    float* scal = (float*)(vimg + kt * H * kImg);
    sigmoid_bwd_prepare_kernel<<<dim3((unsigned)(qt > kt ? qt : kt), (unsigned)(4 * H)), 256, 0, st>>>(q, g, k, v, drow, rowsum, N, L, H, qimg, gimg,
                                                                                                     kimg, vimg, scal);
    DIF_LAUNCH_OK();
    static bool attr = false;
    if (!attr) {
        DIF_CUDA_OK(cudaFuncSetAttribute(sigmoid_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBwd));
        DIF_CUDA_OK(cudaFuncSetAttribute(sigmoid_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBwd));
        attr = true;
    }
    {
        SigBwdArgs a{};
        a.own_a = qimg; a.own_b = gimg; a.str_a = kimg; a.str_b = vimg; a.drow = drow; a.rowsum = rowsum;
        a.n_own = N; a.n_str = L; a.H = H; a.split = split_k; a.out0 = dq; a.part0 = part_dq;
        sigmoid_bwd_tc_kernel<false><<<dim3((unsigned)qt, (unsigned)H, (unsigned)split_k), kBwdThreads, kSmemBwd, st>>>(a);
        DIF_LAUNCH_OK();
    }
    {
        SigBwdArgs a{};
        a.own_a = kimg; a.own_b = vimg; a.str_a = qimg; a.str_b = gimg; a.str_scal = scal;
        a.n_own = L; a.n_str = N; a.H = H; a.split = split_q; a.out0 = dv; a.out1 = dk; a.part0 = part_dv; a.part1 = part_dk;
        sigmoid_bwd_tc_kernel<true><<<dim3((unsigned)kt, (unsigned)H, (unsigned)split_q), kBwdThreads, kSmemBwd, st>>>(a);
        DIF_LAUNCH_OK();
    }
    return DIF_OK;
}

}  // namespace dif
