// kernel='sigmoid' forward on the tensor cores (sm_100a) -- full_attention_conv(..., 'sigmoid'),
// node classification/difformer.py:45-56:   P = sigmoid(Q K^T),  out = (P V) / rowsum(P).
//
// Flash-style: one CTA owns 128 query rows of one head (and one slice of the keys), loops over 128-key tiles and
// never materialises [N, L]:
//   S  = Q K^T      tcgen05.mma M=128 N=128 K=64, Q and K split into bf16 hi+lo (3 MMAs per product: the scores go
//                   through a sigmoid, so they need ~fp32 accuracy), accumulator in TMEM (two S buffers)
//   P  = sigmoid(S) 256 threads (thread = query row x half of the keys): tcgen05.ld -> ex2/rcp -> bf16 hi+lo ->
//                   tcgen05.st back into the S buffer it came from (P in (0,1): no running max).  P is split like
//                   the other operands: the output is a mean of signed values, so a 2^-9 rounding of the weights
//                   would show up at the 1e-3 level.
//   O += P V        tcgen05.mma M=128 N=128 K=128, A = P hi/lo read from TENSOR MEMORY, B = [Vhi | Vlo] (shared
//                   memory, MN-major: keys are the K index); the epilogue adds the two 64-column halves
//
// What bounds it: with every operand in shared memory the kernel was limited by the 128 B/clk shared-memory port
// (per 128x128 tile: 240 KB of MMA operand reads + 64 KB of P stores + 64 KB of TMA writes = 2900 clk against
// 2050 clk of MUFU and 1540 clk of tensor math).  So the A operands live in tensor memory: P is written over its
// own S buffer (32-bit column = two consecutive keys; every 32-key chunk becomes 16 columns hi | 16 columns lo), and Q hi/lo
// (constant per CTA, pre-scaled by -log2 e) sits in 64 more columns.  Shared memory then only carries the K / V
// tiles: 64 KB in by TMA and 96 KB of B-operand reads per tile.
//
// K and V are converted ONCE per call into the operand images the MMA reads (bf16 hi|lo, 128B-swizzled, one 32 KB
// block per 128-key tile and head: sigmoid_prepare_kernel), so the main kernel streams them with one TMA bulk copy
// per tile instead of re-converting them for every query tile.
// Warp roles: 0-7 sigmoid, 8 K loader, 9 V loader, 10 MMA issuer.  The MMA thread issues S(j+1) before P V(j), so
// the sigmoid of tile j overlaps S(j+1) and P V(j-1).  tcgen05.mma instructions of one CTA execute in issue order,
// which is what makes the S/P aliasing safe: S(j+2) is issued after P V(j), the last reader of that buffer.
// Small N: the key range is split over gridDim.z and the un-normalised partials are combined in fixed order
// (sigmoid.cu), like the FFMA kernel.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <type_traits>

namespace dif {
namespace {

#ifndef DIF_SIG_PAIR
#define DIF_SIG_PAIR 1
#endif
constexpr int kT = 128;                  // query rows per CTA = keys per tile
constexpr int kOpT = kT * 128;           // one bf16 [128 rows][64] operand tile: 16 KB
constexpr int kImg = 2 * kOpT;           // hi | lo image of one tile: 32 KB
constexpr int kSigWarps = 11, kSigThreads = kSigWarps * 32;
constexpr int kSmemSig = 2 * kImg + 2 * kImg + 1024;       // 2 K stages + 2 V stages
constexpr int kColO = 2 * kT, kColQ = 3 * kT;              // TMEM columns: S0/P0 | S1/P1 | O (P Vhi | P Vlo) | Qhi (32) | Qlo (32)

// optional timeline (DIF_SIG_DEBUG=1): clock64 stamps of CTA 0: [role][tile][slot]; roles: sigmoid warp 0, MMA thread
#define DIF_SIG_STAMP(slot)                                                                                     \
    do {                                                                                                        \
        if (p.dbg != nullptr && tid == 0 && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && i < 16) p.dbg[i * 4 + (slot)] = clock64(); \
    } while (0)
#define DIF_MMA_STAMP(slot)                                                                                     \
    do {                                                                                                        \
        if (p.dbg != nullptr && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && i < 16) p.dbg[64 + i * 4 + (slot)] = clock64(); \
    } while (0)

struct SigTcArgs {
    long long* dbg;
    const float* q;
    const uint8_t *kimg, *vimg;   // [H | Hv][ltiles] x 32 KB operand images
    int64_t N, L;
    int H, Hv, ksplit;
    float *out, *rowsum;     // ksplit == 1: final results
    float *pout, *prs;       // ksplit  > 1: [ksplit][N,H,64] un-normalised sums, [ksplit][N,H] row sums
};

// Eight sigmoids from pre-scaled scores x = -s log2(e) (the scale is folded into Q): p = 1 / (1 + 2^x), two MUFU
// ops per element and branch-free (ex2 overflow -> +inf -> rcp -> 0; underflow -> 1).  Everything else is packed
// f32x2 arithmetic (FADD2 / FFMA2): the sigmoid warps are issue-bound, not MUFU-bound.
template <bool kMasked>
__device__ __forceinline__ void sigmoid8(const uint32_t* x, int nvalid, float2 (&rs)[2], uint32_t* h, uint32_t* l) {
    float2 pr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float2 e;
#if DIF_SIG_PAIR
        // three MUFU ops for two sigmoids: 1/(1+a) = (1+b) / ((1+a)(1+b)); x is clamped at 62 so that the product
        // (1+a)(1+b) <= 2^124 stays finite (sigmoid < 2e-19 there)
        float r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(fminf(__uint_as_float(x[2 * k]), 62.f)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(fminf(__uint_as_float(x[2 * k + 1]), 62.f)));
        e = __fadd2_rn(e, make_float2(1.f, 1.f));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e.x * e.y));
        pr[k] = make_float2(e.y * r, e.x * r);
#else
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(__uint_as_float(x[2 * k])));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(__uint_as_float(x[2 * k + 1])));
        e = __fadd2_rn(e, make_float2(1.f, 1.f));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(pr[k].x) : "f"(e.x));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(pr[k].y) : "f"(e.y));
#endif
        if (kMasked) {
            pr[k].x = 2 * k < nvalid ? pr[k].x : 0.f;
            pr[k].y = 2 * k + 1 < nvalid ? pr[k].y : 0.f;
        }
        rs[k & 1] = __fadd2_rn(rs[k & 1], pr[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        h[k] = bf2_bits(pr[k].x, pr[k].y);
        const float2 hf = make_float2(__uint_as_float(h[k] << 16), __uint_as_float(h[k] & 0xffff0000u));
        const float2 d = __ffma2_rn(hf, make_float2(-1.f, -1.f), pr[k]);      // exact: p - bf16(p)
        l[k] = bf2_bits(d.x, d.y);
    }
}

// [128 rows][64 floats] of one head -> bf16 hi/lo image (rows of 128 B, 8-row swizzle atoms).  The same layout
// serves Q / K as K-major operands and V as the MN-major B operand (keys = K index).  256 threads.
template <bool kToShared>
__device__ __forceinline__ void convert_tile(const float* src, int heads, int head, int64_t row0, int64_t nrows, int tid,
                                             uint32_t s_hi, uint8_t* g_hi, float scale = 1.f) {
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
        uint4 hi, lo;
        if (kToShared) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[j][i] *= scale;
        }
        split8(x[j], hi, lo);
        const uint32_t off = sw128(t >> 3, t & 7);
        if (kToShared) {
            sts128(s_hi + off, hi);
            sts128(s_hi + kOpT + off, lo);
        } else {
            *reinterpret_cast<uint4*>(g_hi + off) = hi;
            *reinterpret_cast<uint4*>(g_hi + kOpT + off) = lo;
        }
    }
}

// grid (ltiles, H + Hv): K heads first, then V heads
__global__ void __launch_bounds__(256) sigmoid_prepare_kernel(const float* __restrict__ k, const float* __restrict__ v, int64_t L, int H,
                                                              int Hv, uint8_t* __restrict__ kimg, uint8_t* __restrict__ vimg) {
    const int64_t t = blockIdx.x, ltiles = gridDim.x;
    const int y = blockIdx.y;
    if (y < H) convert_tile<false>(k, H, y, t * kT, L, threadIdx.x, 0, kimg + ((int64_t)y * ltiles + t) * kImg);
    else convert_tile<false>(v, Hv, y - H, t * kT, L, threadIdx.x, 0, vimg + ((int64_t)(y - H) * ltiles + t) * kImg);
}

__global__ void __launch_bounds__(kSigThreads, 1) sigmoid_fwd_tc_kernel(const __grid_constant__ SigTcArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* Kst = base;                      // 2 x (Khi | Klo)
    uint8_t* Vst = Kst + 2 * kImg;            // 2 x (Vhi | Vlo)
    __shared__ uint64_t qfull, kfull[2], kempty[2], vfull[2], vempty[2], sfull[2], pfull[2], done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.H, h = blockIdx.y, hv = (p.Hv == H) ? h : 0;
    const int64_t n0 = (int64_t)blockIdx.x * kT;
    const int64_t ltiles = (p.L + kT - 1) / kT, per = (ltiles + p.ksplit - 1) / p.ksplit;
    const int64_t t0 = (int64_t)blockIdx.z * per, t1 = min(ltiles, t0 + per);
    const int T = (int)max((int64_t)0, t1 - t0);        // key tiles of this CTA

    if (tid == 0) {
        mbar_init(&qfull, 8);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&kfull[s], 1); mbar_init(&kempty[s], 1); mbar_init(&vfull[s], 1); mbar_init(&vempty[s], 1);
            mbar_init(&sfull[s], 1); mbar_init(&pfull[s], 8);
        }
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 10) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp < 8) {
        // ===== sigmoid: thread = (query row r, key half)
        const int quad = warp & 3, half = warp >> 2, r = quad * 32 + lane;
        const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
        const int64_t row = n0 + r;
        {   // Q row -> -log2(e) q as bf16 hi (key half 0 warps) / lo (half 1 warps), 32 packed columns each, into TMEM
            uint32_t w[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float x[8];
                if (row < p.N) ldg256_keep(p.q + (row * H + h) * kDim + c * 8, x);
                else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[e] = 0.f;
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] *= -1.4426950408889634f;
                uint4 hi, lo;
                split8(x, hi, lo);
                const uint4 u = half == 0 ? hi : lo;
                w[4 * c] = u.x; w[4 * c + 1] = u.y; w[4 * c + 2] = u.z; w[4 * c + 3] = u.w;
            }
            tmem_st32(tlane + kColQ + half * 32, w);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&qfull);
        }
        float2 rs[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        for (int i = 0; i < T; ++i) {
            const int sbuf = i & 1;
            const uint32_t tcol = tlane + sbuf * kT + half * 64;
            const int valid = (int)min((int64_t)kT, p.L - (t0 + i) * kT) - half * 64;      // keys of this half that exist
            mbar_wait(&sfull[sbuf], (i >> 1) & 1);
            tc_fence_after();
            DIF_SIG_STAMP(0);
            uint32_t x0[32], x1[32];
            tmem_ld32(tcol, x0);
            tmem_ld32(tcol + 32, x1);
            tmem_ld_wait32(x0);
            tmem_ld_wait32(x1);
            DIF_SIG_STAMP(1);
            // Each 32-key chunk writes its P words over its OWN 32 score columns -- [16 words hi | 16 words lo] -- so
            // no thread ever overwrites scores that another thread has not read yet.
            auto chunk = [&](uint32_t (&x)[32], int c) {
                uint32_t w[32];
                if (valid - 32 * c >= 32) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) sigmoid8<false>(&x[j], 0, rs, &w[j >> 1], &w[16 + (j >> 1)]);
                } else {                                        // only the last key tile has keys >= L
#pragma unroll
                    for (int j = 0; j < 32; j += 8) sigmoid8<true>(&x[j], valid - 32 * c - j, rs, &w[j >> 1], &w[16 + (j >> 1)]);
                }
                tmem_st32(tcol + 32 * c, w);
            };
            chunk(x0, 0);
            chunk(x1, 1);
            DIF_SIG_STAMP(2);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&pfull[sbuf]);
            DIF_SIG_STAMP(3);
        }
        // ---- O / rowsum: the two key halves of a row exchange their partial row sums through (now free) K memory
        mbar_wait(&done, 0);
        tc_fence_after();
        float* rsx = reinterpret_cast<float*>(Kst);
        rsx[half * kT + r] = (rs[0].x + rs[0].y) + (rs[1].x + rs[1].y);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float rowsum = rsx[r] + rsx[kT + r];
        float* o_dst = p.ksplit > 1 ? p.pout + (int64_t)blockIdx.z * p.N * H * kDim : p.out;
        float* r_dst = p.ksplit > 1 ? p.prs + (int64_t)blockIdx.z * p.N * H : p.rowsum;
        const float inv = p.ksplit > 1 ? 1.f : 1.f / rowsum;
        uint32_t o[32], o2[32];                           // this thread: output columns [half * 32, +32) of P Vhi and P Vlo
        if (T > 0) {
            tmem_ld32(tlane + kColO + half * 32, o);
            tmem_ld32(tlane + kColO + kDim + half * 32, o2);
            tmem_ld_wait32(o);
            tmem_ld_wait32(o2);
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) + __uint_as_float(o2[j]));
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = 0u;
        }
        if (row < p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o_dst + (row * H + h) * kDim + half * 32 + j) =
                    make_float4(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv, __uint_as_float(o[j + 2]) * inv,
                                __uint_as_float(o[j + 3]) * inv);
            if (half == 0) r_dst[row * H + h] = rowsum;
        }
    } else if (warp == 8 || warp == 9) {
        // ===== loaders: one 32 KB bulk copy per tile (warp 8: K images, warp 9: V images)
        if (lane == 0) {
            const bool isk = warp == 8;
            const uint8_t* img = isk ? p.kimg + ((int64_t)h * ltiles + t0) * kImg : p.vimg + ((int64_t)hv * ltiles + t0) * kImg;
            uint64_t* full = isk ? kfull : vfull;
            uint64_t* empty = isk ? kempty : vempty;
            const uint32_t dst = smem_u32(isk ? Kst : Vst);
            for (int i = 0; i < T; ++i) {
                const int s = i & 1;
                if (i >= 2) mbar_wait(&empty[s], ((i >> 1) - 1) & 1);
                mbar_expect_tx(&full[s], kImg);
                tma_load_1d(dst + s * kImg, img + (int64_t)i * kImg, kImg, &full[s]);
            }
        }
    } else {
        // ===== MMA issuer (A operands from tensor memory): the whole warp runs the loop, one elected lane issues
        const bool leader = elect_one();
        const uint32_t idS = make_idesc(kT, kT, 0, 0);          // S = Q K^T : both operands K-major
        const uint32_t idO = make_idesc(kT, 2 * kDim, 0, 1);    // [P Vhi | P Vlo] += P [Vhi | Vlo] : A = P (TMEM), B = V MN-major, N = 128
        auto issue_S = [&](int i) {
            const int s = i & 1;
            mbar_wait(&kfull[s], (i >> 1) & 1);
            tc_fence_after();
            const uint32_t sb = smem_u32(Kst) + s * kImg;
            if (!leader) return;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {                     // 16 features per step = 8 packed TMEM columns of Q
                const uint32_t qhi = tmem + kColQ + ks * 8, qlo = tmem + kColQ + 32 + ks * 8;
                const uint64_t khi = make_desc(sb + ks * 32, kKmajLBO, kKmajSBO), klo = make_desc(sb + kOpT + ks * 32, kKmajLBO, kKmajSBO);
                umma_ts(tmem + s * kT, qhi, khi, idS, ks > 0 ? 1u : 0u);
                umma_ts(tmem + s * kT, qlo, khi, idS, 1u);
                umma_ts(tmem + s * kT, qhi, klo, idS, 1u);
            }
            umma_commit(&sfull[s]);
            umma_commit(&kempty[s]);
        };
        auto issue_PV = [&](int i) {
            const int s = i & 1;
            mbar_wait(&vfull[s], (i >> 1) & 1);
            mbar_wait(&pfull[s], (i >> 1) & 1);
            tc_fence_after();
            if (!leader) return;
            DIF_MMA_STAMP(2);
            const uint32_t sb = smem_u32(Vst) + s * kImg, pt = tmem + s * kT;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {                     // 16 keys per step = 8 packed TMEM columns of P
                // B = [Vhi | Vlo]: the two 64-wide MN blocks of the V image (16 KB apart = LBO) as ONE N = 128 operand.
                // An MMA with A in tensor memory costs 64 clk whatever N is (4 KB of A at 64 B/clk), so two N = 128
                // MMAs per step (P hi, P lo) replace three N = 64 ones; the epilogue adds the two 64-column halves.
                const uint64_t vb = make_desc(sb + ks * 2048, kOpT, 1024);
                const uint32_t ph = pt + (ks >> 1) * 32 + (ks & 1) * 8;      // 32-key chunk: [16 words hi | 16 words lo]
                umma_ts(tmem + kColO, ph, vb, idO, (i > 0 || ks > 0) ? 1u : 0u);
                umma_ts(tmem + kColO, ph + 16, vb, idO, 1u);
            }
            umma_commit(&vempty[s]);
        };
        mbar_wait(&qfull, 0);
        tc_fence_after();
        if (T > 0) issue_S(0);
        for (int i = 0; i < T; ++i) {
            if (leader) DIF_MMA_STAMP(0);
            if (i + 1 < T) issue_S(i + 1);
            if (leader) DIF_MMA_STAMP(1);
            issue_PV(i);
            if (leader) DIF_MMA_STAMP(3);
            __syncwarp();
        }
        if (leader) { if (T > 0) umma_commit(&done); else mbar_arrive(&done); }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 10) tmem_dealloc(tmem, 512);
}

}  // namespace

bool sigmoid_tc_supported(int64_t N, int64_t L, int H, int Hv, int M, int D) {
    return N >= 1 && L >= 1 && H >= 1 && H + Hv <= 65535 && (Hv == H || Hv == 1) && M == kDim && D == kDim;
}

// Key split: CTAs = query tiles x heads x s run in waves of one CTA per SM (225 KB of shared memory each); pick the
// s that minimises waves x (key tiles per CTA + fixed per-CTA cost), with a small charge for the combine pass.
int sigmoid_tc_ksplit(int64_t N, int64_t L, int H) {
    const int64_t ctas = ((N + kT - 1) / kT) * H, ltiles = (L + kT - 1) / kT, sms = sm_count();
    int best = 1;
    double best_cost = 1e300;
    for (int64_t s = 1; s <= ltiles && s <= 32; ++s) {
        const int64_t waves = (ctas * s + sms - 1) / sms, per = (ltiles + s - 1) / s;
        const double cost = (double)waves * ((double)per + 1.5) + (s > 1 ? 0.5 + 0.1 * (double)s : 0.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = (int)s; }
    }
    return best;
}

// bytes of the K / V operand images (after the key-split partials in the workspace), incl. 1 KB alignment slack
int64_t sigmoid_tc_image_bytes(int64_t L, int H, int Hv) { return ((L + kT - 1) / kT) * (int64_t)(H + Hv) * kImg + 1024; }

int sigmoid_fwd_tc(const float* q, const float* k, const float* v, int64_t N, int64_t L, int H, int Hv,
                   float* out, float* rowsum, float* pout, float* prs, int ksplit, void* images, cudaStream_t st) {
    DIF_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) & 31) == 0 && ((uintptr_t)out & 15) == 0, DIF_EARG,
                "sigmoid(tcgen05): q/k/v must be 32-byte aligned");
    const int64_t ltiles = (L + kT - 1) / kT;
    uint8_t* kimg = (uint8_t*)(((uintptr_t)images + 1023) & ~(uintptr_t)1023);
    uint8_t* vimg = kimg + ltiles * H * kImg;
    sigmoid_prepare_kernel<<<dim3((unsigned)ltiles, (unsigned)(H + Hv)), 256, 0, st>>>(k, v, L, H, Hv, kimg, vimg);
    DIF_LAUNCH_OK();
    SigTcArgs a{};
    a.q = q; a.kimg = kimg; a.vimg = vimg; a.N = N; a.L = L; a.H = H; a.Hv = Hv; a.ksplit = ksplit;
    a.out = out; a.rowsum = rowsum; a.pout = pout; a.prs = prs;
    static long long* dbg = nullptr;
    static const bool want_dbg = getenv("DIF_SIG_DEBUG") != nullptr;
    if (want_dbg && dbg == nullptr) { DIF_CUDA_OK(cudaMalloc(&dbg, 128 * sizeof(long long))); DIF_CUDA_OK(cudaMemset(dbg, 0, 128 * sizeof(long long))); }
    a.dbg = dbg;
    static bool attr = false;
    if (!attr) { DIF_CUDA_OK(cudaFuncSetAttribute(sigmoid_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemSig)); attr = true; }
    dim3 grid((unsigned)((N + kT - 1) / kT), (unsigned)H, (unsigned)ksplit);
    sigmoid_fwd_tc_kernel<<<grid, kSigThreads, kSmemSig, st>>>(a);
    DIF_LAUNCH_OK();
    if (want_dbg) {            // debugging aid only: synchronises
        long long hbuf[128];
        DIF_CUDA_OK(cudaStreamSynchronize(st));
        DIF_CUDA_OK(cudaMemcpy(hbuf, dbg, sizeof(hbuf), cudaMemcpyDeviceToHost));
        const long long t00 = hbuf[0];
        for (int i = 0; i < 16; ++i)
            fprintf(stderr, "sig tile %2d  sigmoid: sfull %6lld ld %6lld math %6lld signalled %6lld | mma: top %6lld S(i+1) issued %6lld pfull %6lld PV issued %6lld\n", i,
                    hbuf[i * 4] - t00, hbuf[i * 4 + 1] - t00, hbuf[i * 4 + 2] - t00, hbuf[i * 4 + 3] - t00, hbuf[64 + i * 4] - t00,
                    hbuf[64 + i * 4 + 1] - t00, hbuf[64 + i * 4 + 2] - t00, hbuf[64 + i * 4 + 3] - t00);
    }
    return DIF_OK;
}

}  // namespace dif
