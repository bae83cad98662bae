// kernel='sigmoid' forward on the tensor cores (sm_100a) -- full_attention_conv(..., 'sigmoid'),
// node classification/difformer.py:45-56:   P = sigmoid(Q K^T),  out = (P V) / rowsum(P).
//
// Flash-style: one CTA owns 128 query rows of one head (and one slice of the keys), loops over 128-key tiles and
// never materialises [N, L]:
//   S  = Q K^T      tcgen05.mma M=128 N=128 K=64, Q and K split into bf16 hi+lo (3 MMAs per product: the scores go
//                   through a sigmoid, so they need ~fp32 accuracy), accumulator in TMEM (two S buffers)
//   P  = sigmoid(S) 256 threads (thread = query row x half of the keys): tcgen05.ld -> ex2/rcp -> bf16 hi+lo ->
//                   128B-swizzled shared memory (P in (0,1): no running max).  P is split like the other operands:
//                   the output is a mean of signed values, so a 2^-9 rounding of the weights would show up at the
//                   1e-3 level.
//   O += P V        tcgen05.mma M=128 N=64 K=128, A = P hi/lo (K-major), B = V hi/lo (MN-major: keys are the K index)
//
// K and V are converted ONCE per call into the operand images the MMA reads (bf16 hi|lo, 128B-swizzled, one 32 KB
// block per 128-key tile and head: sigmoid_prepare_kernel), so the main kernel streams them with one TMA bulk copy
// per tile instead of re-converting them for every query tile.
// Warp roles: 0-7 sigmoid, 8 K loader, 9 V loader, 10 MMA issuer.  The sigmoid (MUFU-bound) is the critical
// resource, so shared memory goes to a double-buffered P (the sigmoid of tile j+1 never waits for the P V product of
// tile j) and K / V get one stage each: the K stage is released as soon as S = QK^T is done, the V stage after P V,
// and the next tile streams in while the sigmoid runs.  The MMA thread issues S(j+1) before P V(j).  Small N: the key range is split over gridDim.z and the un-normalised partials are
// combined in fixed order (sigmoid.cu), like the FFMA kernel.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <type_traits>

namespace dif {
namespace {

constexpr int kT = 128;                  // query rows per CTA = keys per tile
constexpr int kOpT = kT * 128;           // one bf16 [128 rows][64] operand tile: 16 KB
constexpr int kImg = 2 * kOpT;           // hi | lo image of one tile: 32 KB
constexpr int kSigWarps = 11, kSigThreads = kSigWarps * 32;
constexpr int kPBuf = 2 * kImg;          // one P buffer: Phi (keys 0-63 | 64-127) | Plo (same)
constexpr int kSmemSig = kImg + kImg + kImg + 2 * kPBuf + 1024;          // Q + K + V + 2 P buffers

struct SigTcArgs {
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
__device__ __forceinline__ void sigmoid8(const uint32_t* x, int nvalid, float2 (&rs)[2], uint4& hi, uint4& lo) {
    float2 pr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float2 e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(__uint_as_float(x[2 * k])));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(__uint_as_float(x[2 * k + 1])));
        e = __fadd2_rn(e, make_float2(1.f, 1.f));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(pr[k].x) : "f"(e.x));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(pr[k].y) : "f"(e.y));
        if (kMasked) {
            pr[k].x = 2 * k < nvalid ? pr[k].x : 0.f;
            pr[k].y = 2 * k + 1 < nvalid ? pr[k].y : 0.f;
        }
        rs[k & 1] = __fadd2_rn(rs[k & 1], pr[k]);
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        h[k] = bf2_bits(pr[k].x, pr[k].y);
        const float2 hf = make_float2(__uint_as_float(h[k] << 16), __uint_as_float(h[k] & 0xffff0000u));
        const float2 d = __ffma2_rn(hf, make_float2(-1.f, -1.f), pr[k]);      // exact: p - bf16(p)
        l[k] = bf2_bits(d.x, d.y);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
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
    uint8_t* Qop = base;                      // Qhi | Qlo
    uint8_t* Kst = Qop + kImg;                // Khi | Klo   (one stage: the next tile is fetched while the sigmoid runs)
    uint8_t* Vst = Kst + kImg;                // Vhi | Vlo
    uint8_t* Pop = Vst + kImg;                // 2 x (Phi keys 0-63 | Phi keys 64-127 | Plo keys 0-63 | Plo keys 64-127)
    __shared__ uint64_t qfull, kfull, kempty, vfull, vempty, sfull[2], sempty[2], pfull[2], pempty[2], done;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.H, h = blockIdx.y, hv = (p.Hv == H) ? h : 0;
    const int64_t n0 = (int64_t)blockIdx.x * kT;
    const int64_t ltiles = (p.L + kT - 1) / kT, per = (ltiles + p.ksplit - 1) / p.ksplit;
    const int64_t t0 = (int64_t)blockIdx.z * per, t1 = min(ltiles, t0 + per);
    const int T = (int)max((int64_t)0, t1 - t0);        // key tiles of this CTA

    if (tid == 0) {
        mbar_init(&qfull, 8);
        mbar_init(&kfull, 1); mbar_init(&kempty, 1); mbar_init(&vfull, 1); mbar_init(&vempty, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&sfull[s], 1); mbar_init(&sempty[s], 8); mbar_init(&pfull[s], 8); mbar_init(&pempty[s], 1); }
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 10) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp < 8) {
        // ===== sigmoid: thread = (query row r, key half); S row -> p = sigmoid(s) -> bf16 hi/lo P row
        const int quad = warp & 3, half = warp >> 2, r = quad * 32 + lane;
        convert_tile<true>(p.q, H, h, n0, p.N, tid, smem_u32(Qop), nullptr, -1.4426950408889634f);   // S = -log2(e) Q K^T
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&qfull);

        float2 rs[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        for (int i = 0; i < T; ++i) {
            const int sbuf = i & 1;
            const int64_t l0 = (t0 + i) * kT;
            const uint32_t pbase = smem_u32(Pop) + sbuf * kPBuf + half * kOpT;
            mbar_wait(&sfull[sbuf], (i >> 1) & 1);
            tc_fence_after();
            if (i >= 2) mbar_wait(&pempty[sbuf], ((i >> 1) - 1) & 1);    // P V of tile i-2 has consumed this P buffer
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + sbuf * kT + half * 64;
            const int valid = (int)min((int64_t)kT, p.L - l0) - half * 64;      // keys of this half that exist
            auto tile = [&](auto masked_t) {
                constexpr bool kMasked = decltype(masked_t)::value;       // only the last key tile has keys >= L
                uint32_t sa[32], sb[32];
                tmem_ld32(taddr, sa);
                tmem_ld32(taddr + 32, sb);
                tmem_ld_wait32(sa);
                tmem_ld_wait32(sb);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sempty[sbuf]);      // S is in registers: the MMA may overwrite this buffer
#pragma unroll
                for (int j = 0; j < 64; j += 8) {
                    uint4 hi, lo;
                    sigmoid8<kMasked>(j < 32 ? &sa[j & 31] : &sb[j & 31], valid - j, rs, hi, lo);
                    const uint32_t off = sw128(r, j >> 3);      // 16-byte chunk of this half's 64-key row
                    sts128(pbase + off, hi);
                    sts128(pbase + 2 * kOpT + off, lo);
                }
            };
            if (valid >= 64) tile(std::false_type{}); else tile(std::true_type{});
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&pfull[sbuf]);
        }
        // ---- O / rowsum: the two key halves of a row exchange their partial row sums through (now free) P memory
        mbar_wait(&done, 0);
        tc_fence_after();
        float* rsx = reinterpret_cast<float*>(Pop);
        rsx[half * kT + r] = (rs[0].x + rs[0].y) + (rs[1].x + rs[1].y);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float rowsum = rsx[r] + rsx[kT + r];
        const int64_t row = n0 + r;
        float* o_dst = p.ksplit > 1 ? p.pout + (int64_t)blockIdx.z * p.N * H * kDim : p.out;
        float* r_dst = p.ksplit > 1 ? p.prs + (int64_t)blockIdx.z * p.N * H : p.rowsum;
        const float inv = p.ksplit > 1 ? 1.f : 1.f / rowsum;
        uint32_t o[32];                                   // this thread: output columns [half * 32, +32)
        if (T > 0) {
            tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + 2 * kT + half * 32, o);
            tmem_ld_wait32(o);
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
            uint64_t* full = isk ? &kfull : &vfull;
            uint64_t* empty = isk ? &kempty : &vempty;
            const uint32_t dst = smem_u32(isk ? Kst : Vst);
            for (int i = 0; i < T; ++i) {
                if (i >= 1) mbar_wait(empty, (i - 1) & 1);
                mbar_expect_tx(full, kImg);
                tma_load_1d(dst, img + (int64_t)i * kImg, kImg, full);
            }
        }
    } else if (lane == 0) {
        // ===== MMA issuer
        const uint32_t idS = make_idesc(kT, kT, 0, 0);          // S = Q K^T : both operands K-major
        const uint32_t idO = make_idesc(kT, kDim, 0, 1);        // O += P V  : A = P K-major, B = V MN-major
        const uint32_t qb = smem_u32(Qop);
        auto issue_S = [&](int i) {
            const int s = i & 1;
            mbar_wait(&kfull, i & 1);
            if (i >= 2) mbar_wait(&sempty[s], ((i >> 1) - 1) & 1);
            tc_fence_after();
            const uint32_t sb = smem_u32(Kst);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t qhi = make_desc(qb + ks * 32, kKmajLBO, kKmajSBO), qlo = make_desc(qb + kOpT + ks * 32, kKmajLBO, kKmajSBO);
                const uint64_t khi = make_desc(sb + ks * 32, kKmajLBO, kKmajSBO), klo = make_desc(sb + kOpT + ks * 32, kKmajLBO, kKmajSBO);
                umma(tmem + s * kT, qhi, khi, idS, ks > 0 ? 1u : 0u);
                umma(tmem + s * kT, qlo, khi, idS, 1u);
                umma(tmem + s * kT, qhi, klo, idS, 1u);
            }
            umma_commit(&sfull[s]);
            umma_commit(&kempty);
        };
        auto issue_PV = [&](int i) {
            const int s = i & 1;
            mbar_wait(&vfull, i & 1);
            mbar_wait(&pfull[s], (i >> 1) & 1);
            tc_fence_after();
            const uint32_t sb = smem_u32(Vst), pb = smem_u32(Pop) + s * kPBuf;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {                     // 16 keys per step
                const uint64_t pa = make_desc(pb + (ks >> 2) * kOpT + (ks & 3) * 32, kKmajLBO, kKmajSBO);
                const uint64_t pl = make_desc(pb + (2 + (ks >> 2)) * kOpT + (ks & 3) * 32, kKmajLBO, kKmajSBO);
                const uint64_t vhi = make_desc(sb + ks * 2048, kOpT, 1024), vlo = make_desc(sb + kOpT + ks * 2048, kOpT, 1024);
                umma(tmem + 2 * kT, pa, vhi, idO, (i > 0 || ks > 0) ? 1u : 0u);
                umma(tmem + 2 * kT, pa, vlo, idO, 1u);
                umma(tmem + 2 * kT, pl, vhi, idO, 1u);
            }
            umma_commit(&pempty[s]);
            umma_commit(&vempty);
        };
        mbar_wait(&qfull, 0);
        if (T > 0) issue_S(0);
        for (int i = 0; i < T; ++i) {
            if (i + 1 < T) issue_S(i + 1);
            issue_PV(i);
        }
        if (T > 0) umma_commit(&done); else mbar_arrive(&done);
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
    static bool attr = false;
    if (!attr) { DIF_CUDA_OK(cudaFuncSetAttribute(sigmoid_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemSig)); attr = true; }
    dim3 grid((unsigned)((N + kT - 1) / kT), (unsigned)H, (unsigned)ksplit);
    sigmoid_fwd_tc_kernel<<<grid, kSigThreads, kSmemSig, st>>>(a);
    DIF_LAUNCH_OK();
    return DIF_OK;
}

}  // namespace dif
