// kernel='simple' -- tcgen05 / TMEM path for sm_100a (H = 4, M = D = 64, fp32 in / fp32 out).
//
// Reference path replaced: full_attention_conv(..., 'simple'), node classification/difformer.py:18-39.
//
// Both contractions of the O(N) kernel run on the 5th-gen tensor cores:
//   pass 1  S_h  = K_h^T V_h      contraction over NODES  -> operands MN-major, M = N = 128 (two heads
//                                  stacked; the two diagonal 64x64 blocks of D are S_{2p}, S_{2p+1})
//   pass 2  q_n (c S_h | c z_h)   contraction over m      -> operands K-major, M = 128 rows, N = 80
//                                  (64 columns of S plus the z column = the denominator)
// Inputs are fp32; the reference tolerance (1e-3, also on the intermediates) rules out TF32
// (truncation bias ~1e-3), so every fp32 value x is split on the fly into bf16 hi + bf16 lo
// (x = hi + lo + O(2^-17 x)) and each product uses 3 MMAs (hi*hi + hi*lo + lo*hi), error ~2^-16.
// The split runs on the CUDA cores of the warps that stream the rows from HBM (256-bit loads),
// which write the bf16 tiles straight into the 128B-swizzled UMMA layout in shared memory;
// an elected thread issues tcgen05.mma with the accumulators in TMEM; mbarrier rings couple
// producers -> MMA -> epilogue.  The tensor pipe needs ~20% of the HBM time, so the kernels are
// HBM-bound by design (roofline: 4*H*D*4 B per node).
//
// Warp roles (13 warps, 1 CTA per SM, persistent over contiguous row ranges / 128-row tiles):
//   pass 1: warps 0-7 K/V producers (+ sum k, sum v, sum k^2), warps 8-11 stream Q for sum q^2,
//           warp 12 MMA issuer; epilogue: warps 0-3 TMEM -> per-CTA record (deterministic 2-stage reduce)
//   pass 2: warps 0-7 Q producers, warps 8-11 epilogue (TMEM -> registers -> (acc+u)/den -> HBM,
//           optionally the fused layer epilogue), warp 12 MMA issuer.
#include <cuda.h>        // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)
#include <cuda_bf16.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace dif {

int simple_finalize_fwd(const float* ws, int nchunks, int H, int Hv, int M, int D, float* partials, cudaStream_t st);

namespace {

constexpr int kH = 4;
constexpr int kDim = 64;
constexpr int kRowF = kH * kDim;      // floats per node row (256)
constexpr int kWarps = 13;
constexpr int kThreadsTC = kWarps * 32;

// ---- descriptor conventions (verified on hardware by csrc/probe_umma.cu) --------------------
constexpr uint32_t kSwizzle128 = 2;
// K-major SW128: 8-row groups 1024 B apart (SBO); LBO unused
constexpr uint32_t kKmajLBO = 0, kKmajSBO = 1024;
// MN-major SW128: 64-element (128 B) MN blocks are LBO apart, 8-k groups SBO apart
#ifndef DIF_MN_LBO_IS_MNSTRIDE
#define DIF_MN_LBO_IS_MNSTRIDE 1
#endif

// ---- optional in-kernel timeline (DIF_TC_DEBUG_TIMES=1): thread 0 of every CTA stamps %globaltimer
__device__ __forceinline__ uint64_t gtime() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define DIF_STAMP(buf, slot)                                                        \
    do {                                                                            \
        if ((buf) != nullptr && threadIdx.x == 0) (buf)[blockIdx.x * 8 + (slot)] = gtime(); \
    } while (0)

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(bar);
    while (!done) {
        asm volatile("{\n\t.reg .pred pq;\n\tmbarrier.try_wait.parity.shared::cta.b64 pq, [%1], %2;\n\tselp.b32 %0, 1, 0, pq;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(cols) : "memory");
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;              // descriptor version: Blackwell
    d |= (uint64_t)kSwizzle128 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;                        // D format f32
    d |= 1u << 7;                        // A format bf16
    d |= 1u << 10;                       // B format bf16
    d |= (uint32_t)a_mn << 15;           // A major: 0 = K, 1 = MN
    d |= (uint32_t)b_mn << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pp;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    return r;
}
// tcgen05.ld is asynchronous: its destination registers are only valid after wait::ld.  The
// registers are threaded through the wait as read-write operands so the compiler cannot schedule
// a consumer above it.
__device__ __forceinline__ void tmem_ld_wait32(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                   "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                   "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}
__device__ __forceinline__ void tmem_ld_wait1(uint32_t& r) { asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r) :: "memory"); }

// streaming 256-bit global load (data is consumed once: no L1 allocation, evict-first in L2)
__device__ __forceinline__ void ldg256_stream(const float* p, float (&r)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p));
}
// 256-bit load that may stay in L2 (pass 1 reads Q only for its norm; pass 2 reads it again)
__device__ __forceinline__ void ldg256_keep(const float* p, float (&r)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p));
}
// fire-and-forget bulk prefetch of a contiguous global range into L2 (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l2_hint(const void* p, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" :: "l"(p), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t policy_evict_first_() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 256-bit load with an explicit L2 eviction policy
__device__ __forceinline__ void ldg256_policy(const float* p, float (&r)[8], uint64_t policy) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p), "l"(policy));
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// TMA 2-D tensor store: shared (128B-swizzled box) -> global, bulk-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(map), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 :: "l"(map), "r"(smem_src), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t bf2_bits(float lo_elem, float hi_elem) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo_elem, hi_elem);   // .x (low 16 bits) = first element
    return *reinterpret_cast<uint32_t*>(&h);
}
// x[0..7] -> 8 bf16 hi (16 B) + 8 bf16 lo (16 B), x = hi + lo + O(2^-17 |x|)
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        h[j] = bf2_bits(x[2 * j], x[2 * j + 1]);
        const float h0 = __uint_as_float(h[j] << 16), h1 = __uint_as_float(h[j] & 0xffff0000u);
        l[j] = bf2_bits(x[2 * j] - h0, x[2 * j + 1] - h1);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// byte offset of 16-byte chunk c of row r in a [rows][128 B] tile, 128B swizzle (Swizzle<3,4,3>)
__device__ __forceinline__ uint32_t sw128(int r, int c) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((c ^ r) & 7) << 4));
}

// ------------------------------------------------------------------------------------------
// pass 1
// ------------------------------------------------------------------------------------------
constexpr int kR1 = 16;                               // nodes per stage = one UMMA K step
constexpr int kHeadTile1 = kR1 * 128;                 // 2048 B: [16 nodes][64 bf16] of one head
constexpr int kOp1 = kH * kHeadTile1;                 // 8192 B per operand (Khi / Klo / Vhi / Vlo)
constexpr int kStage1 = 4 * kOp1;                     // 32 KB
constexpr int kNS1 = 4;
constexpr int kSmem1 = kNS1 * kStage1 + 1024;

template <bool RING, bool EVICT_FIRST>
__global__ void __launch_bounds__(kThreadsTC, 1)
reduce_tc_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int64_t N,
                 int rows_per_cta, float* __restrict__ ws, int64_t ws_len, int pf_dist, int q_keep) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* stages = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[kNS1], empty[kNS1], done;
    __shared__ uint32_t tmem_slot;
    __shared__ float part[16];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = min(N, r0 + (int64_t)rows_per_cta);
    const int iters = r1 > r0 ? (int)((r1 - r0 + kR1 - 1) / kR1) : 0;

    if (tid == 0) {
        for (int s = 0; s < kNS1; ++s) { mbar_init(&full[s], 8); mbar_init(&empty[s], 1); }
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) tmem_alloc(&tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    float zacc[8], uacc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { zacc[i] = 0.f; uacc[i] = 0.f; }
    float ss = 0.f;     // producers: sum k^2 ; Q warps: sum q^2

    if (warp < 8) {
        if (!RING) {
            // variant A: whole-iteration double buffer (loads of it+1 issued, then it converted)
            float kc[2][8], vc[2][8], kn[2][8], vn[2][8];
            auto load = [&](int it, float (&kk)[2][8], float (&vv)[2][8]) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int64_t row = r0 + (int64_t)it * kR1 + warp + 8 * j;
                    if (row < r1) {
                        if (EVICT_FIRST) { ldg256_stream(k + row * kRowF + lane * 8, kk[j]); ldg256_stream(v + row * kRowF + lane * 8, vv[j]); }
                        else { ldg256_keep(k + row * kRowF + lane * 8, kk[j]); ldg256_keep(v + row * kRowF + lane * 8, vv[j]); }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { kk[j][i] = 0.f; vv[j][i] = 0.f; }
                    }
                }
            };
            if (iters > 0) load(0, kc, vc);
            const uint32_t stage_base = smem_u32(stages);
            for (int it = 0; it < iters; ++it) {
                if (it + 1 < iters) load(it + 1, kn, vn);
                const int s = it % kNS1;
                if (it >= kNS1) mbar_wait(&empty[s], ((it / kNS1) - 1) & 1);
                const uint32_t sb = stage_base + s * kStage1;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint32_t off = (uint32_t)((lane >> 3) * kHeadTile1 + j * 1024 + warp * 128 + ((((lane & 7) ^ warp) & 7) << 4));
                    uint4 hi, lo;
                    split8(kc[j], hi, lo);
                    sts128(sb + 0 * kOp1 + off, hi);
                    sts128(sb + 1 * kOp1 + off, lo);
                    split8(vc[j], hi, lo);
                    sts128(sb + 2 * kOp1 + off, hi);
                    sts128(sb + 3 * kOp1 + off, lo);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        zacc[i] += kc[j][i];
                        uacc[i] += vc[j][i];
                        ss = fmaf(kc[j][i], kc[j][i], ss);
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
                if (it + 1 < iters) {
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int i = 0; i < 8; ++i) { kc[j][i] = kn[j][i]; vc[j][i] = vn[j][i]; }
                }
            }
        } else {
            // ===== K/V producers: warp w owns nodes w and w+8 of every 16-node stage, lane l owns columns 8l..8l+7.
            // Register ring of two iterations x 4 chunks (K/V x 2 nodes): as soon as a chunk is converted its
            // registers are refilled with the load of iteration it+2, so ~8 x 32 B per thread stay in flight.
            float buf[2][4][8];
            auto issue = [&](int it, int c, float (&dst)[8]) {
                if (it >= iters) return;
                const int64_t row = r0 + (int64_t)it * kR1 + warp + 8 * (c >> 1);
                if (row < r1) {
                    if (EVICT_FIRST) ldg256_stream(((c & 1) ? v : k) + row * kRowF + lane * 8, dst);
                    else ldg256_keep(((c & 1) ? v : k) + row * kRowF + lane * 8, dst);
                } else {
    #pragma unroll
                    for (int i = 0; i < 8; ++i) dst[i] = 0.f;
                }
            };
    #pragma unroll
            for (int c = 0; c < 4; ++c) issue(0, c, buf[0][c]);
    #pragma unroll
            for (int c = 0; c < 4; ++c) issue(1, c, buf[1][c]);
            const uint32_t stage_base = smem_u32(stages);
            for (int it0 = 0; it0 < iters; it0 += 2) {
    #pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int it = it0 + half;
                    if (it < iters) {
                        const int s = it % kNS1;
                        if (it >= kNS1) mbar_wait(&empty[s], ((it / kNS1) - 1) & 1);
                        const uint32_t sb = stage_base + s * kStage1;
    #pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            // chunk c: node r = warp + 8*(c>>1) of the stage ((r>>3) = c>>1, (r&7) = warp); K for even c, V for odd c
                            const uint32_t off = (uint32_t)((lane >> 3) * kHeadTile1 + (c >> 1) * 1024 + warp * 128 + ((((lane & 7) ^ warp) & 7) << 4));
                            uint4 hi, lo;
                            split8(buf[half][c], hi, lo);
                            sts128(sb + ((c & 1) ? 2 : 0) * kOp1 + off, hi);
                            sts128(sb + ((c & 1) ? 3 : 1) * kOp1 + off, lo);
    #pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                if (c & 1) {
                                    uacc[i] += buf[half][c][i];
                                } else {
                                    zacc[i] += buf[half][c][i];
                                    ss = fmaf(buf[half][c][i], buf[half][c][i], ss);
                                }
                            }
                            issue(it + 2, c, buf[half][c]);
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full[s]);
                    }
                }
            }
        }
    } else if (warp < 12) {
        // ===== Q stream: sum of squares only (the Frobenius norm of difformer.py:20)
        const int64_t n8 = (r1 > r0 ? (r1 - r0) : 0) * (kRowF / 8);
        const float* base = q + r0 * kRowF;
        const int t = tid - 256;
        const uint64_t qpol = policy_evict_last();
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int64_t i = t;
        for (; i + 7 * 128 < n8; i += 8 * 128) {
            float x[8][8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (q_keep) ldg256_policy(base + (i + u * 128) * 8, x[u], qpol);
                else ldg256_keep(base + (i + u * 128) * 8, x[u]);
            }
#pragma unroll
            for (int u = 0; u < 8; u += 4)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    a0 = fmaf(x[u][e], x[u][e], a0); a1 = fmaf(x[u + 1][e], x[u + 1][e], a1);
                    a2 = fmaf(x[u + 2][e], x[u + 2][e], a2); a3 = fmaf(x[u + 3][e], x[u + 3][e], a3);
                }
        }
        for (; i < n8; i += 128) {
            float x0[8];
            ldg256_keep(base + i * 8, x0);
#pragma unroll
            for (int e = 0; e < 8; ++e) a0 = fmaf(x0[e], x0[e], a0);
        }
        ss = (a0 + a1) + (a2 + a3);
    } else if (lane == 0) {
        // ===== MMA issuer
        const uint32_t idesc = make_idesc(128, 128, 1, 1);
#if DIF_MN_LBO_IS_MNSTRIDE
        const uint32_t lbo = kHeadTile1, sbo = 1024;
#else
        const uint32_t lbo = 1024, sbo = kHeadTile1;
#endif
        const uint32_t stage_base = smem_u32(stages);
        for (int it = 0; it < iters; ++it) {
            const int s = it % kNS1;
            if (pf_dist > 0) {
                // L2 prefetch of the rows `pf_dist` stages ahead (K, V, Q: 16 KB each, contiguous): the producers'
                // register loads then hit L2, which roughly triples the bandwidth one load slot sustains
                const int64_t prow = r0 + (int64_t)(it + pf_dist) * kR1;
                if (prow < r1) {
                    const uint32_t bytes = (uint32_t)(min((int64_t)kR1, r1 - prow) * kRowF * 4);
                    if (q_keep) {       // K,V are dead after this pass, Q is re-read by pass 2: tell the L2
                        prefetch_l2_hint(k + prow * kRowF, bytes, policy_evict_first_());
                        prefetch_l2_hint(v + prow * kRowF, bytes, policy_evict_first_());
                        prefetch_l2_hint(q + prow * kRowF, bytes, policy_evict_last());
                    } else {
                        prefetch_l2(k + prow * kRowF, bytes);
                        prefetch_l2(v + prow * kRowF, bytes);
                        prefetch_l2(q + prow * kRowF, bytes);
                    }
                }
            }
            mbar_wait(&full[s], (it / kNS1) & 1);
            tc_fence_after();
            const uint32_t sb = stage_base + s * kStage1;
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const uint32_t ho = p * 2 * kHeadTile1;          // heads 2p, 2p+1
                const uint64_t khi = make_desc(sb + 0 * kOp1 + ho, lbo, sbo), klo = make_desc(sb + 1 * kOp1 + ho, lbo, sbo);
                const uint64_t vhi = make_desc(sb + 2 * kOp1 + ho, lbo, sbo), vlo = make_desc(sb + 3 * kOp1 + ho, lbo, sbo);
                umma(tmem + p * 128, khi, vhi, idesc, it > 0 ? 1u : 0u);
                umma(tmem + p * 128, khi, vlo, idesc, 1u);
                umma(tmem + p * 128, klo, vhi, idesc, 1u);
            }
            umma_commit(&empty[s]);
        }
        if (iters > 0) umma_commit(&done); else mbar_arrive(&done);
    }

    // ===== epilogue: per-CTA record [S | z | u | sq slots | sk slots]
    __syncwarp();
    mbar_wait(&done, 0);
    tc_fence_after();
    ss = warp_sum(ss);
    if (lane == 0) part[warp] = ss;
    float* red = reinterpret_cast<float*>(stages);      // all MMAs have completed: stage memory is free
    if (warp < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            red[warp * kRowF + lane * 8 + i] = zacc[i];
            red[8 * kRowF + warp * kRowF + lane * 8 + i] = uacc[i];
        }
    }
    __syncthreads();
    float* rec = ws + (int64_t)blockIdx.x * ws_len;
    const int64_t offZ = (int64_t)kH * kDim * kDim, offU = offZ + kH * kDim, offSq = offU + kH * kDim;
    if (tid < kRowF) {
        float z = 0.f, u = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { z += red[w * kRowF + tid]; u += red[8 * kRowF + w * kRowF + tid]; }
        rec[offZ + tid] = z;
        rec[offU + tid] = u;
    }
    if (tid == 0) {
        float sk = 0.f, sq = 0.f;
        for (int w = 0; w < 8; ++w) sk += part[w];
        for (int w = 8; w < 12; ++w) sq += part[w];
        for (int h = 0; h < kH; ++h) { rec[offSq + h] = h == 0 ? sq : 0.f; rec[offSq + kH + h] = h == 0 ? sk : 0.f; }
    }
    if (warp < 4) {
        // D_p rows 0-63 x cols 0-63 = S_{2p}; rows 64-127 x cols 64-127 = S_{2p+1}; warp w reads lanes 32w..32w+31
        const int hp = warp >> 1, m = (warp * 32 + lane) & 63;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            float* dst = rec + ((int64_t)(2 * p + hp) * kDim + m) * kDim;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t r[32];
                if (iters > 0) {
                    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + p * 128 + hp * 64 + c0, r);
                    tmem_ld_wait32(r);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(dst + c0 + j) =
                        make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------
// pass 1, TMA-staged variant: the K/V/Q rows of a stage (16 nodes x 3 x 1 KB, contiguous in HBM) are
// fetched with cp.async.bulk into an fp32 staging ring (mbarrier complete_tx), so the loads in
// flight are bounded by shared memory (3 x 48 KB per SM), not by registers / L1 miss tracking.
// Converter warps read the staging rows (conflict-free LDS.128), split to bf16 hi/lo and write the
// swizzled UMMA operand ring.
// ------------------------------------------------------------------------------------------
constexpr int kBOpBytes = 80 * 128;                   // one (head, hi|lo) B-operand tile of pass 2: 80 rows x 128 B
constexpr int kShardMaxRanks = 16;
struct ShardArgs {            // multi-GPU: peer-mapped exchange buffers [2 data slots | flags], see csrc/comm.cu
    float* bufs[kShardMaxRanks];
    int rank, world;
    unsigned long long seq;
    int64_t slot_floats;
};
constexpr int kNSG = 3;                               // staging stages
constexpr int kStgT = kR1 * 1024;                     // 16 KB: 16 rows of one tensor
constexpr int kStg = 3 * kStgT;                       // K | V | Q
constexpr int kNO = 2;                                // operand stages
constexpr int kSmem1T = kNSG * kStg + kNO * kStage1 + 1024;
constexpr int kThreadsT = 10 * 32;                    // warps 0-7 converters, 8 TMA issuer, 9 MMA issuer

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_1d_hint(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" :: "r"(addr), "r"(a), "r"(b) : "memory");
}
// 4 floats -> 4 bf16 hi (8 B) + 4 bf16 lo (8 B)
__device__ __forceinline__ void split4(const float4& x, uint32_t (&hi)[2], uint32_t (&lo)[2]) {
    hi[0] = bf2_bits(x.x, x.y);
    hi[1] = bf2_bits(x.z, x.w);
    lo[0] = bf2_bits(x.x - __uint_as_float(hi[0] << 16), x.y - __uint_as_float(hi[0] & 0xffff0000u));
    lo[1] = bf2_bits(x.z - __uint_as_float(hi[1] << 16), x.w - __uint_as_float(hi[1] & 0xffff0000u));
}

__global__ void __launch_bounds__(kThreadsT, 1)
reduce_tma_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int64_t N,
                  int rows_per_cta, float* __restrict__ ws, int64_t ws_len, unsigned long long* __restrict__ flags,
                  unsigned long long epoch, float* __restrict__ partials, uint8_t* __restrict__ prepared,
                  int pf_tiles, int pf_grid, int l2_hints, const ShardArgs sh, uint64_t* __restrict__ dbg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* stg = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* ops = stg + kNSG * kStg;
    __shared__ uint64_t sfull[kNSG], sempty[kNSG], ofull[kNO], oempty[kNO], done, tail_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float part[16];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = min(N, r0 + (int64_t)rows_per_cta);
    const int iters = r1 > r0 ? (int)((r1 - r0 + kR1 - 1) / kR1) : 0;
    DIF_STAMP(dbg, 0);

    if (tid == 0) {
        for (int s = 0; s < kNSG; ++s) { mbar_init(&sfull[s], 1); mbar_init(&sempty[s], 8); }
        for (int s = 0; s < kNO; ++s) { mbar_init(&ofull[s], 8); mbar_init(&oempty[s], 1); }
        mbar_init(&done, 1);
        mbar_init(&tail_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) tmem_alloc(&tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    DIF_STAMP(dbg, 1);

    float zacc[8], uacc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { zacc[i] = 0.f; uacc[i] = 0.f; }
    float ssk = 0.f, ssq = 0.f;

    if (warp < 8) {
        // ===== converters: warp w owns nodes w and w+8 of every stage; lane l owns columns 4l..4l+3 and 128+4l..131+4l
        const uint32_t stg_base = smem_u32(stg), ops_base = smem_u32(ops);
        for (int it = 0; it < iters; ++it) {
            const int s = it % kNSG, o = it % kNO;
            const int nrows = (int)min((int64_t)kR1, r1 - (r0 + (int64_t)it * kR1));
            mbar_wait(&sfull[s], (it / kNSG) & 1);
            if (it == 0) DIF_STAMP(dbg, 2);
            float4 x[2][3][2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int node = warp + 8 * j;
#pragma unroll
                for (int t = 0; t < 3; ++t)
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        x[j][t][g] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (node < nrows) x[j][t][g] = lds128(stg_base + s * kStg + t * kStgT + node * 1024 + g * 512 + lane * 16);
                    }
            }
            if (it >= kNO) mbar_wait(&oempty[o], ((it / kNO) - 1) & 1);
            const uint32_t ob = ops_base + o * kStage1;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    // column 128g + 4 lane: head = 2g + (lane>>4), m = 4 (lane & 15): chunk16 = (lane & 15) >> 1, half = lane & 1
                    const uint32_t off = (uint32_t)((2 * g + (lane >> 4)) * kHeadTile1 + j * 1024 + warp * 128 +
                                                    ((((lane & 15) >> 1) ^ warp) & 7) * 16 + (lane & 1) * 8);
                    uint32_t hi[2], lo[2];
                    split4(x[j][0][g], hi, lo);
                    sts64(ob + 0 * kOp1 + off, hi[0], hi[1]);
                    sts64(ob + 1 * kOp1 + off, lo[0], lo[1]);
                    split4(x[j][1][g], hi, lo);
                    sts64(ob + 2 * kOp1 + off, hi[0], hi[1]);
                    sts64(ob + 3 * kOp1 + off, lo[0], lo[1]);
                    const float4 kk = x[j][0][g], vv = x[j][1][g], qq = x[j][2][g];
                    zacc[4 * g + 0] += kk.x; zacc[4 * g + 1] += kk.y; zacc[4 * g + 2] += kk.z; zacc[4 * g + 3] += kk.w;
                    uacc[4 * g + 0] += vv.x; uacc[4 * g + 1] += vv.y; uacc[4 * g + 2] += vv.z; uacc[4 * g + 3] += vv.w;
                    ssk = fmaf(kk.x, kk.x, ssk); ssk = fmaf(kk.y, kk.y, ssk); ssk = fmaf(kk.z, kk.z, ssk); ssk = fmaf(kk.w, kk.w, ssk);
                    ssq = fmaf(qq.x, qq.x, ssq); ssq = fmaf(qq.y, qq.y, ssq); ssq = fmaf(qq.z, qq.z, ssq); ssq = fmaf(qq.w, qq.w, ssq);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&ofull[o]); mbar_arrive(&sempty[s]); }
        }
    } else if (warp == 8) {
        if (lane == 0) {
            // ===== TMA issuer: three 16 KB bulk copies per stage
            const uint32_t stg_base = smem_u32(stg);
            for (int it = 0; it < iters; ++it) {
                const int s = it % kNSG;
                if (it >= kNSG) mbar_wait(&sempty[s], ((it / kNSG) - 1) & 1);
                const int64_t row = r0 + (int64_t)it * kR1;
                const uint32_t bytes = (uint32_t)(min((int64_t)kR1, r1 - row) * kRowF * 4);
                mbar_expect_tx(&sfull[s], 3 * bytes);
                if (l2_hints) {     // K, V are dead after this pass; Q is read again by pass 2
                    tma_load_1d_hint(stg_base + s * kStg + 0 * kStgT, k + row * kRowF, bytes, &sfull[s], policy_evict_first_());
                    tma_load_1d_hint(stg_base + s * kStg + 1 * kStgT, v + row * kRowF, bytes, &sfull[s], policy_evict_first_());
                    tma_load_1d_hint(stg_base + s * kStg + 2 * kStgT, q + row * kRowF, bytes, &sfull[s], policy_evict_last());
                } else {
                    tma_load_1d(stg_base + s * kStg + 0 * kStgT, k + row * kRowF, bytes, &sfull[s]);
                    tma_load_1d(stg_base + s * kStg + 1 * kStgT, v + row * kRowF, bytes, &sfull[s]);
                    tma_load_1d(stg_base + s * kStg + 2 * kStgT, q + row * kRowF, bytes, &sfull[s]);
                }
            }
        }
    } else if (lane == 0) {
        // ===== MMA issuer
        const uint32_t idesc = make_idesc(128, 128, 1, 1);
        const uint32_t lbo = kHeadTile1, sbo = 1024;
        const uint32_t ops_base = smem_u32(ops);
        for (int it = 0; it < iters; ++it) {
            const int o = it % kNO;
            mbar_wait(&ofull[o], (it / kNO) & 1);
            tc_fence_after();
            const uint32_t sb = ops_base + o * kStage1;
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const uint32_t ho = p * 2 * kHeadTile1;
                const uint64_t khi = make_desc(sb + 0 * kOp1 + ho, lbo, sbo), klo = make_desc(sb + 1 * kOp1 + ho, lbo, sbo);
                const uint64_t vhi = make_desc(sb + 2 * kOp1 + ho, lbo, sbo), vlo = make_desc(sb + 3 * kOp1 + ho, lbo, sbo);
                umma(tmem + p * 128, khi, vhi, idesc, it > 0 ? 1u : 0u);
                umma(tmem + p * 128, khi, vlo, idesc, 1u);
                umma(tmem + p * 128, klo, vhi, idesc, 1u);
            }
            umma_commit(&oempty[o]);
        }
        if (iters > 0) umma_commit(&done); else mbar_arrive(&done);
    }

    // ===== tail: per-CTA record in the partials layout [S | z | u | sq | sk], then the cross-CTA sum fused in:
    // every CTA publishes its record (flag = epoch), waits for all flags (the grid is launched
    // cooperatively, all CTAs are resident), TMA-loads "its" column slice of all records into shared
    // memory and sums it in fixed order (fp64) -> partials (deterministic, no float atomics).
    // The S / z entries are also emitted as the bf16 hi/lo, 128B-swizzled B operand image of pass 2.
    __syncwarp();
    if (warp == 0) DIF_STAMP(dbg, 3);
    mbar_wait(&done, 0);
    tc_fence_after();
    DIF_STAMP(dbg, 4);
    ssk = warp_sum(ssk);
    ssq = warp_sum(ssq);
    if (lane == 0 && warp < 8) { part[warp] = ssk; part[8 + warp] = ssq; }
    float* red = reinterpret_cast<float*>(ops);         // all MMAs have completed: operand memory is free
    if (warp < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int col = (i < 4) ? 4 * lane + i : 128 + 4 * lane + (i - 4);
            red[warp * kRowF + col] = zacc[i];
            red[8 * kRowF + warp * kRowF + col] = uacc[i];
        }
    }
    __syncthreads();
    float* rec = ws + (int64_t)blockIdx.x * ws_len;
    constexpr int offZ = kH * kDim * kDim, offU = offZ + kH * kDim, offSq = offU + kH * kDim, kP = offSq + 2;
    if (tid < kRowF) {
        float z = 0.f, u = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { z += red[w * kRowF + tid]; u += red[8 * kRowF + w * kRowF + tid]; }
        rec[offZ + tid] = z;
        rec[offU + tid] = u;
    }
    if (tid == 0) {
        float sk = 0.f, sq = 0.f;
        for (int w = 0; w < 8; ++w) { sk += part[w]; sq += part[8 + w]; }
        rec[offSq] = sq;
        rec[offSq + 1] = sk;
        for (int64_t i = kP; i < ws_len; ++i) rec[i] = 0.f;
    }
    if (warp < 8) {
        // D_p rows 0-63 x cols 0-63 = S_{2p}, rows 64-127 x cols 64-127 = S_{2p+1}.  Warp w reads TMEM lanes
        // 32(w%4)..+31 (its quadrant); warps 0-3 take head pair p = 0, warps 4-7 p = 1.  256-bit stores.
        const int wq = warp & 3, p = warp >> 2;
        const int hp = wq >> 1, m = (wq * 32 + lane) & 63;
        float* dst = rec + ((int64_t)(2 * p + hp) * kDim + m) * kDim;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t r[32];
            if (iters > 0) {
                tmem_ld32(tmem + ((uint32_t)(wq * 32) << 16) + p * 128 + hp * 64 + c0, r);
                tmem_ld_wait32(r);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 8)
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                             :: "l"(dst + c0 + j), "r"(r[j]), "r"(r[j + 1]), "r"(r[j + 2]), "r"(r[j + 3]), "r"(r[j + 4]), "r"(r[j + 5]),
                                "r"(r[j + 6]), "r"(r[j + 7]) : "memory");
        }
    }
    // ---- publish the record
    __threadfence();
    __syncthreads();
    DIF_STAMP(dbg, 5);
    const int grid = gridDim.x;
    if (tid == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(flags + blockIdx.x), "l"(epoch) : "memory");
    if (tid == 64 && pf_tiles > 0 && (int)blockIdx.x < pf_grid) {
        // HBM is idle while the grid exchanges its records: pull the first Q tiles pass 2 will read (tiles
        // b, b + G, ... of "its" CTA b) into L2 now, so pass 2 starts from L2 hits instead of a cold ramp
        const int64_t ntiles = (N + 127) / 128;
        for (int i = 0; i < pf_tiles; ++i) {
            const int64_t tile = blockIdx.x + (int64_t)i * pf_grid;
            if (tile >= ntiles) break;
            const int64_t prow = tile * 128;
            const int64_t nrows = min((int64_t)128, N - prow);
            for (int64_t r = 0; r < nrows; r += 16)
                prefetch_l2(q + (prow + r) * kRowF, (uint32_t)(min((int64_t)16, nrows - r) * kRowF * 4));
        }
    }
    // ---- column slices of the record: kSlices fixed slices of `chunk` floats (multiples of 16 B), slice sl is
    //      owned by CTA sl % grid -- the partition does not depend on this rank's grid, so slices line up across ranks
    constexpr int kSlices = 148;
    const int chunk = (int)((((ws_len + kSlices - 1) / kSlices) + 3) & ~(int64_t)3);
    float* sbuf = reinterpret_cast<float*>(stg);        // [grid][chunk] fp32, the staging ring is idle now
    const bool sharded = sh.world > 1;
    const int xslot = (int)(sh.seq & 1);
    bool waited = false;
    uint32_t tail_phase = 0;
    for (int sl = blockIdx.x; sl < kSlices; sl += grid) {
        const int64_t j0 = (int64_t)sl * chunk;
        const int slice = (int)max((int64_t)0, min(ws_len, j0 + chunk) - j0);
        if (slice <= 0) break;
        if (!waited) {
            for (int r = tid; r < grid; r += kThreadsT) {
                unsigned long long f;
                do {
                    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(f) : "l"(flags + r) : "memory");
                    if (f != epoch) __nanosleep(64);
                } while (f != epoch);
            }
            asm volatile("fence.proxy.async;" ::: "memory");   // acquired (generic proxy) before the bulk (async proxy) reads
            waited = true;
        }
        __syncthreads();                                  // also: previous slice's readers of sbuf are done
        if (tid == 0) mbar_expect_tx(&tail_bar, (uint32_t)grid * (uint32_t)slice * 4u);
        __syncthreads();
        for (int r = tid; r < grid; r += kThreadsT)
            tma_load_1d(smem_u32(sbuf) + (uint32_t)r * chunk * 4, ws + (int64_t)r * ws_len + j0, (uint32_t)slice * 4u, &tail_bar);
        mbar_wait(&tail_bar, tail_phase);
        tail_phase ^= 1;
        // local sum of this slice over the grid's records
        float local = 0.f;
        const int t = tid;                                // chunk <= kThreadsT is asserted on the host
        const int64_t j = j0 + t;
        const bool live = t < slice && j < kP;
        if (live) {
            // four independent chains (records r = 4i + k), combined in a fixed order: deterministic, 4x shorter latency
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int r = 0;
            for (; r + 3 < grid; r += 4) {
                a0 += (double)sbuf[(r + 0) * chunk + t];
                a1 += (double)sbuf[(r + 1) * chunk + t];
                a2 += (double)sbuf[(r + 2) * chunk + t];
                a3 += (double)sbuf[(r + 3) * chunk + t];
            }
            for (; r < grid; ++r) a0 += (double)sbuf[r * chunk + t];
            local = (float)((a0 + a1) + (a2 + a3));
        }
        float sum = local;
        if (sharded) {
            // ---- cross-GPU: the same kernel finishes the all-reduce over NVLink, slice by slice.  Publish the local
            // slice in this rank's peer-mapped buffer, raise flag (slot, rank, slice) in every peer, wait for the
            // peers' flags, add the ranks' slices in rank order (bit-identical on all ranks).
            float* mine = sh.bufs[sh.rank] + (int64_t)xslot * sh.slot_floats;
            if (live) mine[j] = local;
            __threadfence_system();
            __syncthreads();
            if (tid < sh.world) {
                unsigned long long* f = reinterpret_cast<unsigned long long*>(sh.bufs[tid] + 2 * sh.slot_floats) +
                                        ((size_t)(xslot * kShardMaxRanks + sh.rank) * 256 + sl);
                asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(f), "l"(sh.seq) : "memory");
                const unsigned long long* w = reinterpret_cast<const unsigned long long*>(sh.bufs[sh.rank] + 2 * sh.slot_floats) +
                                              ((size_t)(xslot * kShardMaxRanks + tid) * 256 + sl);
                unsigned long long got;
                do {
                    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(w) : "memory");
                } while (got != sh.seq);
            }
            __syncthreads();
            if (live) {
                sum = 0.f;
                for (int r = 0; r < sh.world; ++r) {
                    float x;
                    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(x) : "l"(sh.bufs[r] + (int64_t)xslot * sh.slot_floats + j) : "memory");
                    sum += x;
                }
            }
        }
        if (live) {
            partials[j] = sum;
            if (prepared != nullptr && j < offU) {
                // B operand image of pass 2 (un-scaled; pass 2 applies c = 1/(|Q||K|) in its epilogue):
                // S[h][m][d] -> row n = d, k = m of head h ; z[h][m] -> row 64
                int h, n, m;
                if (j < offZ) { h = (int)(j >> 12); m = (int)(j >> 6) & 63; n = (int)j & 63; }
                else { h = (int)(j - offZ) >> 6; m = (int)(j - offZ) & 63; n = kDim; }
                const __nv_bfloat16 hi = __float2bfloat16_rn(sum);
                const __nv_bfloat16 lo = __float2bfloat16_rn(sum - __bfloat162float(hi));
                uint8_t* img = prepared + (size_t)h * 2 * kBOpBytes + sw128(n, m >> 3) + (m & 7) * 2;
                *reinterpret_cast<__nv_bfloat16*>(img) = hi;
                *reinterpret_cast<__nv_bfloat16*>(img + kBOpBytes) = lo;
            }
        }
    }
    if (prepared != nullptr && blockIdx.x == grid - 1) {
        // zero rows 65..79 of every (head, hi/lo) tile: 15 rows x 128 B, contiguous after row 64 inside the last 8-row group...
        // rows 64..71 live in group 8 (bytes 8192..9215), rows 72..79 in group 9: zero everything except row 64
        for (int i = tid; i < kH * 2 * 15 * 8; i += kThreadsT) {
            const int c = i & 7, rr = (i >> 3) % 15 + 65, t = i / (8 * 15);
            *reinterpret_cast<uint4*>(prepared + (size_t)t * kBOpBytes + sw128(rr, c)) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    tc_fence_before();
    __syncthreads();
    DIF_STAMP(dbg, 6);
    if (warp == 9) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------
// pass 2
// ------------------------------------------------------------------------------------------
constexpr int kTile2 = 128;                           // rows per tile = UMMA M
constexpr int kQOp = kTile2 * 128;                    // 16 KB: [128 rows][64 bf16] of one head
constexpr int kStage2 = 2 * kQOp;                     // Qhi | Qlo
constexpr int kNS2 = 3;
constexpr int kBN = 80;                               // UMMA N: 64 columns of S + z column + padding
constexpr int kBOp = kBN * 128;                       // 10 KB
constexpr int kBBytes = kH * 2 * kBOp;                // 80 KB: per head hi | lo
constexpr int kNAcc = 4, kAccCols = 128;              // TMEM accumulator ring (4 x 128 columns)
constexpr int kOutBox = 32 * 128;                     // TMA store box: 32 rows x 32 floats, 128B swizzle
constexpr int kOutStage = 4 * 2 * kOutBox;            // per epilogue warp: two boxes (column halves of a head)
constexpr int kSmem2 = kBBytes + kNS2 * kStage2 + kOutStage + kH * kDim * 4 + 1024;

struct ApplyTcArgs {
    const float* q;
    const float* partials;
    float n_total;
    int64_t N;
    float* out;
    int tiles_per_cta;      // > 0: CTA b owns tiles [b*tpc, (b+1)*tpc) -- the row range it reduced in pass 1 -- and walks
                            //      them backwards, so the Q rows pass 1 touched last are re-read first (L2 hits)
    int pf_tiles;           // L2 prefetch distance in tiles (0 = off)
    uint64_t* dbg;          // optional timeline buffer
    const uint8_t* prepared; // optional B operand image written by the fused pass-1 tail (un-scaled S|z, bf16 hi/lo, swizzled)
    int store_hint;         // 1: TMA stores carry an L2 evict_first policy (output is not re-read; keeps Q resident)
    dif_epilogue_t ep;
};

template <int MODE, bool RING>
__global__ void __launch_bounds__(kThreadsTC, 1) apply_tc_kernel(const __grid_constant__ ApplyTcArgs p, const __grid_constant__ CUtensorMap out_map) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* Bop = base;                               // [h][hi|lo][80 rows][128 B]
    uint8_t* stages = base + kBBytes;
    uint8_t* ostage = stages + kNS2 * kStage2;                       // [4 warps][2 boxes][32 rows][128 B], 1024-aligned
    float* us = reinterpret_cast<float*>(ostage + kOutStage);        // [H][64]
    __shared__ uint64_t full[kNS2], empty[kNS2], tfull[kNAcc], tempty[kNAcc], bbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t ntiles = (p.N + kTile2 - 1) / kTile2;
    const bool contiguous = p.tiles_per_cta > 0;
    const int64_t first_tile = contiguous ? (int64_t)blockIdx.x * p.tiles_per_cta : blockIdx.x;
    int my_tiles;
    if (contiguous) {
        const int64_t rem = ntiles - first_tile;
        my_tiles = rem <= 0 ? 0 : (rem < p.tiles_per_cta ? (int)rem : p.tiles_per_cta);
    } else {
        my_tiles = blockIdx.x < ntiles ? (int)((ntiles - 1 - blockIdx.x) / gridDim.x + 1) : 0;
    }
    const int nsc = my_tiles * kH;                     // (tile, head) stages of this CTA
    auto tile_of = [&](int sc) -> int64_t {
        const int i = sc >> 2;
        return contiguous ? first_tile + (my_tiles - 1 - i) : first_tile + (int64_t)i * gridDim.x;
    };

    DIF_STAMP(p.dbg, 0);
    if (tid == 32 && my_tiles > 0) {
        // the B-operand prologue below takes a few us: have the first tiles of Q on their way to L2 meanwhile
        for (int i = 0; i < 2 && i < my_tiles; ++i) {
            const int64_t prow = tile_of(4 * i) * kTile2;
            const int64_t nrows = min((int64_t)kTile2, p.N - prow);
            for (int64_t r = 0; r < nrows; r += 16)
                prefetch_l2(p.q + (prow + r) * kRowF, (uint32_t)(min((int64_t)16, nrows - r) * kRowF * 4));
        }
    }
    if (tid == 0) {
        for (int s = 0; s < kNS2; ++s) { mbar_init(&full[s], 8); mbar_init(&empty[s], 1); }
        for (int s = 0; s < kNAcc; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
        mbar_init(&bbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (p.prepared != nullptr) {
            // the B operands were prepared by pass 1: one 80 KB TMA fetch instead of a transposing prologue
            mbar_expect_tx(&bbar, (uint32_t)kBBytes);
            for (int i = 0; i < kH * 2; ++i)
                tma_load_1d(smem_u32(Bop) + i * kBOp, p.prepared + (size_t)i * kBOp, (uint32_t)kBOp, &bbar);
        }
    }
    if (warp == 12) tmem_alloc(&tmem_slot, 512);

    // ---- B operands: row n < 64: c*S[h][:, n] ; row 64: c*z[h] ; rows 65..79: 0   (K-major SW128, hi/lo split)
    const int64_t offZ = (int64_t)kH * kDim * kDim, offU = offZ + kH * kDim, offSq = offU + kH * kDim;
    const float c = 1.f / (sqrtf(p.partials[offSq]) * sqrtf(p.partials[offSq + 1]));
    const float cscale = p.prepared != nullptr ? c : 1.f;   // prepared operands are un-scaled: the epilogue applies c
    if (p.prepared == nullptr) {
        // all loads of a thread's (up to 7) tasks are issued before the first use: two L2 round trips, not 50
        constexpr int kTasks = kH * 8 * kBN, kPer = (kTasks + kThreadsTC - 1) / kThreadsTC;
        float x[kPer][8];
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int task = tid + u * kThreadsTC;
            const int n = task % kBN, hc = task / kBN, ch = hc & 7, h = hc >> 3;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = ch * 8 + i;
                x[u][i] = 0.f;
                if (task < kTasks) {
                    if (n < kDim) x[u][i] = __ldg(p.partials + ((int64_t)h * kDim + m) * kDim + n);
                    else if (n == kDim) x[u][i] = __ldg(p.partials + offZ + h * kDim + m);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int task = tid + u * kThreadsTC;
            if (task < kTasks) {
                const int n = task % kBN, hc = task / kBN, ch = hc & 7, h = hc >> 3;
#pragma unroll
                for (int i = 0; i < 8; ++i) x[u][i] *= c;
                uint4 hi, lo;
                split8(x[u], hi, lo);
                const uint32_t off = (uint32_t)(h * 2 * kBOp) + sw128(n, ch);
                sts128(smem_u32(Bop) + off, hi);
                sts128(smem_u32(Bop) + kBOp + off, lo);
            }
        }
    }
    for (int i = tid; i < kH * kDim; i += kThreadsTC) us[i] = p.partials[offU + i];
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    DIF_STAMP(p.dbg, 1);

    if (warp < 8) {
        // ===== Q producers: stage = (tile, head): 128 rows x 256 B; task t -> row t>>3, chunk t&7.
        // Two-stage register ring, refilled chunk by chunk (see pass 1).
        float buf[2][4][8];
        auto issue = [&](int sc, int j, float (&dst)[8]) {
            if (sc >= nsc) return;
            const int64_t tile = tile_of(sc);
            const int t = tid + 256 * j;
            const int64_t row = tile * kTile2 + (t >> 3);
            if (row < p.N) {
                ldg256_stream(p.q + row * kRowF + (sc & 3) * kDim + (t & 7) * 8, dst);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) dst[i] = 0.f;
            }
        };
#pragma unroll
        for (int j = 0; j < 4; ++j) issue(0, j, buf[0][j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) issue(1, j, buf[1][j]);
        const uint32_t stage_base = smem_u32(stages);
        if (RING) {
            for (int sc0 = 0; sc0 < nsc; sc0 += 2) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int sc = sc0 + half;
                    if (sc < nsc) {
                        const int s = sc % kNS2;
                        if (sc >= kNS2) mbar_wait(&empty[s], ((sc / kNS2) - 1) & 1);
                        const uint32_t sb = stage_base + s * kStage2;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int t = tid + 256 * j;
                            uint4 hi, lo;
                            split8(buf[half][j], hi, lo);
                            const uint32_t off = sw128(t >> 3, t & 7);
                            sts128(sb + off, hi);
                            sts128(sb + kQOp + off, lo);
                            issue(sc + 2, j, buf[half][j]);
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&full[s]);
                    }
                }
            }
        } else {
            // whole-stage double buffer: buf[0] = current, buf[1] = next (loads of sc+1 issued, then sc converted)
            for (int sc = 0; sc < nsc; ++sc) {
                const int s = sc % kNS2;
                if (sc >= kNS2) mbar_wait(&empty[s], ((sc / kNS2) - 1) & 1);
                const uint32_t sb = stage_base + s * kStage2;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int t = tid + 256 * j;
                    uint4 hi, lo;
                    split8(buf[0][j], hi, lo);
                    const uint32_t off = sw128(t >> 3, t & 7);
                    sts128(sb + off, hi);
                    sts128(sb + kQOp + off, lo);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) buf[0][j][i] = buf[1][j][i];
                    issue(sc + 2, j, buf[1][j]);
                }
            }
        }
    } else if (warp < 12) {
        // ===== epilogue: thread = one row of the tile; accumulator lane = 32*(warp%4) + lane.
        // Results leave through TMA: each lane writes its row into a 128B-swizzled staging box
        // (conflict-free st.shared), one lane issues cp.async.bulk.tensor stores (rows beyond N are
        // clipped by the tensor map).  No scattered st.global on the LSU.
        const int ew = warp - 8;
        const uint32_t obox = smem_u32(ostage) + ew * 2 * kOutBox;
        float hs[MODE == 1 ? kDim : 1];
        for (int sc = 0; sc < nsc; ++sc) {
            const int64_t tile = tile_of(sc);
            const int h = sc & 3, slot = sc % kNAcc;
            mbar_wait(&tfull[slot], (sc / kNAcc) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(ew * 32) << 16) + slot * kAccCols;
            uint32_t qz_bits = tmem_ld1(taddr + kDim);            // column 64 = q^.z^
            tmem_ld_wait1(qz_bits);
            const float inv_den = 1.f / (fmaf(__uint_as_float(qz_bits), cscale, p.n_total));   // one division per (row, head)
            if (MODE == 1 && h == 0) {
#pragma unroll
                for (int i = 0; i < kDim; ++i) hs[i] = 0.f;
            }
            if (MODE == 0 || h == kH - 1) {       // staging is about to be rewritten: previous TMA reads must be done
                if (lane == 0) tma_wait_read0();
                __syncwarp();
            }
#pragma unroll
            for (int c0 = 0; c0 < kDim; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c0, r);
                tmem_ld_wait32(r);
                if (c0 == 32) {          // everything of this slot is in registers: hand the accumulator back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[slot]);
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 u4 = *reinterpret_cast<const float4*>(us + h * kDim + c0 + j);
                    float4 o;
                    o.x = fmaf(__uint_as_float(r[j]), cscale, u4.x) * inv_den;
                    o.y = fmaf(__uint_as_float(r[j + 1]), cscale, u4.y) * inv_den;
                    o.z = fmaf(__uint_as_float(r[j + 2]), cscale, u4.z) * inv_den;
                    o.w = fmaf(__uint_as_float(r[j + 3]), cscale, u4.w) * inv_den;
                    if (MODE == 0) {
                        sts128(obox + (c0 >> 5) * kOutBox + sw128(lane, j >> 2),
                               make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w)));
                    } else {
                        hs[c0 + j] += o.x; hs[c0 + j + 1] += o.y; hs[c0 + j + 2] += o.z; hs[c0 + j + 3] += o.w;
                    }
                }
            }
            if (MODE == 1 && h == kH - 1) {
                const int64_t row = tile * kTile2 + ew * 32 + lane;
                const bool ok = row < p.N;
#pragma unroll
                for (int j = 0; j < kDim; j += 4) {
                    float4 o = make_float4(hs[j] * p.ep.attn_scale, hs[j + 1] * p.ep.attn_scale, hs[j + 2] * p.ep.attn_scale,
                                           hs[j + 3] * p.ep.attn_scale);
                    for (int a = 0; a < p.ep.n_add; ++a) {
                        if (ok) {
                            const float4 x = ldg4(p.ep.add[a] + row * kDim + j);
                            const float s = p.ep.add_scale[a];
                            o.x = fmaf(s, x.x, o.x); o.y = fmaf(s, x.y, o.y); o.z = fmaf(s, x.z, o.z); o.w = fmaf(s, x.w, o.w);
                        }
                    }
                    sts128(obox + (j >> 5) * kOutBox + sw128(lane, (j & 31) >> 2),
                           make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w)));
                }
            }
            if (MODE == 0 || h == kH - 1) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    const int col = MODE == 0 ? h * kDim : 0;
                    const int row0 = (int)(tile * kTile2) + ew * 32;
                    if (p.store_hint) {
                        const uint64_t pol = policy_evict_first();
                        tma_store_2d_hint(&out_map, obox, col, row0, pol);
                        tma_store_2d_hint(&out_map, obox + kOutBox, col + 32, row0, pol);
                    } else {
                        tma_store_2d(&out_map, obox, col, row0);
                        tma_store_2d(&out_map, obox + kOutBox, col + 32, row0);
                    }
                    tma_commit();
                }
            }
        }
        if (lane == 0) tma_wait_all0();      // stores must have landed before the CTA exits
    } else if (lane == 0) {
        // ===== MMA issuer: per (tile, head): 4 K-steps x (hi*hi + lo*hi + hi*lo), M=128 N=80 K=16
        const uint32_t idesc = make_idesc(kTile2, kBN, 0, 0);
        const uint32_t stage_base = smem_u32(stages), b_base = smem_u32(Bop);
        if (p.prepared != nullptr) mbar_wait(&bbar, 0);
        for (int sc = 0; sc < nsc; ++sc) {
            const int s = sc % kNS2, slot = sc % kNAcc, h = sc & 3;
            if (p.pf_tiles > 0 && h == 0 && sc + 4 * p.pf_tiles < nsc) {
                const int64_t ptile = tile_of(sc + 4 * p.pf_tiles);
                const int64_t prow = ptile * kTile2;
                const int64_t nrows = min((int64_t)kTile2, p.N - prow);
                for (int64_t r = 0; r < nrows; r += 16)
                    prefetch_l2(p.q + (prow + r) * kRowF, (uint32_t)(min((int64_t)16, nrows - r) * kRowF * 4));
            }
            if (sc >= kNAcc) mbar_wait(&tempty[slot], ((sc / kNAcc) - 1) & 1);
            mbar_wait(&full[s], (sc / kNS2) & 1);
            tc_fence_after();
            const uint32_t sb = stage_base + s * kStage2, bb = b_base + h * 2 * kBOp;
            const uint32_t d = tmem + slot * kAccCols;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t qhi = make_desc(sb + ks * 32, kKmajLBO, kKmajSBO), qlo = make_desc(sb + kQOp + ks * 32, kKmajLBO, kKmajSBO);
                const uint64_t bhi = make_desc(bb + ks * 32, kKmajLBO, kKmajSBO), blo = make_desc(bb + kBOp + ks * 32, kKmajLBO, kKmajSBO);
                umma(d, qhi, bhi, idesc, ks > 0 ? 1u : 0u);
                umma(d, qlo, bhi, idesc, 1u);
                umma(d, qhi, blo, idesc, 1u);
            }
            umma_commit(&empty[s]);
            umma_commit(&tfull[slot]);
        }
    }
    __syncwarp();
    if (warp == 0) DIF_STAMP(p.dbg, 3);
    tc_fence_before();
    __syncthreads();
    DIF_STAMP(p.dbg, 5);
    if (warp == 12) tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// fp32 [rows][cols] row-major tensor, box = 32 rows x 32 floats (128 B), 128B swizzle
int make_out_map(CUtensorMap* map, float* base, int64_t rows, int64_t cols) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        DIF_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
        DIF_REQUIRE(f && q == cudaDriverEntryPointSuccess, DIF_ECUDA, "cuTensorMapEncodeTiled not available in this driver");
        fn = (EncodeTiledFn)f;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DIF_REQUIRE(r == CUDA_SUCCESS, DIF_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return DIF_OK;
}

int tc_grid(int64_t units) {
    const int sms = sm_count();
    return (int)(units < sms ? (units < 1 ? 1 : units) : sms);
}

// Row partition shared by both passes: contiguous ranges of whole 128-row tiles, one per CTA.
int tc_rows_per_cta(int64_t N, int* grid) {
    int g = tc_grid((N + kTile2 - 1) / kTile2);
    int64_t rpc = (N + g - 1) / g;
    rpc = (rpc + kTile2 - 1) / kTile2 * kTile2;
    g = (int)((N + rpc - 1) / rpc);
    *grid = g;
    return (int)rpc;
}

// DIF_TC_DEBUG_TIMES=1: per-CTA %globaltimer stamps, summarised on stderr after a device sync (debug only)
uint64_t* dbg_buffer() {
    static uint64_t* buf = nullptr;
    static int on = -1;
    if (on < 0) { const char* e = getenv("DIF_TC_DEBUG_TIMES"); on = (e && atoi(e)) ? 1 : 0; }
    if (!on) return nullptr;
    if (!buf) cudaMalloc(&buf, 256 * 8 * sizeof(uint64_t));
    cudaMemset(buf, 0, 256 * 8 * sizeof(uint64_t));
    return buf;
}
void dbg_report(const char* name, uint64_t* buf, int grid) {
    if (!buf) return;
    cudaDeviceSynchronize();
    static uint64_t h[256 * 8];
    cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost);
    uint64_t t0 = ~0ull;
    for (int b = 0; b < grid; ++b) if (h[b * 8] && h[b * 8] < t0) t0 = h[b * 8];
    fprintf(stderr, "[%s] slot: min/avg/max us since first CTA start\n", name);
    for (int s = 0; s < 7; ++s) {
        double mn = 1e30, mx = 0, sum = 0; int n = 0;
        for (int b = 0; b < grid; ++b) { if (!h[b * 8 + s]) continue; double t = (h[b * 8 + s] - t0) * 1e-3; mn = t < mn ? t : mn; mx = t > mx ? t : mx; sum += t; ++n; }
        if (n) fprintf(stderr, "  stamp %d: %7.2f %7.2f %7.2f  (n=%d)\n", s, mn, sum / n, mx, n);
    }
}

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

}  // namespace

bool simple_tc_supported(int64_t N, int H, int Hv, int M, int D) {
    return N >= 1 && H == kH && Hv == kH && M == kDim && D == kDim;
}

int64_t simple_tc_workspace_bytes(int64_t N, int H, int Hv, int M, int D) {
    int grid;
    tc_rows_per_cta(N, &grid);
    // records + one 64-bit ready flag per CTA
    return (int64_t)grid * SimpleLayout{H, Hv, M, D}.wsLen() * (int64_t)sizeof(float) + (int64_t)grid * 8 + 64;
}

int64_t simple_tc_prepared_bytes(int H, int Hv, int M, int D) {
    if (env_int("DIF_TC_P1_TMA", 1) == 0) return 0;      // only the TMA-staged pass 1 emits the operand image
    return (H == kH && Hv == kH && M == kDim && D == kDim) ? (int64_t)kBBytes : 0;
}

int simple_reduce_tc(const float* q, const float* k, const float* v, int64_t N, int H, int Hv, int M, int D,
                     float* partials, void* prepared, void* ws, int64_t ws_bytes, cudaStream_t st,
                     void* const* peer_bufs, int rank, int world, unsigned long long seq) {
    DIF_REQUIRE(simple_tc_supported(N, H, Hv, M, D), DIF_EUNSUPPORTED, "tcgen05 path: unsupported shape");
    DIF_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) & 31) == 0, DIF_EARG, "tcgen05 path: q/k/v must be 32-byte aligned");
    int grid;
    const int rpc = tc_rows_per_cta(N, &grid);
    const SimpleLayout L{H, Hv, M, D};
    DIF_REQUIRE(ws_bytes >= (int64_t)grid * L.wsLen() * 4 + (int64_t)grid * 8, DIF_EARG, "simple_reduce(tcgen05): workspace too small");
    DIF_REQUIRE(prepared == nullptr || ((uintptr_t)prepared & 15) == 0, DIF_EARG, "simple_reduce(tcgen05): prepared buffer must be 16-byte aligned");
    static const int use_tma = env_int("DIF_TC_P1_TMA", 1);      // 1 = TMA-staged producer (default), 0 = register-path producer
    if (use_tma) {
        // cooperative launch: the fused cross-CTA sum spins on per-CTA flags, so all CTAs must be co-resident
        // (grid <= #SMs, 1 CTA/SM); the runtime refuses the launch otherwise instead of deadlocking
        static std::atomic<unsigned long long> epoch_src{0x9E3779B97F4A7C15ull ^ (unsigned long long)(uintptr_t)&epoch_src};
        unsigned long long epoch = epoch_src.fetch_add(0x632BE59BD9B4E019ull) | 1ull;
        DIF_CUDA_OK(cudaFuncSetAttribute(reduce_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem1T));
        uint64_t* dbg = dbg_buffer();
        float* wsf = (float*)ws;
        int64_t ws_len = L.wsLen();
        unsigned long long* flags = (unsigned long long*)(wsf + (int64_t)grid * ws_len);
        uint8_t* prep = (uint8_t*)prepared;
        int rpc_ = rpc;
        static const int tail_pf = env_int("DIF_TC_TAIL_PREFETCH", 0);
        int pf_tiles = (env_int("DIF_TC_P2_VARIANT", 4) & 2) ? 0 : tail_pf;     // matches pass 2's interleaved tile order only
        int pf_grid = tc_grid((N + kTile2 - 1) / kTile2);
        int l2_hints = env_int("DIF_TC_P1_HINTS", 1);
        ShardArgs sh{};
        sh.world = 1;
        if (peer_bufs != nullptr && world > 1) {
            DIF_REQUIRE(world <= kShardMaxRanks && rank >= 0 && rank < world && seq > 0, DIF_EARG, "simple_reduce(sharded): bad rank/world/seq");
            for (int r = 0; r < world; ++r) { DIF_REQUIRE(peer_bufs[r], DIF_EARG, "simple_reduce(sharded): null peer buffer"); sh.bufs[r] = (float*)peer_bufs[r]; }
            sh.rank = rank; sh.world = world; sh.seq = seq;
            sh.slot_floats = (SimpleLayout{H, Hv, M, D}.len() + 63) & ~(int64_t)63;
        }
        static_assert(116 <= kThreadsT, "one thread per slice element");
        void* args[] = {(void*)&q, (void*)&k, (void*)&v, (void*)&N, (void*)&rpc_, (void*)&wsf, (void*)&ws_len, (void*)&flags,
                        (void*)&epoch, (void*)&partials, (void*)&prep, (void*)&pf_tiles, (void*)&pf_grid, (void*)&l2_hints, (void*)&sh, (void*)&dbg};
        DIF_CUDA_OK(cudaLaunchCooperativeKernel((const void*)reduce_tma_kernel, dim3(grid), dim3(kThreadsT), args, (size_t)kSmem1T, st));
        dbg_report("reduce_tma", dbg, grid);
        return DIF_OK;
    }
    DIF_REQUIRE(peer_bufs == nullptr || world <= 1, DIF_EUNSUPPORTED, "simple_reduce(sharded) needs the TMA-staged pass 1 (DIF_TC_P1_TMA=1)");
    static const int variant = env_int("DIF_TC_P1_VARIANT", 2);   // tuning switches: 1 = register ring, 2 = K/V evict_first, 4 = Q evict_last (+ policy-hinted prefetch)
    static const int pf = env_int("DIF_TC_P1_PREFETCH", 0);       // L2 prefetch distance in 16-row stages (0 = off)
#define DIF_P1(R, E)                                                                                                  \
    do {                                                                                                              \
        DIF_CUDA_OK(cudaFuncSetAttribute(reduce_tc_kernel<R, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem1)); \
        reduce_tc_kernel<R, E><<<grid, kThreadsTC, kSmem1, st>>>(q, k, v, N, rpc, (float*)ws, L.wsLen(), pf, (variant >> 2) & 1);             \
    } while (0)
    switch (variant & 3) {
        case 0: DIF_P1(false, false); break;
        case 1: DIF_P1(true, false); break;
        case 2: DIF_P1(false, true); break;
        default: DIF_P1(true, true); break;
    }
#undef DIF_P1
    DIF_LAUNCH_OK();
    return simple_finalize_fwd((const float*)ws, grid, H, Hv, M, D, partials, st);
}

int simple_apply_tc(const float* q, const float* partials, const void* prepared, double n_total, int64_t N, int H, int Hv, int M, int D,
                    float* out, const dif_epilogue_t* ep, cudaStream_t st) {
    DIF_REQUIRE(simple_tc_supported(N, H, Hv, M, D), DIF_EUNSUPPORTED, "tcgen05 path: unsupported shape");
    DIF_REQUIRE(((uintptr_t)q & 31) == 0 && ((uintptr_t)out & 15) == 0, DIF_EARG, "tcgen05 path: q must be 32-byte, out 16-byte aligned");
    ApplyTcArgs a{};
    a.q = q; a.partials = partials; a.n_total = (float)n_total; a.N = N; a.out = out;
    a.prepared = (const uint8_t*)prepared;
    DIF_REQUIRE(prepared == nullptr || ((uintptr_t)prepared & 15) == 0, DIF_EARG, "simple_apply(tcgen05): prepared buffer must be 16-byte aligned");
    if (ep) a.ep = *ep; else { a.ep.mode = 0; a.ep.n_add = 0; }
    DIF_REQUIRE(a.ep.mode == 0 || a.ep.mode == 1, DIF_EARG, "simple_apply: epilogue mode %d", a.ep.mode);
    DIF_REQUIRE(N < (1ll << 31), DIF_EUNSUPPORTED, "tcgen05 path: N must fit a 32-bit TMA coordinate");
    // tuning switches: 1 = register ring, 2 = contiguous reversed tile order (pass-1 partition), 4 = evict_first stores
    static const int variant = env_int("DIF_TC_P2_VARIANT", 4);
    int grid;
    if (variant & 2) {
        a.tiles_per_cta = tc_rows_per_cta(N, &grid) / kTile2;
    } else {
        a.tiles_per_cta = 0;
        grid = tc_grid((N + kTile2 - 1) / kTile2);
    }
    a.store_hint = (variant & 4) ? 1 : 0;
    a.pf_tiles = env_int("DIF_TC_P2_PREFETCH", 1);
    a.dbg = dbg_buffer();
    CUtensorMap map;
    int rc = make_out_map(&map, out, N, a.ep.mode == 0 ? (int64_t)kH * kDim : (int64_t)kDim);
    if (rc) return rc;
#define DIF_P2(MODE, R)                                                                                                    \
    do {                                                                                                                   \
        DIF_CUDA_OK(cudaFuncSetAttribute(apply_tc_kernel<MODE, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2));  \
        apply_tc_kernel<MODE, R><<<grid, kThreadsTC, kSmem2, st>>>(a, map);                                                \
    } while (0)
    if (a.ep.mode == 0) { if (variant & 1) DIF_P2(0, true); else DIF_P2(0, false); }
    else                { if (variant & 1) DIF_P2(1, true); else DIF_P2(1, false); }
#undef DIF_P2
    dbg_report("apply_tc", a.dbg, grid);
    DIF_LAUNCH_OK();
    return DIF_OK;
}

}  // namespace dif
