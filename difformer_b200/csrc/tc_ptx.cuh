// PTX wrappers shared by the tcgen05 kernels (simple_sm100.cu, sigmoid_sm100.cu): mbarrier, tcgen05 (alloc / mma /
// commit / ld), TMA (bulk + tensor), UMMA descriptors, bf16 hi/lo split, 128B swizzle.  sm_100a only.
#pragma once
#include <cuda.h>        // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace dif {
namespace {

constexpr int kDim = 64;
constexpr int kWarps = 13;
constexpr int kThreadsTC = kWarps * 32;
constexpr int kDbgSlots = 16;                       // per-CTA %globaltimer stamps of the debug timeline

// ---- descriptor conventions (verified on hardware by csrc/probe_umma.cu) --------------------
constexpr uint32_t kSwizzle128 = 2;
// K-major SW128: 8-row groups 1024 B apart (SBO); LBO unused
constexpr uint32_t kKmajLBO = 0, kKmajSBO = 1024;
// MN-major SW128: 64-element (128 B) MN blocks are LBO apart, 8-k groups SBO (= 1024 B) apart

// ---- optional in-kernel timeline (DIF_TC_DEBUG_TIMES=1): thread 0 of every CTA stamps %globaltimer
__device__ __forceinline__ uint64_t gtime() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t smid() { uint32_t r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define DIF_STAMP(buf, slot)                                                        \
    do {                                                                            \
        if ((buf) != nullptr && threadIdx.x == 0) {                                 \
            (buf)[blockIdx.x * kDbgSlots + (slot)] = gtime();                       \
            if ((slot) == 0) (buf)[blockIdx.x * kDbgSlots + kDbgSlots - 1] = smid() + 1; \
        }                                                                           \
    } while (0)

// the same from any single thread the caller elects (role warps other than warp 0)
#define DIF_STAMP_ANY(buf, slot)                                                    \
    do {                                                                            \
        if ((buf) != nullptr) (buf)[blockIdx.x * kDbgSlots + (slot)] = gtime();     \
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
// One lane of a converged warp (elect.sync).  Issuing tcgen05.mma under this predicate -- instead of under
// `lane == 0` -- lets the compiler emit the UTCHMMA directly; a plain lane test makes it wrap every MMA in a
// thread-serialising loop (~70 clk per instruction on the issuing thread).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\tselp.u32 %0, 1, 0, pe;\n\t}" : "=r"(pred));
    return pred != 0;
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
// A operand in tensor memory ("TS" form): lane = row m, each 32-bit column holds two K-consecutive 16-bit elements
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, pp;\n\t}"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                    "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
                    "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
                    "r"(r[30]), "r"(r[31])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
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
// TMA 2-D tensor load: global box -> shared (128B-swizzled), completion on an mbarrier (rows beyond the tensor are zero-filled)
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                 :: "r"(smem_dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
// Ampere-style asynchronous copy global -> shared, 16 bytes per thread (no registers held while in flight)
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
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



}  // namespace

// fp32 [rows][cols] row-major tensor map, box = 32 rows x 32 floats (128 B), 128B swizzle (cached); simple_sm100.cu
int make_out_map(CUtensorMap* map, float* base, int64_t rows, int64_t cols);

}  // namespace dif
