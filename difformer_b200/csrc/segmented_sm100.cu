// Batched-graph kernel='simple' forward on the tensor cores (sm_100a) -- TransConv.full_attention(..., 'simple', n_nodes),
// physical particle/difformer-v2.py:80-111: every graph g of the batch attends only inside itself,
//     out_i = sum_{j in g(i)} (1 + c q_i.k_j) v_j / sum_{j in g(i)} (1 + c q_i.k_j),      c = 1 / (|Q|_F |K|_F) over the whole batch.
//
// The particle graphs have 10-40 nodes: per graph the O(n) form (S_g = K_g^T V_g, 64 x 64 x n) is more arithmetic than the direct
// O(n^2) form, and one warp per graph (segmented.cu) is issue-bound at ~7000 instructions per graph.  Here whole graphs are packed
// into 128-row tiles (contiguous row ranges, no data movement: the plan below) and a tile runs as block-diagonal dense attention:
//   Sc = Q K^T            tcgen05.mma M = N = 128, K = 64, both operands K-major, bf16 hi/lo split (3 MMAs per product)
//   W  = mask o (1 + c Sc) 256 threads (row x column half): tcgen05.ld, the row's graph is a column range [gs, ge) -> no per-column
//                         lookup; W is split into bf16 hi + lo (hi = 1 exactly for the usual tiny c s, lo carries c s) and written
//                         as the K-major A operand of the second product; the fp32 row sums are the denominators
//   O  = W [Vhi | Vlo]    tcgen05.mma M = 128, N = 128, K = 128: B = the V tile read MN-major (nodes = K index); the epilogue adds the
//                         two 64-column halves and divides by the row sum
// Warp roles: 0-7 load + split Q, K, V (global -> registers -> swizzled operands), 8-15 W pass + epilogue, 16 MMA issuer.
//
// Plan (dif_segmented_plan_build, once per batch layout): tile b holds the graphs whose FIRST row lies in [b S, (b+1) S) with
// S = 129 - max_nodes, so a tile spans < S + max_nodes = 129 rows, tiles are independent of each other (no sequential packing) and
// the fill is S / 128 (70 % at max_nodes = 40); row_range[r] = (first row, end row) of the graph of row r.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dif {
int make_out_map(CUtensorMap* map, float* base, int64_t rows, int64_t cols);      // simple_sm100.cu: [rows, cols] fp32, box 32 x 32, 128B swizzle (cached)

namespace {

constexpr int kST = 128;                 // rows per tile
constexpr int kSOp = kST * 128;          // one bf16 operand tile [128 rows][64]: 16 KB
constexpr int kSegTcWarps = 17, kSegTcThreads = kSegTcWarps * 32;
constexpr int kSegTcSmem = 3 * 2 * kSOp + 4 * kSOp + 1024;      // Q, K, V (hi | lo) + W (hi: 2 K blocks, lo: 2 K blocks)

struct SegTcArgs {
    const float *q, *k, *v;
    const int2* row_range;
    const int* tile_row0;
    int ntiles;
    const float* norms;                  // [sum q^2, sum k^2] over the whole batch
    float* out;
    unsigned long long* dbg;             // optional timeline of CTA 0 (DIF_SEG_DEBUG=1): [tile < 8][event < 8] %globaltimer
};

#define SEG_STAMP(ev)                                                                                   \
    do {                                                                                                \
        if (p.dbg != nullptr && blockIdx.x == 0 && it < 8) p.dbg[it * 8 + (ev)] = gtime();             \
    } while (0)

__device__ __forceinline__ void seg_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }

__global__ void seg_plan_kernel(const int32_t* __restrict__ seg, int B, int64_t N, int S, int ntiles, int* __restrict__ tile_row0,
                                int2* __restrict__ row_range) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B) return;
    const int s = seg[g], e = seg[g + 1];
    if (e <= s) return;                                       // empty graph: owns no rows
    for (int r = s; r < e; ++r) row_range[r] = make_int2(s, e);
    int pb = -1;                                              // bucket of the previous non-empty graph
    if (s > 0) {
        int gp = g - 1;
        while (gp > 0 && seg[gp] == s) --gp;                  // skip empty graphs
        pb = seg[gp] / S;
    }
    const int b = s / S;
    for (int bb = pb + 1; bb <= b; ++bb) tile_row0[bb] = s;   // this graph opens bucket b (and any empty buckets before it)
    if (e == (int)N)
        for (int bb = b + 1; bb <= ntiles; ++bb) tile_row0[bb] = (int)N;
}

// 128 rows x 64 floats of `src` starting at row0 (rows >= row1 read as zero) -> bf16 hi | lo K-major SW128 operand at s_hi.  256 threads.
__device__ __forceinline__ void seg_load(const float* src, int64_t row0, int64_t row1, int tid, float (&x)[4][8]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int t = tid + 256 * j;
        const int64_t row = row0 + (t >> 3);
        if (row < row1) ldg256_keep(src + row * kDim + (t & 7) * 8, x[j]);
        else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[j][i] = 0.f;
        }
    }
}
__device__ __forceinline__ void seg_store(uint32_t s_hi, int tid, const float (&x)[4][8]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int t = tid + 256 * j;
        uint4 hi, lo;
        split8(x[j], hi, lo);
        const uint32_t off = sw128(t >> 3, t & 7);
        sts128(s_hi + off, hi);
        sts128(s_hi + kSOp + off, lo);
    }
}

// Output rows leave through TMA: a warp's 32 rows x 32 columns are staged in the (by then idle) weight buffer, in the very 4 KB the warp
// itself fills during the weight pass, and stored with one cp.async.bulk.tensor; 32 scattered 16-byte st.global per instruction cost a
// wavefront each on the LSU that the operand warps need.  Warps whose rows straddle the end of the tile store directly.
__global__ void __launch_bounds__(kSegTcThreads, 1) seg_fwd_tc_kernel(const __grid_constant__ SegTcArgs p, const __grid_constant__ CUtensorMap out_map) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t Qop = smem_u32(base), Kop = Qop + 2 * kSOp, Vop = Kop + 2 * kSOp, Whi = Vop + 2 * kSOp, Wlo = Whi + 2 * kSOp;
    __shared__ uint64_t qk_full, v_full, s_full, w_full, o_full, s_free, o_free;
    __shared__ uint32_t tmem_slot;
    __shared__ float den_s[2][2][kST];                        // [tile parity][column half][row]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int my_tiles = (int)blockIdx.x < p.ntiles ? (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        mbar_init(&qk_full, 8); mbar_init(&v_full, 8); mbar_init(&w_full, 8);
        mbar_init(&s_full, 1); mbar_init(&o_full, 1);
        mbar_init(&s_free, 8); mbar_init(&o_free, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) tmem_alloc(&tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t tmemS = tmem, tmemO = tmem + kST;          // columns 0-127: Sc ; 128-255: O = [W Vhi | W Vlo]

    if (warp < 8) {
        // ===== producers: Q, K then V of the tile -> operands.  Q and K may be overwritten once the previous tile's first product has
        // completed (s_full), V once its second product has (o_full); the loads themselves are issued before those waits.
        // Register schedule (two sets of 32 floats per thread, swapped every tile): on entry A = Q(it), B = K(it); V(it) and Q(it+1) are
        // requested together right after the first product's operands are published, K(it+1) right after V is; the rows of tile it+2
        // are pulled into L2 meanwhile, so the register loads are L2 hits.
        float xa[4][8], xb[4][8];
        auto l2_ahead = [&](int tile) {
            if (tid != 0 || tile >= p.ntiles) return;
            const int64_t r0 = p.tile_row0[tile], r1 = p.tile_row0[tile + 1];
            if (r1 <= r0) return;
            const uint32_t bytes = (uint32_t)((r1 - r0) * kDim * 4);
            prefetch_l2(p.q + r0 * kDim, bytes);
            prefetch_l2(p.k + r0 * kDim, bytes);
            prefetch_l2(p.v + r0 * kDim, bytes);
        };
        auto body = [&](int it, float (&A)[4][8], float (&B)[4][8]) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int64_t r0 = p.tile_row0[tile], r1 = p.tile_row0[tile + 1];
            const bool has_next = it + 1 < my_tiles;
            int64_t n0 = 0, n1 = 0;
            if (has_next) { n0 = p.tile_row0[tile + gridDim.x]; n1 = p.tile_row0[tile + gridDim.x + 1]; }
            l2_ahead(tile + 2 * gridDim.x);
            if (it > 0) mbar_wait(&s_full, (it - 1) & 1);
            seg_store(Qop, tid, A);
            seg_store(Kop, tid, B);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&qk_full);
            if (tid == 0) SEG_STAMP(0);
            seg_load(p.v, r0, r1, tid, A);
            if (has_next) seg_load(p.q, n0, n1, tid, B);
            if (it > 0) mbar_wait(&o_full, (it - 1) & 1);
            seg_store(Vop, tid, A);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&v_full);
            if (tid == 0) SEG_STAMP(1);
            if (has_next) seg_load(p.k, n0, n1, tid, A);       // exit: B = Q(it+1), A = K(it+1)
        };
        if (my_tiles > 0) {
            const int64_t r0 = p.tile_row0[blockIdx.x], r1 = p.tile_row0[blockIdx.x + 1];
            seg_load(p.q, r0, r1, tid, xa);
            seg_load(p.k, r0, r1, tid, xb);
            l2_ahead(blockIdx.x + gridDim.x);
        }
        for (int it = 0; it < my_tiles; it += 2) {
            body(it, xa, xb);
            if (it + 1 < my_tiles) body(it + 1, xb, xa);
        }
    } else if (warp < 16) {
        // ===== W pass + epilogue: thread = (tile row i, column half).  TMEM lane = 32 quad + lane.
        // These eight warps are the critical resource of a tile (scores -> W -> wait for W V -> output), so: the tile's row range and
        // the row's graph come from registers loaded during the previous tile's wait; both 32-column TMEM loads of a phase are in
        // flight together; 8-column groups outside the row's graph are stored as zeros without touching the scores.
        const int ew = warp - 8, quad = ew & 3, half = ew >> 2;
        const int i = quad * 32 + lane;
        const uint32_t tlane = (uint32_t)(quad * 32) << 16;
        const float c = 1.f / (sqrtf(p.norms[0]) * sqrtf(p.norms[1]));
        int r0 = 0, r1 = 0, gs = 0, ge = 0;                    // tile rows [r0, r1); this row's graph = columns [gs, ge) of the tile
        auto tile_rows = [&](int it_, int& a0, int& a1, int& s_, int& e_) {
            const int tile = blockIdx.x + it_ * gridDim.x;
            a0 = p.tile_row0[tile];
            a1 = p.tile_row0[tile + 1];
            s_ = e_ = 0;
            if (a0 + i < a1) { const int2 rg = p.row_range[a0 + i]; s_ = rg.x - a0; e_ = rg.y - a0; }
        };
        if (my_tiles > 0) tile_rows(0, r0, r1, gs, ge);
        for (int it = 0; it < my_tiles; ++it) {
            const int row = r0 + i;
            const bool valid = row < r1;
            mbar_wait(&s_full, it & 1);
            tc_fence_after();
            if (ew == 0 && lane == 0) SEG_STAMP(2);
            uint32_t ra[32], rb[32];
            tmem_ld32(tmemS + tlane + 64 * half, ra);
            tmem_ld32(tmemS + tlane + 64 * half + 32, rb);
            tmem_ld_wait32(ra);
            tmem_ld_wait32(rb);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&s_free); tma_wait_read0(); }      // scores in registers; the previous output store has read this warp's part of the buffer
            __syncwarp();
            float den = 0.f;
#pragma unroll
            for (int g8 = 0; g8 < 8; ++g8) {
                const int jb = 64 * half + 8 * g8;
                const uint32_t off = (uint32_t)(half * kSOp) + sw128(i, g8);          // K block = column half
                if (jb + 8 <= gs || jb >= ge) {
                    sts128(Whi + off, make_uint4(0u, 0u, 0u, 0u));
                    sts128(Wlo + off, make_uint4(0u, 0u, 0u, 0u));
                } else {
                    float w[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int j = jb + e;
                        const uint32_t sc = g8 < 4 ? ra[8 * g8 + e] : rb[8 * (g8 - 4) + e];
                        w[e] = (j >= gs && j < ge) ? fmaf(c, __uint_as_float(sc), 1.f) : 0.f;
                        den += w[e];
                    }
                    uint4 hi, lo;
                    split8(w, hi, lo);
                    sts128(Whi + off, hi);
                    sts128(Wlo + off, lo);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&w_full);
            if (ew == 0 && lane == 0) SEG_STAMP(3);
            den_s[it & 1][half][i] = den;
            int nr0 = 0, nr1 = 0, ngs = 0, nge = 0;
            if (it + 1 < my_tiles) tile_rows(it + 1, nr0, nr1, ngs, nge);     // in flight while the second product runs
            seg_bar_sync(1 + quad, 64);                        // the two column halves of the rows 32 quad .. +31
            const float inv = 1.f / (den_s[it & 1][0][i] + den_s[it & 1][1][i]);
            mbar_wait(&o_full, it & 1);
            tc_fence_after();
            if (ew == 0 && lane == 0) SEG_STAMP(4);
            tmem_ld32(tmemO + tlane + 32 * half, ra);
            tmem_ld32(tmemO + tlane + kDim + 32 * half, rb);
            tmem_ld_wait32(ra);
            tmem_ld_wait32(rb);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free);
            const bool whole = r0 + 32 * quad + 32 <= r1;      // warp-uniform: all 32 rows of this warp belong to the tile
            const uint32_t box = Whi + (uint32_t)(half * kSOp + quad * 4096);
            float* dst = p.out + (int64_t)row * kDim + 32 * half;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 o;
                o.x = (__uint_as_float(ra[j]) + __uint_as_float(rb[j])) * inv;
                o.y = (__uint_as_float(ra[j + 1]) + __uint_as_float(rb[j + 1])) * inv;
                o.z = (__uint_as_float(ra[j + 2]) + __uint_as_float(rb[j + 2])) * inv;
                o.w = (__uint_as_float(ra[j + 3]) + __uint_as_float(rb[j + 3])) * inv;
                if (whole) sts128(box + sw128(lane, j >> 2), make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w)));
                else if (valid) *reinterpret_cast<float4*>(dst + j) = o;
            }
            if (whole) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { tma_store_2d(&out_map, box, 32 * half, r0 + 32 * quad); tma_commit(); }
            }
            if (ew == 0 && lane == 0) SEG_STAMP(5);
            r0 = nr0; r1 = nr1; gs = ngs; ge = nge;
        }
        if (lane == 0) tma_wait_all0();                        // stores must have landed before the CTA exits
    } else if (lane == 0) {
        // ===== MMA issuer
        const uint32_t idS = make_idesc(kST, kST, 0, 0);          // Sc = Q K^T: both operands K-major
        const uint32_t idO = make_idesc(kST, 2 * kDim, 0, 1);     // [W Vhi | W Vlo]: A = W K-major, B = V MN-major (hi | lo: two 64-blocks, LBO apart)
        const uint32_t idOl = make_idesc(kST, kDim, 0, 1);        // N = 64: the hi block of V only
        for (int it = 0; it < my_tiles; ++it) {
            if (it > 0) mbar_wait(&s_free, (it - 1) & 1);
            mbar_wait(&qk_full, it & 1);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t qhi = make_desc(Qop + ks * 32, kKmajLBO, kKmajSBO), qlo = make_desc(Qop + kSOp + ks * 32, kKmajLBO, kKmajSBO);
                const uint64_t khi = make_desc(Kop + ks * 32, kKmajLBO, kKmajSBO), klo = make_desc(Kop + kSOp + ks * 32, kKmajLBO, kKmajSBO);
                umma(tmemS, qhi, khi, idS, ks > 0 ? 1u : 0u);
                umma(tmemS, qlo, khi, idS, 1u);
                umma(tmemS, qhi, klo, idS, 1u);
            }
            umma_commit(&s_full);
            SEG_STAMP(6);
            if (it > 0) mbar_wait(&o_free, (it - 1) & 1);
            mbar_wait(&v_full, it & 1);
            mbar_wait(&w_full, it & 1);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const uint32_t wo = (uint32_t)((ks >> 2) * kSOp + (ks & 3) * 32);
                const uint64_t whi = make_desc(Whi + wo, kKmajLBO, kKmajSBO), wlo = make_desc(Wlo + wo, kKmajLBO, kKmajSBO);
                const uint64_t vb = make_desc(Vop + ks * 2048, kSOp, 1024);
                umma(tmemO, whi, vb, idO, ks > 0 ? 1u : 0u);      // Whi [Vhi | Vlo]
                umma(tmemO, wlo, vb, idOl, 1u);                    // Wlo Vhi onto the first 64 columns (Wlo Vlo is below fp32 resolution)
            }
            umma_commit(&o_full);
            SEG_STAMP(7);
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem, 256);
}


// ------------------------------------------------------------------------------------------------------------------------------------
// Backward of the tiled form (same plan, same tiles).  With W = mask o (1 + c S), den = W 1, out = W V / den and G = dL/dout:
//     dV  = W'^T G,   W' = diag(1/den) W
//     dW  = mask o (diag(1/den) G V^T + dden 1^T),   dden_i = -(g_i . out_i) / den_i
//     dS  = c dW,     dQ' = dS K,   dK' = dS^T Q,    t = sum dW o (c S)    (c = 1/(|Q||K|) is global: dQ = dQ' - Q t/|Q|^2, dK alike,
//                                                     applied by seg_fixup_rows_kernel once t has been summed over the batch)
// Per tile: [S = Q K^T | P = G V^T] -> pass 1 of the eight weight warps writes W' (bf16 hi/lo, K-major) -> dV = W'^T G (W' read as the
// MN-major A operand) -> pass 2 re-reads S and P from tensor memory and overwrites the same buffer with dS -> dQ' = dS K and
// dK' = dS^T Q -> the three 64-column accumulators are stored.  Tensor memory: S | P | dV | dQ | dK = 448 columns.
struct SegBwdTcArgs {
    const float *q, *k, *v, *g, *out;
    const int2* row_range;
    const int* tile_row0;
    int ntiles;
    const float* norms;
    float *dq, *dk, *dv;
    float* part;                         // per-CTA share of t: part[2 cta] = part[2 cta + 1]
    unsigned long long* dbg;             // optional timeline of CTA 0 (DIF_SEG_DEBUG=1)
};

constexpr int kSegBwdSmem = 4 * 2 * kSOp + 4 * kSOp + 1024;     // Q, K, V, G (hi | lo) + the W' / dS buffer (hi: 2 K blocks, lo: 2 K blocks)

__global__ void __launch_bounds__(kSegTcThreads, 1) seg_bwd_tc_kernel(const __grid_constant__ SegBwdTcArgs p, const __grid_constant__ CUtensorMap dv_map,
                                                                       const __grid_constant__ CUtensorMap dq_map, const __grid_constant__ CUtensorMap dk_map) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t Qop = smem_u32(base), Kop = Qop + 2 * kSOp, Vop = Kop + 2 * kSOp, Gop = Vop + 2 * kSOp, Xhi = Gop + 2 * kSOp, Xlo = Xhi + 2 * kSOp;
    __shared__ uint64_t in_full, sp_full, x1_full, dv_full, x2_full, c_full, e_free;
    __shared__ uint32_t tmem_slot;
    __shared__ float den_s[2][2][kST], gdo_s[2][2][kST];      // [tile parity][column half][row]: row sums of W, g . out
    __shared__ float t_red[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int my_tiles = (int)blockIdx.x < p.ntiles ? (p.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        mbar_init(&in_full, 8); mbar_init(&x1_full, 8); mbar_init(&x2_full, 8); mbar_init(&e_free, 8);
        mbar_init(&sp_full, 1); mbar_init(&dv_full, 1); mbar_init(&c_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t tmS = tmem, tmP = tmem + 128, tmDV = tmem + 256, tmDQ = tmem + 320, tmDK = tmem + 384;

    if (warp < 8) {
        // ===== producers.  V is free once S | P of the previous tile are complete, G once its dV is, Q and K once its dQ | dK are.
        // Two register sets: V, G of the next tile are requested as soon as this tile's operands are published.
        float xa[4][8], xb[4][8];
        auto rows = [&](int it_, int64_t& a0, int64_t& a1) {
            const int tile = blockIdx.x + it_ * gridDim.x;
            a0 = p.tile_row0[tile];
            a1 = p.tile_row0[tile + 1];
        };
        auto l2_ahead = [&](int it_) {
            if (tid != 0 || it_ >= my_tiles) return;
            int64_t a0, a1;
            rows(it_, a0, a1);
            if (a1 <= a0) return;
            const uint32_t bytes = (uint32_t)((a1 - a0) * kDim * 4);
            prefetch_l2(p.q + a0 * kDim, bytes); prefetch_l2(p.k + a0 * kDim, bytes);
            prefetch_l2(p.v + a0 * kDim, bytes); prefetch_l2(p.g + a0 * kDim, bytes);
        };
        int64_t r0 = 0, r1 = 0;
        if (my_tiles > 0) {
            rows(0, r0, r1);
            seg_load(p.v, r0, r1, tid, xa);
            seg_load(p.g, r0, r1, tid, xb);
            l2_ahead(1);
        }
        for (int it = 0; it < my_tiles; ++it) {
            if (it > 0) mbar_wait(&sp_full, (it - 1) & 1);
            seg_store(Vop, tid, xa);
            if (it > 0) mbar_wait(&dv_full, (it - 1) & 1);
            seg_store(Gop, tid, xb);
            seg_load(p.q, r0, r1, tid, xa);
            seg_load(p.k, r0, r1, tid, xb);
            l2_ahead(it + 2);
            if (it > 0) mbar_wait(&c_full, (it - 1) & 1);
            seg_store(Qop, tid, xa);
            seg_store(Kop, tid, xb);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&in_full);
            if (tid == 0) SEG_STAMP(0);
            if (it + 1 < my_tiles) {
                rows(it + 1, r0, r1);
                seg_load(p.v, r0, r1, tid, xa);
                seg_load(p.g, r0, r1, tid, xb);
            }
        }
    } else if (warp < 16) {
        // ===== weight passes + epilogue: thread = (tile row i, column half)
        const int ew = warp - 8, quad = ew & 3, half = ew >> 2;
        const int i = quad * 32 + lane;
        const uint32_t tlane = (uint32_t)(quad * 32) << 16;
        const float c = 1.f / (sqrtf(p.norms[0]) * sqrtf(p.norms[1]));
        int r0 = 0, r1 = 0, gs = 0, ge = 0;
        float gdo = 0.f;                                       // this thread's half of g_i . out_i
        auto tile_rows = [&](int it_, int& a0, int& a1, int& s_, int& e_, float& gd_) {
            const int tile = blockIdx.x + it_ * gridDim.x;
            a0 = p.tile_row0[tile];
            a1 = p.tile_row0[tile + 1];
            s_ = e_ = 0;
            gd_ = 0.f;
            if (a0 + i < a1) {
                const int2 rg = p.row_range[a0 + i];
                s_ = rg.x - a0; e_ = rg.y - a0;
                const float* gp = p.g + (int64_t)(a0 + i) * kDim + 32 * half;
                const float* op = p.out + (int64_t)(a0 + i) * kDim + 32 * half;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 a = ldg4(gp + 4 * j), b = ldg4(op + 4 * j);
                    gd_ = fmaf(a.x, b.x, gd_); gd_ = fmaf(a.y, b.y, gd_); gd_ = fmaf(a.z, b.z, gd_); gd_ = fmaf(a.w, b.w, gd_);
                }
            }
        };
        if (my_tiles > 0) tile_rows(0, r0, r1, gs, ge, gdo);
        float t_acc = 0.f;
        for (int it = 0; it < my_tiles; ++it) {
            const int row = r0 + i;
            const bool valid = row < r1;
            // ---- pass 1: W' = diag(1/den) mask o (1 + c S).  Two sweeps over the row's 64 scores (row sum, then the scaled weights):
            // tensor memory is re-read rather than 64 values kept in registers; 8-column groups outside the row's graph cost nothing.
            mbar_wait(&sp_full, it & 1);
            tc_fence_after();
            if (ew == 0 && lane == 0) SEG_STAMP(1);
            uint32_t ra[32], rb[32];
            float den = 0.f;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                tmem_ld32(tmS + tlane + 64 * half + 32 * ch, ra);
                tmem_ld_wait32(ra);
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    const int jb = 64 * half + 32 * ch + 8 * g8;
                    if (jb + 8 <= gs || jb >= ge) continue;
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        den += (jb + e >= gs && jb + e < ge) ? fmaf(c, __uint_as_float(ra[8 * g8 + e]), 1.f) : 0.f;
                }
            }
            den_s[it & 1][half][i] = den;
            gdo_s[it & 1][half][i] = gdo;
            if (lane == 0) tma_wait_read0();                   // the previous tile's output stores have read this warp's part of the buffer
            seg_bar_sync(1 + quad, 64);
            const float inv = valid ? 1.f / (den_s[it & 1][0][i] + den_s[it & 1][1][i]) : 0.f;
            const float dden = -(gdo_s[it & 1][0][i] + gdo_s[it & 1][1][i]) * inv;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                tmem_ld32(tmS + tlane + 64 * half + 32 * ch, ra);
                tmem_ld_wait32(ra);
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    const int jb = 64 * half + 32 * ch + 8 * g8;
                    const uint32_t off = (uint32_t)(half * kSOp) + sw128(i, 4 * ch + g8);
                    if (jb + 8 <= gs || jb >= ge) {
                        sts128(Xhi + off, make_uint4(0u, 0u, 0u, 0u));
                        sts128(Xlo + off, make_uint4(0u, 0u, 0u, 0u));
                    } else {
                        float w[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            w[e] = (jb + e >= gs && jb + e < ge) ? fmaf(c, __uint_as_float(ra[8 * g8 + e]), 1.f) * inv : 0.f;
                        uint4 hi, lo;
                        split8(w, hi, lo);
                        sts128(Xhi + off, hi);
                        sts128(Xlo + off, lo);
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&x1_full);
            if (ew == 0 && lane == 0) SEG_STAMP(2);
            // ---- pass 2: dS = c mask o (P / den + dden) over the same buffer, once dV = W'^T G has read it
            mbar_wait(&dv_full, it & 1);
            tc_fence_after();
            if (ew == 0 && lane == 0) SEG_STAMP(3);
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                tmem_ld32(tmS + tlane + 64 * half + 32 * ch, ra);
                tmem_ld32(tmP + tlane + 64 * half + 32 * ch, rb);
                tmem_ld_wait32(ra);
                tmem_ld_wait32(rb);
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    const int jb = 64 * half + 32 * ch + 8 * g8;
                    const uint32_t off = (uint32_t)(half * kSOp) + sw128(i, 4 * ch + g8);
                    if (jb + 8 <= gs || jb >= ge) {
                        sts128(Xhi + off, make_uint4(0u, 0u, 0u, 0u));
                        sts128(Xlo + off, make_uint4(0u, 0u, 0u, 0u));
                    } else {
                        float ds[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float dw = (jb + e >= gs && jb + e < ge) ? fmaf(__uint_as_float(rb[8 * g8 + e]), inv, dden) : 0.f;
                            const float cdw = c * dw;
                            t_acc = fmaf(cdw, __uint_as_float(ra[8 * g8 + e]), t_acc);
                            ds[e] = cdw;
                        }
                        uint4 hi, lo;
                        split8(ds, hi, lo);
                        sts128(Xhi + off, hi);
                        sts128(Xlo + off, lo);
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&x2_full);
            if (ew == 0 && lane == 0) SEG_STAMP(4);
            int nr0 = 0, nr1 = 0, ngs = 0, nge = 0;
            float ngdo = 0.f;
            if (it + 1 < my_tiles) tile_rows(it + 1, nr0, nr1, ngs, nge, ngdo);       // in flight while dQ | dK run
            // ---- epilogue: dV, dQ', dK' rows
            mbar_wait(&c_full, it & 1);
            tc_fence_after();
            if (ew == 0 && lane == 0) SEG_STAMP(5);
            // rows leave through TMA, staged in this warp's own 4 KB of the (now idle) W' / dS buffer (see seg_fwd_tc_kernel)
            const bool whole = r0 + 32 * quad + 32 <= r1;
            const uint32_t box0 = Xhi + (uint32_t)(half * kSOp + quad * 4096), box1 = Xlo + (uint32_t)(half * kSOp + quad * 4096);
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                const uint32_t box = o == 1 ? box1 : box0;             // two boxes: the store of dV is being read while dQ' is staged
                tmem_ld32((o == 0 ? tmDV : o == 1 ? tmDQ : tmDK) + tlane + 32 * half, ra);
                tmem_ld_wait32(ra);
                if (o == 2) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&e_free);
                }
                if (whole) {
                    if (o == 2) {                                     // dV's store (two groups back) has read box0
                        if (lane == 0) tma_wait_read1();
                        __syncwarp();
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4) sts128(box + sw128(lane, j >> 2), make_uint4(ra[j], ra[j + 1], ra[j + 2], ra[j + 3]));
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { tma_store_2d(o == 0 ? &dv_map : o == 1 ? &dq_map : &dk_map, box, 32 * half, r0 + 32 * quad); tma_commit(); }
                } else if (valid) {
                    float* dst = (o == 0 ? p.dv : o == 1 ? p.dq : p.dk) + (int64_t)row * kDim + 32 * half;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<uint4*>(dst + j) = make_uint4(ra[j], ra[j + 1], ra[j + 2], ra[j + 3]);
                }
            }
            if (ew == 0 && lane == 0) SEG_STAMP(6);
            r0 = nr0; r1 = nr1; gs = ngs; ge = nge; gdo = ngdo;
        }
        if (lane == 0) tma_wait_all0();                        // stores must have landed before the CTA exits
        // this CTA's share of t (fixed order: lanes, then warps)
        for (int o = 16; o > 0; o >>= 1) t_acc += __shfl_xor_sync(0xffffffffu, t_acc, o);
        if (lane == 0) t_red[ew] = t_acc;
        asm volatile("bar.sync 5, 256;" ::: "memory");
        if (ew == 0 && lane == 0) {
            float t = 0.f;
            for (int w = 0; w < 8; ++w) t += t_red[w];
            p.part[2 * blockIdx.x] = t;
            p.part[2 * blockIdx.x + 1] = t;
        }
    } else if (lane == 0) {
        // ===== MMA issuer
        const uint32_t idKK = make_idesc(kST, kST, 0, 0);         // S, P: both operands K-major, N = 128
        const uint32_t idMM = make_idesc(kST, kDim, 1, 1);        // dV = W'^T G, dK' = dS^T Q: A and B MN-major, N = 64
        const uint32_t idKM = make_idesc(kST, kDim, 0, 1);        // dQ' = dS K: A K-major, B MN-major, N = 64
        for (int it = 0; it < my_tiles; ++it) {
            mbar_wait(&in_full, it & 1);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t qhi = make_desc(Qop + ks * 32, kKmajLBO, kKmajSBO), qlo = make_desc(Qop + kSOp + ks * 32, kKmajLBO, kKmajSBO);
                const uint64_t khi = make_desc(Kop + ks * 32, kKmajLBO, kKmajSBO), klo = make_desc(Kop + kSOp + ks * 32, kKmajLBO, kKmajSBO);
                umma(tmS, qhi, khi, idKK, ks > 0 ? 1u : 0u);
                umma(tmS, qlo, khi, idKK, 1u);
                umma(tmS, qhi, klo, idKK, 1u);
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ghi = make_desc(Gop + ks * 32, kKmajLBO, kKmajSBO), glo = make_desc(Gop + kSOp + ks * 32, kKmajLBO, kKmajSBO);
                const uint64_t vhi = make_desc(Vop + ks * 32, kKmajLBO, kKmajSBO), vlo = make_desc(Vop + kSOp + ks * 32, kKmajLBO, kKmajSBO);
                umma(tmP, ghi, vhi, idKK, ks > 0 ? 1u : 0u);
                umma(tmP, glo, vhi, idKK, 1u);
                umma(tmP, ghi, vlo, idKK, 1u);
            }
            umma_commit(&sp_full);
            mbar_wait(&x1_full, it & 1);
            if (it > 0) mbar_wait(&e_free, (it - 1) & 1);          // the previous tile's dV | dQ | dK have been read
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {          // K index = tile row i: 16 rows of the buffer / of G per step
                const uint64_t xhi = make_desc(Xhi + ks * 2048, kSOp, 1024), xlo = make_desc(Xlo + ks * 2048, kSOp, 1024);
                const uint64_t ghi = make_desc(Gop + ks * 2048, kSOp, 1024), glo = make_desc(Gop + kSOp + ks * 2048, kSOp, 1024);
                umma(tmDV, xhi, ghi, idMM, ks > 0 ? 1u : 0u);
                umma(tmDV, xlo, ghi, idMM, 1u);
                umma(tmDV, xhi, glo, idMM, 1u);
            }
            umma_commit(&dv_full);
            mbar_wait(&x2_full, it & 1);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {          // dQ': K index = column j of dS (K-major A), rows j of K (MN-major B)
                const uint32_t wo = (uint32_t)((ks >> 2) * kSOp + (ks & 3) * 32);
                const uint64_t shi = make_desc(Xhi + wo, kKmajLBO, kKmajSBO), slo = make_desc(Xlo + wo, kKmajLBO, kKmajSBO);
                const uint64_t khi = make_desc(Kop + ks * 2048, kSOp, 1024), klo = make_desc(Kop + kSOp + ks * 2048, kSOp, 1024);
                umma(tmDQ, shi, khi, idKM, ks > 0 ? 1u : 0u);
                umma(tmDQ, slo, khi, idKM, 1u);
                umma(tmDQ, shi, klo, idKM, 1u);
            }
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {          // dK': K index = row i of dS (MN-major A), rows i of Q (MN-major B)
                const uint64_t shi = make_desc(Xhi + ks * 2048, kSOp, 1024), slo = make_desc(Xlo + ks * 2048, kSOp, 1024);
                const uint64_t qhi = make_desc(Qop + ks * 2048, kSOp, 1024), qlo = make_desc(Qop + kSOp + ks * 2048, kSOp, 1024);
                umma(tmDK, shi, qhi, idMM, ks > 0 ? 1u : 0u);
                umma(tmDK, slo, qhi, idMM, 1u);
                umma(tmDK, shi, qlo, idMM, 1u);
            }
            umma_commit(&c_full);
            SEG_STAMP(7);
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem, 512);
}

// t = sum of the per-CTA shares (fixed order) -> scal[0] = scal[1]
__global__ void seg_bwd_tc_sum_kernel(const float* __restrict__ part, int n, float* __restrict__ scal) {
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += (double)part[2 * i];
        scal[0] = scal[1] = (float)t;
    }
}

// dq -= q t / |Q|^2, dk -= k t / |K|^2 over all rows.  npart > 0: t is still spread over the CTAs' shares (`part`, single-GPU call): every
// block adds them up itself in the same fixed order and block 0 publishes the total; npart == 0: t has been reduced (and all-reduced) into scal.
__global__ void __launch_bounds__(256) seg_fixup_rows_kernel(const float4* __restrict__ q, const float4* __restrict__ k, float4* __restrict__ dq,
                                                             float4* __restrict__ dk, int64_t count, float* __restrict__ scal,
                                                             const float* __restrict__ part, int npart, const float* __restrict__ norms) {
    __shared__ float tsh;
    float tq = 0.f, tk = 0.f;
    if (npart > 0) {
        if (threadIdx.x < 32) {
            double t = 0.0;
            for (int i = threadIdx.x; i < npart; i += 32) t += (double)part[2 * i];
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (threadIdx.x == 0) {
                tsh = (float)t;
                if (blockIdx.x == 0) scal[0] = scal[1] = (float)t;
            }
        }
        __syncthreads();
        tq = tk = tsh;
    } else {
        tq = scal[0];
        tk = scal[1];
    }
    const float aq = tq / norms[0], ak = tk / norms[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 q4 = q[i], k4 = k[i];
        float4 a = dq[i], b = dk[i];
        a.x = fmaf(-aq, q4.x, a.x); a.y = fmaf(-aq, q4.y, a.y); a.z = fmaf(-aq, q4.z, a.z); a.w = fmaf(-aq, q4.w, a.w);
        b.x = fmaf(-ak, k4.x, b.x); b.y = fmaf(-ak, k4.y, b.y); b.z = fmaf(-ak, k4.z, b.z); b.w = fmaf(-ak, k4.w, b.w);
        dq[i] = a;
        dk[i] = b;
    }
}

}  // namespace

// plan = tile_row0 [ntiles + 1] (int32, padded to 16 bytes) | row_range [N] (int2)
static int seg_plan_layout(int64_t N, int max_nodes, int* S, int* ntiles, int64_t* off_rows) {
    if (N < 1 || N >= (1ll << 31) || max_nodes < 1 || max_nodes > kST) return 1;
    *S = kST + 1 - max_nodes;
    *ntiles = (int)((N + *S - 1) / *S);
    *off_rows = (((int64_t)(*ntiles + 1) * 4 + 15) / 16) * 16;
    return 0;
}

int64_t segmented_plan_bytes(int64_t N, int max_nodes) {
    int S, nt;
    int64_t off;
    if (seg_plan_layout(N, max_nodes, &S, &nt, &off)) return 0;
    return off + N * 8;
}

int segmented_plan_build(const int32_t* seg_ptr, int B, int64_t N, int max_nodes, void* plan, cudaStream_t st) {
    int S, nt;
    int64_t off;
    DIF_REQUIRE(!seg_plan_layout(N, max_nodes, &S, &nt, &off), DIF_EUNSUPPORTED, "segmented plan: needs 1 <= max_nodes <= 128 and N < 2^31");
    seg_plan_kernel<<<(B + 255) / 256, 256, 0, st>>>(seg_ptr, B, N, S, nt, (int*)plan, (int2*)((uint8_t*)plan + off));
    DIF_LAUNCH_OK();
    return DIF_OK;
}

int segmented_fwd_tc(const float* q, const float* k, const float* v, const void* plan, int64_t N, int max_nodes, const float* norms, float* out,
                     cudaStream_t st) {
    int S, nt;
    int64_t off;
    DIF_REQUIRE(!seg_plan_layout(N, max_nodes, &S, &nt, &off), DIF_EUNSUPPORTED, "segmented_fwd(tcgen05): needs 1 <= max_nodes <= 128 and N < 2^31");
    DIF_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) & 31) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)plan & 15) == 0, DIF_EARG,
                "segmented_fwd(tcgen05): q / k / v must be 32-byte, out and plan 16-byte aligned");
    static bool attr_set = false;
    if (!attr_set) {
        DIF_CUDA_OK(cudaFuncSetAttribute(seg_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSegTcSmem));
        attr_set = true;
    }
    int dev = 0, sms = 148;
    DIF_CUDA_OK(cudaGetDevice(&dev));
    DIF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static unsigned long long* dbg = nullptr;
    static const bool debug = getenv("DIF_SEG_DEBUG") && atoi(getenv("DIF_SEG_DEBUG"));
    if (debug && !dbg) DIF_CUDA_OK(cudaMalloc(&dbg, 64 * sizeof(unsigned long long)));
    if (debug) DIF_CUDA_OK(cudaMemsetAsync(dbg, 0, 64 * sizeof(unsigned long long), st));
    SegTcArgs a{q, k, v, (const int2*)((const uint8_t*)plan + off), (const int*)plan, nt, norms, out, debug ? dbg : nullptr};
    CUtensorMap omap;
    int rc = make_out_map(&omap, out, N, kDim);
    if (rc) return rc;
    seg_fwd_tc_kernel<<<nt < sms ? nt : sms, kSegTcThreads, kSegTcSmem, st>>>(a, omap);
    DIF_LAUNCH_OK();
    if (debug) {   // CTA 0: events per tile in ns since its first stamp: qk published, v published, scores seen, W published, O seen, epilogue done, MMA1 issued, MMA2 issued
        unsigned long long h[64];
        DIF_CUDA_OK(cudaStreamSynchronize(st));
        DIF_CUDA_OK(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull;
        for (int i = 0; i < 64; ++i) if (h[i] && h[i] < t0) t0 = h[i];
        fprintf(stderr, "[seg_fwd_tc] tile: qk_full v_full | s_seen w_full o_seen epi_done | mma1_issued mma2_issued (ns)\n");
        for (int t = 0; t < 8; ++t) {
            fprintf(stderr, "  %d:", t);
            for (int e = 0; e < 8; ++e) fprintf(stderr, " %7lld", h[t * 8 + e] ? (long long)(h[t * 8 + e] - t0) : -1ll);
            fprintf(stderr, "\n");
        }
    }
    return DIF_OK;
}

// phase 0: everything; phase 1: up to the batch-wide scalar t (scal[0] = scal[1], this rank's graphs) ; phase 2: the t terms only.
// `part`: 2 floats per CTA, `scal`: 2 floats.
int segmented_bwd_tc(const float* q, const float* k, const float* v, const float* g, const float* out, const void* plan, int64_t N, int max_nodes,
                     const float* norms, float* dq, float* dk, float* dv, float* part, float* scal, int phase, cudaStream_t st) {
    int S, nt;
    int64_t off;
    DIF_REQUIRE(!seg_plan_layout(N, max_nodes, &S, &nt, &off), DIF_EUNSUPPORTED, "segmented_bwd(tcgen05): needs 1 <= max_nodes <= 128 and N < 2^31");
    DIF_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)g) & 31) == 0 &&
                (((uintptr_t)out | (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv | (uintptr_t)plan) & 15) == 0, DIF_EARG,
                "segmented_bwd(tcgen05): q / k / v / g must be 32-byte, out / dq / dk / dv / plan 16-byte aligned");
    int dev = 0, sms = 148;
    DIF_CUDA_OK(cudaGetDevice(&dev));
    DIF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = nt < sms ? nt : sms;
    if (phase != 2) {
        static bool attr_set = false;
        if (!attr_set) {
            DIF_CUDA_OK(cudaFuncSetAttribute(seg_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSegBwdSmem));
            attr_set = true;
        }
        static unsigned long long* dbg = nullptr;
        static const bool debug = getenv("DIF_SEG_DEBUG") && atoi(getenv("DIF_SEG_DEBUG"));
        if (debug && !dbg) DIF_CUDA_OK(cudaMalloc(&dbg, 64 * sizeof(unsigned long long)));
        if (debug) DIF_CUDA_OK(cudaMemsetAsync(dbg, 0, 64 * sizeof(unsigned long long), st));
        SegBwdTcArgs a{q, k, v, g, out, (const int2*)((const uint8_t*)plan + off), (const int*)plan, nt, norms, dq, dk, dv, part, debug ? dbg : nullptr};
        CUtensorMap mv, mq, mk;
        int rc = make_out_map(&mv, dv, N, kDim);
        if (!rc) rc = make_out_map(&mq, dq, N, kDim);
        if (!rc) rc = make_out_map(&mk, dk, N, kDim);
        if (rc) return rc;
        seg_bwd_tc_kernel<<<grid, kSegTcThreads, kSegBwdSmem, st>>>(a, mv, mq, mk);
        DIF_LAUNCH_OK();
        if (debug) {
            unsigned long long h[64];
            DIF_CUDA_OK(cudaStreamSynchronize(st));
            DIF_CUDA_OK(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
            unsigned long long t0 = ~0ull;
            for (int i = 0; i < 64; ++i) if (h[i] && h[i] < t0) t0 = h[i];
            fprintf(stderr, "[seg_bwd_tc] tile: in_full | sp_seen x1_full dv_seen x2_full c_seen epi_done | mmaC_issued (ns)\n");
            for (int t = 0; t < 8; ++t) {
                fprintf(stderr, "  %d:", t);
                for (int e = 0; e < 8; ++e) fprintf(stderr, " %7lld", h[t * 8 + e] ? (long long)(h[t * 8 + e] - t0) : -1ll);
                fprintf(stderr, "\n");
            }
        }
        if (phase == 1) {            // sharded graphs: the caller all-reduces scal before phase 2
            seg_bwd_tc_sum_kernel<<<1, 32, 0, st>>>(part, grid, scal);
            DIF_LAUNCH_OK();
        }
    }
    if (phase == 1) return DIF_OK;
    seg_fixup_rows_kernel<<<sms * 8, 256, 0, st>>>((const float4*)q, (const float4*)k, (float4*)dq, (float4*)dk, N * (kDim / 4), scal, part,
                                                   phase == 0 ? grid : 0, norms);
    DIF_LAUNCH_OK();
    return DIF_OK;
}

}  // namespace dif
