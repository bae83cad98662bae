// Batched-graph 'simple' attention -- TransConv.full_attention, physical particle/difformer-v2.py:80-111.
//
// The reference pads every graph to [B, maxN, H, D] with Python loops over the B graphs
// (make_batch_mask / make_batch, :8-20; 164 ms per call at B = 8192) and runs the einsums on the
// padded tensors.  Semantics: per graph g with n_g rows,
//     out_n = (q^_n S^_g + u_g) / (q^_n z^_g + n_g)         (:96-109)
// with S_g, z_g, u_g summed over the rows of g only, but ||Q||_F, ||K||_F taken over the whole batch
// (:82-83).  Here: no padding and no host loop -- seg_ptr[B+1] on device, one CTA per graph, which
// builds S_g in registers/shared memory and applies it to its own rows straight away.
//
// Backward = per-graph a-1b with two global scalars (t_q, t_k):
//   phase A: per graph t_q,g and t_k,g -> summed in fixed order (deterministic)
//   phase B: per graph recompute S_g, dS_g and emit dq, dk, dv.
#include <algorithm>

#include "common.cuh"
#include "tile.cuh"

namespace dif {
namespace {

constexpr int kSegRows = 32;

struct SegArgs {
    const float *q, *k, *v, *g, *out;
    const int32_t* seg;
    const float* norms;   // [sum q^2, sum k^2]
    const float* scal;    // bwd phase B: [t_q, t_k]
    int64_t N;
    int B, H, Hv, M, D;
    float *o, *dq, *dk, *dv;
    float* part;          // bwd phase A: [B][2]
    int min_rows;         // fwd CTA kernel: only graphs with more rows than this (0 = all)
};

__device__ __forceinline__ void seg_load(float* __restrict__ dst, int ld, const float* __restrict__ src, int64_t row0, int nr,
                                         int heads, int head, int W) {
    const int w4 = W >> 2;
    for (int idx = threadIdx.x; idx < kSegRows * w4; idx += kThreads) {
        const int r = idx / w4, c4 = idx - r * w4;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nr) x = ldg4(src + ((row0 + r) * heads + head) * W + 4 * c4);
        *reinterpret_cast<float4*>(dst + r * ld + 4 * c4) = x;
    }
}

// S[M][D] (+ z[M], u[D]) of rows [s,e) of head h into shared memory (raw, un-normalised).
// `wrow` (optional, shared) weights the z sum per row; A rows come from `a`, B rows from sb tiles
// already prepared by the caller when PREP (backward: B = dnum).
template <int TPT>
__device__ __forceinline__ void seg_store_tiles(float (&acc)[TPT][4][4], float* __restrict__ Sd, int D, int tilesD, int ntile, float scale) {
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
        const int tile = threadIdx.x + t * kThreads;
        if (tile < ntile) {
            const int mi = tile / tilesD, di = tile - mi * tilesD;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(Sd + (4 * mi + i) * D + 4 * di) =
                    make_float4(acc[t][i][0] * scale, acc[t][i][1] * scale, acc[t][i][2] * scale, acc[t][i][3] * scale);
        }
    }
}

template <int TPT>
__device__ __forceinline__ void zero_acc(float (&acc)[TPT][4][4]) {
#pragma unroll
    for (int t = 0; t < TPT; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.f;
}

// forward reduce of one graph/head: Ws = c*S_g, zs = c*z_g, us = u_g
template <int TPT>
__device__ __forceinline__ void seg_reduce_kv(const SegArgs& p, int64_t s, int64_t e, int h, int hv, float c,
                                              float* Ws, float* zs, float* us, float* ta, float* tb) {
    const int M = p.M, D = p.D, lda = M + 4, ldb = D + 4;
    const int tid = threadIdx.x, tilesD = D >> 2, ntile = (M >> 2) * tilesD;
    float acc[TPT][4][4];
    zero_acc<TPT>(acc);
    float zacc = 0.f, uacc = 0.f;
    for (int64_t r0 = s; r0 < e; r0 += kSegRows) {
        const int nr = (int)min((int64_t)kSegRows, e - r0);
        seg_load(ta, lda, p.k, r0, nr, p.H, h, M);
        seg_load(tb, ldb, p.v, r0, nr, p.Hv, hv, D);
        __syncthreads();
#pragma unroll
        for (int t = 0; t < TPT; ++t) {
            const int tile = tid + t * kThreads;
            if (tile < ntile) tile_atb_acc(ta, lda, tb, ldb, kSegRows, tile / tilesD, tile % tilesD, acc[t]);
        }
        if (tid < M) { for (int r = 0; r < kSegRows; ++r) zacc += ta[r * lda + tid]; }
        else if (tid < M + D) { for (int r = 0; r < kSegRows; ++r) uacc += tb[r * ldb + tid - M]; }
        __syncthreads();
    }
    seg_store_tiles<TPT>(acc, Ws, D, tilesD, ntile, c);
    if (tid < M) zs[tid] = zacc * c;
    else if (tid < M + D) us[tid - M] = uacc;
    __syncthreads();
}

template <int TPT>
__global__ void __launch_bounds__(kThreads) seg_fwd_kernel(SegArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int M = p.M, D = p.D, H = p.H;
    const int lda = M + 4, ldb = D + 4;
    float* Ws = smem;                       // [M][D]
    float* zs = Ws + M * D;                 // [M]
    float* us = zs + M;                     // [D]
    float* ta = us + D;                     // [32][M+4]
    float* tb = ta + kSegRows * lda;        // [32][D+4]
    float* sden = tb + kSegRows * ldb;      // [32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float c = 1.f / (sqrtf(p.norms[0]) * sqrtf(p.norms[1]));
    const int tilesJ = D >> 2, ntileA = (kSegRows >> 2) * tilesJ;
    // min_rows > 0: the small graphs were taken by seg_fwd_warp_kernel and this kernel mostly skips.  Then every CTA
    // scans kThreads graphs at once (one coalesced load of their sizes) instead of walking them one by one.
    const bool scan = p.min_rows > 0;
    const int step = scan ? gridDim.x * kThreads : gridDim.x;
    for (int base = scan ? blockIdx.x * kThreads : blockIdx.x; base < p.B; base += step) {
    int g_end = base + 1;
    if (scan) {
        const int gi = base + tid;
        const bool big = gi < p.B && (p.seg[gi + 1] - p.seg[gi]) > p.min_rows;
        if (!__syncthreads_or(big)) continue;
        g_end = min(base + kThreads, p.B);
    }
    for (int g = base; g < g_end; ++g) {
        const int64_t s = p.seg[g], e = p.seg[g + 1];
        if (e - s <= p.min_rows) continue;          // small graphs are handled by seg_fwd_warp_kernel
        const float ng = (float)(e - s);
        for (int h = 0; h < H; ++h) {
            const int hv = (p.Hv == H) ? h : 0;
            seg_reduce_kv<TPT>(p, s, e, h, hv, c, Ws, zs, us, ta, tb);
            for (int64_t r0 = s; r0 < e; r0 += kSegRows) {
                const int nr = (int)min((int64_t)kSegRows, e - r0);
                seg_load(ta, lda, p.q, r0, nr, H, h, M);
                __syncthreads();
                for (int r = warp; r < kSegRows; r += kThreads / 32) {
                    float qz = 0.f;
                    for (int i = lane; i < M; i += 32) qz = fmaf(ta[r * lda + i], zs[i], qz);
                    qz = warp_sum(qz);
                    if (lane == 0) sden[r] = qz + ng;
                }
                __syncthreads();
                for (int tile = tid; tile < ntileA; tile += kThreads) {
                    const int ri = tile / tilesJ, ji = tile - ri * tilesJ;
                    float acc[4][4];
                    tile_mm(ta, lda, Ws, D, M, ri, ji, acc);
                    const float4 u4 = *reinterpret_cast<const float4*>(us + 4 * ji);
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int r = 4 * ri + a;
                        if (r >= nr) continue;
                        const float den = sden[r];
                        *reinterpret_cast<float4*>(p.o + ((r0 + r) * H + h) * D + 4 * ji) =
                            make_float4((acc[a][0] + u4.x) / den, (acc[a][1] + u4.y) / den, (acc[a][2] + u4.z) / den, (acc[a][3] + u4.w) / den);
                    }
                }
                __syncthreads();
            }
        }
    }
    }
}

// ---- small graphs (n_g <= kWarpMaxRows), M == D == 64: one WARP per graph, direct O(n^2) form
//      out_n = sum_l (1 + c q_n.k_l) v_l / sum_l (1 + c q_n.k_l)        (same value as (c q S + u)/(c q z + n))
// which costs 2 n^2 64 MACs per graph instead of 2 n 64^2 -- cheaper for n < 64 -- with no block-level
// synchronisation at all: two lanes per query row, K/V rows of the graph broadcast from a per-warp shared buffer.
constexpr int kWarpMaxRows = 64, kWarpsPerCta = 4, kWarpStage = 16;
constexpr int kWarpBufFloats = 2 * kWarpStage * 64;      // one staging buffer: K rows | V rows

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

__global__ void __launch_bounds__(kWarpsPerCta * 32, 3) seg_fwd_warp_kernel(SegArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* bufs = smem + warp * (2 * kWarpBufFloats);     // two staging buffers per warp
    const int H = p.H;
    const float c = 1.f / (sqrtf(p.norms[0]) * sqrtf(p.norms[1]));
    const int gw = blockIdx.x * kWarpsPerCta + warp, nw = gridDim.x * kWarpsPerCta;
    for (int g = gw; g < p.B; g += nw) {
        const int64_t s = p.seg[g], e = p.seg[g + 1];
        const int n = (int)(e - s);
        if (n <= 0 || n > kWarpMaxRows) continue;
        for (int h = 0; h < H; ++h) {
            const int hv = (p.Hv == H) ? h : 0;
            // lane = (row pair rp = lane >> 1, column half = lane & 1): 2 query rows x 32 columns per lane, 32 rows per
            // pass.  Every K/V value read from shared memory feeds two rows (halves the LDS count), and the two column
            // halves of a staged row are interleaved in 16-byte chunks so that the two addresses a quarter-warp reads
            // fall into different banks (a 128-bit broadcast load costs one wavefront per quarter-warp, not two).
            const int half = lane & 1, rp = lane >> 1;
            // K/V rows are staged 16 at a time with cp.async into a double buffer: the copy of stage t+1 is in flight
            // while stage t is consumed, with a fixed number of (predicated) copies per lane -- the first version
            // staged through registers with a data-dependent trip count and spent half its time in exposed
            // global-load latency (profiles/r1_segmented.md).
            auto stage_load = [&](int buf, int l0) {
                const int nl = min(kWarpStage, n - l0);
                float* Kb = bufs + buf * kWarpBufFloats;
                float* Vb = Kb + kWarpStage * 64;
#pragma unroll
                for (int it = 0; it < (kWarpStage * 16) / 32; ++it) {
                    const int idx = lane + 32 * it, r = idx >> 4, c4 = idx & 15;
                    if (r < nl) {
                        // chunk c4 (columns 4 c4 ..) of half (c4 >> 3) goes to interleaved slot 2 (c4 & 7) + (c4 >> 3)
                        const int slot = 2 * (c4 & 7) + (c4 >> 3);
                        cp_async16(Kb + r * 64 + 4 * slot, p.k + ((s + l0 + r) * H + h) * 64 + 4 * c4);
                        cp_async16(Vb + r * 64 + 4 * slot, p.v + ((s + l0 + r) * p.Hv + hv) * 64 + 4 * c4);
                    }
                }
                cp_async_commit();
            };
            const int S = (n + kWarpStage - 1) / kWarpStage, P = (n + 31) >> 5, total = S * P;
            __syncwarp();                                   // the previous graph / head has finished with the buffers
            stage_load(0, 0);
            float4 qa[8], qb[8], aa[8], ab[8];
            float dena = 0.f, denb = 0.f;
            int st = 0, q0 = 0;
            for (int t = 0; t < total; ++t) {
                const int qr0 = q0 + 2 * rp, qr1 = qr0 + 1;
                if (st == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        qa[i] = qr0 < n ? ldg4(p.q + ((s + qr0) * H + h) * 64 + 32 * half + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                        qb[i] = qr1 < n ? ldg4(p.q + ((s + qr1) * H + h) * 64 + 32 * half + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                        aa[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    dena = 0.f;
                    denb = 0.f;
                }
                if (t + 1 < total) stage_load((t + 1) & 1, (st + 1 == S ? 0 : st + 1) * kWarpStage);
                else cp_async_commit();                     // empty group: keeps the wait below uniform
                cp_async_wait1();                           // everything but the newest group has landed => stage t
                __syncwarp();
                const float* Ks = bufs + (t & 1) * kWarpBufFloats;
                const float* Vs = Ks + kWarpStage * 64;
                const int nl = min(kWarpStage, n - st * kWarpStage);
                for (int l = 0; l < nl; ++l) {
                    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;      // independent chains (FMA latency)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 k4 = *reinterpret_cast<const float4*>(Ks + l * 64 + 4 * (2 * i + half));
                        a0 = fmaf(qa[i].x, k4.x, a0); a1 = fmaf(qa[i].y, k4.y, a1);
                        a0 = fmaf(qa[i].z, k4.z, a0); a1 = fmaf(qa[i].w, k4.w, a1);
                        b0 = fmaf(qb[i].x, k4.x, b0); b1 = fmaf(qb[i].y, k4.y, b1);
                        b0 = fmaf(qb[i].z, k4.z, b0); b1 = fmaf(qb[i].w, k4.w, b1);
                    }
                    float sa = a0 + a1, sb = b0 + b1;
                    sa += __shfl_xor_sync(0xffffffffu, sa, 1);            // the other column half of the rows
                    sb += __shfl_xor_sync(0xffffffffu, sb, 1);
                    const float wa = fmaf(c, sa, 1.f), wb = fmaf(c, sb, 1.f);
                    dena += wa;
                    denb += wb;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 v4 = *reinterpret_cast<const float4*>(Vs + l * 64 + 4 * (2 * i + half));
                        aa[i].x = fmaf(wa, v4.x, aa[i].x); aa[i].y = fmaf(wa, v4.y, aa[i].y);
                        aa[i].z = fmaf(wa, v4.z, aa[i].z); aa[i].w = fmaf(wa, v4.w, aa[i].w);
                        ab[i].x = fmaf(wb, v4.x, ab[i].x); ab[i].y = fmaf(wb, v4.y, ab[i].y);
                        ab[i].z = fmaf(wb, v4.z, ab[i].z); ab[i].w = fmaf(wb, v4.w, ab[i].w);
                    }
                }
                __syncwarp();                               // all lanes are done with this buffer (refilled at t + 2)
                if (++st == S) {
                    if (qr0 < n) {
                        float* o = p.o + ((s + qr0) * H + h) * 64 + 32 * half;
                        const float inv = 1.f / dena;
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            *reinterpret_cast<float4*>(o + 4 * i) = make_float4(aa[i].x * inv, aa[i].y * inv, aa[i].z * inv, aa[i].w * inv);
                    }
                    if (qr1 < n) {
                        float* o = p.o + ((s + qr1) * H + h) * 64 + 32 * half;
                        const float inv = 1.f / denb;
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            *reinterpret_cast<float4*>(o + 4 * i) = make_float4(ab[i].x * inv, ab[i].y * inv, ab[i].z * inv, ab[i].w * inv);
                    }
                    st = 0;
                    q0 += 32;
                }
            }
        }
    }
}

// ---- backward of the small graphs, same warp-per-graph direct form.  With w_nl = 1 + c q_n.k_l, d_n = sum_l w_nl,
//      out_n = sum_l w_nl v_l / d_n and D_n = g_n.out_n:
//          dw_nl = (g_n.v_l - D_n) / d_n
//          dq_n  = c sum_l dw_nl k_l  - q_n t / |Q|^2         dk_l = c sum_n dw_nl q_n  - k_l t / |K|^2
//          dv_l  = sum_n (w_nl / d_n) g_n                      t    = sum_{g,n,l} dw_nl c q_n.k_l   (c = 1/(|Q||K|) is global)
//      The warp kernel writes dq, dk without the t terms, dv, and its graph's share of t; seg_bwd_fixup_kernel subtracts
//      the t terms once t has been summed over all graphs (fixed order).
//      lane = (row = lane >> 1, column half = lane & 1): 16 rows per pass, 32 columns per lane.
template <class F>
__device__ __forceinline__ void seg_warp_pipeline(float* bufs, int lane, const float* a_src, int a_heads, int a_head, const float* b_src,
                                                  int b_heads, int b_head, int64_t s, int n, F&& body) {
    const int S = (n + kWarpStage - 1) / kWarpStage;
    auto load = [&](int buf, int l0) {
        const int nl = min(kWarpStage, n - l0);
        float* Ab = bufs + buf * kWarpBufFloats;
        float* Bb = Ab + kWarpStage * 64;
#pragma unroll
        for (int it = 0; it < (kWarpStage * 16) / 32; ++it) {
            const int idx = lane + 32 * it, r = idx >> 4, c4 = idx & 15;
            if (r < nl) {
                const int slot = 2 * (c4 & 7) + (c4 >> 3);       // column halves interleaved in 16-byte chunks (see forward)
                cp_async16(Ab + r * 64 + 4 * slot, a_src + ((s + l0 + r) * a_heads + a_head) * 64 + 4 * c4);
                cp_async16(Bb + r * 64 + 4 * slot, b_src + ((s + l0 + r) * b_heads + b_head) * 64 + 4 * c4);
            }
        }
        cp_async_commit();
    };
    __syncwarp();                                   // everybody is done with the buffers
    load(0, 0);
    for (int t = 0; t < S; ++t) {
        if (t + 1 < S) load((t + 1) & 1, (t + 1) * kWarpStage);
        else cp_async_commit();
        cp_async_wait1();
        __syncwarp();
        const float* A = bufs + (t & 1) * kWarpBufFloats;
        const float* Bm = A + kWarpStage * 64;
        const int nl = min(kWarpStage, n - t * kWarpStage);
        for (int l = 0; l < nl; ++l) body(A + l * 64, Bm + l * 64, t * kWarpStage + l);
        __syncwarp();
    }
}

__device__ __forceinline__ float dot32(const float4 (&a)[8], const float4 (&b)[8]) {
    float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x0 = fmaf(a[i].x, b[i].x, x0); x1 = fmaf(a[i].y, b[i].y, x1);
        x2 = fmaf(a[i].z, b[i].z, x2); x3 = fmaf(a[i].w, b[i].w, x3);
    }
    return (x0 + x1) + (x2 + x3);
}
__device__ __forceinline__ void lds_row(const float* row, int half, float4 (&x)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = *reinterpret_cast<const float4*>(row + 4 * (2 * i + half));
}
__device__ __forceinline__ void ldg_row(const float* row32, bool live, float4 (&x)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = live ? ldg4(row32 + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void axpy32(float a, const float4 (&x)[8], float4 (&y)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        y[i].x = fmaf(a, x[i].x, y[i].x); y[i].y = fmaf(a, x[i].y, y[i].y);
        y[i].z = fmaf(a, x[i].z, y[i].z); y[i].w = fmaf(a, x[i].w, y[i].w);
    }
}

__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) seg_bwd_warp_kernel(SegArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* bufs = smem + warp * (2 * kWarpBufFloats + 2 * kWarpMaxRows);
    float* rowD = bufs + 2 * kWarpBufFloats;              // [64] D_n = g_n . out_n
    float* rowI = rowD + kWarpMaxRows;                    // [64] 1 / d_n
    const int H = p.H;
    const bool bcast = (p.Hv != H);
    const float c = 1.f / (sqrtf(p.norms[0]) * sqrtf(p.norms[1]));
    const int half = lane & 1, rr = lane >> 1;
    const int gw = blockIdx.x * kWarpsPerCta + warp, nw = gridDim.x * kWarpsPerCta;
    for (int g = gw; g < p.B; g += nw) {
        const int64_t s = p.seg[g], e = p.seg[g + 1];
        const int n = (int)(e - s);
        if (n > kWarpMaxRows) continue;
        if (n <= 0) {                                       // empty graph: its share of t is zero
            if (lane == 0) { p.part[2 * (int64_t)g] = 0.f; p.part[2 * (int64_t)g + 1] = 0.f; }
            continue;
        }
        float tacc = 0.f;
        for (int h = 0; h < H; ++h) {
            const int hv = bcast ? 0 : h;
            // ---- z = sum_l k_l (this lane's 32 columns)
            float4 z[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            seg_warp_pipeline(bufs, lane, p.k, H, h, p.v, p.Hv, hv, s, n, [&](const float* kr, const float*, int) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 k4 = *reinterpret_cast<const float4*>(kr + 4 * (2 * i + half));
                    z[i].x += k4.x; z[i].y += k4.y; z[i].z += k4.z; z[i].w += k4.w;
                }
            });
            // ---- phase A: lanes own query rows; loop over the keys: dq, t, and the per-row scalars D_n, 1/d_n
            for (int q0 = 0; q0 < n; q0 += 16) {
                const int r = q0 + rr;
                const bool live = r < n;
                float4 qv[8], gv[8], acc[8];
                ldg_row(p.q + ((s + r) * H + h) * 64 + 32 * half, live, qv);
                ldg_row(p.g + ((s + r) * H + h) * 64 + 32 * half, live, gv);
                ldg_row(p.out + ((s + r) * H + h) * 64 + 32 * half, live, acc);      // out row, only for D_n
                float Dn = dot32(gv, acc), qz = dot32(qv, z);
                Dn += __shfl_xor_sync(0xffffffffu, Dn, 1);
                qz += __shfl_xor_sync(0xffffffffu, qz, 1);
                const float invd = 1.f / fmaf(c, qz, (float)n);
                if (live && half == 0) { rowD[r] = Dn; rowI[r] = invd; }
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                seg_warp_pipeline(bufs, lane, p.k, H, h, p.v, p.Hv, hv, s, n, [&](const float* kr, const float* vr, int) {
                    float4 k4[8], v4[8];
                    lds_row(kr, half, k4);
                    lds_row(vr, half, v4);
                    float sc = dot32(qv, k4), tt = dot32(gv, v4);
                    sc += __shfl_xor_sync(0xffffffffu, sc, 1);
                    tt += __shfl_xor_sync(0xffffffffu, tt, 1);
                    const float dw = (tt - Dn) * invd;
                    axpy32(dw, k4, acc);
                    if (live && half == 0) tacc = fmaf(dw * c, sc, tacc);
                });
                if (live) {
                    float* o = p.dq + ((s + r) * H + h) * 64 + 32 * half;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *reinterpret_cast<float4*>(o + 4 * i) = make_float4(acc[i].x * c, acc[i].y * c, acc[i].z * c, acc[i].w * c);
                }
            }
            // ---- phase B: lanes own key rows; loop over the queries: dk, dv
            for (int l0 = 0; l0 < n; l0 += 16) {
                const int r = l0 + rr;
                const bool live = r < n;
                float4 kv[8], vv[8], dka[8], dva[8];
                ldg_row(p.k + ((s + r) * H + h) * 64 + 32 * half, live, kv);
                ldg_row(p.v + ((s + r) * p.Hv + hv) * 64 + 32 * half, live, vv);
#pragma unroll
                for (int i = 0; i < 8; ++i) { dka[i] = make_float4(0.f, 0.f, 0.f, 0.f); dva[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
                seg_warp_pipeline(bufs, lane, p.q, H, h, p.g, H, h, s, n, [&](const float* qr, const float* gr, int nn) {
                    float4 q4[8], g4[8];
                    lds_row(qr, half, q4);
                    lds_row(gr, half, g4);
                    float sc = dot32(q4, kv), tt = dot32(g4, vv);
                    sc += __shfl_xor_sync(0xffffffffu, sc, 1);
                    tt += __shfl_xor_sync(0xffffffffu, tt, 1);
                    const float invd = rowI[nn];
                    const float dw = (tt - rowD[nn]) * invd;
                    axpy32(dw, q4, dka);
                    axpy32(fmaf(c, sc, 1.f) * invd, g4, dva);
                });
                if (live) {
                    float* ok = p.dk + ((s + r) * H + h) * 64 + 32 * half;
                    float* ov = p.dv + ((s + r) * p.Hv + hv) * 64 + 32 * half;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        *reinterpret_cast<float4*>(ok + 4 * i) = make_float4(dka[i].x * c, dka[i].y * c, dka[i].z * c, dka[i].w * c);
                        float4 o = dva[i];
                        if (bcast && h > 0) {     // V shared by all heads: accumulate (same lane, same address, fixed order)
                            const float4 prev = *reinterpret_cast<const float4*>(ov + 4 * i);
                            o.x += prev.x; o.y += prev.y; o.z += prev.z; o.w += prev.w;
                        }
                        *reinterpret_cast<float4*>(ov + 4 * i) = o;
                    }
                }
            }
        }
        tacc = warp_sum(tacc);
        if (lane == 0) { p.part[2 * (int64_t)g] = tacc; p.part[2 * (int64_t)g + 1] = tacc; }
    }
}

// dq -= q t_q/|Q|^2, dk -= k t_k/|K|^2 on the rows of the small graphs (one warp per graph, after t has been summed)
__global__ void __launch_bounds__(256) seg_bwd_fixup_kernel(SegArgs p) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const float aq = p.scal[0] / p.norms[0], ak = p.scal[1] / p.norms[1];
    for (int g = gw; g < p.B; g += nw) {
        const int64_t s = p.seg[g], e = p.seg[g + 1];
        const int n = (int)(e - s);
        if (n <= 0 || n > kWarpMaxRows) continue;
        const int64_t base = s * p.H * 16, count = (int64_t)n * p.H * 16;      // float4 elements (M = 64)
        for (int64_t i = lane; i < count; i += 32) {
            const float4 q4 = ldg4(p.q + 4 * (base + i)), k4 = ldg4(p.k + 4 * (base + i));
            float4* dq = reinterpret_cast<float4*>(p.dq) + base + i;
            float4* dk = reinterpret_cast<float4*>(p.dk) + base + i;
            float4 a = *dq, b = *dk;
            a.x = fmaf(-aq, q4.x, a.x); a.y = fmaf(-aq, q4.y, a.y); a.z = fmaf(-aq, q4.z, a.z); a.w = fmaf(-aq, q4.w, a.w);
            b.x = fmaf(-ak, k4.x, b.x); b.y = fmaf(-ak, k4.y, b.y); b.z = fmaf(-ak, k4.z, b.z); b.w = fmaf(-ak, k4.w, b.w);
            *dq = a;
            *dk = b;
        }
    }
}

// backward reduce of one graph/head given Ws=c*S, zs=c*z, us=u: dS (raw, scaled by `scale` into
// dWs), dzs = scale*dz, dus = du; returns this thread's t_q contribution and, through `tk`, its
// share of sum S o dS + z . dz (both raw*c products).
template <int TPT>
__device__ __forceinline__ void seg_reduce_bwd(const SegArgs& p, int64_t s, int64_t e, int h, float ng, float scale,
                                               const float* Ws, const float* zs, const float* us,
                                               float* dWs, float* dzs, float* dus,
                                               float* tq_, float* tg, float* to, float* sw, float& tq, float& tk) {
    const int M = p.M, D = p.D, H = p.H, lda = M + 4, ldb = D + 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tilesD = D >> 2, ntile = (M >> 2) * tilesD;
    float acc[TPT][4][4];
    zero_acc<TPT>(acc);
    float zacc = 0.f, uacc = 0.f;
    for (int64_t r0 = s; r0 < e; r0 += kSegRows) {
        const int nr = (int)min((int64_t)kSegRows, e - r0);
        seg_load(tq_, lda, p.q, r0, nr, H, h, M);
        seg_load(tg, ldb, p.g, r0, nr, H, h, D);
        seg_load(to, ldb, p.out, r0, nr, H, h, D);
        __syncthreads();
        for (int r = warp; r < kSegRows; r += kThreads / 32) {
            float qz = 0.f, go = 0.f, gu = 0.f;
            for (int i = lane; i < M; i += 32) qz = fmaf(tq_[r * lda + i], zs[i], qz);
            for (int i = lane; i < D; i += 32) {
                const float gg = tg[r * ldb + i];
                go = fmaf(gg, to[r * ldb + i], go);
                gu = fmaf(gg, us[i], gu);
            }
            qz = warp_sum(qz); go = warp_sum(go); gu = warp_sum(gu);
            const float inv = 1.f / (qz + ng);
            const float dden = -go * inv;
            for (int i = lane; i < D; i += 32) tg[r * ldb + i] *= inv;
            if (lane == 0) { sw[r] = dden; if (r < nr) tq += go - inv * gu + dden * qz; }
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < TPT; ++t) {
            const int tile = tid + t * kThreads;
            if (tile < ntile) tile_atb_acc(tq_, lda, tg, ldb, kSegRows, tile / tilesD, tile % tilesD, acc[t]);
        }
        if (tid < M) { for (int r = 0; r < kSegRows; ++r) zacc = fmaf(tq_[r * lda + tid], sw[r], zacc); }
        else if (tid < M + D) { for (int r = 0; r < kSegRows; ++r) uacc += tg[r * ldb + tid - M]; }
        __syncthreads();
    }
    // t_k share: (c S) o dS_raw  + (c z) . dz_raw
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
        const int tile = tid + t * kThreads;
        if (tile < ntile) {
            const int mi = tile / tilesD, di = tile - mi * tilesD;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) tk = fmaf(Ws[(4 * mi + i) * D + 4 * di + j], acc[t][i][j], tk);
        }
    }
    if (tid < M) tk = fmaf(zs[tid], zacc, tk);
    if (dWs) {
        seg_store_tiles<TPT>(acc, dWs, D, tilesD, ntile, scale);
        if (tid < M) dzs[tid] = zacc * scale;
        else if (tid < M + D) dus[tid - M] = uacc;
    }
    __syncthreads();
}

template <int TPT>
__global__ void __launch_bounds__(kThreads) seg_bwd_scalars_kernel(SegArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int M = p.M, D = p.D, H = p.H, lda = M + 4, ldb = D + 4;
    float* Ws = smem;
    float* zs = Ws + M * D;
    float* us = zs + M;
    float* ta = us + D;
    float* tb = ta + kSegRows * lda;
    float* tc = tb + kSegRows * ldb;
    float* sw = tc + kSegRows * ldb;
    float* red = sw + kSegRows;
    const float c = 1.f / (sqrtf(p.norms[0]) * sqrtf(p.norms[1]));
    for (int g = blockIdx.x; g < p.B; g += gridDim.x) {
        const int64_t s = p.seg[g], e = p.seg[g + 1];
        if (p.min_rows > 0 && e - s <= p.min_rows) continue;     // small graphs: seg_bwd_warp_kernel wrote their share
        float tq = 0.f, tk = 0.f;
        if (e > s) {
            for (int h = 0; h < H; ++h) {
                const int hv = (p.Hv == H) ? h : 0;
                seg_reduce_kv<TPT>(p, s, e, h, hv, c, Ws, zs, us, ta, tb);
                seg_reduce_bwd<TPT>(p, s, e, h, (float)(e - s), 1.f, Ws, zs, us, nullptr, nullptr, nullptr, ta, tb, tc, sw, tq, tk);
            }
        }
        const float a = block_sum(tq, red);
        const float b = block_sum(tk, red);
        if (threadIdx.x == 0) { p.part[2 * (int64_t)g] = a; p.part[2 * (int64_t)g + 1] = b; }
    }
}

__global__ void seg_sum_pairs_kernel(const float* __restrict__ part, int64_t n, float* __restrict__ out2) {
    __shared__ float red[33];
    double a = 0.0, b = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { a += (double)part[2 * i]; b += (double)part[2 * i + 1]; }
    const float fa = block_sum((float)a, red);
    const float fb = block_sum((float)b, red);
    if (threadIdx.x == 0) { out2[0] = fa; out2[1] = fb; }
}

template <int TPT>
__global__ void __launch_bounds__(kThreads) seg_bwd_main_kernel(SegArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int M = p.M, D = p.D, H = p.H, lda = M + 4, ldb = D + 4;
    float* Ws = smem;                      // c*S_g
    float* dWs = Ws + M * D;               // c*dS_g
    float* zs = dWs + M * D;               // c*z
    float* us = zs + M;                    // u
    float* dzs = us + D;                   // c*dz
    float* dus = dzs + M;                  // du
    float* ta = dus + D;                   // [32][M+4]  q / k
    float* tb = ta + kSegRows * lda;       // [32][D+4]  g->dnum / v
    float* tc = tb + kSegRows * ldb;       // [32][D+4]  out
    float* sw = tc + kSegRows * ldb;       // [32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float sq = p.norms[0], sk = p.norms[1];
    const float c = 1.f / (sqrtf(sq) * sqrtf(sk));
    const float tq_over = p.scal[0] / sq, tk_over = p.scal[1] / sk;
    const int tilesM = M >> 2, ntM = (kSegRows >> 2) * tilesM;   // outputs with M columns (dq, dk)
    const int tilesD = D >> 2, ntD = (kSegRows >> 2) * tilesD;   // dv
    const int BS = M >> 2;                                        // interleave stride of tile_abt B rows
    const bool bcast = (p.Hv != H);
    for (int g = blockIdx.x; g < p.B; g += gridDim.x) {
        const int64_t s = p.seg[g], e = p.seg[g + 1];
        if (e <= s || e - s <= p.min_rows) continue;             // small graphs are handled by seg_bwd_warp_kernel
        const float ng = (float)(e - s);
        for (int h = 0; h < H; ++h) {
            const int hv = bcast ? 0 : h;
            float dummy_q = 0.f, dummy_k = 0.f;
            seg_reduce_kv<TPT>(p, s, e, h, hv, c, Ws, zs, us, ta, tb);
            seg_reduce_bwd<TPT>(p, s, e, h, ng, c, Ws, zs, us, dWs, dzs, dus, ta, tb, tc, sw, dummy_q, dummy_k);
            for (int64_t r0 = s; r0 < e; r0 += kSegRows) {
                const int nr = (int)min((int64_t)kSegRows, e - r0);
                // ---- dq = dnum (cS)^T + dden (cz) - q t_q/sq
                seg_load(ta, lda, p.q, r0, nr, H, h, M);
                seg_load(tb, ldb, p.g, r0, nr, H, h, D);
                seg_load(tc, ldb, p.out, r0, nr, H, h, D);
                __syncthreads();
                for (int r = warp; r < kSegRows; r += kThreads / 32) {
                    float qz = 0.f, go = 0.f;
                    for (int i = lane; i < M; i += 32) qz = fmaf(ta[r * lda + i], zs[i], qz);
                    for (int i = lane; i < D; i += 32) go = fmaf(tb[r * ldb + i], tc[r * ldb + i], go);
                    qz = warp_sum(qz); go = warp_sum(go);
                    const float inv = 1.f / (qz + ng);
                    for (int i = lane; i < D; i += 32) tb[r * ldb + i] *= inv;
                    if (lane == 0) sw[r] = -go * inv;
                }
                __syncthreads();
                for (int tile = tid; tile < ntM; tile += kThreads) {
                    const int ri = tile / tilesM, ci = tile - ri * tilesM;
                    float acc[4][4];
                    // acc[a][b] = sum_d dnum[4ri+a][d] * cS[ci + BS*b][d]
                    if (BS == 16) tile_abt<16>(tb, ldb, Ws, D, D, ri, ci, acc);
                    else if (BS == 8) tile_abt<8>(tb, ldb, Ws, D, D, ri, ci, acc);
                    else tile_abt<4>(tb, ldb, Ws, D, D, ri, ci, acc);
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int r = 4 * ri + a;
                        if (r >= nr) continue;
                        const float dd = sw[r];
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const int m = ci + BS * b;
                            p.dq[((r0 + r) * H + h) * M + m] = acc[a][b] + dd * zs[m] - ta[r * lda + m] * tq_over;
                        }
                    }
                }
                __syncthreads();
                // ---- dk = v (c dS)^T + c dz - k t_k/sk ;  dv = k (c dS) + du
                seg_load(ta, lda, p.k, r0, nr, H, h, M);
                seg_load(tb, ldb, p.v, r0, nr, p.Hv, hv, D);
                __syncthreads();
                for (int tile = tid; tile < ntM; tile += kThreads) {
                    const int ri = tile / tilesM, ci = tile - ri * tilesM;
                    float acc[4][4];
                    if (BS == 16) tile_abt<16>(tb, ldb, dWs, D, D, ri, ci, acc);
                    else if (BS == 8) tile_abt<8>(tb, ldb, dWs, D, D, ri, ci, acc);
                    else tile_abt<4>(tb, ldb, dWs, D, D, ri, ci, acc);
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int r = 4 * ri + a;
                        if (r >= nr) continue;
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const int m = ci + BS * b;
                            p.dk[((r0 + r) * H + h) * M + m] = acc[a][b] + dzs[m] - ta[r * lda + m] * tk_over;
                        }
                    }
                }
                for (int tile = tid; tile < ntD; tile += kThreads) {
                    const int ri = tile / tilesD, ji = tile - ri * tilesD;
                    float acc[4][4];
                    tile_mm(ta, lda, dWs, D, M, ri, ji, acc);
                    const float4 u4 = *reinterpret_cast<const float4*>(dus + 4 * ji);
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int r = 4 * ri + a;
                        if (r >= nr) continue;
                        float4* dst = reinterpret_cast<float4*>(p.dv + ((r0 + r) * p.Hv + hv) * D + 4 * ji);
                        float4 o = make_float4(acc[a][0] + u4.x, acc[a][1] + u4.y, acc[a][2] + u4.z, acc[a][3] + u4.w);
                        if (bcast && h > 0) {     // V shared by all heads: accumulate (same thread, same address, fixed order)
                            const float4 prev = *dst;
                            o.x += prev.x; o.y += prev.y; o.z += prev.z; o.w += prev.w;
                        }
                        *dst = o;
                    }
                }
                __syncthreads();
            }
        }
    }
}

__global__ void sumsq2_stage1(const float* __restrict__ q, const float* __restrict__ k, int64_t count, float* __restrict__ part) {
    __shared__ float red[33];
    float a = 0.f, b = 0.f;
    const int64_t n4 = count >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 x = ldg4(q + 4 * i), y = ldg4(k + 4 * i);
        a += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        b += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
    }
    if (blockIdx.x == 0)
        for (int64_t i = 4 * n4 + threadIdx.x; i < count; i += blockDim.x) { a += q[i] * q[i]; b += k[i] * k[i]; }
    const float fa = block_sum(a, red);
    const float fb = block_sum(b, red);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = fa; part[2 * blockIdx.x + 1] = fb; }
}

constexpr int kSumBlocks = 592;

int seg_check(int64_t N, int B, int H, int Hv, int M, int D, bool bwd) {
    DIF_REQUIRE(N >= 1 && B >= 1 && H >= 1, DIF_EARG, "segmented: bad N/B/H");
    DIF_REQUIRE(Hv == H || Hv == 1, DIF_EARG, "segmented: Hv=%d must equal H=%d or 1", Hv, H);
    DIF_REQUIRE(M >= 4 && D >= 4 && (M % 4) == 0 && (D % 4) == 0 && M <= 128 && D <= 128, DIF_EUNSUPPORTED,
                "segmented: need M,D multiples of 4 in [4,128] (got M=%d D=%d)", M, D);
    if (bwd)
        DIF_REQUIRE((M == 16 || M == 32 || M == 64) && D <= 64, DIF_EUNSUPPORTED,
                    "segmented backward: need M in {16,32,64} and D <= 64 (got M=%d D=%d)", M, D);
    return DIF_OK;
}

template <typename K>
int seg_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) DIF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return DIF_OK;
}

}  // namespace
}  // namespace dif

using namespace dif;

extern "C" int64_t dif_segmented_workspace_bytes(int32_t B) {
    const int64_t n = (int64_t)(B > kSumBlocks ? B : kSumBlocks) * 2 + 2 + 2 * 512;     // + per-CTA shares of the tensor-core backward
    return n * (int64_t)sizeof(float);
}

extern "C" int dif_sumsq2(const float* q, const float* k, int64_t count, float* norms, void* workspace, int64_t workspace_bytes, void* stream) {
    DIF_REQUIRE(q && k && norms && workspace && count >= 1, DIF_EARG, "sumsq2: bad argument");
    DIF_REQUIRE(workspace_bytes >= (int64_t)kSumBlocks * 2 * 4, DIF_EARG, "sumsq2: workspace too small");
    DIF_REQUIRE((((uintptr_t)q | (uintptr_t)k) & 15) == 0, DIF_EARG, "sumsq2: pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    sumsq2_stage1<<<kSumBlocks, 256, 0, st>>>(q, k, count, (float*)workspace);
    DIF_LAUNCH_OK();
    seg_sum_pairs_kernel<<<1, 1024, 0, st>>>((const float*)workspace, kSumBlocks, norms);
    DIF_LAUNCH_OK();
    return DIF_OK;
}

extern "C" int dif_segmented_simple_fwd(const float* q, const float* k, const float* v, const int32_t* seg_ptr, int32_t B,
                                        const float* norms, int64_t N, int H, int Hv, int M, int D, float* out, void* stream) {
    int rc = seg_check(N, B, H, Hv, M, D, false);
    if (rc) return rc;
    DIF_REQUIRE(q && k && v && seg_ptr && norms && out, DIF_EARG, "segmented_fwd: null pointer");
    SegArgs a{};
    a.q = q; a.k = k; a.v = v; a.seg = seg_ptr; a.norms = norms; a.N = N; a.B = B; a.H = H; a.Hv = Hv; a.M = M; a.D = D; a.o = out;
    const size_t smem = ((size_t)M * D + M + D + (size_t)kSegRows * (M + 4) + (size_t)kSegRows * (D + 4) + kSegRows) * sizeof(float);
    const int grid = B < 148 * 16 ? B : 148 * 16;
    const int ntile = (M / 4) * (D / 4);
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 64 && D == 64) {
        // graphs with <= 64 rows: one warp per graph, direct form (the particle datasets: 10-40 nodes per graph)
        const size_t wsmem = (size_t)kWarpsPerCta * 2 * kWarpBufFloats * sizeof(float);      // double-buffered K/V stages
        static bool attr = false;
        if (!attr) { DIF_CUDA_OK(cudaFuncSetAttribute(seg_fwd_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem)); attr = true; }
        const int wgrid = (int)std::min<int64_t>(((int64_t)B + kWarpsPerCta - 1) / kWarpsPerCta, 148 * 16);
        seg_fwd_warp_kernel<<<wgrid, kWarpsPerCta * 32, wsmem, st>>>(a);
        DIF_LAUNCH_OK();
        a.min_rows = kWarpMaxRows;       // the CTA kernel below only takes the larger graphs
    }
    const int grid_cta = a.min_rows > 0 ? (grid < 296 ? grid : 296) : grid;     // mostly skipping: a small grid is enough
    if (ntile <= kThreads) { if ((rc = seg_smem(seg_fwd_kernel<1>, smem))) return rc; seg_fwd_kernel<1><<<grid_cta, kThreads, smem, st>>>(a); }
    else if (ntile <= 2 * kThreads) { if ((rc = seg_smem(seg_fwd_kernel<2>, smem))) return rc; seg_fwd_kernel<2><<<grid, kThreads, smem, st>>>(a); }
    else { if ((rc = seg_smem(seg_fwd_kernel<4>, smem))) return rc; seg_fwd_kernel<4><<<grid, kThreads, smem, st>>>(a); }
    DIF_LAUNCH_OK();
    return DIF_OK;
}

// phase 0: everything.  Graphs sharded over ranks: phase 1 stops once the batch-wide scalars (t_q, t_k) of THIS rank's graphs
// are at workspace float offset 2 B; the caller all-reduces those two floats; phase 2 finishes (dq, dk, dv) with the totals.
extern "C" int dif_segmented_simple_bwd_phase(const float* q, const float* k, const float* v, const float* g, const float* out,
                                              const int32_t* seg_ptr, int32_t B, const float* norms,
                                              int64_t N, int H, int Hv, int M, int D, float* dq, float* dk, float* dv,
                                              void* workspace, int64_t workspace_bytes, int phase, void* stream) {
    DIF_REQUIRE(phase >= 0 && phase <= 2, DIF_EARG, "segmented_bwd: phase %d", phase);
    int rc = seg_check(N, B, H, Hv, M, D, true);
    if (rc) return rc;
    DIF_REQUIRE(q && k && v && g && out && seg_ptr && norms && dq && dk && dv && workspace, DIF_EARG, "segmented_bwd: null pointer");
    DIF_REQUIRE(workspace_bytes >= ((int64_t)B * 2 + 2) * 4, DIF_EARG, "segmented_bwd: workspace too small");
    float* part = (float*)workspace;
    float* scal = part + 2 * (int64_t)B;
    SegArgs a{};
    a.q = q; a.k = k; a.v = v; a.g = g; a.out = out; a.seg = seg_ptr; a.norms = norms; a.scal = scal;
    a.N = N; a.B = B; a.H = H; a.Hv = Hv; a.M = M; a.D = D; a.dq = dq; a.dk = dk; a.dv = dv; a.part = part;
    const int grid = B < 148 * 16 ? B : 148 * 16;
    cudaStream_t st = (cudaStream_t)stream;
    const bool warp_path = (M == 64 && D == 64);
    if (warp_path && phase != 2) {
        // graphs with <= 64 rows: one warp per graph, direct form; writes dq, dk (without the t terms), dv and part[g]
        const size_t wsmem = (size_t)kWarpsPerCta * (2 * kWarpBufFloats + 2 * kWarpMaxRows) * sizeof(float);
        static bool attr = false;
        if (!attr) { DIF_CUDA_OK(cudaFuncSetAttribute(seg_bwd_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem)); attr = true; }
        const int wgrid = (int)std::min<int64_t>(((int64_t)B + kWarpsPerCta - 1) / kWarpsPerCta, 148 * 16);
        seg_bwd_warp_kernel<<<wgrid, kWarpsPerCta * 32, wsmem, st>>>(a);
        DIF_LAUNCH_OK();
    }
    if (warp_path) a.min_rows = kWarpMaxRows;       // the CTA kernels below only take the larger graphs
    // M,D <= 64 => (M/4)*(D/4) <= 256 => one 4x4 tile per thread
    if (phase != 2) {
        const size_t smem = ((size_t)M * D + M + D + (size_t)kSegRows * (M + 4) + 2 * (size_t)kSegRows * (D + 4) + kSegRows + 33) * sizeof(float);
        if ((rc = seg_smem(seg_bwd_scalars_kernel<1>, smem))) return rc;
        seg_bwd_scalars_kernel<1><<<grid, kThreads, smem, st>>>(a);
        DIF_LAUNCH_OK();
        seg_sum_pairs_kernel<<<1, 1024, 0, st>>>(part, B, scal);
        DIF_LAUNCH_OK();
    }
    if (phase == 1) return DIF_OK;
    {
        const size_t smem = (2 * (size_t)M * D + 2 * (size_t)(M + D) + (size_t)kSegRows * (M + 4) + 2 * (size_t)kSegRows * (D + 4) + kSegRows) * sizeof(float);
        if ((rc = seg_smem(seg_bwd_main_kernel<1>, smem))) return rc;
        seg_bwd_main_kernel<1><<<grid, kThreads, smem, st>>>(a);
        DIF_LAUNCH_OK();
    }
    if (warp_path) {
        seg_bwd_fixup_kernel<<<148 * 8, 256, 0, st>>>(a);
        DIF_LAUNCH_OK();
    }
    return DIF_OK;
}

extern "C" int dif_segmented_simple_bwd(const float* q, const float* k, const float* v, const float* g, const float* out,
                                        const int32_t* seg_ptr, int32_t B, const float* norms,
                                        int64_t N, int H, int Hv, int M, int D, float* dq, float* dk, float* dv,
                                        void* workspace, int64_t workspace_bytes, void* stream) {
    return dif_segmented_simple_bwd_phase(q, k, v, g, out, seg_ptr, B, norms, N, H, Hv, M, D, dq, dk, dv, workspace, workspace_bytes, 0, stream);
}
