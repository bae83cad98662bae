// Register-tiled shared-memory matmul building blocks shared by the FFMA kernels.
#pragma once
#include "common.cuh"

namespace dif {

constexpr int kThreads = 256;
constexpr int kAppRows = 64;   // rows per tile in the apply-like kernels

static __device__ __forceinline__ void fma4x4(float (&acc)[4][4], const float4& a, const float4& b) {
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
}

// ------------------------------------------------------------------------------------------
// shared tile matmul: Y[4ri+a][4ji+b] = sum_i Xs[(4ri+a)*ldx + i] * Ws[i*ldw + 4ji + b]
// ------------------------------------------------------------------------------------------
static __device__ __forceinline__ void tile_mm(const float* __restrict__ Xs, int ldx, const float* __restrict__ Ws, int ldw,
                                        int I, int ri, int ji, float (&acc)[4][4]) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const float* x0 = Xs + (4 * ri) * ldx;
    const float* w0 = Ws + 4 * ji;
#pragma unroll 2
    for (int i4 = 0; i4 < I; i4 += 4) {
        float4 x[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) x[a] = *reinterpret_cast<const float4*>(x0 + a * ldx + i4);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const float4 w = *reinterpret_cast<const float4*>(w0 + (i4 + ii) * ldw);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float xv = ii == 0 ? x[a].x : ii == 1 ? x[a].y : ii == 2 ? x[a].z : x[a].w;
                acc[a][0] = fmaf(xv, w.x, acc[a][0]);
                acc[a][1] = fmaf(xv, w.y, acc[a][1]);
                acc[a][2] = fmaf(xv, w.z, acc[a][2]);
                acc[a][3] = fmaf(xv, w.w, acc[a][3]);
            }
        }
    }
}

static __device__ __forceinline__ void load_rows(float* __restrict__ dst, int ld, const float* __restrict__ src, int64_t row0,
                                          int64_t N, int heads, int head, int W) {
    // dst[r][0..W) = src[(row0+r), head, 0..W)   for r < kAppRows (zero beyond N)
    const int w4 = W >> 2;
    for (int idx = threadIdx.x; idx < kAppRows * w4; idx += kThreads) {
        const int r = idx / w4, c4 = idx - r * w4;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < N) x = ldg4(src + ((row0 + r) * heads + head) * W + 4 * c4);
        *reinterpret_cast<float4*>(dst + r * ld + 4 * c4) = x;
    }
}


// accumulating variant: acc += X W  (no zeroing)
static __device__ __forceinline__ void tile_mm_acc(const float* __restrict__ Xs, int ldx, const float* __restrict__ Ws, int ldw,
                                                   int I, int ri, int ji, float (&acc)[4][4]) {
    const float* x0 = Xs + (4 * ri) * ldx;
    const float* w0 = Ws + 4 * ji;
#pragma unroll 2
    for (int i4 = 0; i4 < I; i4 += 4) {
        float4 x[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) x[a] = *reinterpret_cast<const float4*>(x0 + a * ldx + i4);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const float4 w = *reinterpret_cast<const float4*>(w0 + (i4 + ii) * ldw);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float xv = ii == 0 ? x[a].x : ii == 1 ? x[a].y : ii == 2 ? x[a].z : x[a].w;
                acc[a][0] = fmaf(xv, w.x, acc[a][0]);
                acc[a][1] = fmaf(xv, w.y, acc[a][1]);
                acc[a][2] = fmaf(xv, w.z, acc[a][2]);
                acc[a][3] = fmaf(xv, w.w, acc[a][3]);
            }
        }
    }
}

// row-row dot tile (A B^T): acc[a][b] = sum_i As[(4ri+a)*lda + i] * Bs[(ci + BS*b)*ldb + i]
// B rows are interleaved (ci, ci+BS, ci+2BS, ci+3BS with BS = rows_of_B/4) so that the lanes of a
// warp (consecutive ci) read consecutive shared-memory rows: 2 wavefronts per float4 instead of 8.
template <int BS>
static __device__ __forceinline__ void tile_abt(const float* __restrict__ As, int lda, const float* __restrict__ Bs, int ldb,
                                                int I, int ri, int ci, float (&acc)[4][4]) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    const float* a0 = As + (4 * ri) * lda;
    const float* b0 = Bs + ci * ldb;
#pragma unroll 2
    for (int i4 = 0; i4 < I; i4 += 4) {
        float4 x[4], y[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) x[a] = *reinterpret_cast<const float4*>(a0 + a * lda + i4);
#pragma unroll
        for (int b = 0; b < 4; ++b) y[b] = *reinterpret_cast<const float4*>(b0 + (BS * b) * ldb + i4);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                acc[a][b] = fmaf(x[a].x, y[b].x, acc[a][b]);
                acc[a][b] = fmaf(x[a].y, y[b].y, acc[a][b]);
                acc[a][b] = fmaf(x[a].z, y[b].z, acc[a][b]);
                acc[a][b] = fmaf(x[a].w, y[b].w, acc[a][b]);
            }
    }
}

// outer-product accumulation over R rows: acc[i][j] += sum_r As[r*lda + 4ai + i] * Bs[r*ldb + 4bi + j]  (A^T B)
static __device__ __forceinline__ void tile_atb_acc(const float* __restrict__ As, int lda, const float* __restrict__ Bs, int ldb,
                                                    int R, int ai, int bi, float (&acc)[4][4]) {
#pragma unroll 8
    for (int r = 0; r < R; ++r) {
        const float4 a4 = *reinterpret_cast<const float4*>(As + r * lda + 4 * ai);
        const float4 b4 = *reinterpret_cast<const float4*>(Bs + r * ldb + 4 * bi);
        fma4x4(acc, a4, b4);
    }
}

}  // namespace dif
