// kernel='simple' -- generic FFMA kernels (any H, M%4==0, D%4==0, M,D<=128).
//
// Reference path replaced: full_attention_conv(..., 'simple'), node classification/difformer.py:18-39.
// This is the shape-general CUDA path (and the baseline the tcgen05 path in simple_sm100.cu is
// validated against).  Structure shared by both:  reduce -> finalize -> apply.
//
//   reduce   : grid (row chunks, H).  Per CTA: S_h += K_h^T V_hv over its rows with a 4x4 register
//              tile per thread, plus z_h = sum k, u = sum v, sum k^2, sum q^2.  One record per
//              chunk in the workspace -- no float atomics, deterministic.
//   finalize : sums the chunk records in fixed order (double accumulation) -> partials.
//   apply    : per 64-row tile, per head: out = (q (cS) + u) / (q (cz) + N), optional fused layer
//              epilogue (head mean + addends).
//
// Backward (analytic) has the same two-pass shape: bwd_reduce (dS, dz, du, t_q),
// scalars (t_k), bwd_dq / bwd_dkv streaming kernels.
#include <math.h>

#include "common.cuh"
#include "tile.cuh"

namespace dif {

namespace {

constexpr int kRedRows = 32;   // rows per shared-memory tile in the reduce kernels

// ------------------------------------------------------------------------------------------
// reduce (forward: A = k, B = v ; backward: A = q, B = g / den with per-row weight dden)
// ------------------------------------------------------------------------------------------
struct ReduceArgs {
    const float* a;         // [N,H,M]
    const float* b;         // fwd v [N,Hv,D] ; bwd g [N,H,D]
    const float* q;         // fwd: q (sum of squares only)
    const float* out;       // bwd: saved forward output [N,H,D]
    const float* partials;  // bwd: forward partials
    float n_total;
    int64_t N;
    int H, Hb, M, D;
    int Hv_fwd;             // bwd: number of V heads of the forward pass (layout of `partials`)
    int rows_per_cta;
    int64_t ws_len;
    float* ws;
};

template <int TPT, bool BWD>
__global__ void __launch_bounds__(kThreads) reduce_kernel(ReduceArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int M = p.M, D = p.D, H = p.H;
    const int h = blockIdx.y;
    const int lda = M + 4, ldb = D + 4;
    float* sa = smem;
    float* sb = sa + kRedRows * lda;
    float* so = sb + kRedRows * ldb;                      // BWD only
    float* sw = so + (BWD ? kRedRows * ldb : 0);          // BWD only: per-row dden
    float* szc = sw + (BWD ? kRedRows : 0);               // BWD only: c * z[h]
    float* su = szc + (BWD ? M : 0);                      // BWD only: u[hv]
    float* red = su + (BWD ? D : 0);                      // 33 floats
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tilesD = D >> 2, ntile = (M >> 2) * tilesD;

    float acc[TPT][4][4];
#pragma unroll
    for (int t = 0; t < TPT; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.f;
    float zacc = 0.f, uacc = 0.f, ss_a = 0.f, ss_q = 0.f, tq = 0.f;

    const int Hb = p.Hb;                                  // heads of the B tensor (fwd: Hv, bwd: H)
    const int hbb = (Hb == H) ? h : 0;
    const int64_t offZ = (int64_t)H * M * D, offU = offZ + (int64_t)H * M;   // same in both layouts
    const int64_t row_begin = (int64_t)blockIdx.x * p.rows_per_cta;
    const int64_t row_end = min(p.N, row_begin + (int64_t)p.rows_per_cta);

    if (BWD) {
        // backward prologue: c*z[h] and u[hv] of the forward partials into shared memory
        const int64_t oSq = offU + (int64_t)p.Hv_fwd * D;
        const float c = 1.f / (sqrtf(p.partials[oSq]) * sqrtf(p.partials[oSq + 1]));
        const int hv_fwd = (p.Hv_fwd == H) ? h : 0;
        for (int i = tid; i < M; i += kThreads) szc[i] = p.partials[offZ + (int64_t)h * M + i] * c;
        for (int i = tid; i < D; i += kThreads) su[i] = p.partials[offU + (int64_t)hv_fwd * D + i];
    }

    for (int64_t row0 = row_begin; row0 < row_end; row0 += kRedRows) {
        const int nr = (int)min((int64_t)kRedRows, row_end - row0);
        const int m4 = M >> 2, d4 = D >> 2;
        for (int idx = tid; idx < kRedRows * m4; idx += kThreads) {
            const int r = idx / m4, c4 = idx - r * m4;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nr) {
                x = ldg4(p.a + ((row0 + r) * H + h) * M + 4 * c4);
                if (!BWD) {
                    ss_a += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
                    const float4 y = ldg4(p.q + ((row0 + r) * H + h) * M + 4 * c4);
                    ss_q += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
                }
            }
            *reinterpret_cast<float4*>(sa + r * lda + 4 * c4) = x;
        }
        for (int idx = tid; idx < kRedRows * d4; idx += kThreads) {
            const int r = idx / d4, c4 = idx - r * d4;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
            if (r < nr) {
                x = ldg4(p.b + ((row0 + r) * Hb + hbb) * D + 4 * c4);
                if (BWD) y = ldg4(p.out + ((row0 + r) * H + h) * D + 4 * c4);
            }
            *reinterpret_cast<float4*>(sb + r * ldb + 4 * c4) = x;
            if (BWD) *reinterpret_cast<float4*>(so + r * ldb + 4 * c4) = y;
        }
        __syncthreads();
        if (BWD) {
            // per-row scalars: den = q^.z^ + N ; dnum = g/den ; dden = -(g.out)/den
            for (int r = warp; r < kRedRows; r += kThreads / 32) {
                float qz = 0.f, go = 0.f, gu = 0.f;
                for (int i = lane; i < M; i += 32) qz = fmaf(sa[r * lda + i], szc[i], qz);
                for (int i = lane; i < D; i += 32) {
                    const float gg = sb[r * ldb + i];
                    go = fmaf(gg, so[r * ldb + i], go);
                    gu = fmaf(gg, su[i], gu);
                }
                qz = warp_sum(qz); go = warp_sum(go); gu = warp_sum(gu);
                const float inv = 1.f / (qz + p.n_total);
                const float dden = -go * inv;
                for (int i = lane; i < D; i += 32) sb[r * ldb + i] *= inv;
                if (lane == 0) {
                    sw[r] = dden;
                    // <q^, dq^> contribution: sum_d dnum (out*den - u) + dden (den - N)
                    if (r < nr) tq += go - inv * gu + dden * qz;
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int t = 0; t < TPT; ++t) {
            const int tile = tid + t * kThreads;
            if (tile < ntile) {
                const int mi = tile / tilesD, di = tile - mi * tilesD;
#pragma unroll 8
                for (int r = 0; r < kRedRows; ++r) {
                    const float4 a4 = *reinterpret_cast<const float4*>(sa + r * lda + 4 * mi);
                    const float4 b4 = *reinterpret_cast<const float4*>(sb + r * ldb + 4 * di);
                    fma4x4(acc[t], a4, b4);
                }
            }
        }
        if (tid < M) {
#pragma unroll 8
            for (int r = 0; r < kRedRows; ++r) zacc = fmaf(sa[r * lda + tid], BWD ? sw[r] : 1.f, zacc);
        } else if (tid < M + D) {
#pragma unroll 8
            for (int r = 0; r < kRedRows; ++r) uacc += sb[r * ldb + tid - M];
        }
        __syncthreads();
    }

    float* rec = p.ws + (int64_t)blockIdx.x * p.ws_len;
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
        const int tile = tid + t * kThreads;
        if (tile < ntile) {
            const int mi = tile / tilesD, di = tile - mi * tilesD;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(rec + ((int64_t)h * M + 4 * mi + i) * D + 4 * di) =
                    make_float4(acc[t][i][0], acc[t][i][1], acc[t][i][2], acc[t][i][3]);
        }
    }
    if (tid < M) {
        rec[offZ + (int64_t)h * M + tid] = zacc;
    } else if (tid < M + D) {
        if (BWD) rec[offU + (int64_t)h * D + tid - M] = uacc;                   // du is per head
        else if (Hb == H || h == 0) rec[offU + (int64_t)hbb * D + tid - M] = uacc;
    }
    if (BWD) {
        const float t = block_sum(tq, red);
        if (tid == 0) rec[offU + (int64_t)H * D + h] = t;
    } else {
        const int64_t oSq = offU + (int64_t)Hb * D;
        const float s1 = block_sum(ss_q, red);
        const float s2 = block_sum(ss_a, red);
        if (tid == 0) { rec[oSq + h] = s1; rec[oSq + H + h] = s2; }
    }
}

// sums chunk records -> partials.  main part: element-wise; tail: `nscal` scalars, each the sum of
// H per-head slots.  256 threads = 8 outputs x 32 chunk groups: every thread issues its (<= 5 at 148
// records) loads back to back, the groups are combined in fixed order in shared memory
// (deterministic, fp64 accumulation).  ~2 L2 round trips instead of a 148-long dependent chain.
constexpr int kFinOut = 8, kFinGroups = 32;
__global__ void __launch_bounds__(256) finalize_kernel(const float* __restrict__ ws, int nchunks, int64_t ws_len, int64_t main_len,
                                                       int nscal, int H, float* __restrict__ out) {
    __shared__ double red[kFinGroups][kFinOut + 1];
    const int jl = threadIdx.x & (kFinOut - 1), cg = threadIdx.x / kFinOut;
    const int64_t j = (int64_t)blockIdx.x * kFinOut + jl;
    double s = 0.0;
    if (j < main_len) {
        const float* p = ws + j;
        int c = cg;
        for (; c + 3 * kFinGroups < nchunks; c += 4 * kFinGroups) {
            const float a0 = p[(int64_t)c * ws_len], a1 = p[(int64_t)(c + kFinGroups) * ws_len];
            const float a2 = p[(int64_t)(c + 2 * kFinGroups) * ws_len], a3 = p[(int64_t)(c + 3 * kFinGroups) * ws_len];
            s += (double)a0; s += (double)a1; s += (double)a2; s += (double)a3;
        }
        for (; c < nchunks; c += kFinGroups) s += (double)p[(int64_t)c * ws_len];
    } else if (j < main_len + nscal) {
        const int k = (int)(j - main_len);
        for (int c = cg; c < nchunks; c += kFinGroups)
            for (int h = 0; h < H; ++h) s += (double)ws[(int64_t)c * ws_len + main_len + (int64_t)k * H + h];
    }
    red[cg][jl] = s;
    __syncthreads();
    if (cg == 0 && j < main_len + nscal) {
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < kFinGroups; ++g) t += red[g][jl];
        out[j] = (float)t;
    }
}

// ------------------------------------------------------------------------------------------
// apply (forward pass 2)
// ------------------------------------------------------------------------------------------
struct ApplyArgs {
    const float* q;
    const float* partials;
    float n_total;
    int64_t N;
    int H, Hv, M, D;
    float* out;
    dif_epilogue_t ep;
};

template <int TPT, int MODE>
__global__ void __launch_bounds__(kThreads) apply_kernel(ApplyArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int M = p.M, D = p.D, H = p.H;
    const int ldx = M + 4;
    float* Ws = smem;                 // [M][D]  c*S[h]
    float* zs = Ws + M * D;           // [M]     c*z[h]
    float* us = zs + M;               // [D]     u[hv]
    float* Xs = us + D;               // [R][M+4]
    float* sden = Xs + kAppRows * ldx;  // [R]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SimpleLayout L{H, p.Hv, M, D};
    const float c = 1.f / (sqrtf(p.partials[L.offSq()]) * sqrtf(p.partials[L.offSk()]));
    const int64_t row0 = (int64_t)blockIdx.x * kAppRows;
    const int tilesJ = D >> 2, ntile = (kAppRows >> 2) * tilesJ;

    float hsum[TPT][4][4];
    if (MODE == 1) {
#pragma unroll
        for (int t = 0; t < TPT; ++t)
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) hsum[t][a][b] = 0.f;
    }

    for (int h = 0; h < H; ++h) {
        const int hv = (p.Hv == H) ? h : 0;
        const float* S = p.partials + L.offS() + (int64_t)h * M * D;
        for (int idx = tid; idx < (M * D) >> 2; idx += kThreads) {
            float4 s = ldg4(S + 4 * idx);
            s.x *= c; s.y *= c; s.z *= c; s.w *= c;
            *reinterpret_cast<float4*>(Ws + 4 * idx) = s;
        }
        for (int i = tid; i < M; i += kThreads) zs[i] = p.partials[L.offZ() + (int64_t)h * M + i] * c;
        for (int i = tid; i < D; i += kThreads) us[i] = p.partials[L.offU() + (int64_t)hv * D + i];
        load_rows(Xs, ldx, p.q, row0, p.N, H, h, M);
        __syncthreads();
        for (int r = warp; r < kAppRows; r += kThreads / 32) {
            float qz = 0.f;
            for (int i = lane; i < M; i += 32) qz = fmaf(Xs[r * ldx + i], zs[i], qz);
            qz = warp_sum(qz);
            if (lane == 0) sden[r] = qz + p.n_total;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < TPT; ++t) {
            const int tile = tid + t * kThreads;
            if (tile < ntile) {
                const int ri = tile / tilesJ, ji = tile - ri * tilesJ;
                float acc[4][4];
                tile_mm(Xs, ldx, Ws, D, M, ri, ji, acc);
                const float4 u4 = *reinterpret_cast<const float4*>(us + 4 * ji);
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int64_t row = row0 + 4 * ri + a;
                    const float den = sden[4 * ri + a];
                    float4 o;
                    o.x = (acc[a][0] + u4.x) / den;
                    o.y = (acc[a][1] + u4.y) / den;
                    o.z = (acc[a][2] + u4.z) / den;
                    o.w = (acc[a][3] + u4.w) / den;
                    if (MODE == 0) {
                        if (row < p.N) *reinterpret_cast<float4*>(p.out + (row * H + h) * D + 4 * ji) = o;
                    } else {
                        hsum[t][a][0] += o.x; hsum[t][a][1] += o.y; hsum[t][a][2] += o.z; hsum[t][a][3] += o.w;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (MODE == 1) {
#pragma unroll
        for (int t = 0; t < TPT; ++t) {
            const int tile = tid + t * kThreads;
            if (tile < ntile) {
                const int ri = tile / tilesJ, ji = tile - ri * tilesJ;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int64_t row = row0 + 4 * ri + a;
                    if (row >= p.N) continue;
                    float4 o = make_float4(hsum[t][a][0] * p.ep.attn_scale, hsum[t][a][1] * p.ep.attn_scale,
                                           hsum[t][a][2] * p.ep.attn_scale, hsum[t][a][3] * p.ep.attn_scale);
                    for (int j = 0; j < p.ep.n_add; ++j) {
                        const float4 x = ldg4(p.ep.add[j] + row * D + 4 * ji);
                        const float s = p.ep.add_scale[j];
                        o.x = fmaf(s, x.x, o.x); o.y = fmaf(s, x.y, o.y); o.z = fmaf(s, x.z, o.z); o.w = fmaf(s, x.w, o.w);
                    }
                    *reinterpret_cast<float4*>(p.out + row * D + 4 * ji) = o;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward streaming kernels
// ------------------------------------------------------------------------------------------
struct BwdArgs {
    const float *q, *k, *v, *g, *out, *partials, *bwd;
    float n_total;
    int64_t N;
    int H, Hv, M, D;
    float *dq, *dk, *dv;
};

__global__ void bwd_scalars_kernel(const float* __restrict__ partials, float* __restrict__ bwd, int H, int Hv, int M, int D) {
    __shared__ float red[33];
    const SimpleLayout L{H, Hv, M, D};
    const BwdLayout B{H, M, D};
    float s = 0.f;
    const int64_t n = L.offU();   // S and z are contiguous in both layouts with identical offsets
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s = fmaf(partials[i], bwd[i], s);
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
        const float c = 1.f / (sqrtf(partials[L.offSq()]) * sqrtf(partials[L.offSk()]));
        bwd[B.offTk()] = c * s;
    }
}

// dq[n,h,:] = c (dnum S^T + dden z) - q t_q / sum(q^2)
template <int TPT>
__global__ void __launch_bounds__(kThreads) bwd_dq_kernel(BwdArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int M = p.M, D = p.D, H = p.H;
    const int ldq = M + 4, ldg = D + 4;
    float* Ws = smem;                     // [D][M]   c*S[h]^T
    float* zs = Ws + M * D;               // [M]
    float* us = zs + M;                   // [D]
    float* Xq = us + D;                   // [R][M+4]
    float* Xg = Xq + kAppRows * ldq;      // [R][D+4]  g -> dnum
    float* Xo = Xg + kAppRows * ldg;      // [R][D+4]
    float* sdd = Xo + kAppRows * ldg;     // [R]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SimpleLayout L{H, p.Hv, M, D};
    const BwdLayout B{H, M, D};
    const float sq = p.partials[L.offSq()];
    const float c = 1.f / (sqrtf(sq) * sqrtf(p.partials[L.offSk()]));
    const float tq_over = p.bwd[B.offTq()] / sq;
    const int64_t row0 = (int64_t)blockIdx.x * kAppRows;
    const int tilesJ = M >> 2, ntile = (kAppRows >> 2) * tilesJ;

    for (int h = 0; h < H; ++h) {
        const int hv = (p.Hv == H) ? h : 0;
        const float* S = p.partials + L.offS() + (int64_t)h * M * D;
        for (int idx = tid; idx < M * D; idx += kThreads) {
            const int m = idx / D, d = idx - m * D;
            Ws[d * M + m] = __ldg(S + idx) * c;
        }
        for (int i = tid; i < M; i += kThreads) zs[i] = p.partials[L.offZ() + (int64_t)h * M + i] * c;
        for (int i = tid; i < D; i += kThreads) us[i] = p.partials[L.offU() + (int64_t)hv * D + i];
        load_rows(Xq, ldq, p.q, row0, p.N, H, h, M);
        load_rows(Xg, ldg, p.g, row0, p.N, H, h, D);
        load_rows(Xo, ldg, p.out, row0, p.N, H, h, D);
        __syncthreads();
        for (int r = warp; r < kAppRows; r += kThreads / 32) {
            float qz = 0.f, go = 0.f;
            for (int i = lane; i < M; i += 32) qz = fmaf(Xq[r * ldq + i], zs[i], qz);
            for (int i = lane; i < D; i += 32) go = fmaf(Xg[r * ldg + i], Xo[r * ldg + i], go);
            qz = warp_sum(qz); go = warp_sum(go);
            const float inv = 1.f / (qz + p.n_total);
            for (int i = lane; i < D; i += 32) Xg[r * ldg + i] *= inv;
            if (lane == 0) sdd[r] = -go * inv;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < TPT; ++t) {
            const int tile = tid + t * kThreads;
            if (tile < ntile) {
                const int ri = tile / tilesJ, ji = tile - ri * tilesJ;
                float acc[4][4];
                tile_mm(Xg, ldg, Ws, M, D, ri, ji, acc);
                const float4 z4 = *reinterpret_cast<const float4*>(zs + 4 * ji);
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int64_t row = row0 + 4 * ri + a;
                    if (row >= p.N) continue;
                    const float dd = sdd[4 * ri + a];
                    const float4 qv = *reinterpret_cast<const float4*>(Xq + (4 * ri + a) * ldq + 4 * ji);
                    float4 o;
                    o.x = acc[a][0] + dd * z4.x - qv.x * tq_over;
                    o.y = acc[a][1] + dd * z4.y - qv.y * tq_over;
                    o.z = acc[a][2] + dd * z4.z - qv.z * tq_over;
                    o.w = acc[a][3] + dd * z4.w - qv.w * tq_over;
                    *reinterpret_cast<float4*>(p.dq + (row * H + h) * M + 4 * ji) = o;
                }
            }
        }
        __syncthreads();
    }
}

// dk[l,h,:] = c (v dS^T + dz) - k t_k / sum(k^2) ;  dv[l,hv,:] = sum_h c k dS + du
template <int TPT_K, int TPT_V>
__global__ void __launch_bounds__(kThreads) bwd_dkv_kernel(BwdArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int M = p.M, D = p.D, H = p.H;
    const int ldk = M + 4, ldv = D + 4;
    float* Ws = smem;                     // [D][M] then [M][D]
    float* dzs = Ws + M * D;              // [M]  c*dz[h]
    float* dus = dzs + M;                 // [D]  du[h]
    float* Xk = dus + D;                  // [R][M+4]
    float* Xv = Xk + kAppRows * ldk;      // [R][D+4]
    const int tid = threadIdx.x;
    const SimpleLayout L{H, p.Hv, M, D};
    const BwdLayout B{H, M, D};
    const float sk = p.partials[L.offSk()];
    const float c = 1.f / (sqrtf(p.partials[L.offSq()]) * sqrtf(sk));
    const float tk_over = p.bwd[B.offTk()] / sk;
    const int64_t row0 = (int64_t)blockIdx.x * kAppRows;
    const int tilesK = M >> 2, ntileK = (kAppRows >> 2) * tilesK;
    const int tilesV = D >> 2, ntileV = (kAppRows >> 2) * tilesV;
    const bool bcast = (p.Hv != H);

    float vsum[TPT_V][4][4];
#pragma unroll
    for (int t = 0; t < TPT_V; ++t)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) vsum[t][a][b] = 0.f;

    for (int h = 0; h < H; ++h) {
        const int hv = bcast ? 0 : h;
        const float* dS = p.bwd + B.offS() + (int64_t)h * M * D;
        for (int idx = tid; idx < M * D; idx += kThreads) {
            const int m = idx / D, d = idx - m * D;
            Ws[d * M + m] = __ldg(dS + idx) * c;
        }
        for (int i = tid; i < M; i += kThreads) dzs[i] = p.bwd[B.offZ() + (int64_t)h * M + i] * c;
        for (int i = tid; i < D; i += kThreads) dus[i] = p.bwd[B.offU() + (int64_t)h * D + i];
        load_rows(Xk, ldk, p.k, row0, p.N, H, h, M);
        if (h == 0 || !bcast) load_rows(Xv, ldv, p.v, row0, p.N, p.Hv, hv, D);
        __syncthreads();
        // ---- dk
#pragma unroll
        for (int t = 0; t < TPT_K; ++t) {
            const int tile = tid + t * kThreads;
            if (tile < ntileK) {
                const int ri = tile / tilesK, ji = tile - ri * tilesK;
                float acc[4][4];
                tile_mm(Xv, ldv, Ws, M, D, ri, ji, acc);
                const float4 z4 = *reinterpret_cast<const float4*>(dzs + 4 * ji);
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int64_t row = row0 + 4 * ri + a;
                    if (row >= p.N) continue;
                    const float4 kv = *reinterpret_cast<const float4*>(Xk + (4 * ri + a) * ldk + 4 * ji);
                    float4 o;
                    o.x = acc[a][0] + z4.x - kv.x * tk_over;
                    o.y = acc[a][1] + z4.y - kv.y * tk_over;
                    o.z = acc[a][2] + z4.z - kv.z * tk_over;
                    o.w = acc[a][3] + z4.w - kv.w * tk_over;
                    *reinterpret_cast<float4*>(p.dk + (row * H + h) * M + 4 * ji) = o;
                }
            }
        }
        __syncthreads();
        // ---- dv : W = c*dS[h]  [M][D]
        for (int idx = tid; idx < (M * D) >> 2; idx += kThreads) {
            float4 s = ldg4(dS + 4 * idx);
            s.x *= c; s.y *= c; s.z *= c; s.w *= c;
            *reinterpret_cast<float4*>(Ws + 4 * idx) = s;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < TPT_V; ++t) {
            const int tile = tid + t * kThreads;
            if (tile < ntileV) {
                const int ri = tile / tilesV, ji = tile - ri * tilesV;
                float acc[4][4];
                tile_mm(Xk, ldk, Ws, D, M, ri, ji, acc);
                const float4 u4 = *reinterpret_cast<const float4*>(dus + 4 * ji);
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const float o0 = acc[a][0] + u4.x, o1 = acc[a][1] + u4.y, o2 = acc[a][2] + u4.z, o3 = acc[a][3] + u4.w;
                    if (bcast) {
                        vsum[t][a][0] += o0; vsum[t][a][1] += o1; vsum[t][a][2] += o2; vsum[t][a][3] += o3;
                    } else {
                        const int64_t row = row0 + 4 * ri + a;
                        if (row < p.N)
                            *reinterpret_cast<float4*>(p.dv + (row * H + h) * D + 4 * ji) = make_float4(o0, o1, o2, o3);
                    }
                }
            }
        }
        __syncthreads();
    }
    if (bcast) {
#pragma unroll
        for (int t = 0; t < TPT_V; ++t) {
            const int tile = tid + t * kThreads;
            if (tile < ntileV) {
                const int ri = tile / tilesV, ji = tile - ri * tilesV;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int64_t row = row0 + 4 * ri + a;
                    if (row < p.N)
                        *reinterpret_cast<float4*>(p.dv + row * D + 4 * ji) =
                            make_float4(vsum[t][a][0], vsum[t][a][1], vsum[t][a][2], vsum[t][a][3]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
int check_shape(int64_t N, int H, int Hv, int M, int D) {
    DIF_REQUIRE(N >= 1 && H >= 1 && H <= 1024, DIF_EARG, "simple: bad N=%lld or H=%d", (long long)N, H);
    DIF_REQUIRE(Hv == H || Hv == 1, DIF_EARG, "simple: Hv=%d must equal H=%d or 1 (difformer.py:120)", Hv, H);
    DIF_REQUIRE(M >= 4 && D >= 4 && (M % 4) == 0 && (D % 4) == 0 && M <= 128 && D <= 128, DIF_EUNSUPPORTED,
                "simple (generic): need M,D multiples of 4 in [4,128], got M=%d D=%d", M, D);
    return DIF_OK;
}

int reduce_chunks(int64_t N, int H, int* rows_per_cta) {
    const int64_t target = ((int64_t)sm_count() * 4 + H - 1) / H;   // ~4 CTAs per SM over all heads
    const int64_t max_chunks = (N + kRedRows - 1) / kRedRows;
    int64_t chunks = target < max_chunks ? target : max_chunks;
    if (chunks < 1) chunks = 1;
    int64_t rpc = (N + chunks - 1) / chunks;
    rpc = (rpc + kRedRows - 1) / kRedRows * kRedRows;
    chunks = (N + rpc - 1) / rpc;
    *rows_per_cta = (int)rpc;
    return (int)chunks;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) DIF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return DIF_OK;
}

template <bool BWD>
int launch_reduce(const ReduceArgs& a, int nchunks, cudaStream_t st) {
    const int M = a.M, D = a.D;
    size_t smem = (size_t)kRedRows * (M + 4) + (size_t)kRedRows * (D + 4) + 33;
    if (BWD) smem += (size_t)kRedRows * (D + 4) + kRedRows + M + D;
    smem *= sizeof(float);
    const int ntile = (M / 4) * (D / 4);
    dim3 grid(nchunks, a.H);
    int rc;
    if (ntile <= kThreads) {
        if ((rc = set_smem(reduce_kernel<1, BWD>, smem))) return rc;
        reduce_kernel<1, BWD><<<grid, kThreads, smem, st>>>(a);
    } else if (ntile <= 2 * kThreads) {
        if ((rc = set_smem(reduce_kernel<2, BWD>, smem))) return rc;
        reduce_kernel<2, BWD><<<grid, kThreads, smem, st>>>(a);
    } else {
        if ((rc = set_smem(reduce_kernel<4, BWD>, smem))) return rc;
        reduce_kernel<4, BWD><<<grid, kThreads, smem, st>>>(a);
    }
    DIF_LAUNCH_OK();
    return DIF_OK;
}

}  // namespace

int64_t simple_generic_workspace_bytes(int64_t N, int H, int Hv, int M, int D) {
    int rpc;
    const int chunks = reduce_chunks(N, H, &rpc);
    const SimpleLayout L{H, Hv, M, D};
    const BwdLayout B{H, M, D};
    const int64_t rec = L.wsLen() > B.wsLen() ? L.wsLen() : B.wsLen();
    return (int64_t)chunks * rec * (int64_t)sizeof(float);
}

int simple_finalize_fwd(const float* ws, int nchunks, int H, int Hv, int M, int D, float* partials, cudaStream_t st);

int simple_reduce_generic(const float* q, const float* k, const float* v, int64_t N, int H, int Hv, int M, int D,
                          float* partials, void* ws, int64_t ws_bytes, cudaStream_t st) {
    int rc = check_shape(N, H, Hv, M, D);
    if (rc) return rc;
    int rpc;
    const int chunks = reduce_chunks(N, H, &rpc);
    const SimpleLayout L{H, Hv, M, D};
    DIF_REQUIRE(ws_bytes >= (int64_t)chunks * L.wsLen() * 4, DIF_EARG, "simple_reduce: workspace too small (%lld)", (long long)ws_bytes);
    ReduceArgs a{};
    a.a = k; a.b = v; a.q = q; a.N = N; a.H = H; a.Hb = Hv; a.M = M; a.D = D;
    a.rows_per_cta = rpc; a.ws_len = L.wsLen(); a.ws = (float*)ws;
    if ((rc = launch_reduce<false>(a, chunks, st))) return rc;
    return simple_finalize_fwd((const float*)ws, chunks, H, Hv, M, D, partials, st);
}

int simple_finalize_fwd(const float* ws, int nchunks, int H, int Hv, int M, int D, float* partials, cudaStream_t st) {
    const SimpleLayout L{H, Hv, M, D};
    const int64_t main_len = L.offSq();
    const int blocks = (int)((main_len + 2 + kFinOut - 1) / kFinOut);
    finalize_kernel<<<blocks, 256, 0, st>>>(ws, nchunks, L.wsLen(), main_len, 2, H, partials);
    DIF_LAUNCH_OK();
    return DIF_OK;
}

int simple_apply_generic(const float* q, const float* partials, double n_total, int64_t N, int H, int Hv, int M, int D,
                         float* out, const dif_epilogue_t* ep, cudaStream_t st) {
    int rc = check_shape(N, H, Hv, M, D);
    if (rc) return rc;
    ApplyArgs a{};
    a.q = q; a.partials = partials; a.n_total = (float)n_total; a.N = N; a.H = H; a.Hv = Hv; a.M = M; a.D = D; a.out = out;
    if (ep) a.ep = *ep; else { a.ep = dif_epilogue_t{}; }
    DIF_REQUIRE(a.ep.mode == 0 || a.ep.mode == 1, DIF_EARG, "simple_apply: epilogue mode %d", a.ep.mode);
    DIF_REQUIRE(a.ep.n_add >= 0 && a.ep.n_add <= 3, DIF_EARG, "simple_apply: n_add %d", a.ep.n_add);
    DIF_REQUIRE(a.ep.mode == 0 || (a.ep.ln_weight == nullptr && a.ep.relu == 0 && a.ep.gcn_rowptr == nullptr), DIF_EUNSUPPORTED,
                "simple_apply: the LayerNorm / ReLU tail and the in-epilogue gcn term are fused on the tcgen05 kernels only");
    const size_t smem = ((size_t)M * D + M + D + (size_t)kAppRows * (M + 4) + kAppRows) * sizeof(float);
    const int grid = (int)((N + kAppRows - 1) / kAppRows);
    const int ntile = (kAppRows / 4) * (D / 4);
#define DIF_APPLY(TPT, MODE)                                                       \
    do {                                                                           \
        if ((rc = set_smem(apply_kernel<TPT, MODE>, smem))) return rc;             \
        apply_kernel<TPT, MODE><<<grid, kThreads, smem, st>>>(a);                  \
    } while (0)
    if (ntile <= kThreads) { if (a.ep.mode == 0) DIF_APPLY(1, 0); else DIF_APPLY(1, 1); }
    else                   { if (a.ep.mode == 0) DIF_APPLY(2, 0); else DIF_APPLY(2, 1); }
#undef DIF_APPLY
    DIF_LAUNCH_OK();
    return DIF_OK;
}

}  // namespace dif

// ------------------------------------------------------------------------------------------
// C ABI: backward of 'simple' (generic kernels serve every shape)
// ------------------------------------------------------------------------------------------
using namespace dif;

extern "C" int64_t dif_simple_bwd_partials_len(int H, int M, int D) { return BwdLayout{H, M, D}.len(); }

namespace dif {
void simple_bwd_scalars(const float* partials, float* bwd_partials, int H, int Hv, int M, int D, cudaStream_t st) {
    bwd_scalars_kernel<<<1, 1024, 0, st>>>(partials, bwd_partials, H, Hv, M, D);
}
}  // namespace dif

// floats of per-(node, head) scratch the tcgen05 backward passes between its two calls (0: generic path only)
extern "C" int64_t dif_simple_bwd_rowscal_len(int64_t N, int H, int Hv, int M, int D) {
    return simple_tc_supported(N, H, Hv, M, D) ? simple_tc_rowscal_floats(N, H) : 0;
}

extern "C" int dif_simple_bwd_reduce(const float* q, const float* g, const float* out, const float* partials,
                                     double n_total, int64_t N, int H, int Hv, int M, int D,
                                     float* bwd_partials, float* rowscal, void* workspace, int64_t workspace_bytes, int impl, void* stream) {
    DIF_REQUIRE(q && g && out && partials && bwd_partials && workspace, DIF_EARG, "simple_bwd_reduce: null pointer");
    if (impl != DIF_IMPL_GENERIC && rowscal != nullptr && simple_tc_supported(N, H, Hv, M, D))
        return simple_bwd_reduce_tc(q, g, out, partials, n_total, N, H, bwd_partials, rowscal, workspace, workspace_bytes, (cudaStream_t)stream);
    DIF_REQUIRE(impl != DIF_IMPL_TCGEN05, DIF_EUNSUPPORTED, "simple_bwd_reduce: tcgen05 path needs a tcgen05 shape and the rowscal buffer");
    int rc = check_shape(N, H, Hv, M, D);
    if (rc) return rc;
    int rpc;
    const int chunks = reduce_chunks(N, H, &rpc);
    const BwdLayout B{H, M, D};
    DIF_REQUIRE(workspace_bytes >= (int64_t)chunks * B.wsLen() * 4, DIF_EARG, "simple_bwd_reduce: workspace too small");
    ReduceArgs a{};
    a.a = q; a.b = g; a.out = out; a.partials = partials; a.n_total = (float)n_total;
    a.N = N; a.H = H; a.Hb = H; a.Hv_fwd = Hv; a.M = M; a.D = D;
    a.rows_per_cta = rpc; a.ws_len = B.wsLen(); a.ws = (float*)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = launch_reduce<true>(a, chunks, st))) return rc;
    const int64_t main_len = B.offTq();
    const int blocks = (int)((main_len + 1 + kFinOut - 1) / kFinOut);
    finalize_kernel<<<blocks, 256, 0, st>>>((const float*)workspace, chunks, B.wsLen(), main_len, 1, H, bwd_partials);
    DIF_LAUNCH_OK();
    return DIF_OK;
}

extern "C" int dif_simple_bwd_apply(const float* q, const float* k, const float* v, const float* g, const float* out,
                                    const float* partials, float* bwd_partials, const float* rowscal, double n_total,
                                    int64_t N, int H, int Hv, int M, int D,
                                    float* dq, float* dk, float* dv, int impl, void* stream) {
    DIF_REQUIRE(q && k && v && g && out && partials && bwd_partials && dq && dk && dv, DIF_EARG, "simple_bwd_apply: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (impl != DIF_IMPL_GENERIC && rowscal != nullptr && simple_tc_supported(N, H, Hv, M, D)) {
        bwd_scalars_kernel<<<1, 1024, 0, st>>>(partials, bwd_partials, H, Hv, M, D);      // t_k from the (reduced) dS
        DIF_LAUNCH_OK();
        return simple_bwd_apply_tc(q, k, v, g, partials, bwd_partials, rowscal, N, H, dq, dk, dv, st);
    }
    DIF_REQUIRE(impl != DIF_IMPL_TCGEN05, DIF_EUNSUPPORTED, "simple_bwd_apply: tcgen05 path needs a tcgen05 shape and the rowscal buffer");
    int rc = check_shape(N, H, Hv, M, D);
    if (rc) return rc;
    bwd_scalars_kernel<<<1, 1024, 0, st>>>(partials, bwd_partials, H, Hv, M, D);
    DIF_LAUNCH_OK();
    BwdArgs a{q, k, v, g, out, partials, bwd_partials, (float)n_total, N, H, Hv, M, D, dq, dk, dv};
    const int grid = (int)((N + kAppRows - 1) / kAppRows);
    {
        const size_t smem = ((size_t)M * D + M + D + (size_t)kAppRows * (M + 4) + 2 * (size_t)kAppRows * (D + 4) + kAppRows) * sizeof(float);
        const int ntile = (kAppRows / 4) * (M / 4);
        if (ntile <= kThreads) {
            if ((rc = set_smem(bwd_dq_kernel<1>, smem))) return rc;
            bwd_dq_kernel<1><<<grid, kThreads, smem, st>>>(a);
        } else {
            if ((rc = set_smem(bwd_dq_kernel<2>, smem))) return rc;
            bwd_dq_kernel<2><<<grid, kThreads, smem, st>>>(a);
        }
        DIF_LAUNCH_OK();
    }
    {
        const size_t smem = ((size_t)M * D + M + D + (size_t)kAppRows * (M + 4) + (size_t)kAppRows * (D + 4)) * sizeof(float);
        const int tk = (kAppRows / 4) * (M / 4) <= kThreads ? 1 : 2;
        const int tv = (kAppRows / 4) * (D / 4) <= kThreads ? 1 : 2;
#define DIF_DKV(A, B_)                                                         \
    do {                                                                       \
        if ((rc = set_smem(bwd_dkv_kernel<A, B_>, smem))) return rc;           \
        bwd_dkv_kernel<A, B_><<<grid, kThreads, smem, st>>>(a);                \
    } while (0)
        if (tk == 1 && tv == 1) DIF_DKV(1, 1);
        else if (tk == 1) DIF_DKV(1, 2);
        else if (tv == 1) DIF_DKV(2, 1);
        else DIF_DKV(2, 2);
#undef DIF_DKV
        DIF_LAUNCH_OK();
    }
    return DIF_OK;
}
