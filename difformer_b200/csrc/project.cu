// The per-layer algebra of the folded projections (reference: node classification/difformer.py:115-140 computes
// Q = x Wq^T + bq, K = x Wk^T + bk, V = x Wv^T + bv and hands them to full_attention_conv :18-39).
//
// Pass 1 of 'simple' run on x itself gives the Gram matrix G = X^T X [64,64] and the column sums s = X^T 1 [64].  Everything the
// second pass needs follows from (G, s) and the layer's weights with a few 64 x 64 products per head, in fp64 (FP64 FMAs: 1 M per
// head -- microseconds on a handful of SMs, where the same algebra through torch ops is ~50 launches):
//
//     T   = Wk_h G                                   S_h = T Wv_h^T + (Wk_h s) bv_h^T + bk_h (Wv_h s)^T + n bk_h bv_h^T
//     z_h = Wk_h s + n bk_h                          u_h = Wv_h s + n bv_h
//     sum k^2 = sum_h <T, Wk_h> + 2 bk_h.(Wk_h s) + n |bk_h|^2          (sum q^2 alike with Wq, bq)
//     A_h = Wq_h^T S_h     a_h = bq_h^T S_h     w_h = Wq_h^T z_h     beta_h = bq_h . z_h
//     c   = 1 / sqrt(sum q^2 sum k^2)
//     vpartials = [A | w | u + c a | sum q^2 | sum k^2]       n_total_vec[h] = n + c beta_h
//
// project_head_kernel: grid (4 column slabs, H heads), also the head mean of Wv; project_finish_kernel: the cross-head scalar c.
#include "common.cuh"

namespace dif {
namespace {

constexpr int kC = 64;                 // hidden size (in = out = 64)
constexpr int kSlabs = 4;              // 16 columns of S / A per block
constexpr int kSlabW = kC / kSlabs;
constexpr int kThreads = 256;

// workspace (doubles) per head: a[64] | u[64] | beta | sk | sq[kSlabs]
constexpr int kWsHead = 2 * kC + 2 + kSlabs;

struct ProjectArgs {
    const float* gram;                 // partials layout of (H=1, Hv=1, 64, 64): G[4096] | s[64] | s[64] | sum x^2 | sum x^2
    const float *Wq, *bq, *Wk, *bk, *Wv, *bv;     // nn.Linear layout: weight [H*64, 64] row-major, bias [H*64]; Wv null: V_h = x
    double n;
    int H;
    float* vpartials;
    float* vbar_partials;              // partials layout of (1, 1, 64, 64) that makes pass 2 compute mean_h V = x wbar^T + bbar
    double* ws;
};

__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];      // fixed order: deterministic
    return t;
}

// Shared-memory matrices are fp64 (widened once on the way in: an F2F per FMA would cost more than the DFMA) with a leading dimension of
// 66 doubles, and every product reads BOTH operands k-major (row k = one contraction index): per k a thread loads a few consecutive
// doubles of each operand (LDS.128, broadcast across the warp's other threads) and does a register-tiled outer product.
constexpr int kLd = kC + 2;

__global__ void __launch_bounds__(kThreads) project_head_kernel(ProjectArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sG = reinterpret_cast<double*>(smem_raw);         // G[j][c]
    double* sWkT = sG + kC * kLd;                             // Wk^T[j][m]
    double* sWq = sWkT + kC * kLd;                            // Wq[m][c]
    double* sWvT = sWq + kC * kLd;                            // Wv^T[c][d]
    double* sTT = sWvT + kC * kLd;                            // T^T[c][m], T = Wk G
    double* sS = sTT + kC * kLd;                              // S slab [m][kSlabW]
    double* sv = sS + kC * kSlabW;                            // s | ks | vs | qs | bk | bv | bq | z : 8 x 64
    double* red = sv + 8 * kC;                                // 8
    double *s_ = sv, *ks = sv + kC, *vs = sv + 2 * kC, *qs = sv + 3 * kC, *bk = sv + 4 * kC, *bv = sv + 5 * kC, *bq = sv + 6 * kC,
           *z = sv + 7 * kC;

    const int t = threadIdx.x, slab = blockIdx.x, h = blockIdx.y;
    const int H = p.H;
#pragma unroll 4
    for (int i = t; i < kC * kC; i += kThreads) {
        const int r = i / kC, c = i % kC;
        sG[r * kLd + c] = p.gram[i];
        sWkT[c * kLd + r] = p.Wk[(size_t)h * kC * kC + i];
        sWq[r * kLd + c] = p.Wq[(size_t)h * kC * kC + i];
        sWvT[c * kLd + r] = p.Wv ? (double)p.Wv[(size_t)h * kC * kC + i] : (r == c ? 1.0 : 0.0);
    }
    if (t < kC) {
        s_[t] = p.gram[kC * kC + t];
        bk[t] = p.bk[h * kC + t];
        bq[t] = p.bq[h * kC + t];
        bv[t] = p.Wv ? (double)p.bv[h * kC + t] : 0.0;
    }
    __syncthreads();

    if (t < 3 * kC) {                                         // ks = Wk s, vs = Wv s, qs = Wq s
        const int r = t & (kC - 1);
        double acc = 0.0;
        if (t < kC) { for (int j = 0; j < kC; ++j) acc += sWkT[j * kLd + r] * s_[j]; ks[r] = acc; }
        else if (t < 2 * kC) { for (int j = 0; j < kC; ++j) acc += sWvT[j * kLd + r] * s_[j]; vs[r] = acc; }
        else { for (int j = 0; j < kC; ++j) acc += sWq[r * kLd + ((j + r) & (kC - 1))] * s_[(j + r) & (kC - 1)]; qs[r] = acc; }
    }

    // T = Wk G (64 x 64): thread -> rows 4 ty .. +3, columns {2 tx, 2 tx + 1, 32 + 2 tx, 33 + 2 tx}
    double skp = 0.0;
    {
        const int ty = t >> 4, tx = t & 15;
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 8
        for (int k = 0; k < kC; ++k) {
            const double2 a0 = *reinterpret_cast<const double2*>(sWkT + k * kLd + 4 * ty), a1 = *reinterpret_cast<const double2*>(sWkT + k * kLd + 4 * ty + 2);
            const double2 b0 = *reinterpret_cast<const double2*>(sG + k * kLd + 2 * tx), b1 = *reinterpret_cast<const double2*>(sG + k * kLd + 32 + 2 * tx);
            const double a[4] = {a0.x, a0.y, a1.x, a1.y}, b[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = 4 * ty + i, c = (j < 2 ? 2 * tx + j : 32 + 2 * tx + (j - 2));
                sTT[c * kLd + m] = acc[i][j];
                skp += acc[i][j] * sWkT[c * kLd + m];         // <T, Wk>
            }
    }
    // this slab's 16 rows of P = Wq^T Wq (sum q^2 = <G, P>): thread -> rows slab*16 + 4 ty .. +3 (ty < 4), column tx (64)
    double sqp = 0.0;
    {
        const int ty = t >> 6, tx = t & 63, j0 = slab * kSlabW + 4 * ty;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
        for (int k = 0; k < kC; ++k) {
            const double2 a0 = *reinterpret_cast<const double2*>(sWq + k * kLd + j0), a1 = *reinterpret_cast<const double2*>(sWq + k * kLd + j0 + 2);
            const double b = sWq[k * kLd + tx];
            acc[0] += a0.x * b; acc[1] += a0.y * b; acc[2] += a1.x * b; acc[3] += a1.y * b;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) sqp += acc[i] * sG[(j0 + i) * kLd + tx];
    }
    __syncthreads();                                          // sTT, ks, vs, qs complete
    if (t < kC) z[t] = ks[t] + p.n * bk[t];

    // S slab: S[m][d] = sum_c T[m][c] Wv[d][c] + ks[m] bv[d] + bk[m] vs[d] + n bk[m] bv[d]; thread -> row m = t >> 2, 4 columns
    {
        const int m = t >> 2, dl0 = (t & 3) * 4, d0 = slab * kSlabW + dl0;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
        for (int c = 0; c < kC; ++c) {
            const double tv = sTT[c * kLd + m];
            const double2 b0 = *reinterpret_cast<const double2*>(sWvT + c * kLd + d0), b1 = *reinterpret_cast<const double2*>(sWvT + c * kLd + d0 + 2);
            acc[0] += tv * b0.x; acc[1] += tv * b0.y; acc[2] += tv * b1.x; acc[3] += tv * b1.y;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int d = d0 + i;
            sS[m * kSlabW + dl0 + i] = acc[i] + ks[m] * bv[d] + bk[m] * vs[d] + p.n * bk[m] * bv[d];
        }
    }
    __syncthreads();

    // A slab: A[c][d] = sum_m Wq[m][c] S[m][d]; thread -> row c = t >> 2, 4 columns
    {
        const int c = t >> 2, dl0 = (t & 3) * 4;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
        for (int m = 0; m < kC; ++m) {
            const double w = sWq[m * kLd + c];
            const double2 b0 = *reinterpret_cast<const double2*>(sS + m * kSlabW + dl0), b1 = *reinterpret_cast<const double2*>(sS + m * kSlabW + dl0 + 2);
            acc[0] += w * b0.x; acc[1] += w * b0.y; acc[2] += w * b1.x; acc[3] += w * b1.y;
        }
        float4 o = make_float4((float)acc[0], (float)acc[1], (float)acc[2], (float)acc[3]);
        *reinterpret_cast<float4*>(p.vpartials + (size_t)h * kC * kC + c * kC + slab * kSlabW + dl0) = o;
    }
    double* wsh = p.ws + (size_t)h * kWsHead;
    if (t < kSlabW) {                                         // a[d] = bq^T S, u[d] = vs + n bv
        const int d = slab * kSlabW + t;
        double acc = 0.0;
        for (int m = 0; m < kC; ++m) acc += bq[m] * sS[m * kSlabW + t];
        wsh[d] = acc;
        wsh[kC + d] = vs[d] + p.n * bv[d];
    }
    if (slab == 0 && t >= 64 && t < 64 + kC) {                // w[c] = Wq^T z
        const int c = t - 64;
        double acc = 0.0;
        for (int m = 0; m < kC; ++m) acc += sWq[m * kLd + c] * z[m];
        p.vpartials[(size_t)H * kC * kC + h * kC + c] = (float)acc;
    }
    if (h == 0) {
        // mean_h V = x wbar^T + bbar (the input of the gcn term, difformer.py:139) as a pass-2 problem with one head:
        // S[m][d] = mean_h Wv_h[d][m], z = 0, u = bbar, sum q^2 = sum k^2 = 1 (c = 1), denominator constant 1
        float* vb = p.vbar_partials;
        for (int i = slab * (kC * kC / kSlabs) + t; i < (slab + 1) * (kC * kC / kSlabs); i += kThreads) {
            const int m = i / kC, d = i % kC;
            double acc = 0.0;
            if (p.Wv) { for (int hh = 0; hh < H; ++hh) acc += (double)p.Wv[((size_t)hh * kC + d) * kC + m]; acc /= H; }
            else acc = m == d ? 1.0 : 0.0;
            vb[i] = (float)acc;
        }
        if (slab == 0 && t < kC) {
            double acc = 0.0;
            if (p.Wv) { for (int hh = 0; hh < H; ++hh) acc += (double)p.bv[hh * kC + t]; acc /= H; }
            vb[kC * kC + t] = 0.f;
            vb[kC * kC + kC + t] = (float)acc;
            if (t < 2) vb[kC * kC + 2 * kC + t] = 1.f;
        }
    }
    // scalars
    double e_q = 0.0, e_k = 0.0, e_b = 0.0;
    if (t < kC) {
        e_k = 2.0 * bk[t] * ks[t] + p.n * bk[t] * bk[t];
        e_q = 2.0 * bq[t] * qs[t] + p.n * bq[t] * bq[t];
        e_b = bq[t] * z[t];
    }
    const double sk = block_sum(skp + e_k, red);
    const double sq = block_sum(sqp + (slab == 0 ? e_q : 0.0), red);
    const double beta = block_sum(e_b, red);
    if (t == 0) {
        wsh[2 * kC + 2 + slab] = sq;
        if (slab == 0) {
            wsh[2 * kC] = beta;
            wsh[2 * kC + 1] = sk;
        }
    }
}

struct FinishArgs {
    const double* ws;
    const float *Wv, *bv;
    double n;
    int H;
    float *vpartials, *nvec;
};

__global__ void __launch_bounds__(kThreads) project_finish_kernel(FinishArgs p) {
    const int t = threadIdx.x, H = p.H;
    double sq = 0.0, sk = 0.0;
    for (int h = 0; h < H; ++h) {                             // every thread: the same fixed-order sum
        const double* w = p.ws + (size_t)h * kWsHead;
        sk += w[2 * kC + 1];
        for (int j = 0; j < kSlabs; ++j) sq += w[2 * kC + 2 + j];
    }
    const double c = 1.0 / sqrt(sq * sk);
    float* uo = p.vpartials + (size_t)H * kC * kC + H * kC;
    for (int i = t; i < H * kC; i += kThreads) {
        const double* w = p.ws + (size_t)(i / kC) * kWsHead;
        uo[i] = (float)(w[kC + (i % kC)] + c * w[i % kC]);
    }
    if (t < H) p.nvec[t] = (float)(p.n + c * p.ws[(size_t)t * kWsHead + 2 * kC]);
    if (t == H) p.nvec[H] = 1.f;                              // the denominator constant of the vbar problem
    if (t == 0) {
        uo[H * kC] = (float)sq;
        uo[H * kC + 1] = (float)sk;
    }
}

constexpr size_t kSmemHead = (5 * kC * kLd + kC * kSlabW + 8 * kC + 8) * sizeof(double);

}  // namespace

int64_t simple_project_workspace_bytes(int H) { return (int64_t)H * kWsHead * (int64_t)sizeof(double); }

int simple_project(const float* gram, const float* Wq, const float* bq, const float* Wk, const float* bk, const float* Wv, const float* bv,
                   double n_total, int H, float* vpartials, float* nvec, float* vbar_partials, void* ws, cudaStream_t st) {
    static bool attr_set[64] = {};
    int dev = 0;
    DIF_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        DIF_CUDA_OK(cudaFuncSetAttribute(project_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemHead));
        if (dev < 64) attr_set[dev] = true;
    }
    ProjectArgs a{gram, Wq, bq, Wk, bk, Wv, bv, n_total, H, vpartials, vbar_partials, reinterpret_cast<double*>(ws)};
    project_head_kernel<<<dim3(kSlabs, H), kThreads, kSmemHead, st>>>(a);
    DIF_LAUNCH_OK();
    FinishArgs f{reinterpret_cast<const double*>(ws), Wv, bv, n_total, H, vpartials, nvec};
    project_finish_kernel<<<1, kThreads, 0, st>>>(f);
    DIF_LAUNCH_OK();
    return DIF_OK;
}

}  // namespace dif
