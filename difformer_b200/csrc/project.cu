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

// D(8x8) += A(8x4) B(4x8), fp64 on the tensor cores
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];      // fixed order: deterministic
    return t;
}

// Shared-memory matrices are fp64 (widened once on the way in) with a leading dimension of 66 doubles, and every product reads BOTH
// operands k-major (row k = one contraction index), which is what the m8n8k4 fragments want.
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

    // The four 64^3-class products run on the FP64 tensor cores (mma.sync m8n8k4): the vector DFMA pipe of this part sustains only a few
    // FMAs per clock and SM (the scalar version of this kernel spent 29 us there).  Fragments: A(8x4) a = [lane/4][lane%4],
    // B(4x8) b = [lane%4][lane/4], C(8x8) c0,c1 = [lane/4][2 (lane%4) + {0,1}]; both operands are read k-major from shared memory.
    const int warp = t >> 5, lane = t & 31, gid = lane >> 2, tig = lane & 3;
    // T = Wk G (64 x 64): warp -> rows 8 warp .. +7, all 64 columns
    double skp = 0.0;
    {
        double acc[8][2];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll 4
        for (int ks = 0; ks < 16; ++ks) {
            const double a = sWkT[(4 * ks + tig) * kLd + 8 * warp + gid];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) dmma(acc[nt], a, sG[(4 * ks + tig) * kLd + 8 * nt + gid]);
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int m = 8 * warp + gid, c = 8 * nt + 2 * tig + e;
                sTT[c * kLd + m] = acc[nt][e];
                skp += acc[nt][e] * sWkT[c * kLd + m];        // <T, Wk>
            }
    }
    // this slab's 16 rows of P = Wq^T Wq (sum q^2 = <G, P>): warp -> row tile warp / 4, column tiles 2 (warp % 4), +1
    double sqp = 0.0;
    {
        const int j0 = slab * kSlabW + 8 * (warp >> 2), n0 = 16 * (warp & 3);
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll 4
        for (int ks = 0; ks < 16; ++ks) {
            const double a = sWq[(4 * ks + tig) * kLd + j0 + gid];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) dmma(acc[nt], a, sWq[(4 * ks + tig) * kLd + n0 + 8 * nt + gid]);
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) sqp += acc[nt][e] * sG[(j0 + gid) * kLd + n0 + 8 * nt + 2 * tig + e];
    }
    __syncthreads();                                          // sTT, ks, vs, qs complete
    if (t < kC) z[t] = ks[t] + p.n * bk[t];

    // S slab (64 x 16): S[m][d] = sum_c T[m][c] Wv[d][c] + ks[m] bv[d] + bk[m] vs[d] + n bk[m] bv[d]; warp -> rows 8 warp .. +7
    {
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll 4
        for (int ks_ = 0; ks_ < 16; ++ks_) {
            const double a = sTT[(4 * ks_ + tig) * kLd + 8 * warp + gid];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) dmma(acc[nt], a, sWvT[(4 * ks_ + tig) * kLd + slab * kSlabW + 8 * nt + gid]);
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int m = 8 * warp + gid, dl = 8 * nt + 2 * tig + e, d = slab * kSlabW + dl;
                sS[m * kSlabW + dl] = acc[nt][e] + ks[m] * bv[d] + bk[m] * vs[d] + p.n * bk[m] * bv[d];
            }
    }
    __syncthreads();

    // A slab (64 x 16): A[c][d] = sum_m Wq[m][c] S[m][d]; warp -> rows c = 8 warp .. +7
    {
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll 4
        for (int ks_ = 0; ks_ < 16; ++ks_) {
            const double a = sWq[(4 * ks_ + tig) * kLd + 8 * warp + gid];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) dmma(acc[nt], a, sS[(4 * ks_ + tig) * kSlabW + 8 * nt + gid]);
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int c = 8 * warp + gid, d = slab * kSlabW + 8 * nt + 2 * tig;
            *reinterpret_cast<float2*>(p.vpartials + (size_t)h * kC * kC + c * kC + d) = make_float2((float)acc[nt][0], (float)acc[nt][1]);
        }
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
    if (h == 0 && p.vbar_partials != nullptr) {
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

// The same one-head pass-2 problem from the weights alone (no dependence on G): lets the value branch of a layer (mean_h V -> SpMM) start
// before / beside the Gram pass.  vb = [S (64 x 64) | z = 0 | u = bbar | 1 | 1], one[0] = 1 (the denominator constant).
__global__ void __launch_bounds__(kThreads) project_values_kernel(const float* __restrict__ Wv, const float* __restrict__ bv, int H,
                                                                  float* __restrict__ vb, float* __restrict__ one) {
    const int t = blockIdx.x * kThreads + threadIdx.x;
    for (int i = t; i < kC * kC; i += gridDim.x * kThreads) {
        const int m = i / kC, d = i % kC;
        double acc = 0.0;
        if (Wv) { for (int hh = 0; hh < H; ++hh) acc += (double)Wv[((size_t)hh * kC + d) * kC + m]; acc /= H; }
        else acc = m == d ? 1.0 : 0.0;
        vb[i] = (float)acc;
    }
    if (t < kC) {
        double acc = 0.0;
        if (Wv) { for (int hh = 0; hh < H; ++hh) acc += (double)bv[hh * kC + t]; acc /= H; }
        vb[kC * kC + t] = 0.f;
        vb[kC * kC + kC + t] = (float)acc;
        if (t < 2) vb[kC * kC + 2 * kC + t] = 1.f;
        if (t == 0) one[0] = 1.f;
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

int simple_project_values(const float* Wv, const float* bv, int H, float* vbar_partials, float* one, cudaStream_t st) {
    project_values_kernel<<<4, kThreads, 0, st>>>(Wv, bv, H, vbar_partials, one);
    DIF_LAUNCH_OK();
    return DIF_OK;
}

}  // namespace dif
