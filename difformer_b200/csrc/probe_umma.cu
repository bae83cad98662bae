// Stand-alone probe: checks the tcgen05 shared-memory descriptor conventions used by simple_sm100.cu
// against exact integer-valued bf16 matrices.  Run on the B200 box: build/probe_umma
//
//   mode 0 : A K-major SW128 [M=128 rows][K=64], B K-major SW128 [N=80 rows][K=64]      (pass 2)
//   mode 1 : A MN-major SW128 (M=128 = 2 heads x 64, K=32 nodes), B MN-major SW128 N=128 (pass 1)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
    d |= (uint64_t)layout << 61;     // 2 = SWIZZLE_128B
    return d;
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;                    // D = f32
    d |= 1u << 7;                    // A = bf16
    d |= 1u << 10;                   // B = bf16
    d |= (uint32_t)a_mn << 15;
    d |= (uint32_t)b_mn << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

// element (row r, 16-byte chunk c, elem e) of a [rows][64 bf16] tile in the SW128 layout
__device__ __forceinline__ uint32_t sw128_off(int r, int c, int e) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((c ^ (r & 7)) & 7) << 4) + e * 2);
}

struct Params {
    int mode;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    int a_mn, b_mn;
    int M, N, ksteps;
    uint32_t a_kstep_bytes, b_kstep_bytes;
};

// A source: mode 0: A[m][k] (m<128,k<64) ; mode 1: K[node][mm] node<32, mm<128 (2 heads x 64)
// B source: mode 0: B[n][k] (n<80,k<64)  ; mode 1: V[node][nn] node<32, nn<128
__global__ void __launch_bounds__(128) probe(Params p, const float* __restrict__ Asrc, const float* __restrict__ Bsrc, float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;               // up to 16 KB
    uint8_t* sB = smem + 16384;       // up to 16 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < 32768 / 4; i += 128) ((uint32_t*)smem)[i] = 0;
    __syncthreads();
    if (p.mode == 0) {
        for (int i = tid; i < 128 * 64; i += 128) {
            const int r = i / 64, k = i % 64;
            *(__nv_bfloat16*)(sA + sw128_off(r, k >> 3, k & 7)) = __float2bfloat16(Asrc[i]);
        }
        for (int i = tid; i < 80 * 64; i += 128) {
            const int r = i / 64, k = i % 64;
            *(__nv_bfloat16*)(sB + sw128_off(r, k >> 3, k & 7)) = __float2bfloat16(Bsrc[i]);
        }
    } else {
        // per-head tiles [32 nodes][64 mm] at head*4096 bytes
        for (int i = tid; i < 32 * 128; i += 128) {
            const int node = i / 128, mm = i % 128, head = mm >> 6, m = mm & 63;
            *(__nv_bfloat16*)(sA + head * 4096 + sw128_off(node, m >> 3, m & 7)) = __float2bfloat16(Asrc[i]);
            *(__nv_bfloat16*)(sB + head * 4096 + sw128_off(node, m >> 3, m & 7)) = __float2bfloat16(Bsrc[i]);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    if (tid == 0) {
        const uint32_t idesc = make_idesc(p.M, p.N, p.a_mn, p.b_mn);
        for (int ks = 0; ks < p.ksteps; ++ks) {
            const uint64_t ad = make_desc(smem_u32(sA) + ks * p.a_kstep_bytes, p.a_lbo, p.a_sbo, 2);
            const uint64_t bd = make_desc(smem_u32(sB) + ks * p.b_kstep_bytes, p.b_lbo, p.b_sbo, 2);
            const uint32_t accum = ks > 0 ? 1u : 0u;
            asm volatile("{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pp;\n\t}"
                         :: "r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(accum) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    // wait for the MMAs (phase 0)
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred pq;\n\tmbarrier.try_wait.parity.shared::cta.b64 pq, [%1], %2;\n\tselp.b32 %0, 1, 0, pq;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // each warp reads its 32 lanes, 128 columns in 4 chunks of 32
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                       "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                       "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[(size_t)tid * 128 + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128));
}

static double run(const Params& p, const std::vector<float>& A, const std::vector<float>& B, const std::vector<float>& ref, int cols) {
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, 128 * 128 * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, 128 * 128 * 4));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024));
    probe<<<1, 128, 34 * 1024>>>(p, dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  kernel failed: %s\n", cudaGetErrorString(e)); exit(2); }
    std::vector<float> D(128 * 128);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int i = 0; i < p.M; ++i) for (int j = 0; j < cols; ++j) {
        double d = fabs((double)D[i * 128 + j] - (double)ref[i * 128 + j]);
        if (d > maxerr) maxerr = d;
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return maxerr;
}

int main() {
    srand(1);
    {   // ---------------- mode 0: K-major
        std::vector<float> A(128 * 64), B(80 * 64), ref(128 * 128, 0.f);
        for (auto& x : A) x = (float)(rand() % 7 - 3);
        for (auto& x : B) x = (float)(rand() % 5 - 2);
        for (int i = 0; i < 128; ++i) for (int j = 0; j < 80; ++j) { float s = 0; for (int k = 0; k < 64; ++k) s += A[i * 64 + k] * B[j * 64 + k]; ref[i * 128 + j] = s; }
        Params p{}; p.mode = 0; p.M = 128; p.N = 80; p.ksteps = 4; p.a_mn = 0; p.b_mn = 0; p.a_kstep_bytes = 32; p.b_kstep_bytes = 32;
        const uint32_t cand[][2] = {{0, 1024}, {16, 1024}, {1024, 0}, {1024, 1024}};
        for (auto& c : cand) {
            p.a_lbo = p.b_lbo = c[0]; p.a_sbo = p.b_sbo = c[1];
            printf("mode0 K-major SW128 lbo=%u sbo=%u : max|err| = %g\n", c[0], c[1], run(p, A, B, ref, 80));
        }
    }
    {   // ---------------- mode 1: MN-major, D[mm][nn] = sum_node K[node][mm] V[node][nn]
        std::vector<float> A(32 * 128), B(32 * 128), ref(128 * 128, 0.f);
        for (auto& x : A) x = (float)(rand() % 7 - 3);
        for (auto& x : B) x = (float)(rand() % 5 - 2);
        for (int i = 0; i < 128; ++i) for (int j = 0; j < 128; ++j) { float s = 0; for (int n = 0; n < 32; ++n) s += A[n * 128 + i] * B[n * 128 + j]; ref[i * 128 + j] = s; }
        Params p{}; p.mode = 1; p.M = 128; p.N = 128; p.ksteps = 2; p.a_mn = 1; p.b_mn = 1; p.a_kstep_bytes = 2048; p.b_kstep_bytes = 2048;
        // head tiles are 4096 B apart (32 nodes x 128 B); 8-node groups 1024 B apart
        const uint32_t cand[][2] = {{4096, 1024}, {1024, 4096}, {4096, 0}, {0, 4096}};
        for (auto& c : cand) {
            p.a_lbo = p.b_lbo = c[0]; p.a_sbo = p.b_sbo = c[1];
            printf("mode1 MN-major SW128 lbo=%u sbo=%u : max|err| = %g\n", c[0], c[1], run(p, A, B, ref, 128));
        }
    }
    return 0;
}
