// Which shared-memory float does tcgen05.mma read as A[m][k] for a given smem descriptor / major-ness?
// One CTA, M = 128, N = 8, K = 8 (one kind::tf32 instruction).  B = identity (W[n][k] = (n == k), K-major, no
// swizzle ambiguity: written through the known-good K-major SW128 layout), A region = 32 KB of floats whose value
// is their own index (two launches: low 10 bits, high bits -- TF32 keeps 10 mantissa bits), so D[m][n] = index of the
// float the tensor core took for A[m][k = n].   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 8;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (sm_100)
    d |= (uint64_t)layout << 61;     // 0 none, 2 SW128, 4 SW64, 6 SW32
    return d;
}
__device__ __forceinline__ uint32_t b_off(int n, int k) {   // K-major SW128
    return (n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 2)) ^ (n & 7)) << 4) + (k & 3) * 4;
}

__global__ void __launch_bounds__(128) probe(float* __restrict__ D, int high, uint32_t lbo, uint32_t sbo, uint32_t layout, int a_mn, uint32_t start) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    float* sA = reinterpret_cast<float*>(smem);            // 32 KB
    uint8_t* sB = smem + 32768;                            // 1 KB: 8 rows x 128 B
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "n"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 8192; i += 128) sA[i] = high ? (float)(i >> 10) : (float)(i & 1023);
    for (int i = tid; i < 256; i += 128) reinterpret_cast<float*>(sB)[i] = 0.f;
    __syncthreads();
    if (tid < 8) *reinterpret_cast<float*>(sB + b_off(tid, tid)) = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? (1u << 15) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t da = make_desc(smem_u32(sA) + start, lbo, sbo, layout);
        const uint64_t db = make_desc(smem_u32(sB), 16, 1024, 2);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    }
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[tid * N + j] = done ? __uint_as_float(r[j]) : -1.0f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(32));
}

int main() {
    float* dD;
    cudaMalloc(&dD, M * N * 4);
    const size_t smem = 32768 + 1024 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    struct Cfg { const char* name; uint32_t lbo, sbo, layout; int a_mn; uint32_t start; };
    const Cfg cfgs[] = {
        {"K-major SW128 lbo16 sbo1024", 16, 1024, 2, 0, 0},
        {"MN-major SW128_BASE32B lbo1024 sbo4096", 1024, 4096, 1, 1, 0},
        {"MN-major SW128_BASE32B lbo4096 sbo1024", 4096, 1024, 1, 1, 0},
        {"MN-major SW128_BASE32B lbo512 sbo2048", 512, 2048, 1, 1, 0},
        {"MN-major SW128_BASE32B lbo2048 sbo512", 2048, 512, 1, 1, 0},
        {"MN-major layout3", 1024, 4096, 3, 1, 0},
        {"MN-major layout5", 1024, 4096, 5, 1, 0},
        {"MN-major layout7", 1024, 4096, 7, 1, 0},
    };
    std::vector<float> lo(M * N), hi(M * N);
    for (const Cfg& c : cfgs) {
        for (int pass = 0; pass < 2; ++pass) {
            cudaMemset(dD, 0, M * N * 4);
            probe<<<1, 128, smem>>>(dD, pass, c.lbo, c.sbo, c.layout, c.a_mn, c.start);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
            cudaMemcpy((pass ? hi : lo).data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        }
        printf("== %s  (float index read for A[m][k])\n", c.name);
        const int ms[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 31, 32, 33, 63, 64, 96, 127};
        for (int m : ms) {
            printf("  m=%3d:", m);
            for (int k = 0; k < 8; ++k) printf(" %5d", (int)(lo[m * N + k] + 1024.0f * hi[m * N + k]));
            printf("\n");
        }
    }
    return 0;
}
