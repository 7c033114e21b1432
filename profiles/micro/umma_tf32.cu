// Stage-0 study for the fused Detect conv + decode kernel: one CTA, one tile
//   D[m][n] = sum_k X[k][m] * W[n][k]      M = 128 spatial positions, N = 256 channels, K = KT*32
// A = X^T straight from the NCHW activation (MN-major, fp32 bits used as TF32), B = W (K-major),
// both in the 128-byte-swizzled canonical UMMA layouts, filled with 16-byte cp.async copies;
// accumulator in TMEM, read back with tcgen05.ld.   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int M = 128, N = 256, KB = 32;   // one k-block = 32 tf32 = 128 bytes

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
// A tile (MN-major, SW128): atom = 32 m x 8 k = 1024 B; m-blocks LBO apart, k-groups SBO apart
__device__ __forceinline__ uint32_t a_off(int m, int k, uint32_t lbo, uint32_t sbo) {
    return (m >> 5) * lbo + (k >> 3) * sbo + (k & 7) * 128 + ((((m & 31) >> 2) ^ (k & 7)) << 4) + (m & 3) * 4;
}
// B tile (K-major, SW128): row n = 128 B (32 k), 8 rows per atom, atoms SBO = 1024 B apart
__device__ __forceinline__ uint32_t b_off(int n, int k) {
    return (n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 2)) ^ (n & 7)) << 4) + (k & 3) * 4;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;          // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(128) umma_tile(const float* __restrict__ X, int ldx, const float* __restrict__ W, int K,
                                                 float* __restrict__ D, int dbg) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int KT = K / KB;
    uint8_t* sA = smem;                                   // KT x [128 m x 32 k] = KT x 16 KB
    uint8_t* sB = smem + (size_t)KT * 16384;              // KT x [256 n x 32 k] = KT x 32 KB
    const uint32_t A_LBO = 1024, A_SBO = 4096;            // per 16 KB k-block: 4 m-blocks contiguous, 4 k-groups
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // fill: A chunk = (k, 4 consecutive m); B chunk = (n, 4 consecutive k)
    for (int kb = 0; kb < KT; ++kb) {
        if (dbg & 8) {                                     // A K-major (debug): plain stores, element by element
            for (int e = tid; e < 128 * 32; e += 128) {
                const int m = e & 127, k = e >> 7;
                *reinterpret_cast<float*>(sA + kb * 16384 + b_off(m, k)) = X[(size_t)(kb * KB + k) * ldx + m];
            }
        } else
        for (int e = tid; e < 32 * 32; e += 128) {        // 32 k x 32 m-chunks
            const int k = e >> 5, m4 = (e & 31) << 2;
            cp16(smem_u32(sA + kb * 16384) + a_off(m4, k, A_LBO, A_SBO), X + (size_t)(kb * KB + k) * ldx + m4);
        }
        if (dbg & 32) {
            for (int e = tid; e < 256 * 32; e += 128) {
                const int n = e & 255, k = e >> 8;
                *reinterpret_cast<float*>(sB + kb * 32768 + a_off(n, k, 1024, 8192)) = W[(size_t)n * K + kb * KB + k];
            }
        } else
        for (int e = tid; e < 256 * 8; e += 128) {        // 256 n x 8 k-chunks
            const int n = e >> 3, k4 = (e & 7) << 2;
            cp16(smem_u32(sB + kb * 32768) + b_off(n, k4), W + (size_t)n * K + kb * KB + k4);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (dbg & 1) {   // pattern: D[m][n] = m + n/1000 in TMEM, then the MMAs accumulate on top
        for (int c0 = 0; c0 < N; c0 += 4) {
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(taddr),
                         "r"(__float_as_uint(tid + (c0 + 0) * 1e-3f)), "r"(__float_as_uint(tid + (c0 + 1) * 1e-3f)),
                         "r"(__float_as_uint(tid + (c0 + 2) * 1e-3f)), "r"(__float_as_uint(tid + (c0 + 3) * 1e-3f)) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (tid == 0 && (dbg & 4)) printf("tmem_base = 0x%08x sA=0x%x sB=0x%x\n", tmem, smem_u32(sA), smem_u32(sB));
    if (tid == 0 && (dbg & 64)) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const long long t0 = clock64();
        const int REP = 256;
        for (int rep = 0; rep < REP; ++rep)
            for (int kb = 0; kb < KT; ++kb)
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t da = make_desc(smem_u32(sA + kb * 16384) + ks * 32, 16, 1024);
                    const uint64_t db = make_desc(smem_u32(sB + kb * 32768) + ks * 32, 16, 1024);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
                }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
        const long long t1 = clock64();
        printf("%d MMAs (M128 N256 K8 tf32): %lld cycles = %.1f cycles per MMA\n", REP * KT * 4, t1 - t0, (double)(t1 - t0) / (REP * KT * 4));
    } else
    if (tid == 0 && !(dbg & 2)) {
        // idesc: D = F32, A = B = TF32, A MN-major, B K-major, N = 256, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((dbg & 8) ? 0u : (1u << 15)) | ((dbg & 32) ? (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int kb = 0; kb < KT; ++kb)
            for (int ks = 0; ks < 4; ++ks) {              // UMMA_K = 8 for tf32
                const uint64_t da = (dbg & 8) ? make_desc(smem_u32(sA + kb * 16384) + ks * 32, 16, 1024)
                                              : (dbg & 16) ? make_desc(smem_u32(sA + kb * 16384) + ks * A_SBO, A_SBO, A_LBO)
                                              : make_desc(smem_u32(sA + kb * 16384) + ks * A_SBO, A_LBO, A_SBO);
                const uint64_t db = (dbg & 32) ? make_desc(smem_u32(sB + kb * 32768) + ks * 8192, 1024, 8192)
                                               : make_desc(smem_u32(sB + kb * 32768) + ks * 32, 16, 1024);
                const uint32_t acc = ((kb | ks) || (dbg & 1)) ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    }
    if (tid == 0 && (dbg & 2)) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    // everyone waits for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w reads TMEM lanes 32w..32w+31 (row m = tid), 256 columns in chunks of 32
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                       "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                       "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(256));
}

static float tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y; }

int main(int argc, char** argv) {
    const int K = argc > 1 ? atoi(argv[1]) : 64, ldx = 6400; const int dbg = argc > 2 ? atoi(argv[2]) : 0;
    std::vector<float> X((size_t)K * ldx), W((size_t)N * K), D((size_t)M * N);
    srand(1);
    for (auto& v : X) v = (rand() % 2001 - 1000) / 500.0f;
    for (auto& v : W) v = (rand() % 2001 - 1000) / 2000.0f;
    float *dX, *dW, *dD;
    cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, D.size() * 4);
    const size_t smem = (size_t)(K / KB) * (16384 + 32768) + 1024;
    cudaFuncSetAttribute(umma_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_tile<<<1, 128, smem>>>(dX + 256, ldx, dW, K, dD, dbg);     // tile starting at spatial position 256
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0; int bad = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)tf32(X[(size_t)k * ldx + 256 + m]) * (double)tf32(W[(size_t)n * K + k]);
            const double err = fabs(ref - D[(size_t)m * N + n]);
            if (err > maxerr) maxerr = err;
            if (fabs(ref) > maxref) maxref = fabs(ref);
            if (!(err <= 1e-3 * (1 + fabs(ref))) && bad++ < 5) printf("  m=%d n=%d got %g want %g\n", m, n, D[(size_t)m * N + n], ref);
        }
    printf("D[0][0..3] = %g %g %g %g   D[5][7] = %g  D[127][255] = %g\n", D[0], D[1], D[2], D[3], D[5 * N + 7], D[127 * N + 255]);
    printf("K=%d max |err| = %.3g (max |ref| = %.3g), mismatches = %d\n", K, maxerr, maxref, bad);
    return bad != 0;
}
