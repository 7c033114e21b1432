// Does TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B write a [32 k][32 s] fp32 box the way tcgen05 reads an MN-major TF32
// operand (SWIZZLE_128B_BASE32B: 32-byte chunk c of k-row r at c ^ (r & 3), profiles/micro/umma_probe.cu)?
// One CTA loads one box of X[k][s] (value = k * 1000 + s) and dumps shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a profiles/micro/tma_probe.cu -o tma_probe   (driver API via cudaGetDriverEntryPoint)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void __launch_bounds__(32) probe(const __grid_constant__ CUtensorMap tm, float* __restrict__ out, int c0, int c1) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(&bar), sd = (uint32_t)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(sb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sb), "r"(4096) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(sd), "l"(&tm), "r"(c0), "r"(c1), "r"(sb) : "memory");
    }
    __syncwarp();
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(sb), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 1024; i += 32) out[i] = done ? reinterpret_cast<float*>(smem)[i] : -1.0f;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int K = 64, S = 400;                        // X[k][s], 400 positions per plane (a stride-32 level)
    std::vector<float> X((size_t)K * S);
    for (int k = 0; k < K; ++k) for (int s = 0; s < S; ++s) X[(size_t)k * S + s] = k * 1000.0f + s;
    float *dX, *dO;
    cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dO, 4096);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr);
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    const CUtensorMapSwizzle modes[] = {CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B_FLIP_8B};
    const char* names[] = {"128B_ATOM_32B", "128B", "128B_ATOM_32B_FLIP_8B"};
    for (int mi = 0; mi < 3; ++mi) {
        CUtensorMap tm;
        const cuuint64_t dims[2] = {(cuuint64_t)S, (cuuint64_t)K};
        const cuuint64_t strides[1] = {(cuuint64_t)S * 4};
        const cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dX, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         modes[mi], CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", names[mi], (int)r); continue; }
        for (int test = 0; test < 2; ++test) {
            const int c0 = test ? 384 : 64, c1 = test ? 32 : 0;   // second test: 16 of the 32 positions are past the plane
            probe<<<1, 32, 4096 + 1024>>>(tm, dO, c0, c1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", names[mi], cudaGetErrorString(e)); return 1; }
            std::vector<float> o(1024);
            cudaMemcpy(o.data(), dO, 4096, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int k = 0; k < 32; ++k)
                for (int m = 0; m < 32; ++m) {
                    const int idx = k * 32 + (m ^ ((k & 3) << 3));                       // BASE32B: float index of (m, k)
                    const float want = (c0 + m < S) ? (c1 + k) * 1000.0f + (c0 + m) : 0.0f;
                    if (o[idx] != want) ++bad;
                }
            printf("%s box at (s=%d, k=%d): %d of 1024 floats differ from the BASE32B layout; row 1 = %g %g %g %g %g %g %g %g | %g ...\n",
                   names[mi], c0, c1, bad, o[32], o[33], o[34], o[35], o[36], o[37], o[38], o[39], o[40]);
        }
    }
    return 0;
}
