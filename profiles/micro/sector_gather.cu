// Design study for the sparse filter: how many scattered 32-byte sectors per second can a B200 fetch from HBM?
// The sparse Detect filter gathers 85 sectors per surviving prediction row from an NCHW tensor (one per channel plane,
// 25.6 KB apart), so its floor is a sector RATE, not a byte rate.  Patterns (buffer 548 MB = one 64-image conv-output set):
//   random   every load hits an independent random sector
//   rows     a warp reads 85 sectors strided by 25 600 B (one prediction row), 3 loads per lane, rows random
//   rows x4  as above, 4 rows (12 loads per lane) in flight per warp
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sector_gather sector_gather.cu ; run: ./sector_gather
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int PER_THREAD>
__global__ void random_sectors(const float* __restrict__ buf, uint32_t nsectors, float* out, int rounds) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    for (int r = 0; r < rounds; ++r) {
        float v[PER_THREAD];
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) {
            const uint32_t s = mix(tid * 131u + r * 977u + k * 7919u) % nsectors;
            v[k] = __ldg(buf + (size_t)s * 8);
        }
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) acc += v[k];
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int ROWS>
__global__ void row_sectors(const float* __restrict__ buf, int planes, int nynx, float* out, int rounds) {
    // plane = 85 channels x nynx floats; a warp gathers ROWS random (plane, s) rows per round
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int r = 0; r < rounds; ++r) {
        float v[ROWS][3];
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            const uint32_t h = mix(warp * 131u + r * 977u + j * 7919u);
            const size_t base = (size_t)(h % planes) * 85 * nynx + (mix(h) % nynx);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int c = lane + 32 * k;
                v[j][k] = c < 85 ? __ldg(buf + base + (size_t)c * nynx) : 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < ROWS; ++j) acc += v[j][0] + v[j][1] + v[j][2];
    }
    if (acc == 123.456f) out[0] = acc;
}

int main() {
    const size_t bytes = 64ull * 25200 * 85 * 4;
    float *buf, *out;
    cudaMalloc(&buf, bytes); cudaMalloc(&out, 4);
    cudaMemset(buf, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char* name, double sectors, float ms) {
        printf("%-34s %8.1f us  %7.1f G sectors/s  %7.1f GB/s of 32-byte sectors\n", name, ms * 1e3, sectors / ms / 1e6, sectors * 32 / ms / 1e6);
    };
    const uint32_t nsec = (uint32_t)(bytes / 32);
    for (int blocks_per_sm : {4, 8}) {
        const int grid = 148 * blocks_per_sm, rounds = 8;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            random_sectors<8><<<grid, 256>>>(buf, nsec, out, rounds);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        char nm[64]; snprintf(nm, 64, "random, 8 loads/thread, %d blk/SM", blocks_per_sm);
        report(nm, (double)grid * 256 * 8 * rounds, ms);
    }
    const int planes = 64 * 3, nynx = 6400;   // the 80x80 level of 64 images
    for (int blocks_per_sm : {4, 8}) {
        const int grid = 148 * blocks_per_sm, rounds = 8;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            row_sectors<1><<<grid, 256>>>(buf, planes, nynx, out, rounds);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        char nm[64]; snprintf(nm, 64, "rows x1, %d blk/SM", blocks_per_sm);
        report(nm, (double)grid * 8 * 85 * rounds, ms);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            row_sectors<4><<<grid, 256>>>(buf, planes, nynx, out, rounds);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        cudaEventElapsedTime(&ms, e0, e1);
        snprintf(nm, 64, "rows x4, %d blk/SM", blocks_per_sm);
        report(nm, (double)grid * 8 * 85 * 4 * rounds, ms);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
