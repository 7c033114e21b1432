// Microbenchmark: which load/store path moves a 64x85 fp32 tile fastest on B200?
// (design study for detect_decode_kernel; not part of the library)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o transpose_bw transpose_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#ifndef TS_
#define TS_ 64
#endif
constexpr int NO = 85, TS = TS_, THR = 256, QN = TS / 4, JN = 3 * QN / 8;
constexpr int TILE = NO * TS;   // 5440 floats

__device__ __forceinline__ float sig(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}
__device__ __forceinline__ void cpa4(uint32_t d, const float* s) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(s) : "memory"); }
__device__ __forceinline__ void cpa16(uint32_t d, const float* s) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(s) : "memory"); }
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// E1: plain elementwise, one float4 x4 per thread (like a torch elementwise kernel)
__global__ void e1_plain(const float4* __restrict__ in, float4* __restrict__ out, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x * 4 + threadIdx.x;
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) if (i + u * blockDim.x < n4) v[u] = in[i + u * blockDim.x];
#pragma unroll
    for (int u = 0; u < 4; ++u) if (i + u * blockDim.x < n4) {
        float4 r = make_float4(sig(v[u].x), sig(v[u].y), sig(v[u].z), sig(v[u].w));
        out[i + u * blockDim.x] = r;
    }
}

// mode: 0 = LDGSTS.32 transposing (plane-strided source), 1 = LDGSTS.128 contiguous source, 2 = LDGSTS.32 contiguous source (no transposition),
//       3 = LDG.128 contiguous -> STS.128
// wmode: 0 = sigmoid + STG.128, 1 = no store (read-only), 2 = no load (write only)
template <int MODE>
__device__ __forceinline__ void prefetch(const float* in, int nynx, size_t plane0, int s0, int t, float* tile) {
    if (MODE == 0) {
        const int r = threadIdx.x & 63, cg = threadIdx.x >> 6;
        const float* src = in + plane0 + (size_t)cg * nynx + s0 + r;
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile + r * NO + cg);
        for (int c = cg; c < NO; c += 4, src += (size_t)4 * nynx, dst += 16) cpa4(dst, src);
    } else if (MODE == 4) {
        const float* src0 = in + plane0 + s0;
        for (int e = threadIdx.x; e < NO * QN; e += THR) {
            const int c = e / QN, q = e % QN;
            cpa16((uint32_t)__cvta_generic_to_shared(tile + c * TS + ((q ^ (c & 7)) << 2)), src0 + (size_t)c * nynx + 4 * q);
        }
    } else if (MODE == 1) {
        const float* src = in + (size_t)t * TILE;
        for (int e = threadIdx.x; e < TILE / 4; e += THR) cpa16((uint32_t)__cvta_generic_to_shared(tile + 4 * e), src + 4 * e);
    } else if (MODE == 2) {
        const float* src = in + (size_t)t * TILE;
        for (int e = threadIdx.x; e < TILE; e += THR) cpa4((uint32_t)__cvta_generic_to_shared(tile + e), src + e);
    } else {
        const float4* src = reinterpret_cast<const float4*>(in + (size_t)t * TILE);
        for (int e = threadIdx.x; e < TILE / 4; e += THR) reinterpret_cast<float4*>(tile)[e] = src[e];
    }
    commit();
}

template <int MODE, int WMODE>
__global__ void __launch_bounds__(THR) persist(const float* __restrict__ in, float* __restrict__ out, int tiles_per_img, int nynx, int total) {
    extern __shared__ __align__(16) float sm[];
    int t = blockIdx.x;
    if (t >= total) return;
    auto loc = [&](int tt, size_t& plane0, int& s0) {
        const int b = tt / tiles_per_img, ti = tt - b * tiles_per_img;     // one level: 3 anchors x (nynx/64) tiles
        const int tpa = nynx / TS, a = ti / tpa;
        s0 = (ti - a * tpa) * TS;
        plane0 = ((size_t)(b * 3 + a) * NO) * nynx;
    };
    size_t p0; int s0;
    loc(t, p0, s0);
    if (WMODE != 2) prefetch<MODE>(in, nynx, p0, s0, t, sm);
    for (int k = 0; t < total; ++k, t += gridDim.x) {
        float* tile = sm + (k & 1) * TILE;
        wait0();
        __syncthreads();
        const int tn = t + gridDim.x;
        if (tn < total && WMODE != 2) { loc(tn, p0, s0); prefetch<MODE>(in, nynx, p0, s0, tn, sm + ((k + 1) & 1) * TILE); }
        float* o = out + (size_t)t * TILE;
        float4 v[(TILE / 4 + THR - 1) / THR];
#pragma unroll
        for (int it = 0; it < (TILE / 4 + THR - 1) / THR; ++it) { const int e = threadIdx.x + it * THR; if (e < TILE / 4) v[it] = reinterpret_cast<const float4*>(tile)[e]; }
#pragma unroll
        for (int it = 0; it < (TILE / 4 + THR - 1) / THR; ++it) {
            const int e = threadIdx.x + it * THR;
            if (e < TILE / 4) {
                float4 r = make_float4(sig(v[it].x), sig(v[it].y), sig(v[it].z), sig(v[it].w));
                if (WMODE != 1 || r.x == 123.f) reinterpret_cast<float4*>(o)[e] = r;
            }
        }
    }
}

// (A): LDGSTS.128 strided planes -> [c][64] swizzled; LDS.128 along rows, lanes over channels; 4 scalar STG per thread
// (C): same loads; smem->smem transposition into the linear [row][85] layout; linear LDS.128 -> STG.128
template <int VAR>
__global__ void __launch_bounds__(THR) persistA(const float* __restrict__ in, float* __restrict__ out, int tiles_per_img, int nynx, int total) {
    extern __shared__ __align__(16) float sm[];
    float* lin = sm + 2 * TILE;
    int t = blockIdx.x;
    if (t >= total) return;
    auto loc = [&](int tt, size_t& plane0, int& s0) {
        const int b = tt / tiles_per_img, ti = tt - b * tiles_per_img;
        const int tpa = nynx / TS, a = ti / tpa;
        s0 = (ti - a * tpa) * TS;
        plane0 = ((size_t)(b * 3 + a) * NO) * nynx;
    };
    size_t p0; int s0;
    loc(t, p0, s0);
    prefetch<4>(in, nynx, p0, s0, t, sm);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = 0; t < total; ++k, t += gridDim.x) {
        float* tile = sm + (k & 1) * TILE;
        wait0();
        __syncthreads();
        const int tn = t + gridDim.x;
        if (tn < total) { loc(tn, p0, s0); prefetch<4>(in, nynx, p0, s0, tn, sm + ((k + 1) & 1) * TILE); }
        float* o = out + (size_t)t * TILE;
        float4 v[JN];
#pragma unroll
        for (int j = 0; j < JN; ++j) {
            const int c = 32 * (j / (QN / 8)) + lane, q = w + 8 * (j % (QN / 8));
            if (c < NO) v[j] = *reinterpret_cast<const float4*>(tile + c * TS + ((q ^ (c & 7)) << 2));
        }
        if (VAR == 0) {
#pragma unroll
            for (int j = 0; j < JN; ++j) {
                const int c = 32 * (j / (QN / 8)) + lane, q = w + 8 * (j % (QN / 8));
                if (c < NO) {
                    float* p = o + (4 * q) * NO + c;
#ifdef CS
                    asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(sig(v[j].x)) : "memory");
                    asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p + NO), "f"(sig(v[j].y)) : "memory");
                    asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p + 2 * NO), "f"(sig(v[j].z)) : "memory");
                    asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p + 3 * NO), "f"(sig(v[j].w)) : "memory");
#else
                    p[0] = sig(v[j].x); p[NO] = sig(v[j].y); p[2 * NO] = sig(v[j].z); p[3 * NO] = sig(v[j].w);
#endif
                }
            }
        } else if (VAR == 2) {
            // lin double-buffered by tile parity; the bulk store of tile k-2 must have read its buffer
            float* l2 = lin + (k & 1) * TILE;
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
#pragma unroll
            for (int j = 0; j < JN; ++j) {
                const int c = 32 * (j / (QN / 8)) + lane, q = w + 8 * (j % (QN / 8));
                if (c < NO) {
                    float* p = l2 + (4 * q) * NO + c;
                    p[0] = sig(v[j].x); p[NO] = sig(v[j].y); p[2 * NO] = sig(v[j].z); p[3 * NO] = sig(v[j].w);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             :: "l"(o), "r"((uint32_t)__cvta_generic_to_shared(l2)), "r"(TILE * 4) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
#pragma unroll
            for (int j = 0; j < JN; ++j) {
                const int c = 32 * (j / (QN / 8)) + lane, q = w + 8 * (j % (QN / 8));
                if (c < NO) {
                    float* p = lin + (4 * q) * NO + c;
                    p[0] = sig(v[j].x); p[NO] = sig(v[j].y); p[2 * NO] = sig(v[j].z); p[3 * NO] = sig(v[j].w);
                }
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < (TILE / 4 + THR - 1) / THR; ++it) {
                const int e = threadIdx.x + it * THR;
                if (e < TILE / 4) reinterpret_cast<float4*>(o)[e] = reinterpret_cast<const float4*>(lin)[e];
            }
        }
    }
    if (VAR == 2 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F> float timeit(F f, int iters = 20) {
    for (int i = 0; i < 3; ++i) f();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / iters * 1e3f;
}

int main(int argc, char** argv) {
    const int B = 64, nynx = 6400;                       // one P3-like level, 3 anchors: 300 tiles/img
    const int tiles_per_img = 3 * nynx / TS, total = B * tiles_per_img;
    const size_t n = (size_t)total * TILE;
    float *in, *out;
    cudaMalloc(&in, n * 4); cudaMalloc(&out, n * 4);
    cudaMemset(in, 0, n * 4);
    const double mb = n * 4 / 1e6;
    printf("%.1f MB each way, %d tiles\n", mb, total);
    float us = timeit([&] { e1_plain<<<(unsigned)((n / 4 + 1023) / 1024), 256>>>((const float4*)in, (float4*)out, n / 4); });
    printf("E1 plain elementwise sigmoid        %8.1f us  %7.1f GB/s (r+w)\n", us, 2 * mb / us * 1e3 / 1e3);
    const size_t smem = 2 * TILE * 4;
#define RUN(MODE, WMODE, name)                                                                     \
    for (int bps = 1; bps <= 5; ++bps) {                                                           \
        if ((size_t)bps * (smem + 1024) > 227 * 1024) break;                                       \
        cudaFuncSetAttribute(persist<MODE, WMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        float u = timeit([&] { persist<MODE, WMODE><<<148 * bps, THR, smem>>>(in, out, tiles_per_img, nynx, total); }); \
        printf("%-34s bps=%d %8.1f us  %7.1f GB/s\n", name, bps, u, (WMODE == 0 ? 2 : 1) * mb / u);  \
    }
    RUN(0, 0, "LDGSTS.32 transposing, r+w");
    RUN(1, 0, "LDGSTS.128 contiguous, r+w");
    RUN(2, 0, "LDGSTS.32 contiguous, r+w");
    RUN(3, 0, "LDG.128->STS.128 contiguous, r+w");
    RUN(0, 1, "LDGSTS.32 transposing, read only");
    RUN(1, 1, "LDGSTS.128 contiguous, read only");
    RUN(1, 2, "write only (sigmoid of smem)");
#define RUNA(VAR, name, SM)                                                                        \
    for (int bps = 1; bps <= 5; ++bps) {                                                           \
        if ((size_t)bps * ((SM) + 1024) > 227 * 1024) break;                                       \
        cudaFuncSetAttribute(persistA<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM)); \
        float u = timeit([&] { persistA<VAR><<<148 * bps, THR, (SM)>>>(in, out, tiles_per_img, nynx, total); }); \
        printf("%-34s bps=%d %8.1f us  %7.2f TB/s\n", name, bps, u, 2 * mb / u);                   \
    }
    RUNA(0, "(A) LDGSTS.128 + scalar STG", 2 * TILE * 4);
    RUNA(1, "(C) LDGSTS.128 + smem transp + STG.128", 3 * TILE * 4);
    if (TS == 64) RUNA(2, "(C') LDGSTS.128 + smem transp + bulk store", 4 * TILE * 4);
    RUN(4, 1, "LDGSTS.128 strided, read only");
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
