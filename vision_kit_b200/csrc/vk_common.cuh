// Shared host/device helpers of libvk_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <atomic>

#include "../../include/vk_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvk_b200 is written for sm_100a (B200) only"
#endif

namespace vk {

constexpr int kNumSMs = 148;  // B200

void set_error(const char* fmt, ...);
int fail_arg(const char* fmt, ...);        // sets message, returns VK_E_ARG
int fail_code(int code, const char* fmt, ...);
int check_launch(const char* what);        // cudaGetLastError -> return code
void count_launch(int n = 1);

inline cudaStream_t as_stream(vk_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) / occupancy queries cost microseconds of host time per
// call and would race between host threads if every launch set its own size: the library raises a
// kernel's limit only when a launch needs more than any launch before it (monotone, under a mutex) and
// remembers occupancy answers per (kernel, block size, shared memory).
int ensure_dyn_smem(const void* func, size_t bytes, const char* who);        // VK_OK or an error code
int blocks_per_sm(const void* func, int threads, size_t dyn_smem);           // cached occupancy (>= 1)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

// 128-bit streaming load that does not allocate in L1 (data is touched once).
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f32(const void* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
// element loads of the filter / decode kernels: fp16 and bf16 inputs are up-cast exactly
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <class T>
__device__ __forceinline__ float ld_elem(const T* p) { return to_f32(__ldg(p)); }

// streaming stores (evict-first: the output is not re-read by this kernel)
__device__ __forceinline__ void st_stream_f4(void* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_f32(void* p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}
// predicated 8-byte store: the predicate guards the instruction, not a branch around it
__device__ __forceinline__ void st_pred_u2(void* p, uint32_t a, uint32_t b, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q st.global.v2.u32 [%0], {%1,%2};\n\t}"
                 :: "l"(p), "r"(a), "r"(b), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ void st_stream_u2(void* p, uint2 v) {
    asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// The sigmoid every kernel of the path shares, so that the fused decode+filter and the
// two-step decode -> filter paths produce identical bits.  Four instructions: FMUL,
// MUFU.EX2, FADD, MUFU.RCP (ex2.approx <= 2 ulp, rcp.approx <= 1 ulp: relative error below
// 4e-7, inside the 1e-5 decode tolerance of SURVEY.md A.3; .ftz only touches results < 1.2e-38).
__device__ __forceinline__ float sigmoidf_vk(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(x, -1.4426950408889634f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fadd_rn(1.0f, e)));
    return r;
}

// Detect decode of one element; every product/sum rounded separately.
//   c: 0=cx 1=cy 2=w 3=h >=4: probability.  g: grid index of that axis.  anchor: pixels.
__device__ __forceinline__ float decode_elem(float logit, int c, float g, float stride,
                                             float anchor, int variant) {
    const float s = sigmoidf_vk(logit);
    if (c >= 4) return s;
    const float t = __fmul_rn(s, 2.0f);
    if (c < 2) {
        const float u = (variant == VK_HEAD_V5)
                            ? __fadd_rn(t, __fsub_rn(g, 0.5f))            // yolov5.py:68,88
                            : __fadd_rn(__fsub_rn(t, 0.5f), g);           // yolov7.py:80
        return __fmul_rn(u, stride);
    }
    return __fmul_rn(__fmul_rn(t, t), anchor);                            // yolov5.py:69
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Exclusive scan of one int per thread over a block of up to 1024 threads.
// `wsum` is shared scratch of 33 ints.  Returns the exclusive prefix; *total = block sum.
__device__ __forceinline__ int block_excl_scan(int v, int* wsum, int* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    const int inc = warp_incl_scan(v, lane);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        int x = (lane < nw) ? wsum[lane] : 0;
        int xi = warp_incl_scan(x, lane);
        wsum[lane] = xi - x;
        if (lane == 31) wsum[32] = xi;
    }
    __syncthreads();
    const int r = wsum[w] + inc - v;
    *total = wsum[32];
    __syncthreads();
    return r;
}
#endif  // __CUDACC__

}  // namespace vk
