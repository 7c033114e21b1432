// Error reporting, launch accounting and the small element-wise entry points.
#include "vk_common.cuh"

#include <map>
#include <mutex>
#include <tuple>

namespace vk {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int fail_arg(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return VK_E_ARG;
}
int fail_code(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return VK_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static std::mutex g_attr_mutex;
static std::map<std::pair<int, const void*>, size_t> g_smem_set;                       // (device, kernel) -> limit set
static std::map<std::tuple<int, const void*, int, size_t>, int> g_occupancy;

int ensure_dyn_smem(const void* func, size_t bytes, const char* who) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    size_t& cur = g_smem_set[{dev, func}];
    if (bytes <= cur) return VK_OK;
    // (static + dynamic shared memory above 48 KB needs the opt-in even when the dynamic part alone is below it)
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail_code((int)e, "%s: %zu B of shared memory: %s", who, bytes, cudaGetErrorString(e));
    cur = bytes;
    return VK_OK;
}

int blocks_per_sm(const void* func, int threads, size_t dyn_smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    auto key = std::make_tuple(dev, func, threads, dyn_smem);
    auto it = g_occupancy.find(key);
    if (it != g_occupancy.end()) return it->second;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, func, threads, dyn_smem) != cudaSuccess || n < 1) n = 1;
    g_occupancy[key] = n;
    return n;
}

// utils/image_proc.py:76-79: coords[:, [0,2]] -= pad[0]; coords[:, [1,3]] -= pad[1];
// coords[:, :4] /= gain; clip_coords (utils/bboxes.py:50-59).  float32 ops, one rounding each.
__global__ void scale_coords_kernel(float* __restrict__ c, int n, int row_stride, float pad_w,
                                    float pad_h, float gain, int subtract_pad, float clip_w,
                                    float clip_h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 4) return;
    const int r = i >> 2, k = i & 3;
    float* p = c + (size_t)r * row_stride + k;
    float v = *p;
    if (subtract_pad) v = __fsub_rn(v, (k & 1) ? pad_h : pad_w);
    v = __fdiv_rn(v, gain);
    if (clip_w >= 0.f) {
        const float hi = (k & 1) ? clip_h : clip_w;
        v = fminf(fmaxf(v, 0.f), hi);
    }
    *p = v;
}

// utils/bboxes.py:103-111
__global__ void cxcywh_to_xyxy_kernel(const float4* __restrict__ in, float4* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 b = in[i];
    const float hw = __fmul_rn(b.z, 0.5f), hh = __fmul_rn(b.w, 0.5f);
    out[i] = make_float4(__fsub_rn(b.x, hw), __fsub_rn(b.y, hh), __fadd_rn(b.x, hw), __fadd_rn(b.y, hh));
}

}  // namespace vk

using namespace vk;

extern "C" int vk_version(void) { return 100; }
extern "C" const char* vk_last_error(void) { return g_err; }
extern "C" uint64_t vk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" int vk_build_arch(void) { return 100; }

extern "C" int vk_scale_coords(float* coords, int n, int row_stride, float pad_w, float pad_h,
                               float gain, int subtract_pad, float clip_w, float clip_h,
                               vk_stream_t stream) {
    if (n == 0) return VK_OK;
    if (!coords || n < 0 || row_stride < 4) return fail_arg("vk_scale_coords: bad argument");
    scale_coords_kernel<<<ceil_div(n * 4, 256), 256, 0, as_stream(stream)>>>(
        coords, n, row_stride, pad_w, pad_h, gain, subtract_pad, clip_w, clip_h);
    count_launch();
    return check_launch("scale_coords_kernel");
}

extern "C" int vk_cxcywh_to_xyxy(const float* in, float* out, int n, vk_stream_t stream) {
    if (n == 0) return VK_OK;
    if (!in || !out || n < 0) return fail_arg("vk_cxcywh_to_xyxy: bad argument");
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15)
        return fail_arg("vk_cxcywh_to_xyxy: pointers must be 16-byte aligned");
    cxcywh_to_xyxy_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), n);
    count_launch();
    return check_launch("cxcywh_to_xyxy_kernel");
}
