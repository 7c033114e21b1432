// Letterbox: resize (OpenCV INTER_LINEAR 8-bit fixed point, bit-exact) + constant border +
// optional BGR->RGB + HWC->CHW + /255 + cast, one launch for a whole batch of sources of
// mixed sizes.  Replaces utils/image_proc.py:12-60 `resize`, demo/processing.py:45-52
// `preprocess` and the normalise of core/train/det_trainer.py:74-75 (SURVEY.md §8a1, a1', a2).
//
// HBM-bound integer/byte work: algorithmic bytes per image = src_h*src_w*3 read +
// 3*out_h*out_w*sizeof(out) written.  No tensor cores.
#include "vk_common.cuh"

#include <math.h>
#include <string.h>

namespace vk {

// ---------------------------------------------------------------------------------------
// coefficient tables: xtab[b][x] = {3*xs, 3*min(xs+1, sw-1), a0, a1}, ytab[b][y] =
// {r0, r1, b0, b1} for interior coordinates.  float64/float32 chain with
// every operation rounded separately, exactly as OpenCV's resize.cpp builds xofs/ialpha.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void linear_coeff(int d, int dn, int sn, bool is_x, int* s_out,
                                             int* w0, int* w1) {
    const double scale = __ddiv_rn(1.0, __ddiv_rn((double)dn, (double)sn));
    float f = __double2float_rn(__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5));
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (is_x) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= sn - 1) { s = sn - 1; f = 0.f; }
    }
    *s_out = s;
    *w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));  // cvRound: half-even
    *w1 = __float2int_rn(__fmul_rn(f, 2048.0f));
}

__global__ void __launch_bounds__(256)
lb_tables_kernel(const VkLbDesc* __restrict__ descs, int4* __restrict__ xtab,
                 int4* __restrict__ ytab, int out_h, int out_w) {
    const int b = blockIdx.x;
    const VkLbDesc d = descs[b];
    if (d.new_h == d.src_h && d.new_w == d.src_w) return;  // image_proc.py:42
    int4* xt = xtab + (size_t)b * out_w;
    int4* yt = ytab + (size_t)b * out_h;
    for (int i = threadIdx.x; i < d.new_w + d.new_h; i += blockDim.x) {
        int s, w0, w1;
        if (i < d.new_w) {
            linear_coeff(i, d.new_w, d.src_w, true, &s, &w0, &w1);
            xt[i] = make_int4(3 * s, 3 * min(s + 1, d.src_w - 1), w0, w1);
        } else {
            const int y = i - d.new_w;
            linear_coeff(y, d.new_h, d.src_h, false, &s, &w0, &w1);
            yt[y] = make_int4(min(max(s, 0), d.src_h - 1), min(max(s + 1, 0), d.src_h - 1),
                              w0, w1);
        }
    }
}

// uint8 / 255 in float32, correctly rounded (== IEEE division for all 256 inputs;
// tests/test_host_logic.py proves it exhaustively): q0 = v*r, one Newton correction.
__device__ __forceinline__ float norm255(float v) {
    const float r = 0.003921568859368563f;  // RN(1/255)
    const float q = __fmul_rn(v, r);
    const float e = __fmaf_rn(-q, 255.0f, v);
    return __fmaf_rn(e, r, q);
}

template <int FMT> struct OutT;
template <> struct OutT<VK_LB_F32_NCHW> { using type = float; };
template <> struct OutT<VK_LB_BF16_NCHW> { using type = __nv_bfloat16; };
template <> struct OutT<VK_LB_U8_NHWC> { using type = uint8_t; };

template <int FMT>
__device__ __forceinline__ void store_px(typename OutT<FMT>::type* dst, size_t plane, size_t off_chw,
                                         size_t off_hwc, int v0, int v1, int v2) {
    if constexpr (FMT == VK_LB_F32_NCHW) {
        st_stream_f32(dst + off_chw, norm255((float)v0));
        st_stream_f32(dst + off_chw + plane, norm255((float)v1));
        st_stream_f32(dst + off_chw + 2 * plane, norm255((float)v2));
    } else if constexpr (FMT == VK_LB_BF16_NCHW) {
        dst[off_chw] = __float2bfloat16_rn(norm255((float)v0));
        dst[off_chw + plane] = __float2bfloat16_rn(norm255((float)v1));
        dst[off_chw + 2 * plane] = __float2bfloat16_rn(norm255((float)v2));
    } else {
        dst[off_hwc] = (uint8_t)v0;
        dst[off_hwc + 1] = (uint8_t)v1;
        dst[off_hwc + 2] = (uint8_t)v2;
    }
}

constexpr int kLbRows = 8;      // canvas rows per block
constexpr int kLbThreads = 320; // 640-wide canvas = 2 columns per thread, no idle lanes

// One block = kLbRows canvas rows of one image.  Block-uniform choice between
//   (a) the vector copy path: no resize and everything 4-pixel aligned -> 3 x 32-bit
//       loads (12 source bytes = 4 pixels) and one 128-bit store per plane;
//   (b) the general path: one thread per canvas column, bilinear taps through L1.
template <int FMT>
__global__ void __launch_bounds__(kLbThreads)
lb_kernel(const VkLbDesc* __restrict__ descs, const int4* __restrict__ xtab,
          const int4* __restrict__ ytab, int out_h, int out_w, int swap_rb, uint32_t pad_rgb,
          typename OutT<FMT>::type* __restrict__ dst_all) {
    using T = typename OutT<FMT>::type;
    const int b = blockIdx.y;
    const int y_begin = blockIdx.x * kLbRows;
    const int y_end = min(y_begin + kLbRows, out_h);
    const VkLbDesc d = descs[b];
    const size_t plane = (size_t)out_h * out_w;
    T* dst = dst_all + (size_t)b * 3 * plane;
    const int p0 = pad_rgb & 255, p1 = (pad_rgb >> 8) & 255, p2 = (pad_rgb >> 16) & 255;
    const int c0 = swap_rb ? 2 : 0, c2 = swap_rb ? 0 : 2;  // source byte of output channel 0/2
    const bool resize = !(d.new_h == d.src_h && d.new_w == d.src_w);

    const bool vec_ok = (FMT != VK_LB_U8_NHWC) && !resize && ((d.left & 3) == 0) &&
                        ((d.new_w & 3) == 0) && ((out_w & 3) == 0) &&
                        ((reinterpret_cast<uintptr_t>(d.src) & 3) == 0) && ((d.pitch & 3) == 0) &&
                        ((reinterpret_cast<uintptr_t>(dst_all) & 15) == 0);
    if (vec_ok) {
        const int gpr = out_w >> 2;  // 4-pixel groups per row
        const int items = (y_end - y_begin) * gpr;
        float fp[3];
        fp[0] = norm255((float)p0); fp[1] = norm255((float)p1); fp[2] = norm255((float)p2);
        for (int i = threadIdx.x; i < items; i += kLbThreads) {
            const int ry = i / gpr;
            const int x = (i - ry * gpr) << 2;
            const int y = y_begin + ry;
            const int sy = y - d.top, sx = x - d.left;
            float4 o[3];
            if (sy >= 0 && sy < d.new_h && sx >= 0 && sx < d.new_w) {
                const uint8_t* p = d.src + (size_t)sy * d.pitch + (size_t)sx * 3;
                const uint32_t w0 = ld_stream_u32(p), w1 = ld_stream_u32(p + 4),
                               w2 = ld_stream_u32(p + 8);
                // byte i of the 12-byte group: pixel k, source channel j sits at i = 3k + j
                auto B = [&](int i) -> float {
                    const uint32_t w = i < 4 ? w0 : (i < 8 ? w1 : w2);
                    return norm255((float)((w >> (8 * (i & 3))) & 255u));
                };
                const float4 q0 = make_float4(B(0), B(3), B(6), B(9));
                const float4 q1 = make_float4(B(1), B(4), B(7), B(10));
                const float4 q2 = make_float4(B(2), B(5), B(8), B(11));
                o[0] = swap_rb ? q2 : q0;
                o[1] = q1;
                o[2] = swap_rb ? q0 : q2;
            } else {
                o[0] = make_float4(fp[0], fp[0], fp[0], fp[0]);
                o[1] = make_float4(fp[1], fp[1], fp[1], fp[1]);
                o[2] = make_float4(fp[2], fp[2], fp[2], fp[2]);
            }
            const size_t off = (size_t)y * out_w + x;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if constexpr (FMT == VK_LB_F32_NCHW) {
                    st_stream_f4(dst + off + c * plane, o[c]);
                } else if constexpr (FMT == VK_LB_BF16_NCHW) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(o[c].x, o[c].y);
                    __nv_bfloat162 hi = __floats2bfloat162_rn(o[c].z, o[c].w);
                    uint2 u;
                    u.x = *reinterpret_cast<uint32_t*>(&lo);
                    u.y = *reinterpret_cast<uint32_t*>(&hi);
                    st_stream_u2(dst + off + c * plane, u);
                }
            }
        }
        return;
    }

    const int4* xt = xtab + (size_t)b * out_w;
    const int4* yt = ytab + (size_t)b * out_h;
    for (int x = threadIdx.x; x < out_w; x += kLbThreads) {
        const int sx = x - d.left;
        const bool in_x = sx >= 0 && sx < d.new_w;
        int4 xc = make_int4(3 * sx, 3 * sx, 2048, 0);
        if (in_x && resize) xc = __ldg(xt + sx);
        for (int y = y_begin; y < y_end; ++y) {
            const int sy = y - d.top;
            int v0 = p0, v1 = p1, v2 = p2;
            if (in_x && sy >= 0 && sy < d.new_h) {
                if (resize) {
                    const int4 yc = __ldg(yt + sy);
                    const uint8_t* r0 = d.src + (size_t)yc.x * d.pitch;
                    const uint8_t* r1 = d.src + (size_t)yc.y * d.pitch;
                    int v[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int h0 = (int)__ldg(r0 + xc.x + c) * xc.z + (int)__ldg(r0 + xc.y + c) * xc.w;
                        const int h1 = (int)__ldg(r1 + xc.x + c) * xc.z + (int)__ldg(r1 + xc.y + c) * xc.w;
                        v[c] = ((((yc.z * (h0 >> 4)) >> 16) + ((yc.w * (h1 >> 4)) >> 16) + 2) >> 2);
                    }
                    v0 = v[c0]; v1 = v[1]; v2 = v[c2];
                } else {
                    const uint8_t* p = d.src + (size_t)sy * d.pitch + xc.x;
                    v0 = __ldg(p + c0); v1 = __ldg(p + 1); v2 = __ldg(p + c2);
                }
            }
            const size_t off = (size_t)y * out_w + x;
            store_px<FMT>(dst, plane, off, off * 3, v0, v1, v2);
        }
    }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace vk

using namespace vk;

extern "C" int vk_letterbox_geometry(int src_h, int src_w, int img_h, int img_w, int stride,
                                     int letterbox, int scaleup, int auto_, VkLbGeom* g) {
    if (!g || src_h <= 0 || src_w <= 0 || img_h <= 0 || img_w <= 0 || stride <= 0)
        return fail_arg("vk_letterbox_geometry: bad size");
    // utils/image_proc.py:27-31 -- int / int true division, float64
    double ratio = fmin((double)img_h / (double)src_h, (double)img_w / (double)src_w);
    if (!scaleup) ratio = fmin(ratio, 1.0);
    // :34-35 -- int(round(x)): round-half-even == nearbyint in the default rounding mode
    const int new_w = (int)nearbyint((double)src_w * ratio);
    const int new_h = (int)nearbyint((double)src_h * ratio);
    long dwi = img_w - new_w, dhi = img_h - new_h;  // :36-37
    if (auto_) {                                     // :38-40 np.mod: result has the sign of stride
        dwi = ((dwi % stride) + stride) % stride;
        dhi = ((dhi % stride) + stride) % stride;
    }
    double dw = (double)dwi, dh = (double)dhi;
    int top = 0, left = 0;
    int bottom = (int)nearbyint(dh), right = (int)nearbyint(dw);  // :46-47
    if (letterbox) {                                              // :49-53
        dw = dw / 2.0;
        dh = dh / 2.0;
        top = (int)nearbyint(dh - 0.1);
        bottom = (int)nearbyint(dh + 0.1);
        left = (int)nearbyint(dw - 0.1);
        right = (int)nearbyint(dw + 0.1);
    }
    memset(g, 0, sizeof(*g));
    g->ratio = ratio;
    g->pad_w = dw;
    g->pad_h = dh;
    g->new_w = new_w;
    g->new_h = new_h;
    g->top = top; g->bottom = bottom; g->left = left; g->right = right;
    g->out_h = new_h + top + bottom;
    g->out_w = new_w + left + right;
    g->needs_resize = !(new_w == src_w && new_h == src_h);
    if (new_w <= 0 || new_h <= 0) return fail_arg("vk_letterbox_geometry: degenerate target %dx%d", new_w, new_h);
    return VK_OK;
}

static size_t lb_desc_bytes(int batch) { return align_up((size_t)batch * sizeof(VkLbDesc), 256); }

extern "C" size_t vk_letterbox_workspace_bytes(int batch, int out_h, int out_w) {
    if (batch <= 0 || out_h <= 0 || out_w <= 0) return 0;
    return lb_desc_bytes(batch) + (size_t)batch * ((size_t)out_w + out_h) * sizeof(int4);
}

extern "C" int vk_letterbox_batch(const VkLbDesc* descs_host, const VkLbDesc* descs_dev, int batch,
                                  int out_h, int out_w, int swap_rb, uint32_t pad_rgb, int dst_fmt,
                                  void* dst, void* ws, size_t ws_bytes, vk_stream_t stream_) {
    if (batch == 0) return VK_OK;
    if (!descs_host || !dst || batch < 0 || out_h <= 0 || out_w <= 0)
        return fail_arg("vk_letterbox_batch: null/negative argument");
    if (batch > 65535) return fail_code(VK_E_LIMIT, "vk_letterbox_batch: batch %d > 65535", batch);
    if (dst_fmt < 0 || dst_fmt > 2) return fail_arg("vk_letterbox_batch: dst_fmt %d", dst_fmt);
    if (!ws || ws_bytes < vk_letterbox_workspace_bytes(batch, out_h, out_w))
        return fail_code(VK_E_WORKSPACE, "vk_letterbox_batch: workspace %zu < %zu", ws_bytes,
                         vk_letterbox_workspace_bytes(batch, out_h, out_w));
    bool any_resize = false;
    for (int i = 0; i < batch; ++i) {
        const VkLbDesc& d = descs_host[i];
        if (!d.src || d.src_h <= 0 || d.src_w <= 0 || d.new_h <= 0 || d.new_w <= 0 ||
            d.pitch < (int64_t)d.src_w * 3 || d.top < 0 || d.left < 0 ||
            d.top + d.new_h > out_h || d.left + d.new_w > out_w)
            return fail_arg("vk_letterbox_batch: descriptor %d does not fit the %dx%d canvas", i, out_h, out_w);
        any_resize |= !(d.new_h == d.src_h && d.new_w == d.src_w);
    }
    cudaStream_t stream = as_stream(stream_);
    char* w = static_cast<char*>(ws);
    const VkLbDesc* dd = descs_dev;
    if (!dd) {
        cudaError_t e = cudaMemcpyAsync(w, descs_host, (size_t)batch * sizeof(VkLbDesc),
                                        cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return fail_code((int)e, "vk_letterbox_batch: descriptor upload: %s", cudaGetErrorString(e));
        dd = reinterpret_cast<const VkLbDesc*>(w);
    }
    int4* xtab = reinterpret_cast<int4*>(w + lb_desc_bytes(batch));
    int4* ytab = xtab + (size_t)batch * out_w;
    if (any_resize) {
        lb_tables_kernel<<<batch, 256, 0, stream>>>(dd, xtab, ytab, out_h, out_w);
        count_launch();
        if (int rc = check_launch("lb_tables_kernel")) return rc;
    }
    dim3 grid(ceil_div(out_h, kLbRows), batch);
    switch (dst_fmt) {
        case VK_LB_F32_NCHW:
            lb_kernel<VK_LB_F32_NCHW><<<grid, kLbThreads, 0, stream>>>(
                dd, xtab, ytab, out_h, out_w, swap_rb, pad_rgb, static_cast<float*>(dst));
            break;
        case VK_LB_BF16_NCHW:
            lb_kernel<VK_LB_BF16_NCHW><<<grid, kLbThreads, 0, stream>>>(
                dd, xtab, ytab, out_h, out_w, swap_rb, pad_rgb, static_cast<__nv_bfloat16*>(dst));
            break;
        default:
            lb_kernel<VK_LB_U8_NHWC><<<grid, kLbThreads, 0, stream>>>(
                dd, xtab, ytab, out_h, out_w, swap_rb, pad_rgb, static_cast<uint8_t*>(dst));
    }
    count_launch();
    return check_launch("lb_kernel");
}
