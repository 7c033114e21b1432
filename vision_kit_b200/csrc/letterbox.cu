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
// coefficient tables: xtab[b][x] = {3*xs, a0 | a1 << 16} (the second tap is the next pixel,
// clamped to the last one, where OpenCV gives it weight 0), ytab[b][y] = {r0, r1, b0, b1} for
// interior coordinates.  float64/float32 chain with every operation rounded
// separately, exactly as OpenCV's resize.cpp builds xofs/ialpha.
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
lb_tables_kernel(const VkLbDesc* __restrict__ descs, int2* __restrict__ xtab,
                 int4* __restrict__ ytab, int4* __restrict__ ttab, int out_h, int out_w, int tile_rows) {
    const int b = blockIdx.x;
    const VkLbDesc d = descs[b];
    int2* xt = xtab + (size_t)b * out_w;
    int4* yt = ytab + (size_t)b * out_h;
    for (int i = threadIdx.x; i < d.new_w + d.new_h; i += blockDim.x) {
        int s, w0, w1;
        if (i < d.new_w) {
            linear_coeff(i, d.new_w, d.src_w, true, &s, &w0, &w1);
            xt[i] = make_int2(3 * s, w0 | (w1 << 16));
        } else {
            const int y = i - d.new_w;
            linear_coeff(y, d.new_h, d.src_h, false, &s, &w0, &w1);
            yt[y] = make_int4(min(max(s, 0), d.src_h - 1), min(max(s + 1, 0), d.src_h - 1),
                              w0, w1);
        }
    }
    __syncthreads();
    // ttab[b][g] = {first source row, number of source rows, first and one-past-last canvas row inside the image}
    // of the tile of `tile_rows` canvas rows starting at g * tile_rows
    const int tiles_y = (out_h + tile_rows - 1) / tile_rows;
    for (int g = threadIdx.x; g < tiles_y; g += blockDim.x) {
        const int y0 = g * tile_rows, y1 = min(y0 + tile_rows, out_h);
        const int ys = max(y0, d.top), ye = min(y1, d.top + d.new_h);
        int rlo = 0, nsrc = 0;
        if (ys < ye) {
            rlo = yt[ys - d.top].x;
            nsrc = yt[ye - 1 - d.top].y - rlo + 1;
        }
        ttab[(size_t)b * tiles_y + g] = make_int4(rlo, nsrc, ys, ye);
    }
}

// uint8 / 255 in float32, correctly rounded (== IEEE division for all 256 inputs;
// tests/test_host_logic.py proves it exhaustively): 1/255 split in two floats, v*hi + fl(v*lo) in one FMA.
__device__ __forceinline__ float norm255(float v) {
    const float hi = 0.003921568859368563f;   // RN(1/255)
    const float lo = -2.319175823606301e-10f;   // RN(1/255 - hi)
    return __fmaf_rn(v, hi, __fmul_rn(v, lo));
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

// Horizontal pass of one source row for one canvas column: h[c] = P0[c]*a0 + P1[c]*a1 where the
// two source pixels are the 6 bytes starting `sh/8` bytes into the aligned words w0 w1 w2 and
// a01 = a0 | a1 << 16.  PRMT pairs (P0[c], P1[c]) as adjacent bytes, DP2A does the two products.
__device__ __forceinline__ void taps_dp2a(uint32_t w0, uint32_t w1, uint32_t w2, int sh, uint32_t a01, uint32_t h[3]) {
    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
    const uint32_t p01 = __byte_perm(lo, hi, 0x4130);   // [P0c0, P1c0, P0c1, P1c1]
    const uint32_t p2 = __byte_perm(lo, hi, 0x0052);    // [P0c2, P1c2, -, -]
    h[0] = __dp2a_lo(a01, p01, 0u);
    h[1] = __dp2a_hi(a01, p01, 0u);
    h[2] = __dp2a_lo(a01, p2, 0u);
}

#ifndef VK_LB_UNR
#define VK_LB_UNR 2
#endif
#ifndef VK_LB_MINB
#define VK_LB_MINB 3
#endif
constexpr int kLbRows = 8;        // canvas rows per block
constexpr int kLbThreads = 320;   // 640-wide canvas = 2 columns per thread, no idle lanes
#ifndef VK_GEN_ROWS
#define VK_GEN_ROWS 8
#endif
#ifndef VK_GEN_RING_KB
#define VK_GEN_RING_KB 96
#endif
#ifndef VK_GEN_BPS
#define VK_GEN_BPS 2
#endif
#ifndef VK_GEN_UNROLL
#define VK_GEN_UNROLL 1
#endif

constexpr int kGenRows = VK_GEN_ROWS;           // canvas rows per tile of the general kernel
constexpr int kGenCols = 160;                   // 4-column groups per pass: a 640-wide canvas row
constexpr int kGenThreads = 2 * kGenCols;       // two row sets (even / odd rows of the tile)
constexpr int kGenUnroll = VK_GEN_UNROLL;
constexpr int kGenQ = 4;                        // tiles a block keeps in flight
constexpr int kGenRing = VK_GEN_RING_KB * 1024; // staging ring per block (two blocks per SM)

// ---------------------------------------------------------------------------------------
// copy kernel: every image of the batch is a plain copy (no resize) with 4-pixel alignment.
// 3 x 32-bit loads = 4 pixels, byte extract, /255, one 128-bit streaming store per plane.
// ---------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(kLbThreads, (FMT == VK_LB_BF16_NCHW) ? (VK_LB_MINB > 4 ? 4 : VK_LB_MINB) : VK_LB_MINB)
lb_copy_kernel(const VkLbDesc* __restrict__ descs, int out_h, int out_w, int swap_rb, uint32_t pad_rgb,
               typename OutT<FMT>::type* __restrict__ dst_all) {
    using T = typename OutT<FMT>::type;
    // pixels per thread: every plane store is 128 bits (4 floats or 8 bfloat16)
    constexpr int PX = (FMT == VK_LB_BF16_NCHW) ? 8 : 4;
    constexpr int NW = PX * 3 / 4;               // source words per group
    const int b = blockIdx.y;
    const int y_begin = blockIdx.x * kLbRows;
    const int y_end = min(y_begin + kLbRows, out_h);
    const VkLbDesc d = descs[b];
    const size_t plane = (size_t)out_h * out_w;
    T* dst = dst_all + (size_t)b * 3 * plane;
    const int p0 = pad_rgb & 255, p1 = (pad_rgb >> 8) & 255, p2 = (pad_rgb >> 16) & 255;
    const int gpr = out_w / PX;                  // groups per row
    const int items = (y_end - y_begin) * gpr;
    // pad value per SOURCE channel slot (the store swaps planes for BGR input)
    const float fq0 = norm255((float)(swap_rb ? p2 : p0)), fq1 = norm255((float)p1), fq2 = norm255((float)(swap_rb ? p0 : p2));
    // UNR items per pass: all their loads are issued before the first conversion, so a thread keeps
    // UNR x 12 (24) bytes of reads and UNR x 3 plane stores in flight
    constexpr int UNR = VK_LB_UNR;
    for (int i0 = threadIdx.x; i0 < items; i0 += UNR * kLbThreads) {
        uint32_t w[UNR][NW];
        bool in[UNR];
        size_t off[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int i = i0 + u * kLbThreads;
            const int ry = i / gpr;
            const int x = (i - ry * gpr) * PX;
            const int y = y_begin + ry;
            const int sy = y - d.top, sx = x - d.left;
            off[u] = (size_t)y * out_w + x;
            in[u] = i < items && sy >= 0 && sy < d.new_h && sx >= 0 && sx < d.new_w;
            if (in[u]) {
                const uint8_t* p = d.src + (size_t)sy * d.pitch + (size_t)sx * 3;
#pragma unroll
                for (int k = 0; k < NW; ++k) w[u][k] = ld_stream_u32(p + 4 * k);
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if (i0 + u * kLbThreads >= items) break;
            float o[3][PX];
            if (in[u]) {
#pragma unroll
                for (int k = 0; k < PX; ++k) {       // byte 3k + j = pixel k, source channel j
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const int bi = 3 * k + j;
                        o[j][k] = norm255((float)((w[u][bi >> 2] >> (8 * (bi & 3))) & 255u));   // source channel order; the swap happens at the store
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < PX; ++k) { o[0][k] = fq0; o[1][k] = fq1; o[2][k] = fq2; }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                T* const pl = dst + off[u] + (size_t)((swap_rb && c != 1) ? 2 - c : c) * plane;   // BGR->RGB = plane swap
                if constexpr (FMT == VK_LB_F32_NCHW) {
                    st_stream_f4(pl, make_float4(o[c][0], o[c][1], o[c][2], o[c][3]));
                } else if constexpr (FMT == VK_LB_BF16_NCHW) {
                    uint32_t uu[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(o[c][2 * k], o[c][2 * k + 1]);
                        uu[k] = *reinterpret_cast<uint32_t*>(&h2);
                    }
                    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};"
                                 :: "l"(pl), "r"(uu[0]), "r"(uu[1]), "r"(uu[2]), "r"(uu[3]) : "memory");
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// general kernel (any source size, alignment and pitch; images without resize use it with identity tables,
// weights 2048/0 reproduce the source byte exactly).
//
// Persistent blocks, each walking tiles of kGenRows canvas rows of one image (tile t = row group t / B of
// image t % B, so that a block sees all the batch's sizes).  The source rows a tile needs are staged in a
// shared-memory RING: a tile takes exactly the bytes it needs, and a block keeps issuing the asynchronous
// copies of later tiles (up to kGenQ, as far as the ring has room) while it evaluates the oldest one.  What
// bounds this kernel is the number of bytes in flight per SM; with one fixed worst-case staging buffer per
// block (the round-1 kernel) a mixed batch kept ~27 KB per SM in flight and ran at 0.4-0.6 of the copy peak.
//
//   1. stage: 16-byte LDGSTS chunks; a staged row keeps its source alignment modulo 16 so that aligned
//      global chunks are aligned shared chunks.  DRAM sees only coalesced loads; nothing is read before
//      the image's first aligned word or past its last byte.
//   2. per canvas column the two horizontal taps are 6 consecutive bytes of a staged row: 2-3 LDS.32, funnel
//      shift, PRMT to pair the bytes, DP2A for P0*a0 + P1*a1.  Clamped taps at the right edge have weight 0
//      (OpenCV resets the fraction there), so every column takes this path.
//   3. vertical pass with IMAD.HI on weights pre-shifted by 16, value/255 by the two-FMA division.
//
// Tiles whose source span exceeds the ring (down-scales beyond ~2.5x of 1280-wide sources) take byte taps
// through L1 instead.
// ---------------------------------------------------------------------------------------
struct GenTile {
    int b, y_begin, nrows;
    int ys, ye;            // canvas rows of the tile that intersect the image: [ys, ye)
    int rlo, nsrc;         // source rows they touch: [rlo, rlo + nsrc)
    int stride;            // bytes per staged row (multiple of 16)
    int bytes;             // nsrc * stride
    int off;               // ring offset of the first staged row, -1 = not staged (direct taps)
    int acct;              // ring bytes the tile gives back (its own + the tail it skipped when wrapping)
    int left, new_w;       // interior columns on the canvas: [left, left + new_w)
};

template <int FMT>
__global__ void __launch_bounds__(kGenThreads, VK_GEN_BPS)
lb_general_kernel(const VkLbDesc* __restrict__ descs, const int2* __restrict__ xtab,
                  const int4* __restrict__ ytab, const int4* __restrict__ ttab, int batch, int out_h, int out_w,
                  int swap_rb, uint32_t pad_rgb, typename OutT<FMT>::type* __restrict__ dst_all) {
    using T = typename OutT<FMT>::type;
    extern __shared__ __align__(16) uint8_t ring[];
    __shared__ GenTile s_q[kGenQ];
    __shared__ int4 s_y[kGenQ][kGenRows];        // {ring byte offset of row r0's first pixel, of r1's, b0<<16, b1<<16}
    __shared__ float s_lut[256];
    const int tid = threadIdx.x;
    const int tiles_y = (out_h + kGenRows - 1) / kGenRows, total = tiles_y * batch;
    const size_t plane = (size_t)out_h * out_w;
    const int p0 = pad_rgb & 255, p1 = (pad_rgb >> 8) & 255, p2 = (pad_rgb >> 16) & 255;
    const int c0 = swap_rb ? 2 : 0, c2 = swap_rb ? 0 : 2;  // source byte of output channel 0/2
    const int pa = swap_rb ? p2 : p0, pc = swap_rb ? p0 : p2;                    // pad value per source byte
    for (int i = tid; i < 256; i += kGenThreads) s_lut[i] = norm255((float)i);

    // tile t = row group t % tiles_y of image t / tiles_y; its table entry and descriptor are loaded one tile ahead
    auto describe = [&](int b, int g, const int4 ti, const VkLbDesc& d) {
        GenTile q;
        q.b = b;
        q.y_begin = g * kGenRows;
        q.nrows = min(kGenRows, out_h - q.y_begin);
        q.rlo = ti.x; q.nsrc = ti.y; q.ys = ti.z; q.ye = ti.w;
        q.stride = (3 * d.src_w + 36 + 15) & ~15;
        q.bytes = q.nsrc * q.stride;
        q.off = 0; q.acct = 0;
        q.left = d.left; q.new_w = d.new_w;
        return q;
    };

    // asynchronous copies of a tile's source rows into the ring, and its row table
    auto issue = [&](const GenTile& q, const VkLbDesc& d, int slot) {
        if (tid == 0) s_q[slot] = q;
        if (tid < kGenRows) {
            const int y = q.y_begin + tid;
            int4 e = make_int4(-1, -1, 0, 0);
            if (q.off >= 0 && y >= q.ys && y < q.ye) {
                const int4 yc = __ldg(ytab + (size_t)q.b * out_h + (y - d.top));
                // byte position of a row's first pixel inside its staged row: its address mod 16
                const int a0 = (int)(reinterpret_cast<uintptr_t>(d.src + (size_t)yc.x * d.pitch) & 15);
                const int a1 = (int)(reinterpret_cast<uintptr_t>(d.src + (size_t)yc.y * d.pitch) & 15);
                e = make_int4(q.off + (yc.x - q.rlo) * q.stride + a0, q.off + (yc.y - q.rlo) * q.stride + a1,
                              yc.z << 16, yc.w << 16);
            }
            s_y[slot][tid] = e;
        }
        if (q.off < 0) return;
        const uint8_t* img_end = d.src + (size_t)(d.src_h - 1) * d.pitch + (size_t)3 * d.src_w;
        const int lane = tid & 31;
        for (int j = tid >> 5; j < q.nsrc; j += kGenThreads / 32) {   // one warp per source row
            const uint8_t* g = d.src + (size_t)(q.rlo + j) * d.pitch;
            const int a = (int)(reinterpret_cast<uintptr_t>(g) & 3);
            const uint8_t* ga = g - a;                                // first aligned word of the row
            const int k = (int)((reinterpret_cast<uintptr_t>(ga) >> 2) & 3);
            const int nwj = (a + 3 * d.src_w + 3) >> 2;               // aligned words holding the row
            uint32_t* srow = reinterpret_cast<uint32_t*>(ring + q.off + j * q.stride);
            const bool last_row = (q.rlo + j == d.src_h - 1);         // only there a word can cross img_end
            const uint8_t* gk = ga - 4 * k;                           // global address of staging word 0 (16-byte aligned)
            const int c_first = k ? 1 : 0;                            // chunks [c_first, c_full) are whole and inside the row
            int c_full = (k + nwj) >> 2;
            if (last_row && c_full > c_first && gk + 16 * c_full > img_end) --c_full;
            const unsigned sa0 = (unsigned)__cvta_generic_to_shared(srow);
            for (int c = c_first + lane; c < c_full; c += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa0 + 16u * c), "l"(gk + 16 * c) : "memory");
            // the ragged ends: words [k, 4) of chunk 0, the words after the last whole chunk, three zero words (taps read past the row)
            if (lane < 11) {
                int w;                                                // staging word index
                if (lane < 4) w = lane;                               // head chunk
                else w = 4 * c_full + (lane - 4);                     // up to 7 tail words (one dropped chunk + a partial one)
                const bool head_ok = lane < 4 && k && w >= k && w < k + nwj;
                const bool tail_ok = lane >= 4 && w >= k && w < k + nwj && (c_full >= c_first);
                if (head_ok || tail_ok) {
                    const uint8_t* wp = gk + 4 * w;
                    if (!last_row || wp + 4 <= img_end) {
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sa0 + 4u * w), "l"(wp) : "memory");
                    } else {                                          // the image's very last word: no over-read
                        uint32_t v = 0;
                        for (int r = 0; r < 4 && wp + r < img_end; ++r) v |= (uint32_t)__ldg(wp + r) << (8 * r);
                        srow[w] = v;
                    }
                }
            } else if (lane < 14) {
                srow[k + nwj + (lane - 11)] = 0u;
            }
        }
    };

    // (explicit 32-bit shared addresses: through generic pointers the compiler rebuilds the shared window base,
    // an S2R, inside the row loop)
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    auto pixel = [&](const int4 ye, int tb, uint32_t a01, int& va, int& vb, int& vc) {
        const uint32_t t0 = ring_s + (uint32_t)(ye.x + tb), t1 = ring_s + (uint32_t)(ye.y + tb);   // byte addresses of the taps
        uint32_t u0, u1, u2, w0, w1, w2;
        asm volatile("ld.shared.u32 %0, [%3];\n\tld.shared.u32 %1, [%3+4];\n\tld.shared.u32 %2, [%3+8];"
                     : "=r"(u0), "=r"(u1), "=r"(u2) : "r"(t0 & ~3u));
        asm volatile("ld.shared.u32 %0, [%3];\n\tld.shared.u32 %1, [%3+4];\n\tld.shared.u32 %2, [%3+8];"
                     : "=r"(w0), "=r"(w1), "=r"(w2) : "r"(t1 & ~3u));
        uint32_t h0[3], h1[3];
        taps_dp2a(u0, u1, u2, (int)(t0 << 3), a01, h0);               // the funnel shift takes the amount modulo 32
        taps_dp2a(w0, w1, w2, (int)(t1 << 3), a01, h1);
        // ((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2 >> 2
        va = (int)((__umulhi((uint32_t)ye.z, h0[0] >> 4) + __umulhi((uint32_t)ye.w, h1[0] >> 4) + 2u) >> 2);
        vb = (int)((__umulhi((uint32_t)ye.z, h0[1] >> 4) + __umulhi((uint32_t)ye.w, h1[1] >> 4) + 2u) >> 2);
        vc = (int)((__umulhi((uint32_t)ye.z, h0[2] >> 4) + __umulhi((uint32_t)ye.w, h1[2] >> 4) + 2u) >> 2);
    };
    const uint32_t sy_s = (uint32_t)__cvta_generic_to_shared(&s_y[0][0]);
    auto row_entry = [&](int slot, int r) {
        int4 e;
        asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w)
                     : "r"(sy_s + 16u * (uint32_t)(slot * kGenRows + r)));
        return e;
    };

    const bool vec4 = ((out_w & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst_all) & 15) == 0);
    // column coefficients of the thread's first 4-column group, kept across the tiles of one image
    int cb = -1, ctb[4] = {0, 0, 0, 0};
    uint32_t ca01[4] = {0, 0, 0, 0};
    bool cin[4] = {false, false, false, false};
    auto compute = [&](const GenTile& q, int slot) {
        T* dst = dst_all + (size_t)q.b * 3 * plane;
        const int2* xt = xtab + (size_t)q.b * out_w;
        if (q.off < 0) {
            const VkLbDesc d = descs[q.b];
            // ---- byte taps through L1 (source span larger than the ring)
            const int4* yt = ytab + (size_t)q.b * out_h;
            for (int x = tid; x < out_w; x += kGenThreads) {
                const int sx = x - d.left;
                const bool in_x = sx >= 0 && sx < d.new_w;
                const int2 xc = in_x ? __ldg(xt + sx) : make_int2(0, 0);
                const int x0 = xc.x, x1 = min(xc.x + 3, 3 * (d.src_w - 1)), w0 = xc.y & 0xffff, w1 = (int)((uint32_t)xc.y >> 16);
                for (int y = q.y_begin; y < q.y_begin + q.nrows; ++y) {
                    const int sy = y - d.top;
                    int v0 = p0, v1 = p1, v2 = p2;
                    if (in_x && sy >= 0 && sy < d.new_h) {
                        const int4 yc = __ldg(yt + sy);
                        const uint8_t* r0 = d.src + (size_t)yc.x * d.pitch;
                        const uint8_t* r1 = d.src + (size_t)yc.y * d.pitch;
                        int v[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int h0 = (int)__ldg(r0 + x0 + c) * w0 + (int)__ldg(r0 + x1 + c) * w1;
                            const int h1 = (int)__ldg(r1 + x0 + c) * w0 + (int)__ldg(r1 + x1 + c) * w1;
                            v[c] = ((((yc.z * (h0 >> 4)) >> 16) + ((yc.w * (h1 >> 4)) >> 16) + 2) >> 2);
                        }
                        v0 = v[c0]; v1 = v[1]; v2 = v[c2];
                    }
                    const size_t off = (size_t)y * out_w + x;
                    if constexpr (FMT == VK_LB_F32_NCHW) {
                        st_stream_f32(dst + off, s_lut[v0]);
                        st_stream_f32(dst + off + plane, s_lut[v1]);
                        st_stream_f32(dst + off + 2 * plane, s_lut[v2]);
                    } else if constexpr (FMT == VK_LB_BF16_NCHW) {
                        dst[off] = __float2bfloat16_rn(s_lut[v0]);
                        dst[off + plane] = __float2bfloat16_rn(s_lut[v1]);
                        dst[off + 2 * plane] = __float2bfloat16_rn(s_lut[v2]);
                    } else {
                        dst[off * 3] = (uint8_t)v0; dst[off * 3 + 1] = (uint8_t)v1; dst[off * 3 + 2] = (uint8_t)v2;
                    }
                }
            }
            return;
        }
        // ---- taps from the ring.  A thread owns 4 adjacent canvas columns and every other row of the tile:
        // the three plane stores of a row are 128 bits (64 for bf16) and their address arithmetic is shared by
        // 4 pixels; the column coefficients are loaded once per tile.  BGR<->RGB is a swap of plane pointers.
        if (vec4) {
            const size_t oa = (FMT == VK_LB_U8_NHWC) ? 0 : (size_t)c0 * plane;       // plane of source byte 0
            const size_t ob = (FMT == VK_LB_U8_NHWC) ? 0 : plane;
            const size_t oc = (FMT == VK_LB_U8_NHWC) ? 0 : (size_t)c2 * plane;
            const int rsel = tid / kGenCols, xfirst = (tid - rsel * kGenCols) * 4;
            auto load_cols = [&](int x4, int* tb, uint32_t* a01, bool* in_x) {
                const int sx0 = x4 - q.left;
                if (sx0 >= 0 && sx0 + 3 < q.new_w && ((sx0 & 1) == 0)) {               // two 16-byte loads
                    const int4 u = __ldg(reinterpret_cast<const int4*>(xt + sx0)), v = __ldg(reinterpret_cast<const int4*>(xt + sx0 + 2));
                    tb[0] = u.x; a01[0] = (uint32_t)u.y; tb[1] = u.z; a01[1] = (uint32_t)u.w;
                    tb[2] = v.x; a01[2] = (uint32_t)v.y; tb[3] = v.z; a01[3] = (uint32_t)v.w;
                    in_x[0] = in_x[1] = in_x[2] = in_x[3] = true;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int sx = sx0 + j;
                        in_x[j] = sx >= 0 && sx < q.new_w;
                        const int2 xc = in_x[j] ? __ldg(xt + sx) : make_int2(0, 0);
                        tb[j] = xc.x;
                        a01[j] = (uint32_t)xc.y;
                    }
                }
            };
            if (q.b != cb) {
                cb = q.b;
                if (xfirst < out_w) load_cols(xfirst, ctb, ca01, cin);
            }
            for (int x4 = xfirst; x4 < out_w; x4 += kGenCols * 4) {
                int tb[4];
                uint32_t a01[4];
                bool in_x[4];
                if (x4 == xfirst) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { tb[j] = ctb[j]; a01[j] = ca01[j]; in_x[j] = cin[j]; }
                } else {
                    load_cols(x4, tb, a01, in_x);
                }
                const bool all_in = in_x[0] && in_x[1] && in_x[2] && in_x[3];
                uint32_t idx = (uint32_t)((q.y_begin + rsel) * out_w + x4);
#pragma unroll kGenUnroll
                for (int r = rsel; r < q.nrows; r += 2, idx += 2 * out_w) {
                    const int4 ye = row_entry(slot, r);
                    int va[4], vb[4], vc[4];
                    if (all_in && ye.x >= 0) {                         // interior: no per-pixel predicates
#pragma unroll
                        for (int j = 0; j < 4; ++j) pixel(ye, tb[j], a01[j], va[j], vb[j], vc[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            va[j] = pa; vb[j] = p1; vc[j] = pc;
                            if (in_x[j] && ye.x >= 0) pixel(ye, tb[j], a01[j], va[j], vb[j], vc[j]);
                        }
                    }
                    if constexpr (FMT == VK_LB_F32_NCHW) {
                        // value/255 by arithmetic here: the table would cost 12 more shared-memory
                        // wavefronts per thread-row
                        st_stream_f4(dst + oa + idx, make_float4(norm255((float)va[0]), norm255((float)va[1]), norm255((float)va[2]), norm255((float)va[3])));
                        st_stream_f4(dst + ob + idx, make_float4(norm255((float)vb[0]), norm255((float)vb[1]), norm255((float)vb[2]), norm255((float)vb[3])));
                        st_stream_f4(dst + oc + idx, make_float4(norm255((float)vc[0]), norm255((float)vc[1]), norm255((float)vc[2]), norm255((float)vc[3])));
                    } else if constexpr (FMT == VK_LB_BF16_NCHW) {
                        auto pack = [&](const int* v) {
                            __nv_bfloat162 lo = __floats2bfloat162_rn(norm255((float)v[0]), norm255((float)v[1]));
                            __nv_bfloat162 hi = __floats2bfloat162_rn(norm255((float)v[2]), norm255((float)v[3]));
                            uint2 u;
                            u.x = *reinterpret_cast<uint32_t*>(&lo);
                            u.y = *reinterpret_cast<uint32_t*>(&hi);
                            return u;
                        };
                        st_stream_u2(dst + oa + idx, pack(va));
                        st_stream_u2(dst + ob + idx, pack(vb));
                        st_stream_u2(dst + oc + idx, pack(vc));
                    } else {
                        // 4 pixels x 3 interleaved bytes = three aligned 32-bit words
                        int v[12];
#pragma unroll
                        for (int j = 0; j < 4; ++j) { v[3 * j + c0] = va[j]; v[3 * j + 1] = vb[j]; v[3 * j + c2] = vc[j]; }
                        uint32_t* wp = reinterpret_cast<uint32_t*>(dst + (size_t)idx * 3);
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            wp[k] = (uint32_t)v[4 * k] | ((uint32_t)v[4 * k + 1] << 8) | ((uint32_t)v[4 * k + 2] << 16) |
                                    ((uint32_t)v[4 * k + 3] << 24);
                    }
                }
            }
        } else {
            T* const plane_a = dst + (FMT == VK_LB_U8_NHWC ? c0 : (size_t)c0 * plane);   // receives source byte 0
            T* const plane_b = dst + (FMT == VK_LB_U8_NHWC ? 1 : plane);
            T* const plane_c = dst + (FMT == VK_LB_U8_NHWC ? c2 : (size_t)c2 * plane);   // receives source byte 2
            for (int x = tid; x < out_w; x += kGenThreads) {
                const int sx = x - q.left;
                const bool in_x = sx >= 0 && sx < q.new_w;
                const int2 xc = in_x ? __ldg(xt + sx) : make_int2(0, 0);
                uint32_t idx = (uint32_t)(q.y_begin * out_w + x);
                for (int r = 0; r < q.nrows; ++r, idx += out_w) {
                    const int4 ye = row_entry(slot, r);
                    int va = pa, vb = p1, vc = pc;
                    if (in_x && ye.x >= 0) pixel(ye, xc.x, (uint32_t)xc.y, va, vb, vc);
                    if constexpr (FMT == VK_LB_F32_NCHW) {
                        st_stream_f32(plane_a + idx, s_lut[va]);
                        st_stream_f32(plane_b + idx, s_lut[vb]);
                        st_stream_f32(plane_c + idx, s_lut[vc]);
                    } else if constexpr (FMT == VK_LB_BF16_NCHW) {
                        plane_a[idx] = __float2bfloat16_rn(s_lut[va]);
                        plane_b[idx] = __float2bfloat16_rn(s_lut[vb]);
                        plane_c[idx] = __float2bfloat16_rn(s_lut[vc]);
                    } else {
                        plane_a[3 * idx] = (uint8_t)va; plane_b[3 * idx] = (uint8_t)vb; plane_c[3 * idx] = (uint8_t)vc;
                    }
                }
            }
        }
    };

    // ---- the pipeline.  Tile t = row group t / B of image t % B, block k takes t = k, k + grid, ...: every block
    // sees all the batch's sizes.  h = next tile to issue, c = next tile to evaluate.  One barrier per tile: the
    // bytes of tile c - 1 go back to the ring after the barrier that precedes the evaluation of tile c.
    const int t_begin = blockIdx.x, t_end = total, t_step = gridDim.x;
    const int step_g = t_step / batch, step_b = t_step - step_g * batch;
    int hg = t_begin / batch, hb = t_begin - hg * batch;               // row group and image of tile h
    int h = t_begin, c = t_begin, nq = 0, qh = 0, qc = 0;
    int wr = 0, room = kGenRing, pend = 0;                             // ring write offset, free bytes, bytes to give back
    int4 nti = make_int4(0, 0, 0, 0);                                  // table entry and descriptor of tile h
    VkLbDesc nd = {};
    auto look_ahead = [&]() {
        if (h < t_end) { nti = __ldg(ttab + (size_t)hb * tiles_y + hg); nd = descs[hb]; }
    };
    auto issue_some = [&]() {
        while (h < t_end && nq < kGenQ) {
            GenTile q = describe(hb, hg, nti, nd);
            if (q.bytes > kGenRing) {
                q.off = -1;
            } else {
                int off = wr, acct = q.bytes;
                if (wr + q.bytes > kGenRing) { acct += kGenRing - wr; off = 0; }    // skip the tail, start over
                if (acct > room) break;                                // wait for older tiles to leave
                q.off = off; q.acct = acct;
                wr = off + q.bytes; room -= acct;
            }
            issue(q, nd, qh);
            asm volatile("cp.async.commit_group;" ::: "memory");
            qh = (qh + 1) % kGenQ; ++nq; h += t_step;
            hg += step_g; hb += step_b;
            if (hb >= batch) { hb -= batch; ++hg; }
            look_ahead();
        }
    };
    look_ahead();
    while (c < t_end) {
        if (nq == 0) {                                                 // nothing in flight: the ring starts over
            __syncthreads();
            room += pend; pend = 0; wr = 0;
            issue_some();
        }
        switch (nq - 1) {                                              // groups newer than the oldest tile may stay pending
            case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
            case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
            case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
            default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        }
        __syncthreads();                                               // tile c has landed, everyone is done with tile c - 1
        room += pend; pend = 0;
        issue_some();
        const GenTile q = s_q[qc];
        compute(q, qc);
        pend = q.acct;
        qc = (qc + 1) % kGenQ; --nq; c += t_step;
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

// Eval-time ingest: data/datasets/yolo.py:144-160 `load_resized_image`
// (longest side -> max(img_sz), int() truncation, INTER_LINEAR) followed by the validation
// pipeline's albumentations PadIfNeeded (data/augmentations.py:197-200: centre position,
// pad_before = int((min - size) / 2.0)).  Same VkLbGeom as the letterbox, so the same kernels run.
extern "C" int vk_dataset_geometry(int src_h, int src_w, int img_h, int img_w, VkLbGeom* g) {
    if (!g || src_h <= 0 || src_w <= 0 || img_h <= 0 || img_w <= 0)
        return fail_arg("vk_dataset_geometry: bad size");
    const int m_img = img_h > img_w ? img_h : img_w, m_src = src_h > src_w ? src_h : src_w;
    const double r = (double)m_img / (double)m_src;                  // yolo.py:154
    int new_w = src_w, new_h = src_h;
    if (r != 1.0) {                                                  // :155-157
        new_w = (int)((double)src_w * r);
        new_h = (int)((double)src_h * r);
    }
    if (new_w <= 0 || new_h <= 0) return fail_arg("vk_dataset_geometry: degenerate target %dx%d", new_w, new_h);
    if (new_h > img_h || new_w > img_w)
        return fail_arg("vk_dataset_geometry: resized %dx%d exceeds the %dx%d canvas (PadIfNeeded never crops)",
                        new_h, new_w, img_h, img_w);
    const int top = (int)((double)(img_h - new_h) / 2.0), left = (int)((double)(img_w - new_w) / 2.0);
    memset(g, 0, sizeof(*g));
    g->ratio = r;
    g->pad_w = (double)left;
    g->pad_h = (double)top;
    g->new_w = new_w; g->new_h = new_h;
    g->top = top; g->bottom = img_h - new_h - top; g->left = left; g->right = img_w - new_w - left;
    g->out_h = img_h; g->out_w = img_w;
    g->needs_resize = !(new_w == src_w && new_h == src_h);
    return VK_OK;
}

static size_t lb_desc_bytes(int batch) { return align_up((size_t)batch * sizeof(VkLbDesc), 256); }

extern "C" size_t vk_letterbox_workspace_bytes(int batch, int out_h, int out_w) {
    if (batch <= 0 || out_h <= 0 || out_w <= 0) return 0;
    return lb_desc_bytes(batch) + align_up((size_t)batch * out_w * sizeof(int2), 16) + (size_t)batch * out_h * sizeof(int4) +
           (size_t)batch * ceil_div(out_h, kGenRows) * sizeof(int4);
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
    const int px = (dst_fmt == VK_LB_BF16_NCHW) ? 8 : 4;     // pixels per 128-bit plane store of the copy kernel
    bool all_copy = (dst_fmt != VK_LB_U8_NHWC) && (out_w % px == 0) &&
                    ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    for (int i = 0; i < batch; ++i) {
        const VkLbDesc& d = descs_host[i];
        if (!d.src || d.src_h <= 0 || d.src_w <= 0 || d.new_h <= 0 || d.new_w <= 0 ||
            d.pitch < (int64_t)d.src_w * 3 || d.top < 0 || d.left < 0 ||
            d.top + d.new_h > out_h || d.left + d.new_w > out_w)
            return fail_arg("vk_letterbox_batch: descriptor %d does not fit the %dx%d canvas", i, out_h, out_w);
        const bool rs = !(d.new_h == d.src_h && d.new_w == d.src_w);
        all_copy &= !rs && (d.left % px == 0) && (d.new_w % px == 0) &&
                    ((reinterpret_cast<uintptr_t>(d.src) & 3) == 0) && ((d.pitch & 3) == 0);
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
    int2* xtab = reinterpret_cast<int2*>(w + lb_desc_bytes(batch));                  // 16-byte aligned rows when out_w is even
    int4* ytab = reinterpret_cast<int4*>(w + lb_desc_bytes(batch) + align_up((size_t)batch * out_w * sizeof(int2), 16));
    int4* ttab = ytab + (size_t)batch * out_h;
    dim3 grid(ceil_div(out_h, kLbRows), batch);
    if (all_copy) {
        if (dst_fmt == VK_LB_F32_NCHW)
            lb_copy_kernel<VK_LB_F32_NCHW><<<grid, kLbThreads, 0, stream>>>(dd, out_h, out_w, swap_rb, pad_rgb, static_cast<float*>(dst));
        else
            lb_copy_kernel<VK_LB_BF16_NCHW><<<grid, kLbThreads, 0, stream>>>(dd, out_h, out_w, swap_rb, pad_rgb, static_cast<__nv_bfloat16*>(dst));
        count_launch();
        return check_launch("lb_copy_kernel");
    }
    lb_tables_kernel<<<batch, 256, 0, stream>>>(dd, xtab, ytab, ttab, out_h, out_w, kGenRows);
    count_launch();
    if (int rc = check_launch("lb_tables_kernel")) return rc;
    const int tiles = ceil_div(out_h, kGenRows) * batch;
#define VK_LB_LAUNCH(FMT, T)                                                                              \
    do {                                                                                                   \
        const void* fn = reinterpret_cast<const void*>(&lb_general_kernel<FMT>);                           \
        if (int rc = ensure_dyn_smem(fn, kGenRing, "vk_letterbox_batch")) return rc;                       \
        const int bps = blocks_per_sm(fn, kGenThreads, kGenRing);                                          \
        const int nblk = tiles < bps * kNumSMs ? tiles : bps * kNumSMs;                                   \
        lb_general_kernel<FMT><<<nblk, kGenThreads, kGenRing, stream>>>(dd, xtab, ytab, ttab, batch, out_h, out_w, swap_rb, pad_rgb, \
                                                                       static_cast<T*>(dst));              \
    } while (0)
    switch (dst_fmt) {
        case VK_LB_F32_NCHW: VK_LB_LAUNCH(VK_LB_F32_NCHW, float); break;
        case VK_LB_BF16_NCHW: VK_LB_LAUNCH(VK_LB_BF16_NCHW, __nv_bfloat16); break;
        default: VK_LB_LAUNCH(VK_LB_U8_NHWC, uint8_t);
    }
#undef VK_LB_LAUNCH
    count_launch();
    return check_launch("lb_general_kernel");
}
