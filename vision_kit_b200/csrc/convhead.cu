// Fused Detect conv head: vk_conv_decode_filter and its tcgen05 kernels.
#include "decode_common.cuh"

#include <cuda.h>      // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

// =======================================================================================
// Fused Detect head: 1x1 conv (tensor cores) + decode + confidence filter  
//
//   models/heads/yolov5.py:58-69 / yolov7.py:67-81: x[i] = m[i](x[i]) (a 1x1 conv, Cout = na*no),
//   then sigmoid / grid / anchors, then nms()'s candidate selection (utils/image_proc.py:99-151).
//   Here the (B, 255, ny, nx) conv output never exists: one CTA computes
//       D[s][co] = sum_ci X[b][ci][s] * W[co][ci]        s: 128 spatial positions, co: 256 (padded)
//   with tcgen05.mma (kind::tf32, fp32 bits of X and W used as TF32, fp32 accumulate in TMEM).
//   With D transposed like this a TMEM lane is a spatial position, so every epilogue thread owns
//   whole prediction rows (one per anchor) and runs the dense filter's per-row logic straight
//   from tcgen05.ld -- no shared-memory round trip of the logits.
//
//   Operands, both straight from global memory by 16-byte cp.async copies:
//     W[co][ci] is K-major as stored: rows of 32 tf32 = 128 B, 8-row atoms of 1024 B, SWIZZLE_128B.
//     X is [ci][s], s contiguous = MN-major for the A operand.  tcgen05 takes MN-major TF32 only in the
//     SWIZZLE_128B_BASE32B layout (layout type 1; every other layout type yields zeros or an illegal instruction --
//     profiles/micro/umma_probe.cu reads the mapping back from the tensor core): atoms of 4 k-rows x 32 positions
//     (4 x 128 B), the 32-byte chunks of row k at position chunk ^ (k & 3); position blocks LBO apart, k-groups SBO
//     apart.  A 16-byte copy of X[k][m..m+3] lands where the MMA expects it: no transposition on chip.
//   Pipeline per 32-channel block: the copies of block k+1 (X and W) are in flight while block k is multiplied; both
//   operand tiles are double-buffered behind the MMA-completion mbarrier.
//   128 threads, 256 TMEM columns, 96 KB shared memory and no static allocation -> 2 CTAs per SM.
//
//   Candidates, boxes, segment table: exactly vk_decode_filter's format (same VkCandBuf), so
//   vk_nms_batched consumes it unchanged.  Results equal conv-then-vk_decode_filter up to TF32
//   rounding of the logits (tests/test_gpu_parity.py::test_conv_head_*).
// =======================================================================================
namespace vk {

constexpr int kChM = 128, kChN = 256, kChKB = 32, kChThreads = 128;
constexpr int kChABytes = kChM * 128;                 // 16 KB: [8 k-groups][4 position blocks][4 k x 128 B]
constexpr int kChBBytes = kChN * 128;                 // 32 KB
constexpr int kChTail = 64;                           // mbarrier, TMEM base, warp totals
constexpr int kChSmem = 2 * kChABytes + 2 * kChBBytes + kChTail;   // 96 KB + 64 B, no static shared memory: 2 CTAs per SM
constexpr uint32_t kChALbo = 512, kChASbo = 2048;     // A tile: position blocks 512 B apart, k-groups (4 channels) 2048 B apart

struct ConvHead {
    const float* x[VK_MAX_LEVELS];      // (B, cin, ny, nx)
    const float* w[VK_MAX_LEVELS];      // (na*no, cin)
    const float* bias[VK_MAX_LEVELS];   // (na*no) or null
    int cin[VK_MAX_LEVELS];
    int mtile_start[VK_MAX_LEVELS + 1]; // first 128-row tile of each level inside an image
    int mtiles;                         // per image
    int ptile_start[VK_MAX_LEVELS + 1]; // the same in 256-position pair tiles (persistent kernel)
    int ptiles;
};

__device__ __forceinline__ uint32_t ch_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major SW128 tile: row r = 128 B (32 tf32), 16-byte chunk c of row r at position c ^ (r & 7)
__device__ __forceinline__ uint32_t ch_koff(int r, int chunk) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
}
__device__ __forceinline__ uint64_t ch_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;              // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)(1024 >> 4) << 32;    // stride byte offset: 8-row atoms 1024 B apart
    d |= (uint64_t)1 << 46;              // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;              // SWIZZLE_128B
    return d;
}
// MN-major SW128_BASE32B tile of A: element (m, k) of a 32-channel block
__device__ __forceinline__ uint32_t ch_aoff(int m, int k) {
    return (uint32_t)((m >> 5) * kChALbo + (k >> 2) * kChASbo + (k & 3) * 128 + ((((m & 31) >> 2) ^ ((k & 3) << 1)) << 4) + (m & 3) * 4);
}
__device__ __forceinline__ uint64_t ch_desc_a(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)(kChALbo >> 4) << 16;
    d |= (uint64_t)(kChASbo >> 4) << 32;
    d |= (uint64_t)1 << 46;              // descriptor version (sm_100)
    d |= (uint64_t)1 << 61;              // SWIZZLE_128B_BASE32B
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, A MN-major, B K-major, N = 256, M = 128
constexpr uint32_t kChIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(kChN >> 3) << 17) | ((uint32_t)(kChM >> 4) << 24);
// A tile of the bias step (K = 8): A[m][0] = A[m][1] = 1, everything else 0 (rows of equal values: the swizzle does not move them)
__device__ __forceinline__ void ch_bias_a(uint8_t* a_tile, int t128) {
#pragma unroll
    for (int e = t128; e < 256; e += 128) {                // 2 k-groups x 4 position blocks x 4 rows x 8 float4
        const int g = e >> 7, k = (e >> 3) & 3;
        const float v = (g == 0 && k < 2) ? 1.0f : 0.0f;
        *reinterpret_cast<float4*>(a_tile + g * kChASbo + ((e >> 5) & 3) * kChALbo + (e & 31) * 16) = make_float4(v, v, v, v);
    }
}
__device__ __forceinline__ void ch_cp16(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;        // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(n) : "memory");
}
// bounded wait on an mbarrier phase; returns false on time-out (never hangs the GPU)
__device__ __forceinline__ bool ch_wait(uint32_t bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 24); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}

#define VK_TMEM_LD16(r, taddr)                                                                                     \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),  \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
                 : "r"(taddr));                                                                                    \
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

// 32 / 64 consecutive columns of the calling warp's 32 lanes (no wait: the caller waits once for all its loads)
__device__ __forceinline__ void ch_tmem_ld32(uint32_t* r, uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void ch_tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Epilogue of one 128-position tile whose accumulator sits in 256 TMEM columns at `trow` (lane
// offset of the calling warp included): thread `tid` (0..127) = spatial position s0 + tid = TMEM
// lane, one prediction row per anchor.  `bar_id` names a 128-thread barrier shared by the 4 warps
// that run it, `s_wtot` 4 ints of their scratch.  Anchors a_first, a_first + a_step, ... are handled.
__device__ __forceinline__ void ch_bar_sync(int id) { asm volatile("bar.sync %0, 128;" :: "r"(id) : "memory"); }

__device__ __forceinline__ void conv_epilogue(const HeadDev& H, const FilterArgs& A, uint32_t trow, int l, int b, int s0,
                                              int nvalid, int tid, int* s_wtot, int bar_id, bool ok,
                                              int a_first = 0, int a_step = 1) {
    const int lane = tid & 31, warp = tid >> 5;
    const int nynx = H.nynx[l], no = H.no;
    const int nc = A.nc;
    const int sp = s0 + tid;
    const int half = tid >> 6;                                 // 64-row segment of the thread inside the tile
    const bool seg_exists = s0 + 64 * half < nynx;
    struct { int variant, nx; float stride; } const gbase{H.variant, H.nx[l], H.stride[l]};
    for (int a = a_first; a < H.na; a += a_step) {            // (the persistent kernel splits the anchors over two teams)
        const int cb = a * no;
        uint32_t r[16];
        VK_TMEM_LD16(r, trow + cb);                            // x, y, w, h, obj, first 11 classes
        float box_l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) box_l[j] = __uint_as_float(r[j]);
        const float o = sigmoidf_vk(__uint_as_float(r[4]));
        const bool alive = ok && tid < nvalid && o > A.conf;       // image_proc.py:99
        const float obj = alive ? o : 0.0f;
        // pass 1: count (multi-label) or best class.  The sigmoid is monotone, so the 2 MUFU + 4 FP32
        // instructions per class are spent only where the outcome is open: multi-label evaluates
        // p = sigmoid(x) * obj exactly only for logits above logit(conf / obj) - 0.05 (everything below
        // cannot pass p > conf); best-class finds the largest logit first and evaluates the products
        // only within 1e-3 of it (first maximum of the PRODUCTS, as the reference takes it).
        int count = 0;
        float bv = -INFINITY;
        int bj = 0x7fffffff;
        const bool any_alive = __any_sync(0xffffffffu, alive);
        // logit(conf / obj): conf/obj in (0,1) for alive rows; log via MUFU, generous margin below
        float tau = INFINITY;
        if (alive) {
            const float rr = __fdividef(A.conf, obj);
            tau = (rr > 0.f) ? __logf(__fdividef(rr, 1.0f - rr)) - 0.05f : -INFINITY;
        }
        if (any_alive) {
            if (A.multi_label) {
                for (int c0 = 0; c0 < nc; c0 += 16) {
                    uint32_t q[16];
                    VK_TMEM_LD16(q, trow + cb + 5 + c0);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int c = c0 + j;
                        const float x = __uint_as_float(q[j]);
                        if (c < nc && x > tau) {
                            const float p = __fmul_rn(sigmoidf_vk(x), obj);                            // :135
                            count += (p > A.conf && class_allowed(A.class_mask, c)) ? 1 : 0;            // :141,151
                        }
                    }
                }
            } else {
                // Best class = first maximum of the PRODUCTS p = sigmoid(x) * obj (image_proc.py:145).  ONE sweep over
                // the 80 class logits (three 32-column loads: the epilogue is bound by TMEM round trips) keeps
                // the two largest logits and the first index of the largest.  Runner-up more than 1e-3 below (almost
                // always): the arg-max of the logits is the answer and only its product is evaluated.  Otherwise,
                // anywhere in the warp (tcgen05.ld is warp-collective): the products of all classes within 1e-3 of the
                // maximum are compared, as before.
                // (four independent accumulators over interleaved classes, merged at the end)
                float m1, m2;
                {
                    float a1[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, a2[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                    int aj[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
                    // columns cb+5 .. cb+84 in chunks [0,32) [32,64) [48,80): the last one overlaps (class 79 is the
                    // anchor's last column; reading 32 more from class 64 on would leave the accumulator for a = 2).
                    // One chunk at a time: 96 live registers spill in this kernel.
                    uint32_t q[32];
                    auto take = [&](uint32_t u, int c, int k) {
                        const float x = __uint_as_float(u);
                        a2[k] = fmaxf(a2[k], fminf(a1[k], x));
                        aj[k] = x > a1[k] ? c : aj[k];
                        a1[k] = fmaxf(a1[k], x);
                    };
                    auto sweep = [&](const uint32_t* q, int first, int lo, int hi) {   // classes first + j inside [lo, hi)
                        if (first >= lo && first + 32 <= hi) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) take(q[j], first + j, j & 3);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (first + j >= lo && first + j < hi) take(q[j], first + j, j & 3);
                        }
                    };
                    ch_tmem_ld32(q, trow + cb + 5);
                    ch_tmem_wait();
                    sweep(q, 0, 0, nc);
                    if (nc > 32) {
                        ch_tmem_ld32(q, trow + cb + 5 + 32);
                        ch_tmem_wait();
                        sweep(q, 32, 32, nc);
                    }
                    if (nc > 64) {
                        ch_tmem_ld32(q, trow + cb + 5 + nc - 32);
                        ch_tmem_wait();
                        sweep(q, nc - 32, 64, nc);
                    }
                    // merge: largest, runner-up (the smaller of two maxima is a runner-up candidate), index of the largest
                    auto merge = [&](int x, int y) {
                        a2[x] = fmaxf(fmaxf(a2[x], a2[y]), fminf(a1[x], a1[y]));
                        aj[x] = a1[y] > a1[x] ? aj[y] : aj[x];
                        a1[x] = fmaxf(a1[x], a1[y]);
                    };
                    merge(0, 1); merge(2, 3); merge(0, 2);
                    m1 = a1[0]; m2 = a2[0]; bj = aj[0];
                }
                const float near = m1 - 1e-3f;
                if (__any_sync(0xffffffffu, alive && m2 >= near)) {
                    bj = 0x7fffffff;
                    for (int c0 = 0; c0 < nc; c0 += 16) {
                        uint32_t q[16];
                        VK_TMEM_LD16(q, trow + cb + 5 + c0);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float x = __uint_as_float(q[j]);
                            if (c0 + j < nc && x >= near) {
                                const float p = __fmul_rn(sigmoidf_vk(x), obj);                            // :135
                                if (p > bv) { bv = p; bj = c0 + j; }                                        // :145
                            }
                        }
                    }
                } else {
                    bv = __fmul_rn(sigmoidf_vk(m1), obj);                                                  // :135
                }
                count = (alive && bj != 0x7fffffff && bv > A.conf && class_allowed(A.class_mask, bj)) ? 1 : 0;
            }
        }
        // offsets inside the 64-row segment (warps 2*half, 2*half+1), canonical order = row order
        const int incl = warp_incl_scan(count, lane);
        if (lane == 31) s_wtot[warp] = incl;
        ch_bar_sync(bar_id);
        const int base = (warp & 1) ? s_wtot[warp - 1] : 0;
        const int seg_total = s_wtot[2 * half] + s_wtot[2 * half + 1];
        const int seg = H.tile_start[l] + a * H.tpa[l] + (s0 >> 6) + half;
        const int row = H.row_base[l] + a * nynx + sp;             // prediction row inside the image
        // pass 2 (warp-uniform: tcgen05.ld is a warp-collective): recompute the products and store
        uint2* wp = reinterpret_cast<uint2*>(A.cand + (size_t)b * A.cap) + (size_t)seg * A.tile_cap + base + (incl - count);
        const uint32_t idx0 = (uint32_t)(row * nc);
        if (A.multi_label) {
            if (__any_sync(0xffffffffu, count > 0)) {
                for (int c0 = 0; c0 < nc; c0 += 16) {
                    uint32_t q[16];
                    VK_TMEM_LD16(q, trow + cb + 5 + c0);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int c = c0 + j;
                        if (c < nc && count > 0 && __uint_as_float(q[j]) > tau) {
                            const float p = __fmul_rn(sigmoidf_vk(__uint_as_float(q[j])), obj);
                            if (p > A.conf && class_allowed(A.class_mask, c)) *wp++ = make_uint2(__float_as_uint(p), idx0 + (uint32_t)c);
                        }
                    }
                }
            }
        } else if (count > 0) {
            *wp = make_uint2(__float_as_uint(bv), idx0 + (uint32_t)bj);
        }
        if (count > 0) {
            const int gy = sp / gbase.nx, gx = sp - gy * gbase.nx;
            const float aw = H.anchors[l][2 * a], ah = H.anchors[l][2 * a + 1];
            A.boxes[(size_t)b * A.rows + row] =
                xyxy_from_cxcywh(decode_elem(box_l[0], 0, (float)gx, gbase.stride, aw, gbase.variant),
                                 decode_elem(box_l[1], 1, (float)gy, gbase.stride, ah, gbase.variant),
                                 decode_elem(box_l[2], 2, 0.f, gbase.stride, aw, gbase.variant),
                                 decode_elem(box_l[3], 3, 0.f, gbase.stride, ah, gbase.variant));
        }
        if ((tid & 63) == 0 && seg_exists) {
            A.seg_count[(size_t)b * A.segs + seg] = seg_total;
            if (seg == 0) A.flags[b] = cand_flags(A);
            if (seg_total) atomicAdd(A.counts + b, seg_total);
        }
        ch_bar_sync(bar_id);                                       // s_wtot is reused by the next anchor
    }
}

__global__ void __launch_bounds__(kChThreads, 2)
conv_decode_filter_kernel(const HeadDev H, const ConvHead C, const FilterArgs A, int* __restrict__ fault) {
    // no static shared memory: the dynamic window then starts 1024-byte aligned (the swizzled tiles need it)
    // and 2 x (96 KB + 64 B + 1 KB reserved) fits one SM
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                                      // 2 x 16 KB
    uint8_t* sB = sA + 2 * kChABytes;                        // 2 x 32 KB
    uint64_t& s_bar = *reinterpret_cast<uint64_t*>(sB + 2 * kChBBytes);
    uint32_t& s_tmem = *reinterpret_cast<uint32_t*>(sB + 2 * kChBBytes + 8);
    int* s_wtot = reinterpret_cast<int*>(sB + 2 * kChBBytes + 16);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    int l = 0;
#pragma unroll
    for (int i = 1; i < VK_MAX_LEVELS; ++i)
        if (i < H.nl && (int)blockIdx.x >= C.mtile_start[i]) l = i;
    const int s0 = ((int)blockIdx.x - C.mtile_start[l]) * kChM;
    const int nynx = H.nynx[l], cin = C.cin[l], no = H.no, cout = H.na * no;
    const int nvalid = min(kChM, nynx - s0);
    const float* __restrict__ X = C.x[l] + (size_t)b * cin * nynx + s0;
    const float* __restrict__ W = C.w[l];
    const int nkb = cin / kChKB;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(ch_smem(&s_tmem)), "n"(kChN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(ch_smem(&s_bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const float* __restrict__ bias = C.bias[l];

    // copy loops: thread-constant parts of the addresses hoisted (element e = tid + i * 128)
    const int xk0 = tid >> 5, xm4 = (tid & 31) << 2;                       // X: channel xk0 + 4 i, positions xm4..+3
    const bool xvalid = xm4 < nvalid;
    const float* const xsrc0 = X + (size_t)xk0 * nynx + (xvalid ? xm4 : 0);
    const uint32_t xdst0 = ch_aoff(xm4, xk0);                              // + i * kChASbo: channel xk0 + 4 i is row xk0 of k-group i
    const int wn0 = tid >> 3, wc = tid & 7;                                // W: row wn0 + 16 i, chunk wc
    const uint32_t wdst0 = ch_koff(wn0, wc);                               // (n & 7) does not change with i: + i * 2048
    auto issue_x = [&](int kb) {
        // X block: 32 channels x 32 chunks of 4 positions, straight into the MN-major tile; positions past the plane are zeros
        const uint32_t sa = ch_smem(sA + (kb & 1) * kChABytes) + xdst0;
        const float* src = xsrc0 + (size_t)kb * kChKB * nynx;
#pragma unroll
        for (int i = 0; i < 8; ++i) ch_cp16(sa + i * kChASbo, src + (size_t)(4 * i) * nynx, xvalid);
    };
    auto issue_w = [&](int kb) {
        // W block: 256 rows x 8 chunks of 4 channels, straight into the swizzled K-major tile
        const uint32_t sb = ch_smem(sB + (kb & 1) * kChBBytes) + wdst0;
        const float* src = W + kb * kChKB + 4 * wc;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int n = wn0 + 16 * i;
            ch_cp16(sb + i * 2048, src + (size_t)(n < cout ? n : 0) * cin, n < cout);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue_x(0);
    issue_w(0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    bool ok = true;
    // one copy group per block (X then W): at block kb everything but the newest group must have landed
    for (int kb = 0; kb < nkb; ++kb) {
        // the MMAs of block kb-1 read the operand buffers that block kb+1 reuses
        if (kb > 0) ok &= ch_wait(ch_smem(&s_bar), (uint32_t)((kb - 1) & 1));
        if (kb + 1 < nkb) {
            issue_x(kb + 1);
            issue_w(kb + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes (cp.async) -> async proxy
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                      // block kb landed for everyone
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = ch_smem(sA + (kb & 1) * kChABytes), b0 = ch_smem(sB + (kb & 1) * kChBBytes);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {                  // UMMA_K = 8 tf32: two k-groups of A, 32 bytes inside W's 128-byte rows
                const uint32_t acc = (kb | ks) ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             :: "r"(tmem), "l"(ch_desc_a(a0 + ks * 2 * kChASbo)), "l"(ch_desc(b0 + ks * 32)), "r"(kChIdesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(ch_smem(&s_bar)) : "memory");
        }
    }
    // The bias rides on the tensor cores as one more K = 8 step: A gets the columns [1, 1, 0, ...], W the
    // columns [hi, lo, 0, ...] with bias = hi + lo split into two TF32 values (fp32-accurate sum), built in
    // the operand buffers that block nkb would use (both free: their last readers have completed).
    int last_phase = (nkb - 1) & 1;
    if (bias != nullptr) {
        // a parity wait only distinguishes adjacent phases: pass the last block's phase before the
        // bias step can complete the next one
        ok &= ch_wait(ch_smem(&s_bar), (uint32_t)last_phase);
        uint8_t* ea = sA + (nkb & 1) * kChABytes;
        uint8_t* eb = sB + (nkb & 1) * kChBBytes;
        ch_bias_a(ea, tid);
        for (int n = tid; n < kChN; n += kChThreads) {
            const float v = n < cout ? __ldg(bias + n) : 0.0f;
            const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            *reinterpret_cast<float4*>(eb + ch_koff(n, 0)) = make_float4(hi, __fsub_rn(v, hi), 0.f, 0.f);
            *reinterpret_cast<float4*>(eb + ch_koff(n, 1)) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         :: "r"(tmem), "l"(ch_desc_a(ch_smem(ea))), "l"(ch_desc(ch_smem(eb))), "r"(kChIdesc), "r"(1u) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(ch_smem(&s_bar)) : "memory");
        }
        last_phase = nkb & 1;
    }
    ok &= ch_wait(ch_smem(&s_bar), (uint32_t)last_phase);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok && tid == 0) atomicExch(fault, 1);

    // ---------------- epilogue: thread = spatial position s0 + tid = TMEM lane; one prediction row per anchor
    conv_epilogue(H, A, tmem + ((uint32_t)(warp * 32) << 16), l, b, s0, nvalid, tid, s_wtot, 0, ok);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(kChN));
}

// ---------------------------------------------------------------------------------------
// Warp-specialised persistent variant on CTA PAIRS (cluster of 2, tcgen05 cta_group::2), operands by TMA.
// What bounds this kernel is the traffic that fills shared memory: every 128-position tile needs all of W
// (256 x cin), twice the bytes of its own X block.  A pair computes two tiles with one M = 256 instruction stream:
// each CTA stages the X block of ITS tile and only ITS HALF of W (128 of the 256 output channels); the tensor cores
// of both SMs read both halves.  W traffic per tile is halved.
//   warp 16     producer, one lane: per 32-channel block four cp.async.bulk.tensor boxes of X ([32 k][32 s], swizzle
//               128B_ATOM_32B = the MN-major BASE32B operand layout, profiles/micro/tma_probe.cu) and one of W
//               ([128 n][32 k], swizzle 128B), 32 KB on one transaction barrier; kWsSlots-deep ring; positions past
//               a plane and output channels past na*no arrive as zeros (TMA out-of-bounds fill)
//   warp 17     forwarder, one lane: "this CTA's block landed" -> arrive on the LEADER's READY barrier
//   warp 18     (leader CTA only) MMA issuer, one lane: tcgen05.mma.cta_group::2 into one of TWO 256-column
//               accumulators (each CTA's TMEM receives its own tile's 128 rows); the bias as one more K = 8 step from
//               constant tiles built at kernel start; tcgen05.commit multicast to both CTAs -> "slot free" and
//               "accumulator full" mbarriers
//   warps 0-15  two epilogue groups (one per accumulator) of two 4-warp teams (alternate anchors) draining tile i while
//               the mainloop of tile i+1 runs; "accumulator empty" is signalled on the leader's barrier by both CTAs
// A pair's two tiles are the two halves of 256 consecutive positions of one level (the last pair of a level may have
// an empty second half: its X is zeros and its epilogue writes nothing).
// Every mbarrier phase is waited in order by its consumer (a parity wait cannot tell phase n from
// n + 2); every wait is bounded and an abort flag stops all roles if one ever times out.
// ---------------------------------------------------------------------------------------
constexpr int kWsThreads = 19 * 32;
constexpr int kWsSlots = 6;                                          // ring depth of the operand tiles
constexpr int kWsBBytes = 128 * 128;                                 // this CTA's half of a W block: 128 rows x 128 B
constexpr int kWsOffB = kWsSlots * kChABytes;                        // 96 KB
constexpr int kWsOffBiasA = kWsOffB + kWsSlots * kWsBBytes;          // 192 KB: A tile of the bias step (4 KB)
constexpr int kWsOffBiasB = kWsOffBiasA + 4096;                      // W tile of the bias step: level l in the columns 8 l .. 8 l + 7
constexpr int kWsOffBar = kWsOffBiasB + kWsBBytes;
static_assert(VK_MAX_LEVELS * 8 <= kChKB, "bias columns of all levels fit one 32-column tile");
constexpr int kWsSmem = kWsOffBar + 320;                               // barriers (192 B), TMEM base, abort flag, 64 B of scan scratch
constexpr uint32_t kWsStageTx = kChABytes + kWsBBytes;               // bytes one block brings: 4 X boxes + 1 W box
// instruction descriptor of the pair: M = 256 (128 rows per CTA), N = 256, A MN-major, B K-major
constexpr uint32_t kWsIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(kChN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
// A tile as TMA writes it: one box = 32 positions x 32 channels = 8 k-groups of 512 B; the 4 position blocks 4096 B apart
__device__ __forceinline__ uint64_t ch_desc_a_tma(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)(4096 >> 4) << 16;    // leading byte offset: position blocks
    d |= (uint64_t)(512 >> 4) << 32;     // stride byte offset: k-groups
    d |= (uint64_t)1 << 46;              // descriptor version (sm_100)
    d |= (uint64_t)1 << 61;              // SWIZZLE_128B_BASE32B
    return d;
}

struct ConvMaps {                        // one tensor map per level: X as (s, b * cin + ci), W as (ci, co)
    CUtensorMap x[VK_MAX_LEVELS], w[VK_MAX_LEVELS];
};

struct WsTile {
    int valid, b, l, s0, nvalid, nkb;
};

// pair-tile tp (256 positions of one level of one image), this CTA's half
__device__ __forceinline__ WsTile ws_tile(const HeadDev& H, const ConvHead& C, int tp, int total, int rank) {
    WsTile q;
    q.valid = tp < total;
    if (!q.valid) { q.b = q.l = q.s0 = q.nvalid = q.nkb = 0; return q; }
    q.b = tp / C.ptiles;
    const int pt = tp - q.b * C.ptiles;
    int l = 0;
#pragma unroll
    for (int i = 1; i < VK_MAX_LEVELS; ++i)
        if (i < H.nl && pt >= C.ptile_start[i]) l = i;
    q.l = l;
    q.s0 = ((pt - C.ptile_start[l]) * 2 + rank) * kChM;
    q.nvalid = max(0, min(kChM, H.nynx[l] - q.s0));
    q.nkb = C.cin[l] / kChKB;
    return q;
}

// arrive on the barrier at the same offset in the leader CTA (rank 0) of the pair.  Default (CTA-scope) semantics on
// both sides, as CUTLASS's ClusterBarrier does: the operands are read by the tensor cores through the async proxy,
// never through the leader's L1; cluster-scope acquire / release compiled to CCTL.IVALL + MEMBAR.ALL.GPU on every
// wait and arrive (225 -> 172 us per 64 images in the LDGSTS version of this kernel).
__device__ __forceinline__ void ws_arrive_leader(uint32_t bar) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(0));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(remote) : "memory");
}
__device__ __forceinline__ void ws_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// POLL = true: mbarrier.test_wait in a tight loop (single lanes whose hand-offs happen once per 0.3 us k-block);
// false: try_wait, which lets the hardware suspend the (many) waiting threads.
template <bool POLL>
__device__ __forceinline__ bool ws_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
    for (int spin = 0; spin < (POLL ? (1 << 26) : (1 << 22)); ++spin) {
        uint32_t done;
        if (POLL)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return true;
        if ((spin & 1023) == 1023 && *abort_flag) return false;
    }
    *abort_flag = 1;
    return false;
}
__device__ __forceinline__ void ws_tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWsThreads, 1)
conv_decode_filter_ws_kernel(const HeadDev H, const ConvHead C, const FilterArgs A, const __grid_constant__ ConvMaps M,
                             int total_pairs, int* __restrict__ fault) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + kWsOffB;
    uint8_t* sBiasA = smem + kWsOffBiasA;
    uint8_t* sBiasB = smem + kWsOffBiasB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWsOffBar);   // full[S] ready[S] done[S] tfull[2] tempty[2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kWsOffBar + 192);
    volatile int* s_abort = reinterpret_cast<volatile int*>(smem + kWsOffBar + 196);
    int* s_wtot = reinterpret_cast<int*>(smem + kWsOffBar + 208);     // [2 accumulators][2 teams][4]
    const uint32_t bar0 = ch_smem(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };                  // own: this CTA's copies of the slot landed (transaction count)
    auto READY = [&](int s) { return bar0 + 8u * (kWsSlots + s); };    // leader: both CTAs' operands of the slot landed
    auto DONE = [&](int s) { return bar0 + 8u * (2 * kWsSlots + s); }; // both: the MMAs reading the slot completed
    auto TFULL = [&](int g) { return bar0 + 8u * (3 * kWsSlots + g); };        // both: accumulator g holds a finished tile
    auto TEMPTY = [&](int g) { return bar0 + 8u * (3 * kWsSlots + 2 + g); };   // leader: both CTAs drained accumulator g
    static_assert(8 * (3 * kWsSlots + 4) <= 192, "barrier block");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cout = H.na * H.no;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (warp == 18) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(ch_smem(s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    if (tid == 0) {
        for (int i = 0; i < kWsSlots; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(FULL(i)), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(READY(i)), "r"(2));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(DONE(i)), "r"(1));
        }
        for (int g = 0; g < 2; ++g) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(TFULL(g)), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(TEMPTY(g)), "r"(16));
        }
        *s_abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // The bias rides on the tensor cores as one more K = 8 step per tile: A = [1, 1, 0...], W = [hi, lo, 0...] with
    // bias = hi + lo split into two TF32 values (fp32-accurate sum).  Constant tiles, built once.
    if (tid < 128) {
        ch_bias_a(sBiasA, tid);
        for (int l = 0; l < H.nl; ++l) {
            const float* bias = C.bias[l];
            const int n = (int)rank * 128 + tid;
            const float v = (bias && n < cout) ? __ldg(bias + n) : 0.0f;
            const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            *reinterpret_cast<float4*>(sBiasB + ch_koff(tid, 2 * l)) = make_float4(hi, __fsub_rn(v, hi), 0.f, 0.f);
            *reinterpret_cast<float4*>(sBiasB + ch_koff(tid, 2 * l + 1)) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    ws_cluster_sync();                                                 // the peer's barriers and bias tiles exist before anyone uses them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *s_tmem;

    if (warp == 16) {
        // ------------------------------------------------------------------ producer (one lane)
        if (lane == 0) {
            bool ok = true;
            int j = 0;
            for (int i = 0; ok; ++i) {
                const WsTile q = ws_tile(H, C, pair0 + i * npairs, total_pairs, (int)rank);
                if (!q.valid) break;
                const int row0 = q.b * C.cin[q.l];                     // first row of the image in the (s, b*cin + ci) view
                for (int kb = 0; kb < q.nkb && ok; ++kb, ++j) {
                    const int s = j % kWsSlots;
                    if (j >= kWsSlots) ok &= ws_wait<true>(DONE(s), (uint32_t)((j / kWsSlots - 1) & 1), s_abort);   // the slot's last readers completed
                    if (!ok) break;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(FULL(s)), "r"(kWsStageTx) : "memory");
                    const uint32_t a0 = ch_smem(sA + s * kChABytes);
#pragma unroll
                    for (int mb = 0; mb < 4; ++mb) ws_tma_2d(a0 + mb * 4096, &M.x[q.l], q.s0 + 32 * mb, row0 + kb * kChKB, FULL(s));
                    ws_tma_2d(ch_smem(sB + s * kWsBBytes), &M.w[q.l], kb * kChKB, (int)rank * 128, FULL(s));
                }
            }
        }
    } else if (warp == 17) {
        // ------------------------------------------------------------------ forwarder (one lane)
        if (lane == 0) {
            bool ok = true;
            int j = 0;
            for (int i = 0; ok; ++i) {
                const WsTile q = ws_tile(H, C, pair0 + i * npairs, total_pairs, (int)rank);
                if (!q.valid) break;
                for (int kb = 0; kb < q.nkb && ok; ++kb, ++j) {
                    const int s = j % kWsSlots;
                    ok &= ws_wait<true>(FULL(s), (uint32_t)((j / kWsSlots) & 1), s_abort);
                    if (ok) ws_arrive_leader(READY(s));
                }
            }
        }
    } else if (warp == 18) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA)
        if (lane == 0 && rank == 0) {
            bool ok = true;
            int j = 0;
            for (int i = 0; ok; ++i) {
                const WsTile q = ws_tile(H, C, pair0 + i * npairs, total_pairs, 0);
                if (!q.valid) break;
                const int g = i & 1;
                if (i >= 2) ok &= ws_wait<true>(TEMPTY(g), (uint32_t)(((i - 2) >> 1) & 1), s_abort);   // both epilogues drained tile i-2
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc_addr = tmem + (uint32_t)(g * kChN);
                for (int kb = 0; kb < q.nkb && ok; ++kb, ++j) {
                    const int s = j % kWsSlots;
                    ok &= ws_wait<true>(READY(s), (uint32_t)((j / kWsSlots) & 1), s_abort);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = ch_smem(sA + s * kChABytes), b0 = ch_smem(sB + s * kWsBBytes);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {                           // K = 8: two k-groups (512 B each) of A, 32 B of W's rows
                        const uint32_t acc = (kb | ks) ? 1u : 0u;
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                     :: "r"(acc_addr), "l"(ch_desc_a_tma(a0 + ks * 1024)), "l"(ch_desc(b0 + ks * 32)), "r"(kWsIdesc), "r"(acc) : "memory");
                    }
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 :: "r"(DONE(s)), "h"((uint16_t)3) : "memory");
                }
                if (ok) {                                                      // bias step, then the tile is complete
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 :: "r"(acc_addr), "l"(ch_desc_a(ch_smem(sBiasA))), "l"(ch_desc(ch_smem(sBiasB) + q.l * 32)),
                                    "r"(kWsIdesc), "r"(1u) : "memory");
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 :: "r"(TFULL(g)), "h"((uint16_t)3) : "memory");
                }
            }
        }
    } else if (warp < 16) {
        // ------------------------------------------------------------------ epilogue: accumulator g = warps 8g .. 8g+7, two
        // teams of 4 warps (one per 32-lane quarter of TMEM) taking alternate anchors.  An accumulator is refilled only
        // after its previous tile is drained, so a tile costs (mainloop + epilogue) / 2 whenever the epilogue is the
        // longer of the two: the epilogue's LATENCY counts, and two teams cut it by a third (three anchors).
        const int g = warp >> 3, team = (warp >> 2) & 1, ewarp = warp & 3, etid = ewarp * 32 + lane;
        bool ok = true;
        for (int i = g; ; i += 2) {
            const WsTile q = ws_tile(H, C, pair0 + i * npairs, total_pairs, (int)rank);
            if (!q.valid) break;
            ok &= ws_wait<false>(TFULL(g), (uint32_t)((i >> 1) & 1), s_abort);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifndef VK_CONV_NO_EPILOGUE      // (tuning builds: how long the mainloop alone takes)
            conv_epilogue(H, A, tmem + ((uint32_t)(ewarp * 32) << 16) + (uint32_t)(g * kChN), q.l, q.b, q.s0, q.nvalid,
                          etid, s_wtot + 4 * (2 * g + team), 3 + 2 * g + team, ok, team, 2);
#endif
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) ws_arrive_leader(TEMPTY(g));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0 && *s_abort) atomicExch(fault, 1);
    ws_cluster_sync();                                                 // neither CTA leaves while the other may still read its shared memory
    if (warp == 18) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512));
}

}  // namespace vk

using namespace vk;

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*VkEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_conv_maps(const HeadDev& H, const ConvHead& C, int batch, ConvMaps* maps) {
    static VkEncodeTiled enc = nullptr;
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn)
            return fail_code(VK_E_LIMIT, "vk_conv_decode_filter: the driver has no cuTensorMapEncodeTiled");
        enc = reinterpret_cast<VkEncodeTiled>(fn);
    }
    memset(maps, 0, sizeof(*maps));
    const cuuint32_t ones[2] = {1, 1};
    for (int l = 0; l < H.nl; ++l) {
        // X: element (s, r) of the (nynx, B * cin) view = X[b][ci][s], r = b * cin + ci; box = 32 positions x 32 channels,
        // written in the MN-major operand layout (32-byte swizzle atoms); positions past the plane are zeros
        const cuuint64_t xd[2] = {(cuuint64_t)H.nynx[l], (cuuint64_t)batch * C.cin[l]}, xs[1] = {(cuuint64_t)H.nynx[l] * 4};
        const cuuint32_t xb[2] = {32, 32};
        CUresult r = enc(&maps->x[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(C.x[l]), xd, xs, xb, ones,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail_code(VK_E_LIMIT, "vk_conv_decode_filter: tensor map of the level-%d features: CUresult %d", l, (int)r);
        // W: element (ci, co); box = 32 channels x 128 output channels, K-major 128-byte swizzle; rows past na*no are zeros
        const cuuint64_t wd[2] = {(cuuint64_t)C.cin[l], (cuuint64_t)(H.na * H.no)}, ws[1] = {(cuuint64_t)C.cin[l] * 4};
        const cuuint32_t wb[2] = {32, 128};
        r = enc(&maps->w[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(C.w[l]), wd, ws, wb, ones,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail_code(VK_E_LIMIT, "vk_conv_decode_filter: tensor map of the level-%d weights: CUresult %d", l, (int)r);
    }
    return VK_OK;
}

extern "C" int vk_conv_decode_filter(const VkHeadCfg* cfg, const float* const* feats, const int32_t* cin,
                                     const float* const* weights, const float* const* biases, int batch,
                                     float conf_thres, int multi_label, const uint32_t* class_mask, int kernel,
                                     const VkCandBuf* out, int32_t* fault, vk_stream_t stream_) {
    HeadDev H;
    if (int rc = make_head(cfg, &H, "vk_conv_decode_filter")) return rc;
    if (batch == 0) return VK_OK;
    if (!feats || !cin || !weights || !fault || batch < 0) return fail_arg("vk_conv_decode_filter: null/negative argument");
    if (kernel != VK_CONV_TILE && kernel != VK_CONV_PERSISTENT) return fail_arg("vk_conv_decode_filter: kernel %d", kernel);
    if (!(conf_thres >= 0.f && conf_thres <= 1.f)) return fail_arg("vk_conv_decode_filter: conf_thres %g outside [0,1]", conf_thres);
    if (batch > 65535) return fail_code(VK_E_LIMIT, "vk_conv_decode_filter: batch %d > 65535", batch);
    if (H.na * H.no > kChN || (H.na - 1) * H.no + 5 + 16 * ceil_div(H.nc, 16) > kChN ||
        (H.nc <= 64 && (H.na - 1) * H.no + 5 + 64 > kChN))            // the epilogue's 32-column loads stay inside the accumulator
        return fail_code(VK_E_LIMIT, "vk_conv_decode_filter: %d output channels do not fit %d TMEM columns", H.na * H.no, kChN);
    if (int rc = check_cand(out, H.rows, H.tiles, H.nc, multi_label, "vk_conv_decode_filter")) return rc;
    ConvHead C;
    memset(&C, 0, sizeof(C));
    int mt = 0, pt = 0;
    for (int l = 0; l < H.nl; ++l) {
        if (!feats[l] || !weights[l]) return fail_arg("vk_conv_decode_filter: level %d is NULL", l);
        if (cin[l] <= 0 || cin[l] % kChKB) return fail_code(VK_E_LIMIT, "vk_conv_decode_filter: cin[%d] = %d is not a multiple of %d", l, cin[l], kChKB);
        if (H.nynx[l] % 4 || (reinterpret_cast<uintptr_t>(feats[l]) & 15) || (reinterpret_cast<uintptr_t>(weights[l]) & 15))
            return fail_code(VK_E_LIMIT, "vk_conv_decode_filter: level %d needs ny*nx %% 4 == 0 and 16-byte aligned tensors", l);
        C.x[l] = feats[l]; C.w[l] = weights[l]; C.bias[l] = biases ? biases[l] : nullptr; C.cin[l] = cin[l];
        C.mtile_start[l] = mt;
        C.ptile_start[l] = pt;
        mt += ceil_div(H.nynx[l], kChM);
        pt += ceil_div(H.nynx[l], 2 * kChM);
    }
    for (int l = H.nl; l <= VK_MAX_LEVELS; ++l) { C.mtile_start[l] = mt; C.ptile_start[l] = pt; }
    C.mtiles = mt;
    C.ptiles = pt;
    cudaStream_t stream = as_stream(stream_);
    if (int rc = reset_cand(out, batch, "vk_conv_decode_filter", stream)) return rc;
    cudaError_t e = cudaMemsetAsync(fault, 0, sizeof(int32_t), stream);
    if (e != cudaSuccess) return fail_code((int)e, "vk_conv_decode_filter: memset: %s", cudaGetErrorString(e));
    FilterArgs A = make_filter_args(out, batch, conf_thres, multi_label, class_mask);
    if (kernel == VK_CONV_PERSISTENT) {
        const int total = pt * batch;                                  // pair tiles
        ConvMaps maps;
        if (int rc = make_conv_maps(H, C, batch, &maps)) return rc;
        if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&conv_decode_filter_ws_kernel), kWsSmem, "vk_conv_decode_filter")) return rc;
        const int pairs = total < kNumSMs / 2 ? total : kNumSMs / 2;
        conv_decode_filter_ws_kernel<<<2 * pairs, kWsThreads, kWsSmem, stream>>>(H, C, A, maps, total, fault);
        count_launch();
        return check_launch("conv_decode_filter_ws_kernel");
    }
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&conv_decode_filter_kernel), kChSmem, "vk_conv_decode_filter")) return rc;
    conv_decode_filter_kernel<<<dim3(mt, batch), kChThreads, kChSmem, stream>>>(H, C, A, fault);
    count_launch();
    return check_launch("conv_decode_filter_kernel");
}
