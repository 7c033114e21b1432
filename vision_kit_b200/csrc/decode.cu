// Detect-head decode and the confidence filter (SURVEY.md §8 a3, a4, a5 first half).
//
//   vk_detect_decode   conv outputs (B, na*no, ny, nx) -> pred (B, rows, no): a transposing
//                      copy with the sigmoid/grid/anchor decode fused in.  HBM-bound:
//                      8 568 000 B read + 8 568 000 B written per 640x640 image.
//   vk_decode_filter   the same tiles, but only rows with obj > conf are decoded and only
//                      candidates leave the SM: pred is never materialised.  Two kernels with
//                      identical results: one warp per tile gathering just the surviving rows
//                      (demo thresholds), and a persistent whole-tile kernel (eval thresholds).
//   vk_filter_pred     drop-in filter for an existing pred tensor (`nms(prediction)`).
//
// Candidate order: the reference's candidate list is ordered (row asc, class asc)
// (`nonzero`, utils/image_proc.py:141-143).  Here every tile writes its candidates, in that
// order, into the fixed slot range its tile owns (64 rows x nc slots, or 64 in best-class
// mode) and records its count in a segment table indexed by tile, so slot order IS the
// canonical order.  No tile ever waits for another, the buffer cannot overflow, and the
// consumer (nms.cu) breaks score ties by slot.
#include "decode_common.cuh"

namespace vk {


// ---------------------------------------------------------------------------------------
// materialised decode
// ---------------------------------------------------------------------------------------
// Persistent blocks, each walking tiles t = blockIdx.x, +gridDim.x, ... of the whole batch with a
// two-deep shared-memory pipeline: while tile k is decoded and stored, the loads of tile k+1 are
// already in flight, so every resident block keeps a full tile (64*no*4 B) of reads outstanding.
//
//  in   16-byte async copies (LDGSTS.128; 4-byte ones move only ~13 B/clk/SM on B200 and are
//       kept for unaligned planes) of 4 consecutive rows of one channel plane into tile[c][64].
//       The 16-byte chunk q of channel c sits at chunk position q ^ (c & 7), so that both the
//       copies (8 lanes = 8 chunks of one channel) and the reads below (8 lanes = one chunk of 8
//       consecutive channels) touch all 32 banks.
//  out  a warp takes 4 rows x 32 channels: one LDS.128 per lane (its channel, 4 rows), 4 sigmoids,
//       4 scalar stores -- each store instruction writes 128 contiguous bytes of one pred row.
//       Lanes 0-3 of the first channel group hold the box channels and decode them instead.
//  Tile descriptors (the integer divisions that locate a tile) are computed by one thread, two
//  tiles ahead, and broadcast through shared memory.  profiles/micro/transpose_bw.cu is the design
//  study behind these choices (6.4 TB/s for this structure vs 4.7 TB/s with 4-byte copies).
__device__ __forceinline__ void cp_async_4(uint32_t smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_dst), "l"(gsrc) : "memory");
}

struct DecTile {
    const void* src;      // first logit of the tile: channel 0, row s0 of plane (b, a); element type T
    float* out;           // pred chunk
    float* raw;           // raw chunk or null
    int nynx, nvalid, nx, gy0, gx0;   // (gy0, gx0): grid cell of the tile's first row
    int vec;              // 16-byte copies allowed
    int raw_bulk;         // raw chunk leaves as one bulk copy from shared memory (16-byte aligned, size % 16 == 0)
    float stride, aw, ah;
};

// Tile descriptors are produced by ONE thread, two tiles ahead of their use; that thread's warp is
// what the block waits for at the next barrier, so the arithmetic avoids integer division: the
// (image, tile) cursor advances by a precomputed (quotient, remainder) step, the anchor comes
// from at most na-1 subtractions and the grid row from a corrected float reciprocal.
struct DecCursor {
    int b, ti;            // image and tile inside the image of the NEXT descriptor to produce
    int step_b, step_t;   // gridDim.x = step_b * tiles + step_t
    int t;                // linear tile index of (b, ti)
};

__device__ __forceinline__ DecCursor cursor_begin(const HeadDev& H, int t0, int step) {
    DecCursor c;
    c.b = t0 / H.tiles; c.ti = t0 - c.b * H.tiles;
    c.step_b = step / H.tiles; c.step_t = step - c.step_b * H.tiles;
    c.t = t0;
    return c;
}

__device__ __forceinline__ void cursor_next(const HeadDev& H, DecCursor& c) {
    c.b += c.step_b; c.ti += c.step_t;
    if (c.ti >= H.tiles) { c.ti -= H.tiles; ++c.b; }
    c.t += c.step_b * H.tiles + c.step_t;
}

struct TilePos {          // what the filter kernel needs on top of DecTile
    int b, seg, row0, s0; // image, tile index inside the image, first pred row, first row inside the plane
};

template <class T>
__device__ __forceinline__ void decode_tile_at(const HeadDev& H, float* pred, const DecCursor& cur, DecTile* d,
                                               TilePos* loc = nullptr) {
    const int b = cur.b;
    int l = 0;
#pragma unroll
    for (int i = 1; i < VK_MAX_LEVELS; ++i)
        if (i < H.nl && cur.ti >= H.tile_start[i]) l = i;
    int rel = cur.ti - H.tile_start[l], a = 0;
    const int tpa = H.tpa[l];
    while (rel >= tpa) { rel -= tpa; ++a; }
    const int no = H.no, nynx = H.nynx[l], nx = H.nx[l];
    const int s0 = rel * kTileS;
    const int nvalid = min(kTileS, nynx - s0);
    const int row0 = H.row_base[l] + a * nynx + s0;
    int gy0 = (int)((float)s0 * __frcp_rn((float)nx));      // s0 < 2^24: off by at most one, fixed below
    int gx0 = s0 - gy0 * nx;
    if (gx0 < 0) { --gy0; gx0 += nx; }
    if (gx0 >= nx) { ++gy0; gx0 -= nx; }
    d->src = static_cast<const T*>(H.lv[l]) + ((size_t)(b * H.na + a) * no) * nynx + s0;
    d->out = pred ? pred + ((size_t)b * H.rows + row0) * no : nullptr;
    d->raw = H.raw[l] ? H.raw[l] + (((size_t)b * H.na + a) * nynx + s0) * no : nullptr;
    d->nynx = nynx; d->nvalid = nvalid; d->nx = nx;
    d->gy0 = gy0; d->gx0 = gx0;
    if (loc) { loc->b = b; loc->seg = cur.ti; loc->row0 = row0; loc->s0 = s0; }
    d->raw_bulk = d->raw != nullptr && ((reinterpret_cast<uintptr_t>(d->raw) & 15) == 0) && (((nvalid * no) & 3) == 0);
    d->vec = ((nynx & (16 / (int)sizeof(T) - 1)) == 0) && ((reinterpret_cast<uintptr_t>(H.lv[l]) & 15) == 0);
    d->stride = H.stride[l]; d->aw = H.anchors[l][2 * a]; d->ah = H.anchors[l][2 * a + 1];
}

__device__ __forceinline__ int swz(int c, int s) { return c * kTileS + ((((s >> 2) ^ c) & 7) << 2 | (s & 32) | (s & 3)); }

__device__ __forceinline__ void decode_prefetch(const DecTile& d, float* tile, int no) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tile);
    const float* const tsrc = static_cast<const float*>(d.src);
    if (d.vec) {
        const int q = threadIdx.x & 15, c0 = threadIdx.x >> 4;      // 16 chunks per channel, 16 channels per pass
        if (4 * q < d.nvalid) {
            const float* src = tsrc + (size_t)c0 * d.nynx + 4 * q;
            const size_t step = (size_t)(kDecThreads / 16) * d.nynx;
            // c advances by 16: (c & 7) and with it the chunk position stay the same, dst moves 16 channel rows
            uint32_t dst = base + 4u * (uint32_t)swz(c0, 4 * q);
            for (int c = c0; c < no; c += kDecThreads / 16, src += step, dst += 16 * kTileS * 4) cp_async_16(dst, src);
        }
    } else {
        const int r = threadIdx.x & (kTileS - 1), c0 = threadIdx.x >> 6;
        if (r < d.nvalid) {
            const float* src = tsrc + (size_t)c0 * d.nynx + r;
            const size_t step = (size_t)(kDecThreads / kTileS) * d.nynx;
            for (int c = c0; c < no; c += kDecThreads / kTileS, src += step)
                cp_async_4(base + 4u * (uint32_t)swz(c, r), src);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// Tile layout of detect_decode_kernel per input element type.
//   float       [c][64] floats, 16-byte chunks swizzled (swz), 16-byte async copies
//   half types  [c][68] elements (136-byte pitch: lanes over channels reading 4 rows = 8 bytes are
//               conflict-free per half-warp), 8-byte async copies of 4 rows
template <class T>
struct DecLayout {
    static constexpr int kPitch = kTileS + 4;                                   // elements
    static __device__ __forceinline__ int tile_bytes(int no) { return (no * kPitch * 2 + 15) & ~15; }
    static __device__ __forceinline__ void prefetch(const DecTile& d, unsigned char* tile, int no) {
        const uint32_t base = (uint32_t)__cvta_generic_to_shared(tile);
        const T* const tsrc = static_cast<const T*>(d.src);
        const bool vec8 = ((d.nynx & 3) == 0) && ((reinterpret_cast<uintptr_t>(tsrc) & 7) == 0);
        if (vec8) {
            const int q = threadIdx.x & 15, c0 = threadIdx.x >> 4;              // 16 chunks of 4 rows per channel
            if (4 * q < d.nvalid)
                for (int c = c0; c < no; c += kDecThreads / 16)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;"
                                 :: "r"(base + 2u * (uint32_t)(c * kPitch + 4 * q)), "l"(tsrc + (size_t)c * d.nynx + 4 * q) : "memory");
        } else {
            T* const t = reinterpret_cast<T*>(tile);
            for (int e = threadIdx.x; e < no * kTileS; e += kDecThreads) {
                const int c = e >> 6, r = e & 63;
                if (r < d.nvalid) t[c * kPitch + r] = tsrc[(size_t)c * d.nynx + r];
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    static __device__ __forceinline__ float4 load4(const unsigned char* tile, int c, int q) {   // rows 4q..4q+3 of channel c
        const uint2 u = *reinterpret_cast<const uint2*>(tile + 2 * (c * kPitch + 4 * q));
        const T* e = reinterpret_cast<const T*>(&u);
        return make_float4(to_f32(e[0]), to_f32(e[1]), to_f32(e[2]), to_f32(e[3]));
    }
    static __device__ __forceinline__ float elem(const unsigned char* tile, int c, int s) {
        return to_f32(reinterpret_cast<const T*>(tile)[c * kPitch + s]);
    }
};
template <>
struct DecLayout<float> {
    static __device__ __forceinline__ int tile_bytes(int no) { return no * kTileS * 4; }
    static __device__ __forceinline__ void prefetch(const DecTile& d, unsigned char* tile, int no) {
        decode_prefetch(d, reinterpret_cast<float*>(tile), no);
    }
    static __device__ __forceinline__ float4 load4(const unsigned char* tile, int c, int q) {
        return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(tile) + swz(c, 4 * q));
    }
    static __device__ __forceinline__ float elem(const unsigned char* tile, int c, int s) {
        return reinterpret_cast<const float*>(tile)[swz(c, s)];
    }
};

// NCG > 0: no <= 32*NCG channels, item loop unrolled; NCG == 0: runtime loop over channel groups.
template <class T, int NCG>
__global__ void __launch_bounds__(kDecThreads, 3)
detect_decode_kernel(const HeadDev H, float* __restrict__ pred, int total_tiles, int have_lin) {
    extern __shared__ __align__(16) unsigned char tiles_sm[];  // 2 x tile of logits (DecLayout<T>) (+ [kTileS][no] floats of raw staging)
    __shared__ DecTile s_dt[3];
    const int no = H.no;
    const int tile_bytes = DecLayout<T>::tile_bytes(no);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int t = blockIdx.x;
    if (t >= total_tiles) return;
    __shared__ DecCursor cur;                       // thread 0 only (shared: keeps it out of everyone's registers)
    if (threadIdx.x == 0) {
        cur = cursor_begin(H, t, gridDim.x);
        decode_tile_at<T>(H, pred, cur, &s_dt[0]);
        cursor_next(H, cur);
        if (cur.t < total_tiles) decode_tile_at<T>(H, pred, cur, &s_dt[1]);
        cursor_next(H, cur);
    }
    __syncthreads();
    DecLayout<T>::prefetch(s_dt[0], tiles_sm, no);
    for (int k = 0; t < total_tiles; ++k, t += gridDim.x) {
        const unsigned char* tile = tiles_sm + (k & 1) * tile_bytes;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (have_lin && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();                      // tile k landed; everyone is done with the other buffer (and `lin`)
        const int slot = k % 3;
        if (t + (int)gridDim.x < total_tiles)
            DecLayout<T>::prefetch(s_dt[slot == 2 ? 0 : slot + 1], tiles_sm + ((k + 1) & 1) * tile_bytes, no);
        if (threadIdx.x == 0) {                                            // slot (k+2)%3 was tile k-1's: free
            if (cur.t < total_tiles) decode_tile_at<T>(H, pred, cur, &s_dt[slot == 0 ? 2 : slot - 1]);
            cursor_next(H, cur);
        }
        const DecTile& d = s_dt[slot];
        const int nvalid = d.nvalid;
        float* __restrict__ out = d.out;
        float* __restrict__ raw = d.raw;
        float* const lin = reinterpret_cast<float*>(tiles_sm + 2 * tile_bytes);   // raw logits of the tile in output order
        const bool raw_bulk = have_lin && d.raw_bulk;

        // Box channels.  Warp w owns rows 4w..4w+3 and 4(w+8)..4(w+8)+3 in the first channel group:
        // 8 rows x 4 box channels = one element per lane, decoded here and handed to lanes 0-3 of
        // the two items by shuffles, so that a row's first 128 bytes leave in ONE store instruction
        // (the sector holding channels 0-7 is never written in two pieces).
        float boxv;
        {
            const int rl = lane >> 2, cb = lane & 3;
            const int br = (rl < 4) ? 4 * w + rl : 4 * (w + 8) + rl - 4;
            const int nx = d.nx;
            int gx = d.gx0 + br, gy = d.gy0;
            if (gx >= nx) { const int wq = gx / nx; gy += wq; gx -= wq * nx; }
            const float anc = (cb & 1) ? d.ah : d.aw;
            boxv = decode_elem(DecLayout<T>::elem(tile, cb, br), cb, (float)((cb & 1) ? gy : gx), d.stride, anc, H.variant);
        }
        auto item = [&](int cgp, int q, const float4 v) {
            // rows 4q..4q+3 of channel c = 32*cgp + lane
            const int c = 32 * cgp + lane;
            const int r0 = 4 * q;
            const int left = nvalid - r0;
            float r[4] = {sigmoidf_vk(v.x), sigmoidf_vk(v.y), sigmoidf_vk(v.z), sigmoidf_vk(v.w)};
            if (cgp == 0) {
                const int src0 = (q >= 8 ? 16 : 0) + (lane & 3);
#pragma unroll
                for (int jr = 0; jr < 4; ++jr) {
                    const float bx = __shfl_sync(0xffffffffu, boxv, src0 + 4 * jr);
                    if (lane < 4) r[jr] = bx;
                }
            }
            if (c < no) {
                float* po = out + r0 * no + c;
                if (left >= 4) {
                    st_stream_f32(po, r[0]); st_stream_f32(po + no, r[1]);
                    st_stream_f32(po + 2 * no, r[2]); st_stream_f32(po + 3 * no, r[3]);
                } else {
#pragma unroll
                    for (int jr = 0; jr < 4; ++jr) if (jr < left) st_stream_f32(po + jr * no, r[jr]);
                }
                if (raw_bulk) {
                    float* pl = lin + r0 * no + c;                   // rows past nvalid stay inside the buffer
                    pl[0] = v.x; pl[no] = v.y; pl[2 * no] = v.z; pl[3 * no] = v.w;
                } else if (raw) {
                    float* pr = raw + r0 * no + c;
                    const float l4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int jr = 0; jr < 4; ++jr) if (jr < left) st_stream_f32(pr + jr * no, l4[jr]);
                }
            }
        };
        // every lane of a warp enters item() together (shuffles inside): q is warp-uniform
        if (NCG > 0) {
            float4 v[2 * (NCG > 0 ? NCG : 1)];
#pragma unroll
            for (int j = 0; j < 2 * NCG; ++j) {
                const int c = 32 * (j >> 1) + lane, q = w + 8 * (j & 1);
                v[j] = (c < no) ? DecLayout<T>::load4(tile, c, q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 2 * NCG; ++j) {
                const int q = w + 8 * (j & 1);
                if (4 * q < nvalid) item(j >> 1, q, v[j]);
            }
        } else {
            for (int cgp = 0; 32 * cgp < no; ++cgp)
                for (int q = w; q < 16; q += kWarps) {
                    const int c = 32 * cgp + lane;
                    if (4 * q < nvalid)
                        item(cgp, q, (c < no) ? DecLayout<T>::load4(tile, c, q) : make_float4(0.f, 0.f, 0.f, 0.f));
                }
        }
        if (raw_bulk) {
            // generic-proxy writes of `lin` -> visible to the bulk-copy engine -> one 64*no*4-byte store
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             :: "l"(raw), "r"((uint32_t)__cvta_generic_to_shared(lin)), "r"(nvalid * no * 4) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (have_lin && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}


// ---------------------------------------------------------------------------------------
// confidence filter, sparse form (demo thresholds: a few percent of the rows survive).
//
// One WARP owns one 64-row tile from its objectness to its segment-table entry, so nothing in
// the kernel waits on a block barrier and a tile's candidates are ordered by construction:
//
//   1  objectness of the tile: two coalesced loads per lane, sigmoid, two ballots -> a 64-bit
//      survivor mask (utils/image_proc.py:99)
//   2  the surviving rows are gathered four at a time: lane j reads channels j, j+32, j+64 of a
//      row (one 32-byte sector each from the NCHW conv output; coalesced from a pred tensor),
//      so a warp keeps up to 12 loads per lane in flight
//   3  per row: class products on all lanes, best class by a warp arg-max that prefers the
//      lower class on ties (= the first maximum the reference's `max(1)` returns, :145), or
//      multi-label ballots in channel order (`nonzero` order, :141-143); lane 0 decodes the box
//   4  candidates go to the tile's own slot range in order; the tile's count joins the image's
//      counter with a fire-and-forget atomic (nothing in the kernel waits for an atomic's result)
//
// With ~1.8 surviving rows per tile the kernel is bound by the rate at which HBM serves scattered sectors
// (85 per surviving row, each in a different DRAM page; profiles/micro/sector_gather.cu).
// ---------------------------------------------------------------------------------------
#ifndef VK_DENSE_BPS
#define VK_DENSE_BPS 4
#endif
#ifndef VK_ROW_BATCH
#define VK_ROW_BATCH 4
#endif
#ifndef VK_ROWS_BPS
#define VK_ROWS_BPS 8
#endif
#ifndef VK_ROWS_WARPS
#define VK_ROWS_WARPS 4
#endif
#ifndef VK_ROWS_PERSISTENT
#define VK_ROWS_PERSISTENT 0
#endif
constexpr int kRowBatch = VK_ROW_BATCH;      // surviving rows whose gathers are in flight together
constexpr int kRowWarps = VK_ROWS_WARPS;     // warps (= tiles in flight) per block of the sparse kernels

struct PlaneGeom {       // fused path: what is needed to decode a box from logits
    int variant, nx, s0; // s0: spatial index of the tile's first row inside its plane
    float stride, aw, ah;
    __device__ __forceinline__ float4 box(float l0, float l1, float l2, float l3, int sp) const {
        const int gy = sp / nx, gx = sp - gy * nx;
        return xyxy_from_cxcywh(decode_elem(l0, 0, (float)gx, stride, aw, variant),
                                decode_elem(l1, 1, (float)gy, stride, ah, variant),
                                decode_elem(l2, 2, 0.f, stride, aw, variant),
                                decode_elem(l3, 3, 0.f, stride, ah, variant));
    }
};

// ---- where the 5+nc values of row r of a tile live and how they turn into numbers
template <class T>
struct LogitRows {       // conv output plane (b, a): element (row r, channel c) at base[c * nynx + r]
    const T* base; int nynx; PlaneGeom g;
    __device__ __forceinline__ float obj_val(float raw) const { return sigmoidf_vk(raw); }
    __device__ __forceinline__ float raw(int r, int c) const { return ld_elem(base + (size_t)c * nynx + r); }
    __device__ __forceinline__ float prob(float x) const { return sigmoidf_vk(x); }
    __device__ __forceinline__ float4 box(float l0, float l1, float l2, float l3, int r) const {
        return g.box(l0, l1, l2, l3, g.s0 + r);
    }
};
template <class T>
struct PredRows {        // decoded prediction rows: element (row r, channel c) at base[r * no + c]
    const T* base; int no;
    __device__ __forceinline__ float obj_val(float raw) const { return raw; }
    __device__ __forceinline__ float raw(int r, int c) const { return ld_elem(base + (size_t)r * no + c); }
    __device__ __forceinline__ float prob(float x) const { return x; }
    __device__ __forceinline__ float4 box(float l0, float l1, float l2, float l3, int) const {
        return xyxy_from_cxcywh(l0, l1, l2, l3);
    }
};

// Everything a warp needs to emit candidates of its tile.
struct TileOut {
    uint2* cand;         // the tile's slot range
    float4* boxes;       // the image's boxes
    int row0;            // first prediction row of the tile
    int cnt;             // candidates written so far (warp-uniform)
};

// One surviving row whose channels lane, lane+32, ... are in x[0..NK): emits its candidates.
template <class Src, int NK, bool ML>
__device__ __forceinline__ void emit_row(const Src& S, const FilterArgs& A, TileOut& O, int r, float o, const float* x) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int nc = A.nc;
    const int row = O.row0 + r;
    int total;
    if (ML) {
        float p[NK];
        unsigned bal[NK];
        total = 0;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const int cls = lane + 32 * k - 5;
            p[k] = __fmul_rn(S.prob(x[k]), o);                                                   // image_proc.py:135
            const bool f = cls >= 0 && cls < nc && p[k] > A.conf && class_allowed(A.class_mask, cls);   // :141,151
            bal[k] = __ballot_sync(0xffffffffu, f);
            total += __popc(bal[k]);
        }
        if (total == 0) return;
        int pos = O.cnt;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            if ((bal[k] >> lane) & 1u)
                O.cand[pos + __popc(bal[k] & lt)] = make_uint2(__float_as_uint(p[k]), (uint32_t)(row * nc + lane + 32 * k - 5));
            pos += __popc(bal[k]);
        }
    } else {
        float bv = -INFINITY;
        int bj = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const int cls = lane + 32 * k - 5;
            const float pk = __fmul_rn(S.prob(x[k]), o);
            if (cls >= 0 && cls < nc && pk > bv) { bv = pk; bj = cls; }                          // first max of the lane
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {                                                 // first max of the row (:145)
            const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
            if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
        }
        total = (bj != 0x7fffffff && bv > A.conf && class_allowed(A.class_mask, bj)) ? 1 : 0;    // :147,151
        if (total == 0) return;
        if (lane == 0) O.cand[O.cnt] = make_uint2(__float_as_uint(bv), (uint32_t)(row * nc + bj));
    }
    const float l1 = __shfl_sync(0xffffffffu, x[0], 1), l2 = __shfl_sync(0xffffffffu, x[0], 2),
                l3 = __shfl_sync(0xffffffffu, x[0], 3);
    if (lane == 0) O.boxes[row] = S.box(x[0], l1, l2, l3, r);
    O.cnt += total;
}

// The tile's surviving rows (objectness o0: rows 0-31, o1: rows 32-63, already evaluated) -> candidates
// -> segment count.  NK > 0: no <= 32 * NK channels, rows gathered kRowBatch at a time; NK == 0: any
// channel count, one row at a time.
template <class Src, int NK, bool ML>
__device__ __forceinline__ void filter_tile_rows(const Src& S, const FilterArgs& A, int b, int seg, int row0,
                                                 float o0, float o1) {
    const int lane = threadIdx.x & 31;
    const int no = A.nc + 5;
    const unsigned m0 = __ballot_sync(0xffffffffu, o0 > A.conf);                                 // image_proc.py:99
    const unsigned m1 = __ballot_sync(0xffffffffu, o1 > A.conf);
    unsigned long long mask = (unsigned long long)m0 | ((unsigned long long)m1 << 32);
    TileOut O;
    O.cand = reinterpret_cast<uint2*>(A.cand + (size_t)b * A.cap) + (size_t)seg * A.tile_cap;
    O.boxes = A.boxes + (size_t)b * A.rows;
    O.row0 = row0;
    O.cnt = 0;
    if (NK > 0) {
        constexpr int NKx = NK > 0 ? NK : 1;
        while (mask) {
            int rr[kRowBatch];
            float x[kRowBatch][NKx];
#pragma unroll
            for (int j = 0; j < kRowBatch; ++j) {
                rr[j] = mask ? __ffsll((long long)mask) - 1 : -1;
                if (mask) mask &= mask - 1;
#pragma unroll
                for (int k = 0; k < NKx; ++k) {
                    const int c = lane + 32 * k;
                    x[j][k] = (rr[j] >= 0 && c < no) ? S.raw(rr[j], c) : 0.0f;
                }
            }
#pragma unroll
            for (int j = 0; j < kRowBatch; ++j) {
                if (rr[j] < 0) break;                                  // warp-uniform
                const float o = __shfl_sync(0xffffffffu, rr[j] < 32 ? o0 : o1, rr[j] & 31);
                emit_row<Src, NKx, ML>(S, A, O, rr[j], o, x[j]);
            }
        }
    } else {
        // any class count: the row is walked 32 channels at a time with running state
        const unsigned lt = (1u << lane) - 1u;
        const int nc = A.nc;
        while (mask) {
            const int r = __ffsll((long long)mask) - 1;
            mask &= mask - 1;
            const float o = __shfl_sync(0xffffffffu, r < 32 ? o0 : o1, r & 31);
            const int row = row0 + r;
            const float x0 = (lane < no) ? S.raw(r, lane) : 0.0f;
            if (ML) {
                int pos = O.cnt;
                for (int c0 = 0; c0 < no; c0 += 32) {
                    const int c = c0 + lane, cls = c - 5;
                    const float pv = (c < no) ? __fmul_rn(S.prob(c0 ? S.raw(r, c) : x0), o) : 0.0f;
                    const bool f = cls >= 0 && cls < nc && pv > A.conf && class_allowed(A.class_mask, cls);
                    const unsigned bal = __ballot_sync(0xffffffffu, f);
                    if (f) O.cand[pos + __popc(bal & lt)] = make_uint2(__float_as_uint(pv), (uint32_t)(row * nc + cls));
                    pos += __popc(bal);
                }
                if (pos == O.cnt) continue;
                O.cnt = pos;
            } else {
                float bv = -INFINITY;
                int bj = 0x7fffffff;
                for (int c0 = 0; c0 < no; c0 += 32) {
                    const int c = c0 + lane, cls = c - 5;
                    if (c < no && cls >= 0) {
                        const float pv = __fmul_rn(S.prob(c0 ? S.raw(r, c) : x0), o);
                        if (pv > bv) { bv = pv; bj = cls; }
                    }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
                    const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
                    if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
                }
                if (!(bj != 0x7fffffff && bv > A.conf && class_allowed(A.class_mask, bj))) continue;
                if (lane == 0) O.cand[O.cnt] = make_uint2(__float_as_uint(bv), (uint32_t)(row * nc + bj));
                O.cnt += 1;
            }
            const float l1 = __shfl_sync(0xffffffffu, x0, 1), l2 = __shfl_sync(0xffffffffu, x0, 2),
                        l3 = __shfl_sync(0xffffffffu, x0, 3);
            if (lane == 0) O.boxes[row] = S.box(x0, l1, l2, l3, r);
        }
    }
    if (lane == 0) {
        A.seg_count[(size_t)b * A.segs + seg] = O.cnt;
        if (O.cnt) atomicAdd(A.counts + b, O.cnt);                    // result unused: fire-and-forget
        if (seg == 0) A.flags[b] = cand_flags(A);
    }
}

// ---- tile locators: which tile a linear index is, where its values live
template <class T>
struct LogitTileRef {
    const T* base;       // channel 0, row s0 of plane (b, a)
    int b, seg, l, a, s0, nvalid, row0, nynx;
    __device__ __forceinline__ float obj_raw(int r) const { return r < nvalid ? ld_elem(base + (size_t)4 * nynx + r) : 0.0f; }
};
template <class T>
struct LogitLocator {
    const HeadDev& H;
    typedef LogitTileRef<T> Ref;
    typedef LogitRows<T> Src;
    __device__ __forceinline__ Ref locate(int t) const {
        Ref q;
        q.b = t / H.tiles; q.seg = t - q.b * H.tiles;
        int l = 0;
#pragma unroll
        for (int i = 1; i < VK_MAX_LEVELS; ++i)
            if (i < H.nl && q.seg >= H.tile_start[i]) l = i;
        const int rel = q.seg - H.tile_start[l];
        q.l = l; q.a = rel / H.tpa[l];
        q.s0 = (rel - q.a * H.tpa[l]) * kTileS;
        q.nynx = H.nynx[l];
        q.nvalid = min(kTileS, q.nynx - q.s0);
        q.row0 = H.row_base[l] + q.a * q.nynx + q.s0;
        q.base = static_cast<const T*>(H.lv[l]) + ((size_t)(q.b * H.na + q.a) * H.no) * q.nynx + q.s0;
        return q;
    }
    __device__ __forceinline__ Src source(const Ref& q) const {
        return Src{q.base, q.nynx, PlaneGeom{H.variant, H.nx[q.l], q.s0, H.stride[q.l], H.anchors[q.l][2 * q.a], H.anchors[q.l][2 * q.a + 1]}};
    }
};
template <class T>
struct PredTileRef {
    const T* base;       // row row0 of image b
    int b, seg, nvalid, row0, no;
    __device__ __forceinline__ float obj_raw(int r) const { return r < nvalid ? ld_elem(base + (size_t)r * no + 4) : 0.0f; }
};
template <class T>
struct PredLocator {
    const T* pred; int no, rows, segs;
    typedef PredTileRef<T> Ref;
    typedef PredRows<T> Src;
    __device__ __forceinline__ Ref locate(int t) const {
        Ref q;
        q.b = t / segs; q.seg = t - q.b * segs;
        q.row0 = q.seg * kTileS;
        q.nvalid = min(kTileS, rows - q.row0);
        q.no = no;
        q.base = pred + ((size_t)q.b * rows + q.row0) * no;
        return q;
    }
    __device__ __forceinline__ Src source(const Ref& q) const { return Src{q.base, q.no}; }
};

// One tile per warp: the hardware block scheduler balances the very uneven tiles (a tile costs one more
// dependent round trip per four surviving rows).  Persistent warps with the next tile's objectness
// prefetched (VK_ROWS_PERSISTENT) were measured slower: static striding lets the slowest warp decide.
template <class Loc, int NK, bool ML>
__device__ __forceinline__ void rows_loop(const Loc& L, const FilterArgs& A, int total_tiles) {
    const int lane = threadIdx.x & 31;
    int t = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (t >= total_tiles) return;
    typename Loc::Ref cur = L.locate(t);
    float r0 = cur.obj_raw(lane), r1 = cur.obj_raw(lane + 32);
#if VK_ROWS_PERSISTENT
    const int nwarps = gridDim.x * kRowWarps;
    for (;;) {
        const int tn = t + nwarps;
        const bool more = tn < total_tiles;
        typename Loc::Ref nxt = cur;
        float n0 = 0.0f, n1 = 0.0f;
        if (more) {
            nxt = L.locate(tn);
            n0 = nxt.obj_raw(lane);
            n1 = nxt.obj_raw(lane + 32);
        }
        const typename Loc::Src S = L.source(cur);
        const float o0 = (lane < cur.nvalid) ? S.obj_val(r0) : -1.0f;
        const float o1 = (lane + 32 < cur.nvalid) ? S.obj_val(r1) : -1.0f;
        filter_tile_rows<typename Loc::Src, NK, ML>(S, A, cur.b, cur.seg, cur.row0, o0, o1);
        if (!more) break;
        cur = nxt; r0 = n0; r1 = n1; t = tn;
    }
#else
    const typename Loc::Src S = L.source(cur);
    const float o0 = (lane < cur.nvalid) ? S.obj_val(r0) : -1.0f;
    const float o1 = (lane + 32 < cur.nvalid) ? S.obj_val(r1) : -1.0f;
    filter_tile_rows<typename Loc::Src, NK, ML>(S, A, cur.b, cur.seg, cur.row0, o0, o1);
#endif
}

template <class T, int NK, bool ML>
__global__ void __launch_bounds__(32 * kRowWarps, VK_ROWS_BPS)
decode_filter_rows_kernel(const HeadDev H, const FilterArgs A, int total_tiles) {
    rows_loop<LogitLocator<T>, NK, ML>(LogitLocator<T>{H}, A, total_tiles);
}

template <class T, int NK, bool ML>
__global__ void __launch_bounds__(32 * kRowWarps, VK_ROWS_BPS)
filter_pred_rows_kernel(const T* __restrict__ pred, int no, const FilterArgs A, int total_tiles) {
    rows_loop<PredLocator<T>, NK, ML>(PredLocator<T>{pred, no, A.rows, A.segs}, A, total_tiles);
}

// ---------------------------------------------------------------------------------------
// Dense fused filter without shared memory or barriers ("lanes = rows"), the default at eval thresholds.
// NCHW conv outputs are already laid out the way a warp wants to read them: the 64 rows of a tile are contiguous
// inside every channel plane, so lane j reading rows j and j + 32 of one channel is two coalesced 128-byte
// requests and nothing has to be staged or transposed.  One warp owns one tile and makes ONE pass over it,
// channel by channel (8 channels = 16 loads in flight per lane): p = sigmoid(x) * obj, and the lanes whose
// product passes append their candidate at once -- two ballots per class give every one its slot.  That
// writes a tile's candidates class-major instead of row-major, which is allowed: a candidate is (score, id) and
// the consumers order by id (include/vk_b200.h).  Best-class mode keeps a running maximum per row and appends
// once per row.  No staging copies, no swizzled address arithmetic, no block barriers, no scan: ~2500
// instructions per warp and tile where the shared-memory kernel above spends ~5500.
// ---------------------------------------------------------------------------------------
#ifndef VK_LANES_WARPS
#define VK_LANES_WARPS 4
#endif
#ifndef VK_LANES_BPS
#define VK_LANES_BPS 8
#endif
constexpr int kLaneWarps = VK_LANES_WARPS;

template <class T, bool ML>
__global__ void __launch_bounds__(32 * kLaneWarps, VK_LANES_BPS)
decode_filter_lanes_kernel(const HeadDev H, const FilterArgs A, int total_tiles) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int t = blockIdx.x * kLaneWarps + (threadIdx.x >> 5);
    if (t >= total_tiles) return;
    const LogitLocator<T> L{H};
    const LogitTileRef<T> q = L.locate(t);
    const int nc = A.nc, nynx = q.nynx;
    const T* const base = q.base;                         // channel 0, first row of the tile
    const bool v0 = lane < q.nvalid, v1 = lane + 32 < q.nvalid;
    // Rows past the end of a ragged tile read row 0 of the tile instead (no per-load predicate) and carry
    // obj = NaN: every product is NaN and no comparison with it holds.
    const int j0 = v0 ? lane : 0, j1 = v1 ? lane + 32 : 0;
    // objectness (image_proc.py:99); rows at or under the threshold keep obj = 0 and never pass
    const float oa = sigmoidf_vk(ld_elem(base + (size_t)4 * nynx + j0));
    const float ob = sigmoidf_vk(ld_elem(base + (size_t)4 * nynx + j1));
    const float obj0 = v0 ? (oa > A.conf ? oa : 0.0f) : __int_as_float(0x7fc00000);
    const float obj1 = v1 ? (ob > A.conf ? ob : 0.0f) : __int_as_float(0x7fc00000);
    if (ML && !__any_sync(0xffffffffu, obj0 > 0.0f || obj1 > 0.0f) && A.conf >= 0.0f) {   // no row passes :99
        if (lane == 0) {
            A.seg_count[(size_t)q.b * A.segs + q.seg] = 0;
            if (q.seg == 0) A.flags[q.b] = cand_flags(A);
        }
        return;
    }
    const T* src0 = base + (size_t)5 * nynx + j0;         // class plane being fetched, the lane's two rows
    const T* src1 = base + (size_t)5 * nynx + j1;
    uint2* out = reinterpret_cast<uint2*>(A.cand + (size_t)q.b * A.cap) + (size_t)q.seg * A.tile_cap;
    asm volatile("" : "+l"(out));                         // one base register (not re-derived per store)
    const uint32_t id0 = (uint32_t)((q.row0 + lane) * nc), id1 = (uint32_t)((q.row0 + lane + 32) * nc);
    uint32_t run = 0;                                     // candidates of the tile so far (warp-uniform)
    bool any0 = false, any1 = false;                      // the lane's rows produced a candidate
    float bv0 = -INFINITY, bv1 = -INFINITY;               // best class (ML == false)
    int bj0 = 0x7fffffff, bj1 = 0x7fffffff;
    // (the loads of the next 8 channels are issued before the current 8 are evaluated; the planes are walked
    // by pointer increments, classes past nc re-read the last plane and are masked out)
    float x0[8], x1[8];
    auto fetch = [&](int c0, float* a, float* b) {
        if (c0 + 8 <= nc) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a[u] = ld_elem(src0); b[u] = ld_elem(src1);
                src0 += nynx; src1 += nynx;
            }
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool in = c0 + u < nc;
                a[u] = in ? ld_elem(src0) : 0.0f; b[u] = in ? ld_elem(src1) : 0.0f;
                src0 += nynx; src1 += nynx;
            }
        }
    };
    fetch(0, x0, x1);
    uint32_t allowed = 0xffffffffu;                       // class filter bits of the current 32 classes (:151)
    unsigned rows0 = 0, rows1 = 0;                        // ballots of the rows that produced candidates
    for (int c0 = 0; c0 < nc; c0 += 8) {
        float n0[8], n1[8];
        fetch(c0 + 8, n0, n1);
        if ((c0 & 31) == 0) {
            allowed = (ML && A.class_mask) ? __ldg(A.class_mask + (c0 >> 5)) : 0xffffffffu;
            if (nc - c0 < 32) allowed &= (1u << (nc - c0)) - 1u;        // classes past nc never pass
        }
        const uint32_t al = (allowed >> (c0 & 31)) & 0xffu;
        if (ML) {
            // Two ballots per class give every passing lane its slot; the stores are predicated instructions (a
            // branch around each costs more than the store).  Measured alternatives, both slower on the eval
            // workload (12% of all (row, class) pairs pass): skipping classes of a tile whose logits all sit under
            // logit(conf / obj) (almost no class of 64 rows is skippable), and one warp scan per 8 classes with
            // lane-private runs of slots (scattered 8-byte stores).  POPC shares the XU pipe with the sigmoid's EX2
            // and RCP and that pipe is this kernel's limiter (77% busy with four POPCs per class), hence the shuffle.
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float p0 = __fmul_rn(sigmoidf_vk(x0[u]), obj0);                            // image_proc.py:135
                const float p1 = __fmul_rn(sigmoidf_vk(x1[u]), obj1);
                const bool ok = (al >> u) & 1u;
                const bool f0 = ok && p0 > A.conf, f1 = ok && p1 > A.conf;                       // :141
                const unsigned b0 = __ballot_sync(0xffffffffu, f0), b1 = __ballot_sync(0xffffffffu, f1);
                // (the totals come from lane 31 by one shuffle, not from two more POPCs: see below)
                const uint32_t r0 = __popc(b0 & lt), r1 = __popc(b1 & lt);
                const uint32_t tot = __shfl_sync(0xffffffffu, (r0 + (uint32_t)f0) | ((r1 + (uint32_t)f1) << 16), 31);
                const uint32_t n_b0 = run + (tot & 0xffffu);
                st_pred_u2(out + (run + r0), __float_as_uint(p0), id0 + (uint32_t)(c0 + u), f0);
                st_pred_u2(out + (n_b0 + r1), __float_as_uint(p1), id1 + (uint32_t)(c0 + u), f1);
                run = n_b0 + (tot >> 16);
                rows0 |= b0; rows1 |= b1;
            }
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float p0 = __fmul_rn(sigmoidf_vk(x0[u]), obj0);
                const float p1 = __fmul_rn(sigmoidf_vk(x1[u]), obj1);
                const bool ok = (al >> u) & 1u;
                if (ok && p0 > bv0) { bv0 = p0; bj0 = c0 + u; }                                  // first max (:145)
                if (ok && p1 > bv1) { bv1 = p1; bj1 = c0 + u; }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { x0[u] = n0[u]; x1[u] = n1[u]; }
    }
    if (ML) {
        any0 = (rows0 >> lane) & 1u; any1 = (rows1 >> lane) & 1u;
    } else {
        any0 = v0 && bj0 != 0x7fffffff && bv0 > A.conf && class_allowed(A.class_mask, bj0);      // :147,151
        any1 = v1 && bj1 != 0x7fffffff && bv1 > A.conf && class_allowed(A.class_mask, bj1);
        const unsigned b0 = __ballot_sync(0xffffffffu, any0), b1 = __ballot_sync(0xffffffffu, any1);
        if (any0) out[__popc(b0 & lt)] = make_uint2(__float_as_uint(bv0), id0 + (uint32_t)bj0);
        if (any1) out[__popc(b0) + __popc(b1 & lt)] = make_uint2(__float_as_uint(bv1), id1 + (uint32_t)bj1);
        run = (uint32_t)(__popc(b0) + __popc(b1));
    }
    // boxes of the rows that produced candidates
    if (run) {
        const PlaneGeom geom{H.variant, H.nx[q.l], q.s0, H.stride[q.l], H.anchors[q.l][2 * q.a], H.anchors[q.l][2 * q.a + 1]};
        float4* boxes = A.boxes + (size_t)q.b * A.rows + q.row0;
        if (any0) {
            const T* bp = base + lane;
            boxes[lane] = geom.box(ld_elem(bp), ld_elem(bp + nynx), ld_elem(bp + 2 * (size_t)nynx), ld_elem(bp + 3 * (size_t)nynx), q.s0 + lane);
        }
        if (any1) {
            const T* bp = base + lane + 32;
            boxes[lane + 32] = geom.box(ld_elem(bp), ld_elem(bp + nynx), ld_elem(bp + 2 * (size_t)nynx), ld_elem(bp + 3 * (size_t)nynx), q.s0 + lane + 32);
        }
    }
    if (lane == 0) {
        A.seg_count[(size_t)q.b * A.segs + q.seg] = (int)run;
        if (q.seg == 0) A.flags[q.b] = cand_flags(A);
        if (run) atomicAdd(A.counts + q.b, (int)run);
    }
}

// ---------------------------------------------------------------------------------------
// Multi-label dense filter, two phases ("pairs" kernel; the default at eval thresholds; VK_FILTER_DENSE_ONEPASS
// selects the lanes kernel above instead).
// The one-phase loop spends two MUFUs and two POPCs per (row, class) pair although only ~1 pair in 8 becomes a
// candidate at eval thresholds (XU pipe 61% busy, 4900 issue slots per tile).  Here the transcendentals run on
// candidates only:
//   (1) lane j owns rows 2j and 2j+1 of the tile (one 8-byte load per class plane).  Every logit takes a conservative
//       pre-test in the logit domain, !(x <= logit(conf / obj) - margin): one FSET + one LOP3 into a 16-bit mask per
//       8 classes, threshold computed once per row (`logit_floor`).  Survivors go to a per-warp shared-memory list
//       (logit, class << 6 | row) in lane-private runs: one POPC and one warp scan per 8 classes.
//   (2) while the list holds 32 entries, a dense round takes the last 32: every lane computes the exact
//       p = sigmoid(x) * obj > conf of image_proc.py:135,141 for one entry and one ballot gives the passing lanes their
//       slots.  The order of a tile's candidates is free (include/vk_b200.h: consumers order by id).
// Every stored bit comes from the same `sigmoidf_vk` as in the other kernels; the pre-test only has to be a superset
// of the exact test.
// ---------------------------------------------------------------------------------------
#ifndef VK_PAIRS_BPS
#define VK_PAIRS_BPS 8
#endif
constexpr int kPairList = 16 * 32 + 32;       // one 8-class step appends at most 512 entries to at most 31 left over

// Largest logit that certainly fails `fl(sigmoidf_vk(x) * obj) > conf`, minus a margin (needs obj > conf > 0).
//   pass  =>  sigma(x) (1 + d) obj (1 + 2^-24) > conf with |d| <= 4e-7 (vk_common.cuh)  =>  sigma(x) > (conf / obj)(1 - 5e-7)
//         =>  sigma(x) > s := min(fl(fl(conf / obj) * 0.999996f), 0.999f)  =>  x > logit(s) = -ln(1 / s - 1).
// Computed value: u = fl(1 / s) - 1 >= 1.001e-3 carries a relative error <= 6e-8 / (1 - s) + 6e-8 <= 6.1e-5, lg2.approx
// adds <= 2^-22 absolute and the product with ln 2 a relative 6e-8 (|logit| <= 104): the computed logit is within
// 1.1e-4 of the true one; the margin is 1e-3.  s == 0 (conf / obj underflows) gives u = +inf and a floor of -inf:
// every logit goes on to the exact test.
__device__ __forceinline__ float logit_floor(float conf, float obj) {
    const float s = fminf(__fmul_rn(__fdiv_rn(conf, obj), 0.999996f), 0.999f);
    const float u = __fsub_rn(__fdiv_rn(1.0f, s), 1.0f);
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u));
    return __fsub_rn(__fmul_rn(l, -0.6931471805599453f), 1e-3f);
}
// m | (!(x <= thr) ? bit : 0) as FSETP + predicated LOP3 (a NaN threshold passes every logit, a NaN logit goes on
// to the exact test and fails it there)
__device__ __forceinline__ void pretest_bit(float x, float thr, uint32_t bit, uint32_t& m) {
    asm("{\n\t.reg .pred p;\n\tsetp.gtu.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(m) : "f"(x), "f"(thr), "r"(bit));
}
// inclusive warp scan, two instructions per level (the shuffle's own predicate guards the add)
__device__ __forceinline__ int warp_incl_scan_p(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
        asm("{\n\t.reg .s32 t;\n\t.reg .pred p;\n\tshfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t@p add.s32 %0, %0, t;\n\t}"
            : "+r"(v) : "r"(o));
    return v;
}
template <class T> struct Pair;
template <> struct Pair<float> {
    typedef float2 type;
    static __device__ __forceinline__ float2 ld(const float2* p) { return __ldg(p); }
};
template <> struct Pair<__half> {
    typedef __half2 type;
    static __device__ __forceinline__ float2 ld(const __half2* p) { return __half22float2(__ldg(p)); }
};
template <> struct Pair<__nv_bfloat16> {
    typedef __nv_bfloat162 type;
    static __device__ __forceinline__ float2 ld(const __nv_bfloat162* p) { return __bfloat1622float2(__ldg(p)); }
};

// shared memory of one warp of the pairs kernel, addressed with explicit 32-bit shared addresses (through generic
// pointers the compiler rebuilds the shared window base, an S2R + LEA, inside the loops)
constexpr int kPairX = 0;                                  // logits that passed the pre-test
constexpr int kPairC = kPairX + 4 * kPairList;             // ... their class << 6 | row of the tile
constexpr int kPairObj = kPairC + 4 * kPairList;           // objectness of the 64 rows (0: failed :99, NaN: past the tile)
constexpr int kPairHit = kPairObj + 4 * kTileS;            // rows that produced a candidate (bytes)
constexpr int kPairBytes = kPairHit + kTileS;
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

// VEC: every plane has an even number of rows and the level pointers are aligned for two-element loads (the host
// checks): one load per pair.  Otherwise (odd grids such as 13x13, 19x19, 21x21) the two rows are loaded separately.
template <class T, bool VEC>
__global__ void __launch_bounds__(32 * kLaneWarps, VK_PAIRS_BPS)
decode_filter_pairs_kernel(const HeadDev H, const FilterArgs A, int total_tiles) {
    typedef typename Pair<T>::type T2;
    __shared__ __align__(16) uint8_t s_warp[kLaneWarps][kPairBytes];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int t = blockIdx.x * kLaneWarps + warp;
    if (t >= total_tiles) return;
    const LogitLocator<T> L{H};
    const LogitTileRef<T> q = L.locate(t);
    const int nc = A.nc;
    const int nynx = q.nynx;
    const bool v0 = 2 * lane < q.nvalid, v1 = 2 * lane + 1 < q.nvalid;   // (VEC: nvalid is even, v0 == v1)
    // Rows past the end of a ragged tile read row 0 of the tile instead (no per-load predicate) and carry obj = NaN.
    const int j0 = v0 ? 2 * lane : 0;
    const int d1 = v1 ? 1 : -j0;                          // the second row, relative to the first
    const T* const base = q.base + j0;                    // channel 0, the lane's first row
    auto ld2 = [&](const T* p) {
        if (VEC) return Pair<T>::ld(reinterpret_cast<const T2*>(p));
        return make_float2(ld_elem(p), ld_elem(p + d1));
    };
    uint32_t sb = (uint32_t)__cvta_generic_to_shared(s_warp[warp]);
    asm volatile("" : "+r"(sb));                          // one register, not re-derived
    // objectness (image_proc.py:99); rows at or under the threshold keep obj = 0 and never pass
    const float2 o = ld2(base + (size_t)4 * nynx);
    const float oa = sigmoidf_vk(o.x), ob = sigmoidf_vk(o.y);
    const float obj0 = v0 ? (oa > A.conf ? oa : 0.0f) : __int_as_float(0x7fc00000);
    const float obj1 = v1 ? (ob > A.conf ? ob : 0.0f) : __int_as_float(0x7fc00000);
    if (!__any_sync(0xffffffffu, obj0 > 0.0f || obj1 > 0.0f) && A.conf >= 0.0f) {         // no row passes :99
        if (lane == 0) {
            A.seg_count[(size_t)q.b * A.segs + q.seg] = 0;
            if (q.seg == 0) A.flags[q.b] = cand_flags(A);
        }
        return;
    }
    float thr0 = INFINITY, thr1 = INFINITY;               // rows that failed :99 (or lie past the tile) never pass
    if (A.conf > 0.0f) {
        if (obj0 > 0.0f) thr0 = logit_floor(A.conf, obj0);
        if (obj1 > 0.0f) thr1 = logit_floor(A.conf, obj1);
    } else {                                              // conf <= 0: even a product of 0 may pass, test every logit
        if (v0) thr0 = __int_as_float(0x7fc00000);
        if (v1) thr1 = __int_as_float(0x7fc00000);
    }
    sts32(sb + kPairObj + 8 * lane, __float_as_uint(obj0));
    sts32(sb + kPairObj + 8 * lane + 4, __float_as_uint(obj1));
    if (lane < kTileS / 4) sts32(sb + kPairHit + 4 * lane, 0u);
    __syncwarp();
    uint2* out = reinterpret_cast<uint2*>(A.cand + (size_t)q.b * A.cap) + (size_t)q.seg * A.tile_cap;
    asm volatile("" : "+l"(out));                         // one base register (not re-derived per store)
    uint32_t id0 = (uint32_t)q.row0 * (uint32_t)nc;
    asm volatile("" : "+r"(id0));
    uint32_t run = 0;                                     // candidates of the tile so far (warp-uniform)
    int fill = 0;                                         // entries waiting in the list (warp-uniform)
    // one dense round over list entries [at, at + 32)
    auto settle = [&](int at, bool active) {
        const uint32_t e = sb + 4u * (uint32_t)(at + lane);
        const float x = active ? __uint_as_float(lds32(e + kPairX)) : 0.0f;
        const uint32_t code = active ? lds32(e + kPairC) : 0u;
        const uint32_t r = code & 63u;
        const float p = __fmul_rn(sigmoidf_vk(x), __uint_as_float(lds32(sb + kPairObj + 4u * r)));      // image_proc.py:135
        const bool f = active && p > A.conf;                                                     // :141
        const unsigned b = __ballot_sync(0xffffffffu, f);
        if (f) {
            out[run + (uint32_t)__popc(b & lt)] = make_uint2(__float_as_uint(p), id0 + r * (uint32_t)nc + (code >> 6));
            asm volatile("st.shared.u8 [%0], %1;" :: "r"(sb + kPairHit + r), "r"(1u) : "memory");
        }
        run += __popc(b);
    };
    const T* src = base + (size_t)5 * nynx;               // class plane being fetched
    float2 x[8];
    auto fetch = [&](int c0, float2* a) {
        if (c0 + 8 <= nc) {
#pragma unroll
            for (int u = 0; u < 8; ++u) { a[u] = ld2(src); src += nynx; }
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a[u] = c0 + u < nc ? ld2(src) : make_float2(0.0f, 0.0f);
                src += nynx;
            }
        }
    };
    // one step: the next 8 planes are requested, the current 8 tested, their survivors listed, full rounds settled
    auto step = [&](int c0, const float2* cur, float2* nxt) {
        fetch(c0 + 8, nxt);
        uint32_t m = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            pretest_bit(cur[u].x, thr0, 1u << (2 * u), m);
            pretest_bit(cur[u].y, thr1, 2u << (2 * u), m);
        }
        if (A.class_mask != nullptr || c0 + 8 > nc) {     // class filter (:151) and the classes past nc: two bits each
            uint32_t al = A.class_mask ? (__ldg(A.class_mask + (c0 >> 5)) >> (c0 & 31)) & 0xffu : 0xffu;
            if (nc - c0 < 8) al &= (1u << (nc - c0)) - 1u;
            al = (al | (al << 4)) & 0x0f0fu;
            al = (al | (al << 2)) & 0x3333u;
            al = (al | (al << 1)) & 0x5555u;
            m &= al | (al << 1);
        }
        const int cnt = __popc(m);
        const int incl = warp_incl_scan_p(cnt);
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        if (tot) {
            uint32_t w = sb + 4u * (uint32_t)(fill + incl - cnt);
            uint32_t cb = ((uint32_t)c0 << 6) | (uint32_t)(2 * lane);
            asm volatile("" : "+r"(cb));                  // (the compiler re-derives it from %tid per entry otherwise)
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (m & (1u << (2 * u))) { sts32(w + kPairX, __float_as_uint(cur[u].x)); sts32(w + kPairC, cb + (uint32_t)(u << 6)); w += 4; }
                if (m & (2u << (2 * u))) { sts32(w + kPairX, __float_as_uint(cur[u].y)); sts32(w + kPairC, cb + (uint32_t)(u << 6) + 1u); w += 4; }
            }
            fill += tot;
            __syncwarp();
            while (fill >= 32) { fill -= 32; settle(fill, true); }      // the last 32: what stays needs no move
            __syncwarp();
        }
    };
    float2 y[8];
    fetch(0, x);
    for (int c0 = 0; c0 < nc; c0 += 16) {                 // (two steps per trip: the buffers swap without copies)
        step(c0, x, y);
        if (c0 + 8 < nc) step(c0 + 8, y, x);
    }
    if (fill) settle(0, lane < fill);
    __syncwarp();
    uint32_t hit;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(hit) : "r"(sb + kPairHit + 2 * lane) : "memory");
    // boxes of the rows that produced candidates
    if (hit) {
        const PlaneGeom geom{H.variant, H.nx[q.l], q.s0, H.stride[q.l], H.anchors[q.l][2 * q.a], H.anchors[q.l][2 * q.a + 1]};
        float4* boxes = A.boxes + (size_t)q.b * A.rows + q.row0 + 2 * lane;
        const float2 l0 = ld2(base), l1 = ld2(base + nynx), l2 = ld2(base + 2 * (size_t)nynx), l3 = ld2(base + 3 * (size_t)nynx);
        if (hit & 0xffu) boxes[0] = geom.box(l0.x, l1.x, l2.x, l3.x, q.s0 + 2 * lane);
        if (hit >> 8) boxes[1] = geom.box(l0.y, l1.y, l2.y, l3.y, q.s0 + 2 * lane + 1);
    }
    if (lane == 0) {
        A.seg_count[(size_t)q.b * A.segs + q.seg] = (int)run;
        if (q.seg == 0) A.flags[q.b] = cand_flags(A);
        if (run) atomicAdd(A.counts + q.b, (int)run);
    }
}

// ---------------------------------------------------------------------------------------
// Dense variant of filter_pred (the `nms(prediction)` drop-in at eval thresholds): the same
// persistent pipeline and (row, class part) mapping as decode_filter_dense_kernel, reading an
// existing (B, rows, no) prediction tensor.  A tile is 64 consecutive rows = one contiguous run of
// 64*no floats, copied as it is ([row][no], odd pitch: lanes over rows are conflict-free); the
// values are probabilities already, boxes are cxcywh.
// ---------------------------------------------------------------------------------------
constexpr int kParts = kDecThreads / kTileS;   // 4 class parts per row

template <class T, int I, int N>
struct PredProducts {
    static __device__ __forceinline__ void run(float* p, uint32_t base, float obj) {
        float x;
        if (sizeof(T) == 4) {
            asm("ld.shared.f32 %0, [%1+%2];" : "=f"(x) : "r"(base), "n"(I * 4));
        } else {
            unsigned short h;
            asm("ld.shared.u16 %0, [%1+%2];" : "=h"(h) : "r"(base), "n"(I * 2));
            x = to_f32(*reinterpret_cast<const T*>(&h));
        }
        p[I] = __fmul_rn(x, obj);                                                                 // image_proc.py:135
        PredProducts<T, I + 1, N>::run(p, base, obj);
    }
};
template <class T, int N>
struct PredProducts<T, N, N> {
    static __device__ __forceinline__ void run(float*, uint32_t, float) {}
};

// nelem contiguous elements of one tile of prediction rows -> shared memory (16-byte copies when aligned)
template <class T>
__device__ __forceinline__ void pred_prefetch(const T* __restrict__ src, int nelem, T* tile) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tile);
    constexpr int per16 = 16 / (int)sizeof(T);
    if (((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((nelem & (per16 - 1)) == 0)) {
        for (int e = threadIdx.x; per16 * e < nelem; e += kDecThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(base + 16u * e), "l"(src + per16 * e) : "memory");
    } else if (sizeof(T) == 4) {
        for (int e = threadIdx.x; e < nelem; e += kDecThreads)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(base + 4u * e), "l"(src + e) : "memory");
    } else {
        for (int e = threadIdx.x; e < nelem; e += kDecThreads) tile[e] = src[e];
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <class T, int CPP, bool ML>
__global__ void __launch_bounds__(kDecThreads, VK_DENSE_BPS)
filter_pred_dense_kernel(const T* __restrict__ pred, int no, const FilterArgs A, int total_tiles) {
    extern __shared__ __align__(16) unsigned char tiles_raw[];  // 2 x [kTileS][no] prediction rows
    T* const tiles_sm = reinterpret_cast<T*>(tiles_raw);
    __shared__ int s_cnt[kDecThreads];
    __shared__ int s_off[kDecThreads + 1];
    __shared__ int s_wsum[kWarps];
    __shared__ float s_bv[kDecThreads];
    __shared__ int s_bj[kDecThreads];
    const int nc = A.nc;
    const int tile_floats = (kTileS * no + 7) & ~7;        // elements per buffer, 16-byte multiple
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int row = threadIdx.x & (kTileS - 1), qd = threadIdx.x >> 6;
    const int cpp = (nc + kParts - 1) / kParts;
    const int c_lo = min(nc, qd * cpp), c_hi = min(nc, c_lo + cpp);
    const int slot = row * kParts + qd;
    uint32_t allowed = 0;
    for (int i = 0; i < c_hi - c_lo; ++i)
        if (class_allowed(A.class_mask, c_lo + i)) allowed |= 1u << i;
    int t = blockIdx.x;
    if (t >= total_tiles) return;
    // tile t = (image b, segment seg): rows seg*64 .. of image b
    auto tile_src = [&](int tt, int& b, int& seg, int& nvalid) {
        b = tt / A.segs; seg = tt - b * A.segs;
        nvalid = min(kTileS, A.rows - seg * kTileS);
        return pred + ((size_t)b * A.rows + (size_t)seg * kTileS) * no;
    };
    int b, seg, nvalid;
    {
        const T* src = tile_src(t, b, seg, nvalid);
        pred_prefetch<T>(src, nvalid * no, tiles_sm);
    }
    for (int k = 0; t < total_tiles; ++k, t += gridDim.x) {
        const T* tile = tiles_sm + (k & 1) * tile_floats;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                      // tile k landed; everyone is done with the other buffer and the scratch arrays
        if (t + (int)gridDim.x < total_tiles) {
            int b2, seg2, nv2;
            const T* src = tile_src(t + gridDim.x, b2, seg2, nv2);
            pred_prefetch<T>(src, nv2 * no, tiles_sm + ((k + 1) & 1) * tile_floats);
        }
        tile_src(t, b, seg, nvalid);
        const T* prow = tile + row * no;
        const float o = to_f32(prow[4]);
        const float obj = (row < nvalid && o > A.conf) ? o : 0.0f;    // image_proc.py:99 (dead rows: products 0)
        float p[CPP];
        PredProducts<T, 0, CPP>::run(p, (uint32_t)__cvta_generic_to_shared(prow + 5 + c_lo), obj);
        uint32_t flags = 0;
        float bv = -INFINITY;
        int bj = 0x7fffffff;
        if (ML) {
#pragma unroll
            for (int i = 0; i < CPP; ++i)
                if (p[i] > A.conf) flags |= 1u << i;                                              // :141
        } else {
            const int ncls = c_hi - c_lo;
#pragma unroll
            for (int i = 0; i < CPP; ++i)
                if (i < ncls && p[i] > bv) { bv = p[i]; bj = c_lo + i; }                          // first max (:145)
        }
        int count;
        if (ML) {
            flags &= allowed;
            count = __popc(flags);
        } else {
            s_bv[slot] = bv; s_bj[slot] = bj;
            __syncthreads();
            count = 0;
            if (qd == 0) {
#pragma unroll
                for (int q2 = 1; q2 < kParts; ++q2)
                    if (s_bv[slot + q2] > bv) { bv = s_bv[slot + q2]; bj = s_bj[slot + q2]; }
                count = (bj != 0x7fffffff && bv > A.conf && class_allowed(A.class_mask, bj)) ? 1 : 0;   // :147,151
            }
        }
        s_cnt[slot] = count;
        __syncthreads();
        {
            const int v = s_cnt[threadIdx.x];
            const int inc = warp_incl_scan(v, lane);
            s_off[threadIdx.x] = inc - v;
            if (lane == 31) s_wsum[w] = inc;
        }
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int i = 0; i < kWarps; ++i) {
            const int x = s_wsum[i];
            if (i < (slot >> 5)) wbase += x;
            total += x;
        }
        const int tile_base = seg * A.tile_cap;
        const int grow = seg * kTileS + row;                                   // row inside the image
        if (count) {
            uint2* const wp = reinterpret_cast<uint2*>(A.cand + (size_t)b * A.cap) + tile_base + wbase + s_off[slot];
            const uint32_t idx0 = (uint32_t)(grow * nc + c_lo);
            if (ML) {
                uint32_t pos = 0;
#pragma unroll
                for (int i = 0; i < CPP; ++i) {
                    const bool f = (flags & (1u << i)) != 0;
                    if (f) wp[pos] = make_uint2(__float_as_uint(p[i]), idx0 + (uint32_t)i);
                    pos += f;
                }
            } else {
                *wp = make_uint2(__float_as_uint(bv), (uint32_t)(grow * nc + bj));
            }
        }
        if (qd == 0) {
            const int n = s_cnt[slot] + s_cnt[slot + 1] + s_cnt[slot + 2] + s_cnt[slot + 3];
            if (n > 0) A.boxes[(size_t)b * A.rows + grow] = xyxy_from_cxcywh(to_f32(prow[0]), to_f32(prow[1]), to_f32(prow[2]), to_f32(prow[3]));
        }
        if (threadIdx.x == 0) {
            A.seg_count[(size_t)b * A.segs + seg] = total;
            if (seg == 0) A.flags[b] = cand_flags(A);
            if (total) atomicAdd(A.counts + b, total);
        }
    }
}
}  // namespace vk

using namespace vk;

extern "C" int vk_head_rows(const VkHeadCfg* cfg) {
    HeadDev H;
    if (int rc = make_head(cfg, &H, "vk_head_rows")) return rc;
    return H.rows;
}

extern "C" int vk_decode_filter_segments(const VkHeadCfg* cfg) {
    HeadDev H;
    if (int rc = make_head(cfg, &H, "vk_decode_filter_segments")) return rc;
    return H.tiles;
}

extern "C" int vk_cand_tile_slots(int nc, int multi_label) { return kTileS * ((multi_label && nc > 1) ? nc : 1); }

extern "C" int vk_filter_segments(int rows) { return rows > 0 ? ceil_div(rows, kTileS) : 0; }

// dtype dispatch: F(float) / F(__half) / F(__nv_bfloat16)
#define VK_BY_DTYPE(dtype, F)                       \
    do {                                            \
        if ((dtype) == VK_F32) { F(float); }        \
        else if ((dtype) == VK_F16) { F(__half); }  \
        else { F(__nv_bfloat16); }                  \
    } while (0)

static bool bad_dtype(int dtype) { return dtype != VK_F32 && dtype != VK_F16 && dtype != VK_BF16; }
static int elem_size(int dtype) { return dtype == VK_F32 ? 4 : 2; }

template <class T, int NCG>
static int launch_detect_decode(const HeadDev& H, float* pred, int total_tiles, int have_lin, cudaStream_t stream) {
    const size_t tile_bytes = sizeof(T) == 4 ? (size_t)H.no * kTileS * 4 : (((size_t)H.no * (kTileS + 4) * 2 + 15) & ~(size_t)15);
    const size_t smem = 2 * tile_bytes + (have_lin ? (size_t)kTileS * H.no * sizeof(float) : 0);
    if (smem > 200 * 1024) return fail_code(VK_E_LIMIT, "vk_detect_decode: nc=%d needs %zu B of shared memory", H.nc, smem);
    const void* fn = reinterpret_cast<const void*>(&detect_decode_kernel<T, NCG>);
    if (int rc = ensure_dyn_smem(fn, smem, "vk_detect_decode")) return rc;
    int per_sm = blocks_per_sm(fn, kDecThreads, smem);
    if (per_sm > 3) per_sm = 3;                   // more concurrent tile streams cost DRAM locality
    const int grid = total_tiles < per_sm * kNumSMs ? total_tiles : per_sm * kNumSMs;   // persistent: every block is resident
    detect_decode_kernel<T, NCG><<<grid, kDecThreads, smem, stream>>>(H, pred, total_tiles, have_lin);
    count_launch();
    return check_launch("detect_decode_kernel");
}

extern "C" int vk_detect_decode(const VkHeadCfg* cfg, const void* const* levels, int dtype, int batch,
                                float* pred, float* const* raw, vk_stream_t stream) {
    HeadDev H;
    if (int rc = make_head(cfg, &H, "vk_detect_decode")) return rc;
    if (batch == 0) return VK_OK;
    if (!levels || !pred || batch < 0) return fail_arg("vk_detect_decode: null/negative argument");
    if (bad_dtype(dtype)) return fail_arg("vk_detect_decode: dtype %d", dtype);
    if (batch > 65535) return fail_code(VK_E_LIMIT, "vk_detect_decode: batch %d > 65535", batch);
    for (int l = 0; l < H.nl; ++l) {
        if (!levels[l]) return fail_arg("vk_detect_decode: level %d is NULL", l);
        if (reinterpret_cast<uintptr_t>(levels[l]) & (elem_size(dtype) - 1)) return fail_arg("vk_detect_decode: level %d is misaligned", l);
        H.lv[l] = levels[l];
        H.raw[l] = raw ? raw[l] : nullptr;
    }
    const int have_lin = raw != nullptr;
    if ((long)H.tiles * batch > 0x7fffffffL) return fail_code(VK_E_LIMIT, "vk_detect_decode: %d x %d tiles", H.tiles, batch);
    const int total_tiles = H.tiles * batch;
#define VK_DEC_T(T)                                                                                          \
    do {                                                                                                     \
        if (H.no <= 32) return launch_detect_decode<T, 1>(H, pred, total_tiles, have_lin, as_stream(stream));       \
        if (H.no <= 64) return launch_detect_decode<T, 2>(H, pred, total_tiles, have_lin, as_stream(stream));       \
        if (H.no <= 96) return launch_detect_decode<T, 3>(H, pred, total_tiles, have_lin, as_stream(stream));       \
        if (H.no <= 128) return launch_detect_decode<T, 4>(H, pred, total_tiles, have_lin, as_stream(stream));      \
        return launch_detect_decode<T, 0>(H, pred, total_tiles, have_lin, as_stream(stream));                       \
    } while (0)
    VK_BY_DTYPE(dtype, VK_DEC_T);
#undef VK_DEC_T
    return VK_OK;
}

// ---- filter launches.  `dense` = whole tiles staged in shared memory (persistent); else one warp per tile.
static bool pick_dense(int kernel, float conf_thres) {
    return kernel == VK_FILTER_DENSE || kernel == VK_FILTER_DENSE_ONEPASS || (kernel == VK_FILTER_AUTO && conf_thres < 0.05f);
}

template <class T, bool ML>
static int launch_decode_filter_lanes(const HeadDev& H, const FilterArgs& A, int total_tiles, cudaStream_t stream) {
    decode_filter_lanes_kernel<T, ML><<<ceil_div(total_tiles, kLaneWarps), 32 * kLaneWarps, 0, stream>>>(H, A, total_tiles);
    count_launch();
    return check_launch("decode_filter_lanes_kernel");
}

// The pairs kernel loads two rows at once: every plane needs an even number of rows and 2-element alignment.
static bool pairs_ok(const HeadDev& H, size_t elem) {
    for (int l = 0; l < H.nl; ++l)
        if ((H.nynx[l] & 1) || (reinterpret_cast<uintptr_t>(H.lv[l]) & (2 * elem - 1))) return false;
    return true;
}
template <class T>
static int launch_decode_filter_pairs(const HeadDev& H, const FilterArgs& A, int total_tiles, cudaStream_t stream) {
    const int grid = ceil_div(total_tiles, kLaneWarps);
    if (pairs_ok(H, sizeof(T))) decode_filter_pairs_kernel<T, true><<<grid, 32 * kLaneWarps, 0, stream>>>(H, A, total_tiles);
    else decode_filter_pairs_kernel<T, false><<<grid, 32 * kLaneWarps, 0, stream>>>(H, A, total_tiles);
    count_launch();
    return check_launch("decode_filter_pairs_kernel");
}

template <class T, int NK, bool ML>
static int launch_decode_filter_rows(const HeadDev& H, const FilterArgs& A, int total_tiles, cudaStream_t stream) {
    int grid = ceil_div(total_tiles, kRowWarps);
#if VK_ROWS_PERSISTENT
    grid = min(grid, blocks_per_sm(reinterpret_cast<const void*>(&decode_filter_rows_kernel<T, NK, ML>), 32 * kRowWarps, 0) * kNumSMs);
#endif
    decode_filter_rows_kernel<T, NK, ML><<<grid, 32 * kRowWarps, 0, stream>>>(H, A, total_tiles);
    count_launch();
    return check_launch("decode_filter_rows_kernel");
}

extern "C" int vk_decode_filter(const VkHeadCfg* cfg, const void* const* levels, int dtype, int batch,
                                float conf_thres, int multi_label, const uint32_t* class_mask, int kernel,
                                const VkCandBuf* out, vk_stream_t stream_) {
    HeadDev H;
    if (int rc = make_head(cfg, &H, "vk_decode_filter")) return rc;
    if (batch == 0) return VK_OK;
    if (!levels || batch < 0) return fail_arg("vk_decode_filter: null/negative argument");
    if (bad_dtype(dtype)) return fail_arg("vk_decode_filter: dtype %d", dtype);
    if (kernel < VK_FILTER_AUTO || kernel > VK_FILTER_DENSE_ONEPASS) return fail_arg("vk_decode_filter: kernel %d", kernel);
    if (!(conf_thres >= 0.f && conf_thres <= 1.f)) return fail_arg("vk_decode_filter: conf_thres %g outside [0,1]", conf_thres);
    if (batch > 65535) return fail_code(VK_E_LIMIT, "vk_decode_filter: batch %d > 65535", batch);
    if ((long)H.tiles * batch > 0x7fffffffL) return fail_code(VK_E_LIMIT, "vk_decode_filter: %d x %d tiles", H.tiles, batch);
    if (int rc = check_cand(out, H.rows, H.tiles, H.nc, multi_label, "vk_decode_filter")) return rc;
    for (int l = 0; l < H.nl; ++l) {
        if (!levels[l]) return fail_arg("vk_decode_filter: level %d is NULL", l);
        if (reinterpret_cast<uintptr_t>(levels[l]) & (elem_size(dtype) - 1)) return fail_arg("vk_decode_filter: level %d is misaligned", l);
        H.lv[l] = levels[l];
    }
    cudaStream_t stream = as_stream(stream_);
    if (int rc = reset_cand(out, batch, "vk_decode_filter", stream)) return rc;
    const FilterArgs A = make_filter_args(out, batch, conf_thres, multi_label, class_mask);
    const int total_tiles = H.tiles * batch;
    const bool ml = A.multi_label != 0;
    if (pick_dense(kernel, conf_thres)) {
#define VK_DL_T(T)                                                                                  \
        if (ml && kernel != VK_FILTER_DENSE_ONEPASS) return launch_decode_filter_pairs<T>(H, A, total_tiles, stream); \
        return ml ? launch_decode_filter_lanes<T, true>(H, A, total_tiles, stream)                   \
                  : launch_decode_filter_lanes<T, false>(H, A, total_tiles, stream)
        VK_BY_DTYPE(dtype, VK_DL_T);
#undef VK_DL_T
    }
#define VK_DR_NK(T, NK) return ml ? launch_decode_filter_rows<T, NK, true>(H, A, total_tiles, stream) \
                                  : launch_decode_filter_rows<T, NK, false>(H, A, total_tiles, stream)
#define VK_DR_T(T)                                  \
    do {                                            \
        if (H.no <= 32) { VK_DR_NK(T, 1); }         \
        if (H.no <= 64) { VK_DR_NK(T, 2); }         \
        if (H.no <= 96) { VK_DR_NK(T, 3); }         \
        if (H.no <= 128) { VK_DR_NK(T, 4); }        \
        VK_DR_NK(T, 0);                             \
    } while (0)
    VK_BY_DTYPE(dtype, VK_DR_T);
#undef VK_DR_T
#undef VK_DR_NK
    return VK_OK;
}

template <class T, int CPP, bool ML>
static int launch_filter_pred_dense(const void* pred, int no, const FilterArgs& A, int total_tiles, cudaStream_t stream) {
    const size_t dsmem = (2 * (size_t)kTileS * no + (size_t)CPP + 16) * sizeof(T);        // + slack: unrolled class loop
    if (dsmem > 200 * 1024) return fail_code(VK_E_LIMIT, "vk_filter_pred: nc=%d needs %zu B of shared memory", A.nc, dsmem);
    const void* fn = reinterpret_cast<const void*>(&filter_pred_dense_kernel<T, CPP, ML>);
    if (int rc = ensure_dyn_smem(fn, dsmem, "vk_filter_pred")) return rc;
    int per_sm = blocks_per_sm(fn, kDecThreads, dsmem);
    if (per_sm > VK_DENSE_BPS) per_sm = VK_DENSE_BPS;
    const int grid = total_tiles < per_sm * kNumSMs ? total_tiles : per_sm * kNumSMs;
    filter_pred_dense_kernel<T, CPP, ML><<<grid, kDecThreads, dsmem, stream>>>(static_cast<const T*>(pred), no, A, total_tiles);
    count_launch();
    return check_launch("filter_pred_dense_kernel");
}

template <class T, int NK, bool ML>
static int launch_filter_pred_rows(const void* pred, int no, const FilterArgs& A, int total_tiles, cudaStream_t stream) {
    int grid = ceil_div(total_tiles, kRowWarps);
#if VK_ROWS_PERSISTENT
    grid = min(grid, blocks_per_sm(reinterpret_cast<const void*>(&filter_pred_rows_kernel<T, NK, ML>), 32 * kRowWarps, 0) * kNumSMs);
#endif
    filter_pred_rows_kernel<T, NK, ML><<<grid, 32 * kRowWarps, 0, stream>>>(static_cast<const T*>(pred), no, A, total_tiles);
    count_launch();
    return check_launch("filter_pred_rows_kernel");
}

extern "C" int vk_filter_pred(const void* pred, int dtype, int batch, int rows, int nc, float conf_thres,
                              int multi_label, const uint32_t* class_mask, int kernel, const VkCandBuf* out,
                              vk_stream_t stream_) {
    if (batch == 0) return VK_OK;
    if (!pred || batch < 0 || rows <= 0 || nc < 1) return fail_arg("vk_filter_pred: null/negative argument");
    if (bad_dtype(dtype)) return fail_arg("vk_filter_pred: dtype %d", dtype);
    if (reinterpret_cast<uintptr_t>(pred) & (elem_size(dtype) - 1)) return fail_arg("vk_filter_pred: pred is misaligned");
    if (kernel < VK_FILTER_AUTO || kernel > VK_FILTER_DENSE_ONEPASS) return fail_arg("vk_filter_pred: kernel %d", kernel);
    if (!(conf_thres >= 0.f && conf_thres <= 1.f)) return fail_arg("vk_filter_pred: conf_thres %g outside [0,1]", conf_thres);
    if (batch > 65535) return fail_code(VK_E_LIMIT, "vk_filter_pred: batch %d > 65535", batch);
    const int segs = ceil_div(rows, kTileS);
    if (segs > VK_MAX_SEGMENTS) return fail_code(VK_E_LIMIT, "vk_filter_pred: %d rows > %d", rows, VK_MAX_SEGMENTS * kTileS);
    if ((long)segs * batch > 0x7fffffffL) return fail_code(VK_E_LIMIT, "vk_filter_pred: %d x %d tiles", segs, batch);
    if (int rc = check_cand(out, rows, segs, nc, multi_label, "vk_filter_pred")) return rc;
    cudaStream_t stream = as_stream(stream_);
    if (int rc = reset_cand(out, batch, "vk_filter_pred", stream)) return rc;
    const int no = nc + 5;
    const FilterArgs A = make_filter_args(out, batch, conf_thres, multi_label, class_mask);
    const int total_tiles = segs * batch;
    const bool ml = A.multi_label != 0;
    if (pick_dense(kernel, conf_thres) && nc <= 128) {
#define VK_FP_T(T)                                                                                                   \
        do {                                                                                                          \
            if (nc <= 32) return ml ? launch_filter_pred_dense<T, 8, true>(pred, no, A, total_tiles, stream)         \
                                    : launch_filter_pred_dense<T, 8, false>(pred, no, A, total_tiles, stream);       \
            if (nc <= 80) return ml ? launch_filter_pred_dense<T, 20, true>(pred, no, A, total_tiles, stream)        \
                                    : launch_filter_pred_dense<T, 20, false>(pred, no, A, total_tiles, stream);      \
            return ml ? launch_filter_pred_dense<T, 32, true>(pred, no, A, total_tiles, stream)                      \
                      : launch_filter_pred_dense<T, 32, false>(pred, no, A, total_tiles, stream);                    \
        } while (0)
        VK_BY_DTYPE(dtype, VK_FP_T);
#undef VK_FP_T
    }
#define VK_FR_NK(T, NK) return ml ? launch_filter_pred_rows<T, NK, true>(pred, no, A, total_tiles, stream) \
                                  : launch_filter_pred_rows<T, NK, false>(pred, no, A, total_tiles, stream)
#define VK_FR_T(T)                                \
    do {                                          \
        if (no <= 32) { VK_FR_NK(T, 1); }         \
        if (no <= 64) { VK_FR_NK(T, 2); }         \
        if (no <= 96) { VK_FR_NK(T, 3); }         \
        if (no <= 128) { VK_FR_NK(T, 4); }        \
        VK_FR_NK(T, 0);                           \
    } while (0)
    VK_BY_DTYPE(dtype, VK_FR_T);
#undef VK_FR_T
#undef VK_FR_NK
    return VK_OK;
}
