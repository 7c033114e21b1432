// Detect-head decode and the confidence filter (SURVEY.md §8 a3, a4, a5 first half).
//
//   vk_detect_decode   conv outputs (B, na*no, ny, nx) -> pred (B, rows, no): a transposing
//                      copy with the sigmoid/grid/anchor decode fused in.  HBM-bound:
//                      8 568 000 B read + 8 568 000 B written per 640x640 image.
//   vk_decode_filter   the same tiles, but only rows with obj > conf are decoded and only
//                      candidates leave the SM: pred is never materialised.  A tile whose
//                      objectness plane has few survivors gathers just those rows.
//   vk_filter_pred     drop-in filter for an existing pred tensor (`nms(prediction)`).
//
// Candidate order: the reference's candidate list is ordered (row asc, class asc)
// (`nonzero`, utils/image_proc.py:141-143).  Here every tile writes its candidates, in that
// order, into the fixed slot range its tile owns (64 rows x nc slots, or 64 in best-class
// mode) and records (base, count) in a segment table indexed by tile; canonical order =
// segment order x in-segment order.  No tile ever waits for an atomic, the buffer cannot
// overflow, and the consumer (nms.cu) walks the table.
#include "decode_common.cuh"

namespace vk {


struct TileLoc {
    int l, a, s0, nvalid, row0;
};

__device__ __forceinline__ TileLoc locate_tile(const HeadDev& H, int t) {
    TileLoc q;
    int l = 0;
#pragma unroll
    for (int i = 1; i < VK_MAX_LEVELS; ++i)
        if (i < H.nl && t >= H.tile_start[i]) l = i;
    const int rel = t - H.tile_start[l];
    q.l = l;
    q.a = rel / H.tpa[l];
    q.s0 = (rel - q.a * H.tpa[l]) * kTileS;
    q.nvalid = min(kTileS, H.nynx[l] - q.s0);
    q.row0 = H.row_base[l] + q.a * H.nynx[l] + q.s0;
    return q;
}

// ---------------------------------------------------------------------------------------
// materialised decode
// ---------------------------------------------------------------------------------------
// Coalesced load of one tile's logits into shared memory [no][PITCH] (the filter kernels).  Loads are issued in
// batches of four 128-bit requests per thread before the first shared store, so that a block
// keeps ~16 KB in flight instead of one request per thread.
template <int PITCH>
__device__ __forceinline__ void load_tile(float* tile, const float* __restrict__ in, int no, int nynx,
                                          int nvalid, bool vec) {
    if (vec) {
        const int total = no * (kTileS / 4);
        for (int e0 = threadIdx.x; e0 < total; e0 += 4 * kDecThreads) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * kDecThreads;
                const int c = e >> 4, sq = (e & 15) << 2;
                v[u] = (e < total && sq < nvalid) ? ld_stream_f4(in + (size_t)c * nynx + sq)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + u * kDecThreads;
                if (e < total) {
                    float* d = tile + (e >> 4) * PITCH + ((e & 15) << 2);
                    if constexpr ((PITCH & 3) == 0) {          // rows 16-byte aligned: one STS.128
                        *reinterpret_cast<float4*>(d) = v[u];
                    } else {
                        d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
                    }
                }
            }
        }
    } else {
        for (int e = threadIdx.x; e < no * kTileS; e += kDecThreads) {
            const int c = e >> 6, sq = e & 63;
            if (sq < nvalid) tile[c * PITCH + sq] = ld_stream_f32(in + (size_t)c * nynx + sq);
        }
    }
}

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
    const float* src;     // first logit of the tile: channel 0, row s0 of plane (b, a)
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
    d->src = H.lv[l] + ((size_t)(b * H.na + a) * no) * nynx + s0;
    d->out = pred ? pred + ((size_t)b * H.rows + row0) * no : nullptr;
    d->raw = H.raw[l] ? H.raw[l] + (((size_t)b * H.na + a) * nynx + s0) * no : nullptr;
    d->nynx = nynx; d->nvalid = nvalid; d->nx = nx;
    d->gy0 = gy0; d->gx0 = gx0;
    if (loc) { loc->b = b; loc->seg = cur.ti; loc->row0 = row0; loc->s0 = s0; }
    d->raw_bulk = d->raw != nullptr && ((reinterpret_cast<uintptr_t>(d->raw) & 15) == 0) && (((nvalid * no) & 3) == 0);
    d->vec = ((nynx & 3) == 0) && ((reinterpret_cast<uintptr_t>(H.lv[l]) & 15) == 0);
    d->stride = H.stride[l]; d->aw = H.anchors[l][2 * a]; d->ah = H.anchors[l][2 * a + 1];
}

__device__ __forceinline__ int swz(int c, int s) { return c * kTileS + ((((s >> 2) ^ c) & 7) << 2 | (s & 32) | (s & 3)); }

__device__ __forceinline__ void decode_prefetch(const DecTile& d, float* tile, int no) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tile);
    if (d.vec) {
        const int q = threadIdx.x & 15, c0 = threadIdx.x >> 4;      // 16 chunks per channel, 16 channels per pass
        if (4 * q < d.nvalid) {
            const float* src = d.src + (size_t)c0 * d.nynx + 4 * q;
            const size_t step = (size_t)(kDecThreads / 16) * d.nynx;
            // c advances by 16: (c & 7) and with it the chunk position stay the same, dst moves 16 channel rows
            uint32_t dst = base + 4u * (uint32_t)swz(c0, 4 * q);
            for (int c = c0; c < no; c += kDecThreads / 16, src += step, dst += 16 * kTileS * 4) cp_async_16(dst, src);
        }
    } else {
        const int r = threadIdx.x & (kTileS - 1), c0 = threadIdx.x >> 6;
        if (r < d.nvalid) {
            const float* src = d.src + (size_t)c0 * d.nynx + r;
            const size_t step = (size_t)(kDecThreads / kTileS) * d.nynx;
            for (int c = c0; c < no; c += kDecThreads / kTileS, src += step)
                cp_async_4(base + 4u * (uint32_t)swz(c, r), src);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// NCG > 0: no <= 32*NCG channels, item loop unrolled; NCG == 0: runtime loop over channel groups.
template <int NCG>
__global__ void __launch_bounds__(kDecThreads, 3)
detect_decode_kernel(const HeadDev H, float* __restrict__ pred, int total_tiles, int have_lin) {
    extern __shared__ __align__(16) float tiles_sm[];  // 2 x [no][kTileS] logits, chunk-swizzled (+ [kTileS][no] raw staging)
    __shared__ DecTile s_dt[3];
    const int no = H.no;
    const int tile_floats = kTileS * no;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int t = blockIdx.x;
    if (t >= total_tiles) return;
    __shared__ DecCursor cur;                       // thread 0 only (shared: keeps it out of everyone's registers)
    if (threadIdx.x == 0) {
        cur = cursor_begin(H, t, gridDim.x);
        decode_tile_at(H, pred, cur, &s_dt[0]);
        cursor_next(H, cur);
        if (cur.t < total_tiles) decode_tile_at(H, pred, cur, &s_dt[1]);
        cursor_next(H, cur);
    }
    __syncthreads();
    decode_prefetch(s_dt[0], tiles_sm, no);
    for (int k = 0; t < total_tiles; ++k, t += gridDim.x) {
        const float* tile = tiles_sm + (k & 1) * tile_floats;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (have_lin && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();                      // tile k landed; everyone is done with the other buffer (and `lin`)
        const int slot = k % 3;
        if (t + (int)gridDim.x < total_tiles)
            decode_prefetch(s_dt[slot == 2 ? 0 : slot + 1], tiles_sm + ((k + 1) & 1) * tile_floats, no);
        if (threadIdx.x == 0) {                                            // slot (k+2)%3 was tile k-1's: free
            if (cur.t < total_tiles) decode_tile_at(H, pred, cur, &s_dt[slot == 0 ? 2 : slot - 1]);
            cursor_next(H, cur);
        }
        const DecTile& d = s_dt[slot];
        const int nvalid = d.nvalid;
        float* __restrict__ out = d.out;
        float* __restrict__ raw = d.raw;
        float* const lin = tiles_sm + 2 * tile_floats;       // raw logits of the tile in output order
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
            boxv = decode_elem(tile[swz(cb, br)], cb, (float)((cb & 1) ? gy : gx), d.stride, anc, H.variant);
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
                v[j] = (c < no) ? *reinterpret_cast<const float4*>(tile + swz(c, 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
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
                        item(cgp, q, (c < no) ? *reinterpret_cast<const float4*>(tile + swz(c, 4 * q))
                                              : make_float4(0.f, 0.f, 0.f, 0.f));
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
// confidence filter: one block walks a GROUP of up to 8 consecutive tiles of one plane.
//
//   stage A  objectness of the whole group (<= 512 rows, 2 per thread, one DRAM round trip),
//            ordered list of surviving rows per tile
//   stage B  tiles with <= 8 survivors are "sparse": their rows (<= 64 for the group) are
//            gathered into one staging buffer in a single round trip, evaluated, and their
//            candidates written into the slot range of the group's first sparse tile
//   stage C  the remaining "dense" tiles are loaded whole (coalesced planes / rows), one tile
//            at a time, each into its own slot range
//
// At demo thresholds (0.7 % of rows survive) a group costs ~3 dependent memory round trips
// instead of 3 per tile; at eval thresholds every tile is dense and HBM-bound.
// ---------------------------------------------------------------------------------------
#ifndef VK_DENSE_BPS
#define VK_DENSE_BPS 4
#endif
constexpr int kGroupMax = 8;
constexpr int kItems = kTileS;  // rows evaluated per back-end call (64)


struct FilterSmem {
    float obj[kGroupMax * kTileS];
    short tile_list[kGroupMax][kTileS];  // surviving rows of each tile, ascending
    int wcnt[2 * kWarps];
    int tile_np[kGroupMax];
    int tile_first[kGroupMax + 1];       // first sparse item of each tile
    int it_ai[kItems];                   // accessor index (staging slot or tile row)
    int it_row[kItems];                  // prediction row within the image
    float it_obj[kItems];
    int it_cnt[kItems];
    int it_excl[kItems + 1];
    float it_bv[kItems];
    int it_bj[kItems];
    int part[kDecThreads + 1];           // per (item, class part): count, then exclusive offset
    float part_bv[kDecThreads];          // best-class partial maxima
    int part_bj[kDecThreads];
    int wsum[33];
    int base;
};


// ---- accessors: where the 5+nc values of item `ai` live and how they turn into numbers
struct PlaneGeom {       // fused path: what is needed to decode a box from logits
    int variant, nx, s0; // s0: spatial index of the group's first row inside its plane
    float stride, aw, ah;
    __device__ __forceinline__ float4 box(float l0, float l1, float l2, float l3, int sp) const {
        const int gy = sp / nx, gx = sp - gy * nx;
        return xyxy_from_cxcywh(decode_elem(l0, 0, (float)gx, stride, aw, variant),
                                decode_elem(l1, 1, (float)gy, stride, ah, variant),
                                decode_elem(l2, 2, 0.f, stride, aw, variant),
                                decode_elem(l3, 3, 0.f, stride, ah, variant));
    }
};
struct LogitTile {       // dense tile of logits [no][kFiltPitch]; ai = row inside the tile
    float* t; PlaneGeom g; int tile_s0;
    __device__ __forceinline__ float prob(int ai, int c) const { return sigmoidf_vk(t[(5 + c) * kFiltPitch + ai]); }
    __device__ __forceinline__ void put(int ai, int c, float v) { t[(5 + c) * kFiltPitch + ai] = v; }
    __device__ __forceinline__ float get(int ai, int c) const { return t[(5 + c) * kFiltPitch + ai]; }
    __device__ __forceinline__ float4 box(int ai) const {
        return g.box(t[ai], t[kFiltPitch + ai], t[2 * kFiltPitch + ai], t[3 * kFiltPitch + ai], tile_s0 + ai);
    }
};
struct LogitStage {      // gathered rows of logits [slot][no]; sp[slot] = spatial index in the plane
    float* t; int no; PlaneGeom g; const int* sp;
    __device__ __forceinline__ float prob(int ai, int c) const { return sigmoidf_vk(t[ai * no + 5 + c]); }
    __device__ __forceinline__ void put(int ai, int c, float v) { t[ai * no + 5 + c] = v; }
    __device__ __forceinline__ float get(int ai, int c) const { return t[ai * no + 5 + c]; }
    __device__ __forceinline__ float4 box(int ai) const {
        const float* p = t + ai * no;
        return g.box(p[0], p[1], p[2], p[3], sp[ai]);
    }
};
struct PredRows {        // decoded prediction rows [ai][no] (dense tile or gathered rows alike)
    float* t; int no;
    __device__ __forceinline__ float prob(int ai, int c) const { return t[ai * no + 5 + c]; }
    __device__ __forceinline__ void put(int ai, int c, float v) { t[ai * no + 5 + c] = v; }
    __device__ __forceinline__ float get(int ai, int c) const { return t[ai * no + 5 + c]; }
    __device__ __forceinline__ float4 box(int ai) const {
        const float* p = t + ai * no;
        return xyxy_from_cxcywh(p[0], p[1], p[2], p[3]);
    }
};

// Evaluates S.it_*[0..n_items): class products, per-item candidate counts, then the ordered
// candidate writes starting at slot `base` (a range the caller's tile owns).  Ends with
// S.it_excl[0..n_items] valid and S.base = base.  All threads of the block must call it.
//
// Thread = (class part q, item i): lanes run over items, every thread walks its own consecutive
// class range, so counting needs no ballot and the class order inside a row is the part order.
// One exclusive scan over the 256 (item, part) counts gives every thread its first slot.
template <class Acc>
__device__ __forceinline__ void filter_items(FilterSmem& S, Acc& T, const FilterArgs& A, int b, int n_items,
                                             int base) {
    const int tid = threadIdx.x;
    const int nc = A.nc;
    // Q class parts per item, as many as 256 threads allow (4..8): an eval-mode tile with ~43
    // surviving rows runs 5 parts on 215 threads instead of 4 parts on 172
    const int ni = max(n_items, 32);
    const int Q = min(8, kDecThreads / ni);
    const int qd = tid / ni, i = tid - qd * ni;
    const int cpp = (nc + Q - 1) / Q;
    const int c_lo = min(nc, qd * cpp), c_hi = min(nc, c_lo + cpp);
    const bool act = i < n_items && qd < Q;
    const int ai = act ? S.it_ai[i] : 0;
    const int slot = (qd < Q) ? i * Q + qd : tid;       // item-major, part-minor: canonical order; idle threads
                                                        // own the unused tail slots
    {
        int count = 0;
        float bv = -INFINITY;
        int bj = 0x7fffffff;
        if (act) {
            const float obj = S.it_obj[i];
            if (A.multi_label) {
                if (A.class_mask == nullptr) {
#pragma unroll 4
                    for (int c = c_lo; c < c_hi; ++c) {
                        const float prod = __fmul_rn(T.prob(ai, c), obj);              // image_proc.py:135
                        const bool flag = prod > A.conf;                               // :141
                        T.put(ai, c, flag ? prod : -1.0f);
                        count += flag;
                    }
                } else {
                    for (int c = c_lo; c < c_hi; ++c) {
                        const float prod = __fmul_rn(T.prob(ai, c), obj);
                        const bool flag = (prod > A.conf) && class_allowed(A.class_mask, c);   // :141,151
                        T.put(ai, c, flag ? prod : -1.0f);
                        count += flag;
                    }
                }
            } else {
#pragma unroll 4
                for (int c = c_lo; c < c_hi; ++c) {
                    const float prod = __fmul_rn(T.prob(ai, c), obj);
                    if (prod > bv) { bv = prod; bj = c; }      // first max within the part (:145)
                }
            }
        }
        S.part[slot] = count;
        S.part_bv[slot] = bv;
        S.part_bj[slot] = bj;
    }
    __syncthreads();
    if (!A.multi_label) {
        if (act && qd == 0) {                                   // first max across the parts
            float bv = S.part_bv[slot];
            int bj = S.part_bj[slot];
            for (int q2 = 1; q2 < Q; ++q2)
                if (S.part_bv[slot + q2] > bv) { bv = S.part_bv[slot + q2]; bj = S.part_bj[slot + q2]; }
            const bool sel = (bj != 0x7fffffff) && (bv > A.conf) && class_allowed(A.class_mask, bj);  // :147,151
            S.it_bv[i] = bv;
            S.it_bj[i] = bj;
            S.part[slot] = sel ? 1 : 0;
        }
        __syncthreads();
    }
    {   // exclusive scan of the 256 part counts in slot order
        int total;
        const int v = S.part[tid];
        const int ex = block_excl_scan(v, S.wsum, &total);
        S.part[tid] = ex;
        if (tid == 0) {
            S.part[kDecThreads] = total;
            S.base = base;
            if (total) atomicAdd(A.counts + b, total);         // result unused: fire-and-forget
        }
    }
    __syncthreads();
    if (tid <= kItems) {
        const int t = tid;
        S.it_excl[t] = (t < n_items) ? S.part[t * Q] : S.part[kDecThreads];
    }
    if (tid < kItems) S.it_cnt[tid] = (tid < n_items) ? S.part[(tid + 1) * Q] - S.part[tid * Q] : 0;
    uint64_t* cand = A.cand + (size_t)b * A.cap;
    if (act) {
        const int row = S.it_row[i];
        int pos = base + S.part[slot];
        if (A.multi_label) {
            const int end = base + S.part[slot + 1];
            if (pos < end) {
                uint32_t idx = (uint32_t)(row * nc + c_lo);
                if (end <= A.cap) {            // always, with a buffer sized per include/vk_b200.h
                    uint2* wp = reinterpret_cast<uint2*>(cand) + pos;   // .x = score bits, .y = row*nc + cls
                    for (int c = c_lo; c < c_hi; ++c, ++idx) {
                        const float v = T.get(ai, c);
                        if (v >= 0.0f) *wp++ = make_uint2(__float_as_uint(v), idx);
                    }
                } else {
                    for (int c = c_lo; c < c_hi; ++c, ++idx) {
                        const float v = T.get(ai, c);
                        if (v >= 0.0f) {
                            if (pos < A.cap) cand[pos] = ((uint64_t)idx << 32) | __float_as_uint(v);
                            ++pos;
                        }
                    }
                }
            }
        } else if (qd == 0 && S.part[slot + 1] > S.part[slot] && pos < A.cap) {
            cand[pos] = ((uint64_t)(uint32_t)(row * nc + S.it_bj[i]) << 32) | __float_as_uint(S.it_bv[i]);
        }
        if (qd == 0 && S.part[(i + 1) * Q] > S.part[slot]) A.boxes[(size_t)b * A.rows + row] = T.box(ai);
    }
    __syncthreads();
}

// Stage A for both kernels: S.obj holds the group's objectness (-1 for rows past the end).
// Builds the per-tile ordered survivor lists and counts.  Returns the group's survivor count.
__device__ __forceinline__ int survivors(FilterSmem& S, float conf, int ntiles) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool pa = S.obj[threadIdx.x] > conf;                 // image_proc.py:99
    const bool pb = S.obj[threadIdx.x + kDecThreads] > conf;
    const unsigned ma = __ballot_sync(0xffffffffu, pa), mb = __ballot_sync(0xffffffffu, pb);
    if (lane == 0) { S.wcnt[w] = __popc(ma); S.wcnt[kWarps + w] = __popc(mb); }
    const int total = __syncthreads_count(pa) + __syncthreads_count(pb);
    if (total == 0) return 0;
    const unsigned lt = (1u << lane) - 1u;
    if (pa) S.tile_list[w >> 1][__popc(ma & lt) + ((w & 1) ? S.wcnt[w - 1] : 0)] = (short)(threadIdx.x & 63);
    if (pb) S.tile_list[4 + (w >> 1)][__popc(mb & lt) + ((w & 1) ? S.wcnt[kWarps + w - 1] : 0)] = (short)(threadIdx.x & 63);
    if (threadIdx.x < kGroupMax) {
        const int k = threadIdx.x;
        const int base = (k < 4) ? 2 * k : kWarps + 2 * (k - 4);
        S.tile_np[k] = (k < ntiles) ? S.wcnt[base] + S.wcnt[base + 1] : 0;
    }
    __syncthreads();
    return total;
}

__device__ __forceinline__ bool tile_is_sparse(int np, int nvalid) { return np > 0 && np * 8 <= nvalid; }

// Lays the sparse tiles' survivors out as items 0..n (staging slot = item index).
// Returns n (<= 64).  tile_nvalid(k) = rows of tile k.
template <class NV>
__device__ __forceinline__ int plan_sparse_items(FilterSmem& S, int ntiles, int row0_group, NV tile_nvalid) {
    if (threadIdx.x <= kGroupMax) {
        int run = 0;
        for (int k = 0; k < (int)threadIdx.x; ++k)
            if (k < ntiles && tile_is_sparse(S.tile_np[k], tile_nvalid(k))) run += S.tile_np[k];
        S.tile_first[threadIdx.x] = run;
    }
    __syncthreads();
    const int run = S.tile_first[kGroupMax];
    if (threadIdx.x < run) {
        int k = 0;
#pragma unroll
        for (int q = 1; q < kGroupMax; ++q)
            if ((int)threadIdx.x >= S.tile_first[q]) k = q;
        const int r = S.tile_list[k][threadIdx.x - S.tile_first[k]];
        S.it_ai[threadIdx.x] = threadIdx.x;
        S.it_row[threadIdx.x] = row0_group + k * kTileS + r;
        S.it_obj[threadIdx.x] = S.obj[k * kTileS + r];
    }
    __syncthreads();
    return run;
}

// All sparse survivors of a group (<= 64 rows) fit the slot range of its first sparse tile.
template <class NV>
__device__ __forceinline__ int first_sparse_tile(const FilterSmem& S, int ntiles, NV tile_nvalid) {
    for (int k = 0; k < ntiles; ++k)
        if (tile_is_sparse(S.tile_np[k], tile_nvalid(k))) return k;
    return 0;
}

// Items of one dense tile: its survivors, accessor index = row inside the tile.
__device__ __forceinline__ int plan_dense_items(FilterSmem& S, int k, int row0_group) {
    const int np = S.tile_np[k];
    if (threadIdx.x < np) {
        const int r = S.tile_list[k][threadIdx.x];
        S.it_ai[threadIdx.x] = r;
        S.it_row[threadIdx.x] = row0_group + k * kTileS + r;
        S.it_obj[threadIdx.x] = S.obj[k * kTileS + r];
    }
    return np;
}

__device__ __forceinline__ void write_empty_segments(const FilterArgs& A, int b, int seg0, int ntiles) {
    if (threadIdx.x < ntiles) {
        A.seg_base[(size_t)b * A.segs + seg0 + threadIdx.x] = 0;
        A.seg_count[(size_t)b * A.segs + seg0 + threadIdx.x] = 0;
    }
}

// Segment table entries of the group's sparse and empty tiles after filter_items().
template <class NV>
__device__ __forceinline__ void write_sparse_segments(FilterSmem& S, const FilterArgs& A, int b, int seg0,
                                                      int ntiles, int n_items, NV tile_nvalid) {
    if (threadIdx.x < ntiles) {
        const int k = threadIdx.x;
        const int np = S.tile_np[k];
        const bool sparse = tile_is_sparse(np, tile_nvalid(k));
        if (sparse || np == 0) {
            int sb = 0, sc = 0;
            if (sparse && n_items > 0) {
                const int f = S.tile_first[k];
                const int e0 = S.it_excl[f];
                const int e1 = (f + np >= n_items) ? S.it_excl[kItems] : S.it_excl[f + np];
                sb = S.base + e0;
                sc = e1 - e0;
            }
            A.seg_base[(size_t)b * A.segs + seg0 + k] = sb;
            A.seg_count[(size_t)b * A.segs + seg0 + k] = sc;
        }
    }
}

__device__ __forceinline__ void write_dense_segment(FilterSmem& S, const FilterArgs& A, int b, int seg) {
    if (threadIdx.x == 0) {
        A.seg_base[(size_t)b * A.segs + seg] = S.base;
        A.seg_count[(size_t)b * A.segs + seg] = S.it_excl[kItems];
    }
}

__global__ void __launch_bounds__(kDecThreads, 6)
decode_filter_kernel(const HeadDev H, const FilterArgs A) {
    extern __shared__ __align__(16) float buf[];  // dense tile [no][kFiltPitch] or staging [64][no]
    __shared__ FilterSmem S;
    __shared__ int s_sp[kItems];
    const int b = blockIdx.y;
    const int G = A.group;
    // locate the group: planes are (level, anchor); groups never straddle a plane
    int l = 0;
#pragma unroll
    for (int i = 1; i < VK_MAX_LEVELS; ++i)
        if (i < H.nl && (int)blockIdx.x >= H.group_start[i]) l = i;
    const int rel = blockIdx.x - H.group_start[l];
    const int gpa = ceil_div(H.tpa[l], G);
    const int a = rel / gpa;
    const int tile0 = (rel - a * gpa) * G;                 // first tile of the group inside the plane
    const int ntiles = min(G, H.tpa[l] - tile0);
    const int s0 = tile0 * kTileS;
    const int no = H.no, nynx = H.nynx[l];
    const int nrows = min(ntiles * kTileS, nynx - s0);
    const int row0 = H.row_base[l] + a * nynx + s0;
    const int seg0 = H.tile_start[l] + a * H.tpa[l] + tile0;
    const float* __restrict__ in = H.lv[l] + ((size_t)(b * H.na + a) * no) * nynx + s0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    auto tile_nvalid = [&](int k) { return min(kTileS, nrows - k * kTileS); };

    // stage A: objectness plane of the group (coalesced)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = threadIdx.x + h * kDecThreads;
        S.obj[r] = (r < nrows) ? sigmoidf_vk(ld_stream_f32(in + (size_t)4 * nynx + r)) : -1.0f;
    }
    __syncthreads();
    if (survivors(S, A.conf, ntiles) == 0) {
        write_empty_segments(A, b, seg0, ntiles);
        return;
    }
    const PlaneGeom geom{H.variant, H.nx[l], s0, H.stride[l], H.anchors[l][2 * a], H.anchors[l][2 * a + 1]};

    // stage B: sparse tiles, one gather for the whole group
    const int n_sparse = plan_sparse_items(S, ntiles, row0, tile_nvalid);
    if (n_sparse > 0) {
        for (int i = w; i < n_sparse; i += kWarps) {
            const int sp = S.it_row[i] - row0;              // row inside the group
            if (lane == 0) s_sp[i] = s0 + sp;
            for (int c = lane; c < no; c += 32) buf[i * no + c] = __ldg(in + (size_t)c * nynx + sp);
        }
        __syncthreads();
        LogitStage T{buf, no, geom, s_sp};
        filter_items(S, T, A, b, n_sparse, (seg0 + first_sparse_tile(S, ntiles, tile_nvalid)) * A.tile_cap);
    }
    write_sparse_segments(S, A, b, seg0, ntiles, n_sparse, tile_nvalid);
    __syncthreads();

    // stage C: dense tiles, coalesced plane loads
    const bool vec = ((nynx & 3) == 0) && ((reinterpret_cast<uintptr_t>(H.lv[l]) & 15) == 0);
    for (int k = 0; k < ntiles; ++k) {
        const int np = S.tile_np[k], nv = tile_nvalid(k);
        if (np == 0 || tile_is_sparse(np, nv)) continue;
        load_tile<kFiltPitch>(buf, in + k * kTileS, no, nynx, nv, vec);
        const int n_items = plan_dense_items(S, k, row0);
        __syncthreads();
        LogitTile T{buf, geom, s0 + k * kTileS};
        filter_items(S, T, A, b, n_items, (seg0 + k) * A.tile_cap);
        write_dense_segment(S, A, b, seg0 + k);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Dense variant of the fused filter (eval thresholds: most rows survive, every tile is read
// whole).  Same persistent two-deep pipeline and tile layout as detect_decode_kernel: 16-byte
// async copies keep a full tile of reads in flight per block while the previous tile is
// evaluated.  Thread = (class part qd = tid / 64, row = tid % 64): lanes run over rows
// (conflict-free scalar reads of the chunk-swizzled tile), every thread walks its own CPP
// consecutive classes and keeps the products in registers, so there is no second pass over
// shared memory.  One scan over the 256 (row, part) counts in canonical order gives every
// thread its first slot inside the range the tile owns.
// ML = multi_label.  Results are bit-identical to decode_filter_kernel (same sigmoid, same
// product, same order); the host picks the kernel from the threshold only (vk_decode_filter).
// ---------------------------------------------------------------------------------------
constexpr int kParts = kDecThreads / kTileS;   // 4 class parts per row

// p[I] = sigmoid(logit of the thread's I-th class) * obj, the logit read with an immediate offset
// from one of eight base addresses (compile-time recursion: the offset must be a constant).
template <int I, int N>
struct ClassProducts {
    static __device__ __forceinline__ void run(float* p, const uint32_t* tq, float obj) {
        float x;
        asm("ld.shared.f32 %0, [%1+%2];" : "=f"(x) : "r"(tq[I & 7]), "n"(I * kTileS * 4));
        p[I] = __fmul_rn(sigmoidf_vk(x), obj);                                                    // image_proc.py:135
        ClassProducts<I + 1, N>::run(p, tq, obj);
    }
};
template <int N>
struct ClassProducts<N, N> {
    static __device__ __forceinline__ void run(float*, const uint32_t*, float) {}
};

template <int CPP, bool ML>
__global__ void __launch_bounds__(kDecThreads, VK_DENSE_BPS)
decode_filter_dense_kernel(const HeadDev H, const FilterArgs A, int total_tiles) {
    extern __shared__ __align__(16) float tiles_sm[];  // 2 x [no][kTileS] logits, chunk-swizzled
    __shared__ DecTile s_dt[3];
    __shared__ TilePos s_pos[3];
    __shared__ int s_cnt[kDecThreads];      // counts, slot = row * kParts + part (canonical order)
    __shared__ int s_off[kDecThreads + 1];  // exclusive offsets inside each warp's 32 slots
    __shared__ int s_wsum[kWarps];
    __shared__ float s_bv[kDecThreads];     // best-class partials
    __shared__ int s_bj[kDecThreads];
    const int no = H.no, nc = A.nc;
    const int tile_floats = kTileS * no;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int row = threadIdx.x & (kTileS - 1), qd = threadIdx.x >> 6;
    const int cpp = (nc + kParts - 1) / kParts;            // <= CPP
    const int c_lo = min(nc, qd * cpp), c_hi = min(nc, c_lo + cpp);
    const int slot = row * kParts + qd;
    // classes this thread may emit (class filter, :151): bit i <-> class c_lo + i
    uint32_t allowed = 0;
    for (int i = 0; i < c_hi - c_lo; ++i)
        if (class_allowed(A.class_mask, c_lo + i)) allowed |= 1u << i;
    int t = blockIdx.x;
    if (t >= total_tiles) return;
    __shared__ DecCursor cur;                       // thread 0 only (shared: keeps it out of everyone's registers)
    if (threadIdx.x == 0) {
        cur = cursor_begin(H, t, gridDim.x);
        decode_tile_at(H, nullptr, cur, &s_dt[0], &s_pos[0]);
        cursor_next(H, cur);
        if (cur.t < total_tiles) decode_tile_at(H, nullptr, cur, &s_dt[1], &s_pos[1]);
        cursor_next(H, cur);
    }
    __syncthreads();
    decode_prefetch(s_dt[0], tiles_sm, no);
    for (int k = 0; t < total_tiles; ++k, t += gridDim.x) {
        const float* tile = tiles_sm + (k & 1) * tile_floats;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                      // tile k landed; everyone is done with the other buffer and the scratch arrays
        const int slot3 = k % 3;
        if (t + (int)gridDim.x < total_tiles)
            decode_prefetch(s_dt[slot3 == 2 ? 0 : slot3 + 1], tiles_sm + ((k + 1) & 1) * tile_floats, no);
        if (threadIdx.x == 0) {
            if (cur.t < total_tiles) decode_tile_at(H, nullptr, cur, &s_dt[slot3 == 0 ? 2 : slot3 - 1], &s_pos[slot3 == 0 ? 2 : slot3 - 1]);
            cursor_next(H, cur);
        }
        const DecTile& d = s_dt[slot3];
        const TilePos& tp = s_pos[slot3];

        // ---- products of this thread's classes (registers), flags as a bit mask.
        // Element i of the thread is channel ch0 + i of its row; its 16-byte chunk sits at position
        // (row/4 ^ channel) & 7, which repeats with period 8 in i: eight base pointers, immediate offsets.
        const float* trow = tile + ((row & 32) | (row & 3));
        const int rq = row >> 2;
        const float o = sigmoidf_vk(trow[4 * kTileS + (((rq ^ 4) & 7) << 2)]);
        const float obj = (row < d.nvalid && o > A.conf) ? o : 0.0f;  // image_proc.py:99 (dead rows: products 0)
        const int ch0 = 5 + c_lo;
        uint32_t tq[8];
        {
            const uint32_t trow_s = (uint32_t)__cvta_generic_to_shared(trow + ch0 * kTileS);
#pragma unroll
            for (int j = 0; j < 8; ++j) tq[j] = trow_s + ((((uint32_t)(rq ^ (ch0 + j))) & 7u) << 4);
        }
        float p[CPP];
        ClassProducts<0, CPP>::run(p, tq, obj);   // reads past the thread's last class stay inside the (padded) buffer
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
            if (qd == 0) {                                          // first max across the parts
#pragma unroll
                for (int q2 = 1; q2 < kParts; ++q2)
                    if (s_bv[slot + q2] > bv) { bv = s_bv[slot + q2]; bj = s_bj[slot + q2]; }
                count = (bj != 0x7fffffff && bv > A.conf && class_allowed(A.class_mask, bj)) ? 1 : 0;   // :147,151
            }
        }
        s_cnt[slot] = count;
        __syncthreads();
        {   // scan in slot order: thread tid owns slot tid
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
        const int tile_base = tp.seg * A.tile_cap;
        if (count) {
            uint2* const wp = reinterpret_cast<uint2*>(A.cand + (size_t)tp.b * A.cap) + tile_base + wbase + s_off[slot];
            const uint32_t idx0 = (uint32_t)((tp.row0 + row) * nc + c_lo);
            if (ML) {
                uint32_t pos = 0;
#pragma unroll
                for (int i = 0; i < CPP; ++i) {
                    const bool f = (flags & (1u << i)) != 0;
                    if (f) wp[pos] = make_uint2(__float_as_uint(p[i]), idx0 + (uint32_t)i);
                    pos += f;
                }
            } else {
                *wp = make_uint2(__float_as_uint(bv), (uint32_t)((tp.row0 + row) * nc + bj));
            }
        }
        // ---- box of the row if any of its parts produced a candidate
        if (qd == 0) {
            const int n = s_cnt[slot] + s_cnt[slot + 1] + s_cnt[slot + 2] + s_cnt[slot + 3];
            if (n > 0) {
                const PlaneGeom geom{H.variant, d.nx, tp.s0, d.stride, d.aw, d.ah};
                A.boxes[(size_t)tp.b * A.rows + tp.row0 + row] =
                    geom.box(tile[swz(0, row)], tile[swz(1, row)], tile[swz(2, row)], tile[swz(3, row)], tp.s0 + row);
            }
        }
        if (threadIdx.x == 0) {
            A.seg_base[(size_t)tp.b * A.segs + tp.seg] = total ? tile_base : 0;
            A.seg_count[(size_t)tp.b * A.segs + tp.seg] = total;
            if (total) atomicAdd(A.counts + tp.b, total);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Dense variant of filter_pred (the `nms(prediction)` drop-in at eval thresholds): the same
// persistent pipeline and (row, class part) mapping as decode_filter_dense_kernel, reading an
// existing (B, rows, no) prediction tensor.  A tile is 64 consecutive rows = one contiguous run of
// 64*no floats, copied as it is ([row][no], odd pitch: lanes over rows are conflict-free); the
// values are probabilities already, boxes are cxcywh.
// ---------------------------------------------------------------------------------------
template <int I, int N>
struct PredProducts {
    static __device__ __forceinline__ void run(float* p, uint32_t base, float obj) {
        float x;
        asm("ld.shared.f32 %0, [%1+%2];" : "=f"(x) : "r"(base), "n"(I * 4));
        p[I] = __fmul_rn(x, obj);                                                                 // image_proc.py:135
        PredProducts<I + 1, N>::run(p, base, obj);
    }
};
template <int N>
struct PredProducts<N, N> {
    static __device__ __forceinline__ void run(float*, uint32_t, float) {}
};

__device__ __forceinline__ void pred_prefetch(const float* __restrict__ src, int nfloats, float* tile) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tile);
    if (((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((nfloats & 3) == 0)) {
        for (int e = threadIdx.x; 4 * e < nfloats; e += kDecThreads) cp_async_16(base + 16u * e, src + 4 * e);
    } else {
        for (int e = threadIdx.x; e < nfloats; e += kDecThreads) cp_async_4(base + 4u * e, src + e);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int CPP, bool ML>
__global__ void __launch_bounds__(kDecThreads, VK_DENSE_BPS)
filter_pred_dense_kernel(const float* __restrict__ pred, int no, const FilterArgs A, int total_tiles) {
    extern __shared__ __align__(16) float tiles_sm[];  // 2 x [kTileS][no] prediction rows
    __shared__ int s_cnt[kDecThreads];
    __shared__ int s_off[kDecThreads + 1];
    __shared__ int s_wsum[kWarps];
    __shared__ float s_bv[kDecThreads];
    __shared__ int s_bj[kDecThreads];
    const int nc = A.nc;
    const int tile_floats = kTileS * no;
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
        const float* src = tile_src(t, b, seg, nvalid);
        pred_prefetch(src, nvalid * no, tiles_sm);
    }
    for (int k = 0; t < total_tiles; ++k, t += gridDim.x) {
        const float* tile = tiles_sm + (k & 1) * tile_floats;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                      // tile k landed; everyone is done with the other buffer and the scratch arrays
        if (t + (int)gridDim.x < total_tiles) {
            int b2, seg2, nv2;
            const float* src = tile_src(t + gridDim.x, b2, seg2, nv2);
            pred_prefetch(src, nv2 * no, tiles_sm + ((k + 1) & 1) * tile_floats);
        }
        tile_src(t, b, seg, nvalid);
        const float* prow = tile + row * no;
        const float o = prow[4];
        const float obj = (row < nvalid && o > A.conf) ? o : 0.0f;    // image_proc.py:99 (dead rows: products 0)
        float p[CPP];
        PredProducts<0, CPP>::run(p, (uint32_t)__cvta_generic_to_shared(prow + 5 + c_lo), obj);
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
            if (n > 0) A.boxes[(size_t)b * A.rows + grow] = xyxy_from_cxcywh(prow[0], prow[1], prow[2], prow[3]);
        }
        if (threadIdx.x == 0) {
            A.seg_base[(size_t)b * A.segs + seg] = total ? tile_base : 0;
            A.seg_count[(size_t)b * A.segs + seg] = total;
            if (total) atomicAdd(A.counts + b, total);
        }
    }
}

__global__ void __launch_bounds__(kDecThreads, 5)
filter_pred_kernel(const float* __restrict__ pred, int no, const FilterArgs A) {
    extern __shared__ float buf[];  // [64][no]: a dense tile or the gathered rows
    __shared__ FilterSmem S;
    const int b = blockIdx.y;
    const int G = A.group;
    const int seg0 = blockIdx.x * G;
    const int ntiles = min(G, A.segs - seg0);
    const int row0 = seg0 * kTileS;
    const int nrows = min(ntiles * kTileS, A.rows - row0);
    const float* __restrict__ in = pred + ((size_t)b * A.rows + row0) * no;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    auto tile_nvalid = [&](int k) { return min(kTileS, nrows - k * kTileS); };

#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = threadIdx.x + h * kDecThreads;
        S.obj[r] = (r < nrows) ? __ldg(in + (size_t)r * no + 4) : -1.0f;
    }
    __syncthreads();
    if (survivors(S, A.conf, ntiles) == 0) {
        write_empty_segments(A, b, seg0, ntiles);
        return;
    }
    const int n_sparse = plan_sparse_items(S, ntiles, row0, tile_nvalid);
    if (n_sparse > 0) {
        for (int i = w; i < n_sparse; i += kWarps) {
            const float* src = in + (size_t)(S.it_row[i] - row0) * no;
            for (int c = lane; c < no; c += 32) buf[i * no + c] = __ldg(src + c);
        }
        __syncthreads();
        PredRows T{buf, no};
        filter_items(S, T, A, b, n_sparse, (seg0 + first_sparse_tile(S, ntiles, tile_nvalid)) * A.tile_cap);
    }
    write_sparse_segments(S, A, b, seg0, ntiles, n_sparse, tile_nvalid);
    __syncthreads();

    for (int k = 0; k < ntiles; ++k) {
        const int np = S.tile_np[k], nv = tile_nvalid(k);
        if (np == 0 || tile_is_sparse(np, nv)) continue;
        const float* __restrict__ tin = in + (size_t)k * kTileS * no;   // nv*no contiguous floats
        const int n = nv * no;
        if (((reinterpret_cast<uintptr_t>(tin) & 15) == 0) && ((n & 3) == 0)) {
            for (int e = threadIdx.x; e < (n >> 2); e += kDecThreads)
                reinterpret_cast<float4*>(buf)[e] = ld_stream_f4(tin + 4 * (size_t)e);
        } else {
            for (int e = threadIdx.x; e < n; e += kDecThreads) buf[e] = ld_stream_f32(tin + e);
        }
        const int n_items = plan_dense_items(S, k, row0);
        __syncthreads();
        PredRows T{buf, no};
        filter_items(S, T, A, b, n_items, (seg0 + k) * A.tile_cap);
        write_dense_segment(S, A, b, seg0 + k);
        __syncthreads();
    }
}

// Tiles per block: enough blocks for ~2 full waves of 8 resident blocks per SM, at most 8.
static int choose_group(int batch, int tiles) {
    const long blocks = (long)batch * tiles;
    long g = blocks / (2L * kNumSMs * 8);
    if (g < 1) g = 1;
    if (g > kGroupMax) g = kGroupMax;
    return (int)g;
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

extern "C" int vk_detect_decode(const VkHeadCfg* cfg, const float* const* levels, int batch,
                                float* pred, float* const* raw, vk_stream_t stream) {
    HeadDev H;
    if (int rc = make_head(cfg, &H, "vk_detect_decode")) return rc;
    if (batch == 0) return VK_OK;
    if (!levels || !pred || batch < 0) return fail_arg("vk_detect_decode: null/negative argument");
    if (batch > 65535) return fail_code(VK_E_LIMIT, "vk_detect_decode: batch %d > 65535", batch);
    for (int l = 0; l < H.nl; ++l) {
        if (!levels[l]) return fail_arg("vk_detect_decode: level %d is NULL", l);
        H.lv[l] = levels[l];
        H.raw[l] = raw ? raw[l] : nullptr;
    }
    const int have_lin = raw != nullptr;
    const size_t smem = (size_t)(have_lin ? 3 : 2) * kTileS * H.no * sizeof(float);
    if (smem > 200 * 1024) return fail_code(VK_E_LIMIT, "vk_detect_decode: nc=%d needs %zu B of shared memory", H.nc, smem);
    if ((long)H.tiles * batch > 0x7fffffffL) return fail_code(VK_E_LIMIT, "vk_detect_decode: %d x %d tiles", H.tiles, batch);
    const int total_tiles = H.tiles * batch;
#define VK_DEC_LAUNCH(NCG)                                                                                  \
    do {                                                                                                     \
        cudaFuncSetAttribute(detect_decode_kernel<NCG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        int per_sm = 0;                                                                                      \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, detect_decode_kernel<NCG>, kDecThreads, smem); \
        if (per_sm > 3) per_sm = 3;               /* more concurrent tile streams cost DRAM locality */      \
        if (per_sm < 1) per_sm = 1;                                                                          \
        const int grid = total_tiles < per_sm * kNumSMs ? total_tiles : per_sm * kNumSMs;                    \
        detect_decode_kernel<NCG><<<grid, kDecThreads, smem, as_stream(stream)>>>(H, pred, total_tiles, have_lin); \
    } while (0)
    // persistent: every block is resident
    if (H.no <= 32) VK_DEC_LAUNCH(1);
    else if (H.no <= 64) VK_DEC_LAUNCH(2);
    else if (H.no <= 96) VK_DEC_LAUNCH(3);
    else if (H.no <= 128) VK_DEC_LAUNCH(4);
    else VK_DEC_LAUNCH(0);
#undef VK_DEC_LAUNCH
    count_launch();
    return check_launch("detect_decode_kernel");
}

extern "C" int vk_decode_filter(const VkHeadCfg* cfg, const float* const* levels, int batch,
                                float conf_thres, int multi_label, const uint32_t* class_mask,
                                const VkCandBuf* out, vk_stream_t stream_) {
    HeadDev H;
    if (int rc = make_head(cfg, &H, "vk_decode_filter")) return rc;
    if (batch == 0) return VK_OK;
    if (!levels || batch < 0) return fail_arg("vk_decode_filter: null/negative argument");
    if (!(conf_thres >= 0.f && conf_thres <= 1.f)) return fail_arg("vk_decode_filter: conf_thres %g outside [0,1]", conf_thres);
    if (batch > 65535) return fail_code(VK_E_LIMIT, "vk_decode_filter: batch %d > 65535", batch);
    if (int rc = check_cand(out, H.rows, H.tiles, H.nc, multi_label, "vk_decode_filter")) return rc;
    for (int l = 0; l < H.nl; ++l) {
        if (!levels[l]) return fail_arg("vk_decode_filter: level %d is NULL", l);
        H.lv[l] = levels[l];
    }
    cudaStream_t stream = as_stream(stream_);
    cudaError_t e = cudaMemsetAsync(out->counts, 0, (size_t)batch * sizeof(int32_t), stream);
    if (e != cudaSuccess) return fail_code((int)e, "vk_decode_filter: memset: %s", cudaGetErrorString(e));
    const size_t smem = (size_t)H.no * kFiltPitch * sizeof(float);
    if (smem > 200 * 1024) return fail_code(VK_E_LIMIT, "vk_decode_filter: nc=%d needs %zu B of shared memory", H.nc, smem);
    cudaFuncSetAttribute(decode_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FilterArgs A = make_filter_args(out, conf_thres, multi_label, class_mask);
    A.group = choose_group(batch, H.tiles);
    int groups = 0;
    for (int l = 0; l < H.nl; ++l) {
        H.group_start[l] = groups;
        groups += H.na * ceil_div(H.tpa[l], A.group);
    }
    for (int l = H.nl; l <= VK_MAX_LEVELS; ++l) H.group_start[l] = groups;
    // Eval thresholds (most rows survive) -> the dense, persistent kernel; otherwise the group
    // kernel that gathers only surviving rows.  Both give identical bits; vk_set_filter_kernel()
    // overrides the choice (tests run every case through both).
    const int mode = filter_mode();
    const bool dense = mode == VK_FILTER_DENSE || (mode == VK_FILTER_AUTO && conf_thres < 0.05f);
    if (dense && H.nc <= 128 && (long)H.tiles * batch <= 0x7fffffffL) {
        const int cpp_max = H.nc <= 32 ? 8 : H.nc <= 80 ? 20 : 32;
        // + one part's worth of rows: the unrolled class loop may read past channel no-1
        const size_t dsmem = (2 * (size_t)kTileS * H.no + (size_t)kTileS * cpp_max) * sizeof(float);
        const int total_tiles = H.tiles * batch;
#define VK_DF_LAUNCH(CPP, ML)                                                                                      \
        do {                                                                                                        \
            cudaFuncSetAttribute(decode_filter_dense_kernel<CPP, ML>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem); \
            int per_sm = 0;                                                                                         \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_filter_dense_kernel<CPP, ML>, kDecThreads, dsmem); \
            if (per_sm > VK_DENSE_BPS) per_sm = VK_DENSE_BPS;                                                       \
            if (per_sm < 1) per_sm = 1;                                                                             \
            const int grid = total_tiles < per_sm * kNumSMs ? total_tiles : per_sm * kNumSMs;                       \
            decode_filter_dense_kernel<CPP, ML><<<grid, kDecThreads, dsmem, stream>>>(H, A, total_tiles);           \
        } while (0)
#define VK_DF_CPP(CPP) do { if (A.multi_label) VK_DF_LAUNCH(CPP, true); else VK_DF_LAUNCH(CPP, false); } while (0)
        if (H.nc <= 32) VK_DF_CPP(8);                 // classes per part = ceil(nc / 4)
        else if (H.nc <= 80) VK_DF_CPP(20);
        else VK_DF_CPP(32);
#undef VK_DF_CPP
#undef VK_DF_LAUNCH
        count_launch();
        return check_launch("decode_filter_dense_kernel");
    }
    decode_filter_kernel<<<dim3(groups, batch), kDecThreads, smem, stream>>>(H, A);
    count_launch();
    return check_launch("decode_filter_kernel");
}

extern "C" int vk_filter_pred(const float* pred, int batch, int rows, int nc, float conf_thres,
                              int multi_label, const uint32_t* class_mask, const VkCandBuf* out,
                              vk_stream_t stream_) {
    if (batch == 0) return VK_OK;
    if (!pred || batch < 0 || rows <= 0 || nc < 1) return fail_arg("vk_filter_pred: null/negative argument");
    if (!(conf_thres >= 0.f && conf_thres <= 1.f)) return fail_arg("vk_filter_pred: conf_thres %g outside [0,1]", conf_thres);
    if (batch > 65535) return fail_code(VK_E_LIMIT, "vk_filter_pred: batch %d > 65535", batch);
    const int segs = ceil_div(rows, kTileS);
    if (segs > VK_MAX_SEGMENTS) return fail_code(VK_E_LIMIT, "vk_filter_pred: %d rows > %d", rows, VK_MAX_SEGMENTS * kTileS);
    if (int rc = check_cand(out, rows, segs, nc, multi_label, "vk_filter_pred")) return rc;
    cudaStream_t stream = as_stream(stream_);
    cudaError_t e = cudaMemsetAsync(out->counts, 0, (size_t)batch * sizeof(int32_t), stream);
    if (e != cudaSuccess) return fail_code((int)e, "vk_filter_pred: memset: %s", cudaGetErrorString(e));
    const int no = nc + 5;
    const size_t smem = (size_t)no * kTileS * sizeof(float);
    if (smem > 200 * 1024) return fail_code(VK_E_LIMIT, "vk_filter_pred: nc=%d needs %zu B of shared memory", nc, smem);
    cudaFuncSetAttribute(filter_pred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FilterArgs A = make_filter_args(out, conf_thres, multi_label, class_mask);
    const int mode = filter_mode();
    const bool dense = mode == VK_FILTER_DENSE || (mode == VK_FILTER_AUTO && conf_thres < 0.05f);
    if (dense && nc <= 128 && (long)segs * batch <= 0x7fffffffL) {
        const int cpp_max = nc <= 32 ? 8 : nc <= 80 ? 20 : 32;
        const size_t dsmem = (2 * (size_t)kTileS * no + (size_t)cpp_max + 8) * sizeof(float);   // + slack: unrolled class loop
        const int total_tiles = segs * batch;
#define VK_FP_LAUNCH(CPP, ML)                                                                                      \
        do {                                                                                                        \
            cudaFuncSetAttribute(filter_pred_dense_kernel<CPP, ML>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem); \
            int per_sm = 0;                                                                                         \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, filter_pred_dense_kernel<CPP, ML>, kDecThreads, dsmem); \
            if (per_sm > VK_DENSE_BPS) per_sm = VK_DENSE_BPS;                                                       \
            if (per_sm < 1) per_sm = 1;                                                                             \
            const int grid = total_tiles < per_sm * kNumSMs ? total_tiles : per_sm * kNumSMs;                       \
            filter_pred_dense_kernel<CPP, ML><<<grid, kDecThreads, dsmem, stream>>>(pred, no, A, total_tiles);      \
        } while (0)
#define VK_FP_CPP(CPP) do { if (A.multi_label) VK_FP_LAUNCH(CPP, true); else VK_FP_LAUNCH(CPP, false); } while (0)
        if (nc <= 32) VK_FP_CPP(8);
        else if (nc <= 80) VK_FP_CPP(20);
        else VK_FP_CPP(32);
#undef VK_FP_CPP
#undef VK_FP_LAUNCH
        count_launch();
        return check_launch("filter_pred_dense_kernel");
    }
    A.group = choose_group(batch, segs);
    filter_pred_kernel<<<dim3(ceil_div(segs, A.group), batch), kDecThreads, smem, stream>>>(pred, no, A);
    count_launch();
    return check_launch("filter_pred_kernel");
}
