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
// order, into a slot range claimed with one atomicAdd, and records (base, count) in a
// segment table indexed by tile; canonical order = segment order x in-segment order.  The
// consumer (nms.cu) walks the table, so results do not depend on which tile won the atomic.
#include "vk_common.cuh"

namespace vk {

constexpr int kTileS = 64;        // spatial positions (= prediction rows) per tile
constexpr int kTilePitch = kTileS + 1;
constexpr int kDecThreads = 256;
constexpr int kWarps = kDecThreads / 32;

struct HeadDev {
    int variant, nl, na, nc, no, rows, tiles;
    int ny[VK_MAX_LEVELS], nx[VK_MAX_LEVELS], nynx[VK_MAX_LEVELS];
    int row_base[VK_MAX_LEVELS], tile_start[VK_MAX_LEVELS + 1], tpa[VK_MAX_LEVELS];
    float stride[VK_MAX_LEVELS];
    float anchors[VK_MAX_LEVELS][2 * VK_MAX_ANCHORS];
    const float* lv[VK_MAX_LEVELS];
    float* raw[VK_MAX_LEVELS];
};

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
__global__ void __launch_bounds__(kDecThreads)
detect_decode_kernel(const HeadDev H, float* __restrict__ pred) {
    extern __shared__ float tile[];  // [no][kTilePitch] logits
    const int b = blockIdx.y;
    const TileLoc q = locate_tile(H, blockIdx.x);
    const int no = H.no, nynx = H.nynx[q.l];
    const float* __restrict__ in = H.lv[q.l] + ((size_t)(b * H.na + q.a) * no) * nynx + q.s0;

    const bool vec = ((nynx & 3) == 0) && ((reinterpret_cast<uintptr_t>(H.lv[q.l]) & 15) == 0);
    if (vec) {
        for (int e = threadIdx.x; e < no * (kTileS / 4); e += kDecThreads) {
            const int c = e >> 4, s = (e & 15) << 2;
            if (s < q.nvalid) {
                const float4 v = ld_stream_f4(in + (size_t)c * nynx + s);
                float* d = tile + c * kTilePitch + s;
                d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            }
        }
    } else {
        for (int e = threadIdx.x; e < no * kTileS; e += kDecThreads) {
            const int c = e >> 6, s = e & 63;
            if (s < q.nvalid) tile[c * kTilePitch + s] = ld_stream_f32(in + (size_t)c * nynx + s);
        }
    }
    __syncthreads();

    const int nx = H.nx[q.l];
    const float stride = H.stride[q.l];
    const float aw = H.anchors[q.l][2 * q.a], ah = H.anchors[q.l][2 * q.a + 1];
    const int nout = q.nvalid * no;
    float* __restrict__ out = pred + ((size_t)b * H.rows + q.row0) * no;
    float* __restrict__ raw = H.raw[q.l]
                                  ? H.raw[q.l] + (((size_t)b * H.na + q.a) * nynx + q.s0) * no
                                  : nullptr;
    for (int o = threadIdx.x; o < nout; o += kDecThreads) {
        const int s = o / no, c = o - s * no;
        const float logit = tile[c * kTilePitch + s];
        float g = 0.f, anc = 0.f;
        if (c < 4) {
            const int sp = q.s0 + s;
            const int gy = sp / nx, gx = sp - gy * nx;
            g = (float)((c & 1) ? gy : gx);
            anc = (c & 1) ? ah : aw;
        }
        st_stream_f32(out + o, decode_elem(logit, c, g, stride, anc, H.variant));
        if (raw) st_stream_f32(raw + o, logit);
    }
}

// ---------------------------------------------------------------------------------------
// confidence filter back-end shared by the fused and the drop-in kernels
// ---------------------------------------------------------------------------------------
struct FilterArgs {
    float conf;
    int multi_label;
    const uint32_t* class_mask;  // dev or null
    uint64_t* cand;
    float4* boxes;
    int32_t* counts;
    int32_t* seg_base;
    int32_t* seg_count;
    int cap, rows, segs, nc;
};

struct FilterSmem {
    float obj[kTileS];
    int cnt[kTileS];
    int excl[kTileS];
    int list[kTileS];   // passing rows, ascending
    float best_v[kTileS];
    int best_j[kTileS];
    int npass, base;
};

__device__ __forceinline__ bool class_allowed(const uint32_t* m, int c) {
    return m == nullptr || ((__ldg(m + (c >> 5)) >> (c & 31)) & 1u);
}

// Ordered list of the rows whose flag is set (warps 0 and 1 cover kTileS = 64 rows).
__device__ __forceinline__ void build_pass_list(FilterSmem& S, bool pass) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __shared__ int first_half;
    if (w < 2) {
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (w == 0 && lane == 0) first_half = __popc(m);
        __syncwarp();
        // second warp needs the first warp's count: publish through shared memory below
        S.cnt[threadIdx.x] = 0;
        if (pass) S.excl[threadIdx.x] = __popc(m & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (w < 2 && pass) S.list[S.excl[threadIdx.x] + (w ? first_half : 0)] = threadIdx.x;
    __syncthreads();
}

// Phases 2-4: evaluate passing rows, claim slots, write candidates.
//   val(r, c): class probability (c in [0, nc)) of tile row r -- already multiplied? no: raw prob.
//   put(r, c, v) / get(r, c): scratch for the product (aliases the tile).
//   box(r): xyxy of tile row r.
template <class Tile>
__device__ __forceinline__ void filter_backend(FilterSmem& S, Tile& T, const FilterArgs& A, int b,
                                               int seg, int row0) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nc = A.nc;
    const int npass = S.npass;
    // phase 2
    for (int i = w; i < npass; i += kWarps) {
        const int r = S.list[i];
        const float obj = S.obj[r];
        int count = 0;
        if (A.multi_label) {
            for (int c0 = 0; c0 < nc; c0 += 32) {
                const int c = c0 + lane;
                bool flag = false;
                if (c < nc) {
                    const float prod = __fmul_rn(T.prob(r, c), obj);      // image_proc.py:135
                    flag = (prod > A.conf) && class_allowed(A.class_mask, c);  // :141,151
                    T.put(r, c, flag ? prod : -1.0f);
                }
                count += __popc(__ballot_sync(0xffffffffu, flag));
            }
        } else {
            float bv = -INFINITY;
            int bj = 0x7fffffff;
            for (int c0 = 0; c0 < nc; c0 += 32) {
                const int c = c0 + lane;
                if (c < nc) {
                    const float prod = __fmul_rn(T.prob(r, c), obj);
                    if (prod > bv) { bv = prod; bj = c; }     // first max within the lane
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {                     // first max across lanes
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
                if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
            }
            const bool sel = (bj != 0x7fffffff) && (bv > A.conf) && class_allowed(A.class_mask, bj);  // :145-151
            count = sel ? 1 : 0;
            if (lane == 0) { S.best_v[r] = bv; S.best_j[r] = bj; }
        }
        if (lane == 0) S.cnt[r] = count;
    }
    __syncthreads();
    // phase 3: exclusive scan of the 64 row counts, one atomic per tile
    if (w == 0) {
        const int a = S.cnt[2 * lane], c = S.cnt[2 * lane + 1];
        const int inc = warp_incl_scan(a + c, lane);
        S.excl[2 * lane] = inc - a - c;
        S.excl[2 * lane + 1] = inc - c;
        if (lane == 31) {
            const int total = inc;
            const int base = total ? atomicAdd(A.counts + b, total) : 0;
            S.base = base;
            A.seg_base[(size_t)b * A.segs + seg] = base;
            A.seg_count[(size_t)b * A.segs + seg] = total;
        }
    }
    __syncthreads();
    // phase 4
    uint64_t* cand = A.cand + (size_t)b * A.cap;
    for (int i = w; i < npass; i += kWarps) {
        const int r = S.list[i];
        if (S.cnt[r] == 0) continue;
        const int row = row0 + r;
        int pos = S.base + S.excl[r];
        if (A.multi_label) {
            for (int c0 = 0; c0 < nc; c0 += 32) {
                const int c = c0 + lane;
                const float v = (c < nc) ? T.get(r, c) : -1.0f;
                const bool flag = v >= 0.0f;
                const unsigned m = __ballot_sync(0xffffffffu, flag);
                const int p = pos + __popc(m & ((1u << lane) - 1u));
                if (flag && p < A.cap)
                    cand[p] = ((uint64_t)(uint32_t)(row * nc + c) << 32) | __float_as_uint(v);
                pos += __popc(m);
            }
        } else if (lane == 0 && pos < A.cap) {
            cand[pos] = ((uint64_t)(uint32_t)(row * nc + S.best_j[r]) << 32) | __float_as_uint(S.best_v[r]);
        }
        if (lane == 0) A.boxes[(size_t)b * A.rows + row] = T.box(r);
    }
}

__device__ __forceinline__ float4 xyxy_from_cxcywh(float cx, float cy, float w, float h) {
    const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);  // utils/bboxes.py:103-111
    return make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
}

// tile of logits [no][kTilePitch] (fused path)
struct LogitTile {
    float* t;
    int variant, nx, s0;
    float stride, aw, ah;
    __device__ __forceinline__ float prob(int r, int c) const { return sigmoidf_vk(t[(5 + c) * kTilePitch + r]); }
    __device__ __forceinline__ void put(int r, int c, float v) { t[(5 + c) * kTilePitch + r] = v; }
    __device__ __forceinline__ float get(int r, int c) const { return t[(5 + c) * kTilePitch + r]; }
    __device__ __forceinline__ float4 box(int r) const {
        const int sp = s0 + r;
        const int gy = sp / nx, gx = sp - gy * nx;
        const float cx = decode_elem(t[0 * kTilePitch + r], 0, (float)gx, stride, aw, variant);
        const float cy = decode_elem(t[1 * kTilePitch + r], 1, (float)gy, stride, ah, variant);
        const float w = decode_elem(t[2 * kTilePitch + r], 2, 0.f, stride, aw, variant);
        const float h = decode_elem(t[3 * kTilePitch + r], 3, 0.f, stride, ah, variant);
        return xyxy_from_cxcywh(cx, cy, w, h);
    }
};

// tile of decoded prediction rows [kTileS][no] (drop-in path)
struct PredTile {
    float* t;
    int no;
    __device__ __forceinline__ float prob(int r, int c) const { return t[r * no + 5 + c]; }
    __device__ __forceinline__ void put(int r, int c, float v) { t[r * no + 5 + c] = v; }
    __device__ __forceinline__ float get(int r, int c) const { return t[r * no + 5 + c]; }
    __device__ __forceinline__ float4 box(int r) const {
        const float* p = t + r * no;
        return xyxy_from_cxcywh(p[0], p[1], p[2], p[3]);
    }
};

__global__ void __launch_bounds__(kDecThreads)
decode_filter_kernel(const HeadDev H, const FilterArgs A) {
    extern __shared__ float tile[];  // [no][kTilePitch]
    __shared__ FilterSmem S;
    const int b = blockIdx.y, seg = blockIdx.x;
    const TileLoc q = locate_tile(H, seg);
    const int no = H.no, nynx = H.nynx[q.l];
    const float* __restrict__ in = H.lv[q.l] + ((size_t)(b * H.na + q.a) * no) * nynx + q.s0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

    // phase 0: objectness plane of the tile (coalesced), image_proc.py:99
    bool pass = false;
    if (threadIdx.x < kTileS) {
        float obj = 0.f;
        if (threadIdx.x < q.nvalid) {
            obj = sigmoidf_vk(ld_stream_f32(in + (size_t)4 * nynx + threadIdx.x));
            pass = obj > A.conf;
        }
        S.obj[threadIdx.x] = obj;
    }
    const int npass = __syncthreads_count(pass);
    if (npass == 0) {
        if (threadIdx.x == 0) {
            A.seg_base[(size_t)b * A.segs + seg] = 0;
            A.seg_count[(size_t)b * A.segs + seg] = 0;
        }
        return;
    }
    if (threadIdx.x == 0) S.npass = npass;
    build_pass_list(S, pass);

    // phase 1: bring the logits of the surviving rows into shared memory
    if (npass * 8 > q.nvalid) {  // dense: coalesced planes
        const bool vec = ((nynx & 3) == 0) && ((reinterpret_cast<uintptr_t>(H.lv[q.l]) & 15) == 0);
        if (vec) {
            for (int e = threadIdx.x; e < no * (kTileS / 4); e += kDecThreads) {
                const int c = e >> 4, s = (e & 15) << 2;
                if (s < q.nvalid) {
                    const float4 v = ld_stream_f4(in + (size_t)c * nynx + s);
                    float* d = tile + c * kTilePitch + s;
                    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
                }
            }
        } else {
            for (int e = threadIdx.x; e < no * kTileS; e += kDecThreads) {
                const int c = e >> 6, s = e & 63;
                if (s < q.nvalid) tile[c * kTilePitch + s] = ld_stream_f32(in + (size_t)c * nynx + s);
            }
        }
    } else {                     // sparse: gather the few surviving rows, one sector per element
        for (int i = w; i < npass; i += kWarps) {
            const int r = S.list[i];
            for (int c = lane; c < no; c += 32)
                tile[c * kTilePitch + r] = __ldg(in + (size_t)c * nynx + r);
        }
    }
    __syncthreads();

    LogitTile T{tile, H.variant, H.nx[q.l], q.s0, H.stride[q.l], H.anchors[q.l][2 * q.a],
                H.anchors[q.l][2 * q.a + 1]};
    filter_backend(S, T, A, b, seg, q.row0);
}

__global__ void __launch_bounds__(kDecThreads)
filter_pred_kernel(const float* __restrict__ pred, int no, const FilterArgs A) {
    extern __shared__ float tile[];  // [kTileS][no]
    __shared__ FilterSmem S;
    const int b = blockIdx.y, seg = blockIdx.x;
    const int row0 = seg * kTileS;
    const int nvalid = min(kTileS, A.rows - row0);
    const float* __restrict__ in = pred + ((size_t)b * A.rows + row0) * no;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

    bool pass = false;
    if (threadIdx.x < kTileS) {
        float obj = 0.f;
        if (threadIdx.x < nvalid) {
            obj = __ldg(in + (size_t)threadIdx.x * no + 4);
            pass = obj > A.conf;                                   // image_proc.py:99
        }
        S.obj[threadIdx.x] = obj;
    }
    const int npass = __syncthreads_count(pass);
    if (npass == 0) {
        if (threadIdx.x == 0) {
            A.seg_base[(size_t)b * A.segs + seg] = 0;
            A.seg_count[(size_t)b * A.segs + seg] = 0;
        }
        return;
    }
    if (threadIdx.x == 0) S.npass = npass;
    build_pass_list(S, pass);

    if (npass * 4 > nvalid) {  // dense: the tile is one contiguous chunk of nvalid*no floats
        const int n = nvalid * no;
        if (((reinterpret_cast<uintptr_t>(in) & 15) == 0) && ((n & 3) == 0)) {
            for (int e = threadIdx.x; e < (n >> 2); e += kDecThreads)
                reinterpret_cast<float4*>(tile)[e] = ld_stream_f4(in + 4 * (size_t)e);
        } else {
            for (int e = threadIdx.x; e < n; e += kDecThreads) tile[e] = ld_stream_f32(in + e);
        }
    } else {
        for (int i = w; i < npass; i += kWarps) {
            const int r = S.list[i];
            for (int c = lane; c < no; c += 32) tile[r * no + c] = __ldg(in + (size_t)r * no + c);
        }
    }
    __syncthreads();
    PredTile T{tile, no};
    filter_backend(S, T, A, b, seg, row0);
}

static int make_head(const VkHeadCfg* cfg, HeadDev* H, const char* who) {
    if (!cfg) return fail_arg("%s: cfg is NULL", who);
    if (cfg->nl < 1 || cfg->nl > VK_MAX_LEVELS || cfg->na < 1 || cfg->na > VK_MAX_ANCHORS || cfg->nc < 1)
        return fail_code(VK_E_LIMIT, "%s: nl=%d na=%d nc=%d outside limits", who, cfg->nl, cfg->na, cfg->nc);
    if (cfg->variant != VK_HEAD_V5 && cfg->variant != VK_HEAD_V7) return fail_arg("%s: variant %d", who, cfg->variant);
    memset(H, 0, sizeof(*H));
    H->variant = cfg->variant; H->nl = cfg->nl; H->na = cfg->na; H->nc = cfg->nc; H->no = cfg->nc + 5;
    int rows = 0, tiles = 0;
    for (int l = 0; l < cfg->nl; ++l) {
        if (cfg->ny[l] <= 0 || cfg->nx[l] <= 0) return fail_arg("%s: level %d grid %dx%d", who, l, cfg->ny[l], cfg->nx[l]);
        H->ny[l] = cfg->ny[l]; H->nx[l] = cfg->nx[l]; H->nynx[l] = cfg->ny[l] * cfg->nx[l];
        H->stride[l] = cfg->stride[l];
        for (int k = 0; k < 2 * cfg->na; ++k) H->anchors[l][k] = cfg->anchors[l][k];
        H->row_base[l] = rows;
        H->tile_start[l] = tiles;
        H->tpa[l] = ceil_div(H->nynx[l], kTileS);
        rows += cfg->na * H->nynx[l];
        tiles += cfg->na * H->tpa[l];
    }
    for (int l = cfg->nl; l <= VK_MAX_LEVELS; ++l) H->tile_start[l] = tiles;
    H->rows = rows; H->tiles = tiles;
    if (tiles > VK_MAX_SEGMENTS) return fail_code(VK_E_LIMIT, "%s: %d tiles per image > %d", who, tiles, VK_MAX_SEGMENTS);
    return VK_OK;
}

static int check_cand(const VkCandBuf* o, int rows, int segs, int nc, const char* who) {
    if (!o || !o->cand || !o->boxes || !o->counts || !o->seg_base || !o->seg_count)
        return fail_arg("%s: candidate buffer has a NULL member", who);
    if (o->cap <= 0 || o->rows != rows || o->segs != segs || o->nc != nc)
        return fail_arg("%s: candidate buffer shape (cap=%d rows=%d segs=%d nc=%d) != (rows=%d segs=%d nc=%d)",
                        who, o->cap, o->rows, o->segs, o->nc, rows, segs, nc);
    if ((uint64_t)rows * (uint64_t)nc > 0xffffffffull) return fail_code(VK_E_LIMIT, "%s: rows*nc overflows 32 bits", who);
    if (reinterpret_cast<uintptr_t>(o->boxes) & 15) return fail_arg("%s: boxes must be 16-byte aligned", who);
    return VK_OK;
}

static FilterArgs make_filter_args(const VkCandBuf* o, float conf, int multi_label, const uint32_t* mask) {
    FilterArgs A;
    A.conf = conf;
    A.multi_label = (multi_label && o->nc > 1) ? 1 : 0;   // image_proc.py:111
    A.class_mask = mask;
    A.cand = o->cand;
    A.boxes = reinterpret_cast<float4*>(o->boxes);
    A.counts = o->counts; A.seg_base = o->seg_base; A.seg_count = o->seg_count;
    A.cap = o->cap; A.rows = o->rows; A.segs = o->segs; A.nc = o->nc;
    return A;
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
    const size_t smem = (size_t)H.no * kTilePitch * sizeof(float);
    if (smem > 200 * 1024) return fail_code(VK_E_LIMIT, "vk_detect_decode: nc=%d needs %zu B of shared memory", H.nc, smem);
    cudaFuncSetAttribute(detect_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    detect_decode_kernel<<<dim3(H.tiles, batch), kDecThreads, smem, as_stream(stream)>>>(H, pred);
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
    if (batch > 65535) return fail_code(VK_E_LIMIT, "vk_decode_filter: batch %d > 65535", batch);
    if (int rc = check_cand(out, H.rows, H.tiles, H.nc, "vk_decode_filter")) return rc;
    for (int l = 0; l < H.nl; ++l) {
        if (!levels[l]) return fail_arg("vk_decode_filter: level %d is NULL", l);
        H.lv[l] = levels[l];
    }
    cudaStream_t stream = as_stream(stream_);
    cudaError_t e = cudaMemsetAsync(out->counts, 0, (size_t)batch * sizeof(int32_t), stream);
    if (e != cudaSuccess) return fail_code((int)e, "vk_decode_filter: memset: %s", cudaGetErrorString(e));
    const size_t smem = (size_t)H.no * kTilePitch * sizeof(float);
    if (smem > 200 * 1024) return fail_code(VK_E_LIMIT, "vk_decode_filter: nc=%d needs %zu B of shared memory", H.nc, smem);
    cudaFuncSetAttribute(decode_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const FilterArgs A = make_filter_args(out, conf_thres, multi_label, class_mask);
    decode_filter_kernel<<<dim3(H.tiles, batch), kDecThreads, smem, stream>>>(H, A);
    count_launch();
    return check_launch("decode_filter_kernel");
}

extern "C" int vk_filter_pred(const float* pred, int batch, int rows, int nc, float conf_thres,
                              int multi_label, const uint32_t* class_mask, const VkCandBuf* out,
                              vk_stream_t stream_) {
    if (batch == 0) return VK_OK;
    if (!pred || batch < 0 || rows <= 0 || nc < 1) return fail_arg("vk_filter_pred: null/negative argument");
    if (batch > 65535) return fail_code(VK_E_LIMIT, "vk_filter_pred: batch %d > 65535", batch);
    const int segs = ceil_div(rows, kTileS);
    if (segs > VK_MAX_SEGMENTS) return fail_code(VK_E_LIMIT, "vk_filter_pred: %d rows > %d", rows, VK_MAX_SEGMENTS * kTileS);
    if (int rc = check_cand(out, rows, segs, nc, "vk_filter_pred")) return rc;
    cudaStream_t stream = as_stream(stream_);
    cudaError_t e = cudaMemsetAsync(out->counts, 0, (size_t)batch * sizeof(int32_t), stream);
    if (e != cudaSuccess) return fail_code((int)e, "vk_filter_pred: memset: %s", cudaGetErrorString(e));
    const int no = nc + 5;
    const size_t smem = (size_t)no * kTileS * sizeof(float);
    if (smem > 200 * 1024) return fail_code(VK_E_LIMIT, "vk_filter_pred: nc=%d needs %zu B of shared memory", nc, smem);
    cudaFuncSetAttribute(filter_pred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const FilterArgs A = make_filter_args(out, conf_thres, multi_label, class_mask);
    filter_pred_kernel<<<dim3(segs, batch), kDecThreads, smem, stream>>>(pred, no, A);
    count_launch();
    return check_launch("filter_pred_kernel");
}
