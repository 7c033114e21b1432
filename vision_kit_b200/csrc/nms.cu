// Per-image top-max_nms cut, score sort and greedy class-aware NMS (SURVEY.md §8 a5 second
// half, a5', a6): utils/image_proc.py:154-182 / demo/processing.py:167-197 with
// torchvision.ops.nms's CPU arithmetic.  One CTA (1024 threads) per image; everything between
// the candidate list and the (max_det, 6) output stays in that CTA's shared memory:
//
//   A1  canonical order: exclusive scan of the segment table written by the filter kernels
//   A2  n > max_nms only: exact radix select (11/11/10 bits) of the max_nms-th score; ties at
//       the cut are resolved by canonical position (= the stable argsort the contract fixes)
//   A3  ordered compaction of the selected candidates into shared (score, pos) arrays
//   A4  bitonic sort on (score desc, pos asc)
//   A5  greedy NMS over the sorted list in chunks of 256: each chunk is tested against the
//       kept list (<= max_det boxes in shared memory), then against itself through a 256x256
//       bit matrix resolved by one warp; stops as soon as max_det boxes are kept, which is
//       exact because greedy NMS visits boxes in descending score order (image_proc.py:170).
//
// IoU arithmetic: separate fp32 sub/mul/add/div (no FMA), strict '>' against the python-float
// threshold promoted to double -- implemented as '>' against the largest float32 <= threshold.
#include "vk_common.cuh"

#include <math.h>

namespace vk {

constexpr int kNmsThreads = 1024;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kChunk = 256;
constexpr int kChunkWords = kChunk / 32;
constexpr int kHistBins = 2048;

struct NmsArgs {
    const uint64_t* cand;
    const float4* boxes;
    const int32_t* counts;
    const int32_t* seg_base;
    const int32_t* seg_count;
    int cap, rows, segs, nc;
    float iou_thr;  // largest float <= the double threshold
    int agnostic, max_nms, max_det;
    float max_wh;
    float* dets;
    int32_t* det_counts;
    int64_t* keep_idx;
    int32_t* status;
    uint32_t* sel;  // [batch][P]: compacted position -> physical candidate slot
    int P;          // sort capacity, power of two
};

__device__ __forceinline__ uint32_t order_key(uint32_t fbits) {  // float order -> unsigned order
    return fbits ^ ((fbits >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ uint32_t unorder_key(uint32_t k) {
    return k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu);
}

// torchvision/csrc/ops/cpu/nms_kernel.cpp: suppress j when inter/(area_i+area_j-inter) > thr.
__device__ __forceinline__ bool iou_exceeds(const float4 a, const float aa, const float4 b,
                                            const float ab, const float thr) {
    const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    const float inter = __fmul_rn(w, h);
    if (!(inter > 0.f)) return false;  // 0/x is 0, -0 or NaN: never > thr (thr >= 0)
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
    return ovr > thr;
}

struct NmsScratch {              // lives after the sort arrays in dynamic shared memory
    union {
        struct {                 // A1-A3
            int segoff[VK_MAX_SEGMENTS + 1];
            int segbase[VK_MAX_SEGMENTS];
            union {
                int hist[kHistBins];
                struct { int gt[VK_MAX_SEGMENTS]; int eq[VK_MAX_SEGMENTS]; } c;
            } u;
        } a;
        struct {                 // A5
            float4 kbox[VK_MAX_DET];
            float karea[VK_MAX_DET];
            float4 cbox[kChunk];
            float carea[kChunk];
            uint32_t mask[kChunk][kChunkWords];
            uint8_t csup[kChunk];
            uint8_t newk[kChunk];
        } b;
        unsigned long long xchg[kNmsThreads];  // A4 (small sorts): cross-warp exchange
    };
};

__global__ void __launch_bounds__(kNmsThreads, 1)
nms_image_kernel(const NmsArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int wsum[33];
    __shared__ int s_found_bin, s_found_above, s_kept, s_newcount;

    const int P = A.P;
    uint32_t* skey = reinterpret_cast<uint32_t*>(smem_raw);
    uint16_t* spos = reinterpret_cast<uint16_t*>(smem_raw + (size_t)P * 4);
    NmsScratch& X = *reinterpret_cast<NmsScratch*>(smem_raw + (size_t)P * 6);

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t* cand = A.cand + (size_t)b * A.cap;
    const float4* boxes = A.boxes + (size_t)b * A.rows;
    const int32_t* seg_base = A.seg_base + (size_t)b * A.segs;
    const int32_t* seg_count = A.seg_count + (size_t)b * A.segs;
    uint32_t* sel = A.sel + (size_t)b * P;
    float* dets = A.dets + (size_t)b * A.max_det * 6;
    int64_t* keep_out = A.keep_idx ? A.keep_idx + (size_t)b * A.max_det : nullptr;

    // outputs start zeroed / -1 padded (the reference returns exactly k rows; the host slices)
    for (int i = tid; i < A.max_det * 6; i += kNmsThreads) dets[i] = 0.f;
    if (keep_out)
        for (int i = tid; i < A.max_det; i += kNmsThreads) keep_out[i] = -1;

    const int n_total = A.counts[b];
    const int n = min(n_total, A.cap);
    if (tid == 0 && A.status) A.status[b] = (n_total > A.cap) ? 1 : 0;
    if (n == 0) {
        if (tid == 0) A.det_counts[b] = 0;
        return;
    }

    // ---------------- A1: canonical offsets of the segments
    {
        int carry = 0;
        for (int t0 = 0; t0 < A.segs; t0 += kNmsThreads) {
            const int t = t0 + tid;
            int c = 0;
            if (t < A.segs) {
                c = seg_count[t];
                const int sb = seg_base[t];
                X.a.segbase[t] = sb;
                if (c > 0) c = max(0, min(c, A.cap - sb));  // overflowed tail was never written
            }
            int total;
            const int ex = block_excl_scan(c, wsum, &total);
            if (t < A.segs) X.a.segoff[t] = carry + ex;
            carry += total;
        }
        if (tid == 0) X.a.segoff[A.segs] = carry;
        __syncthreads();
    }

    // ---------------- A2: exact selection of the max_nms-th best score (image_proc.py:161-163)
    const int K = A.max_nms;
    const bool cut = n > K;
    uint32_t tval = 0;
    int need_eq = 0;
    if (cut) {
        uint32_t prefix = 0, pmask = 0;
        int remaining = K;
        const int shifts[3] = {21, 10, 0};
        const int widths[3] = {11, 11, 10};
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
            const int shift = shifts[pass], nb = 1 << widths[pass];
            for (int i = tid; i < kHistBins; i += kNmsThreads) X.a.u.hist[i] = 0;
            __syncthreads();
            for (int i = tid; i < n; i += kNmsThreads) {
                const uint32_t key = order_key((uint32_t)cand[i]);
                if ((key & pmask) == prefix) atomicAdd(&X.a.u.hist[(key >> shift) & (nb - 1)], 1);
            }
            __syncthreads();
            // bins from the top: thread t owns bins nb-1-2t and nb-2-2t
            const int b0 = nb - 1 - 2 * tid, b1 = nb - 2 - 2 * tid;
            const int v0 = (b0 >= 0) ? X.a.u.hist[b0] : 0;
            const int v1 = (b1 >= 0) ? X.a.u.hist[b1] : 0;
            int total;
            const int above = block_excl_scan(v0 + v1, wsum, &total);
            if (above < remaining && remaining <= above + v0) {
                s_found_bin = b0; s_found_above = above;
            } else if (above + v0 < remaining && remaining <= above + v0 + v1) {
                s_found_bin = b1; s_found_above = above + v0;
            }
            __syncthreads();
            prefix |= (uint32_t)s_found_bin << shift;
            pmask |= (uint32_t)(nb - 1) << shift;
            remaining -= s_found_above;
            __syncthreads();
        }
        tval = prefix;        // the max_nms-th best (ordered) score
        need_eq = remaining;  // how many candidates equal to it make the cut (lowest position first)
    }

    // ---------------- A3: ordered compaction into shared memory
    // sel[pos] = row*nc + cls of the candidate at compacted position pos
    const int M = cut ? K : n;
    if (!cut) {
        // every candidate is selected: one thread per canonical position, segment by binary search
        for (int p = tid; p < n; p += kNmsThreads) {
            int lo = 0, hi = A.segs;            // last t with segoff[t] <= p
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (X.a.segoff[mid] <= p) lo = mid; else hi = mid;
            }
            const uint64_t cd = cand[X.a.segbase[lo] + (p - X.a.segoff[lo])];
            skey[p] = order_key((uint32_t)cd);
            spos[p] = (uint16_t)p;
            sel[p] = (uint32_t)(cd >> 32);
        }
    } else {
        for (int t = warp; t < A.segs; t += kNmsWarps) {
            const int cnt = X.a.segoff[t + 1] - X.a.segoff[t];
            const int base = X.a.segbase[t];
            int gt = 0, eq = 0;
            for (int j0 = 0; j0 < cnt; j0 += 32) {
                const int j = j0 + lane;
                uint32_t key = 0;
                const bool ok = j < cnt;
                if (ok) key = order_key((uint32_t)cand[base + j]);
                gt += __popc(__ballot_sync(0xffffffffu, ok && key > tval));
                eq += __popc(__ballot_sync(0xffffffffu, ok && key == tval));
            }
            if (lane == 0) { X.a.u.c.gt[t] = gt; X.a.u.c.eq[t] = eq; }
        }
        __syncthreads();
        int carry_gt = 0, carry_eq = 0;
        for (int t0 = 0; t0 < A.segs; t0 += kNmsThreads) {
            const int t = t0 + tid;
            const int g = (t < A.segs) ? X.a.u.c.gt[t] : 0;
            const int e = (t < A.segs) ? X.a.u.c.eq[t] : 0;
            int tg, te;
            const int xg = block_excl_scan(g, wsum, &tg);
            const int xe = block_excl_scan(e, wsum, &te);
            if (t < A.segs) { X.a.u.c.gt[t] = carry_gt + xg; X.a.u.c.eq[t] = carry_eq + xe; }
            carry_gt += tg; carry_eq += te;
        }
        __syncthreads();
        for (int t = warp; t < A.segs; t += kNmsWarps) {
            const int cnt = X.a.segoff[t + 1] - X.a.segoff[t];
            if (cnt == 0) continue;
            const int base = X.a.segbase[t];
            int eq_before = X.a.u.c.eq[t];
            int pos0 = X.a.u.c.gt[t] + min(eq_before, need_eq);
            for (int j0 = 0; j0 < cnt; j0 += 32) {
                const int j = j0 + lane;
                const bool ok = j < cnt;
                uint64_t cd = 0;
                if (ok) cd = cand[base + j];
                const uint32_t key = order_key((uint32_t)cd);
                const bool is_eq = ok && key == tval;
                const unsigned em = __ballot_sync(0xffffffffu, is_eq);
                const int eq_rank = eq_before + __popc(em & ((1u << lane) - 1u));
                const bool take = ok && (key > tval || (is_eq && eq_rank < need_eq));
                eq_before += __popc(em);
                const unsigned tm = __ballot_sync(0xffffffffu, take);
                const int pos = pos0 + __popc(tm & ((1u << lane) - 1u));
                if (take) {
                    skey[pos] = key;
                    spos[pos] = (uint16_t)pos;
                    sel[pos] = (uint32_t)(cd >> 32);
                }
                pos0 += __popc(tm);
            }
        }
    }
    int Ps = 32;
    while (Ps < M) Ps <<= 1;
    for (int i = M + tid; i < Ps; i += kNmsThreads) { skey[i] = 0u; spos[i] = 0xffffu; }
    __syncthreads();

    // ---------------- A4: bitonic sort, "before" = higher score, then lower position
    if (Ps <= kNmsThreads) {
        // one element per thread in a register: composite = score<<16 | (0xffff - pos), sorted
        // descending; strides below 32 exchange by shuffle, the rest through shared memory
        unsigned long long v = 0ull;
        if (tid < Ps) v = ((unsigned long long)skey[tid] << 16) | (unsigned long long)(0xffffu - spos[tid]);
        __syncthreads();   // skey/spos read before the scratch is reused as exchange buffer
        for (int k = 2; k <= Ps; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                unsigned long long o;
                if (j >= 32) {
                    X.xchg[tid] = v;
                    __syncthreads();
                    o = X.xchg[tid ^ j];
                    __syncthreads();
                } else {
                    o = __shfl_xor_sync(0xffffffffu, v, j);
                }
                const bool keep_max = ((tid & j) == 0) == ((tid & k) == 0);
                v = keep_max ? (v > o ? v : o) : (v < o ? v : o);
            }
        }
        if (tid < Ps) {
            skey[tid] = (uint32_t)(v >> 16);
            spos[tid] = (uint16_t)(0xffffu - (uint32_t)(v & 0xffffull));
        }
        __syncthreads();
    } else {
        for (int k = 2; k <= Ps; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < (Ps >> 1); i += kNmsThreads) {
                    const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                    const int hi = lo | j;
                    const uint32_t ka = skey[lo], kb = skey[hi];
                    const uint16_t pa = spos[lo], pb = spos[hi];
                    const bool hi_before_lo = (kb > ka) || (kb == ka && pb < pa);
                    const bool descending_block = (lo & k) == 0;  // final order: "before" first
                    if (hi_before_lo == descending_block) {
                        skey[lo] = kb; skey[hi] = ka;
                        spos[lo] = pb; spos[hi] = pa;
                    }
                }
                __syncthreads();
            }
        }
    }

    // ---------------- A5: greedy NMS with a kept list, early exit at max_det
    if (tid == 0) s_kept = 0;
    __syncthreads();
    const float thr = A.iou_thr;
    for (int chunk0 = 0; chunk0 < M; chunk0 += kChunk) {
        const int kept0 = s_kept;
        if (kept0 >= A.max_det) break;
        const int cn = min(kChunk, M - chunk0);
        if (tid < kChunk) {
            if (tid < cn) {
                const uint32_t idx = sel[spos[chunk0 + tid]];
                const uint32_t row = idx / (uint32_t)A.nc;
                const float cls = (float)(idx - row * (uint32_t)A.nc);
                const float4 bx = boxes[row];
                const float off = A.agnostic ? 0.f : __fmul_rn(cls, A.max_wh);     // image_proc.py:166
                const float4 ob = make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off),
                                              __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));  // :168
                X.b.cbox[tid] = ob;
                X.b.carea[tid] = __fmul_rn(__fsub_rn(ob.z, ob.x), __fsub_rn(ob.w, ob.y));
            }
            X.b.csup[tid] = 0;
        }
        __syncthreads();
        {   // phase 1: chunk box t vs the kept list, 4 threads per box
            const int t = tid >> 2, q = tid & 3;
            bool sup = false;
            if (t < cn) {
                const float4 cb = X.b.cbox[t];
                const float ca = X.b.carea[t];
                for (int k = q; k < kept0 && !sup; k += 4)
                    sup = iou_exceeds(X.b.kbox[k], X.b.karea[k], cb, ca, thr);
            }
            sup |= __shfl_xor_sync(0xffffffffu, sup, 1);
            sup |= __shfl_xor_sync(0xffffffffu, sup, 2);
            if (q == 0 && t < cn && sup) X.b.csup[t] = 1;
        }
        __syncthreads();
        {   // phase 2: upper-triangular bit matrix of the chunk
            const int i = tid >> 2, q = tid & 3;
            if (i < cn) {
                const bool live = !X.b.csup[i];
                const float4 ib = X.b.cbox[i];
                const float ia = X.b.carea[i];
#pragma unroll
                for (int ww = 0; ww < 2; ++ww) {
                    const int wd = q * 2 + ww;
                    uint32_t m = 0;
                    if (live && wd * 32 + 31 > i) {
                        for (int bit = 0; bit < 32; ++bit) {
                            const int j = wd * 32 + bit;
                            if (j > i && j < cn && iou_exceeds(ib, ia, X.b.cbox[j], X.b.carea[j], thr))
                                m |= 1u << bit;
                        }
                    }
                    X.b.mask[i][wd] = m;
                }
            }
        }
        __syncthreads();
        if (warp == 0) {  // phase 3: sequential resolve; every lane carries all 256 alive bits
            uint32_t alive[kChunkWords];
#pragma unroll
            for (int wd = 0; wd < kChunkWords; ++wd) {
                // lane `bit` of each ballot = "box wd*32+bit is a live candidate"
                const int j = wd * 32 + lane;
                alive[wd] = __ballot_sync(0xffffffffu, j < cn && !X.b.csup[j]);
            }
            int kept = kept0, newc = 0;
            bool done = false;
            const uint4* mrow = reinterpret_cast<const uint4*>(&X.b.mask[0][0]);
            uint4 n0 = mrow[0], n1 = mrow[1];
#pragma unroll
            for (int wd = 0; wd < kChunkWords; ++wd) {
                if (done || wd * 32 >= cn) break;
                const int lim = min(32, cn - wd * 32);
                for (int bit = 0; bit < lim; ++bit) {
                    const int i = wd * 32 + bit;
                    const uint4 m0 = n0, m1 = n1;
                    const int nx = min(i + 1, kChunk - 1);   // prefetch the next row
                    n0 = mrow[2 * nx]; n1 = mrow[2 * nx + 1];
                    if ((alive[wd] >> bit) & 1u) {
                        if (lane == 0) X.b.newk[newc] = (uint8_t)i;
                        ++newc; ++kept;
                        alive[0] &= ~m0.x; alive[1] &= ~m0.y; alive[2] &= ~m0.z; alive[3] &= ~m0.w;
                        alive[4] &= ~m1.x; alive[5] &= ~m1.y; alive[6] &= ~m1.z; alive[7] &= ~m1.w;
                        if (kept >= A.max_det) { done = true; break; }
                    }
                }
            }
            if (lane == 0) { s_kept = kept; s_newcount = newc; }
        }
        __syncthreads();
        {   // phase 4: append to the kept list and emit the detections
            const int newc = s_newcount;
            if (tid < newc) {
                const int i = X.b.newk[tid];
                const int slot = kept0 + tid;
                X.b.kbox[slot] = X.b.cbox[i];
                X.b.karea[slot] = X.b.carea[i];
                const uint32_t idx = sel[spos[chunk0 + i]];
                const uint32_t row = idx / (uint32_t)A.nc;
                const float4 bx = boxes[row];
                float* o = dets + (size_t)slot * 6;     // image_proc.py:182 output[xi] = x[i]
                o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
                o[4] = __uint_as_float(unorder_key(skey[chunk0 + i]));
                o[5] = (float)(idx - row * (uint32_t)A.nc);
                if (keep_out) keep_out[slot] = cut ? (int64_t)(chunk0 + i) : (int64_t)spos[chunk0 + i];
            }
        }
        __syncthreads();
    }
    if (tid == 0) A.det_counts[b] = s_kept;
}

static int next_pow2(int v) {
    int p = 32;
    while (p < v) p <<= 1;
    return p;
}

static int sort_capacity(int max_nms, int cap) { return next_pow2(max_nms < cap ? max_nms : cap); }

}  // namespace vk

using namespace vk;

extern "C" size_t vk_nms_workspace_bytes(int batch, int max_nms) {
    if (batch <= 0 || max_nms <= 0 || max_nms > VK_MAX_NMS) return 0;
    return (size_t)batch * next_pow2(max_nms) * sizeof(uint32_t);
}

extern "C" int vk_nms_batched(const VkCandBuf* c, int batch, float conf_unused, double iou_thres,
                              int agnostic, int max_nms, int max_det, float max_wh, float* dets,
                              int32_t* det_counts, int64_t* keep_idx, int32_t* status, void* ws,
                              size_t ws_bytes, vk_stream_t stream) {
    (void)conf_unused;
    if (batch == 0) return VK_OK;
    if (!c || !c->cand || !c->boxes || !c->counts || !c->seg_base || !c->seg_count || !dets || !det_counts || batch < 0)
        return fail_arg("vk_nms_batched: null/negative argument");
    if (max_nms < 1 || max_nms > VK_MAX_NMS) return fail_code(VK_E_LIMIT, "vk_nms_batched: max_nms %d outside [1,%d]", max_nms, VK_MAX_NMS);
    if (max_det < 1 || max_det > VK_MAX_DET) return fail_code(VK_E_LIMIT, "vk_nms_batched: max_det %d outside [1,%d]", max_det, VK_MAX_DET);
    if (c->segs < 1 || c->segs > VK_MAX_SEGMENTS) return fail_code(VK_E_LIMIT, "vk_nms_batched: %d segments outside [1,%d]", c->segs, VK_MAX_SEGMENTS);
    if (c->cap < 1 || c->rows < 1 || c->nc < 1) return fail_arg("vk_nms_batched: bad candidate buffer shape");
    if (!(iou_thres >= 0.0 && iou_thres <= 1.0)) return fail_arg("vk_nms_batched: iou_thres %g outside [0,1]", iou_thres);
    const int P = sort_capacity(max_nms, c->cap);
    if (!ws || ws_bytes < (size_t)batch * P * sizeof(uint32_t))
        return fail_code(VK_E_WORKSPACE, "vk_nms_batched: workspace %zu < %zu", ws_bytes, (size_t)batch * P * sizeof(uint32_t));
    NmsArgs A;
    A.cand = c->cand; A.boxes = reinterpret_cast<const float4*>(c->boxes); A.counts = c->counts;
    A.seg_base = c->seg_base; A.seg_count = c->seg_count;
    A.cap = c->cap; A.rows = c->rows; A.segs = c->segs; A.nc = c->nc;
    float thr = (float)iou_thres;                       // double compare == float compare against
    if ((double)thr > iou_thres) thr = nextafterf(thr, -INFINITY);  // the largest float <= threshold
    A.iou_thr = thr;
    A.agnostic = agnostic ? 1 : 0; A.max_nms = max_nms; A.max_det = max_det; A.max_wh = max_wh;
    A.dets = dets; A.det_counts = det_counts; A.keep_idx = keep_idx; A.status = status;
    A.sel = static_cast<uint32_t*>(ws); A.P = P;
    const size_t smem = (size_t)P * 6 + sizeof(NmsScratch);
    cudaError_t e = cudaFuncSetAttribute(nms_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail_code((int)e, "vk_nms_batched: %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    nms_image_kernel<<<batch, kNmsThreads, smem, as_stream(stream)>>>(A);
    count_launch();
    return check_launch("nms_image_kernel");
}
