// Per-image top-max_nms cut, score sort and greedy class-aware NMS (SURVEY.md §8 a5 second
// half, a5', a6): utils/image_proc.py:154-182 / demo/processing.py:167-197 with
// torchvision.ops.nms's CPU arithmetic.
//
// Greedy NMS consumes candidates in descending score order and stops at max_det kept boxes
// (image_proc.py:170 truncates afterwards), so the sorted order is only ever needed for a prefix.
// Candidates carry the 64-bit key (ordered score << 32 | ~id), id = row * nc + cls: ascending id is the
// reference's candidate order (`nonzero` order, row then class), so descending keys = the stable descending
// argsort the reference's cut and torchvision's NMS are defined on, and keys are unique.  The key is all a
// candidate is: its box, class and score follow from it, whatever slot a filter kernel wrote it to.
//
//   nms_select_kernel   images with more candidates than their list holds (eval thresholds, ~240 k): a
//                       sampled score histogram gives the largest bound whose estimated count fits half
//                       the list, and a grid-wide streaming pass copies every candidate at or above it
//                       into the image's list.  All SMs work on every image, whatever the batch size.
//   nms_kernel          one CTA per image.  Stages of <= CAP candidates: a radix selection over the
//                       key bits (11/11/10 bits per pass, stopping at the first prefix whose count
//                       fits a stage) finds the stage's lower bound, one pass compacts the stage into
//                       shared memory, a counting sort orders it, boxes are fetched once, and chunks
//                       of 256 run:
//         1. each chunk box against the kept boxes of earlier chunks
//         2. predecessor bit matrix pred[i] = { j < i in the chunk : IoU(j, i) > thr }
//         3. fixed-point resolve: an undecided box with a kept predecessor is removed, one
//            whose predecessors are all removed is kept.  The lowest undecided box is always
//            decidable, so this ends with exactly the sequential greedy result after as many
//            rounds as the longest suppression chain -- not after 256 dependent steps.
//       The stage source is the image's list (built by the select pass) while it lasts, then the
//       segments of the candidate buffer; a small image (demo thresholds) is one stage read directly.  Because keys
//       are unique the selection always terminates: a tie group of thousands of bit-identical
//       scores is split by candidate id.  The exact cut at max_nms falls out of the order (the stage
//       crossing rank max_nms is truncated).
//       Class-aware mode walks per-class hash lists (kept boxes and chunk boxes) so that only
//       same-class pairs are ever tested.  That is exact while every coordinate seen so far
//       lies within +-max_wh/2 (offset boxes of different classes are then disjoint); the
//       first box outside that range switches the image to testing all pairs.
//
// IoU arithmetic: separate fp32 sub/mul/add/div (no FMA), strict '>' against the python-float
// threshold promoted to double -- implemented as '>' against the largest float32 <= threshold.
#include "decode_common.cuh"

#include <math.h>

namespace vk {

constexpr int kChunk = 256;
constexpr int kChunkWords = kChunk / 32;
constexpr int kHistBins = 2048;
constexpr int kHash = 256;          // class hash buckets
constexpr uint32_t kNil = 0xffffu;  // end of a hash list
constexpr int kSelThreads = 256;
constexpr int kSelStage = 128;      // keys a warp of the select pass stages before it reserves list space

struct NmsArgs {
    const uint64_t* cand;
    const float4* boxes;
    const int32_t* counts;      // ctrl row 0
    const int32_t* flags;       // ctrl row 1
    int32_t* list_count;        // ctrl row 2
    int32_t* bound;             // ctrl row 3 (ordered score bits)
    const int32_t* seg_count;
    uint64_t* list;            // [batch][list_cap] keys
    int cap, rows, segs, nc, list_cap;
    float iou_thr;  // largest float <= the double threshold
    int agnostic, max_nms, max_det;
    int prune;      // agnostic NMS: candidates of rows that are already decided are skipped (see nms_kernel)
    float max_wh;
    float* dets;
    int32_t* det_counts;
    int64_t* keep_idx;
    int32_t* status;
};

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

// ---------------------------------------------------------------------------------------
// selection pass (eval thresholds)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int tile_slots_of(int flags) { return (flags >> 8) << 6; }

// true when image b needs the select pass: it has a list, and more candidates than the list holds
__device__ __forceinline__ bool wants_select(const NmsArgs& A, int b) {
    return A.list != nullptr && A.counts[b] > A.list_cap;
}

// Work items of one image for the two passes below: segment t contributes ceil(count / piece) items of
// `piece` consecutive candidates, so warps get equal shares whatever the spread of the segment counts
// (a tile on a cluster of objects holds thousands of candidates, most tiles a few hundred).
// Fills s_cnt[t] and the exclusive item prefix s_pre[t] (s_pre[segs] = total) and returns the total.
__device__ __forceinline__ int build_items(const int32_t* __restrict__ seg_count, int segs, int piece, int* s_cnt,
                                           int* s_pre, int* wsum) {
    int carry = 0;
    for (int t0 = 0; t0 < segs; t0 += kSelThreads) {
        const int t = t0 + threadIdx.x;
        const int c = (t < segs) ? seg_count[t] : 0;
        int total;
        const int ex = block_excl_scan((c + piece - 1) / piece, wsum, &total);
        if (t < segs) { s_cnt[t] = c; s_pre[t] = carry + ex; }
        carry += total;
    }
    if (threadIdx.x == 0) s_pre[segs] = carry;
    __syncthreads();
    return carry;
}
__device__ __forceinline__ int item_segment(const int* s_pre, int segs, int item) {   // last t with s_pre[t] <= item
    int lo = 0, hi = segs;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_pre[mid] <= item) lo = mid; else hi = mid;
    }
    return lo;
}

// Score bound + list of one image, one streaming pass.  The work items of an image are dealt out round-robin
// to its blocks, so a block's share is a uniform sample of the image.  Each block first reads every 8th
// 128-byte line of ITS OWN items (one round trip: ~10 loads in flight per lane), estimates the image's
// score distribution from that (x 8 x blocks), and picks the last histogram bin whose estimated cumulative
// count fits HALF the list (~40 samples at that rank: overflowing the list would need an error of 2x).
// Blocks of one image may pick neighbouring bins; the image's bound is the highest of them (atomicMax), every
// block appends everything at or above ITS bound, so the list holds every candidate at or above the image's
// bound plus a few below it, which nms_kernel skips.  Most candidates of an eval-mode image share a handful
// of low-score bins, so lanes with the same bin elect one of them to add the group's size (match.any)
// instead of colliding 32 ways on one shared-memory counter.
constexpr int kSampleEvery = 8;
constexpr int kSelectPiece = 256;                     // candidates per work item: 16 lines, 8 loads per lane
constexpr int kItemTable = 6144;                      // work items whose segment is looked up in a table
__global__ void __launch_bounds__(kSelThreads)
nms_select_kernel(const NmsArgs A) {
    const int b = blockIdx.y;
    if (!wants_select(A, b)) return;
    __shared__ uint32_t s_hist[VK_HIST_BINS];
    __shared__ int s_cnt[VK_MAX_SEGMENTS], s_pre[VK_MAX_SEGMENTS + 1], wsum[33];
    __shared__ int s_j;
    __shared__ unsigned long long s_stage[kSelThreads / 32][kSelStage];
    __shared__ unsigned short s_item_seg[kItemTable];           // item -> segment (a binary search per item otherwise)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile_cap = tile_slots_of(A.flags[b]);
    const uint2* cand = reinterpret_cast<const uint2*>(A.cand + (size_t)b * A.cap);
    for (int i = threadIdx.x; i < VK_HIST_BINS; i += kSelThreads) s_hist[i] = 0u;
    const int items = build_items(A.seg_count + (size_t)b * A.segs, A.segs, kSelectPiece, s_cnt, s_pre, wsum);
    const bool table = items <= kItemTable;
    if (table) {
        for (int t = threadIdx.x; t < A.segs; t += kSelThreads)
            for (int i = s_pre[t]; i < s_pre[t + 1]; ++i) s_item_seg[i] = (unsigned short)t;
        __syncthreads();
    }
    auto seg_of = [&](int item) { return table ? (int)s_item_seg[item] : item_segment(s_pre, A.segs, item); };
    const int first = blockIdx.x * (kSelThreads / 32) + warp, step = gridDim.x * (kSelThreads / 32);
    // ---- sample: lines 2 and 10 of each of this warp's items (16 lines per item), 32 lanes = 2 lines
    for (int item0 = first; item0 < items; item0 += 4 * step) {          // four items' loads in flight
        int bin[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int item = item0 + u * step;
            bin[u] = -1;
            if (item < items) {
                const int seg = seg_of(item);
                const int j = (item - s_pre[seg]) * kSelectPiece + 16 * (2 + kSampleEvery * (lane >> 4)) + (lane & 15);
                if (j < s_cnt[seg]) bin[u] = hist_bin(cand[(size_t)seg * tile_cap + j].x);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned peers = __match_any_sync(0xffffffffu, bin[u]);
            if (bin[u] >= 0 && lane == __ffs(peers) - 1) atomicAdd(&s_hist[bin[u]], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    // ---- this block's bound: the last bin whose estimated cumulative count fits half the list
    {
        constexpr int BPT = VK_HIST_BINS / kSelThreads;
        uint32_t v[BPT];
        int sum = 0;
#pragma unroll
        for (int k = 0; k < BPT; ++k) { v[k] = s_hist[BPT * threadIdx.x + k]; sum += (int)v[k]; }
        if (threadIdx.x == 0) s_j = -1;
        int total;
        int run = block_excl_scan(sum, wsum, &total);
        const long limit = ((long)A.list_cap / 2) / ((long)kSampleEvery * gridDim.x);
        int best = -1;
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
            run += (int)v[k];
            if (run <= limit) best = BPT * threadIdx.x + k;
        }
        if (best >= 0) atomicMax(&s_j, best);
        __syncthreads();
    }
    const int jb = s_j;
    const uint32_t bound = jb < 0 ? 0xffffffffu : hist_bound(jb);   // no bin fits: nothing from this block, and the image's bound says so
    if (threadIdx.x == 0) {
        atomicMax(reinterpret_cast<unsigned int*>(A.bound) + b, bound);
        atomicOr(const_cast<int32_t*>(A.flags) + b, VK_FLAG_LIST);
    }
    if (jb < 0) return;
    // ---- streaming pass over this block's items.  Selected keys are staged per warp in shared memory and
    //      flushed with ONE reservation on the image's list counter per ~100 keys: a reservation per 32
    //      candidates would serialise thousands of same-address atomics per image.
    unsigned long long* stage = s_stage[warp];
    int staged = 0;                                               // warp-uniform
    uint64_t* list = A.list + (size_t)b * A.list_cap;
    const unsigned lt = (1u << lane) - 1u;
    auto flush = [&]() {
        int base = 0;
        if (lane == 0) base = atomicAdd(&A.list_count[b], staged);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        for (int i = lane; i < staged; i += 32)
            if (base + i < A.list_cap) list[base + i] = stage[i];
        __syncwarp();
        staged = 0;
    };
    // (the loads of a warp's next item are issued before the current one is examined; the two register buffers swap
    // roles instead of being copied.)  Scores of candidates are positive floats (p > conf >= 0), whose bit patterns
    // order like their ordered keys, so the test is one unsigned compare on the raw bits: key >= bound  <=>
    // bits >= braw.  A lane's candidates of an item sit 32 slots apart from its first one: one base pointer and
    // immediate offsets, and `32 u < rem` says which of them exist.
    const uint32_t braw = bound > 0x80000000u ? bound - 0x80000000u : 0u;
    struct Item { const uint2* p; int rem; };
    auto locate = [&](int item) {
        const int seg = seg_of(item);
        const int j0 = (item - s_pre[seg]) * kSelectPiece + lane;
        return Item{cand + ((uint32_t)seg * (uint32_t)tile_cap + (uint32_t)j0), s_cnt[seg] - j0};
    };
    auto fetch = [&](const Item& it, uint2* sc) {
#pragma unroll
        for (int u = 0; u < 8; ++u) sc[u] = (32 * u < it.rem) ? it.p[32 * u] : make_uint2(0u, 0u);
    };
    auto examine = [&](const Item& cur, const uint2* sc) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const bool take = 32 * u < cur.rem && sc[u].x >= braw;
            const unsigned m = __ballot_sync(0xffffffffu, take);
            if (m) {
                if (take) stage[staged + __popc(m & lt)] = ((unsigned long long)order_key(sc[u].x) << 32) | (uint32_t)~sc[u].y;
                staged += __popc(m);
                if (staged > kSelStage - 32) flush();
            }
        }
    };
    if (first < items) {
        Item ia = locate(first), ib = ia;
        uint2 sa[8], sb[8];
        fetch(ia, sa);
        for (int item = first;; item += 2 * step) {
            const bool more1 = item + step < items;
            if (more1) { ib = locate(item + step); fetch(ib, sb); }
            examine(ia, sa);
            if (!more1) break;
            const bool more2 = item + 2 * step < items;
            if (more2) { ia = locate(item + 2 * step); fetch(ia, sa); }
            examine(ib, sb);
            if (!more2) break;
        }
    }
    if (staged) flush();
}

// ---------------------------------------------------------------------------------------
// per-image kernel
// ---------------------------------------------------------------------------------------
struct ScratchB {
    float4* kbox;      // [max_det] kept boxes (class-offset)
    uint32_t* pred;    // [kChunk][kChunkWords]
    uint32_t* kmeta;   // [max_det] cls << 16 | next kept slot of the bucket
    uint32_t* kkeep;   // [max_det] rank (cut) or id of the kept candidate
    int* khead;        // [kHash] newest kept slot per class bucket
    uint32_t* ccnt2;   // [kHash / 2] chunk members per class bucket, two 16-bit counts per word
    uint16_t* cstart;  // [kHash + 2] first member slot of each bucket (+ end)
    uint16_t* members; // [kChunk] chunk rows grouped by class bucket
    uint32_t* kw;      // [2][3][kChunkWords] ping-pong kept | removed | undecided bits
    uint32_t* kfinal;  // [kChunkWords] kept bits of the resolved chunk
    uint8_t* state;    // [kChunk] 0 undecided, 1 kept, 2 removed
    __host__ __device__ static size_t bytes(int max_det) {
        return (size_t)max_det * 16 + kChunk * kChunkWords * 4 + (size_t)max_det * 8 + kHash * 4 + kHash * 2 +
               (kHash + 2) * 2 + kChunk * 2 + kChunkWords * 4 * 7 + kChunk;
    }
    __device__ ScratchB(unsigned char* p, int max_det) {
        kbox = reinterpret_cast<float4*>(p);
        pred = reinterpret_cast<uint32_t*>(kbox + max_det);
        kmeta = pred + kChunk * kChunkWords;
        kkeep = kmeta + max_det;
        khead = reinterpret_cast<int*>(kkeep + max_det);
        ccnt2 = reinterpret_cast<uint32_t*>(khead + kHash);
        cstart = reinterpret_cast<uint16_t*>(ccnt2 + kHash / 2);
        members = cstart + kHash + 2;
        kw = reinterpret_cast<uint32_t*>(members + kChunk);
        kfinal = kw + 6 * kChunkWords;
        state = reinterpret_cast<uint8_t*>(kfinal + kChunkWords);
    }
};

__device__ __forceinline__ float box_area(const float4 b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

// torchvision/csrc/ops/cpu/nms_kernel.cpp: suppress j when inter/(area_i+area_j-inter) > thr.
// The division is only evaluated when a 1e-6-wide guard band around the threshold cannot
// decide: fl(inter/den) > thr is implied by inter > fl(den*thr)*(1+1e-6) and excluded by
// inter < fl(den*thr)*(1-1e-6) (each product/quotient is within 2^-24 relative of exact).
// (Written without early returns: the callers run four of these per lane as independent chains, and a return per
// comparison turned them into divergent branches.  Only the rare undecided case branches, to the division.)
__device__ __forceinline__ bool iou_exceeds(const float4 a, const float aa, const float4 b,
                                            const float ab, const float thr) {
    const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    const float inter = __fmul_rn(w, h);
    const bool pos = inter > 0.f;      // otherwise 0/x is 0, -0 or NaN: never > thr (thr >= 0)
    const float den = __fsub_rn(__fadd_rn(aa, ab), inter);
    const float t = __fmul_rn(den, thr);
    const bool ok = t > 1e-30f && den < 1e30f;   // den > 0 and no under/overflow in the products
    const bool hi = inter > __fmul_rn(t, 1.000001f);
    const bool lo = inter < __fmul_rn(t, 0.999999f);
    bool r = pos && ok && hi;
    if (pos && !(ok && (hi || lo))) r = __fdiv_rn(inter, den) > thr;
    return r;
}

#ifdef VK_NMS_PROFILE
__device__ long long* g_nms_timing = nullptr;   // [batch][32] clock64 stamps (profiling builds only)
#define VK_STAMP(k) do { if (g_nms_timing && threadIdx.x == 0 && ((k) < 22 || (k) == 30)) g_nms_timing[(size_t)blockIdx.x * 32 + (k)] = clock64(); } while (0)
#else
#define VK_STAMP(k) do { } while (0)
#endif

// Everything a chunk needs: the sorted stage in shared memory and where the results go.
struct ChunkCtx {
    ScratchB XB;
    const unsigned long long* keys;   // [CAP] sorted keys of the stage
    const float4* sbox;               // [CAP] class-offset boxes of the stage
    const uint32_t* sidx;             // [CAP] row * nc + cls
    const uint16_t* scls;             // [CAP]
    const float4* boxes;              // the image's boxes (global)
    float* dets;
    int nc, agnostic, max_det, rank_base;
    bool cut, want_keep;
    float thr;
    uint32_t* done;                   // agnostic pruning: bitmap of the rows that are decided, or null
};

// One chunk of <= 256 sorted candidates [chunk0, chunk0 + cn) of the stage against the kept list.
// Returns the new kept count.  All T threads call it; ends with a block barrier.
template <int T>
__device__ __forceinline__ int nms_chunk(const ChunkCtx& C, int chunk0, int cn, int kept0, bool by_class, int mark = -1) {
    constexpr int NW = T / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float4* cbox = C.sbox + chunk0;
    const uint16_t* ccls = C.scls + chunk0;
    for (int i = tid; i < kHash / 2; i += T) C.XB.ccnt2[i] = 0u;
    for (int i = tid; i < kChunk; i += T) {
        int st0 = (i < cn) ? 0 : 2;                           // rows past the end never matter
        if (C.done != nullptr && i < cn) {                    // agnostic: the row was decided by an earlier chunk of this
            const uint32_t row = C.sidx[chunk0 + i] / (uint32_t)C.nc;      // stage, this candidate goes the same way
            if ((C.done[row >> 5] >> (row & 31)) & 1u) st0 = 2;
        }
        C.XB.state[i] = (uint8_t)st0;
#pragma unroll
        for (int wd = 0; wd < kChunkWords; ++wd) C.XB.pred[i * kChunkWords + wd] = 0u;
    }
    __syncthreads();
    if (by_class) {
        // group the chunk's rows by class bucket: counts -> starts -> member slots
        for (int i = tid; i < cn; i += T) {
            const uint32_t cls = ccls[i];
            atomicAdd(&C.XB.ccnt2[(cls & (kHash - 1)) >> 1], 1u << (16 * (cls & 1)));
        }
        __syncthreads();
        if (warp == 0) {
            int c[8], sum = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int bk = lane * 8 + k;
                c[k] = (int)((C.XB.ccnt2[bk >> 1] >> (16 * (bk & 1))) & 0xffffu);
                sum += c[k];
            }
            int base = warp_incl_scan(sum, lane) - sum;
#pragma unroll
            for (int k = 0; k < 8; ++k) { C.XB.cstart[lane * 8 + k] = (uint16_t)base; base += c[k]; }
            if (lane == 31) C.XB.cstart[kHash] = (uint16_t)base;
        }
        __syncthreads();
        for (int i = tid; i < cn; i += T) {
            const uint32_t bk = ccls[i] & (kHash - 1);
            const uint32_t old = atomicSub(&C.XB.ccnt2[bk >> 1], 1u << (16 * (bk & 1)));
            const uint32_t v = (old >> (16 * (bk & 1))) & 0xffffu;      // 1..count: a unique slot
            C.XB.members[C.XB.cstart[bk] + v - 1] = (uint16_t)i;
        }
        if (mark >= 0) VK_STAMP(mark);
        // 1: same-class kept boxes only (hash list of the kept set)
        for (int i = tid; i < cn; i += T) {
            const float4 cb = cbox[i];
            const float ca = box_area(cb);
            const uint32_t cc = ccls[i];
            bool sup = false;
            for (uint32_t k = (uint32_t)C.XB.khead[cc & (kHash - 1)]; k != kNil && !sup;) {
                const uint32_t meta = C.XB.kmeta[k];
                if ((meta >> 16) == cc) {
                    const float4 kb = C.XB.kbox[k];
                    sup = iou_exceeds(kb, box_area(kb), cb, ca, C.thr);
                }
                k = meta & 0xffffu;
            }
            if (sup) C.XB.state[i] = 2;
        }
        __syncthreads();
        if (mark >= 0) VK_STAMP(mark + 1);
        // 2: same-class predecessors inside the chunk: T / 256 threads per row, each walking its share of
        //    the members of the row's class bucket
        {
            constexpr int TPR = T / kChunk > 0 ? T / kChunk : 1;
            for (int i = tid % kChunk; i < cn; i += (T < kChunk ? T : kChunk)) {
                if (C.XB.state[i] != 0) continue;
                const uint32_t ic = ccls[i];
                const uint32_t bk = ic & (kHash - 1);
                const int e = C.XB.cstart[bk + 1];
                const float4 ib = cbox[i];
                const float ia = box_area(ib);
                for (int m = C.XB.cstart[bk] + tid / kChunk; m < e; m += 2 * TPR) {     // two independent chains per step
                    const int ja = C.XB.members[m];
                    const bool two = m + TPR < e;
                    const int jb_ = C.XB.members[two ? m + TPR : m];
                    const bool ta = ja < i && ccls[ja] == ic && C.XB.state[ja] == 0;
                    const bool tb = two && jb_ < i && ccls[jb_] == ic && C.XB.state[jb_] == 0;
                    const float4 ba = cbox[ja], bb = cbox[jb_];
                    const bool ha = ta && iou_exceeds(ba, box_area(ba), ib, ia, C.thr);
                    const bool hb = tb && iou_exceeds(bb, box_area(bb), ib, ia, C.thr);
                    if (ha) atomicOr(&C.XB.pred[i * kChunkWords + (ja >> 5)], 1u << (ja & 31));
                    if (hb) atomicOr(&C.XB.pred[i * kChunkWords + (jb_ >> 5)], 1u << (jb_ & 31));
                }
            }
        }
    } else {
        // 1: all kept boxes.  Warp = (32-row block, share of the kept list); the kept box
        //    is a shared-memory broadcast, every lane tests its own row against it.
        {
            constexpr int NP = NW / 8;                        // shares of the kept list per row block
            const int i = (warp % 8) * 32 + lane, part = warp / 8;
            if (i < cn) {
                const float4 cb = cbox[i];
                const float ca = box_area(cb);
                bool sup = false;
                for (int k = part; k < kept0 && !sup; k += 2 * NP) {       // two independent chains per step
                    const float4 ka = C.XB.kbox[k];
                    const bool second = k + NP < kept0;
                    const float4 kb = C.XB.kbox[second ? k + NP : k];
                    const bool sa = iou_exceeds(ka, box_area(ka), cb, ca, C.thr);
                    const bool sb = second && iou_exceeds(kb, box_area(kb), cb, ca, C.thr);
                    sup = sa || sb;
                }
                if (sup) C.XB.state[i] = 2;
            }
        }
        __syncthreads();
        // 2: warp task = (32-row block rb, 32-column word wd <= rb, half of the word): 72 half tiles.  (Whole tiles
        //    were 36 tasks on 32 warps: two rounds, the second one four tiles wide; halves take 2.25 rounds of half
        //    the length.  An agnostic eval image spends most of its time here: all pairs of 256 boxes per chunk.)
        uint16_t* const pred16 = reinterpret_cast<uint16_t*>(C.XB.pred);
        for (int t = warp; t < 72; t += NW) {
            const int half = t & 1;
            int rb = 0, wd = t >> 1;
            while (wd > rb) { wd -= rb + 1; ++rb; }
            const int i = rb * 32 + lane;
            const int j0 = wd * 32 + 16 * half;
            uint32_t m = 0;
            if (i < cn && C.XB.state[i] == 0) {
                const float4 ib = cbox[i];
                const float ia = box_area(ib);
                const int jn = min(16, i - j0);             // columns j < i only
                for (int bit = 0; bit < jn; bit += 4) {     // four independent chains per step
                    bool hit[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = j0 + bit + u;
                        hit[u] = false;
                        if (bit + u < jn && C.XB.state[j] == 0) {      // already removed: cannot suppress
                            const float4 jb = cbox[j];
                            hit[u] = iou_exceeds(jb, box_area(jb), ib, ia, C.thr);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (hit[u]) m |= 1u << (bit + u);
                }
            }
            pred16[(i * kChunkWords + wd) * 2 + half] = (uint16_t)m;      // (little-endian halves of the 32-bit word)
        }
    }
    __syncthreads();
    if (mark >= 0) VK_STAMP(mark + 2);
    // 3: fixed-point resolve; thread i < 256 owns row i.  Only those 8 warps iterate (named
    //    barrier 2, one per round; the word buffers ping-pong so a buffer is rewritten only
    //    after everyone has passed the next barrier).
    int st = 2;
    if (tid < kChunk) {
        uint32_t p[kChunkWords];
        st = C.XB.state[tid];
#pragma unroll
        for (int wd = 0; wd < kChunkWords; ++wd) p[wd] = C.XB.pred[tid * kChunkWords + wd];
        for (int round = 0;; ++round) {
            uint32_t* kwb = C.XB.kw + (round & 1) * 3 * kChunkWords;   // [kept | removed | undecided]
            const unsigned km = __ballot_sync(0xffffffffu, st == 1);
            const unsigned rm = __ballot_sync(0xffffffffu, st == 2);
            const unsigned um = __ballot_sync(0xffffffffu, st == 0);
            if (lane == 0) { kwb[warp] = km; kwb[kChunkWords + warp] = rm; kwb[2 * kChunkWords + warp] = um; }
            asm volatile("bar.sync 2, %0;" :: "n"(kChunk) : "memory");
            uint32_t hit = 0, pend = 0, und = 0;
#pragma unroll
            for (int wd = 0; wd < kChunkWords; ++wd) {
                const uint32_t kw = kwb[wd], rw = kwb[kChunkWords + wd];
                und |= kwb[2 * kChunkWords + wd];
                hit |= p[wd] & kw;
                pend |= p[wd] & ~(kw | rw);
            }
            if (und == 0) {                       // everyone sees the same words: uniform exit
                if (tid < kChunkWords) C.XB.kfinal[tid] = kwb[tid];
                break;
            }
            if (st == 0) {
                if (hit) st = 2;
                else if (!pend) st = 1;
            }
        }
    }
    __syncthreads();
    if (mark >= 0) VK_STAMP(mark + 3);
    // 4: kept boxes, in order, join the kept list and the output (image_proc.py:170-182)
    int total = 0;
#pragma unroll
    for (int wd = 0; wd < kChunkWords; ++wd) total += __popc(C.XB.kfinal[wd]);
    if (tid < kChunk && st == 1) {
        int rank = __popc(C.XB.kfinal[warp] & ((1u << lane) - 1u));
        for (int wd = 0; wd < warp; ++wd) rank += __popc(C.XB.kfinal[wd]);
        const int slot = kept0 + rank;
        if (slot < C.max_det) {
            const int p = chunk0 + tid;
            const uint32_t cls16 = ccls[tid];
            C.XB.kbox[slot] = cbox[tid];
            const uint32_t old = (uint32_t)atomicExch(&C.XB.khead[cls16 & (kHash - 1)], slot);
            C.XB.kmeta[slot] = (cls16 << 16) | (old & 0xffffu);
            const unsigned long long ck = C.keys[p];
            if (C.want_keep) C.XB.kkeep[slot] = C.cut ? (uint32_t)(C.rank_base + p) : ~(uint32_t)ck;
            const uint32_t idx = C.sidx[p];
            const float4 bx = C.boxes[idx / (uint32_t)C.nc];
            float* o = C.dets + (size_t)slot * 6;
            o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
            o[4] = __uint_as_float(unorder_key((uint32_t)(ck >> 32)));
            o[5] = (float)cls16;
        }
    }
    // agnostic pruning: every row of the chunk is decided now (kept, or removed by a kept box): later candidates of
    // these rows -- same box, IoU(self, self) = 1 > thr for a box of positive finite area -- are removed with them,
    // by the next chunks of this stage (above) and by the selection of the next stages
    if (C.done != nullptr && tid < cn) {
        const float a = box_area(cbox[tid]);
        if (a > 0.f && a < INFINITY) {
            const uint32_t row = C.sidx[chunk0 + tid] / (uint32_t)C.nc;
            atomicOr(&C.done[row >> 5], 1u << (row & 31));
        }
    }
    const int kept = min(C.max_det, kept0 + total);
    __syncthreads();
    if (mark >= 0) VK_STAMP(mark + 4);
    return kept;
}

template <int CAP>
static size_t nms_smem_bytes(int threads, int segs, int max_det, int prune_rows) {
    (void)threads;
    return (size_t)CAP * (8 + 8 + 16 + 4 + 2) + (size_t)(kHistBins + 4) * 4 +
           align16((size_t)(segs + 1) * 4) + (prune_rows ? align16((size_t)((prune_rows + 31) / 32) * 4) : 0) +
           align16(ScratchB::bytes(max_det));
}

template <int T, int CAP>
__global__ void __launch_bounds__(T, (T >= 1024 ? 1 : (T >= 512 ? 2 : 3)))
nms_kernel(const NmsArgs A) {
    constexpr int NW = T / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int wsum[33];
    __shared__ int s_bin, s_above, s_cnt, s_all;
    __shared__ unsigned long long s_max, s_min;

    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    float4* sbox = reinterpret_cast<float4*>(keys + CAP);
    unsigned long long* xchg = reinterpret_cast<unsigned long long*>(sbox + CAP);
    uint32_t* sidx = reinterpret_cast<uint32_t*>(xchg + CAP);
    int* hist = reinterpret_cast<int*>(sidx + CAP);
    uint16_t* scls = reinterpret_cast<uint16_t*>(hist + kHistBins + 4);    // (+4: the counting sort's end marker)
    int* segoff = reinterpret_cast<int*>(scls + CAP);     // [segs + 1] exclusive prefix of the segment counts
    uint32_t* done = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(segoff) + align16((size_t)(A.segs + 1) * 4));   // [rows / 32] (pruning)
    ScratchB XB(reinterpret_cast<unsigned char*>(done) + (A.prune ? align16((size_t)((A.rows + 31) / 32) * 4) : 0), A.max_det);

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint2* cand = reinterpret_cast<const uint2*>(A.cand + (size_t)b * A.cap);
    const float4* boxes = A.boxes + (size_t)b * A.rows;
    const int32_t* seg_count = A.seg_count + (size_t)b * A.segs;
    float* dets = A.dets + (size_t)b * A.max_det * 6;
    int64_t* keep_out = A.keep_idx ? A.keep_idx + (size_t)b * A.max_det : nullptr;
    VK_STAMP(0);

    const int n = A.counts[b];
    const int flags = A.flags[b];
    const int tile_cap = tile_slots_of(flags);
    // the image's list, built by the select pass: every candidate at or above `lbound` (and a few below it,
    // which are skipped); -1 = none
    int list_n = -1, list_len = 0;
    uint32_t lbound = 0;
    const uint64_t* list = A.list ? A.list + (size_t)b * A.list_cap : nullptr;
    if (list != nullptr && (flags & VK_FLAG_LIST)) {
        const int lc = A.list_count[b];
        if (lc <= A.list_cap) {
            lbound = (uint32_t)A.bound[b];
            list_len = lc;
            int mine = 0;
            for (int i = tid; i < lc; i += T) mine += (uint32_t)(list[i] >> 32) >= lbound ? 1 : 0;
            if (tid == 0) s_cnt = 0;
            __syncthreads();
            mine = warp_incl_scan(mine, lane);
            if (lane == 31 && mine) atomicAdd(&s_cnt, mine);
            __syncthreads();
            list_n = s_cnt;
            __syncthreads();
        }
    }
    const int K = min(n, A.max_nms);
    const bool cut = n > A.max_nms;

    for (int i = tid; i < kHash; i += T) XB.khead[i] = (int)kNil;
    // segment offsets (exclusive scan of the segment counts): needed to walk the candidate buffer itself
    // and to report torchvision's indices; an image served from its list never loads them
    bool have_segoff = false;
    auto load_segoff = [&]() {
        int carry = 0;
        for (int t0 = 0; t0 < A.segs; t0 += T) {
            const int t = t0 + tid;
            const int c = (t < A.segs) ? seg_count[t] : 0;
            int total;
            const int ex = block_excl_scan(c, wsum, &total);
            if (t < A.segs) segoff[t] = carry + ex;
            carry += total;
        }
        if (tid == 0) segoff[A.segs] = carry;
        __syncthreads();
        have_segoff = true;
    };
    bool use_list = list_n >= 0;
    if (!use_list) load_segoff();
    __syncthreads();

    // Agnostic mode: every class candidate of a prediction row has the row's box, so once one of them is
    // decided -- kept, or removed by a kept box X -- the rest are decided too: IoU(self, self) = 1 > thr removes
    // them in the first case, IoU(X, box) > thr in the second (torchvision visits them later, in score order).
    // Rows with a decided candidate and a box of positive finite area are marked here and their remaining
    // candidates skipped in the selection and compaction passes; multi-label eval has ~10 candidates per row.
    // (Needs thr < 1; ranks still count every candidate, so the max_nms cut stays exact.)
    bool prune_now = A.prune != 0;
    if (A.prune) {
        for (int i = tid; i < (A.rows + 31) / 32; i += T) done[i] = 0u;
    }
    auto pruned = [&](uint32_t idx) -> bool {
        if (!prune_now) return false;
        const uint32_t row = idx / (uint32_t)A.nc;
        return (done[row >> 5] >> (row & 31)) & 1u;
    };
    // every candidate of the current source: f(valid, key, id)
    const unsigned long long lfloor = (unsigned long long)lbound << 32;
    auto for_each_list = [&](auto&& f) {
        for (int i0 = 0; i0 < list_len; i0 += 4 * T) {
            unsigned long long e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * T + tid;
                e[u] = (i < list_len) ? list[i] : 0ull;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) f(i0 + u * T + tid < list_len && e[u] >= lfloor, e[u], ~(uint32_t)e[u]);
        }
    };
    auto for_each_segment = [&](auto&& f) {
        for (int t = warp; t < A.segs; t += NW) {
            const int cnt = segoff[t + 1] - segoff[t];
            if (cnt == 0) continue;
            const uint32_t slot0 = (uint32_t)t * (uint32_t)tile_cap;
            const uint2* cp = cand + slot0;
            for (int j0 = 0; j0 < cnt; j0 += 128) {        // 4 loads in flight per lane
                uint2 s[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 32 * u + lane;
                    s[u] = (j < cnt) ? cp[j] : make_uint2(0u, 0u);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 32 * u + lane;
                    f(j < cnt, ((unsigned long long)order_key(s[u].x) << 32) | (uint32_t)~s[u].y, s[u].y);
                }
            }
        }
    };
    auto for_each = [&](auto&& f) {
        if (use_list) for_each_list(f); else for_each_segment(f);
    };

    unsigned long long U = ~0ull;            // exclusive upper bound of the keys not yet processed
    int rank_base = 0;                       // candidates already processed (= those with key >= U)
    int kept0 = 0;
    bool safe = A.nc <= 65535;
    const float half_wh = 0.5f * A.max_wh;
    ChunkCtx CC{XB, keys, sbox, sidx, scls, boxes, dets, A.nc, A.agnostic, A.max_det, 0, cut, keep_out != nullptr, A.iou_thr,
                A.prune ? done : nullptr};
    VK_STAMP(1);
    int stamp = 2;
    while (rank_base < K && kept0 < A.max_det) {
        if (use_list && rank_base >= list_n) {            // the list is used up
            if (lbound == 0) break;                       // it held every candidate
            use_list = false;                             // U == lbound << 32: the rest comes from the segments
            if (!have_segoff) load_segoff();
        }
        const int remaining = (use_list ? list_n : n) - rank_base;
        unsigned long long v = use_list ? lfloor : 0ull;   // inclusive lower bound of this stage
        if (remaining > CAP) {
            const int tgt = min(CAP / 2, K - rank_base);
            unsigned long long prefix = 0, pmask = 0;
            int rem_t = tgt, above_total = 0;
            bool found = false;
#pragma unroll 1
            for (int pass = 0; pass < 6 && !found; ++pass) {
                const int shift = pass == 0 ? 53 : pass == 1 ? 42 : pass == 2 ? 32 : pass == 3 ? 21 : pass == 4 ? 10 : 0;
                const int nb = (pass == 2 || pass == 5) ? 1024 : 2048;
                for (int i = tid; i < kHistBins; i += T) hist[i] = 0;
                __syncthreads();
                for_each([&](bool ok, unsigned long long key, uint32_t idx) {
                    if (ok && key < U && (key & pmask) == prefix && !pruned(idx))
                        atomicAdd(&hist[(int)(key >> shift) & (nb - 1)], 1);
                });
                __syncthreads();
                constexpr int BPT = kHistBins / T;                       // bins per thread, from the top
                int vals[BPT], sum = 0;
#pragma unroll
                for (int k = 0; k < BPT; ++k) {
                    const int bin = nb - 1 - (BPT * tid + k);
                    vals[k] = (bin >= 0) ? hist[bin] : 0;
                    sum += vals[k];
                }
                int total;
                int run = block_excl_scan(sum, wsum, &total);
                if (pass == 0 && total <= CAP) {              // (pruning) everything still open fits one stage
                    found = true;
                    __syncthreads();
                    break;
                }
#pragma unroll
                for (int k = 0; k < BPT; ++k) {
                    if (run < rem_t && rem_t <= run + vals[k]) { s_bin = nb - 1 - (BPT * tid + k); s_above = run; }
                    run += vals[k];
                }
                __syncthreads();
                const int bin = s_bin, ab = s_above;
                prefix |= (unsigned long long)bin << shift;
                pmask |= (unsigned long long)(nb - 1) << shift;
                const int count_ge = above_total + ab + hist[bin];   // remaining candidates with key >= prefix
                if (count_ge <= CAP) { v = prefix; found = true; }
                above_total += ab;
                rem_t -= ab;
                __syncthreads();
            }
            // keys are unique, so the last pass always finds a bound (count_ge <= rem_t <= CAP / 2)
        }
        VK_STAMP(stamp);
        // ---- compaction of the stage [v, U) into shared memory
        if (tid == 0) { s_cnt = 0; s_all = 0; }
        __syncthreads();
        if (!use_list && rank_base == 0 && n <= CAP) {
            // small image (demo thresholds), its only stage: one thread per candidate, segment by binary search
            for (int p = tid; p < n; p += T) {
                int lo = 0, hi = A.segs;                       // last t with segoff[t] <= p
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (segoff[mid] <= p) lo = mid; else hi = mid;
                }
                const uint2 cd = cand[(size_t)lo * tile_cap + (p - segoff[lo])];
                keys[p] = ((unsigned long long)order_key(cd.x) << 32) | (uint32_t)~cd.y;
            }
            if (tid == 0) { s_cnt = n; s_all = n; }
        } else {
            for_each([&](bool ok, unsigned long long key, uint32_t idx) {
                const bool in_range = ok && key >= v && key < U;
                const bool take = in_range && !pruned(idx);
                const unsigned ma = __ballot_sync(0xffffffffu, in_range);
                const unsigned m = __ballot_sync(0xffffffffu, take);
                if (ma && lane == 0) atomicAdd(&s_all, __popc(ma));     // ranks count every candidate
                if (m) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_cnt, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (take) {
                        const int pos = base + __popc(m & ((1u << lane) - 1u));
                        if (pos < CAP) keys[pos] = key;
                    }
                }
            });
        }
        __syncthreads();
        VK_STAMP(stamp + 1);
        const int cnt = min(s_cnt, CAP);
        const int cnt_all = s_all;
        if (prune_now && rank_base + cnt_all > K) {           // the stage crosses the max_nms cut: redo it unpruned,
            prune_now = false;                                // the cut is defined on the ranks of ALL candidates
            __syncthreads();
            continue;
        }
        // ---- sort, descending: a counting sort on the leading 11 bits in which the stage's keys differ
        //      (bin = 2047 - ((key - min) >> shift), shift from the spread max - min), then every key ranks itself
        //      inside its bin by comparison.  Scores of a stage are spread over its bins, so bins hold a few keys;
        //      a block-wide bitonic or merge network costs 4x as much for 2048 keys.  (Thousands of equal
        //      scores in one stage make the in-bin step quadratic -- still correct.)
        if (cnt > 1) {
            constexpr int E = CAP / T > 0 ? CAP / T : 1;           // keys per thread
            unsigned long long kq[E];
            unsigned long long mx = 0ull, mn = ~0ull;
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const int e = tid + q * T;
                const bool in = e < cnt;
                kq[q] = in ? keys[e] : 0ull;
                mx = (in && kq[q] > mx) ? kq[q] : mx;
                mn = (in && kq[q] < mn) ? kq[q] : mn;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, mx, off);
                const unsigned long long u = __shfl_xor_sync(0xffffffffu, mn, off);
                mx = o > mx ? o : mx;
                mn = u < mn ? u : mn;
            }
            if (tid == 0) { s_max = 0ull; s_min = ~0ull; }
            for (int i = tid; i < kHistBins; i += T) hist[i] = 0;
            __syncthreads();
            if (lane == 0) { atomicMax(&s_max, mx); atomicMin(&s_min, mn); }
            __syncthreads();
            const unsigned long long base = s_min;
            const unsigned long long span = s_max - base;
            const int shift = (span >> 11) == 0ull ? 0 : (64 - __clzll((long long)span)) - 11;   // (span >> shift) < 2048
            int bin[E], ord[E];
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const int e = tid + q * T;
                bin[q] = (kHistBins - 1) - (int)((kq[q] - base) >> shift);
                ord[q] = (e < cnt) ? atomicAdd(&hist[bin[q]], 1) : 0;
            }
            __syncthreads();
            {   // exclusive scan of the bins (bins ascending = keys descending); hist[kHistBins] = cnt
                constexpr int BPT = kHistBins / T;
                int vals[BPT], sum = 0;
#pragma unroll
                for (int k = 0; k < BPT; ++k) { vals[k] = hist[BPT * tid + k]; sum += vals[k]; }
                int total;
                int run = block_excl_scan(sum, wsum, &total);
#pragma unroll
                for (int k = 0; k < BPT; ++k) { hist[BPT * tid + k] = run; run += vals[k]; }
                if (tid == 0) hist[kHistBins] = total;
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < E; ++q)
                if (tid + q * T < cnt) xchg[hist[bin[q]] + ord[q]] = kq[q];
            __syncthreads();
#pragma unroll
            for (int q = 0; q < E; ++q) {
                if (tid + q * T < cnt) {
                    const int s0 = hist[bin[q]], s1 = hist[bin[q] + 1];
                    int r = 0;
                    for (int p = s0; p < s1; ++p) r += xchg[p] > kq[q] ? 1 : 0;
                    keys[s0 + r] = kq[q];
                }
            }
            __syncthreads();
        }
        VK_STAMP(stamp + 2);
        // ---- the stage's boxes, once: candidate id -> (row, class) -> class-offset box
        const int M = min(cnt, K - rank_base);             // truncated at rank max_nms (image_proc.py:161-163)
        bool ok = true;
        for (int p = tid; p < M; p += T) {
            const uint32_t idx = ~(uint32_t)keys[p];
            const uint32_t row = idx / (uint32_t)A.nc;
            const uint32_t cls = idx - row * (uint32_t)A.nc;
            const float4 bx = boxes[row];
            ok = ok && fabsf(bx.x) <= half_wh && fabsf(bx.y) <= half_wh && fabsf(bx.z) <= half_wh && fabsf(bx.w) <= half_wh;
            const float off = A.agnostic ? 0.f : __fmul_rn((float)cls, A.max_wh);       // image_proc.py:166
            sbox[p] = make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));  // :168
            sidx[p] = idx;
            scls[p] = (uint16_t)cls;
        }
        safe = __syncthreads_and(ok) && safe;
        VK_STAMP(stamp + 3);
        // ---- NMS over the stage
        CC.rank_base = rank_base;
        for (int chunk0 = 0; chunk0 < M && kept0 < A.max_det; chunk0 += kChunk)
            kept0 = nms_chunk<T>(CC, chunk0, min(kChunk, M - chunk0), kept0, safe && !A.agnostic, chunk0 == 0 ? stamp + 4 : -1);
        VK_STAMP(stamp + 9);
        stamp += 10;
        if (prune_now) {                                      // rows decided in this stage
            for (int p = tid; p < M; p += T) {
                const float a = box_area(sbox[p]);
                if (a > 0.f && a < INFINITY) {
                    const uint32_t row = sidx[p] / (uint32_t)A.nc;
                    atomicOr(&done[row >> 5], 1u << (row & 31));
                }
            }
            __syncthreads();
        }
        rank_base += cnt_all;
        U = v;
        if (v == 0) break;                   // everything has been processed
    }
    // rows past the count are zero / -1 (the reference returns exactly k rows; the host slices)
    for (int i = kept0 * 6 + tid; i < A.max_det * 6; i += T) dets[i] = 0.f;
    if (keep_out) {
        if (!cut && kept0 > 0) {
            // torchvision's index = position in the reference's candidate list = number of candidates with a
            // smaller id.  One pass over all candidates: a candidate with id x lies below every kept id > x, so it
            // is tallied at b = #(kept ids <= x) and the ranks are the running sums of the tallies.
            if (!have_segoff) load_segoff();
            uint32_t* sk = reinterpret_cast<uint32_t*>(sbox);           // kept ids, ascending     (stage arrays are free now)
            int* tally = reinterpret_cast<int*>(sidx);                  // [kept0]
            for (int k = tid; k < kept0; k += T) {
                const uint32_t me = XB.kkeep[k];
                int r = 0;
                for (int j = 0; j < kept0; ++j) r += XB.kkeep[j] < me ? 1 : 0;
                sk[r] = me;
            }
            for (int k = tid; k < kept0; k += T) tally[k] = 0;
            __syncthreads();
            for_each_segment([&](bool ok, unsigned long long, uint32_t idx) {
                if (!ok) return;
                int lo = 0, hi = kept0;                                 // first j with sk[j] > idx
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (sk[mid] <= idx) lo = mid + 1; else hi = mid;
                }
                if (lo < kept0) atomicAdd(&tally[lo], 1);
            });
            __syncthreads();
            for (int k = tid; k < kept0; k += T) {
                const uint32_t me = XB.kkeep[k];
                int lo = 0, hi = kept0;                                 // position of me in sk
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (sk[mid] < me) lo = mid + 1; else hi = mid;
                }
                long long r = 0;
                for (int j = 0; j <= lo; ++j) r += tally[j];
                keep_out[k] = r;
            }
        } else {
            for (int k = tid; k < kept0; k += T) keep_out[k] = (int64_t)XB.kkeep[k];
        }
        for (int i = kept0 + tid; i < A.max_det; i += T) keep_out[i] = -1;
    }
    if (tid == 0) {
        A.det_counts[b] = kept0;
        if (A.status) A.status[b] = 0;
    }
    VK_STAMP(30);
#ifdef VK_NMS_PROFILE
    if (g_nms_timing && tid == 0) { g_nms_timing[(size_t)blockIdx.x * 32 + 31] = ((long long)n << 32) | (unsigned)rank_base; }
#endif
}

}  // namespace vk

using namespace vk;

#ifdef VK_NMS_PROFILE
// Profiling builds only (not in include/vk_b200.h): dev buffer of [batch][32] int64 receiving clock64()
// phase stamps of the next launches, or NULL.
extern "C" int vkdbg_nms_timing(void* dev_buf) {
    long long* p = static_cast<long long*>(dev_buf);
    cudaError_t e = cudaMemcpyToSymbol(g_nms_timing, &p, sizeof(p));
    return e == cudaSuccess ? VK_OK : (int)e;
}
#endif

template <int T, int CAP>
static int launch_nms(const NmsArgs& A, int batch, int segs, int max_det, cudaStream_t stream) {
    const size_t smem = nms_smem_bytes<CAP>(T, segs, max_det, A.prune ? A.rows : 0);
    if (smem > 220 * 1024)
        return fail_code(VK_E_LIMIT, "vk_nms_batched: max_det=%d segs=%d need %zu B of shared memory", max_det, segs, smem);
    const void* fn = reinterpret_cast<const void*>(&nms_kernel<T, CAP>);
    if (int rc = ensure_dyn_smem(fn, smem, "vk_nms_batched")) return rc;
    nms_kernel<T, CAP><<<batch, T, smem, stream>>>(A);
    count_launch();
    return check_launch("nms_kernel");
}

extern "C" int vk_nms_batched(const VkCandBuf* c, int batch, double iou_thres, int agnostic, int max_nms,
                              int max_det, float max_wh, float* dets, int32_t* det_counts, int64_t* keep_idx,
                              int32_t* status, vk_stream_t stream_) {
    if (batch == 0) return VK_OK;
    if (!c || !c->cand || !c->boxes || !c->ctrl || !c->seg_count || !dets || !det_counts || batch < 0)
        return fail_arg("vk_nms_batched: null/negative argument");
    if (max_nms < 1) return fail_arg("vk_nms_batched: max_nms %d < 1", max_nms);
    if (max_det < 1 || max_det > VK_MAX_DET) return fail_code(VK_E_LIMIT, "vk_nms_batched: max_det %d outside [1,%d]", max_det, VK_MAX_DET);
    if (c->segs < 1 || c->segs > VK_MAX_SEGMENTS) return fail_code(VK_E_LIMIT, "vk_nms_batched: %d segments outside [1,%d]", c->segs, VK_MAX_SEGMENTS);
    if (c->cap < 1 || c->rows < 1 || c->nc < 1) return fail_arg("vk_nms_batched: bad candidate buffer shape");
    if ((c->list != nullptr) != (c->list_cap > 0)) return fail_arg("vk_nms_batched: list and list_cap disagree");
    if (!(iou_thres >= 0.0 && iou_thres <= 1.0)) return fail_arg("vk_nms_batched: iou_thres %g outside [0,1]", iou_thres);
    cudaStream_t stream = as_stream(stream_);
    NmsArgs A;
    A.cand = c->cand; A.boxes = reinterpret_cast<const float4*>(c->boxes);
    A.counts = c->ctrl; A.flags = c->ctrl + (size_t)batch; A.list_count = c->ctrl + 2 * (size_t)batch; A.bound = c->ctrl + 3 * (size_t)batch;
    A.seg_count = c->seg_count; A.list = c->list;
    A.cap = c->cap; A.rows = c->rows; A.segs = c->segs; A.nc = c->nc; A.list_cap = c->list_cap;
    float thr = (float)iou_thres;                       // double compare == float compare against
    if ((double)thr > iou_thres) thr = nextafterf(thr, -INFINITY);  // the largest float <= threshold
    A.iou_thr = thr;
    A.agnostic = agnostic ? 1 : 0; A.max_nms = max_nms; A.max_det = max_det; A.max_wh = max_wh;
    A.prune = (agnostic && thr < 1.0f && keep_idx == nullptr) ? 1 : 0;
    A.dets = dets; A.det_counts = det_counts; A.keep_idx = keep_idx; A.status = status;
    if (c->list) {
        // selection pass for images with more candidates than their list holds (it returns at once for the
        // others): list entries and bound start from zero
        cudaError_t e = cudaMemsetAsync(A.list_count, 0, (size_t)batch * 2 * sizeof(int32_t), stream);
        if (e != cudaSuccess) return fail_code((int)e, "vk_nms_batched: memset: %s", cudaGetErrorString(e));
        // one wave of blocks, every SM streaming candidates whatever the batch
        const int per_sm = blocks_per_sm(reinterpret_cast<const void*>(&nms_select_kernel), kSelThreads, 0);
        int parts = (per_sm * kNumSMs) / batch;
        const int max_parts = ceil_div(c->segs, kSelThreads / 32);
        if (parts > max_parts) parts = max_parts;
        if (parts < 1) parts = 1;
        nms_select_kernel<<<dim3(parts, batch), kSelThreads, 0, stream>>>(A);
        count_launch();
        if (int rc = check_launch("nms_select_kernel")) return rc;
        return launch_nms<1024, 2048>(A, batch, c->segs, max_det, stream);
    }
    // list-less buffers (demo thresholds, a few hundred candidates per image): 512 threads.  Measured inside the
    // bench step, where this kernel runs beside the letterbox and the filter of the next batch: 115 us per step
    // with 512 threads, 118.5 with 1024 (shorter alone, but it then takes a whole SM's registers), 141 with 256.
#ifdef VK_NMS_TUNE
    if (const char* e = getenv("VK_NMS_T")) {              // tuning builds: CTA size of the list-less kernel
        const int t = atoi(e);
        if (t == 256) return launch_nms<256, 1024>(A, batch, c->segs, max_det, stream);
        if (t == 1024) return launch_nms<1024, 1024>(A, batch, c->segs, max_det, stream);
    }
#endif
    return launch_nms<512, 1024>(A, batch, c->segs, max_det, stream);
}
