// Per-image top-max_nms cut, score sort and greedy class-aware NMS (SURVEY.md §8 a5 second
// half, a5', a6): utils/image_proc.py:154-182 / demo/processing.py:167-197 with
// torchvision.ops.nms's CPU arithmetic.  One CTA (1024 threads) per image; everything between
// the candidate list and the (max_det, 6) output stays in that CTA's shared memory:
//
//   A1  canonical order: exclusive scan of the segment table written by the filter kernels
//   A2  n > max_nms only: exact radix select (11/11/10 bits) of the max_nms-th score; ties at
//       the cut are resolved by canonical position (= the stable argsort the contract fixes)
//   A3  ordered compaction of the selected candidates into shared (score, pos) arrays
//   A4  bitonic sort on (score desc, pos asc); up to 1024 elements in registers + shuffles
//   A5  greedy NMS over the sorted list in chunks of 256:
//         1. each chunk box against the kept boxes of earlier chunks
//         2. predecessor bit matrix pred[i] = { j < i in the chunk : IoU(j, i) > thr }
//         3. fixed-point resolve: an undecided box with a kept predecessor is removed, one
//            whose predecessors are all removed is kept.  The lowest undecided box is always
//            decidable, so this ends with exactly the sequential greedy result after as many
//            rounds as the longest suppression chain -- not after 256 dependent steps.
//       Stops as soon as max_det boxes are kept, which is exact because greedy NMS visits
//       boxes in descending score order (image_proc.py:170 truncates afterwards).
//       Class-aware mode walks per-class hash lists (kept boxes and chunk boxes) so that only
//       same-class pairs are ever tested.  That is exact while every coordinate seen so far
//       lies within +-max_wh/2 (offset boxes of different classes are then disjoint); the
//       first box outside that range switches the image to testing all pairs.
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
constexpr int kHash = 256;          // class hash buckets
constexpr uint32_t kNil = 0xffffu;  // end of a hash list
constexpr size_t kSmemLimit = 232448 - 256;  // 227 KB per CTA minus this kernel's static part

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
    uint32_t* sel;  // [batch][P]: compacted position -> row*nc + cls
    int P;          // sort capacity, power of two
};

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

// Scratch after the sort arrays.  Phase A (A1-A3) and phase B (A5) alias; A4 uses `xchg`.
struct ScratchA {
    int* segoff;   // [segs + 1]
    int* segbase;  // [segs]
    int* hist;     // [kHistBins]        (A2)
    int* gt;       // [segs]             (A3, aliases hist)
    int* eq;       // [segs]
    __host__ __device__ static size_t bytes(int segs) {
        const size_t u = (size_t)kHistBins * 4 > (size_t)segs * 8 ? (size_t)kHistBins * 4 : (size_t)segs * 8;
        return align16((size_t)(2 * segs + 1) * 4) + u;
    }
    __device__ ScratchA(unsigned char* p, int segs) {
        segoff = reinterpret_cast<int*>(p);
        segbase = segoff + segs + 1;
        hist = reinterpret_cast<int*>(p + align16((size_t)(2 * segs + 1) * 4));
        gt = hist;
        eq = hist + segs;
    }
};
struct ScratchB {
    float4* kbox;      // [max_det] kept boxes (class-offset)
    float4* cbox;      // [kChunk]
    uint32_t* pred;    // [kChunk][kChunkWords]
    uint32_t* kmeta;   // [max_det] cls << 16 | next kept slot of the bucket
    int* khead;        // [kHash] newest kept slot per class bucket
    uint32_t* ccnt2;   // [kHash / 2] chunk members per class bucket, two 16-bit counts per word
    uint16_t* cstart;  // [kHash + 2] first member slot of each bucket (+ end)
    uint16_t* members; // [kChunk] chunk rows grouped by class bucket
    uint16_t* ccls;    // [kChunk]
    uint32_t* kw;      // [2][3][kChunkWords] ping-pong kept | removed | undecided bits
    uint32_t* kfinal;  // [kChunkWords] kept bits of the resolved chunk
    uint8_t* state;    // [kChunk] 0 undecided, 1 kept, 2 removed
    __host__ __device__ static size_t bytes(int max_det) {
        return (size_t)max_det * 16 + kChunk * 16 + kChunk * kChunkWords * 4 + (size_t)max_det * 4 +
               kHash * 4 + kHash * 2 + (kHash + 2) * 2 + kChunk * 2 * 2 + kChunkWords * 4 * 7 + kChunk;
    }
    __device__ ScratchB(unsigned char* p, int max_det) {
        kbox = reinterpret_cast<float4*>(p);
        cbox = kbox + max_det;
        pred = reinterpret_cast<uint32_t*>(cbox + kChunk);
        kmeta = pred + kChunk * kChunkWords;
        khead = reinterpret_cast<int*>(kmeta + max_det);
        ccnt2 = reinterpret_cast<uint32_t*>(khead + kHash);
        cstart = reinterpret_cast<uint16_t*>(ccnt2 + kHash / 2);
        members = cstart + kHash + 2;
        ccls = members + kChunk;
        kw = reinterpret_cast<uint32_t*>(ccls + kChunk);
        kfinal = kw + 6 * kChunkWords;
        state = reinterpret_cast<uint8_t*>(kfinal + kChunkWords);
    }
};

static size_t nms_smem_bytes(int P, int segs, int max_det) {
    size_t s = ScratchA::bytes(segs);
    if (ScratchB::bytes(max_det) > s) s = ScratchB::bytes(max_det);
    if ((size_t)kNmsThreads * 16 > s) s = (size_t)kNmsThreads * 16;   // A4 ping-pong exchange
    return (size_t)P * 6 + align16(s);
}

__device__ __forceinline__ uint32_t order_key(uint32_t fbits) {  // float order -> unsigned order
    return fbits ^ ((fbits >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ uint32_t unorder_key(uint32_t k) {
    return k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu);
}
__device__ __forceinline__ float box_area(const float4 b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

// torchvision/csrc/ops/cpu/nms_kernel.cpp: suppress j when inter/(area_i+area_j-inter) > thr.
// The division is only evaluated when a 1e-6-wide guard band around the threshold cannot
// decide: fl(inter/den) > thr is implied by inter > fl(den*thr)*(1+1e-6) and excluded by
// inter < fl(den*thr)*(1-1e-6) (each product/quotient is within 2^-24 relative of exact).
__device__ __forceinline__ bool iou_exceeds(const float4 a, const float aa, const float4 b,
                                            const float ab, const float thr) {
    const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    const float inter = __fmul_rn(w, h);
    if (!(inter > 0.f)) return false;  // 0/x is 0, -0 or NaN: never > thr (thr >= 0)
    const float den = __fsub_rn(__fadd_rn(aa, ab), inter);
    const float t = __fmul_rn(den, thr);
    if (t > 1e-30f && den < 1e30f) {   // den > 0 and no under/overflow in the products
        if (inter > __fmul_rn(t, 1.000001f)) return true;
        if (inter < __fmul_rn(t, 0.999999f)) return false;
    }
    return __fdiv_rn(inter, den) > thr;
}

// Everything A5 needs besides the sorted candidates themselves.
struct ChunkCtx {
    ScratchB XB;
    const float4* boxes;
    float* dets;
    int64_t* keep_out;
    int nc, agnostic, max_det;
    float max_wh, half_wh, thr;
};

// One chunk of <= 256 sorted candidates [chunk0, chunk0 + cn) of source S against the kept list.
// S.idx(p) = row*nc + cls, S.okey(p) = ordered score bits, S.keep(p) = index torchvision would
// report.  Returns the new kept count.  All 1024 threads call it; ends with a block barrier.
template <class Src>
__device__ __forceinline__ int nms_chunk(const ChunkCtx& C, const Src& S, int chunk0, int cn, int kept0, bool& safe) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bool ok = true;
    if (tid < kHash / 2) C.XB.ccnt2[tid] = 0u;
    if (tid < kChunk) {
        C.XB.state[tid] = 2;                               // rows past the end never matter
#pragma unroll
        for (int wd = 0; wd < kChunkWords; ++wd) C.XB.pred[tid * kChunkWords + wd] = 0u;
    }
    __syncthreads();
    if (tid < cn) {
        const uint32_t idx = S.idx(chunk0 + tid);
        const uint32_t row = idx / (uint32_t)C.nc;
        const uint32_t cls = idx - row * (uint32_t)C.nc;
        const float4 bx = C.boxes[row];
        ok = fabsf(bx.x) <= C.half_wh && fabsf(bx.y) <= C.half_wh && fabsf(bx.z) <= C.half_wh &&
             fabsf(bx.w) <= C.half_wh;
        const float off = C.agnostic ? 0.f : __fmul_rn((float)cls, C.max_wh);  // image_proc.py:166
        C.XB.cbox[tid] = make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off),
                                   __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));  // :168
        C.XB.ccls[tid] = (uint16_t)cls;
        C.XB.state[tid] = 0;
        atomicAdd(&C.XB.ccnt2[(cls & (kHash - 1)) >> 1], 1u << (16 * (cls & 1)));
    }
    safe = __syncthreads_and(ok) && safe;
    const bool by_class = safe && !C.agnostic;
    if (by_class) {
        // group the chunk's rows by class bucket: counts -> starts -> member slots
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
        if (tid < cn) {
            const uint32_t bk = C.XB.ccls[tid] & (kHash - 1);
            const uint32_t old = atomicSub(&C.XB.ccnt2[bk >> 1], 1u << (16 * (bk & 1)));
            const uint32_t v = (old >> (16 * (bk & 1))) & 0xffffu;      // 1..count: a unique slot
            C.XB.members[C.XB.cstart[bk] + v - 1] = (uint16_t)tid;
        }
        // (the barrier after step 1 below also orders these writes before step 2 reads them)
    }

    if (by_class) {
        // 1: same-class kept boxes only (hash list of the kept set)
        if (tid < cn) {
            const float4 cb = C.XB.cbox[tid];
            const float ca = box_area(cb);
            const uint32_t cc = C.XB.ccls[tid];
            bool sup = false;
            for (uint32_t k = (uint32_t)C.XB.khead[cc & (kHash - 1)]; k != kNil && !sup;) {
                const uint32_t meta = C.XB.kmeta[k];
                if ((meta >> 16) == cc) {
                    const float4 kb = C.XB.kbox[k];
                    sup = iou_exceeds(kb, box_area(kb), cb, ca, C.thr);
                }
                k = meta & 0xffffu;
            }
            if (sup) C.XB.state[tid] = 2;
        }
        __syncthreads();
        // 2: same-class predecessors inside the chunk.  One warp per row, lanes over the
        //    members of the row's class bucket (rows of a crowded class would otherwise walk
        //    a long list one dependent step at a time).
        for (int i = warp; i < cn; i += kNmsWarps) {
            if (C.XB.state[i] != 0) continue;
            const uint32_t ic = C.XB.ccls[i];
            const uint32_t bk = ic & (kHash - 1);
            const int e = C.XB.cstart[bk + 1];
            const float4 ib = C.XB.cbox[i];
            const float ia = box_area(ib);
            for (int m = C.XB.cstart[bk] + lane; m < e; m += 32) {
                const int j = C.XB.members[m];
                if (j >= i || C.XB.ccls[j] != ic || C.XB.state[j] != 0) continue;
                const float4 jb = C.XB.cbox[j];
                if (iou_exceeds(jb, box_area(jb), ib, ia, C.thr))
                    atomicOr(&C.XB.pred[i * kChunkWords + (j >> 5)], 1u << (j & 31));
            }
        }
    } else {
        // 1: all kept boxes.  Warp = (32-row block, quarter of the kept list); the kept box
        //    is a shared-memory broadcast, every lane tests its own row against it.
        {
            const int i = (warp >> 2) * 32 + lane, part = warp & 3;
            if (i < cn) {
                const float4 cb = C.XB.cbox[i];
                const float ca = box_area(cb);
                bool sup = false;
                for (int k = part; k < kept0 && !sup; k += 4) {
                    const float4 kb = C.XB.kbox[k];
                    sup = iou_exceeds(kb, box_area(kb), cb, ca, C.thr);
                }
                if (sup) C.XB.state[i] = 2;
            }
        }
        __syncthreads();
        // 2: warp tile = (32-row block rb, 32-column word wd <= rb): 36 tiles
        for (int t = warp; t < 36; t += kNmsWarps) {
            int rb = 0, wd = t;
            while (wd > rb) { wd -= rb + 1; ++rb; }
            const int i = rb * 32 + lane;
            uint32_t m = 0;
            if (i < cn && C.XB.state[i] == 0) {
                const float4 ib = C.XB.cbox[i];
                const float ia = box_area(ib);
                const int jn = min(32, i - wd * 32);       // columns j < i only
                for (int bit = 0; bit < jn; ++bit) {
                    const int j = wd * 32 + bit;
                    if (C.XB.state[j] != 0) continue;        // already removed: cannot suppress
                    const float4 jb = C.XB.cbox[j];
                    if (iou_exceeds(jb, box_area(jb), ib, ia, C.thr)) m |= 1u << bit;
                }
            }
            if (i < kChunk) C.XB.pred[i * kChunkWords + wd] = m;
        }
    }
    __syncthreads();
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
    // 4: kept boxes, in order, join the kept list and the output (image_proc.py:170-182)
    int total = 0;
#pragma unroll
    for (int wd = 0; wd < kChunkWords; ++wd) total += __popc(C.XB.kfinal[wd]);
    if (tid < kChunk && st == 1) {
        int rank = __popc(C.XB.kfinal[warp] & ((1u << lane) - 1u));
        for (int wd = 0; wd < warp; ++wd) rank += __popc(C.XB.kfinal[wd]);
        const int slot = kept0 + rank;
        if (slot < C.max_det) {
            const uint32_t cls16 = C.XB.ccls[tid];
            C.XB.kbox[slot] = C.XB.cbox[tid];
            const uint32_t old = (uint32_t)atomicExch(&C.XB.khead[cls16 & (kHash - 1)], slot);
            C.XB.kmeta[slot] = (cls16 << 16) | (old & 0xffffu);
            const uint32_t idx = S.idx(chunk0 + tid);
            const uint32_t row = idx / (uint32_t)C.nc;
            const float4 bx = C.boxes[row];
            float* o = C.dets + (size_t)slot * 6;
            o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
            o[4] = __uint_as_float(unorder_key(S.okey(chunk0 + tid)));
            o[5] = (float)(idx - row * (uint32_t)C.nc);
            if (C.keep_out) C.keep_out[slot] = S.keep(chunk0 + tid);
        }
    }
    const int kept = min(C.max_det, kept0 + total);
    __syncthreads();
    return kept;
}

struct SortedArrays {   // the big kernel's view of its sorted candidates
    const uint32_t* skey; const uint16_t* spos; const uint32_t* sel; bool cut;
    __device__ __forceinline__ uint32_t idx(int p) const { return sel[spos[p]]; }
    __device__ __forceinline__ uint32_t okey(int p) const { return skey[p]; }
    __device__ __forceinline__ int64_t keep(int p) const { return cut ? (int64_t)p : (int64_t)spos[p]; }
};

// Optional phase timestamps (clock64) for profiling: [batch][32] written by thread 0 when set.
__device__ long long* g_nms_timing = nullptr;
#define VK_STAMP(k) do { if (timing && tid == 0) timing[(size_t)blockIdx.x * 32 + (k)] = clock64(); } while (0)

__global__ void __launch_bounds__(kNmsThreads, 1)
nms_image_kernel(const NmsArgs A, const int32_t* __restrict__ only_flagged) {
    if (only_flagged && !only_flagged[blockIdx.x]) return;   // the staged kernel already did this image
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int wsum[33];
    __shared__ int s_found_bin, s_found_above;

    const int P = A.P;
    uint32_t* skey = reinterpret_cast<uint32_t*>(smem_raw);
    uint16_t* spos = reinterpret_cast<uint16_t*>(smem_raw + (size_t)P * 4);
    unsigned char* scratch = smem_raw + (size_t)P * 6;
    ScratchA XA(scratch, A.segs);
    ScratchB XB(scratch, A.max_det);
    unsigned long long* xchg = reinterpret_cast<unsigned long long*>(scratch);

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t* cand = A.cand + (size_t)b * A.cap;
    const float4* boxes = A.boxes + (size_t)b * A.rows;
    const int32_t* seg_base = A.seg_base + (size_t)b * A.segs;
    const int32_t* seg_count = A.seg_count + (size_t)b * A.segs;
    uint32_t* sel = A.sel + (size_t)b * P;
    float* dets = A.dets + (size_t)b * A.max_det * 6;
    int64_t* keep_out = A.keep_idx ? A.keep_idx + (size_t)b * A.max_det : nullptr;

    long long* timing = g_nms_timing;
    VK_STAMP(0);

    // ---------------- A1: canonical offsets of the segments
    int n = 0;
    {
        int carry = 0, clipped = 0;
        for (int t0 = 0; t0 < A.segs; t0 += kNmsThreads) {
            const int t = t0 + tid;
            int c = 0;
            if (t < A.segs) {
                c = seg_count[t];
                const int sb = seg_base[t];
                XA.segbase[t] = sb;
                if (c > 0 && sb + c > A.cap) { c = max(0, A.cap - sb); clipped = 1; }  // never happens with a
                                                                                        // buffer sized per the header
            }
            int total;
            const int ex = block_excl_scan(c, wsum, &total);
            if (t < A.segs) XA.segoff[t] = carry + ex;
            carry += total;
        }
        if (tid == 0) XA.segoff[A.segs] = carry;
        n = carry;
        clipped = __syncthreads_or(clipped);
        if (tid == 0 && A.status) A.status[b] = clipped ? 1 : 0;
    }

    VK_STAMP(1);
    // ---------------- A2: exact selection of the max_nms-th best score (image_proc.py:161-163)
    const int K = A.max_nms;
    const bool cut = n > K;
    uint32_t tval = 0;
    int need_eq = 0;
    if (cut) {
        uint32_t prefix = 0, pmask = 0;
        int remaining = K;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
            const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
            const int nb = pass == 2 ? 1024 : 2048;
            for (int i = tid; i < kHistBins; i += kNmsThreads) XA.hist[i] = 0;
            __syncthreads();
            for (int t = warp; t < A.segs; t += kNmsWarps) {       // candidates live in per-tile slot ranges
                const int cnt = XA.segoff[t + 1] - XA.segoff[t];
                const uint64_t* cp = cand + XA.segbase[t];
                for (int j = lane; j < cnt; j += 32) {
                    const uint32_t key = order_key((uint32_t)cp[j]);
                    if ((key & pmask) == prefix) atomicAdd(&XA.hist[(key >> shift) & (nb - 1)], 1);
                }
            }
            __syncthreads();
            // bins from the top: thread t owns bins nb-1-2t and nb-2-2t
            const int b0 = nb - 1 - 2 * tid, b1 = nb - 2 - 2 * tid;
            const int v0 = (b0 >= 0) ? XA.hist[b0] : 0;
            const int v1 = (b1 >= 0) ? XA.hist[b1] : 0;
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

    VK_STAMP(2);
    // ---------------- A3: ordered compaction into shared memory
    // sel[pos] = row*nc + cls of the candidate at compacted position pos
    const int M = cut ? K : n;
    if (!cut) {
        // every candidate is selected: one thread per canonical position, segment by binary search
        for (int p = tid; p < n; p += kNmsThreads) {
            int lo = 0, hi = A.segs;            // last t with segoff[t] <= p
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (XA.segoff[mid] <= p) lo = mid; else hi = mid;
            }
            const uint64_t cd = cand[XA.segbase[lo] + (p - XA.segoff[lo])];
            skey[p] = order_key((uint32_t)cd);
            spos[p] = (uint16_t)p;
            sel[p] = (uint32_t)(cd >> 32);
        }
    } else {
        for (int t = warp; t < A.segs; t += kNmsWarps) {
            const int cnt = XA.segoff[t + 1] - XA.segoff[t];
            const int base = XA.segbase[t];
            int gt = 0, eq = 0;
            for (int j0 = 0; j0 < cnt; j0 += 32) {
                const int j = j0 + lane;
                uint32_t key = 0;
                const bool ok = j < cnt;
                if (ok) key = order_key((uint32_t)cand[base + j]);
                gt += __popc(__ballot_sync(0xffffffffu, ok && key > tval));
                eq += __popc(__ballot_sync(0xffffffffu, ok && key == tval));
            }
            if (lane == 0) { XA.gt[t] = gt; XA.eq[t] = eq; }
        }
        __syncthreads();
        int carry_gt = 0, carry_eq = 0;
        for (int t0 = 0; t0 < A.segs; t0 += kNmsThreads) {
            const int t = t0 + tid;
            const int g = (t < A.segs) ? XA.gt[t] : 0;
            const int e = (t < A.segs) ? XA.eq[t] : 0;
            int tg, te;
            const int xg = block_excl_scan(g, wsum, &tg);
            const int xe = block_excl_scan(e, wsum, &te);
            if (t < A.segs) { XA.gt[t] = carry_gt + xg; XA.eq[t] = carry_eq + xe; }
            carry_gt += tg; carry_eq += te;
        }
        __syncthreads();
        for (int t = warp; t < A.segs; t += kNmsWarps) {
            const int cnt = XA.segoff[t + 1] - XA.segoff[t];
            if (cnt == 0) continue;
            const int base = XA.segbase[t];
            int eq_before = XA.eq[t];
            int pos0 = XA.gt[t] + min(eq_before, need_eq);
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

    VK_STAMP(3);
    // ---------------- A4: bitonic sort, "before" = higher score, then lower position
    if (M > 1 && Ps <= kNmsThreads) {
        // one element per thread in a register: composite = score<<16 | (0xffff - pos), sorted
        // descending; strides below 32 exchange by shuffle, the rest through shared memory.
        // Only the Ps/32 warps that hold elements take part (named barrier 1).
        if (tid < Ps) {
            unsigned long long v = ((unsigned long long)skey[tid] << 16) | (unsigned long long)(0xffffu - spos[tid]);
            int pp = 0;   // ping-pong: a buffer is rewritten only two barriers after it was read
            for (int k = 2; k <= Ps; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    unsigned long long o;
                    if (j >= 32) {
                        unsigned long long* xb = xchg + pp * kNmsThreads;
                        xb[tid] = v;
                        asm volatile("bar.sync 1, %0;" :: "r"(Ps) : "memory");
                        o = xb[tid ^ j];
                        pp ^= 1;
                    } else {
                        o = __shfl_xor_sync(0xffffffffu, v, j);
                    }
                    const bool keep_max = ((tid & j) == 0) == ((tid & k) == 0);
                    v = keep_max ? (v > o ? v : o) : (v < o ? v : o);
                }
            }
            skey[tid] = (uint32_t)(v >> 16);
            spos[tid] = (uint16_t)(0xffffu - (uint32_t)(v & 0xffffull));
        }
        __syncthreads();
    } else if (M > 1) {
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

    VK_STAMP(4);
    // ---------------- A5: greedy NMS with a kept list, early exit at max_det
    for (int i = tid; i < kHash; i += kNmsThreads) XB.khead[i] = (int)kNil;
    int kept0 = 0;
    bool safe = A.nc <= 65535;
    const ChunkCtx CC{XB, boxes, dets, keep_out, A.nc, A.agnostic, A.max_det, A.max_wh, 0.5f * A.max_wh, A.iou_thr};
    const SortedArrays SRC{skey, spos, sel, cut};
    __syncthreads();
    for (int chunk0 = 0; chunk0 < M && kept0 < A.max_det; chunk0 += kChunk) {
        const int cn = min(kChunk, M - chunk0);
        kept0 = nms_chunk(CC, SRC, chunk0, cn, kept0, safe);
        if (chunk0 == 0) VK_STAMP(8);
    }
    VK_STAMP(9);
    // rows past the count are zero / -1 (the reference returns exactly k rows; the host slices)
    for (int i = kept0 * 6 + tid; i < A.max_det * 6; i += kNmsThreads) dets[i] = 0.f;
    if (keep_out)
        for (int i = kept0 + tid; i < A.max_det; i += kNmsThreads) keep_out[i] = -1;
    if (tid == 0) A.det_counts[b] = kept0;
    VK_STAMP(10);
    if (timing && tid == 0) { timing[(size_t)blockIdx.x * 32 + 11] = n; timing[(size_t)blockIdx.x * 32 + 12] = kept0; }
}

// ---------------------------------------------------------------------------------------
// Staged kernel (the default): greedy NMS consumes candidates in descending score order and
// stops at max_det, so the sorted order is only ever needed for a prefix.  Candidates are
// therefore processed in stages of ~1024: a 2-3 pass radix histogram finds a score bound that
// delimits the next stage (whole tie groups, at most kStageCap candidates), one streaming
// pass compacts the stage into shared memory as 64-bit keys (ordered score << 32 | ~canonical
// position), a bitonic sort orders it, and the chunk loop above runs on it.  At eval settings
// (200 k candidates per image, cut at 30 000) one or two stages reach max_det; the 32 768-
// element sort of the one-shot kernel never happens.  The exact cut at max_nms falls out of
// the order: the stage that crosses rank max_nms is truncated there, ties already sorted by
// canonical position.  A tie group larger than kStageCap (thousands of bit-identical scores)
// is the one case this kernel hands to nms_image_kernel through the need_big flag.
// ---------------------------------------------------------------------------------------
constexpr int kStageCap = 2048;
constexpr int kStageTarget = 1024;

struct StagedSrc {
    const unsigned long long* keys;
    const int* segoff;
    const int* segbase;
    const uint64_t* cand;
    int segs, rank_base;
    bool cut;
    __device__ __forceinline__ uint32_t canon(int p) const { return ~(uint32_t)keys[p]; }
    __device__ __forceinline__ uint32_t idx(int p) const {
        const int cp = (int)canon(p);
        int lo = 0, hi = segs;                 // last t with segoff[t] <= cp
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (segoff[mid] <= cp) lo = mid; else hi = mid;
        }
        return (uint32_t)(cand[segbase[lo] + (cp - segoff[lo])] >> 32);
    }
    __device__ __forceinline__ uint32_t okey(int p) const { return (uint32_t)(keys[p] >> 32); }
    __device__ __forceinline__ int64_t keep(int p) const { return cut ? (int64_t)(rank_base + p) : (int64_t)canon(p); }
};

static size_t staged_smem_bytes(int segs, int max_det) {
    return (size_t)kStageCap * 8 + (size_t)kNmsThreads * 16 + align16((size_t)(2 * segs + 2) * 4) +
           (size_t)kHistBins * 4 + align16(ScratchB::bytes(max_det));
}

__global__ void __launch_bounds__(kNmsThreads, 1)
nms_staged_kernel(const NmsArgs A, int32_t* __restrict__ need_big) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int wsum[33];
    __shared__ int s_bin, s_above, s_cnt;

    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    unsigned long long* xchg = keys + kStageCap;
    int* segoff = reinterpret_cast<int*>(xchg + 2 * kNmsThreads);
    int* segbase = segoff + A.segs + 1;
    int* hist = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(segoff) + align16((size_t)(2 * A.segs + 2) * 4));
    ScratchB XB(reinterpret_cast<unsigned char*>(hist + kHistBins), A.max_det);

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t* cand = A.cand + (size_t)b * A.cap;
    const float4* boxes = A.boxes + (size_t)b * A.rows;
    const int32_t* seg_base = A.seg_base + (size_t)b * A.segs;
    const int32_t* seg_count = A.seg_count + (size_t)b * A.segs;
    float* dets = A.dets + (size_t)b * A.max_det * 6;
    int64_t* keep_out = A.keep_idx ? A.keep_idx + (size_t)b * A.max_det : nullptr;
    if (tid == 0) need_big[b] = 0;

    // ---- canonical offsets of the segments (as A1 of the one-shot kernel)
    int n = 0;
    {
        int carry = 0, clipped = 0;
        for (int t0 = 0; t0 < A.segs; t0 += kNmsThreads) {
            const int t = t0 + tid;
            int c = 0;
            if (t < A.segs) {
                c = seg_count[t];
                const int sb = seg_base[t];
                segbase[t] = sb;
                if (c > 0 && sb + c > A.cap) { c = max(0, A.cap - sb); clipped = 1; }
            }
            int total;
            const int ex = block_excl_scan(c, wsum, &total);
            if (t < A.segs) segoff[t] = carry + ex;
            carry += total;
        }
        if (tid == 0) segoff[A.segs] = carry;
        n = carry;
        clipped = __syncthreads_or(clipped);
        if (tid == 0 && A.status) A.status[b] = clipped ? 1 : 0;
    }
    const int K = min(n, A.max_nms);
    const bool cut = n > A.max_nms;

    for (int i = tid; i < kHash; i += kNmsThreads) XB.khead[i] = (int)kNil;
    int kept0 = 0;
    bool safe = A.nc <= 65535;
    const ChunkCtx CC{XB, boxes, dets, keep_out, A.nc, A.agnostic, A.max_det, A.max_wh, 0.5f * A.max_wh, A.iou_thr};
    __syncthreads();

    // streams every candidate of the image: f(ordered key, canonical position)
    auto for_each_candidate = [&](auto&& f) {
        for (int t = warp; t < A.segs; t += kNmsWarps) {
            const int cnt = segoff[t + 1] - segoff[t];
            if (cnt == 0) continue;
            const uint64_t* cp = cand + segbase[t];
            const int c0 = segoff[t];
            for (int j0 = 0; j0 < cnt; j0 += 128) {        // 4 loads in flight per lane
                uint64_t v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 32 * u + lane;
                    v[u] = (j < cnt) ? cp[j] : 0ull;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 32 * u + lane;
                    f(j < cnt, order_key((uint32_t)v[u]), (uint32_t)(c0 + j));
                }
            }
        }
    };

    unsigned long long U = 1ull << 32;       // exclusive upper bound of the scores not yet processed
    int rank_base = 0;                       // candidates already processed (= those with key >= U)
    while (rank_base < K && kept0 < A.max_det) {
        const int remaining = n - rank_base;
        uint32_t v = 0;                      // inclusive lower bound of this stage
        if (remaining > kStageCap) {
            const int tgt = min(kStageTarget, K - rank_base);
            uint32_t prefix = 0, pmask = 0;
            int rem_t = tgt, above_total = 0;
            bool found = false;
#pragma unroll 1
            for (int pass = 0; pass < 3 && !found; ++pass) {
                const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
                const int nb = pass == 2 ? 1024 : 2048;
                for (int i = tid; i < kHistBins; i += kNmsThreads) hist[i] = 0;
                __syncthreads();
                for_each_candidate([&](bool ok, uint32_t key, uint32_t) {
                    if (ok && (unsigned long long)key < U && (key & pmask) == prefix)
                        atomicAdd(&hist[(key >> shift) & (nb - 1)], 1);
                });
                __syncthreads();
                const int b0 = nb - 1 - 2 * tid, b1 = nb - 2 - 2 * tid;   // bins from the top
                const int v0 = (b0 >= 0) ? hist[b0] : 0;
                const int v1 = (b1 >= 0) ? hist[b1] : 0;
                int total;
                const int above = block_excl_scan(v0 + v1, wsum, &total);
                if (above < rem_t && rem_t <= above + v0) { s_bin = b0; s_above = above; }
                else if (above + v0 < rem_t && rem_t <= above + v0 + v1) { s_bin = b1; s_above = above + v0; }
                __syncthreads();
                const int bin = s_bin, ab = s_above;
                prefix |= (uint32_t)bin << shift;
                pmask |= (uint32_t)(nb - 1) << shift;
                const int count_ge = above_total + ab + hist[bin];   // remaining candidates with key >= prefix
                if (count_ge <= kStageCap) { v = prefix; found = true; }
                above_total += ab;
                rem_t -= ab;
                __syncthreads();
            }
            if (!found) {                    // > kStageCap bit-identical scores: one-shot kernel takes over
                if (tid == 0) need_big[b] = 1;
                return;
            }
        }
        // ---- compaction of the stage [v, U) into shared memory
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        if (remaining == n && n <= kNmsThreads) {
            // small image, first (only) stage: one thread per canonical position
            if (tid < n) {
                int lo = 0, hi = A.segs;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (segoff[mid] <= tid) lo = mid; else hi = mid;
                }
                const uint64_t cd = cand[segbase[lo] + (tid - segoff[lo])];
                keys[tid] = ((unsigned long long)order_key((uint32_t)cd) << 32) | (unsigned long long)(~(uint32_t)tid);
            }
            if (tid == 0) s_cnt = n;
        } else {
            for_each_candidate([&](bool ok, uint32_t key, uint32_t canon) {
                const bool take = ok && key >= v && (unsigned long long)key < U;
                const unsigned m = __ballot_sync(0xffffffffu, take);
                if (m) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_cnt, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (take) {
                        const int pos = base + __popc(m & ((1u << lane) - 1u));
                        if (pos < kStageCap) keys[pos] = ((unsigned long long)key << 32) | (unsigned long long)(~canon);
                    }
                }
            });
        }
        __syncthreads();
        const int cnt = min(s_cnt, kStageCap);
        int Ps = 32;
        while (Ps < cnt) Ps <<= 1;
        for (int i = cnt + tid; i < Ps; i += kNmsThreads) keys[i] = 0ull;
        __syncthreads();
        // ---- sort, descending
        if (cnt > 1 && Ps <= kNmsThreads) {
            if (tid < Ps) {
                unsigned long long x = keys[tid];
                int pp = 0;
                for (int k = 2; k <= Ps; k <<= 1) {
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        unsigned long long o;
                        if (j >= 32) {
                            unsigned long long* xb = xchg + pp * kNmsThreads;
                            xb[tid] = x;
                            asm volatile("bar.sync 1, %0;" :: "r"(Ps) : "memory");
                            o = xb[tid ^ j];
                            pp ^= 1;
                        } else {
                            o = __shfl_xor_sync(0xffffffffu, x, j);
                        }
                        const bool keep_max = ((tid & j) == 0) == ((tid & k) == 0);
                        x = keep_max ? (x > o ? x : o) : (x < o ? x : o);
                    }
                }
                keys[tid] = x;
            }
            __syncthreads();
        } else if (cnt > 1) {
            for (int k = 2; k <= Ps; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < (Ps >> 1); i += kNmsThreads) {
                        const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                        const int hi = lo | j;
                        const unsigned long long ka = keys[lo], kb = keys[hi];
                        if ((kb > ka) == ((lo & k) == 0)) { keys[lo] = kb; keys[hi] = ka; }
                    }
                    __syncthreads();
                }
            }
        }
        // ---- NMS over the stage, truncated at rank max_nms
        const int M = min(cnt, K - rank_base);
        const StagedSrc SRC{keys, segoff, segbase, cand, A.segs, rank_base, cut};
        for (int chunk0 = 0; chunk0 < M && kept0 < A.max_det; chunk0 += kChunk)
            kept0 = nms_chunk(CC, SRC, chunk0, min(kChunk, M - chunk0), kept0, safe);
        rank_base += cnt;
        U = v;
        if (v == 0) break;                   // everything has been processed
    }
    // rows past the count are zero / -1 (the reference returns exactly k rows; the host slices)
    for (int i = kept0 * 6 + tid; i < A.max_det * 6; i += kNmsThreads) dets[i] = 0.f;
    if (keep_out)
        for (int i = kept0 + tid; i < A.max_det; i += kNmsThreads) keep_out[i] = -1;
    if (tid == 0) A.det_counts[b] = kept0;
}

static int next_pow2(int v) {
    int p = 32;
    while (p < v) p <<= 1;
    return p;
}

static int sort_capacity(int max_nms, int cap) { return next_pow2(max_nms < cap ? max_nms : cap); }

}  // namespace vk

using namespace vk;

// Profiling hook (not part of the product ABI in include/vk_b200.h): dev buffer of
// [batch][32] int64 receiving clock64() phase stamps of the next launches, or NULL.
extern "C" int vkdbg_nms_timing(void* dev_buf) {
    long long* p = static_cast<long long*>(dev_buf);
    cudaError_t e = cudaMemcpyToSymbol(g_nms_timing, &p, sizeof(p));
    return e == cudaSuccess ? VK_OK : (int)e;
}

extern "C" size_t vk_nms_workspace_bytes(int batch, int max_nms) {
    if (batch <= 0 || max_nms <= 0 || max_nms > VK_MAX_NMS) return 0;
    return (size_t)batch * next_pow2(max_nms) * sizeof(uint32_t) + (size_t)batch * sizeof(int32_t);
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
    const size_t ws_need = (size_t)batch * P * sizeof(uint32_t) + (size_t)batch * sizeof(int32_t);
    if (!ws || ws_bytes < ws_need)
        return fail_code(VK_E_WORKSPACE, "vk_nms_batched: workspace %zu < %zu", ws_bytes, ws_need);
    const size_t smem = nms_smem_bytes(P, c->segs, max_det);
    if (smem > kSmemLimit)
        return fail_code(VK_E_LIMIT, "vk_nms_batched: max_nms=%d max_det=%d segs=%d need %zu B of shared memory (> %zu)",
                         max_nms, max_det, c->segs, smem, kSmemLimit);
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
    int32_t* need_big = reinterpret_cast<int32_t*>(static_cast<char*>(ws) + (size_t)batch * P * sizeof(uint32_t));
    const size_t smem_st = staged_smem_bytes(c->segs, max_det);
    cudaError_t e = cudaFuncSetAttribute(nms_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_st);
    if (e != cudaSuccess) return fail_code((int)e, "vk_nms_batched: %zu B of shared memory: %s", smem_st, cudaGetErrorString(e));
    nms_staged_kernel<<<batch, kNmsThreads, smem_st, as_stream(stream)>>>(A, need_big);
    count_launch();
    if (int rc = check_launch("nms_staged_kernel")) return rc;
    // images with a tie group too large for a stage (need_big) are redone by the one-shot kernel
    e = cudaFuncSetAttribute(nms_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail_code((int)e, "vk_nms_batched: %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    nms_image_kernel<<<batch, kNmsThreads, smem, as_stream(stream)>>>(A, need_big);
    count_launch();
    return check_launch("nms_image_kernel");
}
