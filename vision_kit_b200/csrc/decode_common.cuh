// Declarations shared by decode.cu (Detect decode, confidence filters) and convhead.cu (the fused
// Detect conv head): head geometry, filter arguments / candidate format, argument checks.
#pragma once
#include "vk_common.cuh"

namespace vk {

constexpr int kTileS = 64;        // spatial positions (= prediction rows) per tile
constexpr int kFiltPitch = kTileS + 4;   // filters: rows read with lanes over rows; 16-byte aligned for STS.128
constexpr int kDecThreads = 256;
constexpr int kWarps = kDecThreads / 32;

struct HeadDev {
    int variant, nl, na, nc, no, rows, tiles;
    int ny[VK_MAX_LEVELS], nx[VK_MAX_LEVELS], nynx[VK_MAX_LEVELS];
    int row_base[VK_MAX_LEVELS], tile_start[VK_MAX_LEVELS + 1], tpa[VK_MAX_LEVELS];
    float stride[VK_MAX_LEVELS];
    float anchors[VK_MAX_LEVELS][2 * VK_MAX_ANCHORS];
    const void* lv[VK_MAX_LEVELS];        // element type: the kernel's template parameter
    float* raw[VK_MAX_LEVELS];
};

struct FilterArgs {
    float conf;
    int multi_label;
    const uint32_t* class_mask;  // dev or null
    uint64_t* cand;
    float4* boxes;
    int32_t* counts;             // ctrl row 0: candidates per image (also the list reservation counter)
    int32_t* flags;              // ctrl row 1: VK_FLAG_* | (tile_cap / 64) << 8, written by the tile with seg == 0
    int32_t* seg_count;
    int cap, rows, segs, nc;
    int tile_cap;                // candidate slots each tile owns
};

// float order -> unsigned order (and back); scores are compared as these keys everywhere
__host__ __device__ __forceinline__ uint32_t order_key(uint32_t fbits) {
    return fbits ^ ((fbits >> 31) ? 0xffffffffu : 0x80000000u);
}
__host__ __device__ __forceinline__ uint32_t unorder_key(uint32_t k) {
    return k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu);
}
// histogram bin of a score (monotone: higher scores -> lower bins; everything >= 1.0 and
// anything that is not a positive float lands in bin 0, so "bin <= j" is always "key >= bound(j)")
__host__ __device__ __forceinline__ int hist_bin(uint32_t fbits) {
    return fbits >= 0x3f800000u ? 0 : (int)((0x3f800000u - fbits) >> 20);
}
// smallest ordered key whose bin is <= j
__host__ __device__ __forceinline__ uint32_t hist_bound(int j) {
    const uint32_t span = (uint32_t)(j + 1) << 20;
    return span >= 0x3f800000u ? 0u : order_key(0x3f800000u - span + 1u);
}

// what the tile with seg == 0 stores in the image's flags word: slots per tile (the consumer maps
// slot <-> (segment, position) with it)
__device__ __forceinline__ int cand_flags(const FilterArgs& A) { return (A.tile_cap >> 6) << 8; }

__device__ __forceinline__ bool class_allowed(const uint32_t* m, int c) {
    return m == nullptr || ((__ldg(m + (c >> 5)) >> (c & 31)) & 1u);
}

__device__ __forceinline__ float4 xyxy_from_cxcywh(float cx, float cy, float w, float h) {
    const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);  // utils/bboxes.py:103-111
    return make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
}

inline int make_head(const VkHeadCfg* cfg, HeadDev* H, const char* who) {
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

inline int check_cand(const VkCandBuf* o, int rows, int segs, int nc, int multi_label, const char* who) {
    if (!o || !o->cand || !o->boxes || !o->ctrl || !o->seg_count)
        return fail_arg("%s: candidate buffer has a NULL member", who);
    if (o->cap <= 0 || o->rows != rows || o->segs != segs || o->nc != nc)
        return fail_arg("%s: candidate buffer shape (cap=%d rows=%d segs=%d nc=%d) != (rows=%d segs=%d nc=%d)",
                        who, o->cap, o->rows, o->segs, o->nc, rows, segs, nc);
    if ((uint64_t)rows * (uint64_t)nc > 0x7fffffffull) return fail_code(VK_E_LIMIT, "%s: rows*nc overflows 31 bits", who);
    const long need = (long)segs * kTileS * ((multi_label && nc > 1) ? nc : 1);
    if (o->cap < need)
        return fail_arg("%s: cap %d < %ld (= segs * 64 * %s): every tile owns a fixed slot range", who, o->cap, need,
                        (multi_label && nc > 1) ? "nc" : "1");
    if (reinterpret_cast<uintptr_t>(o->boxes) & 15) return fail_arg("%s: boxes must be 16-byte aligned", who);
    if ((o->list != nullptr) != (o->list_cap > 0)) return fail_arg("%s: list and list_cap disagree", who);
    return VK_OK;
}

inline FilterArgs make_filter_args(const VkCandBuf* o, int batch, float conf, int multi_label, const uint32_t* mask) {
    FilterArgs A;
    A.conf = conf;
    A.multi_label = (multi_label && o->nc > 1) ? 1 : 0;   // image_proc.py:111
    A.class_mask = mask;
    A.cand = o->cand;
    A.boxes = reinterpret_cast<float4*>(o->boxes);
    A.counts = o->ctrl; A.flags = o->ctrl + (size_t)batch; A.seg_count = o->seg_count;
    A.cap = o->cap; A.rows = o->rows; A.segs = o->segs; A.nc = o->nc;
    A.tile_cap = kTileS * (A.multi_label ? o->nc : 1);
    return A;
}

// Zeroes the control words of the candidate buffer (counts, list entries, bound, flags).
inline int reset_cand(const VkCandBuf* o, int batch, const char* who, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(o->ctrl, 0, (size_t)VK_CTRL_WORDS * batch * sizeof(int32_t), stream);
    if (e != cudaSuccess) return fail_code((int)e, "%s: memset: %s", who, cudaGetErrorString(e));
    return VK_OK;
}

}  // namespace vk
