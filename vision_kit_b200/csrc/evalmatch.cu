// Evaluator matching on the device: the per-image body of
// DetEvaluator.evaluate (core/eval/det_evaluator.py:141-178) for a whole batch in one launch --
// un-letterbox + clip of detections and labels (utils/image_proc.py:63-80), torchvision box_iou
// (labels x detections), and the greedy unique matching of process_batch (:274-300) at every IoU
// threshold.  The reference does this with a D2H copy and a NumPy sort + two np.unique calls per
// threshold per image.
//
// One block per image.  Matching without the sort: a detection's surviving match is its best
// class-matching label (the same label at every threshold its IoU passes), and a label keeps the
// LOWEST detection index among the detections whose best label it is (atomicMin in shared memory).
// IoU ties between two labels of one detection: the reference's order comes from an unstable
// argsort (unspecified); here the lower label index wins (oracle/restate.py::process_batch).
#include "vk_common.cuh"

namespace vk {

constexpr int kEvalThreads = 256;

struct ScaleGeom {
    float pad_w, pad_h, gain, clip_w, clip_h;
};

// utils/image_proc.py:67-71: python floats (double), entering float32 tensor arithmetic as float32
__device__ __forceinline__ ScaleGeom scale_geom(int h1, int w1, int h0, int w0) {
    const double gh = (double)h1 / (double)h0, gw = (double)w1 / (double)w0;
    const double gain = gh < gw ? gh : gw;
    ScaleGeom g;
    g.pad_w = (float)(((double)w1 - (double)w0 * gain) / 2.0);
    g.pad_h = (float)(((double)h1 - (double)h0 * gain) / 2.0);
    g.gain = (float)gain;
    g.clip_w = (float)w0;
    g.clip_h = (float)h0;
    return g;
}

// :76-79 + utils/bboxes.py:50-59, one rounding per operation
__device__ __forceinline__ float4 unletterbox(float4 b, const ScaleGeom& g) {
    b.x = fminf(fmaxf(__fdiv_rn(__fsub_rn(b.x, g.pad_w), g.gain), 0.f), g.clip_w);
    b.y = fminf(fmaxf(__fdiv_rn(__fsub_rn(b.y, g.pad_h), g.gain), 0.f), g.clip_h);
    b.z = fminf(fmaxf(__fdiv_rn(__fsub_rn(b.z, g.pad_w), g.gain), 0.f), g.clip_w);
    b.w = fminf(fmaxf(__fdiv_rn(__fsub_rn(b.w, g.pad_h), g.gain), 0.f), g.clip_h);
    return b;
}

__global__ void __launch_bounds__(kEvalThreads)
eval_match_kernel(const float* __restrict__ dets, const int32_t* __restrict__ det_counts, int max_det,
                  const float* __restrict__ labels, const int32_t* __restrict__ label_offsets,
                  const int32_t* __restrict__ img0_hw, int img1_h, int img1_w,
                  const float* __restrict__ iouv, int niou, int max_labels, int prescaled,
                  float* __restrict__ predn, float* __restrict__ labeln, uint8_t* __restrict__ correct) {
    extern __shared__ __align__(16) unsigned char eval_smem[];
    float* s_lab = reinterpret_cast<float*>(eval_smem);                          // [max_labels][5] cls, x1, y1, x2, y2
    float* s_area = s_lab + (size_t)max_labels * 5;                              // [max_labels]
    int* s_win = reinterpret_cast<int*>(s_area + max_labels);                    // [niou][max_labels] winning detection
    __shared__ float s_thr[32];
    const int b = blockIdx.x;
    const int n = min(det_counts[b], max_det);
    const int l0 = label_offsets[b], m = min(label_offsets[b + 1] - l0, max_labels);
    ScaleGeom g{};
    if (!prescaled) g = scale_geom(img1_h, img1_w, img0_hw[2 * b], img0_hw[2 * b + 1]);
    if (threadIdx.x < niou) s_thr[threadIdx.x] = iouv[threadIdx.x];
    // labels: cxcywh -> xyxy (utils/bboxes.py:103-111), un-letterbox, clip (:165-168)
    for (int l = threadIdx.x; l < m; l += kEvalThreads) {
        const float* p = labels + (size_t)(l0 + l) * 6;                          // image, cls, cx, cy, w, h
        float4 bx;
        if (prescaled) {                                                         // process_batch: x1, y1, x2, y2 as given
            bx = make_float4(p[2], p[3], p[4], p[5]);
        } else {
            const float hw = __fmul_rn(p[4], 0.5f), hh = __fmul_rn(p[5], 0.5f);
            bx = unletterbox(make_float4(__fsub_rn(p[2], hw), __fsub_rn(p[3], hh),
                                         __fadd_rn(p[2], hw), __fadd_rn(p[3], hh)), g);
        }
        float* s = s_lab + 5 * l;
        s[0] = p[1]; s[1] = bx.x; s[2] = bx.y; s[3] = bx.z; s[4] = bx.w;
        s_area[l] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
        if (labeln) {
            float* o = labeln + (size_t)(l0 + l) * 5;
            o[0] = p[1]; o[1] = bx.x; o[2] = bx.y; o[3] = bx.z; o[4] = bx.w;
        }
    }
    for (int i = threadIdx.x; i < niou * m; i += kEvalThreads) s_win[i] = 0x7fffffff;
    __syncthreads();
    // detections: un-letterbox (:152-156), best class-matching label by torchvision box_iou
    // (float32: (area1 + area2) - inter, true division; NaN never matches)
    for (int d0 = 0; d0 < n; d0 += kEvalThreads) {
        const int d = d0 + threadIdx.x;
        float best = -1.0f;
        int best_l = -1;
        if (d < n) {
            const float* p = dets + ((size_t)b * max_det + d) * 6;
            const float4 bx = prescaled ? make_float4(p[0], p[1], p[2], p[3])
                                        : unletterbox(make_float4(p[0], p[1], p[2], p[3]), g);
            const float conf = p[4], cls = p[5];
            if (predn) {
                float* o = predn + ((size_t)b * max_det + d) * 6;
                o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = conf; o[5] = cls;
            }
            const float area_d = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
            for (int l = 0; l < m; ++l) {
                const float* s = s_lab + 5 * l;
                if (s[0] != cls) continue;                                       // :284
                const float w = fmaxf(__fsub_rn(fminf(s[3], bx.z), fmaxf(s[1], bx.x)), 0.f);
                const float h = fmaxf(__fsub_rn(fminf(s[4], bx.w), fmaxf(s[2], bx.y)), 0.f);
                const float inter = __fmul_rn(w, h);
                const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(s_area[l], area_d), inter));
                if (iou > best) { best = iou; best_l = l; }                      // first maximum: lowest label index
            }
            if (best_l >= 0)
                for (int i = 0; i < niou; ++i)
                    if (best >= s_thr[i]) atomicMin(&s_win[i * m + best_l], d); // :287,295-297
        }
        __syncthreads();
    }
    // a detection is correct at threshold i iff it is the winner of its best label (:298)
    for (int d0 = 0; d0 < max_det; d0 += kEvalThreads) {
        const int d = d0 + threadIdx.x;
        if (d >= max_det) break;
        // recompute nothing: winners only exist for passing (detection, threshold) pairs
        uint8_t* o = correct + ((size_t)b * max_det + d) * niou;
        for (int i = 0; i < niou; ++i) o[i] = 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < niou * m; i += kEvalThreads) {
        const int d = s_win[i];
        if (d != 0x7fffffff) correct[((size_t)b * max_det + d) * niou + (i / m)] = 1;
    }
}

}  // namespace vk

using namespace vk;

extern "C" size_t vk_eval_match_smem_bytes(int niou, int max_labels) {
    if (niou <= 0 || max_labels < 0) return 0;
    const size_t ml = max_labels > 0 ? max_labels : 1;
    return ml * 6 * sizeof(float) + (size_t)niou * ml * sizeof(int32_t);
}

extern "C" int vk_eval_match(const float* dets, const int32_t* det_counts, int batch, int max_det,
                             const float* labels, const int32_t* label_offsets, int max_labels,
                             const int32_t* img0_hw, int img1_h, int img1_w, int prescaled, const float* iouv, int niou,
                             float* predn, float* labeln, uint8_t* correct, vk_stream_t stream) {
    if (batch == 0) return VK_OK;
    if (!dets || !det_counts || !label_offsets || (!img0_hw && !prescaled) || !iouv || !correct || batch < 0 || max_det <= 0)
        return fail_arg("vk_eval_match: null/negative argument");
    if (max_labels > 0 && !labels) return fail_arg("vk_eval_match: labels is NULL");
    if (niou < 1 || niou > 32) return fail_code(VK_E_LIMIT, "vk_eval_match: niou %d outside 1..32", niou);
    if (!prescaled && (img1_h <= 0 || img1_w <= 0)) return fail_arg("vk_eval_match: canvas %dx%d", img1_h, img1_w);
    const int ml = max_labels > 0 ? max_labels : 1;
    const size_t smem = vk_eval_match_smem_bytes(niou, ml);
    if (smem > 200 * 1024)
        return fail_code(VK_E_LIMIT, "vk_eval_match: %d labels in one image need %zu B of shared memory", max_labels, smem);
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&eval_match_kernel), smem, "vk_eval_match")) return rc;
    eval_match_kernel<<<batch, kEvalThreads, smem, as_stream(stream)>>>(
        dets, det_counts, max_det, labels, label_offsets, img0_hw, img1_h, img1_w, iouv, niou, ml, prescaled,
        predn, labeln, correct);
    count_launch();
    return check_launch("eval_match_kernel");
}
