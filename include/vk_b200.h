/* vk_b200.h -- C-ABI of libvk_b200.so: Vision-Kit's YOLO detection data path on B200.
 *
 * The reference (ArkarPhyo1310/Vision-Kit) has no FFI layer: its boundary is the Python
 * call surface.  The host shims in vision_kit_b200/ keep that surface and
 * reach the kernels only through the entry points declared here, loaded with ctypes.
 * Each entry point cites the reference lines it replaces (paths relative to
 * /root/reference/vision_kit/).
 *
 * Conventions
 *   - return 0 = OK; < 0 = invalid argument (VK_E_*); > 0 = a cudaError_t.
 *     vk_last_error() returns a thread-local message for the last non-zero return.
 *   - the library never allocates or frees device memory: outputs and workspaces are
 *     caller-owned device buffers (sizes from the vk_*_workspace_bytes helpers).
 *   - all device work is enqueued on the given stream; no entry point synchronises.
 *   - "host" pointers are read before the call returns; "dev" pointers are device memory.
 *   - no CPU fallback: every compute entry point launches sm_100a kernels.
 */
#ifndef VK_B200_H_
#define VK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vk_stream_t; /* a cudaStream_t (torch.cuda.current_stream().cuda_stream) */

#define VK_OK 0
#define VK_E_ARG (-1)      /* bad argument */
#define VK_E_WORKSPACE (-2) /* workspace too small */
#define VK_E_LIMIT (-3)    /* a compiled-in limit was exceeded (VK_MAX_*) */

#define VK_MAX_LEVELS 4
#define VK_MAX_ANCHORS 8     /* anchors per level */
#define VK_MAX_DET 1024      /* largest max_det (reference default 300) */
#define VK_MAX_SEGMENTS 2048 /* candidate segments per image (64 rows each) */

/* ---------------------------------------------------------------- library info */
int vk_version(void);                    /* 100*major + minor */
const char* vk_last_error(void);         /* thread-local, never NULL */
uint64_t vk_launch_count(void);          /* kernels launched by this library so far */

int vk_build_arch(void);                 /* 100 => sm_100a */

/* ---------------------------------------------------------------- letterbox
 * Replaces utils/image_proc.py:12-60 `resize` (= demo/processing.py:59-97
 * `ImageProcessor.resize`) and the normalise step of demo/processing.py:45-52
 * `preprocess` / core/train/det_trainer.py:74-75.
 */
typedef struct VkLbGeom {
    double ratio;          /* image_proc.py:27-31 */
    double pad_w, pad_h;   /* the (dw, dh) the reference returns (:55,60): halved when letterbox */
    int32_t new_w, new_h;  /* :34-35, banker's rounding */
    int32_t top, bottom, left, right; /* :46-53 */
    int32_t out_h, out_w;  /* canvas = new + pads (img_sz unless auto) */
    int32_t needs_resize;  /* :42 */
    int32_t reserved;
} VkLbGeom;

/* Host-only scalar set-up, float64 exactly as the reference's Python. */
int vk_letterbox_geometry(int src_h, int src_w, int img_h, int img_w, int stride,
                          int letterbox, int scaleup, int auto_, VkLbGeom* out);

/* Eval-time dataset ingest: `YOLODataset.load_resized_image` (data/datasets/yolo.py:144-160:
 * r = max(img_sz) / max(h, w), target (int(w*r), int(h*r)), INTER_LINEAR) + the validation
 * pipeline's `A.PadIfNeeded` (data/augmentations.py:197-200: centred, before = int(diff / 2.0)).
 * Fills the same VkLbGeom (ratio = r, pad = (left, top)); the result feeds vk_letterbox_batch,
 * whose float32/bf16 NCHW output is `permute(0,3,1,2).float() / 255` of core/train/det_trainer.py:74-75. */
int vk_dataset_geometry(int src_h, int src_w, int img_h, int img_w, VkLbGeom* out);

typedef struct VkLbDesc {  /* one source image; 48 bytes */
    const uint8_t* src;    /* dev, HWC uint8, 3 channels */
    int64_t pitch;         /* bytes between source rows */
    int32_t src_h, src_w;
    int32_t new_h, new_w;  /* resized interior (== src size when no resize) */
    int32_t top, left;     /* interior origin on the canvas */
    int32_t reserved0, reserved1;
} VkLbDesc;

#define VK_LB_F32_NCHW 0   /* float32 (B,3,H,W), value/255 (true division) */
#define VK_LB_BF16_NCHW 1  /* bfloat16 (B,3,H,W), RN(float32(value/255)) */
#define VK_LB_U8_NHWC 2    /* uint8 (B,H,W,3), what `resize` returns */

size_t vk_letterbox_workspace_bytes(int batch, int out_h, int out_w);

/* descs_host: `batch` descriptors (host).  descs_dev: the same array already on the
 * device, or NULL to have the library upload descs_host into the workspace.
 * swap_rb: 1 = source is BGR, write RGB (cv2.cvtColor BGR2RGB, demo/processing.py:47).
 * pad_rgb: border colour in OUTPUT channel order, 0x00BBGGRR-style packing c0|c1<<8|c2<<16.
 * Resize arithmetic = OpenCV INTER_LINEAR 8-bit fixed point, bit-exact. */
int vk_letterbox_batch(const VkLbDesc* descs_host, const VkLbDesc* descs_dev, int batch,
                       int out_h, int out_w, int swap_rb, uint32_t pad_rgb, int dst_fmt,
                       void* dst, void* ws, size_t ws_bytes, vk_stream_t stream);

/* Element type of the tensors a filter / decode call reads (AMP eval: the head's convs emit fp16,
 * scripts/main.py:41).  Half values are up-cast exactly, so results equal the float32 call on the
 * up-cast tensor bit for bit. */
#define VK_F32 0
#define VK_F16 1
#define VK_BF16 2

/* ---------------------------------------------------------------- Detect decode
 * Replaces the eval branch of models/heads/yolov5.py:54-78 and
 * models/heads/yolov7.py:62-90 after the 1x1 conv.
 */
#define VK_HEAD_V5 0 /* xy = (s*2 + (g-0.5))*stride   yolov5.py:68,88 */
#define VK_HEAD_V7 1 /* xy = (s*2 - 0.5 + g)*stride   yolov7.py:80,95 */

typedef struct VkHeadCfg {
    int32_t variant;             /* VK_HEAD_V5 | VK_HEAD_V7 */
    int32_t nl, na, nc;          /* levels, anchors per level, classes; no = nc + 5 */
    int32_t ny[VK_MAX_LEVELS], nx[VK_MAX_LEVELS];
    float stride[VK_MAX_LEVELS];
    float anchors[VK_MAX_LEVELS][2 * VK_MAX_ANCHORS]; /* pixels: (w,h) per anchor */
} VkHeadCfg;

int vk_head_rows(const VkHeadCfg* cfg);  /* sum_l na*ny*nx (25200 at 640) */

/* levels[l]: dev (B, na*no, ny_l, nx_l) contiguous, element type `dtype` (VK_F32 / VK_F16 / VK_BF16).
 * pred: dev float32 (B, rows, no).  raw[l]: dev float32 (B, na, ny_l, nx_l, no) or raw == NULL to skip the permuted logits copy the reference also returns (yolov5.py:60,78). */
int vk_detect_decode(const VkHeadCfg* cfg, const void* const* levels, int dtype, int batch,
                     float* pred, float* const* raw, vk_stream_t stream);

/* ---------------------------------------------------------------- candidates
 * Confidence filter of utils/image_proc.py:99-151 (= demo/processing.py:110-164):
 * obj > conf, cls *= obj, cxcywh->xyxy (utils/bboxes.py:103-111), multi-label
 * `nonzero` or best-class `max`, optional class filter.
 *
 * A candidate set lives in caller-owned device buffers described by VkCandBuf.  Every 64-row
 * tile ("segment") owns the fixed slot range [seg * T, (seg + 1) * T), T = vk_cand_tile_slots(),
 * and writes its seg_count[seg] candidates to the front of it, so the buffer needs
 * cap >= segs * T, can never overflow, and no tile waits for another.  A candidate is the pair
 * (score, id), id = row * nc + cls: ascending id is the reference's candidate order (row, then
 * class -- the order that breaks score ties in its argsort and that torchvision's indices refer
 * to), so the order of the candidates INSIDE a segment is free; segments are in row order.
 *
 * list: optional scratch of vk_nms_batched for images with more than list_cap candidates (eval
 * thresholds, ~240 k per image): a sampled score histogram picks the score bound above which a
 * grid-wide pass copies the candidates, as (ordered score << 32 | ~(row*nc + cls)), into the image's list, and
 * the per-image kernel sorts from that list.  Results never depend on it; allocate it (8192 entries
 * per image is plenty) when such images are expected.
 */
#define VK_HIST_BINS 1024   /* bin = (0x3f800000 - score bits) >> 20: 8 bins per octave below 1.0 */
#define VK_CTRL_WORDS 4     /* ctrl rows: 0 candidate count, 1 flags | slots per tile / 64 << 8, 2 list entries, 3 list bound */
#define VK_FLAG_LIST 2      /* list[b] holds every candidate with ordered score >= bound (set by vk_nms_batched) */

typedef struct VkCandBuf {
    uint64_t* cand;      /* dev [batch][cap]: low 32 = score bits, high 32 = row*nc + cls */
    float* boxes;        /* dev [batch][rows][4] xyxy of rows that produced candidates */
    int32_t* ctrl;       /* dev [VK_CTRL_WORDS][batch]; zeroed by the filter call */
    int32_t* seg_count;  /* dev [batch][segs] candidates of each segment */
    uint64_t* list;      /* dev [batch][list_cap] or NULL */
    int32_t cap;         /* candidate slots per image */
    int32_t rows;        /* prediction rows per image */
    int32_t segs;        /* segments per image (vk_filter_segments / vk_decode_filter_segments) */
    int32_t nc;
    int32_t list_cap;    /* entries per image in `list` (0 with list == NULL) */
    int32_t reserved;
} VkCandBuf;

int vk_cand_tile_slots(int nc, int multi_label);     /* 64 * nc (multi-label, nc > 1) or 64 */
int vk_filter_segments(int rows);                    /* for vk_filter_pred */
int vk_decode_filter_segments(const VkHeadCfg* cfg); /* for vk_decode_filter */

/* Which kernel a filter call uses (per call; both produce identical candidates, boxes and counts):
 * AUTO picks from the threshold (conf < 0.05, where most rows survive: DENSE); SPARSE = one warp per
 * 64-row tile gathering only the rows with obj > conf; DENSE = every logit of every tile is read once, coalesced
 * (multi-label: a logit-domain pre-test, then the exact test on the survivors only). */
#define VK_FILTER_AUTO 0
#define VK_FILTER_SPARSE 1
#define VK_FILTER_DENSE 2
#define VK_FILTER_DENSE_ONEPASS 3   /* DENSE without the pre-test phase (sigmoid on every logit): kept for measurements */

/* class_mask: dev uint32[(nc+31)/32] bitmap of allowed classes or NULL (classes=None). */
int vk_filter_pred(const void* pred, int dtype, int batch, int rows, int nc, float conf_thres,
                   int multi_label, const uint32_t* class_mask, int kernel, const VkCandBuf* out,
                   vk_stream_t stream);

/* Fused Detect decode + confidence filter straight from the conv outputs: the
 * (B, rows, no) prediction tensor is never materialised. */
int vk_decode_filter(const VkHeadCfg* cfg, const void* const* levels, int dtype, int batch,
                     float conf_thres, int multi_label, const uint32_t* class_mask, int kernel,
                     const VkCandBuf* out, vk_stream_t stream);

/* Fused Detect head: the 1x1 conv of models/heads/yolov5.py:58 / yolov7.py:67-71
 * (implicit layers already folded into weights/biases), the decode and nms()'s candidate selection in one
 * kernel -- D[s][co] = X^T W^T on the tensor cores (tcgen05, TF32 inputs, fp32 accumulation in TMEM), the
 * filter straight from TMEM.  The (B, na*no, ny, nx) conv output is never written.
 *   feats[l]   dev (B, cin[l], ny, nx) float32, 16-byte aligned, ny*nx % 4 == 0, cin[l] % 32 == 0
 *   weights[l] dev (na*no, cin[l]) float32 (the conv weight with its trailing 1x1 dropped); biases[l] (na*no) or NULL
 *   out        the candidate buffer of vk_decode_filter (same sizes), consumed by vk_nms_batched
 *   fault      dev int32, set to 1 if a tensor-core completion wait timed out (never observed; the
 *              kernel then terminates instead of hanging)
 *   kernel     VK_CONV_PERSISTENT (one CTA per SM: two producer teams, one MMA-issuing lane, two epilogue
 *              groups over a double-buffered TMEM accumulator) or VK_CONV_TILE (one tile per CTA, 2 CTAs
 *              per SM); identical results.
 * Logits differ from an fp32 conv by TF32 input rounding (~1e-3 relative). */
#define VK_CONV_TILE 0
#define VK_CONV_PERSISTENT 1
int vk_conv_decode_filter(const VkHeadCfg* cfg, const float* const* feats, const int32_t* cin,
                          const float* const* weights, const float* const* biases, int batch,
                          float conf_thres, int multi_label, const uint32_t* class_mask, int kernel,
                          const VkCandBuf* out, int32_t* fault, vk_stream_t stream);

/* ---------------------------------------------------------------- NMS
 * utils/image_proc.py:154-182: top-max_nms cut by score (stable), class offset
 * cls*max_wh, torchvision.ops.nms greedy suppression, [:max_det].
 * iou_thres is the python float; the comparison is made the way torchvision's CPU kernel
 * makes it (float32 IoU promoted to double).
 *
 * dets: dev float32 (B, max_det, 6) [x1,y1,x2,y2,conf,cls], rows >= det_counts[b] zeroed.
 * keep_idx: dev int64 (B, max_det) or NULL -- indices torchvision.ops.nms returned
 *   (into the candidate list the reference handed it), -1 padded.
 * status: dev int32[batch] or NULL; always written 0 (kept for callers that check it: a
 *   candidate buffer sized as above cannot overflow).
 * Two launches: a grid-wide selection pass that builds the per-image top list (only with a `list`;
 * it returns at once for images that fit a stage anyway), and one CTA per image that consumes
 * candidates in descending score order, a stage of <= 2048 at a time, until max_det boxes are kept.
 * Writes ctrl rows 1-3 of the candidate buffer (the list it builds); needs no workspace.
 */
int vk_nms_batched(const VkCandBuf* cand, int batch, double iou_thres, int agnostic, int max_nms,
                   int max_det, float max_wh, float* dets, int32_t* det_counts, int64_t* keep_idx,
                   int32_t* status, vk_stream_t stream);

/* ---------------------------------------------------------------- small ops */
/* utils/image_proc.py:63-80 `scale_coords` (+ utils/bboxes.py:50-59 `clip_coords`):
 * in place on n rows of `row_stride` floats, columns 0..3 = xyxy.  clip_w/clip_h < 0
 * skips the clip (demo/processing.py:99-105). */
int vk_scale_coords(float* coords, int n, int row_stride, float pad_w, float pad_h,
                    float gain, int subtract_pad, float clip_w, float clip_h,
                    vk_stream_t stream);

/* utils/bboxes.py:103-111 `cxcywh_to_xyxy` on n rows of 4 floats (out may alias in). */
int vk_cxcywh_to_xyxy(const float* in, float* out, int n, vk_stream_t stream);

/* ---------------------------------------------------------------- evaluator matching */
/* The per-image body of `DetEvaluator.evaluate` (core/eval/det_evaluator.py:141-178) for a whole
 * batch, one block per image: un-letterbox + clip of detections and labels (`scale_coords`,
 * utils/image_proc.py:63-80), torchvision `box_iou`, and `process_batch` (:274-300) at every IoU
 * threshold.
 *   dets [batch][max_det][6], det_counts [batch]   the NMS output (canvas pixels)
 *   labels [n][6] = image, cls, cx, cy, w, h       canvas pixels (after `targets[:, 2:] *= (w,h,w,h)`, :139),
 *                                                  grouped by image: rows label_offsets[b] .. label_offsets[b+1]
 *   max_labels                                     largest label count of one image (sizes shared memory)
 *   img0_hw [batch][2]                             original (h, w) of every image; img1 = the canvas
 *   prescaled                                      1 = `process_batch` alone (:274-300): dets and labels are already in
 *                                                  one frame, labels carry x1, y1, x2, y2 instead of cx, cy, w, h, img0_hw unused
 *   iouv [niou]                                    thresholds (reference: linspace(0.5, 0.95, 10)), niou <= 32
 * Outputs: predn [batch][max_det][6] and labeln [n][5] = cls, x1, y1, x2, y2 in original pixels
 * (either may be NULL), correct [batch][max_det][niou] (1 = true positive; rows >= count are 0).
 * IoU ties between two labels of one detection go to the lower label index (the reference's order
 * there comes from an unstable argsort). */
size_t vk_eval_match_smem_bytes(int niou, int max_labels);
int vk_eval_match(const float* dets, const int32_t* det_counts, int batch, int max_det,
                  const float* labels, const int32_t* label_offsets, int max_labels,
                  const int32_t* img0_hw, int img1_h, int img1_w, int prescaled, const float* iouv, int niou,
                  float* predn, float* labeln, uint8_t* correct, vk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VK_B200_H_ */
