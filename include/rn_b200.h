/*
 * rn_b200.h -- C-ABI of librn_b200.so: the B200 (sm_100a) implementation of RetinaNet's
 * anchor + detection-head path.
 *
 * The reference (jabhinav/RetinaNet-for-Table-Detection) is pure Python and has no FFI; the
 * functions below are what its Python call sites bind instead of numpy / TensorFlow ops.  Each
 * entry point cites the reference interface it replaces (file:line, relative to the reference
 * tree).  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - every `*_dev` / tensor pointer is caller-owned DEVICE memory (e.g. a torch allocation);
 *     the library never allocates device memory and never synchronises the stream;
 *   - small tables marked "host" are read synchronously during the call;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - all tensors are dense, row-major, float32 unless stated; row counts are 64-bit;
 *   - return value: RN_OK (0) or a negative RN_ERR_* code; rn_last_error() gives the message of
 *     the calling thread's last failure.  Nothing throws.  Re-entrant per (stream, workspace).
 */
#ifndef RN_B200_H
#define RN_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RN_OK             0
#define RN_ERR_BAD_ARG   -1   /* null pointer, non-positive size, unsupported parameter      */
#define RN_ERR_CUDA      -2   /* launch / runtime failure (cudaGetLastError)                 */
#define RN_ERR_WORKSPACE -3   /* workspace smaller than rn_*_workspace_bytes() reports       */

#define RN_MAX_LEVELS 8       /* pyramid levels per anchor table                              */

#define RN_BCE_TF2    0       /* K.binary_crossentropy, tf.keras 2.x / Keras 2.3 form         */
#define RN_BCE_LOGITS 1       /* standalone Keras <= 2.2 form (via logits)                    */

int         rn_version(void);
const char* rn_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * K1  anchor generation + IoU/argmax matching + labels + regression targets, one launch per batch.
 * Replaces: anchors_for_shape (model/anchors.py:169-204) + compute_overlap (model/utils.py:180-211)
 *           + compute_gt_annotations (model/anchors.py:96-117) + bbox_transform (:282-313)
 *           + the per-image loop of anchor_targets_bbox (:68-90).
 *
 * Anchors are either generated in-kernel from the level table (anchors_dev == NULL; 0 bytes read)
 * or read from an explicit (N,4) float64 array (anchors_dev != NULL; API fidelity:
 * anchor_targets_bbox takes `anchors` as an argument).
 *
 *   base_anchors_dev  (num_levels, anchors_per_cell, 4) float64: generate_anchors() per level
 *   level_hw          host, (num_levels, 2) int: feature-map (H_l, W_l)      [guess_shapes]
 *   level_stride      host, (num_levels) int
 *   num_anchors       N; must equal sum_l H_l*W_l*anchors_per_cell when anchors are generated
 *   gt_boxes_dev      (B, Gmax, 4) float64 x1,y1,x2,y2;  gt_labels_dev (B, Gmax) int32 in [0, C];
 *   gt_count_dev      (B) int32, number of valid GT rows per page (0 allowed)
 *   img_hw_dev        (B, 2) int32: each page's own (H, W) for the border-ignore rule
 *                     (model/anchors.py:85-90); NULL disables the rule
 *   neg_overlap/pos_overlap   compared in float32 (max_iou > neg  -> ignore unless >= pos)
 * Outputs
 *   regression_out    (B, N, 5)   float32: 4 targets + state (-1 ignore / 0 bg / 1 fg)
 *   labels_out        (B, N, C+1) float32: one-hot + state
 *   argmax_out        (B, N) int32 or NULL: index of the best-overlapping GT (first max)
 *   npos_out          (B) int32 or NULL: number of state==1 anchors per page
 *   npos_total_out    1 float or NULL: the same count summed over the batch -- directly usable as
 *                     `npos_dev` of the loss kernels (integer-valued, so the float sum is exact)
 */
int rn_anchor_targets(const double* base_anchors_dev, const int* level_hw, const int* level_stride,
                      int num_levels, int anchors_per_cell,
                      const double* anchors_dev, long long num_anchors,
                      const double* gt_boxes_dev, const int* gt_labels_dev, const int* gt_count_dev,
                      const int* img_hw_dev, int B, int Gmax, int C,
                      float neg_overlap, float pos_overlap,
                      float* regression_out, float* labels_out, int* argmax_out, int* npos_out,
                      float* npos_total_out, void* stream);

/* The same with a launch order for the pages: page_order_dev = a permutation of 0 .. B-1 (device int32) or NULL.  The
 * kernel hands its CTAs out page by page, so the pages at the end of the order make the tail of the launch; callers that
 * know the annotations put the heaviest pages (most / largest tables) first.  Results do not depend on the order; an
 * array that is not a permutation leaves pages unwritten (not checked). */
int rn_anchor_targets_ordered(const double* base_anchors_dev, const int* level_hw, const int* level_stride,
                              int num_levels, int anchors_per_cell,
                              const double* anchors_dev, long long num_anchors,
                              const double* gt_boxes_dev, const int* gt_labels_dev, const int* gt_count_dev,
                              const int* img_hw_dev, int B, int Gmax, int C,
                              float neg_overlap, float pos_overlap,
                              float* regression_out, float* labels_out, int* argmax_out, int* npos_out,
                              float* npos_total_out, const int* page_order_dev, void* stream);

/* The training step's form of the same targets (an EXTENSION; rn_anchor_targets stays the drop-in for anchor_targets_bbox):
 * generated anchors, one class, <= 9 anchor types per cell, <= 32 tables per page.  labels_out, npos_out, npos_total_out are
 * exactly what rn_anchor_targets writes; of regression_out (B, N, 5) ONLY the rows of anchors whose final state is 1 are
 * written (bit-identical to rn_anchor_targets' rows), every other row is left untouched.  That is all the smooth-L1 loss
 * ever reads when it takes the anchor state from the label tensor (model/losses.py:72-74 gathers the state == 1 rows;
 * rn_loss_fwd_bwd with RN_LOSS_SHARED_STATE), and it takes 20 of the 28 bytes per anchor and the four fp64 quotients of
 * every non-positive anchor out of the kernel. */
int rn_anchor_targets_sparse(const double* base_anchors_dev, const int* level_hw, const int* level_stride,
                             int num_levels, int anchors_per_cell, long long num_anchors,
                             const double* gt_boxes_dev, const int* gt_labels_dev, const int* gt_count_dev,
                             const int* img_hw_dev, int B, int Gmax,
                             float neg_overlap, float pos_overlap,
                             float* regression_out, float* labels_out, int* npos_out,
                             float* npos_total_out, const int* page_order_dev, void* stream);

/* anchors_for_shape (model/anchors.py:169-204) on the device: (N,4) float64. */
int rn_anchors_f64(const double* base_anchors_dev, const int* level_hw, const int* level_stride,
                   int num_levels, int anchors_per_cell, double* anchors_out, void* stream);

/* compute_overlap (model/utils.py:180-211): (M,4) x (G,4) float64 -> (M,G) float32 IoU. */
int rn_compute_overlap(const double* boxes1_dev, long long M, const double* boxes2_dev, int G,
                       float* iou_out, void* stream);

/* bbox_transform (model/anchors.py:282-313), row-wise: anchors (N,4) and gt_boxes (N,4) float64 ->
 * ((gt - a) / {w,h} - mean) / std as (N,4) float64.  mean4 / std4: host double[4]. */
int rn_bbox_transform(const double* anchors_dev, const double* gt_boxes_dev, long long N,
                      const double* mean4, const double* std4, double* out_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2  losses, forward + backward in one pass.
 * Replaces: _focal (model/losses.py:13-44) and _smooth_l1 (:58-90) and their TF autodiff backward.
 *
 *   R rows = B*N anchors.  npos_dev: device float holding the positive-anchor COUNT to normalise
 *   with (e.g. the all-reduced global count); NULL = count state==1 over these R rows first.
 *   The normaliser is max(1, count).  grad_* may be NULL (forward only).
 *   losses are written as device floats.  workspace: rn_loss_workspace_bytes().
 */
size_t rn_loss_workspace_bytes(void);

int rn_count_positive(const float* y_true, long long R, int row_width, float* npos_out_dev,
                      void* workspace, size_t workspace_bytes, void* stream);

int rn_focal_fwd_bwd(const float* y_true_cls /*(R,C+1)*/, const float* y_pred /*(R,C)*/,
                     long long R, int C, float alpha, float gamma, int bce_mode,
                     const float* npos_dev, float* loss_out_dev, float* grad_out /*(R,C)*/,
                     void* workspace, size_t workspace_bytes, void* stream);

int rn_smooth_l1_fwd_bwd(const float* y_true_reg /*(R,5)*/, const float* y_pred /*(R,4)*/,
                         long long R, float sigma,
                         const float* npos_dev, float* loss_out_dev, float* grad_out /*(R,4)*/,
                         void* workspace, size_t workspace_bytes, void* stream);

/* both losses in ONE launch; losses_out_dev[0]=focal, [1]=smooth_l1, [2]=normaliser used.
 * flags: RN_LOSS_SHARED_STATE = the smooth-L1 part takes the anchor state from y_true_cls's last column
 * instead of y_true_reg's.  anchor_targets_bbox writes the same state into both (model/anchors.py:73-77,
 * 89-90), so for targets that come from rn_anchor_targets the result is identical while 20 B/anchor of
 * reads disappear; leave 0 for arbitrary y_true tensors. */
#define RN_LOSS_SHARED_STATE 1
#define RN_LOSS_NPOS_PEER_BOX 2   /* npos_dev is this rank's peer mailbox (rn_peer_box_create): the normaliser is
                                    the sum of the counts all ranks published for the current step           */
#define RN_LOSS_PEER_LAG1     8   /* with RN_LOSS_NPOS_PEER_BOX: use the step published BEFORE the latest one
                                    (pipelined schedule: K1 + publish of the next batch run ahead of this K2)  */
#define RN_LOSS_PEER_LOSSES   32  /* with RN_LOSS_NPOS_PEER_BOX: the two loss sums are exchanged through the mailbox as well (by
                                   * the kernel's last CTA, added in rank order), so losses_out is the loss of the whole merged
                                   * batch on every rank (model/losses.py:44, :90 with RetinaNet.py:106-112).  At most ONE such
                                   * launch per exchanged step (not for page-chunked launches). */
#define RN_LOSS_PEER_PUBLISH  16  /* with RN_LOSS_NPOS_PEER_BOX on a mailbox prepared by rn_peer_box_bind: FUSED publish --
                                    this launch itself stores the rank's count into every rank's mailbox (CTA 0, P2P
                                    stores over NVLink) before all CTAs wait for the ranks' counts, and completes the
                                    step when its last CTA finishes; rn_peer_publish is not called for such steps.
                                    Send, wait and loss arithmetic are one kernel.  Not combinable with PEER_LAG1.      */
int rn_loss_fwd_bwd(const float* y_true_cls, const float* cls_pred, const float* y_true_reg,
                    const float* reg_pred, long long R, int C,
                    float alpha, float gamma, int bce_mode, float sigma,
                    const float* npos_dev, float* losses_out_dev,
                    float* grad_cls /*(R,C)*/, float* grad_reg /*(R,4)*/,
                    int flags, void* workspace, size_t workspace_bytes, void* stream);

/* N2  the same fused losses fed by the heads' PER-LEVEL outputs (model/defineModel.py:119-123, 163-166, 217):
 * level l contributes cls_levels[l] (B, n_l, 1) and reg_levels[l] (B, n_l, 4); sum_l n_l = N.  No
 * Concatenate(axis=1) copy is needed, and with RN_LOSS_FROM_LOGITS the classification tensors hold the LOGITS
 * (the heads' Activation('sigmoid') is fused: p = 1 / (1 + exp(-z))) and grad_cls_levels receive
 * d loss / d logit.  Targets are the concatenated (B, N, .) tensors rn_anchor_targets writes.
 * Covers the reference's table-detection configuration: C == 1, gamma == 2, RN_BCE_TF2, RN_LOSS_SHARED_STATE.
 * The *_levels arguments are HOST arrays of num_levels device pointers; level_rows is a host array (n_l). */
#define RN_LOSS_FROM_LOGITS 4
int rn_loss_fwd_bwd_levels(const float* y_true_cls, const float* y_true_reg,
                           const float* const* cls_levels, const float* const* reg_levels,
                           const long long* level_rows, int num_levels, int B, int C,
                           float alpha, float gamma, int bce_mode, float sigma,
                           const float* npos_dev, float* losses_out_dev,
                           float* const* grad_cls_levels, float* const* grad_reg_levels,
                           int flags, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Detection head, layer by layer.
 */
/* Anchors layer (model/layers.py:42-53 + TF shift model/utils.py:51-80), all levels concatenated
 * (model/defineModel.py:283-293): float32 anchors (B, N, 4).  base_anchors_f32_dev:
 * (num_levels, anchors_per_cell, 4) float32 (generate_anchors cast to floatx, layers.py:34). */
int rn_anchors_f32(const float* base_anchors_f32_dev, const int* level_hw, const int* level_stride,
                   int num_levels, int anchors_per_cell, int B, float* anchors_out, void* stream);

/* RegressBoxes / bbox_transform_inv (model/layers.py:136-138, model/utils.py:84-112).
 * mean/std: host float[4]. */
int rn_regress_boxes(const float* boxes, const float* deltas, long long R,
                     const float* mean4, const float* std4, float* out, void* stream);

/* ClipBoxes (model/layers.py:157-171): x to [0,width], y to [0,height]. */
int rn_clip_boxes(const float* boxes, long long R, float width, float height, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3+K4+K5  FilterDetections (model/layers.py:177-264, 298-332): score threshold -> per-class
 * sort (score desc, anchor asc) -> greedy NMS (strict >) -> cross-class top-k -> pad with -1.
 *
 * Two front ends share the sort/NMS/merge back end:
 *   rn_filter_detections         boxes (B,N,4) already decoded+clipped  [layer-by-layer API]
 *   rn_decode_filter_detections  fused K3: anchors generated in-kernel (float32, as the Anchors
 *                                layer does), decode (RegressBoxes) + clip (ClipBoxes) + threshold
 *                                in one pass over regression (B,N,4) / classification (B,N,C).
 *   class_specific, nms          flags as in the reference; nms_threshold / score_threshold float32
 *   max_detections               <= 1024
 *   pre_nms_top_k                0 = off (reference behaviour).  >0: visit only the k best
 *                                candidates per (page,class) -- an extension, NOT in the reference.
 *   cand_cap                     capacity of each (page,class) candidate slab; N = always exact.
 *                                Overflow is reported in status_out (see below), never silent.
 * Outputs (padded with -1): out_boxes (B,M,4) f32, out_scores (B,M) f32, out_labels (B,M) i32,
 *   out_indices (B,M) i32 or NULL (anchor index of each detection; used to gather `other`),
 *   status_out_dev (B) i32 or NULL: 0 ok, 1 = a slab overflowed for this page (results invalid).
 */
size_t rn_filter_workspace_bytes(int B, long long N, int C, int class_specific,
                                 long long cand_cap, int max_detections);

int rn_filter_detections(const float* boxes, const float* classification,
                         int B, long long N, int C, int class_specific, int nms,
                         float score_threshold, float nms_threshold, int max_detections,
                         int pre_nms_top_k, long long cand_cap,
                         float* out_boxes, float* out_scores, int* out_labels, int* out_indices,
                         int* status_out_dev, void* workspace, size_t workspace_bytes, void* stream);

int rn_decode_filter_detections(const float* base_anchors_f32_dev, const int* level_hw,
                                const int* level_stride, int num_levels, int anchors_per_cell,
                                const float* regression, const float* classification,
                                int B, long long N, int C,
                                const float* mean4, const float* std4, float clip_width, float clip_height,
                                int class_specific, int nms,
                                float score_threshold, float nms_threshold, int max_detections,
                                int pre_nms_top_k, long long cand_cap,
                                float* out_boxes, float* out_scores, int* out_labels, int* out_indices,
                                int* status_out_dev, void* workspace, size_t workspace_bytes, void* stream);

/* the `other` tensors of filter_detections (model/layers.py:247, :255): out[b, m, :] = other[b, indices[b, m], :] for the kept
 * detections, -1 for the padding (indices < 0).  other (B, N, D) f32, indices (B, M) i32 (out_indices above), out (B, M, D). */
int rn_gather_other(const float* other, const int* indices, int B, long long N, int M, int D, float* out, void* stream);

/* tf.image.non_max_suppression (call site model/layers.py:211) for one box set:
 * out_indices (max_output) i32 in selection order, padded with -1; out_count_dev: 1 int32. */
size_t rn_nms_workspace_bytes(long long K, int max_output);
int rn_nms(const float* boxes /*(K,4)*/, const float* scores /*(K)*/, long long K,
           int max_output, float iou_threshold, int* out_indices, int* out_count_dev,
           void* workspace, size_t workspace_bytes, void* stream);

/* profiling aid: when enabled, k_segment_nms adds its per-phase clock64() ticks (thread 0 of every CTA) to the first
 * 8 uint64 of the filter workspace (select, gather, sort, group fetch, suppression tests, resolve).  Off by default. */
int rn_debug_nms_timing(int enable);
/* measurement hook (process-wide, not for concurrent use): which stages the following filter calls launch -- bit 0 the workspace
 * reset + k_threshold_keys, bit 1 k_segment_nms, bit 2 k_merge_topk; 7 = all (the default; anything else is for timing only).
 * A stage launched alone works on what an earlier full call left in the same workspace, so bench.py can time each kernel as a
 * train of back-to-back launches of that kernel alone (no event between kernels: the interval between two events around ONE
 * short kernel carries 3-5 us of front-end latency that is not the kernel's).  The mask applies to rn_nms' two kernels too. */
int rn_debug_filter_stages(int mask);

/* ---------------------------------------------------------------------------------------------
 * N3  the reference's host post-step on the detections (RetinaNet.py:366-377):
 *     boxes /= image_scale (fp32 division, per page) and the score cut -- the reference walks the
 *     score-sorted detections and stops at the first score < 0.6; count_out[b] is that position
 *     (M when no score is below the cut; padding rows carry score -1).
 *   boxes (B, M, 4), scores (B, M), image_scale_dev (B) float32; boxes_out may alias boxes.
 * ------------------------------------------------------------------------------------------- */
int rn_rescale_cut(const float* boxes, const float* scores, const float* image_scale_dev, int B, int M,
                   float min_score, float* boxes_out, int* count_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * N4  the page-image producer (DetectTablesUtils.py:183-262, preProcessTrainValImages / preProcessSampleImages):
 *     cv2.cvtColor(BGR2GRAY) -> cv2.adaptiveThreshold(255, ADAPTIVE_THRESH_GAUSSIAN_C, THRESH_BINARY, 11, 2) ->
 *     cv2.distanceTransform(DIST_L2 | DIST_L1 | DIST_C, maskSize 5) -> merge -> the 8-bit image imwrite encodes.
 *   bgr_dev (B, H, W, 3) uint8 (cv2.imread order), out_dev (B, H, W, 3) uint8: channel 0 = 5x5 chamfer "L2", 1 = L1, 2 = C,
 *   each rounded half-to-even and saturated at 255; binary_out_dev (B, H, W) uint8 or NULL: the thresholded page.
 *   Bit-exact against OpenCV 4.13 for page widths that are a multiple of 8 (its scalar tail columns round the blur
 *   differently); W < 65535, H * W < 2^31.
 * ------------------------------------------------------------------------------------------- */
size_t rn_preprocess_workspace_bytes(int B, int H, int W);
int rn_preprocess_pages(const unsigned char* bgr_dev, int B, int H, int W, unsigned char* out_dev,
                        unsigned char* binary_out_dev, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * C1  the path's only exchange step: the batch-global positive-anchor count of the two losses
 *     (model/losses.py:40-44, :88-90; with keras.utils.multi_gpu_model the loss sees the merged
 *     batch, RetinaNet.py:106-112), exchanged over NVLink peer memory between the ranks of one node.
 *
 * Each rank creates one mailbox (device memory owned by the library, the only allocation it makes),
 * sends the 64-byte CUDA IPC handle to the other ranks (any transport; the Python host uses
 * torch.distributed.all_gather_object) and opens theirs.  Per step, after K1:
 *     rn_peer_publish(npos_total_dev, box[rank], boxes, rank, world, stream)
 * stores {step, count} into slot `rank` of every rank's mailbox (one 8-byte P2P store per peer), and
 *     rn_loss_fwd_bwd(..., npos_dev = box[rank], ..., flags | RN_LOSS_NPOS_PEER_BOX, ...)
 * waits (on local memory, inside the kernel) until all `world` counts of the step have arrived and
 * uses their sum (RN_LOSS_PEER_LAG1: the sum of the step before the latest, so that the next batch's K1 +
 * publish can be enqueued ahead of this batch's losses and the exchange leaves the critical path; up to 4 steps
 * are kept).  Fused form (the in-order schedule's default): rn_peer_box_bind(box[rank], boxes, rank, world,
 * npos_total_dev) once, then per step only
 *     rn_loss_fwd_bwd(..., npos_dev = box[rank], ..., flags | RN_LOSS_NPOS_PEER_BOX | RN_LOSS_PEER_PUBLISH, ...)
 * -- the loss kernel sends this rank's count itself (no publish launch between K1 and K2).  A step's further loss
 * launches (page chunks) pass RN_LOSS_NPOS_PEER_BOX alone and read the completed step.
 * No NCCL call, no host synchronisation, CUDA-graph capturable.  All ranks must run
 * the same sequence of publish / loss steps.  A wait that outlasts the mailbox's timeout (default 30 s,
 * rn_peer_box_set_timeout) never hangs the GPU: it sets a STICKY error flag in the mailbox (rn_peer_box_status) and the
 * step's losses are NaN; once the flag is set every further wait of this rank gives up at once, until it is cleared.
 * world <= 16 (one NVSwitch domain).
 * ------------------------------------------------------------------------------------------- */
size_t rn_peer_box_bytes(void);
int rn_peer_box_create(int world, void** box_out, void* ipc_handle_out64);
int rn_peer_box_open(const void* ipc_handle64, void** peer_box_out);
int rn_peer_box_close(void* peer_box);
int rn_peer_box_destroy(void* box);
/* records every rank's mailbox pointer (as mapped into this process, boxes_of_all_ranks[rank] == local_box) and this rank in
 * the local mailbox: what the device-side SENDS of a loss launch need (RN_LOSS_PEER_LOSSES, RN_LOSS_PEER_PUBLISH).  Call it
 * once after the peers' mailboxes have been opened.  Synchronous. */
int rn_peer_box_connect(void* local_box, void* const* boxes_of_all_ranks /* host, (world) */, int rank, int world);
/* records in the local mailbox what RN_LOSS_PEER_PUBLISH needs: every rank's mailbox pointer (as mapped into this
 * process, boxes_of_all_ranks[rank] == local_box), this rank, and the device floats it publishes: the loss launch that
 * completes step t (steps count from 1, rn_peer_box_step() + 1 is the next) sends value_even_dev for even t and
 * value_odd_dev for odd t -- pass the same pointer twice unless the targets are double-buffered (pipelined schedule:
 * K1 of the next batch runs concurrently with this batch's loss launch).  Synchronous (cudaMemcpy); the stream that
 * uses the mailbox must be idle. */
int rn_peer_box_bind(void* local_box, void* const* boxes_of_all_ranks /* host, (world) */, int rank, int world,
                     const float* value_even_dev, const float* value_odd_dev);
/* how long a device-side wait on the mailbox may last (seconds, > 0; default 30).  Synchronous. */
int rn_peer_box_set_timeout(void* local_box, double seconds);
/* *timed_out = 1 when a wait of this rank has timed out since the flag was last cleared; clear != 0 resets it.  Synchronous. */
int rn_peer_box_status(void* local_box, int* timed_out, int clear);
/* number of steps this rank has completed / published so far (synchronous read of the device counter) */
int rn_peer_box_step(const void* local_box, unsigned long long* step_out);
int rn_peer_publish(const float* value_dev, void* local_box, void* const* boxes_of_all_ranks /* host, (world) */,
                    int rank, int world, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RN_B200_H */
