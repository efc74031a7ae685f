// Detection-head layers, one streaming kernel each (layer-by-layer API of the reference):
//   D0 Anchors      model/layers.py:42-53 + TF shift model/utils.py:51-80   (float32 arithmetic)
//   D1 RegressBoxes model/layers.py:136-138 -> bbox_transform_inv model/utils.py:84-112
//   D2 ClipBoxes    model/layers.py:157-171
// All are pure HBM streams of 16-byte rows: one float4 per thread, grid sized in multiples of the SM
// count with a grid-stride loop.  Compiled with -fmad=false so the fp32 evaluation order of the
// reference (mul, add, mul, add) is kept and results are bit-identical to an fp32 CPU evaluation.
#include "rn_common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_anchors_f32(const RnLevels lv, const float* base, int N, int B, float* out) {
    const long long total = (long long)N * B;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(e % N);
        int level, cx, cy, a;
        rn_locate(lv, n, level, cx, cy, a);
        const float* b = base + ((size_t)level * lv.anchors_per_cell + a) * 4;
        const float sx = ((float)cx + 0.5f) * (float)lv.stride[level];
        const float sy = ((float)cy + 0.5f) * (float)lv.stride[level];
        rn_stg_stream4(out + e * 4, make_float4(__ldg(b) + sx, __ldg(b + 1) + sy, __ldg(b + 2) + sx, __ldg(b + 3) + sy));
    }
}

struct Norm4 { float mean[4]; float std[4]; };

__global__ void __launch_bounds__(256) k_regress_boxes(const float* boxes, const float* deltas, long long R, const Norm4 nm, float* out) {
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < R; r += (long long)gridDim.x * blockDim.x) {
        const float4 a = rn_ldg_stream4(boxes + r * 4);
        const float4 d = rn_ldg_stream4(deltas + r * 4);
        const float w = a.z - a.x, h = a.w - a.y;
        float4 o;
        o.x = a.x + (d.x * nm.std[0] + nm.mean[0]) * w;
        o.y = a.y + (d.y * nm.std[1] + nm.mean[1]) * h;
        o.z = a.z + (d.z * nm.std[2] + nm.mean[2]) * w;
        o.w = a.w + (d.w * nm.std[3] + nm.mean[3]) * h;
        rn_stg_stream4(out + r * 4, o);
    }
}

__global__ void __launch_bounds__(256) k_clip_boxes(const float* boxes, long long R, float W, float H, float* out) {
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < R; r += (long long)gridDim.x * blockDim.x) {
        float4 v = rn_ldg_stream4(boxes + r * 4);
        v.x = fminf(fmaxf(v.x, 0.0f), W);
        v.y = fminf(fmaxf(v.y, 0.0f), H);
        v.z = fminf(fmaxf(v.z, 0.0f), W);
        v.w = fminf(fmaxf(v.w, 0.0f), H);
        rn_stg_stream4(out + r * 4, v);
    }
}

int stream_grid(long long rows) {
    long long g = (rows + 255) / 256;
    const long long cap = (long long)RN_NUM_SMS * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int rn_anchors_f32(const float* base_anchors_f32_dev, const int* level_hw, const int* level_stride,
                              int num_levels, int anchors_per_cell, int B, float* anchors_out, void* stream) {
    RN_REQUIRE(base_anchors_f32_dev && anchors_out, "NULL pointer");
    RN_REQUIRE(B >= 1, "B must be >= 1");
    RN_REQUIRE(rn_aligned16(anchors_out), "anchors_out must be 16-byte aligned");
    RnLevels lv;
    int rc = rn_make_levels(&lv, level_hw, level_stride, num_levels, anchors_per_cell);
    if (rc) return rc;
    const int N = lv.start[num_levels];
    if (N == 0) return RN_OK;
    k_anchors_f32<<<stream_grid((long long)N * B), 256, 0, (cudaStream_t)stream>>>(lv, base_anchors_f32_dev, N, B, anchors_out);
    return rn_check_launch("rn_anchors_f32");
}

extern "C" int rn_regress_boxes(const float* boxes, const float* deltas, long long R,
                                const float* mean4, const float* std4, float* out, void* stream) {
    RN_REQUIRE(boxes && deltas && out && mean4 && std4, "NULL pointer");
    RN_REQUIRE(R >= 0, "negative row count");
    if (R == 0) return RN_OK;
    RN_REQUIRE(rn_aligned16(boxes) && rn_aligned16(deltas) && rn_aligned16(out), "tensors must be 16-byte aligned");
    Norm4 nm;
    for (int i = 0; i < 4; ++i) { nm.mean[i] = mean4[i]; nm.std[i] = std4[i]; }
    k_regress_boxes<<<stream_grid(R), 256, 0, (cudaStream_t)stream>>>(boxes, deltas, R, nm, out);
    return rn_check_launch("rn_regress_boxes");
}

extern "C" int rn_clip_boxes(const float* boxes, long long R, float width, float height, float* out, void* stream) {
    RN_REQUIRE(boxes && out, "NULL pointer");
    RN_REQUIRE(R >= 0, "negative row count");
    if (R == 0) return RN_OK;
    RN_REQUIRE(rn_aligned16(boxes) && rn_aligned16(out), "tensors must be 16-byte aligned");
    k_clip_boxes<<<stream_grid(R), 256, 0, (cudaStream_t)stream>>>(boxes, R, width, height, out);
    return rn_check_launch("rn_clip_boxes");
}

// ------------------------------------------------------------------------------------------------
// N3  host post-step of the reference (RetinaNet.py:366-377) as an epilogue on the detections:
//   boxes /= image_scale   (fp32 division by the page's resize scale, as numpy does on the float32 array)
//   visible = number of leading detections before the first score < min_score (the reference walks the
//             score-sorted list and breaks there; padding entries have score -1)
// One CTA per page; detections are at most a few hundred rows.
// ------------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) k_rescale_cut(const float* boxes, const float* scores, const float* scale, int M,
                                                     float min_score, float* boxes_out, int* count_out) {
    __shared__ int s_first;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) s_first = M;
    __syncthreads();
    const float sc = __ldg(scale + b);
    int first = M;
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const size_t at = (size_t)b * M + m;
        float4 v = __ldg(reinterpret_cast<const float4*>(boxes) + at);
        v.x = __fdiv_rn(v.x, sc); v.y = __fdiv_rn(v.y, sc); v.z = __fdiv_rn(v.z, sc); v.w = __fdiv_rn(v.w, sc);
        reinterpret_cast<float4*>(boxes_out)[at] = v;
        if (__ldg(scores + at) < min_score && m < first) first = m;     // NaN < x is false, as in the reference's `if`
    }
    if (first < M) atomicMin(&s_first, first);
    __syncthreads();
    if (threadIdx.x == 0) count_out[b] = s_first;
}

}  // namespace

extern "C" int rn_rescale_cut(const float* boxes, const float* scores, const float* image_scale_dev, int B, int M,
                              float min_score, float* boxes_out, int* count_out, void* stream) {
    RN_REQUIRE(boxes && scores && image_scale_dev && boxes_out && count_out, "NULL pointer");
    RN_REQUIRE(B >= 0 && M >= 0, "negative size");
    if (B == 0) return RN_OK;
    RN_REQUIRE(B <= 2147483647 && M <= (1 << 24), "size out of range");
    RN_REQUIRE(rn_aligned16(boxes) && rn_aligned16(boxes_out), "boxes must be 16-byte aligned");
    k_rescale_cut<<<B, 256, 0, (cudaStream_t)stream>>>(boxes, scores, image_scale_dev, M, min_score, boxes_out, count_out);
    return rn_check_launch("rn_rescale_cut");
}

// ------------------------------------------------------------------------------------------------
// `other` tensors of FilterDetections follow the selected anchors (model/layers.py:247, :255): tf.gather(o, indices) for the
// kept detections, tf.pad(..., constant_values=-1) for the rest.  One 4-byte element per thread: rows are (B, N, D) with any
// trailing size D.
// ------------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) k_gather_other(const float* other, const int* idx, long long total, int N, int M, int D,
                                                      float* out) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / D;                        // (b, m)
        const int d = (int)(e - row * D);
        const long long b = row / M;
        const int n = __ldg(idx + row);
        out[e] = n >= 0 ? __ldg(other + ((size_t)b * N + n) * D + d) : -1.0f;
    }
}

}  // namespace

extern "C" int rn_gather_other(const float* other, const int* indices, int B, long long N, int M, int D, float* out, void* stream) {
    RN_REQUIRE(B >= 0 && N >= 0 && M >= 0 && D >= 1, "bad shape");
    const long long total = (long long)B * M * D;
    if (total == 0) return RN_OK;
    RN_REQUIRE(other && indices && out, "NULL pointer");
    RN_REQUIRE(N < (1ll << 31), "N too large");
    const int blocks = (int)((total + 255) / 256 < (long long)RN_NUM_SMS * 8 ? (total + 255) / 256 : (long long)RN_NUM_SMS * 8);
    k_gather_other<<<blocks, 256, 0, (cudaStream_t)stream>>>(other, indices, total, (int)N, M, D, out);
    return rn_check_launch("rn_gather_other");
}
