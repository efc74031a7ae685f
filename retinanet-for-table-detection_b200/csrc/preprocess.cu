// N4: the page-image producer (reference DetectTablesUtils.py:183-262, preProcessTrainValImages / preProcessSampleImages):
//
//     gray   = cv2.cvtColor(img, COLOR_BGR2GRAY)
//     binary = cv2.adaptiveThreshold(gray, 255, ADAPTIVE_THRESH_GAUSSIAN_C, THRESH_BINARY, 11, 2)
//     page   = imwrite(merge(distanceTransform(binary, DIST_L2, 5), distanceTransform(binary, DIST_L1, 5),
//                            distanceTransform(binary, DIST_C, 5)))                     -> uint8 (H, W, 3)
//
// as three kernels over a batch of pages (OpenCV's arithmetic restated: see oracle/preprocess_np.py, pinned against cv2):
//
//   k_gray_threshold   one pass over the BGR bytes: 15-bit fixed-point grey, the separable 11-tap float Gaussian in shared
//                      memory IN OPENCV'S EVALUATION ORDER (row pass left to right with fused multiply-adds, column pass
//                      symmetric with fused multiply-adds: a blurred mean that lands on x.5 flips a pixel, so the order is
//                      part of the result), round-half-even mean, threshold.  HBM: 3 B/pixel read, 1 B/pixel written.
//   k_row_distance     per page row, one warp: g(x) = distance to the nearest zero pixel of the row (ballot + clz / ffs
//                      scans forward and backward).  1 B/pixel read, 2 B/pixel written.
//   k_distance_u8      per pixel: all three metrics are monotone in dx, so the nearest zero of a row is the only one that
//                      matters: DT(x, y) = min over y' of D(g(x, y'), |y - y'|), swept outward from the pixel's own row and
//                      stopped as soon as |y - y'| exceeds the best L1 distance found (every metric is >= |y - y'|) or 255
//                      (the output saturates).  The 5x5 "L2" of OpenCV is the CHAMFER distance (moves 1, 1.4f, 2.1969f) in
//                      closed form; L1 = dx + dy; C = max(dx, dy).  Output: round-half-even, saturated uint8, the bytes
//                      imwrite encodes (a page without any dark pixel comes out BLACK, as OpenCV's conversion of FLT_MAX does).  Reads g from L2 (2 B/pixel/row visited), writes 3 B/pixel.
#include "rn_common.cuh"

namespace {

constexpr int PT_W = 64, PT_H = 32;                         // output tile of k_gray_threshold
constexpr int PT_R = 5;                                     // kernel radius (11 taps)
constexpr int PT_THREADS = 256;
constexpr unsigned short NO_ZERO = 0xFFFF;

// cv2.getGaussianKernel(11, -1, CV_32F) (sigma = 0.3 * ((11 - 1) * 0.5 - 1) + 0.8 = 2.0), bit patterns
__constant__ unsigned c_gauss11[11] = {0x3c10612bu, 0x3cde5c35u, 0x3d855a85u, 0x3df92326u, 0x3e353f0fu, 0x3e4d6105u,
                                       0x3e353f0fu, 0x3df92326u, 0x3d855a85u, 0x3cde5c35u, 0x3c10612bu};

__global__ void __launch_bounds__(PT_THREADS) k_gray_threshold(const unsigned char* __restrict__ bgr, int H, int W,
                                                                unsigned char* __restrict__ binary) {
    // (odd row pitches: in the row pass the lanes of a warp walk DOWN a column of windows, so that neither its 14 window loads
    // nor its 4 stores hit one bank twice)
    __shared__ float s_gray[PT_H + 2 * PT_R][PT_W + 2 * PT_R + 3];     // grey levels of the tile + halo (replicated border)
    __shared__ float s_row[PT_H + 2 * PT_R][PT_W + 1];                 // row pass
    const int page = blockIdx.z;
    const unsigned char* src = bgr + (size_t)page * H * W * 3;
    unsigned char* dst = binary + (size_t)page * H * W;
    const int x0 = blockIdx.x * PT_W, y0 = blockIdx.y * PT_H;
    const int tid = threadIdx.x;
    float k[11];
#pragma unroll
    for (int j = 0; j < 11; ++j) k[j] = __uint_as_float(c_gauss11[j]);
    // grey = (B * 3735 + G * 19235 + R * 9798 + 2^14) >> 15   (OpenCV's 8-bit BGR2GRAY)
    for (int i = tid; i < (PT_H + 2 * PT_R) * (PT_W + 2 * PT_R); i += PT_THREADS) {
        const int ly = i / (PT_W + 2 * PT_R), lx = i - ly * (PT_W + 2 * PT_R);
        const int y = min(max(y0 + ly - PT_R, 0), H - 1), x = min(max(x0 + lx - PT_R, 0), W - 1);     // BORDER_REPLICATE
        const unsigned char* px = src + ((size_t)y * W + x) * 3;
        const int g = (px[0] * 3735 + px[1] * 19235 + px[2] * 9798 + (1 << 14)) >> 15;
        s_gray[ly][lx] = (float)g;
    }
    __syncthreads();
    // row pass, left to right: acc = k0 * p0; acc = fma(p_j, k_j, acc).  A thread produces 4 consecutive outputs of a row
    // from a 14-element window held in registers (14 shared-memory loads instead of 44).
    for (int i = tid; i < (PT_H + 2 * PT_R) * (PT_W / 4); i += PT_THREADS) {
        const int q = i / (PT_H + 2 * PT_R), ly = i - q * (PT_H + 2 * PT_R), lx = q * 4;
        float w[14];
#pragma unroll
        for (int j = 0; j < 14; ++j) w[j] = s_gray[ly][lx + j];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float acc = k[0] * w[o];
#pragma unroll
            for (int j = 1; j < 11; ++j) acc = fmaf(w[o + j], k[j], acc);
            s_row[ly][lx + o] = acc;
        }
    }
    __syncthreads();
    // column pass, symmetric: acc = k5 * c; acc = fma(p_{+j} + p_{-j}, k_{5+j}, acc); then the threshold.  A thread produces
    // 4 vertically adjacent outputs of a column from a 14-element window.
    for (int i = tid; i < (PT_H / 4) * PT_W; i += PT_THREADS) {
        const int lyb = (i / PT_W) * 4, lx = i - (i / PT_W) * PT_W;
        const int x = x0 + lx;
        float w[14];
#pragma unroll
        for (int j = 0; j < 14; ++j) w[j] = s_row[lyb + j][lx];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int y = y0 + lyb + o;
            float acc = k[5] * w[o + PT_R];
#pragma unroll
            for (int j = 1; j <= PT_R; ++j) acc = fmaf(w[o + PT_R + j] + w[o + PT_R - j], k[5 + j], acc);
            if (y < H && x < W) {
                const int mean = min(max(__float2int_rn(acc), 0), 255);     // saturate_cast<uchar>(cvRound(.))
                const int g = (int)s_gray[lyb + o + PT_R][lx + PT_R];
                dst[(size_t)y * W + x] = (g - mean > -2) ? 255 : 0;         // THRESH_BINARY with delta 2
            }
        }
    }
}

// one warp per page row
__global__ void __launch_bounds__(256) k_row_distance(const unsigned char* __restrict__ binary, int rows_total, int H, int W,
                                                       unsigned short* __restrict__ g, int* __restrict__ has_zero) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows_total) return;
    const unsigned char* src = binary + (size_t)row * W;
    unsigned short* out = g + (size_t)row * W;
    int carry = -(1 << 20);                                 // position of the last zero seen so far (none yet)
    for (int c0 = 0; c0 < W; c0 += 32) {
        const int x = c0 + lane;
        const unsigned m = __ballot_sync(0xffffffffu, x < W && src[x] == 0);
        const unsigned upto = m & (0xffffffffu >> (31 - lane));            // zeros at or before this lane
        const int last = upto ? c0 + 31 - __clz(upto) : carry;
        if (x < W) out[x] = (unsigned short)min(x - last, (int)NO_ZERO);
        if (m) carry = c0 + 31 - __clz(m);
    }
    if (carry >= 0 && lane == 0) has_zero[row / H] = 1;     // the page has ink (benign race: every writer stores 1)
    carry = 1 << 20;                                        // position of the next zero to the right (none yet)
    for (int c0 = ((W - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
        const int x = c0 + lane;
        const unsigned m = __ballot_sync(0xffffffffu, x < W && src[x] == 0);
        const unsigned from = m & (0xffffffffu << lane);                   // zeros at or after this lane
        const int next = from ? c0 + __ffs(from) - 1 : carry;
        if (x < W) out[x] = (unsigned short)min((int)out[x], min(next - x, (int)NO_ZERO));
        if (m) carry = c0 + __ffs(m) - 1;
    }
}

// candidates of one row at vertical distance dy with in-row distance gx
__device__ __forceinline__ void dt_candidates(int gx, int dy, float& l2, int& l1, int& cc) {
    const int M = max(gx, dy), m = min(gx, dy);
    l1 = min(l1, gx + dy);
    cc = min(cc, M);
    // the chamfer distance is >= the chessboard distance M (equal only when m == 0), so a candidate whose M is not below the
    // best value so far cannot lower it: most candidates of a sweep skip the float arithmetic (the kernel is issue-bound:
    // ~300 instructions per pixel; staging g in shared memory instead of reading it from L2 changed nothing, 526 vs 461 us)
    if ((float)M < l2) {
        const float c = 2.1969f, b = 1.4f;                  // OpenCV's 5x5 DIST_L2 mask: moves 1, 1.4f, 2.1969f
        const float d = (M >= 2 * m) ? c * (float)m + (float)(M - 2 * m) : c * (float)(M - m) + b * (float)(2 * m - M);
        l2 = fminf(l2, d);
    }
}

__global__ void __launch_bounds__(256) k_distance_u8(const unsigned short* __restrict__ g, const int* __restrict__ has_zero,
                                                      int H, int W, unsigned char* __restrict__ out) {
    const int page = blockIdx.z;
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= W || y >= H) return;
    unsigned char* o = out + ((size_t)page * H * W + (size_t)y * W + x) * 3;
    if (!has_zero[page]) {
        // no dark pixel on the whole page: OpenCV leaves FLT_MAX, whose 8-bit conversion in imwrite (cvRound overflows to
        // INT_MIN) is 0 -- a black page
        o[0] = 0; o[1] = 0; o[2] = 0;
        return;
    }
    const unsigned short* gp = g + (size_t)page * H * W + x;
    float l2 = 1e9f;
    int l1 = 1 << 28, cc = 1 << 28;
    // every metric is >= dy, so rows further away than the best L1 distance cannot matter; beyond 255 the output saturates
    // anyway.  (A/B: fetching four row pairs per round to overlap their latencies was slower, 596 vs 456 us per 16 pages:
    // on a text page most pixels are done after two or three rows.)
    for (int dy = 0; dy <= 255 && dy < l1; ++dy) {
        if (y - dy >= 0) {
            const int gx = gp[(size_t)(y - dy) * W];
            if (gx != NO_ZERO) dt_candidates(gx, dy, l2, l1, cc);
        }
        if (dy > 0 && y + dy < H) {
            const int gx = gp[(size_t)(y + dy) * W];
            if (gx != NO_ZERO) dt_candidates(gx, dy, l2, l1, cc);
        }
    }
    o[0] = (unsigned char)min(__float2int_rn(fminf(l2, 1000.0f)), 255);     // saturate_cast<uchar>(cvRound(.)), merge order b, g, r
    o[1] = (unsigned char)min(l1, 255);
    o[2] = (unsigned char)min(cc, 255);
}

size_t pp_align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t rn_preprocess_workspace_bytes(int B, int H, int W) {
    if (B < 1 || H < 1 || W < 1) return 0;
    const size_t px = (size_t)B * H * W;
    return pp_align256(px) + pp_align256(px * sizeof(unsigned short)) + pp_align256(sizeof(int) * (size_t)B);
}

extern "C" int rn_preprocess_pages(const unsigned char* bgr_dev, int B, int H, int W, unsigned char* out_dev,
                                   unsigned char* binary_out_dev, void* workspace, size_t workspace_bytes, void* stream) {
    RN_REQUIRE(bgr_dev && out_dev && workspace, "NULL pointer");
    RN_REQUIRE(B >= 1 && B <= 65535 && H >= 1 && W >= 1, "bad shape");
    RN_REQUIRE(W < 65535 && (long long)H * W < (1ll << 31), "page too large (W < 65535, H * W < 2^31)");
    if (workspace_bytes < rn_preprocess_workspace_bytes(B, H, W)) return rn_fail(RN_ERR_WORKSPACE, "preprocess workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t px = (size_t)B * H * W;
    unsigned char* binary = binary_out_dev ? binary_out_dev : reinterpret_cast<unsigned char*>(workspace);
    unsigned short* g = reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(workspace) + pp_align256(px));
    int* has_zero = reinterpret_cast<int*>(reinterpret_cast<char*>(g) + pp_align256(px * sizeof(unsigned short)));
    cudaError_t e = cudaMemsetAsync(has_zero, 0, sizeof(int) * (size_t)B, s);
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
    const dim3 tiles((unsigned)((W + PT_W - 1) / PT_W), (unsigned)((H + PT_H - 1) / PT_H), (unsigned)B);
    RN_REQUIRE(tiles.y <= 65535, "page too tall");
    k_gray_threshold<<<tiles, PT_THREADS, 0, s>>>(bgr_dev, H, W, binary);
    int rc = rn_check_launch("k_gray_threshold");
    if (rc) return rc;
    const long long rows = (long long)B * H;
    RN_REQUIRE(rows < (1ll << 31), "too many rows");
    k_row_distance<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(binary, (int)rows, H, W, g, has_zero);
    rc = rn_check_launch("k_row_distance");
    if (rc) return rc;
    const dim3 grid((unsigned)((W + 63) / 64), (unsigned)((H + 3) / 4), (unsigned)B);
    RN_REQUIRE(grid.y <= 65535, "page too tall");
    k_distance_u8<<<grid, 256, 0, s>>>(g, has_zero, H, W, out_dev);
    return rn_check_launch("k_distance_u8");
}
