// K1: fused anchor generation + IoU / first-max argmax matching + labels + regression targets.
//
// One launch covers the whole batch: grid = (anchor tiles, pages).  A CTA owns 256 consecutive
// anchors of one page; the page's GT tables are staged in shared memory in chunks, after culling
// the ones that cannot intersect the tile's bounding box (warp-shuffle min/max reductions), in GT
// order so the reference's "first maximum wins" tie rule is preserved.  Outputs are staged in
// shared memory and written with 128-bit coalesced stores.
//
// Precision pipeline (must match numpy bit for bit, see oracle/anchors_np.py):
//   anchors fp64 (base + centre, one rounding)  ->  IoU fp64  ->  rounded to fp32  ->
//   strict-greater argmax on fp32  ->  fp32 threshold compares  ->  bbox_transform fp64 -> fp32.
// The file is compiled with -fmad=false so no multiply-add is contracted.
//
// fp64 divisions are the expensive part (B200 issues fp64 at half rate and a correctly rounded divide is
// ~20 instructions).  Every quotient on this path is only ever consumed after rounding to fp32, so each is
// first computed with a fast reciprocal (hardware approximation + 2 Newton steps, relative error < 2^-45)
// and accepted when its fp32 rounding cannot depend on that error: the 29 mantissa bits dropped by the
// fp64->fp32 conversion must be further than 2^17 fp64-ulps (2^-35 relative; the error budget, including the
// base-box width substitution of the regression path, is 2^-37) from the rounding midpoint.  Otherwise
// (probability ~5e-4 per value) the exact IEEE expression of the reference is evaluated.  The result is
// bit-identical to always dividing.
#include "rn_common.cuh"
#include <atomic>

namespace {

constexpr int K1_THREADS = 256;
constexpr int K1_WARPS = K1_THREADS / 32;

struct K1Params {
    RnLevels lv;
    const double* base;      // (L, A, 4)
    const double* anchors;   // (N, 4) or nullptr
    int N;
    const double* gt;        // (B, Gmax, 4)
    const int* gt_labels;    // (B, Gmax)
    const int* gt_count;     // (B)
    const int* img_hw;       // (B, 2) or nullptr
    const int* page_order;   // (B) or nullptr: the page CTA row y works on
    int Gmax, C;
    float neg, pos;
    float* reg;              // (B, N, 5)
    float* lab;              // (B, N, C+1)
    int* argmax;             // (B, N) or nullptr
    int* npos;               // (B) or nullptr
    float* npos_total;       // 1 float or nullptr
    int vec_ok;
    int sparse_reg;          // k_anchor_targets_tiles32<.., SPARSE>: only the regression rows of state == 1 anchors are written
    double max_coord;        // upper bound of any anchor-centre coordinate (for the reciprocal table, see wrapper)
};

// cooperative store of `len` floats starting at element `start` of `base`; 128-bit where aligned
template <typename Gen>
__device__ __forceinline__ void store_range(float* base, long long start, int len, bool vec_ok, int nthreads, Gen gen) {
    float* p = base + start;
    int head = vec_ok ? (int)((4 - (start & 3)) & 3) : len;
    if (head > len) head = len;
#pragma unroll 1
    for (int i = threadIdx.x; i < head; i += nthreads) p[i] = gen(i);
    const int nvec = (len - head) >> 2;
#pragma unroll 1
    for (int v = threadIdx.x; v < nvec; v += nthreads) {
        const int i = head + 4 * v;
        rn_stg_stream4(p + i, make_float4(gen(i), gen(i + 1), gen(i + 2), gen(i + 3)));
    }
#pragma unroll 1
    for (int i = head + 4 * nvec + threadIdx.x; i < len; i += nthreads) p[i] = gen(i);
}

// labels rows of C+1 floats (one-hot + state) for `cnt` anchors starting at row `row0`, any C
__device__ __noinline__ void store_labels_generic(float* lab, long long row0, int cnt, int C, bool vec_ok, int nthreads,
                                                  const float* s_state, const int* s_hot) {
    const int CW = C + 1;
    store_range(lab, row0 * CW, cnt * CW, vec_ok, nthreads, [&](int i) {
        const int r = i / CW, c = i - r * CW;
        return c == C ? s_state[r] : (c == s_hot[r] ? 1.0f : 0.0f);
    });
}

__device__ __forceinline__ void make_anchor(const RnLevels& lv, const double* base, int n,
                                            double& x1, double& y1, double& x2, double& y2) {
    int level, cx, cy, a;
    rn_locate(lv, n, level, cx, cy, a);
    const double* b = base + ((size_t)level * lv.anchors_per_cell + a) * 4;
    const double sx = ((double)cx + 0.5) * (double)lv.stride[level];   // exact in fp64
    const double sy = ((double)cy + 0.5) * (double)lv.stride[level];
    x1 = __ldg(b + 0) + sx;
    y1 = __ldg(b + 1) + sy;
    x2 = __ldg(b + 2) + sx;
    y2 = __ldg(b + 3) + sy;
}

// reciprocal with relative error < 2^-39 for positive normal x (garbage in -> rejected by the check below).
// rcp.approx.ftz.f64 has the upper 20 mantissa bits right (|1 - x r0| <= 2^-20); one Newton step squares that:
// x r1 = 1 - e^2, e^2 <= 2^-40, plus two fp64 roundings.  f32_rounding_safe() tolerates 2^-36.
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// v approximates (to < 2^-38 relative) a value that the reference rounds to fp32: true when the fp32
// rounding of v is guaranteed to equal the fp32 rounding of the exact value
__device__ __forceinline__ bool f32_rounding_safe(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    // biased exponent in [900, 1140] (fp32 normals need >= 897): one unsigned range compare
    const bool exp_ok = ((hi & 0x7ff00000u) - (900u << 20)) <= (240u << 20);
    // the 29 bits dropped by fp64 -> fp32 must be further than 2^17 from the rounding midpoint 2^28
    const bool far = (((lo & 0x1fffffffu) - ((1u << 28) - (1u << 17))) > (1u << 18));
    return exp_ok && far;
}

// min / max of ordered (non-NaN) doubles without the NaN fix-up of fmin / fmax
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }

// The exact IEEE expressions of the reference.  Rarely executed and deliberately NOT inlined: a correctly
// rounded fp64 divide is ~100 SASS lines per site, and inlining it at every use made the kernels larger
// than the instruction cache (ncu: 40 % of stall samples were "no instruction").
__device__ __noinline__ float reg_target_exact(double d, double len) { return (float)((d / len) / 0.2); }
__device__ __noinline__ float iou_exact(double inter, double uni) { return (float)(inter / uni); }

// one regression target: ((g - a) / len) / 0.2 rounded to fp32 (model/anchors.py:300-311)
__device__ __forceinline__ float reg_target(double g, double a, double len, double rlen) {
    const double d = g - a;
    const double approx = (d * rlen) * 5.0;
    if (d != 0.0 && f32_rounding_safe(approx)) return (float)approx;
    return reg_target_exact(d, len);                          // also the exact (signed) zero
}

// Regression targets with the reciprocal pre-multiplied by 5 (r5 ~ 5/len to < 2^-40 relative; d == 0 gives exponent 0 ->
// exact path, signed zero) and IoUs:
// two quotients at a time with ONE rarely taken branch behind both conversions instead of a branch (a convergence region
// the scheduler cannot move instructions across) after each: the two dependency chains -- subtract, multiply, mantissa
// check, convert -- overlap.  Same values, same fall-back, bit for bit (profiles/sweep_k1.py: 55.2 -> 53.5 us, equal checksums).
__device__ __forceinline__ void reg_target5_pair(double ga, double aa, double gb, double ab, double len, double r5, float& ta, float& tb) {
    const double da = ga - aa, db = gb - ab;
    const double pa = da * r5, pb = db * r5;
    const bool oka = f32_rounding_safe(pa), okb = f32_rounding_safe(pb);
    ta = (float)pa; tb = (float)pb;
    if (!(oka && okb)) {
        if (!oka) ta = reg_target_exact(da, len);
        if (!okb) tb = reg_target_exact(db, len);
    }
}
__device__ __forceinline__ void iou_pair(double i0, double u0, double i1, double u1, float& iou0, float& iou1) {
    const double q0 = i0 * rcp_fast(u0), q1 = i1 * rcp_fast(u1);
    const bool ok0 = f32_rounding_safe(q0), ok1 = f32_rounding_safe(q1);
    iou0 = (float)q0; iou1 = (float)q1;
    if (!(ok0 && ok1)) {
        if (!ok0) iou0 = iou_exact(i0, u0);
        if (!ok1) iou1 = iou_exact(i1, u1);
    }
}

template <bool EXPLICIT>
__global__ void __launch_bounds__(K1_THREADS) k_anchor_targets(const K1Params p) {
    __shared__ double s_gx1[K1_THREADS], s_gy1[K1_THREADS], s_gx2[K1_THREADS], s_gy2[K1_THREADS], s_ga[K1_THREADS];
    __shared__ int s_gidx[K1_THREADS];
    __shared__ float s_red[K1_WARPS][4];
    __shared__ int s_wcount[K1_WARPS];
    __shared__ float s_reg[K1_THREADS * 5];
    __shared__ float s_state[K1_THREADS];
    __shared__ int s_hot[K1_THREADS];
    __shared__ int s_npos;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = p.page_order ? __ldg(p.page_order + blockIdx.y) : (int)blockIdx.y;   // (heaviest pages first, see the wrapper)
    const int n0 = blockIdx.x * K1_THREADS;
    const int cnt = min(K1_THREADS, p.N - n0);
    const int n = n0 + tid;
    const bool valid = tid < cnt;
    int G = p.gt_count[b];
    G = max(0, min(G, p.Gmax));

    // ---- this thread's anchor ---------------------------------------------------------------
    double ax1 = 0, ay1 = 0, ax2 = 0, ay2 = 0;
    if (valid) {
        if (EXPLICIT) {
            const double2* ap = reinterpret_cast<const double2*>(p.anchors + (size_t)n * 4);
            const double2 lo = __ldg(ap), hi = __ldg(ap + 1);
            ax1 = lo.x; ay1 = lo.y; ax2 = hi.x; ay2 = hi.y;
        } else {
            make_anchor(p.lv, p.base, n, ax1, ay1, ax2, ay2);
        }
    }
    const double aw = ax2 - ax1, ah = ay2 - ay1;
    const double area_a = aw * ah;

    // ---- tile bounding box: fp32 with outward rounding (conservative), warp shuffles then smem ----
    {
        const float inf = __int_as_float(0x7f800000);
        const float mnx = rn_warp_min(valid ? __double2float_rd(ax1) : inf), mny = rn_warp_min(valid ? __double2float_rd(ay1) : inf);
        const float mxx = rn_warp_max(valid ? __double2float_ru(ax2) : -inf), mxy = rn_warp_max(valid ? __double2float_ru(ay2) : -inf);
        if (lane == 0) { s_red[warp][0] = mnx; s_red[warp][1] = mny; s_red[warp][2] = mxx; s_red[warp][3] = mxy; }
        if (tid == 0) s_npos = 0;
    }
    __syncthreads();
    float fx1 = s_red[0][0], fy1 = s_red[0][1], fx2 = s_red[0][2], fy2 = s_red[0][3];
#pragma unroll
    for (int w = 1; w < K1_WARPS; ++w) {
        fx1 = fminf(fx1, s_red[w][0]); fy1 = fminf(fy1, s_red[w][1]);
        fx2 = fmaxf(fx2, s_red[w][2]); fy2 = fmaxf(fy2, s_red[w][3]);
    }
    const double tx1 = (double)fx1, ty1 = (double)fy1, tx2 = (double)fx2, ty2 = (double)fy2;

    // ---- IoU / argmax over the GT tables that can touch this tile ----------------------------
    float best = 0.0f;     // an all-zero IoU row has argmax 0 (numpy first-max)
    int arg = 0;
    const double* gtb = p.gt + (size_t)b * p.Gmax * 4;
    for (int g0 = 0; g0 < G; g0 += K1_THREADS) {
        const int j = g0 + tid;
        bool keep = false;
        double gx1 = 0, gy1 = 0, gx2 = 0, gy2 = 0;
        if (j < G) {
            gx1 = __ldg(gtb + 4 * j); gy1 = __ldg(gtb + 4 * j + 1);
            gx2 = __ldg(gtb + 4 * j + 2); gy2 = __ldg(gtb + 4 * j + 3);
            // a GT that does not reach into the tile box has zero intersection with every anchor in it
            keep = (gx2 > tx1) && (gx1 < tx2) && (gy2 > ty1) && (gy1 < ty2);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        __syncthreads();                       // previous chunk fully consumed
        if (lane == 0) s_wcount[warp] = __popc(bal);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < K1_WARPS; ++w) { const int c = s_wcount[w]; if (w < warp) off += c; total += c; }
        if (keep) {
            const int pos = off + __popc(bal & ((1u << lane) - 1u));   // stable: GT order kept
            s_gx1[pos] = gx1; s_gy1[pos] = gy1; s_gx2[pos] = gx2; s_gy2[pos] = gy2;
            s_ga[pos] = (gx2 - gx1) * (gy2 - gy1);
            s_gidx[pos] = j;
        }
        __syncthreads();
        if (valid) {
            for (int m = 0; m < total; ++m) {
                const double iw = fmin(ax2, s_gx2[m]) - fmax(ax1, s_gx1[m]);
                const double ih = fmin(ay2, s_gy2[m]) - fmax(ay1, s_gy1[m]);
                if (iw > 0.0 && ih > 0.0) {
                    const double inter = iw * ih;
                    const double uni = area_a + s_ga[m] - inter;
                    const double q = inter * rcp_fast(uni);
                    const float iou = f32_rounding_safe(q) ? (float)q : iou_exact(inter, uni);   // == (float)(inter / uni)
                    if (iou > best) { best = iou; arg = s_gidx[m]; }
                }
            }
        }
    }

    // ---- state, one-hot class, regression targets, border rule ------------------------------
    float state = 0.0f;
    int hot = -1;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    if (valid) {
        if (G > 0) {
            const bool is_pos = best >= p.pos;
            const bool is_ign = (best > p.neg) && !is_pos;
            state = is_pos ? 1.0f : (is_ign ? -1.0f : 0.0f);
            if (is_pos) hot = __ldg(p.gt_labels + (size_t)b * p.Gmax + arg);
            const double* g = gtb + 4 * (size_t)arg;
            const double rw = rcp_fast(aw), rh = rcp_fast(ah);
            t0 = reg_target(__ldg(g + 0), ax1, aw, rw);
            t1 = reg_target(__ldg(g + 1), ay1, ah, rh);
            t2 = reg_target(__ldg(g + 2), ax2, aw, rw);
            t3 = reg_target(__ldg(g + 3), ay2, ah, rh);
        }
        if (p.img_hw) {
            const double ccx = (ax1 + ax2) / 2.0, ccy = (ay1 + ay2) / 2.0;
            if (ccx >= (double)p.img_hw[2 * b + 1] || ccy >= (double)p.img_hw[2 * b]) state = -1.0f;
        }
        s_reg[tid * 5 + 0] = t0; s_reg[tid * 5 + 1] = t1; s_reg[tid * 5 + 2] = t2; s_reg[tid * 5 + 3] = t3;
        s_reg[tid * 5 + 4] = state;
        s_state[tid] = state;
        s_hot[tid] = hot;
        if (p.argmax) p.argmax[(size_t)b * p.N + n] = arg;
    }
    if (p.npos || p.npos_total) {
        const unsigned pb = __ballot_sync(0xffffffffu, valid && state == 1.0f);
        if (lane == 0 && pb) atomicAdd(&s_npos, __popc(pb));
    }
    __syncthreads();
    if (tid == 0 && s_npos) {
        if (p.npos) atomicAdd(p.npos + b, s_npos);
        if (p.npos_total) atomicAdd(p.npos_total, (float)s_npos);   // integer-valued: exact, order-independent
    }

    // ---- coalesced write-out ------------------------------------------------------------------
    const long long row0 = (long long)b * p.N + n0;
    store_range(p.reg, row0 * 5, cnt * 5, p.vec_ok != 0, K1_THREADS, [&](int i) { return s_reg[i]; });
    if (p.C == 1) {
        if (valid) {
            float2 v = make_float2(hot == 0 ? 1.0f : 0.0f, state);
            reinterpret_cast<float2*>(p.lab)[row0 + tid] = v;
        }
    } else {
        store_labels_generic(p.lab, row0, cnt, p.C, p.vec_ok != 0, K1_THREADS, s_state, s_hot);
    }
}

// One staged row -> global memory through the TMA engine.  `dst` and `src` are 16-byte aligned and in phase; floats
// [lo, hi) of the staged row are valid (lo < 4).  Whole 16-byte chunks go in one bulk copy, the partial chunks at the
// ends in one 16-byte copy each with a byte mask (sm_100 cp_mask: bit i = byte i of the chunk).
__device__ __forceinline__ void rn_bulk_store_row(float* dst, const float* src, int lo, int hi) {
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(src);
    auto mask_of = [](int a, int b) { return (unsigned short)(((1u << (4 * b)) - 1u) & ~((1u << (4 * a)) - 1u)); };   // floats [a, b) of a chunk
    const int first_full = lo ? 1 : 0, last_full = hi >> 2;     // chunks [first_full, last_full) are whole
    if (last_full < first_full || (last_full == 0 && lo)) {     // everything inside chunk 0
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.cp_mask [%0], [%1], 16, %2;"
                     :: "l"(dst), "r"(s0), "h"(mask_of(lo, hi)) : "memory");
        return;
    }
    if (lo)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.cp_mask [%0], [%1], 16, %2;"
                     :: "l"(dst), "r"(s0), "h"(mask_of(lo, 4)) : "memory");
    if (last_full > first_full)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(dst + 4 * first_full), "r"(s0 + 16u * first_full), "r"(16u * (unsigned)(last_full - first_full)) : "memory");
    if (hi & 3)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.cp_mask [%0], [%1], 16, %2;"
                     :: "l"(dst + 4 * last_full), "r"(s0 + 16u * last_full), "h"(mask_of(0, hi & 3)) : "memory");
}

// ------------------------------------------------------------------------------------------------
// K1 for generated anchors: a CTA owns a 32-cell x KT_ROWS-row tile of one feature map, a WARP one anchor
// type of that tile, a THREAD one (anchor type, column) and walks the tile's rows.
//
// Everything that depends only on the column is computed once per KT_ROWS anchors (x extent, 5/width, the
// x half of every GT overlap test and the intersection width); everything that depends only on the warp is
// warp-uniform (base box, row geometry, the warp's exact bounding box -> one ballot culls 32 GT tables, and
// the surviving tables are walked with warp-uniform control flow).  The per-anchor arithmetic is the one of
// k_anchor_targets, operation for operation.  Results are staged in shared memory in the reference's
// anchor order (cell-major, anchor-minor; each tile row is one contiguous range) and written with 128-bit
// stores.
// ------------------------------------------------------------------------------------------------
constexpr int KT_ROWS = 4;             // the write-out's row decode assumes 4
// (A/B, profiles/sweep_k1.py: a CTA walking 2 / 3 / 4 vertically adjacent 4-row tiles -- tile decode, column geometry,
// staged GT tables and x targets shared by the passes, 13 % fewer instructions at 2 -- took 60.3 / 62.9 / 64.6 us
// against 56.6: the kernel is latency-, not issue-bound, and longer CTAs drain worse.  Requesting the GT chunk before
// the tile decode changed nothing.  A page with one small table costs 43 us per 16 pages, the heaviest 77 us:
// 3/4 of the time is the per-anchor work (targets, state, staging, write-out), not the matching; profiles/k1_page_cost.py.)
constexpr int KT_MAX_A = 24;
constexpr int KT_CHUNK = 256;          // GT tables staged per round

// dynamic shared memory of k_anchor_targets_tiles: staged regression rows, staged label rows, intersection heights
static size_t kt_dyn_smem(int A) {
    return (size_t)KT_ROWS * ((32 * A * 5 + 4) + (32 * A * 2 + 4)) * sizeof(float) + (size_t)A * KT_ROWS * 32 * sizeof(double);
}

struct K1Tiles {
    int tile_start[RN_MAX_LEVELS + 1];  // first tile of each level
    int tiles_x[RN_MAX_LEVELS];
    float inv_tiles_x[RN_MAX_LEVELS];
};

// k_anchor_targets_tiles32: a 1-D grid, page slot by page slot.  The first `coarse_pages` page slots are cut into CTAs of
// K32_XT x tiles (table `c`), the remaining ones -- the END of the launch -- into CTAs of one x tile (table `f`): three times
// as many CTAs of a third of the duration, so the launch's drain (the time between the last CTA's start and the last CTA's
// end, during which the GPU runs empty) shrinks with them.
struct K1Tiles32 {
    K1Tiles c, f;
    int coarse_pages;
};

// MAXA bounds the block size (32 * A threads) so that the register budget can be set per instantiation:
// <9, 3> is the RetinaNet default (288 threads, >= 3 CTAs per SM), <KT_MAX_A, 1> covers the rest
// Write-out of a tile's staged rows when the output tensors are not 16-byte aligned: scalar stores (rows are staged
// with the destination's phase, see the kernel).  Rare path, out of line so that it costs the kernel no registers.
__device__ __noinline__ void write_out_unaligned(float* reg, float* lab, int C, const float* s_reg, const float* s_lab,
                                                 const float* s_state, const int* s_hot, long long tile_row0, int row_anchors,
                                                 int cnt, int nrows, int A, int tid, int nthreads) {
    const int reg_stride = 32 * A * 5 + 4, lab_stride = 32 * A * 2 + 4;
    for (int r = 0; r < nrows; ++r) {
        const long long first = tile_row0 + (long long)r * row_anchors;
        const float* sr = s_reg + r * reg_stride + (int)((first * 5) & 3);
        for (int i = tid; i < cnt * 5; i += nthreads) reg[first * 5 + i] = sr[i];
        if (C == 1) {
            const float* sl = s_lab + r * lab_stride + (int)((first * 2) & 3);
            for (int i = tid; i < cnt * 2; i += nthreads) lab[first * 2 + i] = sl[i];
        } else {
            store_labels_generic(lab, first, cnt, C, false, nthreads, s_state + r * 32 * A, s_hot + r * 32 * A);
        }
    }
}

// Write-out of a tile's staged rows (shared by the tile kernels).  Each tile row is one contiguous anchor range, staged
// with the destination's 16-byte phase.
template <bool C1, bool NO_REG = false>   // NO_REG: the regression rows are not staged (sparse targets): label rows only
__device__ __forceinline__ void tile_write_out(const K1Params& p, const float* s_reg, const float* s_lab, const float* s_state,
                                               const int* s_hot, int b, int lstart, int cy0, int cx0, int W, int A, int ncols,
                                               int nrows, int reg_stride, int lab_stride, int tid, int nthreads) {
    const int cnt = ncols * A;
    const long long tile_row0 = (long long)b * p.N + lstart + ((long long)cy0 * W + cx0) * A;
    if (p.vec_ok) {
        // TMA bulk stores: ONE thread per (row, tensor) hands its staged row to the copy engine --
        // cp.async.bulk.global.shared::cta for the 16-byte aligned interior, the sm_100 .cp_mask form (a byte mask
        // inside one 16-byte chunk) for the partial chunks at the two ends -- instead of every thread looping
        // over LDS.128 / STG.128 pairs (that loop was ~15 % of the kernel's instructions).
        const int job = tid;                                // jobs [0, KT_ROWS): regression rows, [KT_ROWS, 2 KT_ROWS): label rows
        const bool lab_job = job >= KT_ROWS;
        const int r = lab_job ? job - KT_ROWS : job;
        if (job < 2 * KT_ROWS && r < nrows && !(lab_job && !C1) && !(NO_REG && !lab_job)) {
            const int per = lab_job ? 2 : 5;
            const long long start = (tile_row0 + (long long)r * W * A) * per;
            const int len = cnt * per, shift = (int)(start & 3);
            const float* src = lab_job ? s_lab + r * lab_stride : s_reg + r * reg_stride;
            float* dst = (lab_job ? p.lab : p.reg) + (start - shift);
            const int end = shift + len;                    // staged floats [shift, end) are valid
            rn_bulk_store_row(dst, src, shift, end);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the CTA's shared memory may go away after this
        }
        if (!C1) {
            // C > 1: a label row is C + 1 floats, all zero except the one-hot entry of a positive anchor and a non-zero
            // state (~1 % of the anchors).  So the tile rows' label ranges are zero-filled with 128-bit stores -- no
            // per-element row / column arithmetic -- and, after a barrier, the few non-zero entries are written.
            const int CW = p.C + 1;
            for (int r2 = 0; r2 < nrows; ++r2) {
                const long long start = (tile_row0 + (long long)r2 * W * A) * CW;
                const long long len = (long long)cnt * CW;
                float* dst = p.lab + start;
                int head = (int)((4 - (start & 3)) & 3);
                if (head > len) head = (int)len;
                if (tid < head) dst[tid] = 0.0f;
                const long long nvec = (len - head) >> 2;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                for (long long v = tid; v < nvec; v += nthreads) rn_stg_stream4(dst + head + 4 * v, z);
                const long long done = head + 4 * nvec;
                if (tid < (int)(len - done)) dst[done + tid] = 0.0f;
            }
            __syncthreads();                                // orders the fix-ups below after the zero-fill (same CTA)
            for (int j = tid; j < nrows * cnt; j += nthreads) {
                const int r2 = j / cnt, k = j - r2 * cnt;
                const float st = s_state[r2 * 32 * A + k];
                const int hot = s_hot[r2 * 32 * A + k];
                float* rowp = p.lab + (tile_row0 + (long long)r2 * W * A + k) * CW;
                if (st != 0.0f) rowp[p.C] = st;
                if (hot >= 0) rowp[hot] = 1.0f;
            }
        }
        return;
    }
    // unaligned output tensors (not 16-byte aligned: never the case for framework allocations): plain stores, out of line
    write_out_unaligned(p.reg, p.lab, p.C, s_reg, s_lab, s_state, s_hot, tile_row0, W * A, cnt, nrows, A, tid, nthreads);
}

// C1: one class (the table-detection configuration) -- specialised so that the generic label path costs the common
// instantiation no registers
template <int MAXA, int MINB, bool C1, bool AM>   // AM: the argmax tensor is wanted
__global__ void __launch_bounds__(32 * MAXA, MINB) k_anchor_targets_tiles(const K1Params p, const K1Tiles tl) {
    extern __shared__ __align__(16) float s_dyn[];
    __shared__ double s_gx1[KT_CHUNK], s_gy1[KT_CHUNK], s_gx2[KT_CHUNK], s_gy2[KT_CHUNK], s_ga[KT_CHUNK];
    __shared__ int s_glab[KT_CHUNK];
    __shared__ double s_row[KT_MAX_A][KT_ROWS][3];          // per (anchor type, tile row): y1, y2, height
    __shared__ int s_npos;

    const int tid = threadIdx.x, lane = tid & 31, a = tid >> 5;
    const int A = p.lv.anchors_per_cell, L = p.lv.num_levels;
    const int nthreads = 32 * A;
    const int b = p.page_order ? __ldg(p.page_order + blockIdx.y) : (int)blockIdx.y;   // (heaviest pages first, see the wrapper)
    // staging rows are shifted by the destination's misalignment (start & 3 floats) so that 16-byte units of
    // shared memory map to 16-byte units of global memory: the write-out is LDS.128 + STG.128 per unit
    const int reg_stride = 32 * A * 5 + 4, lab_stride = 32 * A * 2 + 4;     // floats per staged tile row (multiples of 4)
    float* s_reg = s_dyn;                                   // [KT_ROWS][reg_stride]
    float* s_lab = s_reg + KT_ROWS * reg_stride;            // C == 1: [KT_ROWS][lab_stride] {one-hot, state} pairs
    float* s_state = s_lab;                                 // C  > 1: [KT_ROWS][32][A] states, then the hot classes
    int* s_hot = reinterpret_cast<int*>(s_lab + KT_ROWS * 32 * A);
    double* s_ih = reinterpret_cast<double*>(s_lab + KT_ROWS * lab_stride);   // [A][KT_ROWS][32]; 8-byte aligned
    if (tid == 0) s_npos = 0;

    // ---- tile -> (level, tile x, tile y): block-uniform -------------------------------------------------
    // (static indices only: dynamic indexing of by-value kernel parameters would force a local-memory copy)
    int level = 0, tstart = 0, tiles_x = tl.tiles_x[0], W = p.lv.w[0], H = p.lv.h[0], istride = p.lv.stride[0], lstart = p.lv.start[0];
    float inv_tiles_x = tl.inv_tiles_x[0];
#pragma unroll
    for (int l = 1; l < RN_MAX_LEVELS; ++l)
        if (l < L && (int)blockIdx.x >= tl.tile_start[l]) {
            level = l; tstart = tl.tile_start[l]; tiles_x = tl.tiles_x[l]; inv_tiles_x = tl.inv_tiles_x[l];
            W = p.lv.w[l]; H = p.lv.h[l]; istride = p.lv.stride[l]; lstart = p.lv.start[l];
        }
    const int t = blockIdx.x - tstart;
    const int ty = rn_div(t, tiles_x, inv_tiles_x);
    const int tx = t - ty * tiles_x;
    const double stride = (double)istride;
    const int cx0 = tx * 32, cy0 = ty * KT_ROWS;
    const int ncols = min(32, W - cx0), nrows = min(KT_ROWS, H - cy0);
    const bool valid_x = lane < ncols;
    int G = p.gt_count[b];
    G = max(0, min(G, p.Gmax));

    // ---- geometry: base box (warp-uniform), column extent (per thread), row extents (warp-uniform) -----
    const double* bs = p.base + ((size_t)level * A + a) * 4;
    const double b0 = __ldg(bs), b1 = __ldg(bs + 1), b2 = __ldg(bs + 2), b3 = __ldg(bs + 3);
    const double sx = ((double)(cx0 + lane) + 0.5) * stride;
    const double ax1 = b0 + sx, ax2 = b2 + sx;
    const double aw = ax2 - ax1;
    // row geometry is the same for the whole warp: keep it in shared memory, not in 24 registers per thread
    if (lane < KT_ROWS) {
        const double sy = ((double)(cy0 + lane) + 0.5) * stride;
        const double y1 = b1 + sy, y2 = b3 + sy;
        s_row[a][lane][0] = y1; s_row[a][lane][1] = y2; s_row[a][lane][2] = y2 - y1;
    }
    __syncwarp();
    const double (*row)[3] = s_row[a];
    const bool match_x = valid_x && (aw > 0.0);
    // exact bounding box of the warp's anchors (first / last valid column, first / last valid row)
    const double wx1 = b0 + ((double)cx0 + 0.5) * stride, wx2 = b2 + ((double)(cx0 + ncols - 1) + 0.5) * stride;

    // ---- matching -------------------------------------------------------------------------------------
    // The y half of every (anchor, table) overlap depends only on (anchor type, tile row, table): it is the
    // same for all 32 lanes of the warp.  So for each group of 32 tables lane j computes the intersection
    // heights of table j with the warp's KT_ROWS rows once (s_ih) plus a row mask; the per-thread loop then
    // only does the x half once per table and one multiply / divide / compare per overlapping row.
    float best[KT_ROWS];
    int arg[KT_ROWS];
#pragma unroll
    for (int r = 0; r < KT_ROWS; ++r) { best[r] = 0.0f; arg[r] = 0; }   // all-zero IoU row -> argmax 0
    double* ihw = s_ih + (size_t)a * (KT_ROWS * 32);       // this warp's [KT_ROWS][32] intersection heights
    const double* gtb = p.gt + (size_t)b * p.Gmax * 4;
    for (int g0 = 0; g0 < G; g0 += KT_CHUNK) {
        const int chunk = min(KT_CHUNK, G - g0);
        if (g0) __syncthreads();                           // previous chunk consumed (nothing has been staged before the first)
        for (int j = tid; j < chunk; j += nthreads) {
            const double gx1 = __ldg(gtb + 4 * (g0 + j)), gy1 = __ldg(gtb + 4 * (g0 + j) + 1);
            const double gx2 = __ldg(gtb + 4 * (g0 + j) + 2), gy2 = __ldg(gtb + 4 * (g0 + j) + 3);
            s_gx1[j] = gx1; s_gy1[j] = gy1; s_gx2[j] = gx2; s_gy2[j] = gy2;
            s_ga[j] = (gx2 - gx1) * (gy2 - gy1);
            s_glab[j] = __ldg(p.gt_labels + (size_t)b * p.Gmax + g0 + j);
        }
        __syncthreads();
        for (int q0 = 0; q0 < chunk; q0 += 32) {
            const int j = q0 + lane;
            unsigned rows_hit = 0u;
            if (j < chunk) {
                const double gx1 = s_gx1[j], gy1 = s_gy1[j], gx2 = s_gx2[j], gy2 = s_gy2[j];
                // empty tables and tables outside the warp's x range have zero intersection with all its anchors
                if ((gx2 > gx1) && (gy2 > gy1) && (gx2 > wx1) && (gx1 < wx2)) {
#pragma unroll
                    for (int r = 0; r < KT_ROWS; ++r) {
                        const double y1 = row[r][0], y2 = row[r][1], hh = row[r][2];
                        ihw[r * 32 + lane] = dmin(y2, gy2) - dmax(y1, gy1);
                        if (r < nrows && gy2 > y1 && gy1 < y2 && hh > 0.0) rows_hit |= 1u << r;
                    }
                }
            }
            unsigned live = __ballot_sync(0xffffffffu, rows_hit != 0u);   // also orders the s_ih writes
            while (live) {                                 // warp-uniform, ascending GT order
                const int ml = __ffs(live) - 1;
                live &= live - 1u;
                const int m = q0 + ml;
                const unsigned rmask = __shfl_sync(0xffffffffu, rows_hit, ml);
                const double g1 = s_gx1[m], g2 = s_gx2[m];
                if (match_x && g2 > ax1 && g1 < ax2) {
                    const double iw = dmin(ax2, g2) - dmax(ax1, g1);
                    const double ga = s_ga[m];
                    if (rmask == (1u << KT_ROWS) - 1u) {
                        // every row of the tile overlaps (the common case inside a table): no per-row branches, so
                        // pairs of reciprocal chains (MUFU -> DFMA -> DFMA -> DMUL) overlap each other
#pragma unroll
                        for (int r = 0; r < KT_ROWS; r += 2) {          // two rows at a time: ILP without spilling
                            const double i0 = iw * ihw[r * 32 + ml], i1 = iw * ihw[(r + 1) * 32 + ml];
                            const double u0 = aw * row[r][2] + ga - i0, u1 = aw * row[r + 1][2] + ga - i1;
                            float iou0, iou1;
                            iou_pair(i0, u0, i1, u1, iou0, iou1);
                            if (iou0 > best[r]) { best[r] = iou0; arg[r] = g0 + m; }
                            if (iou1 > best[r + 1]) { best[r + 1] = iou1; arg[r + 1] = g0 + m; }
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < KT_ROWS; ++r) {
                            if (rmask & (1u << r)) {
                                const double inter = iw * ihw[r * 32 + ml];
                                const double uni = aw * row[r][2] + ga - inter;
                                const double q = inter * rcp_fast(uni);
                                const float iou = f32_rounding_safe(q) ? (float)q : iou_exact(inter, uni);
                                if (iou > best[r]) { best[r] = iou; arg[r] = g0 + m; }
                            }
                        }
                    }
                }
            }
            __syncwarp();                                  // s_ih is rewritten by the next group
        }
    }

    // ---- state, one-hot class, regression targets, border rule ------------------------------------------
    int my_pos = 0;
    // misalignment (in anchors, mod 4) of the tile's first anchor and of one feature-map row; unsigned wrap-around
    // keeps the low two bits right.  Row r is staged shifted by ((al0 + r * alw) * 5) & 3 = (al0 + r * alw) & 3 floats
    // (regression) and ((al0 + r * alw) * 2) & 3 floats (labels).
    const unsigned al0 = ((unsigned)b * (unsigned)p.N + (unsigned)lstart + ((unsigned)cy0 * (unsigned)W + (unsigned)cx0) * (unsigned)A) & 3u;
    const unsigned alw = ((unsigned)W * (unsigned)A) & 3u;
    const bool gt_staged = G <= KT_CHUNK;                   // the (only) chunk is still in shared memory
    // 5/width, 5/height for the regression fast path: the base box's stand in for the anchor's own (they
    // differ by rounding only) when that is far inside the fast path's tolerance (see the wrapper)
    const double bw = __ldg(bs + 2) - __ldg(bs), bh = __ldg(bs + 3) - __ldg(bs + 1);     // re-read: not kept live across the matching loop
    const bool table_ok = (bw > 0.0) && (bh > 0.0) && (p.max_coord < 4096.0 * fmin(bw, bh));
    const double r5w = 5.0 * rcp_fast(table_ok ? bw : aw);
    const double r5h_tab = table_ok ? 5.0 * rcp_fast(bh) : 0.0;
    if (valid_x) {
        bool out_x = false;
        double img_h = 0.0;
        if (p.img_hw) {
            out_x = ((ax1 + ax2) / 2.0) >= (double)p.img_hw[2 * b + 1];
            img_h = (double)p.img_hw[2 * b];
        }
        int prev = -1;                                      // the x targets depend on the column and the table only
        float t0 = 0.f, t2 = 0.f;
        double gy1 = 0.0, gy2 = 0.0;
#pragma unroll
        for (int r = 0; r < KT_ROWS; ++r) {
            if (r < nrows) {
                const double y1 = row[r][0], y2 = row[r][1], hh = row[r][2];
                float state = 0.0f, t1 = 0.f, t3 = 0.f;
                int hot = -1;
                if (G > 0) {
                    const int m = arg[r];
                    const bool is_pos = best[r] >= p.pos;
                    const bool is_ign = (best[r] > p.neg) && !is_pos;
                    state = is_pos ? 1.0f : (is_ign ? -1.0f : 0.0f);
                    if (m != prev) {                        // coordinates and x targets change with the table only
                        prev = m;
                        double gx1, gx2;
                        if (gt_staged) {
                            gx1 = s_gx1[m]; gy1 = s_gy1[m]; gx2 = s_gx2[m]; gy2 = s_gy2[m];
                        } else {
                            const double* g = gtb + 4 * (size_t)m;
                            gx1 = __ldg(g); gy1 = __ldg(g + 1); gx2 = __ldg(g + 2); gy2 = __ldg(g + 3);
                        }
                        reg_target5_pair(gx1, ax1, gx2, ax2, aw, r5w, t0, t2);
                    }
                    if (is_pos) hot = gt_staged ? s_glab[m] : __ldg(p.gt_labels + (size_t)b * p.Gmax + m);
                    const double r5h = table_ok ? r5h_tab : 5.0 * rcp_fast(hh);
                    reg_target5_pair(gy1, y1, gy2, y2, hh, r5h, t1, t3);
                }
                if (p.img_hw && (out_x || ((y1 + y2) / 2.0) >= img_h)) state = -1.0f;
                const int k = lane * A + a;                 // reference order within the tile row
                const unsigned al = al0 + r * alw;          // first anchor of the staged row, mod 4 in the low bits
                float* sr = s_reg + r * reg_stride + (int)(al & 3u) + k * 5;
                sr[0] = t0; sr[1] = t1; sr[2] = t2; sr[3] = t3; sr[4] = state;
                if (C1) {
                    float* sl = s_lab + r * lab_stride + (int)((al & 1u) * 2u) + k * 2;
                    sl[0] = hot == 0 ? 1.0f : 0.0f; sl[1] = state;
                } else {
                    s_state[r * 32 * A + k] = state;
                    s_hot[r * 32 * A + k] = hot;
                }
                my_pos += (state == 1.0f);
                if (AM && p.argmax)
                    p.argmax[(size_t)b * p.N + lstart + ((size_t)(cy0 + r) * W + cx0 + lane) * A + a] = arg[r];
            }
        }
    }
    if (p.npos || p.npos_total) {
        my_pos = rn_warp_sum(my_pos);
        if (lane == 0 && my_pos) atomicAdd(&s_npos, my_pos);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the staged rows are read by the TMA engine below
    __syncthreads();
    if (tid == 0 && s_npos) {
        if (p.npos) atomicAdd(p.npos + b, s_npos);
        if (p.npos_total) atomicAdd(p.npos_total, (float)s_npos);   // integer-valued: exact, order-independent
    }

    tile_write_out<C1>(p, s_reg, s_lab, s_state, s_hot, b, lstart, cy0, cx0, W, A, ncols, nrows, reg_stride, lab_stride, tid, nthreads);
}

// ------------------------------------------------------------------------------------------------
// K1 for pages with at most 32 GT tables and at most 9 anchor types per cell -- the table-detection case.  Same tiling and,
// operation for operation, the same arithmetic as k_anchor_targets_tiles (which stays the general path: any number of
// tables, up to 24 anchor types); what differs is WHO computes what:
//  * one GT group: no chunk / group loops, the staged tables stay valid for the epilogue;
//  * the y half of a regression target depends on (anchor type, tile row, table) only -- it is the same for all 32 lanes of
//    a warp.  After the matching the warp computes it ONCE per (row, table that can be an argmax of the warp: table 0 and the
//    tables that reach the warp), lane = (table slot, row), 8 tables per pass, usually one pass, into a table in shared
//    memory (it takes over the intersection heights' space).  A thread's epilogue then costs one 64-bit shared-memory load
//    per anchor instead of two fp64 quotients with their rounding certificates (model/anchors.py:300-311);
//  * the row half of the border rule (model/anchors.py:85-90) is a per-warp bit mask: 4 lanes + one ballot;
//  * the one-hot label of the (rare) positive anchors is written behind ONE branch after the row loop, so a row's body has
//    no convergence region but the change of the argmax table.
// ------------------------------------------------------------------------------------------------
constexpr int K32_G = 32;
constexpr int K32_A = 9;
#ifndef RN_K32_XT
#define RN_K32_XT 3
#endif
constexpr int K32_XT = RN_K32_XT;            // x tiles (of 32 columns) a CTA walks: everything that depends on (anchor type, rows, table)
                                       // only -- intersection heights, the y-target table, the border mask, the tile decode and the
                                       // staged tables -- is computed once for all of them.  (A/B on one box, profiles/r2h_sweep_k1.log:
                                       // 2 tiles 48.7 us / 31.0 M warp-instructions, 1 tile 51.3 us / 33.4 M, outputs bit-identical.)

// page slots at the end of the launch that are cut into one-tile CTAs (see K1Tiles32).  A one-tile CTA repeats the per-CTA
// work (tile decode, staged tables, intersection heights, y-target table: an empty one takes 3.6 us against 7.0 us for three
// tiles), so only the very last page slot pays: A/B on one box (profiles/r2/r2q_sweep_k1_fine.log), 16 pages of 800 x 1333 /
// 4 pages of 1600 x 2400: 0 fine pages 45.2 / 39.5 us, 1: 44.7 / 38.2, 2: 44.9 / 39.5, 3: 45.1 / 42.0, 4: 45.2 / 43.2, 6: 47.1 us.
#ifndef RN_K32_FINE_PAGES
#define RN_K32_FINE_PAGES 1
#endif
static int k32_fine_pages(int B) {
    return K32_XT > 1 ? (RN_K32_FINE_PAGES < B ? RN_K32_FINE_PAGES : B) : 0;
}

// dynamic shared memory of k_anchor_targets_tiles32: staged regression rows, staged label rows, intersection heights, y-target table
static size_t k32_dyn_smem(int A) {
    return kt_dyn_smem(A) + (K32_XT > 1 ? (size_t)A * KT_ROWS * 32 * sizeof(float2) : 0);
}

#ifdef RN_K1_TRACE
// A/B build only (profiles/k1_cta_trace.py): per CTA {start, end} of %globaltimer, SM id, level -- what the launch's tail is made of
__device__ unsigned long long g_k1_trace[4 * 16384];
#endif

// SPARSE (C1 only): the training step's form.  The smooth-L1 loss reads the regression targets of state == 1 anchors only
// (model/losses.py:72-74 gathers exactly those rows), so only THOSE rows of the (B, N, 5) tensor are written -- straight from
// the rare-positive branch, with the exact IEEE expressions -- and neither the y-target table nor any other anchor's four
// quotients are computed: 8 instead of 28 bytes per anchor leave the SM.  Labels, states, counts: unchanged.
template <bool C1, bool AM, int XT, bool SPARSE = false>
__global__ void __launch_bounds__(32 * K32_A, 3) k_anchor_targets_tiles32(const K1Params p, const K1Tiles32 tl) {
    extern __shared__ __align__(16) float s_dyn[];
    __shared__ double s_gx1[K32_G], s_gy1[K32_G], s_gx2[K32_G], s_gy2[K32_G], s_ga[K32_G];
    __shared__ int s_glab[K32_G];
    __shared__ double s_row[K32_A][KT_ROWS][3];             // per (anchor type, tile row): y1, y2, height
    __shared__ unsigned char s_list[K32_A][32];             // per warp: the tables that can be an argmax, ascending
    __shared__ int s_npos;

    const int tid = threadIdx.x, lane = tid & 31, a = tid >> 5;
    const int A = p.lv.anchors_per_cell, L = p.lv.num_levels;
    const int nthreads = 32 * A;
    const int reg_stride = 32 * A * 5 + 4, lab_stride = 32 * A * 2 + 4;     // floats per staged tile row (multiples of 4)
    float* s_reg = s_dyn;                                   // [KT_ROWS][reg_stride]
    float* s_lab = s_reg + KT_ROWS * reg_stride;            // C == 1: [KT_ROWS][lab_stride] {one-hot, state} pairs
    float* s_state = s_lab;                                 // C  > 1: [KT_ROWS][32][A] states, then the hot classes
    int* s_hot = reinterpret_cast<int*>(s_lab + KT_ROWS * 32 * A);
    double* s_ih = reinterpret_cast<double*>(s_lab + KT_ROWS * lab_stride);   // [A][KT_ROWS][32]; 8-byte aligned
    if (tid == 0) s_npos = 0;
#ifdef RN_K1_TRACE
    unsigned long long trace_t0 = 0;
    if (tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(trace_t0));
#endif

    // ---- tile -> (level, tile x, tile y): block-uniform (static indices only, see k_anchor_targets_tiles) ----
    // CTA number -> (page slot, tile of the page, x tiles this CTA walks): coarse page slots first, fine ones at the end
    const int n_coarse = tl.c.tile_start[RN_MAX_LEVELS], n_fine = tl.f.tile_start[RN_MAX_LEVELS];
    const int split = tl.coarse_pages * n_coarse;
    const bool fine = (int)blockIdx.x >= split;
    const int xt = fine ? 1 : XT;                           // block-uniform
    const int rest = fine ? (int)blockIdx.x - split : (int)blockIdx.x, per_page = fine ? n_fine : n_coarse;
    const int slot_in = rest / per_page;
    const int tile_id = rest - slot_in * per_page;
    const int bslot = (fine ? tl.coarse_pages : 0) + slot_in;
    const int b = p.page_order ? __ldg(p.page_order + bslot) : bslot;   // (heaviest pages first, see the wrapper)
    int level = 0, tstart = 0, tiles_x = fine ? tl.f.tiles_x[0] : tl.c.tiles_x[0];
    int W = p.lv.w[0], H = p.lv.h[0], istride = p.lv.stride[0], lstart = p.lv.start[0];
    float inv_tiles_x = fine ? tl.f.inv_tiles_x[0] : tl.c.inv_tiles_x[0];
#pragma unroll
    for (int l = 1; l < RN_MAX_LEVELS; ++l) {
        const int ts = fine ? tl.f.tile_start[l] : tl.c.tile_start[l];
        if (l < L && tile_id >= ts) {
            level = l; tstart = ts; tiles_x = fine ? tl.f.tiles_x[l] : tl.c.tiles_x[l]; inv_tiles_x = fine ? tl.f.inv_tiles_x[l] : tl.c.inv_tiles_x[l];
            W = p.lv.w[l]; H = p.lv.h[l]; istride = p.lv.stride[l]; lstart = p.lv.start[l];
        }
    }
    const int t = tile_id - tstart;
    const int ty = rn_div(t, tiles_x, inv_tiles_x);
    const int tx = t - ty * tiles_x;
    const double stride = (double)istride;
    const int cxt = tx * (32 * xt), cy0 = ty * KT_ROWS;     // first column of the CTA's xt tiles
    const int nrows = min(KT_ROWS, H - cy0);
    int G = p.gt_count[b];
    G = max(0, min(G, min(p.Gmax, K32_G)));

    // ---- the page's tables -> shared memory (one per thread) -------------------------------------------
    const double* gtb = p.gt + (size_t)b * p.Gmax * 4;
    if (tid < G) {
        const double gx1 = __ldg(gtb + 4 * tid), gy1 = __ldg(gtb + 4 * tid + 1);
        const double gx2 = __ldg(gtb + 4 * tid + 2), gy2 = __ldg(gtb + 4 * tid + 3);
        s_gx1[tid] = gx1; s_gy1[tid] = gy1; s_gx2[tid] = gx2; s_gy2[tid] = gy2;
        s_ga[tid] = (gx2 - gx1) * (gy2 - gy1);
        s_glab[tid] = __ldg(p.gt_labels + (size_t)b * p.Gmax + tid);
    }

    // ---- geometry that does not depend on the column: base box (warp-uniform), row extents (warp-uniform) -----
    const double* bs = p.base + ((size_t)level * A + a) * 4;
    const double b0 = __ldg(bs), b1 = __ldg(bs + 1), b2 = __ldg(bs + 2), b3 = __ldg(bs + 3);
    if (lane < KT_ROWS) {
        const double sy = ((double)(cy0 + lane) + 0.5) * stride;
        const double y1 = b1 + sy, y2 = b3 + sy;
        s_row[a][lane][0] = y1; s_row[a][lane][1] = y2; s_row[a][lane][2] = y2 - y1;
    }
    __syncthreads();                                        // staged tables, row geometry, s_npos
    const double (*row)[3] = s_row[a];

    // ---- lane j prepares table j's intersection heights with the warp's rows (the y half of every overlap is the same for
    //      all columns) and finds out which of the XT tiles the table reaches ------------------------------------------
    double* ihw = s_ih + (size_t)a * (KT_ROWS * 32);       // this warp's [KT_ROWS][32] intersection heights
    unsigned rows_hit = 0u, reach = 0u;                     // reach: bit h = the table's x range meets tile h's anchors
    if (lane < G) {
        const double gx1 = s_gx1[lane], gy1 = s_gy1[lane], gx2 = s_gx2[lane], gy2 = s_gy2[lane];
        if ((gx2 > gx1) && (gy2 > gy1)) {                   // empty tables have zero intersection with everything
#pragma unroll
            for (int h = 0; h < XT; ++h) {
                const int c0 = cxt + 32 * h, nc = min(32, W - c0);
                if (h < xt && nc > 0) {
                    // exact bounding box of the warp's anchors in tile h (first / last valid column)
                    const double wx1 = b0 + ((double)c0 + 0.5) * stride, wx2 = b2 + ((double)(c0 + nc - 1) + 0.5) * stride;
                    if ((gx2 > wx1) && (gx1 < wx2)) reach |= 1u << h;
                }
            }
            if (reach) {
#pragma unroll
                for (int r = 0; r < KT_ROWS; ++r) {
                    const double y1 = row[r][0], y2 = row[r][1], hh = row[r][2];
                    ihw[r * 32 + lane] = dmin(y2, gy2) - dmax(y1, gy1);
                    if (r < nrows && gy2 > y1 && gy1 < y2 && hh > 0.0) rows_hit |= 1u << r;
                }
            }
        }
    }
    __syncwarp();                                           // the intersection heights are visible to the whole warp
    unsigned live_h[XT];
    unsigned cand = 1u;                                     // possible argmax tables of this warp: table 0 (nothing overlaps) + live
#pragma unroll
    for (int h = 0; h < XT; ++h) {
        live_h[h] = __ballot_sync(0xffffffffu, rows_hit != 0u && ((reach >> h) & 1u));
        cand |= live_h[h];
    }

    // ---- y targets per (row, candidate table), computed by the warp once ---------------------------------
    // 5/width, 5/height for the regression fast path: the base box's stand in for the anchor's own (they
    // differ by rounding only) when that is far inside the fast path's tolerance (see the wrapper)
    const double bw = b2 - b0, bh = b3 - b1;
    const bool table_ok = (bw > 0.0) && (bh > 0.0) && (p.max_coord < 4096.0 * fmin(bw, bh));
    // XT == 1: the table takes over the intersection heights' space once the matching is done; XT > 1: its own space
    float2* tyw = XT > 1 ? reinterpret_cast<float2*>(s_ih + (size_t)A * (KT_ROWS * 32)) + (size_t)a * (KT_ROWS * 32)
                         : reinterpret_cast<float2*>(ihw);  // [KT_ROWS][32]: {t1, t3} of (row, table)
    unsigned out_y = 0u;                                    // bit r: the centres of tile row r lie below the page
    auto make_y_table = [&]() {
        const int r = lane & (KT_ROWS - 1);
        const double y1 = row[r][0], y2 = row[r][1], hh = row[r][2];
        if (G > 0) {
            const int ncand = __popc(cand);
            if ((cand >> lane) & 1u) s_list[a][__popc(cand & ((1u << lane) - 1u))] = (unsigned char)lane;
            __syncwarp();
            const double r5h = table_ok ? 5.0 * rcp_fast(bh) : 5.0 * rcp_fast(hh);
            for (int k = lane >> 2; k < ncand; k += 8) {    // lane = (table slot, row)
                const int m = s_list[a][k];
                float t1, t3;
                reg_target5_pair(s_gy1[m], y1, s_gy2[m], y2, hh, r5h, t1, t3);
                tyw[r * 32 + m] = make_float2(t1, t3);
            }
        }
        if (p.img_hw) out_y = __ballot_sync(0xffffffffu, (lane < KT_ROWS) && (((y1 + y2) / 2.0) >= (double)p.img_hw[2 * b]));
        __syncwarp();                                       // the table is complete and visible to the whole warp
    };
    if (XT > 1 && !SPARSE) make_y_table();
    if (SPARSE && p.img_hw) {                               // (the border mask is a by-product of make_y_table)
        const int r = lane & (KT_ROWS - 1);
        out_y = __ballot_sync(0xffffffffu, (lane < KT_ROWS) && (((row[r][0] + row[r][1]) / 2.0) >= (double)p.img_hw[2 * b]));
    }

    int my_pos = 0;
#pragma unroll 1
    for (int half = 0; half < XT; ++half) {
        const int cx0 = cxt + 32 * half;
        const int ncols = min(32, W - cx0);
        if (half >= xt || ncols <= 0) break;                // block-uniform
        const bool valid_x = lane < ncols;
        // ---- column extent (per thread) ----------------------------------------------------------------------------
        const double sx = ((double)(cx0 + lane) + 0.5) * stride;
        const double ax1 = b0 + sx, ax2 = b2 + sx;
        const double aw = ax2 - ax1;
        const bool match_x = valid_x && (aw > 0.0);

        // ---- matching: the live tables are walked in GT order with warp-uniform control flow (first maximum wins) ------
        float best[KT_ROWS];
        int arg[KT_ROWS];
#pragma unroll
        for (int r = 0; r < KT_ROWS; ++r) { best[r] = 0.0f; arg[r] = 0; }   // all-zero IoU row -> argmax 0
        unsigned live = live_h[0];
#pragma unroll
        for (int h = 1; h < XT; ++h) if (half == h) live = live_h[h];
        while (live) {                                      // warp-uniform, ascending GT order
            const int m = __ffs(live) - 1;
            live &= live - 1u;
            const unsigned rmask = __shfl_sync(0xffffffffu, rows_hit, m);
            const double g1 = s_gx1[m], g2 = s_gx2[m];
            if (match_x && g2 > ax1 && g1 < ax2) {
                const double iw = dmin(ax2, g2) - dmax(ax1, g1);
                const double ga = s_ga[m];
                if (rmask == (1u << KT_ROWS) - 1u) {
                    // every row of the tile overlaps (the common case inside a table): no per-row branches, so
                    // pairs of reciprocal chains (MUFU -> DFMA -> DFMA -> DMUL) overlap each other
#pragma unroll
                    for (int r = 0; r < KT_ROWS; r += 2) {
                        const double i0 = iw * ihw[r * 32 + m], i1 = iw * ihw[(r + 1) * 32 + m];
                        const double u0 = aw * row[r][2] + ga - i0, u1 = aw * row[r + 1][2] + ga - i1;
                        float iou0, iou1;
                        iou_pair(i0, u0, i1, u1, iou0, iou1);
                        if (iou0 > best[r]) { best[r] = iou0; arg[r] = m; }
                        if (iou1 > best[r + 1]) { best[r + 1] = iou1; arg[r + 1] = m; }
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < KT_ROWS; ++r) {
                        if (rmask & (1u << r)) {
                            const double inter = iw * ihw[r * 32 + m];
                            const double uni = aw * row[r][2] + ga - inter;
                            const double q = inter * rcp_fast(uni);
                            const float iou = f32_rounding_safe(q) ? (float)q : iou_exact(inter, uni);
                            if (iou > best[r]) { best[r] = iou; arg[r] = m; }
                        }
                    }
                }
            }
        }
        if (XT == 1 && !SPARSE) {
            __syncwarp();                                   // the intersection heights are dead: their space becomes the y-target table
            make_y_table();
        }
        const double r5w = 5.0 * rcp_fast(table_ok ? bw : aw);
        const bool out_x = p.img_hw ? (((ax1 + ax2) / 2.0) >= (double)p.img_hw[2 * b + 1]) : false;

        // ---- state, regression targets, border rule, staging --------------------------------------------------
        // misalignment (in anchors, mod 4) of the tile's first anchor and of one feature-map row; unsigned wrap-around
        // keeps the low two bits right.  Row r is staged shifted by ((al0 + r * alw) * 5) & 3 = (al0 + r * alw) & 3 floats
        // (regression) and ((al0 + r * alw) * 2) & 3 floats (labels).
        const unsigned al0 = ((unsigned)b * (unsigned)p.N + (unsigned)lstart + ((unsigned)cy0 * (unsigned)W + (unsigned)cx0) * (unsigned)A) & 3u;
        const unsigned alw = ((unsigned)W * (unsigned)A) & 3u;
        if (XT > 1 && half > 0) __syncthreads();            // the previous tile's staged rows have been read by the copy engine
        if (valid_x) {
            const int k = lane * A + a;                     // reference order within the tile row
            int prev = -1;                                  // the x targets depend on the column and the table only
            float t0 = 0.f, t2 = 0.f;
            unsigned posbits = 0u, fgbits = 0u;         // fgbits (SPARSE): rows whose FINAL state is 1 (positive and inside the page)
#pragma unroll
            for (int r = 0; r < KT_ROWS; ++r) {
                if (r < nrows) {
                    float state = 0.0f, t1 = 0.f, t3 = 0.f;
                    if (G > 0) {
                        const int m = arg[r];
                        const bool is_pos = best[r] >= p.pos;
                        const bool is_ign = (best[r] > p.neg) && !is_pos;
                        state = is_pos ? 1.0f : (is_ign ? -1.0f : 0.0f);
                        posbits |= is_pos ? (1u << r) : 0u;
                        if (!SPARSE) {
                            if (m != prev) {                // x targets change with the table only
                                prev = m;
                                reg_target5_pair(s_gx1[m], ax1, s_gx2[m], ax2, aw, r5w, t0, t2);
                            }
                            const float2 ty2 = tyw[r * 32 + m];
                            t1 = ty2.x; t3 = ty2.y;
                        }
                    }
                    if (out_x || ((out_y >> r) & 1u)) state = -1.0f;
                    const unsigned al = al0 + r * alw;      // first anchor of the staged row, mod 4 in the low bits
                    if (!SPARSE) {
                        float* sr = s_reg + r * reg_stride + (int)(al & 3u) + k * 5;
                        sr[0] = t0; sr[1] = t1; sr[2] = t2; sr[3] = t3; sr[4] = state;
                    } else if (state == 1.0f) {
                        fgbits |= 1u << r;
                    }
                    if (C1) {
                        *reinterpret_cast<float2*>(s_lab + r * lab_stride + (int)((al & 1u) * 2u) + k * 2) = make_float2(0.0f, state);
                    } else {
                        s_state[r * 32 * A + k] = state;
                        s_hot[r * 32 * A + k] = -1;
                    }
                    my_pos += (state == 1.0f);
                    if (AM && p.argmax)
                        p.argmax[(size_t)b * p.N + lstart + ((size_t)(cy0 + r) * W + cx0 + lane) * A + a] = arg[r];
                }
            }
            if (posbits) {                                  // positives are ~0.2 % of the anchors
#pragma unroll
                for (int r = 0; r < KT_ROWS; ++r) {
                    if (SPARSE && (fgbits & (1u << r))) {
                        // the anchor's regression row, straight to global memory: ((g - a) / len) / 0.2 in the reference's own
                        // IEEE operations (model/anchors.py:300-311) -- what the dense path's fast quotients are certified to equal
                        const int m = arg[r];
                        float* dst = p.reg + ((size_t)b * p.N + lstart + ((size_t)(cy0 + r) * W + cx0 + lane) * A + a) * 5;
                        dst[0] = reg_target_exact(s_gx1[m] - ax1, aw);
                        dst[1] = reg_target_exact(s_gy1[m] - row[r][0], row[r][2]);
                        dst[2] = reg_target_exact(s_gx2[m] - ax2, aw);
                        dst[3] = reg_target_exact(s_gy2[m] - row[r][1], row[r][2]);
                        dst[4] = 1.0f;
                    }
                    if (posbits & (1u << r)) {
                        const int hot = s_glab[arg[r]];
                        if (C1) {
                            if (hot == 0) s_lab[r * lab_stride + (int)(((al0 + r * alw) & 1u) * 2u) + k * 2] = 1.0f;
                        } else {
                            s_hot[r * 32 * A + k] = hot;
                        }
                    }
                }
            }
        }
        const bool last = (half == xt - 1) || (W - (cx0 + 32) <= 0);       // block-uniform
        if (last && (p.npos || p.npos_total)) {
            my_pos = rn_warp_sum(my_pos);
            if (lane == 0 && my_pos) atomicAdd(&s_npos, my_pos);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the staged rows are read by the TMA engine below
        __syncthreads();
        if (last && tid == 0 && s_npos) {
            rn_grid_dependency_wait();                      // the counters are zeroed by the launch in front of this one
            if (p.npos) atomicAdd(p.npos + b, s_npos);
            if (p.npos_total) atomicAdd(p.npos_total, (float)s_npos);   // integer-valued: exact, order-independent
        }
        tile_write_out<C1, SPARSE>(p, s_reg, s_lab, s_state, s_hot, b, lstart, cy0, cx0, W, A, ncols, nrows, reg_stride, lab_stride, tid, nthreads);
    }
#ifdef RN_K1_TRACE
    __syncthreads();
    if (tid == 0) {
        unsigned long long t1; unsigned sm;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
        const unsigned id = blockIdx.x;
        if (id < 16384) {
            g_k1_trace[4 * id] = trace_t0; g_k1_trace[4 * id + 1] = t1; g_k1_trace[4 * id + 2] = sm;
            g_k1_trace[4 * id + 3] = (unsigned long long)level | ((unsigned long long)b << 8) | ((unsigned long long)tile_id << 16);
        }
    }
#endif
}

__global__ void k_anchors_f64(const RnLevels lv, const double* base, int N, double* out) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        double x1, y1, x2, y2;
        make_anchor(lv, base, n, x1, y1, x2, y2);
        double2* o = reinterpret_cast<double2*>(out + (size_t)n * 4);
        o[0] = make_double2(x1, y1);
        o[1] = make_double2(x2, y2);
    }
}

// compute_overlap: one thread per (box1, box2) pair, box2 fastest (row-major (M,G) output)
__global__ void k_compute_overlap(const double* b1, long long M, const double* b2, int G, float* out) {
    const long long total = M * (long long)G;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / G;
        const int j = (int)(e - i * G);
        const double ax1 = b1[4 * i], ay1 = b1[4 * i + 1], ax2 = b1[4 * i + 2], ay2 = b1[4 * i + 3];
        const double gx1 = b2[4 * j], gy1 = b2[4 * j + 1], gx2 = b2[4 * j + 2], gy2 = b2[4 * j + 3];
        const double iw = fmax(0.0, fmin(ax2, gx2) - fmax(ax1, gx1));
        const double ih = fmax(0.0, fmin(ay2, gy2) - fmax(ay1, gy1));
        const double inter = iw * ih;
        const double uni = (ax2 - ax1) * (ay2 - ay1) + (gx2 - gx1) * (gy2 - gy1) - inter;
        out[e] = (float)(inter / uni);
    }
}

struct Norm4d { double mean[4]; double std[4]; };

__global__ void k_bbox_transform(const double* anchors, const double* gt, long long N, const Norm4d nm, double* out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const double ax1 = anchors[4 * i], ay1 = anchors[4 * i + 1], ax2 = anchors[4 * i + 2], ay2 = anchors[4 * i + 3];
        const double aw = ax2 - ax1, ah = ay2 - ay1;
        out[4 * i + 0] = ((gt[4 * i + 0] - ax1) / aw - nm.mean[0]) / nm.std[0];
        out[4 * i + 1] = ((gt[4 * i + 1] - ay1) / ah - nm.mean[1]) / nm.std[1];
        out[4 * i + 2] = ((gt[4 * i + 2] - ax2) / aw - nm.mean[2]) / nm.std[2];
        out[4 * i + 3] = ((gt[4 * i + 3] - ay2) / ah - nm.mean[3]) / nm.std[3];
    }
}

// The tile kernels' static + dynamic shared memory exceeds the 48 KB default: opt in.  The attribute is per DEVICE, so the
// "done" flags are a bit per device ordinal (a process may drive several GPUs; devices >= 64 simply set it every call).
static int k1_opt_in_shared_memory() {
    static std::atomic<unsigned long long> done{0ull};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (bit && (done.load(std::memory_order_acquire) & bit)) return RN_OK;
    const int big = (int)kt_dyn_smem(KT_MAX_A), small = (int)kt_dyn_smem(9);
    cudaError_t ae = cudaFuncSetAttribute(k_anchor_targets_tiles<KT_MAX_A, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_anchor_targets_tiles<KT_MAX_A, 1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_anchor_targets_tiles<9, 3, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, small);
    if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_anchor_targets_tiles<9, 3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, small);
    if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_anchor_targets_tiles32<true, false, K32_XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k32_dyn_smem(9));
    if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_anchor_targets_tiles32<true, true, K32_XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k32_dyn_smem(9));
    if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_anchor_targets_tiles32<false, true, K32_XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k32_dyn_smem(9));
    if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_anchor_targets_tiles32<true, false, K32_XT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k32_dyn_smem(9));
    if (ae != cudaSuccess) return rn_fail(RN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(ae));
    if (bit) done.fetch_or(bit, std::memory_order_release);
    return RN_OK;
}

}  // namespace

extern "C" int rn_bbox_transform(const double* anchors_dev, const double* gt_boxes_dev, long long N,
                                 const double* mean4, const double* std4, double* out_dev, void* stream) {
    RN_REQUIRE(N >= 0, "negative size");
    if (N == 0) return RN_OK;
    RN_REQUIRE(anchors_dev && gt_boxes_dev && mean4 && std4 && out_dev, "NULL pointer");
    Norm4d nm;
    for (int i = 0; i < 4; ++i) { nm.mean[i] = mean4[i]; nm.std[i] = std4[i]; }
    const int blocks = (int)min((N + 255) / 256, (long long)RN_NUM_SMS * 8);
    k_bbox_transform<<<blocks, 256, 0, (cudaStream_t)stream>>>(anchors_dev, gt_boxes_dev, N, nm, out_dev);
    return rn_check_launch("rn_bbox_transform");
}

// page_order_dev: a permutation of 0 .. B-1 (device) or NULL.  CTAs are handed out page by page (blockIdx.y), so the pages
// at the END of the order decide the tail of the launch: with the heaviest pages (most / largest tables) first the tail is
// made of cheap CTAs.  Results do not depend on the order.  (16 pages of 800x1333 whose heaviest page came last: 56.7 us;
// pages whose last three were light: 46 us for more total work -- profiles/r2/final_k1_order.log.)
extern "C" int rn_anchor_targets_ordered(const double* base_anchors_dev, const int* level_hw, const int* level_stride,
                                         int num_levels, int anchors_per_cell,
                                         const double* anchors_dev, long long num_anchors,
                                         const double* gt_boxes_dev, const int* gt_labels_dev, const int* gt_count_dev,
                                         const int* img_hw_dev, int B, int Gmax, int C,
                                         float neg_overlap, float pos_overlap,
                                         float* regression_out, float* labels_out, int* argmax_out, int* npos_out,
                                         float* npos_total_out, const int* page_order_dev, void* stream);
static int anchor_targets_impl(const double* base_anchors_dev, const int* level_hw, const int* level_stride,
                               int num_levels, int anchors_per_cell,
                               const double* anchors_dev, long long num_anchors,
                               const double* gt_boxes_dev, const int* gt_labels_dev, const int* gt_count_dev,
                               const int* img_hw_dev, int B, int Gmax, int C,
                               float neg_overlap, float pos_overlap,
                               float* regression_out, float* labels_out, int* argmax_out, int* npos_out,
                               float* npos_total_out, const int* page_order_dev, unsigned flags, void* stream);

extern "C" int rn_anchor_targets(const double* base_anchors_dev, const int* level_hw, const int* level_stride,
                                 int num_levels, int anchors_per_cell,
                                 const double* anchors_dev, long long num_anchors,
                                 const double* gt_boxes_dev, const int* gt_labels_dev, const int* gt_count_dev,
                                 const int* img_hw_dev, int B, int Gmax, int C,
                                 float neg_overlap, float pos_overlap,
                                 float* regression_out, float* labels_out, int* argmax_out, int* npos_out,
                                 float* npos_total_out, void* stream) {
    return rn_anchor_targets_ordered(base_anchors_dev, level_hw, level_stride, num_levels, anchors_per_cell, anchors_dev, num_anchors,
                                     gt_boxes_dev, gt_labels_dev, gt_count_dev, img_hw_dev, B, Gmax, C, neg_overlap, pos_overlap,
                                     regression_out, labels_out, argmax_out, npos_out, npos_total_out, nullptr, stream);
}

extern "C" int rn_anchor_targets_ordered(const double* base_anchors_dev, const int* level_hw, const int* level_stride,
                                         int num_levels, int anchors_per_cell,
                                         const double* anchors_dev, long long num_anchors,
                                         const double* gt_boxes_dev, const int* gt_labels_dev, const int* gt_count_dev,
                                         const int* img_hw_dev, int B, int Gmax, int C,
                                         float neg_overlap, float pos_overlap,
                                         float* regression_out, float* labels_out, int* argmax_out, int* npos_out,
                                         float* npos_total_out, const int* page_order_dev, void* stream) {
    return anchor_targets_impl(base_anchors_dev, level_hw, level_stride, num_levels, anchors_per_cell, anchors_dev, num_anchors,
                               gt_boxes_dev, gt_labels_dev, gt_count_dev, img_hw_dev, B, Gmax, C, neg_overlap, pos_overlap,
                               regression_out, labels_out, argmax_out, npos_out, npos_total_out, page_order_dev, 0u, stream);
}

extern "C" int rn_anchor_targets_sparse(const double* base_anchors_dev, const int* level_hw, const int* level_stride,
                                        int num_levels, int anchors_per_cell, long long num_anchors,
                                        const double* gt_boxes_dev, const int* gt_labels_dev, const int* gt_count_dev,
                                        const int* img_hw_dev, int B, int Gmax,
                                        float neg_overlap, float pos_overlap,
                                        float* regression_out, float* labels_out, int* npos_out,
                                        float* npos_total_out, const int* page_order_dev, void* stream) {
    RN_REQUIRE(anchors_per_cell >= 1 && anchors_per_cell <= K32_A && Gmax <= K32_G,
               "sparse regression targets: at most %d anchor types per cell and %d tables per page", K32_A, K32_G);
    return anchor_targets_impl(base_anchors_dev, level_hw, level_stride, num_levels, anchors_per_cell, nullptr, num_anchors,
                               gt_boxes_dev, gt_labels_dev, gt_count_dev, img_hw_dev, B, Gmax, 1, neg_overlap, pos_overlap,
                               regression_out, labels_out, nullptr, npos_out, npos_total_out, page_order_dev, 1u, stream);
}

static int anchor_targets_impl(const double* base_anchors_dev, const int* level_hw, const int* level_stride,
                               int num_levels, int anchors_per_cell,
                               const double* anchors_dev, long long num_anchors,
                               const double* gt_boxes_dev, const int* gt_labels_dev, const int* gt_count_dev,
                               const int* img_hw_dev, int B, int Gmax, int C,
                               float neg_overlap, float pos_overlap,
                               float* regression_out, float* labels_out, int* argmax_out, int* npos_out,
                               float* npos_total_out, const int* page_order_dev, unsigned flags, void* stream) {
    RN_REQUIRE(B >= 1 && B <= 65535, "B must be in [1, 65535] (got %d)", B);
    RN_REQUIRE(C >= 1, "C must be >= 1");
    RN_REQUIRE(Gmax >= 0, "Gmax must be >= 0");
    RN_REQUIRE(num_anchors >= 1 && num_anchors < (1ll << 31), "num_anchors out of range");
    RN_REQUIRE(gt_count_dev && regression_out && labels_out, "NULL output / gt_count pointer");
    RN_REQUIRE(Gmax == 0 || (gt_boxes_dev && gt_labels_dev), "NULL GT pointer with Gmax > 0");
    K1Params p;
    if (anchors_dev == nullptr) {
        RN_REQUIRE(base_anchors_dev, "base_anchors_dev is NULL and no explicit anchors given");
        int rc = rn_make_levels(&p.lv, level_hw, level_stride, num_levels, anchors_per_cell);
        if (rc) return rc;
        RN_REQUIRE(p.lv.start[num_levels] == num_anchors, "num_anchors (%lld) does not match the level table (%d)",
                   num_anchors, p.lv.start[num_levels]);
    } else {
        RN_REQUIRE(rn_aligned16(anchors_dev), "anchors_dev must be 16-byte aligned");
        p.lv.num_levels = 0;
    }
    p.base = base_anchors_dev; p.anchors = anchors_dev; p.N = (int)num_anchors;
    p.gt = gt_boxes_dev; p.gt_labels = gt_labels_dev; p.gt_count = gt_count_dev; p.img_hw = img_hw_dev;
    p.page_order = page_order_dev;
    p.Gmax = Gmax; p.C = C; p.neg = neg_overlap; p.pos = pos_overlap;
    p.reg = regression_out; p.lab = labels_out; p.argmax = argmax_out; p.npos = npos_out;
    p.npos_total = npos_total_out;
    p.vec_ok = rn_aligned16(regression_out) && rn_aligned16(labels_out);
    p.sparse_reg = (flags & 1u) ? 1 : 0;
    RN_REQUIRE((reinterpret_cast<uintptr_t>(labels_out) & 7u) == 0, "labels_out must be 8-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    // the two counters are zeroed by ONE small kernel that lets K1 launch behind it at once (rn_common.cuh); K1 touches them at
    // the very end of a CTA (k_anchor_targets_tiles32 waits for the reset there; the other kernels are launched plainly)
    {
        int rc0 = rn_reset_ints(npos_out, B, reinterpret_cast<int*>(npos_total_out), 1, s);
        if (rc0) return rc0;
    }
    dim3 grid((unsigned)((num_anchors + K1_THREADS - 1) / K1_THREADS), (unsigned)B);
    if (anchors_dev) {
        k_anchor_targets<true><<<grid, K1_THREADS, 0, s>>>(p);
    } else if (anchors_per_cell <= KT_MAX_A) {
        const int A = anchors_per_cell;
        // The regression fast path takes 5/width from the base box.  The true width of a generated anchor,
        // RN(b2+s) - RN(b0+s), deviates from b2-b0 by at most ~2 ulp(max coordinate) = 2^-51 max_coord; the
        // fast path tolerates 2^-36 relative, so the kernel uses the base box only where
        // max_coord / min(width, height) < 2^12 (relative deviation < 2^-39) and a per-anchor reciprocal else.
        double max_coord = 1.0;
        K1Tiles tl;
        int tiles = 0;
        for (int l = 0; l < RN_MAX_LEVELS; ++l) { tl.tile_start[l] = 0; tl.tiles_x[l] = 1; tl.inv_tiles_x[l] = 1.0f; }
        for (int l = 0; l < num_levels; ++l) {
            const int h = level_hw[2 * l], w = level_hw[2 * l + 1];
            const double ext = (double)level_stride[l] * (double)((h > w ? h : w) + 1);
            if (ext > max_coord) max_coord = ext;
            tl.tile_start[l] = tiles;
            tl.tiles_x[l] = (w + 31) / 32 > 0 ? (w + 31) / 32 : 1;
            tl.inv_tiles_x[l] = 1.0f / (float)tl.tiles_x[l];
            tiles += ((w + 31) / 32) * ((h + KT_ROWS - 1) / KT_ROWS);
        }
        for (int l = num_levels; l <= RN_MAX_LEVELS; ++l) tl.tile_start[l] = tiles;
        p.max_coord = max_coord;
        const size_t dyn = kt_dyn_smem(A);
        int rc = k1_opt_in_shared_memory();
        if (rc) return rc;
        const dim3 tgrid((unsigned)tiles, (unsigned)B);
        // <= 9 anchor types and <= 32 tables per page (the table-detection case): the y-target-table kernel; the common
        // instantiation (one class, no argmax tensor) is specialised; everything else goes through the general tile kernel
        if (tiles > 0 && A <= K32_A && Gmax <= K32_G) {
            // (a CTA of this kernel walks K32_XT x tiles -- the first page slots -- or one -- the last ones: two tile tables)
            K1Tiles32 t2;
            auto fill = [&](K1Tiles& tt, int xt) {
                int n = 0;
                for (int l = 0; l < RN_MAX_LEVELS; ++l) { tt.tile_start[l] = 0; tt.tiles_x[l] = 1; tt.inv_tiles_x[l] = 1.0f; }
                for (int l = 0; l < num_levels; ++l) {
                    const int h = level_hw[2 * l], w = level_hw[2 * l + 1];
                    const int tx = (w + 32 * xt - 1) / (32 * xt);
                    tt.tile_start[l] = n;
                    tt.tiles_x[l] = tx > 0 ? tx : 1;
                    tt.inv_tiles_x[l] = 1.0f / (float)tt.tiles_x[l];
                    n += tx * ((h + KT_ROWS - 1) / KT_ROWS);
                }
                for (int l = num_levels; l <= RN_MAX_LEVELS; ++l) tt.tile_start[l] = n;
                return n;
            };
            const int n_coarse = fill(t2.c, K32_XT), n_fine = fill(t2.f, 1);
            const int fine_pages = k32_fine_pages(B);
            t2.coarse_pages = B - fine_pages;
            const long long ctas = (long long)t2.coarse_pages * n_coarse + (long long)fine_pages * n_fine;
            if (ctas > 0x7fffffffll) return rn_fail(RN_ERR_BAD_ARG, "rn_anchor_targets: %lld CTAs", ctas);
            const dim3 g2((unsigned)ctas);
            const size_t dyn2 = k32_dyn_smem(A);
            if (p.sparse_reg) return rn_launch_dependent("rn_anchor_targets", k_anchor_targets_tiles32<true, false, K32_XT, true>, g2, dim3(32 * A), dyn2, s, p, t2);
            if (C != 1) return rn_launch_dependent("rn_anchor_targets", k_anchor_targets_tiles32<false, true, K32_XT>, g2, dim3(32 * A), dyn2, s, p, t2);
            else if (argmax_out) return rn_launch_dependent("rn_anchor_targets", k_anchor_targets_tiles32<true, true, K32_XT>, g2, dim3(32 * A), dyn2, s, p, t2);
            else return rn_launch_dependent("rn_anchor_targets", k_anchor_targets_tiles32<true, false, K32_XT>, g2, dim3(32 * A), dyn2, s, p, t2);
        } else if (tiles > 0 && A <= 9) {
            if (C == 1) k_anchor_targets_tiles<9, 3, true, true><<<tgrid, 32 * A, dyn, s>>>(p, tl);
            else k_anchor_targets_tiles<9, 3, false, true><<<tgrid, 32 * A, dyn, s>>>(p, tl);
        } else if (tiles > 0) {
            if (C == 1) k_anchor_targets_tiles<KT_MAX_A, 1, true, true><<<tgrid, 32 * A, dyn, s>>>(p, tl);
            else k_anchor_targets_tiles<KT_MAX_A, 1, false, true><<<tgrid, 32 * A, dyn, s>>>(p, tl);
        }
    } else {
        k_anchor_targets<false><<<grid, K1_THREADS, 0, s>>>(p);
    }
    return rn_check_launch("rn_anchor_targets");
}

extern "C" int rn_anchors_f64(const double* base_anchors_dev, const int* level_hw, const int* level_stride,
                              int num_levels, int anchors_per_cell, double* anchors_out, void* stream) {
    RN_REQUIRE(base_anchors_dev && anchors_out, "NULL pointer");
    RN_REQUIRE(rn_aligned16(anchors_out), "anchors_out must be 16-byte aligned");
    RnLevels lv;
    int rc = rn_make_levels(&lv, level_hw, level_stride, num_levels, anchors_per_cell);
    if (rc) return rc;
    const int N = lv.start[num_levels];
    if (N == 0) return RN_OK;
    const int blocks = min((N + 255) / 256, RN_NUM_SMS * 8);
    k_anchors_f64<<<blocks, 256, 0, (cudaStream_t)stream>>>(lv, base_anchors_dev, N, anchors_out);
    return rn_check_launch("rn_anchors_f64");
}

extern "C" int rn_compute_overlap(const double* boxes1_dev, long long M, const double* boxes2_dev, int G,
                                  float* iou_out, void* stream) {
    RN_REQUIRE(M >= 0 && G >= 0, "negative size");
    if (M == 0 || G == 0) return RN_OK;
    RN_REQUIRE(boxes1_dev && boxes2_dev && iou_out, "NULL pointer");
    const long long total = M * (long long)G;
    const int blocks = (int)min((total + 255) / 256, (long long)RN_NUM_SMS * 16);
    k_compute_overlap<<<blocks, 256, 0, (cudaStream_t)stream>>>(boxes1_dev, M, boxes2_dev, G, iou_out);
    return rn_check_launch("rn_compute_overlap");
}

#ifdef RN_K1_TRACE
extern "C" int rn_debug_k1_trace(unsigned long long* host_out, int ctas) {
    return (int)cudaMemcpyFromSymbol(host_out, g_k1_trace, sizeof(unsigned long long) * 4 * (size_t)ctas);
}
#endif
