// K3 + K4 + K5: FilterDetections (reference model/layers.py:177-264, :298-332) for a whole batch,
// no host synchronisation anywhere.
//
//   K3  k_threshold_compact   one pass over classification (B,N,C): score > thr (strict, fp32);
//                             survivors are decoded (RegressBoxes + ClipBoxes fused, anchors generated
//                             in-kernel in fp32 like the Anchors layer) or gathered from a dense box
//                             tensor, and appended to the (page,class) candidate slab with
//                             warp-aggregated atomics.  The 16-byte regression row of an anchor is only
//                             fetched when that anchor survives the threshold.
//   K4  block radix top-k     inside k_segment_nms: 8-bit MSD radix select over the slab's 64-bit keys
//                             (score bits | ~anchor index) picks the next <= 2048 best candidates, a
//                             shared-memory bitonic network orders them (score desc, anchor asc).
//   K5  bitmask NMS           same kernel: candidates are visited 128 at a time; each is tested against
//                             the boxes selected so far, survivors get a 128x128 suppression bit-matrix
//                             built with __ballot_sync, one warp resolves the greedy order by scanning
//                             set bits only.  Stops at max_detections (TF's max_output_size early stop).
//   k_merge_topk              per page: class-major concatenation + tf.nn.top_k (ties -> earlier
//                             position) done as a C-way merge of the per-class lists; pad with -1.
//
// Semantics restated from tf.image.non_max_suppression / tf.nn.top_k: see oracle/layers_np.py.
// Compiled with -fmad=false: IoU and decode are evaluated in the reference's fp32 operation order.
#include "rn_common.cuh"
#include <stdlib.h>

namespace {

constexpr int K3_THREADS = 256;
constexpr int NMS_THREADS = 1024;
constexpr int NMS_WARPS = NMS_THREADS / 32;
constexpr int NMS_CHUNK = 2048;    // candidates ordered per radix-select round
#ifndef RN_NMS_CHUNK_NEXT
#define RN_NMS_CHUNK_NEXT 256   // A/B at 64 pages x 5 k candidates: 256 -> 190 us, 512 -> 198 us, 1024 -> 197 us, 2048 -> 203 us
#endif
constexpr int NMS_CHUNK_NEXT = RN_NMS_CHUNK_NEXT;   // ... in the rounds after the first
#ifndef RN_NMS_BATCH
#define RN_NMS_BATCH 128         // A/B (64 pages, 5 k candidates each): 64 -> 217 us, 128 -> 203 us, 256 -> 214 us
#endif
constexpr int NMS_BATCH = RN_NMS_BATCH;     // candidates resolved per bit-matrix
constexpr int NMS_WORDS = NMS_BATCH / 32;
constexpr int MAX_DET_LIMIT = 1024;

struct Norm4 { float mean[4]; float std[4]; };

// order-preserving map float -> uint32 (larger float -> larger uint)
__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ unsigned long long make_key(float score, unsigned idx) {
    return ((unsigned long long)f2ord(score) << 32) | (unsigned long long)(0xffffffffu - idx);
}
__device__ __forceinline__ unsigned key_idx(unsigned long long k) { return 0xffffffffu - (unsigned)(k & 0xffffffffu); }
__device__ __forceinline__ float key_score(unsigned long long k) { return ord2f((unsigned)(k >> 32)); }

struct Slabs {
    int* counts;                 // (S)
    unsigned long long* keys;    // (S, cap)
    float4* boxes;               // (S, cap)   [compact mode]   or the caller's (K) boxes [direct mode]
    int* labels;                 // (S, cap)   only for class-agnostic filtering
    long long cap;
};

struct K3Params {
    const float* cls;      // (B, N, C)
    const float* boxes;    // (B, N, 4) dense, or nullptr when decoding
    const float* reg;      // (B, N, 4) regression (decode mode)
    const float* base32;   // (L, A, 4) float32 base anchors (decode mode)
    RnLevels lv;
    Norm4 nm;
    float clipW, clipH;
    int B, N, C;
    int class_specific;
    float thr;
    float inv_c;           // 1 / C for rn_div
    int vec_ok;            // page rows of `cls` are 16-byte aligned
    Slabs sl;
};

template <bool DECODE>
__device__ __forceinline__ float4 candidate_box(const K3Params& p, int b, int n) {
    if (!DECODE) return __ldg(reinterpret_cast<const float4*>(p.boxes) + (size_t)b * p.N + n);
    int level, cx, cy, a;
    rn_locate(p.lv, n, level, cx, cy, a);
    const float* bs = p.base32 + ((size_t)level * p.lv.anchors_per_cell + a) * 4;
    const float sx = ((float)cx + 0.5f) * (float)p.lv.stride[level];
    const float sy = ((float)cy + 0.5f) * (float)p.lv.stride[level];
    const float ax1 = __ldg(bs) + sx, ay1 = __ldg(bs + 1) + sy, ax2 = __ldg(bs + 2) + sx, ay2 = __ldg(bs + 3) + sy;
    // scattered 16-byte rows (~2.5 % of the anchors): no L1 allocation and the smallest L2 fetch granularity
    float4 d;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(reinterpret_cast<const float4*>(p.reg) + (size_t)b * p.N + n));
    const float w = ax2 - ax1, h = ay2 - ay1;
    float4 o;
    o.x = ax1 + (d.x * p.nm.std[0] + p.nm.mean[0]) * w;
    o.y = ay1 + (d.y * p.nm.std[1] + p.nm.mean[1]) * h;
    o.z = ax2 + (d.z * p.nm.std[2] + p.nm.mean[2]) * w;
    o.w = ay2 + (d.w * p.nm.std[3] + p.nm.mean[3]) * h;
    o.x = fminf(fmaxf(o.x, 0.0f), p.clipW);
    o.y = fminf(fmaxf(o.y, 0.0f), p.clipH);
    o.z = fminf(fmaxf(o.z, 0.0f), p.clipW);
    o.w = fminf(fmaxf(o.w, 0.0f), p.clipH);
    return o;
}

// Measured (profiles/k3_probe.py, 64 pages): streaming the 51 MB of scores alone takes 12.5-14.5 us; the 326 k
// candidates add ~13 us, which is the HBM random-access rate for their scattered 16-byte regression rows
// (~22 G rows/s), not instruction issue or atomics -- a persistent, warp-autonomous, software-pipelined variant
// (no block barriers, next chunk's loads in flight during the candidate phase, 4 candidates per lane in flight)
// streamed faster (12.5 us) but took 34 us with candidates and was not kept.
// grid = (tiles of a page, pages).  Two phases per CTA so that the sparse candidates (~2.5 % of the scores)
// never make whole warps walk the divergent decode path:
//   1. every thread streams K3_VEC float4 groups of scores (all loads issued up front) and pushes the
//      survivors of `score > thr` into a shared-memory list (shared-memory atomics);
//   2. the list is consumed densely, one candidate per thread: slot reservation (ONE global atomic per CTA
//      when all candidates feed the same slab, i.e. C == 1 or class-agnostic; one per candidate otherwise),
//      box decode, key/box store.
constexpr int K3_VEC = 4;                                   // float4 groups per thread
constexpr int K3_TILE = K3_THREADS * K3_VEC * 4;            // scores per CTA (class-specific path)

template <bool DECODE>
__global__ void __launch_bounds__(K3_THREADS) k_threshold_compact(const K3Params p) {
    __shared__ int s_elem[K3_TILE];
    __shared__ float s_score[K3_THREADS * K3_VEC];          // class-agnostic path only (score = max over classes)
    __shared__ int s_label[K3_THREADS * K3_VEC];            // class-agnostic path only
    __shared__ int s_count, s_base;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) s_count = 0;
    __syncthreads();
    const bool one_slab = (p.C == 1) || !p.class_specific;
    if (p.class_specific) {
        const int total = p.N * p.C;                        // scores of this page, < 2^31 (checked on the host)
        const float* src = p.cls + (size_t)b * total;
        const int tile0 = blockIdx.x * K3_TILE;
        float sv[K3_VEC][4];
        int cnt[K3_VEC];
#pragma unroll
        for (int g = 0; g < K3_VEC; ++g) {
            const int e0 = tile0 + (g * K3_THREADS + tid) * 4;
            cnt[g] = max(0, min(4, total - e0));
            sv[g][0] = sv[g][1] = sv[g][2] = sv[g][3] = 0.f;
            if (cnt[g] == 4 && p.vec_ok) { const float4 v = rn_ldg_stream4(src + e0); sv[g][0] = v.x; sv[g][1] = v.y; sv[g][2] = v.z; sv[g][3] = v.w; }
            else {
#pragma unroll
                for (int k = 0; k < 4; ++k) if (k < cnt[g]) sv[g][k] = __ldg(src + e0 + k);   // static indices: sv stays in registers
            }
        }
        // survivors of this thread -> warp prefix sum -> ONE shared-memory atomic per warp reserves the list slots
        unsigned hits = 0u;                                 // bit g*4+k
#pragma unroll
        for (int g = 0; g < K3_VEC; ++g)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < cnt[g] && sv[g][k] > p.thr) hits |= 1u << (g * 4 + k);
        const int mine = __popc(hits);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
        int at = 0;
        if (warp_total) {                                   // warp-uniform
            if (lane == 31) at = atomicAdd(&s_count, warp_total);
            at = __shfl_sync(0xffffffffu, at, 31) + incl - mine;
            // only the element index is staged; the score is re-read (an L2 hit) by the thread that takes the
            // candidate, together with its regression row
            while (hits) {
                const int j = __ffs(hits) - 1;
                hits &= hits - 1u;
                const int e = tile0 + (((j >> 2) * K3_THREADS + tid) << 2) + (j & 3);
                s_elem[at++] = e;
#ifdef K3_PREFETCH_ROWS
                // (A/B for the next sweep, off) request the candidate's regression row now: phase 2 is bound by the number of
                // scattered rows in flight, and this puts them in flight one barrier + one list pass earlier
                if (DECODE && p.C == 1)
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const float4*>(p.reg) + (size_t)b * p.N + e));
#endif
            }
        }
    } else {
        // class-agnostic: score = max over classes, label = first argmax (model/layers.py:234-235)
#pragma unroll
        for (int g = 0; g < K3_VEC; ++g) {
            const int n = blockIdx.x * (K3_THREADS * K3_VEC) + g * K3_THREADS + tid;
            if (n < p.N) {
                const float* s = p.cls + ((size_t)b * p.N + n) * p.C;
                float best = __ldg(s);
                int label = 0;
                for (int c = 1; c < p.C; ++c) { const float v = __ldg(s + c); if (v > best) { best = v; label = c; } }
                if (best > p.thr) {
                    const int at = atomicAdd(&s_count, 1);
                    s_elem[at] = n; s_score[at] = best; s_label[at] = label;
                }
            }
        }
    }
    __syncthreads();
    const int found = s_count;
    if (found == 0) return;
    // slot reservation (one global atomic per CTA when all candidates feed one slab) is issued first and only
    // waited for after the candidates' boxes have been fetched and decoded: the two round trips overlap
    if (one_slab && tid == 0) s_base = atomicAdd(p.sl.counts + b, found);
    for (int i0 = 0; i0 < found; i0 += K3_THREADS) {
        const int i = i0 + tid;
        const bool act = i < found;
        int n = 0, c = 0, seg = b;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        float score = 0.f;
        long long slot = 0;
        if (act) {
            const int e = s_elem[i];
            n = e;
            score = p.class_specific ? __ldg(p.cls + (size_t)b * p.N * p.C + e) : s_score[i];
            if (one_slab) {
                if (!p.class_specific) c = s_label[i];
            } else {
                n = rn_div(e, p.C, p.inv_c);
                c = e - n * p.C;
                seg = b * p.C + c;
                slot = atomicAdd(p.sl.counts + seg, 1);
            }
            box = candidate_box<DECODE>(p, b, n);
        }
        if (one_slab && i0 == 0) __syncthreads();           // s_base has arrived
        if (act) {
            if (one_slab) slot = (long long)s_base + i;
            if (slot < p.sl.cap) {                          // dropped when the slab is full; the count keeps
                const size_t dst = (size_t)seg * p.sl.cap + slot;   // growing so k_segment_nms reports the overflow
                p.sl.keys[dst] = make_key(score, (unsigned)n);
                p.sl.boxes[dst] = box;
                if (p.sl.labels) p.sl.labels[dst] = c;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4 + K5
// ------------------------------------------------------------------------------------------------
struct NmsParams {
    Slabs sl;
    int S;                 // segments
    int segs_per_page;     // C (class specific) or 1
    int direct_boxes;      // 1: sl.boxes is the caller's box array indexed by anchor idx (rn_nms)
    int nms;
    float iou_thr;
    int max_det;
    int pre_nms_top_k;
    // per-segment results
    int* kept_count;               // (S)
    unsigned long long* kept_key;  // (S, max_det)
    float4* kept_box;              // (S, max_det)
    int* kept_label;               // (S, max_det)
    int* status;                   // (pages) or nullptr
    unsigned long long* timing;    // 8 phase counters (clock64 ticks of thread 0, summed over CTAs) or nullptr
};

// TF non_max_suppression_op.cc IOU on corner-normalised boxes with precomputed areas
// Exact fast rejects first (they never change the outcome): a non-positive area or an empty intersection
// gives IoU 0; and since IoU <= min(area)/max(area), boxes whose areas differ by more than the threshold
// allows (0.1 % safety margin over fp32 rounding) cannot exceed it.  Only the remaining pairs pay the divide.
__device__ __forceinline__ bool iou_exceeds(const float4 a, const float aa, const float4 b, const float ab, const float thr) {
    if (thr < 0.0f) {                       // degenerate threshold: fall back to the plain formula
        float iou = 0.0f;
        if (aa > 0.0f && ab > 0.0f) {
            const float iw = fmaxf(fminf(a.z, b.z) - fmaxf(a.x, b.x), 0.0f);
            const float ih = fmaxf(fminf(a.w, b.w) - fmaxf(a.y, b.y), 0.0f);
            const float inter = iw * ih;
            iou = inter / (aa + ab - inter);
        }
        return iou > thr;
    }
    if (!(aa > 0.0f && ab > 0.0f)) return false;
    if (fminf(aa, ab) < thr * fmaxf(aa, ab) * 0.999f) return false;
    const float iw = fminf(a.z, b.z) - fmaxf(a.x, b.x);
    if (!(iw > 0.0f)) return false;
    const float ih = fminf(a.w, b.w) - fmaxf(a.y, b.y);
    if (!(ih > 0.0f)) return false;
    const float inter = iw * ih;
    return inter / (aa + ab - inter) > thr;
}


// Descending bitonic sort of n (power of two, 128 <= n <= NMS_CHUNK) (key, slot) pairs in shared memory.
// Each of the first n/2 threads keeps 2 adjacent elements in registers (all 1024 threads are busy at n = 2048, so
// shuffle and shared-memory latencies overlap across 32 warps): compare-exchange distance 1 is thread-local,
// 2..32 are warp shuffles, distances >= 64 go through shared memory.  Keys are unique (zero padding excepted, which
// never swaps).  Must be called by all threads of the CTA.
__device__ __forceinline__ void ce_keep(unsigned long long& a, unsigned& av, unsigned long long b, unsigned bv, bool take_max) {
    const bool sw = take_max ? (b > a) : (b < a);
    if (sw) { a = b; av = bv; }
}

__device__ void sort_chunk_desc(unsigned long long* s_key, unsigned* s_slot, const int n, const int tid) {
    const bool act = tid < (n >> 1);                 // warp-uniform: n/2 is a multiple of 32
    const int i0 = tid * 2;
    unsigned long long k0 = 0ull, k1 = 0ull;
    unsigned v0 = 0u, v1 = 0u;
    if (act) { k0 = s_key[i0]; k1 = s_key[i0 + 1]; v0 = s_slot[i0]; v1 = s_slot[i0 + 1]; }
    for (int k = 2; k <= n; k <<= 1) {
        const bool desc = ((i0 & k) == 0);           // direction of this thread's pair in the k-merge (k >= 2: same for both)
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 64) {
                if (act) { s_key[i0] = k0; s_key[i0 + 1] = k1; s_slot[i0] = v0; s_slot[i0 + 1] = v1; }
                __syncthreads();
                if (act) {
                    const int x0 = i0 ^ j, x1 = (i0 + 1) ^ j;
                    const bool lower = ((i0 & j) == 0);                  // this thread holds the lower index of each pair
                    ce_keep(k0, v0, s_key[x0], s_slot[x0], lower == desc);
                    ce_keep(k1, v1, s_key[x1], s_slot[x1], lower == desc);
                }
                __syncthreads();
            } else if (j >= 2) {
                if (act) {
                    const int lm = j >> 1;
                    const bool lower = ((i0 & j) == 0);
                    const unsigned long long b0 = __shfl_xor_sync(0xffffffffu, k0, lm), b1 = __shfl_xor_sync(0xffffffffu, k1, lm);
                    const unsigned w0 = __shfl_xor_sync(0xffffffffu, v0, lm), w1 = __shfl_xor_sync(0xffffffffu, v1, lm);
                    ce_keep(k0, v0, b0, w0, lower == desc);
                    ce_keep(k1, v1, b1, w1, lower == desc);
                }
            } else if (act) {
                // j == 1: the pair inside the thread.  For k == 2 the direction alternates per pair ((i0 & 2) == 0).
                const bool d = (k == 2) ? ((i0 & 2) == 0) : desc;
                const bool sw = d ? (k0 < k1) : (k0 > k1);
                if (sw) {
                    const unsigned long long tk = k0; k0 = k1; k1 = tk;
                    const unsigned tv = v0; v0 = v1; v1 = tv;
                }
            }
        }
    }
    if (act) { s_key[i0] = k0; s_key[i0 + 1] = k1; s_slot[i0] = v0; s_slot[i0 + 1] = v1; }
    __syncthreads();
}

__global__ void __launch_bounds__(NMS_THREADS, 1) k_segment_nms(const NmsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // dynamic: selected boxes (normalised corners) + areas, sized by max_det
    float4* s_selbox = reinterpret_cast<float4*>(smem_raw);
    float* s_selarea = reinterpret_cast<float*>(s_selbox + p.max_det);

    __shared__ unsigned long long s_key[NMS_CHUNK];
    __shared__ unsigned s_slot[NMS_CHUNK];
    __shared__ float4 s_cbox[NMS_BATCH];      // corner-normalised
    __shared__ float4 s_craw[NMS_BATCH];      // as stored
    __shared__ float s_carea[NMS_BATCH];
    __shared__ int s_alive[NMS_BATCH];
    __shared__ unsigned s_mask[NMS_BATCH * NMS_WORDS];
    __shared__ unsigned short s_pick[NMS_BATCH];
    __shared__ unsigned s_hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_want, s_loaded, s_nsel, s_done;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int seg = blockIdx.x;
    const int total = p.sl.counts[seg];
    const int cnt = (int)min((long long)total, p.sl.cap);
    if (total > cnt && p.status && tid == 0) p.status[seg / p.segs_per_page] = 1;
    const int limit = p.pre_nms_top_k > 0 ? min(cnt, p.pre_nms_top_k) : cnt;
    const unsigned long long* keys = p.sl.keys + (size_t)seg * p.sl.cap;
    const float4* boxes = p.direct_boxes ? p.sl.boxes : p.sl.boxes + (size_t)seg * p.sl.cap;
    const int* labels = p.sl.labels ? p.sl.labels + (size_t)seg * p.sl.cap : nullptr;
    const int seg_label = seg % p.segs_per_page;

    // optional phase timing (RN_NMS_TIMING=1): thread 0 accumulates clock64() deltas per phase
    long long t_mark = p.timing ? clock64() : 0;
#define RN_PHASE(k) do { if (p.timing && tid == 0) { const long long now = clock64(); atomicAdd(p.timing + (k), (unsigned long long)(now - t_mark)); t_mark = now; } } while (0)
    unsigned long long upper = ~0ull;   // keys >= upper have been visited
    int visited = 0, nsel = 0, round = 0;
    if (tid == 0) s_nsel = 0;
    // The slab's keys are read ONCE into registers (8 per thread) when the slab has <= 8192 candidates -- the
    // radix-select passes and the gathers of every round then run out of registers; larger slabs stream the
    // keys from global memory (L2) in every pass.
    constexpr int KPT = 8;
    const bool in_regs = cnt <= KPT * NMS_THREADS;
    unsigned long long rk[KPT];
#pragma unroll
    for (int t = 0; t < KPT; ++t) {
        const int i = t * NMS_THREADS + tid;
        rk[t] = (in_regs && i < cnt) ? __ldcg(keys + i) : 0ull;      // 0 never matches (real keys are > 0)
    }
    __syncthreads();

    while (visited < limit && nsel < p.max_det) {
        // first round: the NMS_CHUNK best candidates.  Later rounds are only reached when those did not yield max_det
        // selections -- typically a few hundred more are needed, not another 2048 -- so they start at NMS_CHUNK_NEXT
        // and double (a select + sort round costs about the same as consuming 1000 candidates)
        const int take = min(round == 0 ? NMS_CHUNK : min(NMS_CHUNK, NMS_CHUNK_NEXT << (round - 1)), limit - visited);
        ++round;
        // ---------------- K4: radix select the `take` largest unvisited keys --------------------
        unsigned long long thr_key = 0ull;
        if (cnt - visited > take) {
            if (tid == 0) { s_prefix = 0ull; s_want = take; s_done = 0; }
            unsigned long long mask = 0ull;
            for (int shift = 56; shift >= 0; shift -= 8) {
                if (tid < 256) s_hist[tid] = 0u;
                __syncthreads();
                const unsigned long long prefix = s_prefix;
                if (in_regs) {
#pragma unroll
                    for (int t = 0; t < KPT; ++t) {
                        const unsigned long long k = rk[t];
                        if (k != 0ull && k < upper && (k & mask) == prefix) atomicAdd(&s_hist[(unsigned)(k >> shift) & 255u], 1u);
                    }
                } else {
                    for (int i = tid; i < cnt; i += NMS_THREADS) {
                        const unsigned long long k = __ldcg(keys + i);
                        if (k < upper && (k & mask) == prefix) atomicAdd(&s_hist[(unsigned)(k >> shift) & 255u], 1u);
                    }
                }
                __syncthreads();
                if (warp == 0) {
                    // suffix sums over the 256 bins, 8 bins per lane, highest digit first
                    unsigned local[8], run = 0;
#pragma unroll
                    for (int t = 0; t < 8; ++t) { local[t] = s_hist[255 - (lane * 8 + t)]; run += local[t]; }
                    unsigned incl = run;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                    unsigned before = incl - run;          // keys in strictly higher digit groups of earlier lanes
                    const unsigned want = (unsigned)s_want;
                    int digit = -1; unsigned rem = 0; bool whole = false;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        if (digit < 0 && before + local[t] >= want && before < want) {
                            digit = 255 - (lane * 8 + t); rem = want - before;
                            whole = (local[t] == rem);     // every key of this digit group is wanted: lower digits do not matter
                        }
                        before += local[t];
                    }
                    if (digit >= 0) { s_prefix = prefix | ((unsigned long long)digit << shift); s_want = (int)rem; s_done = whole ? 1 : 0; }
                }
                mask |= 255ull << shift;
                __syncthreads();
                if (s_done) break;       // block-uniform (scores are mostly distinct: the low 32 index bits rarely need passes)
            }
            thr_key = s_prefix;          // exactly `take` unvisited keys are >= thr_key (its undecided low digits are 0)
        }
        RN_PHASE(0);
        // ---------------- gather the chunk into shared memory ----------------------------------------
        if (tid == 0) s_loaded = 0;
        __syncthreads();
        if (in_regs) {
#pragma unroll
            for (int t = 0; t < KPT; ++t) {
                const unsigned long long k = rk[t];
                if (k != 0ull && k < upper && k >= thr_key) {
                    const int at = atomicAdd(&s_loaded, 1);
                    if (at < NMS_CHUNK) { s_key[at] = k; s_slot[at] = (unsigned)(t * NMS_THREADS + tid); }
                }
            }
        } else {
            for (int i = tid; i < cnt; i += NMS_THREADS) {
                const unsigned long long k = __ldcg(keys + i);
                if (k < upper && k >= thr_key) {
                    const int at = atomicAdd(&s_loaded, 1);
                    if (at < NMS_CHUNK) { s_key[at] = k; s_slot[at] = (unsigned)i; }
                }
            }
        }
        __syncthreads();
        const int loaded = min(s_loaded, NMS_CHUNK);
        int n2 = 128;
        while (n2 < loaded) n2 <<= 1;
        for (int i = loaded + tid; i < n2; i += NMS_THREADS) { s_key[i] = 0ull; s_slot[i] = 0u; }
        __syncthreads();
        RN_PHASE(1);
        // ---------------- bitonic network, descending (keys are unique) ----------------------------
        sort_chunk_desc(s_key, s_slot, n2, tid);
        const int chunk_n = min(loaded, take);
        RN_PHASE(2);
        // ---------------- K5: greedy NMS over the ordered chunk, 256 candidates per round --------------
        for (int s0 = 0; s0 < chunk_n && nsel < p.max_det; s0 += NMS_BATCH) {
            const int bn = min(NMS_BATCH, chunk_n - s0);
            if (tid < NMS_BATCH) {
                int alive = 0;
                if (tid < bn) {
                    const unsigned slot = s_slot[s0 + tid];
                    const float4 r = p.direct_boxes ? __ldg(boxes + key_idx(s_key[s0 + tid])) : __ldcg(boxes + slot);
                    float4 c;
                    c.x = fminf(r.x, r.z); c.y = fminf(r.y, r.w); c.z = fmaxf(r.x, r.z); c.w = fmaxf(r.y, r.w);
                    s_craw[tid] = r; s_cbox[tid] = c;
                    s_carea[tid] = (c.z - c.x) * (c.w - c.y);
                    alive = 1;
                }
                s_alive[tid] = alive;
            }
            __syncthreads();
            if (p.nms) {
                RN_PHASE(3);
                // (a) against everything selected so far: 2 threads per candidate split the list
                {
                    const int c = tid & (NMS_BATCH - 1), part = tid / NMS_BATCH;
                    if (c < bn) {
                        const float4 cb = s_cbox[c];
                        const float ca = s_carea[c];
                        bool dead = false;
                        for (int s = part; s < nsel && !dead; s += NMS_THREADS / NMS_BATCH)
                            dead = iou_exceeds(cb, ca, s_selbox[s], s_selarea[s], p.iou_thr);
                        if (dead) s_alive[c] = 0;
                    }
                }
                __syncthreads();
                RN_PHASE(4);
                // (b) suppression bit-matrix among the survivors (upper triangle), one ballot per word
                unsigned aw[NMS_WORDS];                     // alive columns, 32 per word (warp-uniform)
#pragma unroll
                for (int w = 0; w < NMS_WORDS; ++w) aw[w] = __ballot_sync(0xffffffffu, s_alive[w * 32 + lane] != 0);
                for (int i = warp; i < bn; i += NMS_WARPS) {
                    if (!s_alive[i]) continue;
                    const float4 bi = s_cbox[i];
                    const float ai = s_carea[i];
                    const int w_first = i >> 5;
#pragma unroll
                    for (int w = 0; w < NMS_WORDS; ++w) {      // unrolled: the 8 words are independent
                        if (w < w_first) continue;
                        unsigned word = 0u;
                        if (aw[w]) {
                            const int j = w * 32 + lane;
                            const bool hit = (j > i) && ((aw[w] >> lane) & 1u) &&
                                             iou_exceeds(bi, ai, s_cbox[j], s_carea[j], p.iou_thr);
                            word = __ballot_sync(0xffffffffu, hit);
                        }
                        if (lane == 0) s_mask[i * NMS_WORDS + w] = word;
                    }
                }
                __syncthreads();
            }
            RN_PHASE(5);
            // (c) one warp walks the surviving bits in order
            if (warp == 0) {
                unsigned mine = 0u;
#pragma unroll
                for (int w = 0; w < NMS_WORDS; ++w) {
                    const unsigned word = __ballot_sync(0xffffffffu, s_alive[w * 32 + lane] != 0);
                    if (lane == w) mine = word;
                }
                int ns = nsel;
                while (ns < p.max_det) {
                    const unsigned has = __ballot_sync(0xffffffffu, mine != 0u) & ((1u << NMS_WORDS) - 1u);
                    if (!has) break;
                    const int w0 = __ffs(has) - 1;
                    const unsigned word = __shfl_sync(0xffffffffu, mine, w0);
                    const int bit = __ffs(word) - 1;
                    const int i = w0 * 32 + bit;
                    if (lane == 0) s_pick[ns - nsel] = (unsigned short)i;
                    ++ns;
                    if (lane == w0) mine &= ~(1u << bit);
                    if (p.nms && lane < NMS_WORDS && lane >= w0) mine &= ~s_mask[i * NMS_WORDS + lane];
                }
                if (lane == 0) s_nsel = ns;
            }
            __syncthreads();
            // the picked candidates join the selected list / the kept arrays, one thread each
            for (int t = tid; t < s_nsel - nsel; t += NMS_THREADS) {
                const int i = s_pick[t], ns = nsel + t;
                s_selbox[ns] = s_cbox[i];
                s_selarea[ns] = s_carea[i];
                const size_t at = (size_t)seg * p.max_det + ns;
                p.kept_key[at] = s_key[s0 + i];
                p.kept_box[at] = s_craw[i];
                p.kept_label[at] = labels ? labels[s_slot[s0 + i]] : seg_label;
            }
            __syncthreads();
            nsel = s_nsel;
            RN_PHASE(6);
        }
        visited += take;
        upper = (loaded > 0) ? s_key[chunk_n - 1] : 0ull;
        __syncthreads();
    }
    if (tid == 0) p.kept_count[seg] = nsel;
#undef RN_PHASE
}

// ------------------------------------------------------------------------------------------------
// per page: merge the per-class kept lists (each already ordered) into the global top max_det
// ------------------------------------------------------------------------------------------------
struct MergeParams {
    int pages, segs_per_page, max_det;
    const int* kept_count;
    const unsigned long long* kept_key;
    const float4* kept_box;
    const int* kept_label;
    float* out_boxes;   // (pages, max_det, 4)
    float* out_scores;  // (pages, max_det)
    int* out_labels;    // (pages, max_det)
    int* out_indices;   // (pages, max_det) or nullptr
    int* out_count;     // (pages) or nullptr
};

__global__ void __launch_bounds__(256) k_merge_topk(const MergeParams p) {
    extern __shared__ int s_head[];                 // (segs_per_page)
    __shared__ unsigned long long s_best[8];
    __shared__ int s_bestc[8];
    __shared__ int s_win;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = p.segs_per_page, M = p.max_det;
    for (int c = tid; c < C; c += 256) s_head[c] = 0;
    __syncthreads();
    int produced = 0;
    if (C == 1) {
        const int n = min(p.kept_count[b], M);
        for (int m = tid; m < n; m += 256) {
            const size_t at = (size_t)b * M + m;
            const unsigned long long k = p.kept_key[at];
            reinterpret_cast<float4*>(p.out_boxes)[at] = p.kept_box[at];
            p.out_scores[at] = key_score(k);
            p.out_labels[at] = p.kept_label[at];
            if (p.out_indices) p.out_indices[at] = (int)key_idx(k);
        }
        produced = n;
    } else {
        for (int m = 0; m < M; ++m) {
            // every thread proposes the best head among its classes: (score desc, class asc)
            unsigned long long best = 0ull; int bc = -1;
            for (int c = tid; c < C; c += 256) {
                const int h = s_head[c];
                if (h < p.kept_count[b * C + c]) {
                    const unsigned long long k = p.kept_key[((size_t)b * C + c) * M + h];
                    const unsigned long long cand = (k & 0xffffffff00000000ull) | (unsigned long long)(0xffffffffu - (unsigned)c);
                    if (cand > best) { best = cand; bc = c; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
                if (ob > best) { best = ob; bc = oc; }
            }
            if (lane == 0) { s_best[warp] = best; s_bestc[warp] = bc; }
            __syncthreads();
            if (tid == 0) {
                unsigned long long bb = 0ull; int cc = -1;
                for (int w = 0; w < 8; ++w) if (s_best[w] > bb) { bb = s_best[w]; cc = s_bestc[w]; }
                s_win = cc;
                if (cc >= 0) {
                    const int h = s_head[cc]++;
                    const size_t src = ((size_t)b * C + cc) * M + h;
                    const size_t at = (size_t)b * M + m;
                    const unsigned long long k = p.kept_key[src];
                    reinterpret_cast<float4*>(p.out_boxes)[at] = p.kept_box[src];
                    p.out_scores[at] = key_score(k);
                    p.out_labels[at] = p.kept_label[src];
                    if (p.out_indices) p.out_indices[at] = (int)key_idx(k);
                }
            }
            __syncthreads();
            if (s_win < 0) break;
            ++produced;
        }
    }
    for (int m = produced + tid; m < M; m += 256) {
        const size_t at = (size_t)b * M + m;
        reinterpret_cast<float4*>(p.out_boxes)[at] = make_float4(-1.f, -1.f, -1.f, -1.f);
        p.out_scores[at] = -1.0f;
        p.out_labels[at] = -1;
        if (p.out_indices) p.out_indices[at] = -1;
    }
    if (p.out_count && tid == 0) p.out_count[b] = produced;
}

__global__ void k_keys_from_scores(const float* scores, long long K, unsigned long long* keys, int* count) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < K; i += (long long)gridDim.x * blockDim.x)
        keys[i] = make_key(scores[i], (unsigned)i);
    if (blockIdx.x == 0 && threadIdx.x == 0) *count = (int)K;
}

// ------------------------------------------------------------------------------------------------
// workspace carving (256-byte aligned sections)
// ------------------------------------------------------------------------------------------------
struct FilterWs {
    unsigned long long* timing;
    int* counts; int* kept_count; int* status;
    unsigned long long* keys; float4* boxes; int* labels;
    unsigned long long* kept_key; float4* kept_box; int* kept_label;
    size_t bytes;
};

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

FilterWs carve(void* ws, int B, int S, long long cap, int max_det, bool agnostic) {
    FilterWs w;
    size_t off = 0;
    char* base = reinterpret_cast<char*>(ws);
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align256(bytes); return p; };
    w.timing = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * 8));   // first 64 bytes
    w.counts = reinterpret_cast<int*>(take(sizeof(int) * S));
    w.kept_count = reinterpret_cast<int*>(take(sizeof(int) * S));
    w.status = reinterpret_cast<int*>(take(sizeof(int) * B));
    w.keys = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * (size_t)S * cap));
    w.boxes = reinterpret_cast<float4*>(take(sizeof(float4) * (size_t)S * cap));
    w.labels = agnostic ? reinterpret_cast<int*>(take(sizeof(int) * (size_t)S * cap)) : nullptr;
    w.kept_key = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * (size_t)S * max_det));
    w.kept_box = reinterpret_cast<float4*>(take(sizeof(float4) * (size_t)S * max_det));
    w.kept_label = reinterpret_cast<int*>(take(sizeof(int) * (size_t)S * max_det));
    w.bytes = off;
    return w;
}

size_t nms_dynamic_smem(int max_det) { return (size_t)max_det * (sizeof(float4) + sizeof(float)); }

int run_back_end(const FilterWs& w, int B, int S, int segs_per_page, long long cap, int direct_boxes, const float4* direct,
                 int nms, float nms_thr, int max_det, int pre_nms_top_k,
                 float* out_boxes, float* out_scores, int* out_labels, int* out_indices, int* out_count,
                 int* status, cudaStream_t s) {
    NmsParams np;
    np.sl.counts = w.counts; np.sl.keys = w.keys; np.sl.boxes = direct_boxes ? const_cast<float4*>(direct) : w.boxes;
    np.sl.labels = w.labels; np.sl.cap = cap;
    np.S = S; np.segs_per_page = segs_per_page; np.direct_boxes = direct_boxes; np.nms = nms; np.iou_thr = nms_thr;
    np.max_det = max_det; np.pre_nms_top_k = pre_nms_top_k;
    np.kept_count = w.kept_count; np.kept_key = w.kept_key; np.kept_box = w.kept_box; np.kept_label = w.kept_label;
    np.status = status;
    static const bool timing_on = getenv("RN_NMS_TIMING") != nullptr;
    np.timing = timing_on ? w.timing : nullptr;
    const size_t dyn = nms_dynamic_smem(max_det);
    // static (~43 KB) + dynamic shared memory exceeds the 48 KB default: opt in (227 KB per CTA on sm_100a)
    cudaError_t ae = cudaFuncSetAttribute(k_segment_nms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nms_dynamic_smem(MAX_DET_LIMIT));
    if (ae != cudaSuccess) return rn_fail(RN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(ae));
    k_segment_nms<<<S, NMS_THREADS, dyn, s>>>(np);
    int rc = rn_check_launch("k_segment_nms");
    if (rc) return rc;
    MergeParams mp;
    mp.pages = B; mp.segs_per_page = segs_per_page; mp.max_det = max_det;
    mp.kept_count = w.kept_count; mp.kept_key = w.kept_key; mp.kept_box = w.kept_box; mp.kept_label = w.kept_label;
    mp.out_boxes = out_boxes; mp.out_scores = out_scores; mp.out_labels = out_labels; mp.out_indices = out_indices;
    mp.out_count = out_count;
    k_merge_topk<<<B, 256, sizeof(int) * (size_t)segs_per_page, s>>>(mp);
    return rn_check_launch("k_merge_topk");
}

int filter_common(K3Params kp, bool decode, int nms, float nms_thr, int max_det, int pre_nms_top_k, long long cand_cap,
                  float* out_boxes, float* out_scores, int* out_labels, int* out_indices, int* status_out,
                  void* workspace, size_t workspace_bytes, cudaStream_t s) {
    const int B = kp.B, C = kp.C;
    RN_REQUIRE(B >= 1 && kp.N >= 1 && C >= 1, "bad shape");
    RN_REQUIRE(max_det >= 1 && max_det <= MAX_DET_LIMIT, "max_detections must be in [1, %d]", MAX_DET_LIMIT);
    RN_REQUIRE(cand_cap >= 1, "cand_cap must be >= 1");
    RN_REQUIRE(out_boxes && out_scores && out_labels && workspace, "NULL pointer");
    RN_REQUIRE(rn_aligned16(out_boxes) && rn_aligned16(workspace), "out_boxes / workspace must be 16-byte aligned");
    RN_REQUIRE(rn_aligned16(kp.cls), "classification must be 16-byte aligned");
    RN_REQUIRE(pre_nms_top_k >= 0, "pre_nms_top_k must be >= 0");
    if (cand_cap > kp.N) cand_cap = kp.N;
    const int spp = kp.class_specific ? C : 1;
    const long long S64 = (long long)B * spp;
    RN_REQUIRE(S64 < (1ll << 30), "too many (page, class) segments");
    const int S = (int)S64;
    FilterWs w = carve(workspace, B, S, cand_cap, max_det, !kp.class_specific);
    if (workspace_bytes < w.bytes) return rn_fail(RN_ERR_WORKSPACE, "filter workspace too small: %zu < %zu", workspace_bytes, w.bytes);
    // counts, kept_count, status are contiguous at the front of the workspace
    cudaError_t e = cudaMemsetAsync(w.timing, 0, (size_t)((char*)w.keys - (char*)w.timing), s);
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
    if (status_out) {
        e = cudaMemsetAsync(status_out, 0, sizeof(int) * (size_t)B, s);
        if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
    }
    kp.sl.counts = w.counts; kp.sl.keys = w.keys; kp.sl.boxes = w.boxes; kp.sl.labels = w.labels; kp.sl.cap = cand_cap;
    RN_REQUIRE((long long)kp.N * C < (1ll << 31) - 4096, "N * C too large for one page");
    RN_REQUIRE(B <= 65535, "B must be <= 65535");
    kp.inv_c = 1.0f / (float)C;
    kp.vec_ok = (((long long)kp.N * C) % 4 == 0) ? 1 : 0;
    const long long tiles = kp.class_specific ? ((long long)kp.N * C + K3_TILE - 1) / K3_TILE
                                              : ((long long)kp.N + K3_THREADS * K3_VEC - 1) / (K3_THREADS * K3_VEC);
    const dim3 grid((unsigned)tiles, (unsigned)B);
    if (decode) k_threshold_compact<true><<<grid, K3_THREADS, 0, s>>>(kp);
    else k_threshold_compact<false><<<grid, K3_THREADS, 0, s>>>(kp);
    int rc = rn_check_launch("k_threshold_compact");
    if (rc) return rc;
    return run_back_end(w, B, S, spp, cand_cap, 0, nullptr, nms, nms_thr, max_det, pre_nms_top_k,
                        out_boxes, out_scores, out_labels, out_indices, nullptr,
                        status_out ? status_out : w.status, s);
}

}  // namespace

extern "C" size_t rn_filter_workspace_bytes(int B, long long N, int C, int class_specific, long long cand_cap, int max_detections) {
    if (B < 1 || N < 1 || C < 1 || max_detections < 1) return 0;
    if (cand_cap < 1 || cand_cap > N) cand_cap = N;
    const long long S = (long long)B * (class_specific ? C : 1);
    return carve(nullptr, B, (int)S, cand_cap, max_detections, !class_specific).bytes;
}

extern "C" int rn_filter_detections(const float* boxes, const float* classification,
                                    int B, long long N, int C, int class_specific, int nms,
                                    float score_threshold, float nms_threshold, int max_detections,
                                    int pre_nms_top_k, long long cand_cap,
                                    float* out_boxes, float* out_scores, int* out_labels, int* out_indices,
                                    int* status_out_dev, void* workspace, size_t workspace_bytes, void* stream) {
    RN_REQUIRE(boxes && classification, "NULL input");
    RN_REQUIRE(rn_aligned16(boxes), "boxes must be 16-byte aligned");
    RN_REQUIRE(N < (1ll << 31), "N too large");
    K3Params kp = {};
    kp.cls = classification; kp.boxes = boxes; kp.B = B; kp.N = (int)N; kp.C = C;
    kp.class_specific = class_specific ? 1 : 0; kp.thr = score_threshold;
    return filter_common(kp, false, nms ? 1 : 0, nms_threshold, max_detections, pre_nms_top_k, cand_cap,
                         out_boxes, out_scores, out_labels, out_indices, status_out_dev,
                         workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int rn_decode_filter_detections(const float* base_anchors_f32_dev, const int* level_hw,
                                           const int* level_stride, int num_levels, int anchors_per_cell,
                                           const float* regression, const float* classification,
                                           int B, long long N, int C,
                                           const float* mean4, const float* std4, float clip_width, float clip_height,
                                           int class_specific, int nms,
                                           float score_threshold, float nms_threshold, int max_detections,
                                           int pre_nms_top_k, long long cand_cap,
                                           float* out_boxes, float* out_scores, int* out_labels, int* out_indices,
                                           int* status_out_dev, void* workspace, size_t workspace_bytes, void* stream) {
    RN_REQUIRE(base_anchors_f32_dev && regression && classification && mean4 && std4, "NULL input");
    RN_REQUIRE(rn_aligned16(regression), "regression must be 16-byte aligned");
    K3Params kp = {};
    int rc = rn_make_levels(&kp.lv, level_hw, level_stride, num_levels, anchors_per_cell);
    if (rc) return rc;
    RN_REQUIRE(kp.lv.start[num_levels] == N, "N (%lld) does not match the level table (%d)", N, kp.lv.start[num_levels]);
    kp.cls = classification; kp.reg = regression; kp.base32 = base_anchors_f32_dev;
    for (int i = 0; i < 4; ++i) { kp.nm.mean[i] = mean4[i]; kp.nm.std[i] = std4[i]; }
    kp.clipW = clip_width; kp.clipH = clip_height;
    kp.B = B; kp.N = (int)N; kp.C = C; kp.class_specific = class_specific ? 1 : 0; kp.thr = score_threshold;
    return filter_common(kp, true, nms ? 1 : 0, nms_threshold, max_detections, pre_nms_top_k, cand_cap,
                         out_boxes, out_scores, out_labels, out_indices, status_out_dev,
                         workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" size_t rn_nms_workspace_bytes(long long K, int max_output) {
    if (K < 1 || max_output < 1) return 256;
    FilterWs w = carve(nullptr, 1, 1, K, max_output, false);
    // + scratch outputs of the merge stage (boxes, scores, labels)
    return w.bytes + align256(sizeof(float4) * (size_t)max_output) + 2 * align256(sizeof(float) * (size_t)max_output);
}

extern "C" int rn_nms(const float* boxes, const float* scores, long long K, int max_output, float iou_threshold,
                      int* out_indices, int* out_count_dev, void* workspace, size_t workspace_bytes, void* stream) {
    RN_REQUIRE(out_indices && out_count_dev && workspace, "NULL pointer");
    RN_REQUIRE(K >= 0 && K < (1ll << 31), "K out of range");
    RN_REQUIRE(max_output >= 1 && max_output <= MAX_DET_LIMIT, "max_output must be in [1, %d]", MAX_DET_LIMIT);
    cudaStream_t s = (cudaStream_t)stream;
    if (workspace_bytes < rn_nms_workspace_bytes(K, max_output)) return rn_fail(RN_ERR_WORKSPACE, "nms workspace too small");
    RN_REQUIRE(rn_aligned16(workspace), "workspace must be 16-byte aligned");
    const long long cap = K < 1 ? 1 : K;
    FilterWs w = carve(workspace, 1, 1, cap, max_output, false);
    char* tail = reinterpret_cast<char*>(workspace) + w.bytes;
    float* sc_boxes = reinterpret_cast<float*>(tail); tail += align256(sizeof(float4) * (size_t)max_output);
    float* sc_scores = reinterpret_cast<float*>(tail); tail += align256(sizeof(float) * (size_t)max_output);
    int* sc_labels = reinterpret_cast<int*>(tail);
    cudaError_t e = cudaMemsetAsync(w.timing, 0, (size_t)((char*)w.keys - (char*)w.timing), s);
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
    if (K > 0) {
        RN_REQUIRE(boxes && scores, "NULL input");
        RN_REQUIRE(rn_aligned16(boxes), "boxes must be 16-byte aligned");
        const int blocks = (int)min((K + 255) / 256, (long long)RN_NUM_SMS * 4);
        k_keys_from_scores<<<blocks, 256, 0, s>>>(scores, K, w.keys, w.counts);
        int rc = rn_check_launch("k_keys_from_scores");
        if (rc) return rc;
    }
    return run_back_end(w, 1, 1, 1, cap, 1, reinterpret_cast<const float4*>(boxes), 1, iou_threshold, max_output, 0,
                        sc_boxes, sc_scores, sc_labels, out_indices, out_count_dev, w.status, s);
}
