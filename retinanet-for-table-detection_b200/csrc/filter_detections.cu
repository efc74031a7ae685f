// K3 + K4 + K5: FilterDetections (reference model/layers.py:177-264, :298-332) for a whole batch,
// no host synchronisation anywhere.
//
//   K3  k_threshold_keys      ONE streaming pass over classification (B,N,C): score > thr (strict, fp32); the survivors'
//                             64-bit sort keys (score bits | ~anchor index) are appended to the (page,class) slab --
//                             nothing else is read or written: the kernel is a pure HBM stream of the scores.
//   K4  approximate select    inside k_segment_nms: a 4-way bisection on the key VALUE (three pivots per pass, counts by
//       + bitonic sort        comparison and __reduce_add_sync -- no histogram, no same-address atomics) finds a threshold
//                             that keeps between 1536 and 2048 of the unvisited keys; a bitonic network orders them.
//   K5  bitmask NMS           same kernel.  The boxes of the ordered candidates are fetched LAZILY, 256 at a time and one
//                             group ahead of their use: only candidates that are actually visited are ever decoded
//                             (RegressBoxes + ClipBoxes fused, anchors generated in-kernel in fp32 like the Anchors layer)
//                             or gathered from a dense box tensor.  Greedy NMS runs 32 candidates at a time: warp i owns
//                             candidate i -- its lanes split the list of boxes selected so far, one ballot gives the
//                             32-bit row of the in-batch suppression matrix -- and one warp resolves the order in registers
//                             (no loop at all when the batch has no internal conflict).  Stops at max_detections.
//   k_merge_topk              per page: class-major concatenation + tf.nn.top_k (ties -> earlier position) done as a
//                             C-way merge of the per-class lists; pad with -1.
//
// Semantics restated from tf.image.non_max_suppression / tf.nn.top_k (third-party, un-pinned): see oracle/layers_np.py.
// Compiled with -fmad=false: IoU and decode are evaluated in the reference's fp32 operation order.
#include "rn_common.cuh"
#include <atomic>
#include <string.h>

namespace {

constexpr int K3_THREADS = 256;
constexpr int NMS_THREADS = 1024;
constexpr int NMS_WARPS = NMS_THREADS / 32;
constexpr int NMS_CHUNK = 2048;    // candidates ordered per round (shared-memory capacity of the sorted chunk)
constexpr int NMS_GROUP = 256;     // boxes fetched / decoded per step, one group ahead of the greedy loop
constexpr int NMS_BASE_MAX = 96;  // base anchors staged in shared memory (levels x anchors per cell; 45 by default)
constexpr int NMS_BATCH = 64;      // candidates resolved per greedy step: warp w <-> candidates w and w + 32
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
    int* labels;                 // (S, cap)   only for class-agnostic filtering
    long long cap;
};

// ------------------------------------------------------------------------------------------------
// K3: threshold + key compaction
// ------------------------------------------------------------------------------------------------
struct K3Params {
    const float* cls;      // (B, N, C)
    int B, N, C;
    int class_specific;
    float thr;
    float inv_c;           // 1 / C for rn_div
    int vec_ok;            // page rows of `cls` are 16-byte aligned
    Slabs sl;
};

// grid = (tiles of a page, pages).  Two phases per CTA:
//   1. every thread streams K3_VEC float4 groups of scores (all loads issued up front), the survivors of `score > thr` are
//      ranked with a warp prefix sum and pushed into a shared-memory list (ONE shared-memory atomic per warp);
//   2. the list is consumed densely, one candidate per thread: slot reservation (ONE global atomic per CTA when all
//      candidates feed the same slab, i.e. C == 1 or class-agnostic; one per candidate otherwise) and the key store.
// Algorithmic bytes: 4 per score read + 8 per candidate written (~2.5 % of the scores).
constexpr int K3_VEC = 4;                                   // float4 groups per thread
constexpr int K3_TILE = K3_THREADS * K3_VEC * 4;            // scores per CTA (class-specific path)

__global__ void __launch_bounds__(K3_THREADS) k_threshold_keys(const K3Params p) {
    __shared__ int s_elem[K3_TILE];
    __shared__ float s_score[K3_THREADS * K3_VEC];          // class-agnostic path only (score = max over classes)
    __shared__ int s_label[K3_THREADS * K3_VEC];            // class-agnostic path only
    __shared__ int s_count, s_base;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) s_count = 0;
    __syncthreads();
    const bool one_slab = (p.C == 1) || !p.class_specific;
    const int total = p.N * p.C;                            // scores of this page, < 2^31 (checked on the host)
    if (p.class_specific) {
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
        if (warp_total) {                                   // warp-uniform
            int at = 0;
            if (lane == 31) at = atomicAdd(&s_count, warp_total);
            at = __shfl_sync(0xffffffffu, at, 31) + incl - mine;
            // only the element index is staged; the score is re-read (an L2 hit) by the thread that stores the key
            while (hits) {
                const int j = __ffs(hits) - 1;
                hits &= hits - 1u;
                s_elem[at++] = tile0 + (((j >> 2) * K3_THREADS + tid) << 2) + (j & 3);
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
    if (found == 0) return;                                 // block-uniform
    if (one_slab) {
        if (tid == 0) s_base = atomicAdd(p.sl.counts + b, found);
        __syncthreads();
    }
    for (int i = tid; i < found; i += K3_THREADS) {
        const int e = s_elem[i];
        int n = e, c = 0, seg = b;
        const float score = p.class_specific ? __ldg(p.cls + (size_t)b * total + e) : s_score[i];
        long long slot;
        if (one_slab) {
            if (!p.class_specific) c = s_label[i];
            slot = (long long)s_base + i;
        } else {
            n = rn_div(e, p.C, p.inv_c);
            c = e - n * p.C;
            seg = b * p.C + c;
            slot = atomicAdd(p.sl.counts + seg, 1);
        }
        if (slot < p.sl.cap) {                              // dropped when the slab is full; the count keeps
            const size_t dst = (size_t)seg * p.sl.cap + slot;   // growing so k_segment_nms reports the overflow
            p.sl.keys[dst] = make_key(score, (unsigned)n);
            if (p.sl.labels) p.sl.labels[dst] = c;
        }
    }
}

// The class-specific path (the reference's default) when page rows are 16-byte aligned: a warp-autonomous stream.
// grid = (CTAs per page, pages), about SMs x 4 CTAs in total; a CTA owns a contiguous slice of one page's scores, its 8 warps
// interleave warp-tiles (32 lanes x 4 float4 = 512 consecutive scores) of the slice, each with the four loads of its next
// tile already in flight while it handles the current one -- no block barrier anywhere.  Survivors are ranked inside the
// warp with five ballots (a lane has 0..16 of them) and collected in a warp-private shared-memory buffer; ONE global atomic
// per flush reserves their slab slots and the keys leave with coalesced stores.  (Measured on the way here: one global atomic
// per 128 scores serialises on the page's counter -- 69 us; 128-score tiles are instruction bound -- 143 warp-instructions
// per tile, 23 us; a two-deep cp.async ring per warp instead of registers -- two tiles, 128 KB per SM, in flight -- 19 us instead
// of 17: the bytes in flight are not what limits it.)  C > 1: the candidates of a warp feed different slabs, one atomic each.
// The slab order is arbitrary by design: keys are unique, the NMS kernel orders them.
constexpr int K3S_VEC = 4;                                  // float4 per lane and tile
constexpr int K3S_TILE = 128 * K3S_VEC;                     // scores per warp-tile
constexpr int K3S_CTAS_PER_SM = 4;
constexpr int K3S_BUF = 256;                                // keys a warp buffers before it flushes

__device__ __forceinline__ void k3s_flush(const K3Params& p, int page, const unsigned long long* buf, int n, int lane) {
    __syncwarp();
    int base = 0;
    if (lane == 0) base = atomicAdd(p.sl.counts + page, n);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int i = lane; i < n; i += 32)                      // dropped when the slab is full; the count keeps growing (overflow is reported)
        if ((long long)base + i < p.sl.cap) p.sl.keys[(size_t)page * p.sl.cap + base + i] = buf[i];
    __syncwarp();
}

// sc[i] for a run-time i in [0, 16) without dynamic register indexing (which would put the array in local memory): a
// binary tree of 15 selects
__device__ __forceinline__ float k3s_pick(const float (&sc)[16], int i) {
    float a[8], b[4], c[2];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = (i & 1) ? sc[2 * k + 1] : sc[2 * k];
#pragma unroll
    for (int k = 0; k < 4; ++k) b[k] = (i & 2) ? a[2 * k + 1] : a[2 * k];
#pragma unroll
    for (int k = 0; k < 2; ++k) c[k] = (i & 4) ? b[2 * k + 1] : b[2 * k];
    return (i & 8) ? c[1] : c[0];
}

__global__ void __launch_bounds__(K3_THREADS, K3S_CTAS_PER_SM) k_threshold_keys_stream(const K3Params p, int tiles_per_page, int tiles_per_cta) {
    __shared__ unsigned long long s_buf[K3_THREADS / 32][K3S_BUF];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int page = blockIdx.y;
    const int total = p.N * p.C;                            // scores of one page
    const float* src = p.cls + (size_t)page * total;
    const int t_begin = blockIdx.x * tiles_per_cta, t_end = min(tiles_per_page, t_begin + tiles_per_cta);
    unsigned long long* buf = s_buf[warp];
    const float ninf = __int_as_float(0xff800000);          // out of range: never above any threshold
    int fill = 0;                                           // warp-uniform
    float4 nx[K3S_VEC];
    auto load = [&](int t) {
#pragma unroll
        for (int g = 0; g < K3S_VEC; ++g) {
            nx[g] = make_float4(ninf, ninf, ninf, ninf);
            if (t < t_end) {
                const int e = t * K3S_TILE + g * 128 + lane * 4;
                if (e < total) nx[g] = rn_ldg_stream4(src + e);     // total % 4 == 0: all four or none
            }
        }
    };
    int t = t_begin + warp;
    load(t);
    rn_grid_dependency_wait();                              // the slab counters are being zeroed by the launch in front (first loads already in flight)
    for (; t < t_end; t += K3_THREADS / 32) {
        float sc[K3S_VEC * 4];
#pragma unroll
        for (int g = 0; g < K3S_VEC; ++g) { sc[4 * g] = nx[g].x; sc[4 * g + 1] = nx[g].y; sc[4 * g + 2] = nx[g].z; sc[4 * g + 3] = nx[g].w; }
        load(t + K3_THREADS / 32);                          // the next tile's loads are in flight while this one is handled
        unsigned hits = 0u;                                 // bit 4 g + k: score k of float4 g
#pragma unroll
        for (int i = 0; i < K3S_VEC * 4; ++i) hits |= (sc[i] > p.thr) ? (1u << i) : 0u;
        if (!__any_sync(0xffffffffu, hits != 0u)) continue;  // warp-uniform
        const int e0 = t * K3S_TILE + lane * 4;              // element of sc[4 g + k]: e0 + 128 g + k
        if (p.C == 1) {
            // exclusive prefix of the lanes' survivor counts (0..16 = 5 bits): five ballots, no shuffles
            const unsigned mine = (unsigned)__popc(hits), lt = (1u << lane) - 1u;
            int at = 0, tot = 0;
#pragma unroll
            for (int bit = 0; bit < 5; ++bit) {
                const unsigned bb = __ballot_sync(0xffffffffu, (mine >> bit) & 1u);
                at += __popc(bb & lt) << bit;
                tot += __popc(bb) << bit;
            }
            if (fill + tot > K3S_BUF) { k3s_flush(p, page, buf, fill, lane); fill = 0; }
            if (tot > K3S_BUF) {
                // a dense tile: slots reserved for it alone, keys stored straight from the registers
                int base = 0;
                if (lane == 0) base = atomicAdd(p.sl.counts + page, tot);
                at += __shfl_sync(0xffffffffu, base, 0);
                for (unsigned h = hits; h != 0u; h &= h - 1u) {
                    const int i = __ffs(h) - 1;
                    if (at < p.sl.cap) p.sl.keys[(size_t)page * p.sl.cap + at] = make_key(k3s_pick(sc, i), (unsigned)(e0 + 128 * (i >> 2) + (i & 3)));
                    ++at;
                }
            } else {
                // survivors are sparse (~0.4 per lane and tile): a divergent loop over the set bits costs far fewer issue slots
                // than 16 predicated store blocks
                at += fill;
                for (unsigned h = hits; h != 0u; h &= h - 1u) {
                    const int i = __ffs(h) - 1;
                    buf[at++] = make_key(k3s_pick(sc, i), (unsigned)(e0 + 128 * (i >> 2) + (i & 3)));
                }
                fill += tot;
            }
        } else {
            for (unsigned h = hits; h != 0u; h &= h - 1u) {
                const int i = __ffs(h) - 1;
                const int e = e0 + 128 * (i >> 2) + (i & 3);
                const int n = rn_div(e, p.C, p.inv_c);
                const int seg = page * p.C + (e - n * p.C);
                const long long slot = atomicAdd(p.sl.counts + seg, 1);
                if (slot < p.sl.cap) p.sl.keys[(size_t)seg * p.sl.cap + slot] = make_key(k3s_pick(sc, i), (unsigned)n);
            }
        }
    }
    if (fill) k3s_flush(p, page, buf, fill, lane);
}

// The class-specific path for SEVERAL classes (C > 1, page rows 16-byte aligned).  Neighbouring scores belong to different
// classes here, so the candidates of a tile feed up to C different slabs, and in the warp-autonomous kernel above every
// candidate waits for the return of an atomic on its slab's counter: at C = 80, 5.4 M atomics per 16 pages on 1280 counters, a
// few of which (the classes of a page's tables) take most of them -- same-address atomics serialise in L2.  ncu's source view
// puts 80 % of that launch's stall samples on the atomic (394 us, 0.42 of the HBM peak).  This form collects the keys PER CLASS
// in shared memory instead: a CTA walks its slice of a page in rounds of one 512-score tile per warp, the next round's tile
// already in flight; survivors are appended to their class' list with shared-memory atomics; between two block barriers the
// lists that are at least half full (after the last round: all) are flushed -- a warp owns every eighth list, its lanes issue
// the atomics of all its due lists AT ONCE (one each, reserving the whole run), then the runs leave as coalesced stores.  A
// list that overflows inside a round sends the surplus straight to the slab.
constexpr int K3C_CAP = 32;                                 // keys per class list (one warp-wide store)
constexpr int K3C_FLUSH = 16;                               // a list this full is flushed at the end of the round
constexpr int K3C_MAX_C = 160;                              // C * (CAP * 8 + 4) + flags must fit the 48 KB default

__global__ void __launch_bounds__(K3_THREADS, 4) k_threshold_keys_classes(const K3Params p, int tiles_per_page, int tiles_per_cta) {
    extern __shared__ __align__(16) unsigned char k3c_smem[];
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(k3c_smem);            // [C][CAP]
    int* s_fill = reinterpret_cast<int*>(s_keys + (size_t)p.C * K3C_CAP);                   // [C] appended so far (may exceed CAP)
    unsigned* s_need = reinterpret_cast<unsigned*>(s_fill + p.C);                            // [(C + 31) / 32] lists due this round
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int page = blockIdx.y, C = p.C;
    const int total = p.N * C;
    const float* src = p.cls + (size_t)page * total;
    const int t_begin = blockIdx.x * tiles_per_cta, t_end = min(tiles_per_page, t_begin + tiles_per_cta);
    const int nwords = (C + 31) >> 5;                       // <= 5
    constexpr int STEP = K3_THREADS / 32;
    const float ninf = __int_as_float(0xff800000);          // out of range: never above any threshold
    float4 nx[K3S_VEC];
    auto load = [&](int t) {
#pragma unroll
        for (int g = 0; g < K3S_VEC; ++g) {
            nx[g] = make_float4(ninf, ninf, ninf, ninf);
            if (t < t_end) {
                const int e = t * K3S_TILE + g * 128 + lane * 4;
                if (e < total) nx[g] = rn_ldg_stream4(src + e);     // total % 4 == 0: all four or none
            }
        }
    };
    load(t_begin + warp);
    for (int i = threadIdx.x; i < C; i += K3_THREADS) s_fill[i] = 0;
    for (int i = threadIdx.x; i < nwords; i += K3_THREADS) s_need[i] = 0u;
    rn_grid_dependency_wait();                              // the slab counters are being zeroed by the launch in front
    __syncthreads();
    const unsigned own = 0x01010101u << warp;               // this warp's lists of a 32-list word: bits warp, warp + 8, ...
    for (int t0 = t_begin; t0 < t_end; t0 += STEP) {        // block-uniform rounds
        const int t = t0 + warp;
        float sc[K3S_VEC * 4];
#pragma unroll
        for (int g = 0; g < K3S_VEC; ++g) { sc[4 * g] = nx[g].x; sc[4 * g + 1] = nx[g].y; sc[4 * g + 2] = nx[g].z; sc[4 * g + 3] = nx[g].w; }
        load(t + STEP);                                     // the next round's tile is in flight through this round's barriers
        unsigned hits = 0u;
#pragma unroll
        for (int i = 0; i < K3S_VEC * 4; ++i) hits |= (sc[i] > p.thr) ? (1u << i) : 0u;
        const int e0 = t * K3S_TILE + lane * 4;
        for (unsigned h = hits; h != 0u; h &= h - 1u) {
            const int i = __ffs(h) - 1;
            const int e = e0 + 128 * (i >> 2) + (i & 3);
            const int n = rn_div(e, C, p.inv_c);
            const int c = e - n * C;
            const unsigned long long key = make_key(k3s_pick(sc, i), (unsigned)n);
            const int slot = atomicAdd(&s_fill[c], 1);
            if (slot < K3C_CAP) {
                s_keys[(size_t)c * K3C_CAP + slot] = key;
                if (slot == K3C_FLUSH - 1) atomicOr(&s_need[c >> 5], 1u << (c & 31));
            } else {                                        // the list is full until the end of the round: straight to the slab
                const int seg = page * C + c;
                const long long gs = atomicAdd(p.sl.counts + seg, 1);
                if (gs < p.sl.cap) p.sl.keys[(size_t)seg * p.sl.cap + gs] = key;
            }
        }
        __syncthreads();                                    // the round's appends are complete
        const bool last_round = t0 + STEP >= t_end;
        // this warp's due lists: lane j < 4 * nwords looks at list (word j / 4, the (j % 4)-th of the warp's four bits)
        int c_mine = -1, n_mine = 0, base_mine = 0;
        if (lane < 4 * nwords) {
            const int w = lane >> 2, b = warp + 8 * (lane & 3), c = w * 32 + b;
            const bool due = last_round || ((s_need[w] >> b) & 1u);
            if (due && c < C) {
                n_mine = min(s_fill[c], K3C_CAP);
                if (n_mine > 0) {
                    c_mine = c;
                    base_mine = atomicAdd(p.sl.counts + page * C + c, n_mine);      // all of the warp's due lists at once
                }
            }
        }
        unsigned due_lanes = __ballot_sync(0xffffffffu, c_mine >= 0);
        for (; due_lanes != 0u; due_lanes &= due_lanes - 1u) {
            const int j = __ffs(due_lanes) - 1;
            const int c = __shfl_sync(0xffffffffu, c_mine, j), n = __shfl_sync(0xffffffffu, n_mine, j);
            const int base = __shfl_sync(0xffffffffu, base_mine, j);
            const size_t seg = (size_t)page * C + c;
            if (lane < n && (long long)base + lane < p.sl.cap) p.sl.keys[seg * p.sl.cap + base + lane] = s_keys[(size_t)c * K3C_CAP + lane];
        }
        __syncwarp();
        if (c_mine >= 0) s_fill[c_mine] = 0;
        if (lane < nwords) atomicAnd(&s_need[lane], ~own);   // this warp's flags of every word
        __syncthreads();                                    // flushed lists are empty, their flags cleared
    }
}

// ------------------------------------------------------------------------------------------------
// where a candidate's box comes from: a dense (pages, N, 4) tensor, or the fused decode (Anchors + RegressBoxes + ClipBoxes)
// ------------------------------------------------------------------------------------------------
struct BoxSource {
    const float* rows;     // (pages, N, 4): boxes (dense mode) or regression deltas (decode mode)
    const float* base32;   // (L, A, 4) float32 base anchors (decode mode)
    RnLevels lv;
    Norm4 nm;
    float clipW, clipH;
    int N;
};

// the candidate's 16-byte row.  Scattered (~1 % of the rows are ever visited): no L1 allocation, smallest L2 fetch granularity
__device__ __forceinline__ float4 fetch_row(const BoxSource& s, int page, int n) {
    float4 d;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(reinterpret_cast<const float4*>(s.rows) + (size_t)page * s.N + n));
    return d;
}

// row -> box.  Decode mode: fp32 anchor as the Anchors layer computes it (model/utils.py:51-80), a + (d * std + mean) * len
// in the reference's order (model/utils.py:102-110), clip to [0, W] x [0, H] (model/layers.py:166-169).
template <bool DECODE>
__device__ __forceinline__ float4 finish_box(const BoxSource& s, int n, const float4 d, const float4* s_base = nullptr) {
    if (!DECODE) return d;
    int level, cx, cy, a;
    rn_locate(s.lv, n, level, cx, cy, a);
    const int bi = level * s.lv.anchors_per_cell + a;
    // (s_base: the base anchors staged in shared memory -- a global read here sits in the middle of a group's dependent chain)
    float4 ba;
    if (s_base) ba = s_base[bi];
    else { const float* bs = s.base32 + (size_t)bi * 4; ba = make_float4(__ldg(bs), __ldg(bs + 1), __ldg(bs + 2), __ldg(bs + 3)); }
    const float sx = ((float)cx + 0.5f) * (float)s.lv.stride[level];
    const float sy = ((float)cy + 0.5f) * (float)s.lv.stride[level];
    const float ax1 = ba.x + sx, ay1 = ba.y + sy, ax2 = ba.z + sx, ay2 = ba.w + sy;
    const float w = ax2 - ax1, h = ay2 - ay1;
    float4 o;
    o.x = ax1 + (d.x * s.nm.std[0] + s.nm.mean[0]) * w;
    o.y = ay1 + (d.y * s.nm.std[1] + s.nm.mean[1]) * h;
    o.z = ax2 + (d.z * s.nm.std[2] + s.nm.mean[2]) * w;
    o.w = ay2 + (d.w * s.nm.std[3] + s.nm.mean[3]) * h;
    o.x = fminf(fmaxf(o.x, 0.0f), s.clipW);
    o.y = fminf(fmaxf(o.y, 0.0f), s.clipH);
    o.z = fminf(fmaxf(o.z, 0.0f), s.clipW);
    o.w = fminf(fmaxf(o.w, 0.0f), s.clipH);
    return o;
}

// ------------------------------------------------------------------------------------------------
// K4 + K5
// ------------------------------------------------------------------------------------------------
struct IouTest { float thr, r; int plain; };    // see iou_exceeds()

struct NmsParams {
    Slabs sl;
    BoxSource src;
    int S;                 // segments
    int segs_per_page;     // C (class specific) or 1
    int nms;
    IouTest iou;           // threshold + the derived constants of iou_exceeds()
    int max_det;
    int pre_nms_top_k;
    unsigned key_floor_hi;         // every key's upper (score) word is >= this (the score threshold's image; 0 = unknown)
    // per-segment results
    int* kept_count;               // (S)
    unsigned long long* kept_key;  // (S, max_det)
    float4* kept_box;              // (S, max_det)
    int* kept_label;               // (S, max_det)
    int* status;                   // (pages) or nullptr
    unsigned long long* timing;    // 32 phase counters / event counts (clock64 ticks of thread 0, summed over CTAs) + 56 segment records, or nullptr
};

// TF non_max_suppression_op.cc IOU on corner-normalised boxes: IoU = 0 when an area is not positive, else
// inter / (area_a + area_b - inter); suppression iff IoU > thr (strict).
//   IoU > thr  <=>  inter > thr (area_a + area_b - inter)  <=>  inter > R (area_a + area_b),  R = thr / (1 + thr),
// so every box carries its WEIGHT w = R * area (+inf when the area is not positive: such a box never suppresses nor is
// suppressed) and the test is inter > w_a + w_b: 4 min/max, 2 subtractions, one clamp, one product, one sum.  The fp32
// quotient of the reference is only evaluated when that cannot decide: the IoU grows at least as fast (relatively) as
// inter does, so inter > (w_a + w_b)(1 + 1e-4) implies RN(inter / union) > thr and inter < (w_a + w_b)(1 - 1e-4) implies
// RN(inter / union) <= thr (the fp32 rounding errors of either side are ~1e-7) -- the outcome is bit-identical to always
// dividing.  Thresholds outside [1e-3, 1e3] (and NaN) take the plain formula (`plain`: the weights are the areas then).
__device__ __forceinline__ float box_area(const float4 c) { return (c.z - c.x) * (c.w - c.y); }
__device__ __forceinline__ float box_weight(const float area, const IouTest t) {
    return t.plain ? area : (area > 0.0f ? t.r * area : __int_as_float(0x7f800000));
}

__device__ __noinline__ bool iou_exceeds_exact(const float4 a, const float4 b, float inter, float thr) {
    const float aa = box_area(a), ab = box_area(b);
    return (aa > 0.0f && ab > 0.0f) && (inter / (aa + ab - inter) > thr);
}

__device__ __forceinline__ bool iou_exceeds(const float4 a, const float wa, const float4 b, const float wb, const IouTest t) {
    if (t.plain) {                          // degenerate threshold (kernel-uniform): the plain formula on the areas
        float iou = 0.0f;
        if (wa > 0.0f && wb > 0.0f) {
            const float iw = fmaxf(fminf(a.z, b.z) - fmaxf(a.x, b.x), 0.0f);
            const float ih = fmaxf(fminf(a.w, b.w) - fmaxf(a.y, b.y), 0.0f);
            const float inter = iw * ih;
            iou = inter / (wa + wb - inter);
        }
        return iou > t.thr;
    }
    const float iw = fmaxf(fminf(a.z, b.z) - fmaxf(a.x, b.x), 0.0f);     // clamping ONE side is enough: inter <= 0 then
    const float ih = fminf(a.w, b.w) - fmaxf(a.y, b.y);
    const float inter = iw * ih;
    const float rhs = wa + wb;                                            // > 0, or +inf
    const bool sure = inter > rhs * 1.0001f;
    const bool maybe = !sure && (inter > rhs * 0.9999f);
    if (maybe) return iou_exceeds_exact(a, b, inter, t.thr);             // ~1e-4 of the overlapping pairs
    return sure;
}

// Descending bitonic sort of n (power of two, 128 <= n <= NMS_CHUNK) (key, slot) pairs in shared memory.
// Each of the first n/2 threads keeps 2 adjacent elements in registers (all 1024 threads are busy at n = 2048, so
// shuffle and shared-memory latencies overlap across 32 warps): compare-exchange distance 1 is thread-local,
// 2..32 are warp shuffles, distances >= 64 go through shared memory.  Keys are unique (zero padding excepted, which
// never swaps).  Must be called by all threads of the CTA.
template <bool SLOT>
__device__ __forceinline__ void ce_keep(unsigned long long& a, unsigned& av, unsigned long long b, unsigned bv, bool take_max) {
    const bool sw = take_max ? (b > a) : (b < a);
    if (sw) { a = b; if (SLOT) av = bv; }
}

// SLOT: the 32-bit payload (slab slot, needed for the labels of class-agnostic filtering) travels with the keys
template <bool SLOT>
__device__ __forceinline__ void sort_chunk_desc(unsigned long long* s_key, unsigned* s_slot, const int n, const int tid) {
    const bool act = tid < (n >> 1);                 // warp-uniform: n/2 is a multiple of 32
    const int i0 = tid * 2;
    unsigned long long k0 = 0ull, k1 = 0ull;
    unsigned v0 = 0u, v1 = 0u;
    if (act) { k0 = s_key[i0]; k1 = s_key[i0 + 1]; if (SLOT) { v0 = s_slot[i0]; v1 = s_slot[i0 + 1]; } }
    for (int k = 2; k <= n; k <<= 1) {
        const bool desc = ((i0 & k) == 0);           // direction of this thread's pair in the k-merge (k >= 2: same for both)
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 64) {
                if (act) { s_key[i0] = k0; s_key[i0 + 1] = k1; if (SLOT) { s_slot[i0] = v0; s_slot[i0 + 1] = v1; } }
                __syncthreads();
                if (act) {
                    const int x0 = i0 ^ j, x1 = (i0 + 1) ^ j;
                    const bool lower = ((i0 & j) == 0);                  // this thread holds the lower index of each pair
                    ce_keep<SLOT>(k0, v0, s_key[x0], SLOT ? s_slot[x0] : 0u, lower == desc);
                    ce_keep<SLOT>(k1, v1, s_key[x1], SLOT ? s_slot[x1] : 0u, lower == desc);
                }
                __syncthreads();
            } else if (j >= 2) {
                if (act) {
                    const int lm = j >> 1;
                    const bool lower = ((i0 & j) == 0);
                    const unsigned long long b0 = __shfl_xor_sync(0xffffffffu, k0, lm), b1 = __shfl_xor_sync(0xffffffffu, k1, lm);
                    unsigned w0 = 0u, w1 = 0u;
                    if (SLOT) { w0 = __shfl_xor_sync(0xffffffffu, v0, lm); w1 = __shfl_xor_sync(0xffffffffu, v1, lm); }
                    ce_keep<SLOT>(k0, v0, b0, w0, lower == desc);
                    ce_keep<SLOT>(k1, v1, b1, w1, lower == desc);
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
    if (act) { s_key[i0] = k0; s_key[i0 + 1] = k1; if (SLOT) { s_slot[i0] = v0; s_slot[i0 + 1] = v1; } }
    __syncthreads();
}

// How many of the 2^SH descending keys at `part` precede x in a merge: those > x, and for an element of a pair's SECOND run
// (second = 1) also those == x (only the zero padding is ever equal).  y >= x is y > x - 1; for x = 0 everything precedes.
// A branch-free binary search, fully unrolled: per step ONE 32-bit read of a key's score word at a compile-time offset, one
// compare and one predicated add.  Equal score words (rare) do not decide: such a search is repeated on the full keys.
__device__ __noinline__ int merge_rank_full(const unsigned long long* part, const int sh, const unsigned long long t) {
    int c = 0;
    for (int s = (1 << sh) >> 1; s >= 1; s >>= 1) if (part[c + s - 1] > t) c += s;
    return c + (part[c] > t ? 1 : 0);
}
template <int SH>
__device__ __forceinline__ int merge_rank(const unsigned long long* part, const unsigned long long x, const int second) {
    if (second && x == 0ull) return 1 << SH;
    const unsigned long long t = x - (unsigned long long)second;
    const unsigned th = (unsigned)(t >> 32);
    const unsigned char* base = reinterpret_cast<const unsigned char*>(part) + 4;      // the keys' upper words
    unsigned cb = 0u;                                       // the count, in bytes (8 per key)
    bool tie = false;
#pragma unroll
    for (int s = (1 << SH) >> 1; s >= 1; s >>= 1) {
        const unsigned y = *reinterpret_cast<const unsigned*>(base + cb + (s - 1) * 8);
        tie |= (y == th);
        if (y > th) cb += (unsigned)s * 8u;
    }
    {
        const unsigned y = *reinterpret_cast<const unsigned*>(base + cb);
        tie |= (y == th);
        if (y > th) cb += 8u;
    }
    if (tie) return merge_rank_full(part, SH, t);
    return (int)(cb >> 3);
}

// Merge sort of the chunk, descending, n2p a power of two in [128, NMS_CHUNK] (padding keys are 0 and end up last).
//  1. every warp orders its 64 elements in registers (2 per lane: a bitonic network of 21 stages, 15 of them shuffles) --
//     no barrier, all 32 warps busy;
//  2. log2(n2p / 64) merge levels: an element's place in the merge of its run with the partner run is its index in its own
//     run plus the number of partner elements that precede it -- a branch-free binary search (log2(run) + 1 shared-memory
//     reads; a thread's two elements search in lockstep), then ONE store into the other buffer and ONE barrier per level.
//     Keys are unique; the zero padding is not, so elements of the second run of a pair also count EQUAL partners.
// ~900 instructions per thread and 6 barriers for 2048 keys.  What it replaced (r2q): an LSD radix sort on the score words
// (4 passes of 8 bits for a chunk whose scores span 0.1 .. 0.99, 7 barriers per pass, plus a repair pass for equal scores and a
// bitonic fall-back) took 25 k cycles per chunk, the bitonic network alone 37 k.
// The result is in s_key (s_slot); which buffer step 1 writes to is chosen so that the last level lands there.
template <bool SLOT>
__device__ __forceinline__ void merge_sort_desc(unsigned long long* s_key, unsigned* s_slot, unsigned long long* s_key2, unsigned* s_slot2,
                                                const int n2p, const int tid) {
    int levels = 0;
    for (int r = 64; r < n2p; r <<= 1) ++levels;
    unsigned long long* ka = (levels & 1) ? s_key2 : s_key;
    unsigned long long* kb = (levels & 1) ? s_key : s_key2;
    unsigned* va = (levels & 1) ? s_slot2 : s_slot;
    unsigned* vb = (levels & 1) ? s_slot : s_slot2;
    {
        const bool act = tid < (n2p >> 1);               // warp-uniform: n2p / 2 is a multiple of 32
        const int i0 = tid * 2;
        unsigned long long k0 = 0ull, k1 = 0ull;
        unsigned v0 = 0u, v1 = 0u;
        if (act) {
            k0 = s_key[i0]; k1 = s_key[i0 + 1];
            if (SLOT) { v0 = s_slot[i0]; v1 = s_slot[i0 + 1]; }
#pragma unroll
            for (int k = 2; k <= 64; k <<= 1) {
                const bool desc = (k == 64) || ((i0 & k) == 0);     // the last merge leaves every 64-run descending
#pragma unroll
                for (int j = k >> 1; j > 0; j >>= 1) {
                    if (j >= 2) {
                        const int lm = j >> 1;
                        const bool lower = ((i0 & j) == 0);
                        const unsigned long long b0 = __shfl_xor_sync(0xffffffffu, k0, lm), b1 = __shfl_xor_sync(0xffffffffu, k1, lm);
                        unsigned w0 = 0u, w1 = 0u;
                        if (SLOT) { w0 = __shfl_xor_sync(0xffffffffu, v0, lm); w1 = __shfl_xor_sync(0xffffffffu, v1, lm); }
                        ce_keep<SLOT>(k0, v0, b0, w0, lower == desc);
                        ce_keep<SLOT>(k1, v1, b1, w1, lower == desc);
                    } else {
                        const bool sw = desc ? (k0 < k1) : (k0 > k1);
                        if (sw) {
                            const unsigned long long tk = k0; k0 = k1; k1 = tk;
                            const unsigned tv = v0; v0 = v1; v1 = tv;
                        }
                    }
                }
            }
        }
        __syncthreads();                                 // (s_key may be the destination: everybody has read it)
        if (act) { ka[i0] = k0; ka[i0 + 1] = k1; if (SLOT) { va[i0] = v0; va[i0 + 1] = v1; } }
        __syncthreads();
    }
    const bool two = n2p > NMS_THREADS, one = tid < n2p;
#pragma unroll 1
    for (int sh = 6; (1 << sh) < n2p; ++sh) {
        const int r = 1 << sh;
        const int e0 = tid, e1 = tid + NMS_THREADS;
        const unsigned long long x0 = one ? ka[e0] : 0ull, x1 = two ? ka[e1] : 0ull;
        int c0 = 0, c1 = 0;
        if (one) {
            const int run0 = e0 >> sh, run1 = e1 >> sh;
            const unsigned long long* p0 = ka + ((run0 ^ 1) << sh);
            const unsigned long long* p1 = ka + ((run1 ^ 1) << sh);
            switch (sh) {                                   // (block-uniform)
                case 6: c0 = merge_rank<6>(p0, x0, run0 & 1); if (two) c1 = merge_rank<6>(p1, x1, run1 & 1); break;
                case 7: c0 = merge_rank<7>(p0, x0, run0 & 1); if (two) c1 = merge_rank<7>(p1, x1, run1 & 1); break;
                case 8: c0 = merge_rank<8>(p0, x0, run0 & 1); if (two) c1 = merge_rank<8>(p1, x1, run1 & 1); break;
                case 9: c0 = merge_rank<9>(p0, x0, run0 & 1); if (two) c1 = merge_rank<9>(p1, x1, run1 & 1); break;
                default: c0 = merge_rank<10>(p0, x0, run0 & 1); if (two) c1 = merge_rank<10>(p1, x1, run1 & 1); break;
            }
            const int at0 = ((run0 >> 1) << (sh + 1)) + (e0 & (r - 1)) + c0;
            kb[at0] = x0;
            if (SLOT) vb[at0] = va[e0];
            if (two) {
                const int at1 = ((run1 >> 1) << (sh + 1)) + (e1 & (r - 1)) + c1;
                kb[at1] = x1;
                if (SLOT) vb[at1] = va[e1];
            }
        }
        __syncthreads();
        unsigned long long* tk = ka; ka = kb; kb = tk;
        unsigned* tv = va; va = vb; vb = tv;
    }
}

// ---- bisection helpers: #{keys >= pivot} for three pivots at once -------------------------------------------------
// keys held in registers: visited keys have been zeroed, so "unvisited" needs no test (every pivot is >= 1)
__device__ __forceinline__ void count3_hi(unsigned khi, unsigned q1, unsigned q2, unsigned q3, unsigned& c1, unsigned& c2, unsigned& c3) {
    c1 += (khi >= q1) ? 1u : 0u;
    c2 += (khi >= q2) ? 1u : 0u;
    c3 += (khi >= q3) ? 1u : 0u;
}
__device__ __forceinline__ void count3(unsigned long long k, unsigned long long q1, unsigned long long q2, unsigned long long q3,
                                       unsigned& c1, unsigned& c2, unsigned& c3) {
    const bool live = k != 0ull;
    c1 += (live && k >= q1) ? 1u : 0u;
    c2 += (live && k >= q2) ? 1u : 0u;
    c3 += (live && k >= q3) ? 1u : 0u;
}

// A selected box as the suppression loops read it: corners + weight in ONE 32-byte record (one address register, two reads)
struct __align__(16) SelRec { float4 box; float w; float pad[3]; };

// shared-memory access by 32-bit shared address.  The base of an array carved out of the dynamic shared memory is a value the
// compiler re-derives wherever it is used (S2UR SR_CgaCtaId + 4 uniform instructions, inside the loops); pinned in a register
// once (smem_pin) and advanced by hand, the loops below carry one address and nothing else.
__device__ __forceinline__ unsigned smem_pin(const void* p) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}
__device__ __forceinline__ float4 lds128(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
template <int IMM>
__device__ __forceinline__ float lds32f(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1 + %2];" : "=f"(v) : "r"(a), "n"(IMM));
    return v;
}

// The alive candidates [w0, w0 + wn) of the current group (wn <= 256, w0 a multiple of 32) are tested against the selected
// boxes [s_lo, s_hi): NMS_THREADS / wp threads per candidate split the list (wp = wn rounded up to a power of two >= 32;
// thread-serial loops over broadcast shared-memory reads: every lane does useful work), a hit clears the candidate's alive
// bit (one shared-memory atomic per warp).  Callers separate this from the next read of the alive bits with a block barrier.
// The loop is 17 instructions per test: two reads, the 13 of the test with ONE compare (inter > 0.9999 (w_a + w_b): a hit or
// a near-hit, rare) and the loop's own three.
// (Tried and dropped, r2o: an 8 x 8 grid of per-cell bit masks of the selected boxes, so that a candidate only tests the boxes
// that share a cell with it -- exact, but the lookups and the insertion of table-sized boxes cost more than the tests they
// saved: the NMS kernel went from 115 to 177 us per 64 pages.)
__device__ __forceinline__ void suppress_window(const float4* s_gbox, const float* s_garea, unsigned* s_alive, int w0, int wn,
                                                const SelRec* s_sel, int s_lo, int s_hi, const IouTest thr,
                                                int tid, int lane) {
    const int sh = max(5, 32 - __clz(wn - 1)), wp = 1 << sh;     // wn in [1, 256]
    const int c = w0 + (tid & (wp - 1)), part = tid >> sh, parts = NMS_THREADS >> sh;
    bool dead = false;
    if (c < w0 + wn && ((s_alive[c >> 5] >> (c & 31)) & 1u)) {
        const float4 cb = s_gbox[c];
        const float ca = s_garea[c];
        if (thr.plain) {
            for (int s = s_lo + part; s < s_hi && !dead; s += parts) dead = iou_exceeds(cb, ca, s_sel[s].box, s_sel[s].w, thr);
        } else {
            const unsigned base = smem_pin(s_sel);
            unsigned at = base + (unsigned)(s_lo + part) * (unsigned)sizeof(SelRec);
            const unsigned end = base + (unsigned)s_hi * (unsigned)sizeof(SelRec), step = (unsigned)parts * (unsigned)sizeof(SelRec);
            while (at < end) {
                const float4 b = lds128(at);
                const float wb = lds32f<16>(at);
                at += step;
                const float iw = fmaxf(fminf(cb.z, b.z) - fmaxf(cb.x, b.x), 0.0f);
                const float ih = fminf(cb.w, b.w) - fmaxf(cb.y, b.y);
                const float inter = iw * ih;
                const float rhs = ca + wb;
                if (inter > rhs * 0.9999f) {                // a hit, or too close to call without the quotient
                    if (inter > rhs * 1.0001f || iou_exceeds_exact(cb, b, inter, thr.thr)) { dead = true; break; }
                }
            }
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, dead);
    if (m != 0u && lane == 0) atomicAnd(&s_alive[c >> 5], ~m);      // a warp's 32 candidates are one word of the bit set
}

// alive word `w` of a group when the window [from, from + wn) opens (from is a multiple of 32)
__device__ __forceinline__ unsigned window_bits(int w, int from, int wn) {
    const int left = from + wn - w * 32, skip = from - w * 32;
    return (skip > 0 || left <= 0) ? 0u : (left >= 32 ? ~0u : (1u << left) - 1u);
}
// The next window of a group of gn candidates of which gdone are open: no wider than what can still be needed (about
// max_det - nsel more selections) and no wider than ~6 k pair tests at opening -- a candidate of an open window is tested
// against every later selection as well, so what is opened but never consumed (behind the stopping point) is pure waste.
// A multiple of 32 unless it ends the group; 0 when the group is open to its end.
__device__ __forceinline__ int next_window(int nsel, int gdone, int gn, int max_det) {
    if (gdone >= gn) return 0;
    const int room = max_det - nsel;
    const int by_cost = max(NMS_BATCH, (int)(__fdividef(6144.0f, (float)max(nsel, 24)) + 0.01f) & ~31);
    return min(gn - gdone, min(by_cost, (room + (room >> 2) + 47) & ~31));
}

template <bool DECODE, bool SLOT>
__global__ void __launch_bounds__(NMS_THREADS, 1) k_segment_nms(const NmsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // dynamic: selected boxes (normalised corners) + areas, sized by max_det
    SelRec* s_sel = reinterpret_cast<SelRec*>(smem_raw);
    unsigned long long* s_key2 = reinterpret_cast<unsigned long long*>(s_sel + p.max_det);      // merge sort: second key buffer,
    float4* s_raw = reinterpret_cast<float4*>(s_key2 + NMS_CHUNK);                               // the chunk's boxes as decoded / stored, in order
    unsigned* s_slot2 = reinterpret_cast<unsigned*>(s_raw + NMS_CHUNK);                          // second payload buffer (SLOT)

    __shared__ unsigned long long s_key[NMS_CHUNK];
    __shared__ unsigned s_slot[SLOT ? NMS_CHUNK : 1];
    __shared__ float4 s_gbox[NMS_GROUP];      // the group's boxes, corner-normalised
    __shared__ float s_garea[NMS_GROUP];
    __shared__ unsigned s_alive[NMS_GROUP / 32];   // bit set: candidate of the group neither consumed nor suppressed yet
    __shared__ uint2 s_conf[NMS_BATCH];       // per batch member: the members it conflicts with (IoU > thr), 64 bits
    __shared__ unsigned char s_bpos[NMS_BATCH];   // per batch member: its position in the group
    __shared__ int s_alive_total, s_open;
    __shared__ float4 s_banchor[NMS_BASE_MAX];   // decode mode: the base anchors (levels x anchors per cell), when they fit
    __shared__ unsigned s_cnt[3][4];          // pivot counts of the bisection, three rotating sets
    __shared__ unsigned s_kmax;
    __shared__ int s_loaded, s_nsel;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int seg = blockIdx.x;
    const int total = p.sl.counts[seg];
    const int cnt = (int)min((long long)total, p.sl.cap);
    const int page = seg / p.segs_per_page;
    if (total > cnt && p.status && tid == 0) p.status[page] = 1;
    const int limit = p.pre_nms_top_k > 0 ? min(cnt, p.pre_nms_top_k) : cnt;
    const unsigned long long* keys = p.sl.keys + (size_t)seg * p.sl.cap;
    const int* labels = p.sl.labels ? p.sl.labels + (size_t)seg * p.sl.cap : nullptr;
    const int seg_label = seg - page * p.segs_per_page;

    // optional phase timing (rn_debug_nms_timing(1)): thread 0 accumulates clock64() deltas per phase
    long long t_mark = p.timing ? clock64() : 0;
    const long long t_start = t_mark;
    __shared__ unsigned s_dbg[16];      // (phase timing only: this segment's event counts)
    if (p.timing && tid < 16) s_dbg[tid] = 0u;
#define RN_COUNT(k, v) do { if (p.timing && tid == 0) { atomicAdd(p.timing + (k), (unsigned long long)(v)); s_dbg[(k) - 10] += (unsigned)(v); } } while (0)
#define RN_PHASE(k) do { if (p.timing && tid == 0) { const long long now = clock64(); atomicAdd(p.timing + (k), (unsigned long long)(now - t_mark)); t_mark = now; } } while (0)
    unsigned long long upper = ~0ull;   // keys >= upper have been visited (no key equals ~0: its score bits would be a NaN's)
    int visited = 0, nsel = 0, round = 0, bis = 0;
    if (tid == 0) { s_nsel = 0; s_kmax = 0u; s_loaded = 0; }
    if (tid < 4) s_cnt[0][tid] = 0u;
    const int n_base = DECODE ? p.src.lv.num_levels * p.src.lv.anchors_per_cell : 0;
    const float4* base_tab = (DECODE && n_base <= NMS_BASE_MAX) ? s_banchor : nullptr;
    if (base_tab && tid < n_base) {
        const float* bs = p.src.base32 + (size_t)tid * 4;
        s_banchor[tid] = make_float4(__ldg(bs), __ldg(bs + 1), __ldg(bs + 2), __ldg(bs + 3));
    }
    __syncthreads();
    // The slab's keys are read ONCE into registers (8 per thread) when the slab has <= 8192 candidates -- the
    // bisection passes and the gathers of every round then run out of registers (visited keys are zeroed there);
    // larger slabs stream the keys from global memory (L2) in every pass.
    constexpr int KPT = 8;
    const bool in_regs = cnt <= KPT * NMS_THREADS;
    unsigned long long rk[KPT];
    {
        unsigned mx = 0u;
#pragma unroll
        for (int t = 0; t < KPT; ++t) {
            const int i = t * NMS_THREADS + tid;
            rk[t] = (in_regs && i < cnt) ? __ldcg(keys + i) : 0ull;      // 0 never matches (real keys are > 0)
            mx = max(mx, (unsigned)(rk[t] >> 32));
        }
        mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0 && mx) atomicMax(&s_kmax, mx);
    }
    __syncthreads();
    // every key's score word lies in [key_floor_hi, kmax]: the first bracket of the bisection
    const unsigned long long top = in_regs ? (((unsigned long long)s_kmax + 1ull) << 32) : 0ull;   // 0: no bound known (also on overflow)

    // ---------------- K4: a threshold key that keeps between `lo` and `hi` of the unvisited keys (more than `hi` are left) ------
    // c(x) = #{unvisited k >= x} falls from the number left (> hi) at x = L to 0 (< lo) at x = H; keys are unique, so
    // c steps by one and some x has lo <= c(x) <= hi.  Each pass evaluates c at the three quartile points of [L, H)
    // and either accepts one of them or keeps the quarter that brackets the window.  While the bracket spans at
    // least four values of the keys' upper (score) word the pivots are multiples of 2^32, so a 32-bit compare of
    // that word decides k >= pivot; only ties in the score ever need the 64-bit form.  Called by all threads.
    auto find_threshold = [&](const int lo, const int hi) -> unsigned long long {
        unsigned long long thr_key = 0ull;
        unsigned long long L = (unsigned long long)p.key_floor_hi << 32, H = upper;
        if (top != 0ull && top < H) H = top;
        for (int pass = 0; pass < 96; ++pass, ++bis) {
            const int set = bis % 3;                          // `bis` runs on across calls: the rotation never restarts on a used set
            if (tid < 4) s_cnt[(bis + 1) % 3][tid] = 0u;      // the next pass's set (last read two barriers ago)
            const unsigned long long span = H - L;
            const bool wide = in_regs && (span >> 34) != 0ull;
            unsigned long long q1, q2, q3;
            unsigned c1 = 0u, c2 = 0u, c3 = 0u;
            if (wide) {
                const unsigned long long Lh = L >> 32, sh = span >> 32;
                const unsigned long long h1 = Lh + (sh >> 2), h2 = Lh + (sh >> 1), h3 = h2 + (sh >> 2);
                q1 = h1 << 32; q2 = h2 << 32; q3 = h3 << 32;
#pragma unroll
                for (int t = 0; t < KPT; ++t) count3_hi((unsigned)(rk[t] >> 32), (unsigned)h1, (unsigned)h2, (unsigned)h3, c1, c2, c3);
            } else {
                q1 = L + (span >> 2); q2 = L + (span >> 1); q3 = q2 + (span >> 2);
                if (in_regs) {
#pragma unroll
                    for (int t = 0; t < KPT; ++t) count3(rk[t], q1, q2, q3, c1, c2, c3);
                } else {
                    for (int i = tid; i < cnt; i += NMS_THREADS) {
                        const unsigned long long k = __ldcg(keys + i);
                        if (k < upper) count3(k, q1, q2, q3, c1, c2, c3);
                    }
                }
            }
            c1 = __reduce_add_sync(0xffffffffu, c1);
            c2 = __reduce_add_sync(0xffffffffu, c2);
            c3 = __reduce_add_sync(0xffffffffu, c3);
            if (lane == 0) {
                if (c1) atomicAdd(&s_cnt[set][0], c1);
                if (c2) atomicAdd(&s_cnt[set][1], c2);
                if (c3) atomicAdd(&s_cnt[set][2], c3);
            }
            __syncthreads();
            const int n1 = (int)s_cnt[set][0], n2 = (int)s_cnt[set][1], n3 = (int)s_cnt[set][2];   // n1 >= n2 >= n3
            if (n1 >= lo && n1 <= hi) { thr_key = q1; ++bis; break; }
            if (n2 >= lo && n2 <= hi) { thr_key = q2; ++bis; break; }
            if (n3 >= lo && n3 <= hi) { thr_key = q3; ++bis; break; }
            if (n1 < lo) H = q1;                         // c(L) > hi, c(q1) < lo
            else if (n2 < lo) { L = q1; H = q2; }        // n1 > hi
            else if (n3 < lo) { L = q2; H = q3; }
            else L = q3;                                 // n3 > hi
            thr_key = L;                                 // (only used if the pass limit is ever hit: more than `hi` keys, the gather truncates)
        }
        return thr_key;
    };

    while (visited < limit && nsel < p.max_det) {
        // ---------------- a round after the first: sweep before sorting ------------------------------------------------
        // The first chunk did not yield max_det selections.  When that is because most of it was suppressed by a FEW boxes
        // (clusters of high-scoring anchors around each table), what follows in score order is mostly more of the same:
        // the next <= 2048 candidates are tested against the selected boxes right where they sit -- in the registers, no
        // sort, no order needed: suppression by an already selected box does not depend on the order -- and the suppressed
        // ones are struck out (they count as visited), so that the rounds below sort only what can still be selected.
        if (round > 0 && p.nms && in_regs && limit == cnt && nsel > 0 && nsel * 4 < p.max_det && cnt - visited > 512) {
            unsigned long long thr_s = 0ull;
            if (cnt - visited > NMS_CHUNK) thr_s = find_threshold(NMS_CHUNK - (NMS_CHUNK >> 2), NMS_CHUNK);
#pragma unroll
            for (int t = 0; t < KPT; ++t)                   // the rows are requested first, read (from L2) below
                if (rk[t] != 0ull && rk[t] >= thr_s)
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const float4*>(p.src.rows) + (size_t)page * p.src.N + key_idx(rk[t])));
            int killed = 0;
#pragma unroll
            for (int t = 0; t < KPT; ++t) {
                if (rk[t] != 0ull && rk[t] >= thr_s) {
                    const int n = (int)key_idx(rk[t]);
                    const float4 r = finish_box<DECODE>(p.src, n, fetch_row(p.src, page, n), base_tab);
                    float4 c;
                    c.x = fminf(r.x, r.z); c.y = fminf(r.y, r.w); c.z = fmaxf(r.x, r.z); c.w = fmaxf(r.y, r.w);
                    const float ca = box_weight(box_area(c), p.iou);
                    bool dead = false;
                    for (int sI = 0; sI < nsel && !dead; ++sI) dead = iou_exceeds(c, ca, s_sel[sI].box, s_sel[sI].w, p.iou);
                    if (dead) { rk[t] = 0ull; ++killed; }
                }
            }
            killed = __reduce_add_sync(0xffffffffu, killed);
            if (lane == 0 && killed) atomicAdd(&s_loaded, killed);      // s_loaded is free between rounds (zero here)
            __syncthreads();
            visited += s_loaded;
            __syncthreads();
            if (tid == 0) s_loaded = 0;
            RN_PHASE(6);
            if (visited >= limit) break;
        }
        // The first round takes (up to) a full chunk; later rounds are only reached when its candidates did not yield
        // max_det selections -- typically a few hundred more are needed -- so they start small and double.
        // (with pre_nms_top_k only `need` more candidates may be visited at all: the chunk is sized for them)
        const int need = limit - visited;
        int hi = round == 0 ? NMS_CHUNK : min(NMS_CHUNK, 512 << (round - 1));
        while (hi >= 256 && (hi >> 1) >= need) hi >>= 1;
        const int lo = min(hi - (hi >> 2), need);
        ++round;
        const int remaining = cnt - visited;
        const unsigned long long thr_key = remaining > hi ? find_threshold(lo, hi) : 0ull;     // 0: everything that is left
        RN_PHASE(0);
        // ---------------- gather the chunk into shared memory (one shared-memory atomic per warp and key slot) ----------
        if (tid == 0) s_loaded = 0;
        __syncthreads();
        if (in_regs) {
#pragma unroll
            for (int t = 0; t < KPT; ++t) {
                const unsigned long long k = rk[t];
                const bool in = (k != 0ull) && (k >= thr_key);
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (m) {                                    // warp-uniform
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_loaded, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    const int at = base + __popc(m & ((1u << lane) - 1u));
                    if (in && at < NMS_CHUNK) {
                        s_key[at] = k;
                        if (SLOT) s_slot[at] = (unsigned)(t * NMS_THREADS + tid);
                        // the chunk's rows start their way from DRAM to L2 now: the sort hides the latency of the first group's reads
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const float4*>(p.src.rows) + (size_t)page * p.src.N + key_idx(k)));
                    }
                }
            }
        } else {
            for (int i0 = warp * 32; i0 < cnt; i0 += NMS_THREADS) {      // warp-uniform trip count
                const int i = i0 + lane;
                const unsigned long long k = i < cnt ? __ldcg(keys + i) : 0ull;
                const bool in = (k != 0ull) && (k < upper) && (k >= thr_key);
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (m) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_loaded, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    const int at = base + __popc(m & ((1u << lane) - 1u));
                    if (in && at < NMS_CHUNK) { s_key[at] = k; if (SLOT) s_slot[at] = (unsigned)i; }
                }
            }
        }
        __syncthreads();
        const int loaded = min(s_loaded, NMS_CHUNK);
        if (s_loaded > NMS_CHUNK && p.status && tid == 0) p.status[page] = 2;   // cannot happen (see the bisection); never silent
        int n2p = 128;
        while (n2p < loaded) n2p <<= 1;
        for (int i = loaded + tid; i < n2p; i += NMS_THREADS) { s_key[i] = 0ull; if (SLOT) s_slot[i] = 0u; }
        __syncthreads();
        RN_PHASE(1);
        // ---------------- order the chunk, descending (keys are unique) ----------------------------
        merge_sort_desc<SLOT>(s_key, s_slot, s_key2, s_slot2, n2p, tid);
        const int chunk_n = min(loaded, limit - visited);
        RN_PHASE(2);
        RN_COUNT(12, loaded);
        RN_COUNT(14, 1);
        // ---------------- K5: greedy NMS over the ordered chunk -------------------------------------------
        // The chunk's boxes: every candidate's row is read and decoded NOW, by all threads at once (two candidates per thread:
        // one memory latency and one decode chain per round; the rows were requested from DRAM at gather time).  Decoding group by
        // group, 256 threads at a time, put that chain -- ~1.6 k cycles -- in front of every group (r2v: 8 % of the kernel).
        for (int e = tid; e < chunk_n; e += NMS_THREADS) {
            const int n = (int)key_idx(s_key[e]);
            s_raw[e] = finish_box<DECODE>(p.src, n, fetch_row(p.src, page, n), base_tab);
        }
        __syncthreads();
        RN_PHASE(18);
        for (int g0 = 0; g0 < chunk_n && nsel < p.max_det; g0 += NMS_GROUP) {
            const int gn = min(NMS_GROUP, chunk_n - g0);
            if (tid < gn) {
                const float4 r = s_raw[g0 + tid];
                float4 c;
                c.x = fminf(r.x, r.z); c.y = fminf(r.y, r.w); c.z = fmaxf(r.x, r.z); c.w = fmaxf(r.y, r.w);
                s_gbox[tid] = c;
                s_garea[tid] = box_weight(box_area(c), p.iou);
            }
            // The group is opened window by window: a window's candidates are first tested against everything selected so
            // far (what an earlier selection suppresses leaves before its order is ever looked at), then consumed 64 alive
            // candidates at a time.  A window is no wider than what can still be needed -- about (max_det - nsel) more
            // selections -- so candidates behind the stopping point are (almost) never tested.  One greedy step is four
            // phases between block barriers:
            //   apply    all threads: the step's new selections against the open windows' alive candidates, and -- when the
            //            open windows are about to run dry -- the next window against everything selected
            //   rank     the first 64 alive candidates, in order (the batch)
            //   pairs    the 64 x 64 conflict matrix of the batch (warp w: rows w and w + 32, one ballot per half row)
            //   resolve  warp 0: greedy order by fixpoint iteration in registers, the survivors join the selected list, the batch
            //            is consumed, and the next window (if any) is declared
            int gdone = 0;                                  // candidates of the group whose window has been opened
            int n_old = nsel;                               // selections [n_old, nsel) have not met the open windows yet
            int open_wn = next_window(nsel, 0, gn, p.max_det);      // (uniform) the window the next apply phase opens
            if (tid < NMS_GROUP / 32) s_alive[tid] = window_bits(tid, 0, open_wn);
            __syncthreads();
            RN_PHASE(3);
            while (true) {
                // ---- apply
                if (p.nms) {
                    if (nsel > n_old && gdone > 0)
                        suppress_window(s_gbox, s_garea, s_alive, 0, gdone, s_sel, n_old, nsel, p.iou, tid, lane);
                    if (open_wn > 0 && nsel > 0)
                        suppress_window(s_gbox, s_garea, s_alive, gdone, open_wn, s_sel, 0, nsel, p.iou, tid, lane);
                }
                if (open_wn > 0) { RN_COUNT(11, 1); RN_COUNT(16, open_wn); }
                gdone += open_wn;
                __syncthreads();                            // the alive bits are final
                RN_PHASE(9);
                // ---- rank: one thread per candidate of the group ranks itself among the alive ones (its warp's alive word +
                //      the population counts of the words before it); the first 64 leave their positions in shared memory
                if (tid < NMS_GROUP) {
                    const unsigned word = s_alive[warp];
                    int before = 0;
#pragma unroll
                    for (int w8 = 0; w8 < NMS_GROUP / 32 - 1; ++w8) before += (w8 < warp) ? __popc(s_alive[w8]) : 0;
                    if ((word >> lane) & 1u) {
                        const int rank = before + __popc(word & ((1u << lane) - 1u));
                        if (rank < NMS_BATCH) s_bpos[rank] = (unsigned char)tid;
                    }
                    if (tid == NMS_GROUP - 1) s_alive_total = before + __popc(word);
                }
                __syncthreads();
                RN_PHASE(8);
                const int alive_total = s_alive_total;
                if (alive_total == 0 && gdone >= gn) break; // block-uniform: the group is used up
                const int bn = min(NMS_BATCH, alive_total);
                // ---- pairs: row i of the conflict matrix = the members j != i with IoU(i, j) > thr
                if (bn > 0) {
                    RN_COUNT(10, 1);
                    if (p.nms) {
                        const int pos_lo = lane < bn ? (int)s_bpos[lane] : 0, pos_hi = lane + 32 < bn ? (int)s_bpos[lane + 32] : 0;
                        const float4 b_lo = s_gbox[pos_lo], b_hi = s_gbox[pos_hi];
                        const float w_lo = s_garea[pos_lo], w_hi = s_garea[pos_hi];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int i = warp + 32 * h;
                            if (i < bn) {                   // warp-uniform
                                const int mine = (int)s_bpos[i];
                                const float4 mb = s_gbox[mine];
                                const float mw = s_garea[mine];
                                const bool hit_lo = (lane != i) && (lane < bn) && iou_exceeds(mb, mw, b_lo, w_lo, p.iou);
                                const bool hit_hi = (lane + 32 != i) && (lane + 32 < bn) && iou_exceeds(mb, mw, b_hi, w_hi, p.iou);
                                const unsigned c_lo = __ballot_sync(0xffffffffu, hit_lo), c_hi = __ballot_sync(0xffffffffu, hit_hi);
                                if (lane == 0) s_conf[i] = make_uint2(c_lo, c_hi);
                            }
                        }
                    }
                    __syncthreads();
                }
                RN_PHASE(4);
                // ---- resolve (warp 0).  Lane l owns members l and l + 32.  A member is KEPT once every earlier member that
                //      conflicts with it is known to be suppressed, SUPPRESSED once one of them is known to be kept; each round
                //      settles at least the first unsettled member, usually almost all of them (2-3 rounds).
                n_old = nsel;
                if (warp == 0) {
                    unsigned k_lo = 0u, k_hi = 0u;          // the kept members (uniform)
                    if (bn > 0) {
                        const unsigned v_lo = bn >= 32 ? ~0u : (1u << bn) - 1u;
                        const unsigned v_hi = bn >= 64 ? ~0u : (bn > 32 ? (1u << (bn - 32)) - 1u : 0u);
                        k_lo = v_lo; k_hi = v_hi;
                        if (p.nms) {
                            const uint2 ca = (lane < bn) ? s_conf[lane] : make_uint2(0u, 0u);
                            const uint2 cb = (lane + 32 < bn) ? s_conf[lane + 32] : make_uint2(0u, 0u);
                            const unsigned below = (1u << lane) - 1u;
                            // earlier members in conflict: member l has only low ones, member l + 32 all low ones and some high
                            const unsigned ha = ca.x & below, hb_lo = cb.x, hb_hi = cb.y & below;
                            if (__any_sync(0xffffffffu, (ha | hb_lo | hb_hi) != 0u)) {
                                unsigned s_lo = 0u, s_hi = 0u;      // the suppressed members (uniform)
                                k_lo = 0u; k_hi = 0u;
                                while (((k_lo | s_lo) != v_lo) || ((k_hi | s_hi) != v_hi)) {       // warp-uniform
                                    const bool a_open = (lane < bn) && !(((k_lo | s_lo) >> lane) & 1u);
                                    const bool b_open = (lane + 32 < bn) && !(((k_hi | s_hi) >> lane) & 1u);
                                    const bool a_keep = a_open && (ha & ~s_lo) == 0u;
                                    const bool a_supp = a_open && (ha & k_lo) != 0u;
                                    const bool b_keep = b_open && (hb_lo & ~s_lo) == 0u && (hb_hi & ~s_hi) == 0u;
                                    const bool b_supp = b_open && ((hb_lo & k_lo) != 0u || (hb_hi & k_hi) != 0u);
                                    const unsigned nk_lo = __ballot_sync(0xffffffffu, a_keep), ns_lo = __ballot_sync(0xffffffffu, a_supp);
                                    const unsigned nk_hi = __ballot_sync(0xffffffffu, b_keep), ns_hi = __ballot_sync(0xffffffffu, b_supp);
                                    k_lo |= nk_lo; s_lo |= ns_lo; k_hi |= nk_hi; s_hi |= ns_hi;
                                }
                            }
                        }
                        int extra = __popc(k_lo) + __popc(k_hi) - (p.max_det - nsel);    // TF stops at max_output_size: the first ones in order
                        while (extra-- > 0) { if (k_hi) k_hi &= ~(0x80000000u >> __clz(k_hi)); else k_lo &= ~(0x80000000u >> __clz(k_lo)); }
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const unsigned km = h ? k_hi : k_lo;
                            const int m = lane + 32 * h;
                            if (m < bn) {
                                const int pos = (int)s_bpos[m];
                                if ((km >> lane) & 1u) {
                                    const int at_sel = nsel + (h ? __popc(k_lo) : 0) + __popc(km & ((1u << lane) - 1u));
                                    s_sel[at_sel].box = s_gbox[pos];
                                    s_sel[at_sel].w = s_garea[pos];
                                    const size_t at = (size_t)seg * p.max_det + at_sel;
                                    p.kept_key[at] = s_key[g0 + pos];
                                    p.kept_box[at] = s_raw[g0 + pos];
                                    p.kept_label[at] = (SLOT && labels) ? labels[s_slot[g0 + pos]] : seg_label;
                                }
                                atomicAnd(&s_alive[pos >> 5], ~(1u << (pos & 31)));      // the whole batch is consumed
                            }
                        }
                    }
                    const int nsel_new = nsel + __popc(k_lo) + __popc(k_hi);
                    // the next window: when fewer than a batch of opened candidates can be left (the apply phase may strike more)
                    const int wn = (alive_total - bn < NMS_BATCH && nsel_new < p.max_det) ? next_window(nsel_new, gdone, gn, p.max_det) : 0;
                    __syncwarp();
                    if (lane < NMS_GROUP / 32 && wn > 0) s_alive[lane] |= window_bits(lane, gdone, wn);
                    if (lane == 0) { s_nsel = nsel_new; s_open = wn; }
                }
                __syncthreads();
                RN_PHASE(5);
                nsel = s_nsel;
                open_wn = s_open;
                if (nsel >= p.max_det) break;
            }
            __syncthreads();                                // the group's arrays are rewritten next
            RN_PHASE(19);
        }
        visited += chunk_n;
        upper = (chunk_n > 0) ? s_key[chunk_n - 1] : 0ull;
        if (in_regs) {
#pragma unroll
            for (int t = 0; t < KPT; ++t) if (rk[t] >= upper) rk[t] = 0ull;     // visited
        }
        if (tid == 0) s_loaded = 0;                         // (the sweep of the next round counts into it)
        __syncthreads();
        if (chunk_n == 0) break;                            // defensive: no progress is impossible while visited < limit
    }
    if (tid == 0) p.kept_count[seg] = nsel;
    RN_COUNT(13, nsel);
    RN_COUNT(15, cnt);
    if (p.timing && tid == 0) {
        const unsigned long long ticks = (unsigned long long)(clock64() - t_start);
        atomicMax(p.timing + 7, ticks);                     // the slowest CTA
        if (seg < 56) {                                     // the first segments' own records: ticks, counts 10..16
            unsigned long long* rec = p.timing + 32 + 4 * seg;
            rec[0] = ticks;
            rec[1] = (unsigned long long)s_dbg[5] | ((unsigned long long)s_dbg[2] << 32);     // above threshold | sorted
            rec[2] = (unsigned long long)s_dbg[6] | ((unsigned long long)s_dbg[4] << 32);     // opened | rounds
            rec[3] = (unsigned long long)s_dbg[0] | ((unsigned long long)s_dbg[1] << 32);     // batches | windows
        }
    }
#undef RN_PHASE
#undef RN_COUNT
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

// C > 1, the parallel form of the merge: tf.nn.top_k over the class-major concatenation of the per-class kept lists is the
// top max_det of the composite keys  score word | ~class | ~position  (ties in the score -> lower class, then earlier
// position: the earlier place in the concatenation).  Every class list is already in descending composite order, so
// #{keys >= pivot} is a sum of per-class binary searches: a 4-way bisection on the composite key (thread c searches list c for
// three pivots per pass; the lists sit in shared memory) finds the key of the max_det-th entry exactly, the winners are
// gathered and ordered by one bitonic network.  ~10 us per page instead of max_det serial block-wide argmax rounds (237 us
// at C = 80).  One CTA of 1024 threads per page; needs C * max_det * 8 bytes of shared memory (192 KB at C = 80, M = 300).
constexpr int MERGE_THREADS = 1024;
constexpr size_t MERGE_SMEM_MAX = 226 * 1024;             // of the 227 KB a CTA may have on sm_100 (80 classes x 300 need 203.5 KB)

__device__ __forceinline__ int merge_count_ge(const unsigned long long* list, int n, unsigned long long pivot) {
    int lo = 0, hi = n;                                     // entries [0, lo) are >= pivot, [hi, n) are < pivot
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (list[mid] >= pivot) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(MERGE_THREADS, 1) k_merge_topk_select(const MergeParams p) {
    extern __shared__ __align__(16) unsigned long long s_comp[];        // [C][M] composite keys, then the 2048-entry sort buffer
    __shared__ unsigned s_cnt3[3][4];
    __shared__ int s_total;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int C = p.segs_per_page, M = p.max_det;
    unsigned long long* s_sort = s_comp + (size_t)C * M;
    if (tid == 0) s_total = 0;
    if (tid < 4) s_cnt3[0][tid] = 0u;
    __syncthreads();
    int mine = 0;
    for (int e = tid; e < C * M; e += MERGE_THREADS) {
        const int c = e / M, j = e - c * M;
        unsigned long long comp = 0ull;
        if (j < min(p.kept_count[b * C + c], M)) {
            const unsigned long long k = p.kept_key[((size_t)b * C + c) * M + j];
            comp = (k & 0xffffffff00000000ull) | ((unsigned long long)(0xffffu - (unsigned)c) << 16) | (unsigned long long)(0xffffu - (unsigned)j);
            ++mine;
        }
        s_comp[e] = comp;
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if (lane == 0 && mine) atomicAdd(&s_total, mine);
    __syncthreads();
    const int total = s_total, want = min(M, total);
    __syncthreads();                                        // s_total is reused below
    // ---- the composite key of the want-th entry (everything when total <= M) ----
    unsigned long long thr = 1ull;
    if (total > want) {
        unsigned long long L = 0ull, H = ~0ull;             // c(L) = total > want, c(H) = 0 < want; composite keys are unique
        for (int pass = 0; pass < 80; ++pass) {
            const int set = pass % 3;
            if (tid < 4) s_cnt3[(pass + 1) % 3][tid] = 0u;
            const unsigned long long span = H - L;
            const unsigned long long q1 = L + (span >> 2), q2 = L + (span >> 1), q3 = q2 + (span >> 2);
            unsigned c1 = 0u, c2 = 0u, c3 = 0u;
            for (int c = tid; c < C; c += MERGE_THREADS) {
                const int n = min(p.kept_count[b * C + c], M);
                const unsigned long long* list = s_comp + (size_t)c * M;
                c1 += merge_count_ge(list, n, q1); c2 += merge_count_ge(list, n, q2); c3 += merge_count_ge(list, n, q3);
            }
            c1 = __reduce_add_sync(0xffffffffu, c1); c2 = __reduce_add_sync(0xffffffffu, c2); c3 = __reduce_add_sync(0xffffffffu, c3);
            if (lane == 0) {
                if (c1) atomicAdd(&s_cnt3[set][0], c1);
                if (c2) atomicAdd(&s_cnt3[set][1], c2);
                if (c3) atomicAdd(&s_cnt3[set][2], c3);
            }
            __syncthreads();
            const int n1 = (int)s_cnt3[set][0], n2 = (int)s_cnt3[set][1], n3 = (int)s_cnt3[set][2];    // n1 >= n2 >= n3
            if (n1 == want) { thr = q1; break; }
            if (n2 == want) { thr = q2; break; }
            if (n3 == want) { thr = q3; break; }
            if (n1 < want) H = q1;
            else if (n2 < want) { L = q1; H = q2; }
            else if (n3 < want) { L = q2; H = q3; }
            else L = q3;
            thr = L;
        }
    }
    // ---- gather the winners (a prefix of every class list) and order them ----
    int n2p = 128;
    while (n2p < want) n2p <<= 1;
    for (int i = tid; i < n2p; i += MERGE_THREADS) s_sort[i] = 0ull;
    if (tid == 0) s_total = 0;
    __syncthreads();
    for (int c = tid; c < C; c += MERGE_THREADS) {
        const int n = min(p.kept_count[b * C + c], M);
        const int take = merge_count_ge(s_comp + (size_t)c * M, n, thr);
        if (take) {
            const int at = atomicAdd(&s_total, take);       // any order: the network sorts
            for (int j = 0; j < take && at + j < NMS_CHUNK; ++j) s_sort[at + j] = s_comp[(size_t)c * M + j];
        }
    }
    __syncthreads();
    sort_chunk_desc<false>(s_sort, nullptr, n2p, tid);
    const int produced = min(want, s_total);
    for (int m = tid; m < M; m += MERGE_THREADS) {
        const size_t at = (size_t)b * M + m;
        if (m < produced) {
            const unsigned long long comp = s_sort[m];
            const int c = (int)(0xffffu - (unsigned)((comp >> 16) & 0xffffu)), j = (int)(0xffffu - (unsigned)(comp & 0xffffu));
            const size_t src = ((size_t)b * C + c) * M + j;
            const unsigned long long k = p.kept_key[src];
            reinterpret_cast<float4*>(p.out_boxes)[at] = p.kept_box[src];
            p.out_scores[at] = key_score(k);
            p.out_labels[at] = p.kept_label[src];
            if (p.out_indices) p.out_indices[at] = (int)key_idx(k);
        } else {
            reinterpret_cast<float4*>(p.out_boxes)[at] = make_float4(-1.f, -1.f, -1.f, -1.f);
            p.out_scores[at] = -1.0f;
            p.out_labels[at] = -1;
            if (p.out_indices) p.out_indices[at] = -1;
        }
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
    unsigned long long* keys; int* labels;
    unsigned long long* kept_key; float4* kept_box; int* kept_label;
    size_t bytes;
};

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

FilterWs carve(void* ws, int B, int S, long long cap, int max_det, bool agnostic) {
    FilterWs w;
    size_t off = 0;
    char* base = reinterpret_cast<char*>(ws);
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align256(bytes); return p; };
    w.timing = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * 256));  // first 2 KB
    w.counts = reinterpret_cast<int*>(take(sizeof(int) * S));
    w.kept_count = reinterpret_cast<int*>(take(sizeof(int) * S));
    w.status = reinterpret_cast<int*>(take(sizeof(int) * B));
    w.keys = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * (size_t)S * cap));
    w.labels = agnostic ? reinterpret_cast<int*>(take(sizeof(int) * (size_t)S * cap)) : nullptr;
    w.kept_key = reinterpret_cast<unsigned long long*>(take(sizeof(unsigned long long) * (size_t)S * max_det));
    w.kept_box = reinterpret_cast<float4*>(take(sizeof(float4) * (size_t)S * max_det));
    w.kept_label = reinterpret_cast<int*>(take(sizeof(int) * (size_t)S * max_det));
    w.bytes = off;
    return w;
}

// host image of f2ord(): scores > thr have upper key words >= this
unsigned host_f2ord(float f) {
    unsigned u;
    memcpy(&u, &f, sizeof(u));
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

size_t nms_dynamic_smem(int max_det) {      // selected boxes + weights, the merge sort's second buffers
    return (size_t)max_det * sizeof(SelRec) + (size_t)NMS_CHUNK * (sizeof(unsigned long long) + sizeof(float4) + sizeof(unsigned));
}

std::atomic<int> g_phase_timing{0};
// measurement hook (rn_debug_filter_stages): which of the three stages a filter call launches -- bit 0: the workspace reset +
// k_threshold_keys, bit 1: k_segment_nms, bit 2: k_merge_topk.  A stage run alone works on what an earlier full call left in
// the workspace (the NMS kernel only reads the slabs, the merge only the kept lists), so every kernel can be timed as a train
// of back-to-back launches without events in between.
std::atomic<int> g_stages{7};

// static (~34 KB) + dynamic shared memory can exceed the 48 KB default: opt in once per DEVICE (bit per device ordinal)
int nms_opt_in_shared_memory() {
    static std::atomic<unsigned long long> done{0ull};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (bit && (done.load(std::memory_order_acquire) & bit)) return RN_OK;
    const int big = (int)nms_dynamic_smem(MAX_DET_LIMIT);
    cudaError_t ae = cudaFuncSetAttribute(k_segment_nms<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_segment_nms<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_segment_nms<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_segment_nms<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (ae != cudaSuccess) return rn_fail(RN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(ae));
    if (bit) done.fetch_or(bit, std::memory_order_release);
    return RN_OK;
}

int merge_opt_in_shared_memory() {
    static std::atomic<unsigned long long> done{0ull};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (bit && (done.load(std::memory_order_acquire) & bit)) return RN_OK;
    e = cudaFuncSetAttribute(k_merge_topk_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MERGE_SMEM_MAX);
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (bit) done.fetch_or(bit, std::memory_order_release);
    return RN_OK;
}

int run_back_end(const FilterWs& w, const BoxSource& src, bool decode, int B, int S, int segs_per_page, long long cap,
                 int nms, float nms_thr, int max_det, int pre_nms_top_k, unsigned key_floor_hi,
                 float* out_boxes, float* out_scores, int* out_labels, int* out_indices, int* out_count,
                 int* status, cudaStream_t s) {
    NmsParams np;
    np.sl.counts = w.counts; np.sl.keys = w.keys; np.sl.labels = w.labels; np.sl.cap = cap;
    np.src = src;
    np.S = S; np.segs_per_page = segs_per_page; np.nms = nms; np.iou.thr = nms_thr; np.iou.plain = !(nms_thr >= 1e-3f && nms_thr <= 1e3f);
    np.iou.r = np.iou.plain ? 0.0f : (float)((double)nms_thr / (1.0 + (double)nms_thr));
    np.max_det = max_det; np.pre_nms_top_k = pre_nms_top_k; np.key_floor_hi = key_floor_hi;
    np.kept_count = w.kept_count; np.kept_key = w.kept_key; np.kept_box = w.kept_box; np.kept_label = w.kept_label;
    np.status = status;
    np.timing = g_phase_timing.load(std::memory_order_relaxed) ? w.timing : nullptr;
    const size_t dyn = nms_dynamic_smem(max_det);
    int rc = nms_opt_in_shared_memory();
    if (rc) return rc;
    const int stages = g_stages.load(std::memory_order_relaxed);
    const bool slot = w.labels != nullptr;              // only class-agnostic filtering reads per-candidate labels
    if (!(stages & 2)) { /* measurement: NMS stage skipped */ }
    else if (decode) { if (slot) k_segment_nms<true, true><<<S, NMS_THREADS, dyn, s>>>(np); else k_segment_nms<true, false><<<S, NMS_THREADS, dyn, s>>>(np); }
    else { if (slot) k_segment_nms<false, true><<<S, NMS_THREADS, dyn, s>>>(np); else k_segment_nms<false, false><<<S, NMS_THREADS, dyn, s>>>(np); }
    rc = rn_check_launch("k_segment_nms");
    if (rc) return rc;
    MergeParams mp;
    mp.pages = B; mp.segs_per_page = segs_per_page; mp.max_det = max_det;
    mp.kept_count = w.kept_count; mp.kept_key = w.kept_key; mp.kept_box = w.kept_box; mp.kept_label = w.kept_label;
    mp.out_boxes = out_boxes; mp.out_scores = out_scores; mp.out_labels = out_labels; mp.out_indices = out_indices;
    mp.out_count = out_count;
    const size_t sel_smem = ((size_t)segs_per_page * max_det + NMS_CHUNK) * sizeof(unsigned long long);
    if (!(stages & 4)) {
        // measurement: merge stage skipped
    } else if (segs_per_page > 1 && segs_per_page <= 65535 && sel_smem <= MERGE_SMEM_MAX) {
        rc = merge_opt_in_shared_memory();
        if (rc) return rc;
        k_merge_topk_select<<<B, MERGE_THREADS, sel_smem, s>>>(mp);
    } else {
        k_merge_topk<<<B, 256, sizeof(int) * (size_t)segs_per_page, s>>>(mp);
    }
    return rn_check_launch("k_merge_topk");
}

int filter_common(K3Params kp, const BoxSource& src, bool decode, int nms, float nms_thr, int max_det, int pre_nms_top_k,
                  long long cand_cap, float* out_boxes, float* out_scores, int* out_labels, int* out_indices, int* status_out,
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
    const bool first_stage = (g_stages.load(std::memory_order_relaxed) & 1) != 0;    // (always, outside measurements)
    // timing, counts, kept_count, status are contiguous at the front of the workspace: they and the caller's status words are
    // zeroed by ONE small kernel that lets the threshold kernel launch behind it at once (rn_common.cuh)
    if (first_stage) {
        int rc0 = rn_reset_ints(reinterpret_cast<int*>(w.timing), (long long)(((char*)w.keys - (char*)w.timing) / sizeof(int)),
                                status_out, B, s);
        if (rc0) return rc0;
    }
    kp.sl.counts = w.counts; kp.sl.keys = w.keys; kp.sl.labels = w.labels; kp.sl.cap = cand_cap;
    RN_REQUIRE((long long)kp.N * C < (1ll << 31) - 4096, "N * C too large for one page");
    RN_REQUIRE(B <= 65535, "B must be <= 65535");
    kp.inv_c = 1.0f / (float)C;
    kp.vec_ok = (((long long)kp.N * C) % 4 == 0) ? 1 : 0;
    int rc = RN_OK;
    const long long page_tiles = ((long long)kp.N * C + K3S_TILE - 1) / K3S_TILE;
    if (!first_stage) {
        // measurement: the slabs of an earlier call are reused
    } else if (kp.class_specific && kp.vec_ok && C > 1 && C <= K3C_MAX_C) {
        // several classes: keys collected per class in shared memory, flushed in runs (k_threshold_keys_classes)
        long long per_page = (RN_NUM_SMS * K3S_CTAS_PER_SM) / B;
        const long long most = (page_tiles + K3_THREADS / 32 - 1) / (K3_THREADS / 32);
        if (per_page > most) per_page = most;
        if (per_page < 1) per_page = 1;
        const int tiles_per_cta = (int)((page_tiles + per_page - 1) / per_page);
        const dim3 grid((unsigned)((page_tiles + tiles_per_cta - 1) / tiles_per_cta), (unsigned)B);
        const size_t dyn = (size_t)C * (K3C_CAP * sizeof(unsigned long long) + sizeof(int)) + sizeof(unsigned) * (size_t)((C + 31) / 32);
        int rc1 = rn_launch_dependent("k_threshold_keys_classes", k_threshold_keys_classes, grid, dim3(K3_THREADS), dyn, s, kp, (int)page_tiles, tiles_per_cta);
        if (rc1) return rc1;
    } else if (kp.class_specific && kp.vec_ok) {
        // about SMs x 4 CTAs over all pages, each CTA a contiguous slice of one page (at least one tile per warp)
        // ONE wave: no more CTAs than the GPU holds at once (640 CTAs on 592 slots ran a second, nearly empty wave: 16 us
        // instead of 12)
        long long per_page = (RN_NUM_SMS * K3S_CTAS_PER_SM) / B;
        const long long most = (page_tiles + K3_THREADS / 32 - 1) / (K3_THREADS / 32);
        if (per_page > most) per_page = most;
        if (per_page < 1) per_page = 1;
        const int tiles_per_cta = (int)((page_tiles + per_page - 1) / per_page);
        const dim3 grid((unsigned)((page_tiles + tiles_per_cta - 1) / tiles_per_cta), (unsigned)B);
        int rc1 = rn_launch_dependent("k_threshold_keys_stream", k_threshold_keys_stream, grid, dim3(K3_THREADS), 0, s, kp, (int)page_tiles, tiles_per_cta);
        if (rc1) return rc1;
    } else {
        // class-agnostic filtering (max over classes per anchor) or unaligned page rows: the tile-per-CTA kernel
        const long long tiles = kp.class_specific ? ((long long)kp.N * C + K3_TILE - 1) / K3_TILE
                                                  : ((long long)kp.N + K3_THREADS * K3_VEC - 1) / (K3_THREADS * K3_VEC);
        const dim3 grid((unsigned)tiles, (unsigned)B);
        k_threshold_keys<<<grid, K3_THREADS, 0, s>>>(kp);
    }
    rc = rn_check_launch("k_threshold_keys");
    if (rc) return rc;
    return run_back_end(w, src, decode, B, S, spp, cand_cap, nms, nms_thr, max_det, pre_nms_top_k, host_f2ord(kp.thr),
                        out_boxes, out_scores, out_labels, out_indices, nullptr,
                        status_out ? status_out : w.status, s);
}

}  // namespace

extern "C" int rn_debug_nms_timing(int enable) {
    g_phase_timing.store(enable ? 1 : 0, std::memory_order_relaxed);
    return RN_OK;
}

extern "C" int rn_debug_filter_stages(int mask) {
    g_stages.store(mask & 7, std::memory_order_relaxed);
    return RN_OK;
}

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
    RN_REQUIRE(N >= 1 && N < (1ll << 31), "N out of range");
    K3Params kp = {};
    kp.cls = classification; kp.B = B; kp.N = (int)N; kp.C = C;
    kp.class_specific = class_specific ? 1 : 0; kp.thr = score_threshold;
    BoxSource src = {};
    src.rows = boxes; src.N = (int)N;
    return filter_common(kp, src, false, nms ? 1 : 0, nms_threshold, max_detections, pre_nms_top_k, cand_cap,
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
    BoxSource src = {};
    int rc = rn_make_levels(&src.lv, level_hw, level_stride, num_levels, anchors_per_cell);
    if (rc) return rc;
    RN_REQUIRE(src.lv.start[num_levels] == N, "N (%lld) does not match the level table (%d)", N, src.lv.start[num_levels]);
    src.rows = regression; src.base32 = base_anchors_f32_dev; src.N = (int)N;
    for (int i = 0; i < 4; ++i) { src.nm.mean[i] = mean4[i]; src.nm.std[i] = std4[i]; }
    src.clipW = clip_width; src.clipH = clip_height;
    K3Params kp = {};
    kp.cls = classification; kp.B = B; kp.N = (int)N; kp.C = C; kp.class_specific = class_specific ? 1 : 0; kp.thr = score_threshold;
    return filter_common(kp, src, true, nms ? 1 : 0, nms_threshold, max_detections, pre_nms_top_k, cand_cap,
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
    BoxSource src = {};
    src.rows = boxes; src.N = (int)cap;
    return run_back_end(w, src, false, 1, 1, 1, cap, 1, iou_threshold, max_output, 0, 0u,
                        sc_boxes, sc_scores, sc_labels, out_indices, out_count_dev, w.status, s);
}
