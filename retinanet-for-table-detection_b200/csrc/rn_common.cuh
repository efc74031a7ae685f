// Shared helpers for librn_b200 (sm_100a).  Internal header; the public C-ABI is include/rn_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/rn_b200.h"

#define RN_NUM_SMS 148   // B200: 2 dies x 74 SMs; grids of streaming kernels are sized in multiples of this

// ---------------------------------------------------------------------------------------------
// error reporting (thread-local message, negative return codes; nothing throws)
// ---------------------------------------------------------------------------------------------
char* rn_error_buffer();

static inline int rn_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(rn_error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define RN_REQUIRE(cond, ...) do { if (!(cond)) return rn_fail(RN_ERR_BAD_ARG, __VA_ARGS__); } while (0)

static inline int rn_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return RN_OK;
}

static inline bool rn_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------------------------
// pyramid level table, passed to kernels by value
// ---------------------------------------------------------------------------------------------
struct RnLevels {
    int num_levels;
    int anchors_per_cell;
    int h[RN_MAX_LEVELS];
    int w[RN_MAX_LEVELS];
    int stride[RN_MAX_LEVELS];
    int start[RN_MAX_LEVELS + 1];   // first anchor index of each level; start[num_levels] = N
    float inv_w[RN_MAX_LEVELS];     // 1 / w[l], for the float-assisted exact integer division below
    float inv_a;                    // 1 / anchors_per_cell
};

static inline int rn_make_levels(RnLevels* t, const int* level_hw, const int* level_stride,
                                 int num_levels, int anchors_per_cell) {
    RN_REQUIRE(level_hw && level_stride, "level table is NULL");
    RN_REQUIRE(num_levels >= 1 && num_levels <= RN_MAX_LEVELS, "num_levels must be in [1, %d]", RN_MAX_LEVELS);
    RN_REQUIRE(anchors_per_cell >= 1 && anchors_per_cell <= 64, "anchors_per_cell must be in [1, 64]");
    t->num_levels = num_levels;
    t->anchors_per_cell = anchors_per_cell;
    long long n = 0;
    for (int l = 0; l < RN_MAX_LEVELS; ++l) { t->h[l] = 0; t->w[l] = 0; t->stride[l] = 1; t->start[l] = 0; t->inv_w[l] = 1.0f; }
    t->inv_a = 1.0f / (float)anchors_per_cell;
    for (int l = 0; l < num_levels; ++l) {
        RN_REQUIRE(level_hw[2 * l] >= 0 && level_hw[2 * l + 1] >= 0 && level_stride[l] > 0, "bad level %d", l);
        t->h[l] = level_hw[2 * l];
        t->w[l] = level_hw[2 * l + 1];
        t->stride[l] = level_stride[l];
        t->inv_w[l] = t->w[l] > 0 ? 1.0f / (float)t->w[l] : 1.0f;
        t->start[l] = (int)n;
        n += (long long)t->h[l] * t->w[l] * anchors_per_cell;
        RN_REQUIRE(n < (1ll << 31), "too many anchors per page");
    }
    for (int l = num_levels; l <= RN_MAX_LEVELS; ++l) t->start[l] = (int)n;
    return RN_OK;
}

// exact r / d for 0 <= r, d > 0: float estimate (inv = 1/d) + one correction step; the estimate is within
// +-1 of the quotient while r < 2^24, a hardware integer division is used beyond that
__device__ __forceinline__ int rn_div(int r, int d, float inv) {
    if (r >= (1 << 24)) return r / d;
    int q = (int)((float)r * inv);
    const int rem = r - q * d;
    if (rem < 0) --q;
    else if (rem >= d) ++q;
    return q;
}

// anchor index -> (level, cell x, cell y, anchor-in-cell).  Order: level-major, then y, x, a
// (reference model/anchors.py:195-202, :232-236).
__device__ __forceinline__ void rn_locate(const RnLevels& t, int n, int& level, int& cx, int& cy, int& a) {
    level = 0;
#pragma unroll
    for (int l = 1; l < RN_MAX_LEVELS; ++l)
        if (l < t.num_levels && n >= t.start[l]) level = l;
    const int r = n - t.start[level];
    const int cell = rn_div(r, t.anchors_per_cell, t.inv_a);
    a = r - cell * t.anchors_per_cell;
    cy = rn_div(cell, t.w[level], t.inv_w[level]);
    cx = cell - cy * t.w[level];
}

// ---------------------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T rn_warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float rn_warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float rn_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming (evict-first) 128-bit global accesses: the big tensors on this path are touched once
__device__ __forceinline__ float4 rn_ldg_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void rn_stg_stream4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch (sm_90+): "reset, then the real kernel" without the launch latency in between.
// k_rn_reset zeroes up to two small int ranges (counters, status words) and allows its dependents to launch at once; the
// kernel behind it is launched with rn_launch_dependent (programmatic stream serialisation) so that its CTAs are placed and
// run their prologue while the reset is still in flight, and calls rn_grid_dependency_wait() before it first touches what the
// reset writes.  None of the other kernels triggers early, so a dependent launched behind one of them is ordered like a plain
// launch.  Works eagerly and under stream capture (programmatic edges in the graph).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rn_grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void rn_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

static __global__ void k_rn_reset(int* a, int na, int* b, int nb) {
    rn_launch_dependents();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < na; i += gridDim.x * blockDim.x) a[i] = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += gridDim.x * blockDim.x) b[i] = 0;
}

// zero `na` ints at a and `nb` ints at b (either may be empty) on the stream; one launch
static inline int rn_reset_ints(int* a, long long na, int* b, long long nb, cudaStream_t s) {
    if (a == nullptr) na = 0;
    if (b == nullptr) nb = 0;
    if (na + nb <= 0) return RN_OK;
    const long long most = na > nb ? na : nb;
    const int blocks = (int)(most > 256 * 64 ? 64 : (most + 255) / 256);
    k_rn_reset<<<blocks, 256, 0, s>>>(a, (int)na, b, (int)nb);
    return rn_check_launch("k_rn_reset");
}

template <typename... KArgs, typename... Args>
static inline int rn_launch_dependent(const char* what, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                      cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return rn_check_launch(what);
}
