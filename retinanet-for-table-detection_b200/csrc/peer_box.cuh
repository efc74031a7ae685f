// Layout of a rank's peer mailbox (see peer_box.cu) and the device-side read used by K2's prologue.
#pragma once
#include "rn_common.cuh"

#define RN_MAX_WORLD 16
#define RN_PEER_SLOTS 4      // steps in flight: see the overwrite argument in peer_box.cu

struct RnPeerBox {
    unsigned long long step;                        // last step this rank published (bumped on the device)
    int world;
    int pad;
    unsigned long long slots[RN_PEER_SLOTS][RN_MAX_WORLD];   // [step % 4][rank] = step << 32 | float bits of the rank's value
};

__device__ __forceinline__ unsigned long long rn_globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Sum over ranks of the values published for this rank's current step (lag 0) or the one before it (lag 1).  Called by ONE WHOLE WARP: lane r waits for
// rank r's slot (all slots are polled concurrently, on LOCAL memory), then the values are added with shuffles --
// they are integer-valued floats below 2^24, so the sum is exact and the same on every rank whatever the order.
// A peer that never publishes (crashed rank) must not hang the GPU: after ~2 s the result is NaN, which the
// caller's losses then carry.
__device__ __forceinline__ float rn_peer_box_sum_warp(const RnPeerBox* box, int lag) {
    const volatile RnPeerBox* b = box;
    const unsigned long long step = b->step - (unsigned long long)lag;   // lag 1: the step published before the latest one
    const int world = b->world, lane = threadIdx.x & 31;
    float v = 0.0f;
    if (lane < world) {
        const unsigned long long t0 = rn_globaltimer_ns();
        unsigned long long w = b->slots[step & (RN_PEER_SLOTS - 1)][lane];
        while ((w >> 32) != (step & 0xffffffffull)) {
            if (rn_globaltimer_ns() - t0 > 2000000000ull) { w = 0x7fc00000ull; break; }
            w = b->slots[step & (RN_PEER_SLOTS - 1)][lane];
        }
        v = __uint_as_float((unsigned)(w & 0xffffffffull));
    }
    return rn_warp_sum(v);
}
