// Layout of a rank's peer mailbox (see peer_box.cu) and the device-side read used by K2's prologue.
#pragma once
#include "rn_common.cuh"

#define RN_MAX_WORLD 16
#define RN_PEER_SLOTS 4      // steps in flight: see the overwrite argument in peer_box.cu

struct RnPeerBox {
    unsigned long long step;                        // last step this rank published (bumped on the device)
    int world;
    int rank;                                       // fused publish (rn_peer_box_bind): this rank,
    const float* value[2];                          //   the device float it publishes for a step: value[step & 1] (K1's
                                                    //   positive count; two so that double-buffered targets alternate),
    RnPeerBox* peers[RN_MAX_WORLD];                 //   and every rank's mailbox as mapped into this process
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
//
// publish = true (fused publish, K2 launched with RN_LOSS_PEER_PUBLISH): there is no separate publish kernel.  `step`
// then counts COMPLETED steps -- the last CTA of the kernel bumps it, see finish_block() in losses.cu -- and the step
// being exchanged is step + 1: warp 0 of CTA 0 first stores this rank's {step + 1, value} into every rank's mailbox
// (P2P stores over NVLink, lane r -> rank r), then every CTA waits for all ranks' words of step + 1 as usual.
__device__ __forceinline__ float rn_peer_box_sum_warp(const RnPeerBox* box, int lag, bool publish = false) {
    const volatile RnPeerBox* b = box;
    const unsigned long long step = publish ? b->step + 1ull : b->step - (unsigned long long)lag;   // lag 1: the step published before the latest one
    const int world = b->world, lane = threadIdx.x & 31;
    if (publish && blockIdx.x == 0 && lane < world) {
        const unsigned long long word = (step << 32) | (unsigned long long)__float_as_uint(__ldcg(box->value[step & 1]));
        volatile unsigned long long* slot = &box->peers[lane]->slots[step & (RN_PEER_SLOTS - 1)][box->rank];
        *slot = word;                               // one aligned 8-byte store per peer: count and step arrive together
    }
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
