// Layout of a rank's peer mailbox (see peer_box.cu) and the device-side exchanges used by K2.
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
    // loss sums (RN_LOSS_PEER_LOSSES): [step % 4][rank][0 / 1] = step << 32 | float bits of the rank's focal / smooth-L1 sum
    unsigned long long loss_slots[RN_PEER_SLOTS][RN_MAX_WORLD][2];
    unsigned long long timeout_ns;                  // how long a wait may last before it gives up (rn_peer_box_set_timeout)
    unsigned int error;                             // STICKY: a wait of this rank timed out (rn_peer_box_status reads and clears it)
};

__device__ __forceinline__ unsigned long long rn_globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// One lane waits until the 64-bit word at `slot` (LOCAL memory, written by a peer over NVLink) carries `step` in its upper half
// and returns its lower half.  A peer that never publishes (crashed / stalled rank) must not hang the GPU: after the mailbox's
// timeout -- or as soon as ANOTHER wait of this rank has timed out -- the mailbox's sticky error flag is set and NaN bits are
// returned; the host turns the flag into an exception (PeerCounter.check), and the losses of the step are NaN.
__device__ __forceinline__ unsigned rn_peer_wait_word(const RnPeerBox* box, const volatile unsigned long long* slot, unsigned long long step) {
    const volatile RnPeerBox* b = box;
    unsigned long long w = *slot;
    if ((w >> 32) == (step & 0xffffffffull)) return (unsigned)(w & 0xffffffffull);
    const unsigned long long t0 = rn_globaltimer_ns(), limit = b->timeout_ns;
    unsigned spins = 0u;
    for (;;) {
        w = *slot;
        if ((w >> 32) == (step & 0xffffffffull)) return (unsigned)(w & 0xffffffffull);
        if ((++spins & 63u) == 0u) {                // the clock and the flag are looked at every 64 polls
            if (b->error != 0u || rn_globaltimer_ns() - t0 > limit) {
                atomicExch(const_cast<unsigned int*>(&box->error), 1u);
                return 0x7fc00000u;
            }
        }
    }
}

// Sum over ranks of the values published for this rank's current step (lag 0) or the one before it (lag 1).  Called by ONE
// WHOLE WARP: lane r waits for rank r's slot (all slots are polled concurrently, on LOCAL memory), then the values are added
// with shuffles -- they are integer-valued floats below 2^24, so the sum is exact and the same on every rank whatever the order.
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
    if (lane < world) v = __uint_as_float(rn_peer_wait_word(box, &b->slots[step & (RN_PEER_SLOTS - 1)][lane], step));
    return rn_warp_sum(v);
}

// The loss sums of step `step`, exchanged the same way at the END of the loss kernel (RN_LOSS_PEER_LOSSES): called by ONE WHOLE
// WARP of the kernel's last CTA with this rank's two sums; lane r stores them into rank r's mailbox, waits for rank r's pair in
// its own, and the pairs are added in RANK ORDER in fp64 -- every rank gets the same bits.  (A rank's last CTA finishes within
// the skew of the ranks' loss kernels, which all started from the same count exchange.)
__device__ __forceinline__ void rn_peer_box_sum_losses_warp(const RnPeerBox* box, unsigned long long step, double& focal, double& sl1) {
    const volatile RnPeerBox* b = box;
    const int world = b->world, lane = threadIdx.x & 31;
    const int s = (int)(step & (RN_PEER_SLOTS - 1));
    float f = 0.0f, l = 0.0f;
    if (lane < world) {
        volatile unsigned long long* dst = &box->peers[lane]->loss_slots[s][box->rank][0];
        dst[0] = (step << 32) | (unsigned long long)__float_as_uint((float)focal);
        dst[1] = (step << 32) | (unsigned long long)__float_as_uint((float)sl1);
        f = __uint_as_float(rn_peer_wait_word(box, &b->loss_slots[s][lane][0], step));
        l = __uint_as_float(rn_peer_wait_word(box, &b->loss_slots[s][lane][1], step));
    }
    double tf = 0.0, ts = 0.0;
    for (int r = 0; r < world; ++r) {               // fixed order: identical sums on every rank
        tf += (double)__shfl_sync(0xffffffffu, f, r);
        ts += (double)__shfl_sync(0xffffffffu, l, r);
    }
    focal = tf; sl1 = ts;
}
