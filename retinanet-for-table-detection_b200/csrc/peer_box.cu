// Peer mailbox: the one exchange step of the path (the batch-global positive-anchor count, C1 in DESIGN.md) done
// over NVLink peer memory instead of a NCCL all-reduce.
//
// Every rank owns a mailbox in its own HBM, mapped into the other ranks of the node through CUDA IPC.  After K1
// a one-CTA kernel bumps the rank's step counter and stores {step, count} as ONE 64-bit word into slot `rank` of
// EVERY rank's mailbox (P2P stores over NVLink / NVSwitch, one per peer).  K2's prologue (csrc/losses.cu) reads
// its own mailbox -- local memory -- until all `world` slots carry the current step, and adds the counts in rank
// order (integer-valued floats: exact and identical on every rank).  No NCCL launch, no host round trip, and both
// kernels stay capturable in CUDA graphs (the step number lives in device memory, not in a kernel argument).
//
// Fused publish (rn_peer_box_bind + RN_LOSS_PEER_PUBLISH, the in-order schedule's default): the publish kernel
// disappears -- K2's CTA 0 stores the word into every mailbox in its prologue, right before all CTAs start waiting,
// and K2's last CTA bumps the step counter; see rn_peer_box_sum_warp().  Send, wait and the loss arithmetic are ONE kernel.
//
// With RN_LOSS_PEER_LOSSES the two loss sums make the same trip at the END of the loss kernel (its last CTA; loss_slots), so
// that every rank's `losses` is the loss of the merged batch.  A wait that outlasts the mailbox's timeout (default 30 s,
// rn_peer_box_set_timeout) sets a STICKY error flag (rn_peer_box_status) and yields NaN losses; it never hangs the GPU.
//
// Slots are indexed by step % 4.  Two schedules are supported, identical on all ranks:
//   in order    K1(s) publish(s) K2(s) K1(s+1) publish(s+1) K2(s+1) ...           K2 reads its latest step (lag 0)
//   pipelined   K1(s+1) publish(s+1) K2(s) K1(s+2) publish(s+2) K2(s+1) ...       K2 reads the step before (lag 1):
//               the next batch's targets and count are produced while this batch's losses wait for nothing --
//               the exchange latency and the skew between ranks leave the critical path.
// Overwrite safety (pipelined, the stricter case): rank r stores step s+4 into slot s % 4 only after its own K2(s+2)
// has returned, which has seen every rank p's publish(s+2), which p issues after its K2(s) in stream order -- so
// every rank has consumed step s before anyone overwrites it.  All ranks must run the same sequence of steps.
#include <stddef.h>
#include "rn_common.cuh"
#include "peer_box.cuh"

namespace {

struct PeerPtrs { RnPeerBox* p[RN_MAX_WORLD]; };

__global__ void k_peer_publish(const float* value, RnPeerBox* local, const PeerPtrs peers, int rank, int world) {
    __shared__ unsigned long long s_step;
    if (threadIdx.x == 0) {
        s_step = local->step + 1ull;
        local->step = s_step;                       // read by this rank's K2, later in the same stream
    }
    __syncthreads();
    const unsigned long long step = s_step;
    const unsigned long long word = (step << 32) | (unsigned long long)__float_as_uint(__ldcg(value));
    if ((int)threadIdx.x < world) {
        volatile unsigned long long* slot = &peers.p[threadIdx.x]->slots[step & (RN_PEER_SLOTS - 1)][rank];
        *slot = word;                               // one aligned 8-byte store per peer: count and step arrive together
    }
}

}  // namespace

extern "C" size_t rn_peer_box_bytes(void) { return sizeof(RnPeerBox); }

extern "C" int rn_peer_box_create(int world, void** box_out, void* ipc_handle_out64) {
    RN_REQUIRE(world >= 1 && world <= RN_MAX_WORLD, "world must be in [1, %d]", RN_MAX_WORLD);
    RN_REQUIRE(box_out && ipc_handle_out64, "NULL pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(RnPeerBox));
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
    RnPeerBox init = {};
    init.world = world;
    init.timeout_ns = 30ull * 1000000000ull;          // see rn_peer_box_set_timeout
    e = cudaMemcpy(d, &init, sizeof(init), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(ipc_handle_out64), d);
    if (e != cudaSuccess) { cudaFree(d); return rn_fail(RN_ERR_CUDA, "peer box setup: %s", cudaGetErrorString(e)); }
    *box_out = d;
    return RN_OK;
}

extern "C" int rn_peer_box_set_timeout(void* local_box, double seconds) {
    RN_REQUIRE(local_box, "NULL pointer");
    RN_REQUIRE(seconds > 0.0 && seconds < 1e9, "timeout must be positive");
    const unsigned long long ns = (unsigned long long)(seconds * 1e9);
    cudaError_t e = cudaMemcpy(reinterpret_cast<char*>(local_box) + offsetof(RnPeerBox, timeout_ns), &ns, sizeof(ns), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "peer box timeout: %s", cudaGetErrorString(e));
    return RN_OK;
}

extern "C" int rn_peer_box_status(void* local_box, int* timed_out, int clear) {
    RN_REQUIRE(local_box && timed_out, "NULL pointer");
    unsigned int flag = 0u;
    char* at = reinterpret_cast<char*>(local_box) + offsetof(RnPeerBox, error);
    cudaError_t e = cudaMemcpy(&flag, at, sizeof(flag), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && flag && clear) { const unsigned int zero = 0u; e = cudaMemcpy(at, &zero, sizeof(zero), cudaMemcpyHostToDevice); }
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "peer box status: %s", cudaGetErrorString(e));
    *timed_out = flag ? 1 : 0;
    return RN_OK;
}

extern "C" int rn_peer_box_step(const void* local_box, unsigned long long* step_out) {
    RN_REQUIRE(local_box && step_out, "NULL pointer");
    cudaError_t e = cudaMemcpy(step_out, reinterpret_cast<const char*>(local_box) + offsetof(RnPeerBox, step), sizeof(unsigned long long),
                               cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "peer box step: %s", cudaGetErrorString(e));
    return RN_OK;
}

extern "C" int rn_peer_box_bind(void* local_box, void* const* boxes_of_all_ranks, int rank, int world,
                                const float* value_even_dev, const float* value_odd_dev) {
    RN_REQUIRE(local_box && boxes_of_all_ranks && value_even_dev && value_odd_dev, "NULL pointer");
    RN_REQUIRE(world >= 1 && world <= RN_MAX_WORLD && rank >= 0 && rank < world, "bad rank / world (%d / %d)", rank, world);
    RN_REQUIRE(boxes_of_all_ranks[rank] == local_box, "boxes_of_all_ranks[rank] must be the local box");
    RnPeerBox host = {};
    host.rank = rank;
    host.value[0] = value_even_dev;
    host.value[1] = value_odd_dev;
    for (int r = 0; r < world; ++r) {
        RN_REQUIRE(boxes_of_all_ranks[r] != nullptr, "box of rank %d is NULL", r);
        host.peers[r] = reinterpret_cast<RnPeerBox*>(boxes_of_all_ranks[r]);
    }
    // rank, value and peers are contiguous in the struct: one copy, `step`, `world` and the slots stay untouched
    const size_t off = offsetof(RnPeerBox, rank), len = offsetof(RnPeerBox, slots) - off;
    cudaError_t e = cudaMemcpy(reinterpret_cast<char*>(local_box) + off, reinterpret_cast<const char*>(&host) + off, len,
                               cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "peer box bind: %s", cudaGetErrorString(e));
    return RN_OK;
}

extern "C" int rn_peer_box_connect(void* local_box, void* const* boxes_of_all_ranks, int rank, int world) {
    RN_REQUIRE(local_box && boxes_of_all_ranks, "NULL pointer");
    RN_REQUIRE(world >= 1 && world <= RN_MAX_WORLD && rank >= 0 && rank < world, "bad rank / world (%d / %d)", rank, world);
    RN_REQUIRE(boxes_of_all_ranks[rank] == local_box, "boxes_of_all_ranks[rank] must be the local box");
    RnPeerBox host = {};
    for (int r = 0; r < world; ++r) {
        RN_REQUIRE(boxes_of_all_ranks[r] != nullptr, "box of rank %d is NULL", r);
        host.peers[r] = reinterpret_cast<RnPeerBox*>(boxes_of_all_ranks[r]);
    }
    cudaError_t e = cudaMemcpy(reinterpret_cast<char*>(local_box) + offsetof(RnPeerBox, rank), &rank, sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = cudaMemcpy(reinterpret_cast<char*>(local_box) + offsetof(RnPeerBox, peers), host.peers, sizeof(host.peers), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "peer box connect: %s", cudaGetErrorString(e));
    return RN_OK;
}

extern "C" int rn_peer_box_open(const void* ipc_handle64, void** peer_box_out) {
    RN_REQUIRE(ipc_handle64 && peer_box_out, "NULL pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, sizeof(h));
    void* d = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&d, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    *peer_box_out = d;
    return RN_OK;
}

extern "C" int rn_peer_box_close(void* peer_box) {
    if (!peer_box) return RN_OK;
    cudaError_t e = cudaIpcCloseMemHandle(peer_box);
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
    return RN_OK;
}

extern "C" int rn_peer_box_destroy(void* box) {
    if (!box) return RN_OK;
    cudaError_t e = cudaFree(box);
    if (e != cudaSuccess) return rn_fail(RN_ERR_CUDA, "cudaFree: %s", cudaGetErrorString(e));
    return RN_OK;
}

extern "C" int rn_peer_publish(const float* value_dev, void* local_box, void* const* boxes_of_all_ranks, int rank, int world,
                               void* stream) {
    RN_REQUIRE(value_dev && local_box && boxes_of_all_ranks, "NULL pointer");
    RN_REQUIRE(world >= 1 && world <= RN_MAX_WORLD && rank >= 0 && rank < world, "bad rank / world (%d / %d)", rank, world);
    PeerPtrs pp;
    for (int r = 0; r < RN_MAX_WORLD; ++r) pp.p[r] = nullptr;
    for (int r = 0; r < world; ++r) {
        RN_REQUIRE(boxes_of_all_ranks[r] != nullptr, "box of rank %d is NULL", r);
        pp.p[r] = reinterpret_cast<RnPeerBox*>(boxes_of_all_ranks[r]);
    }
    RN_REQUIRE(boxes_of_all_ranks[rank] == local_box, "boxes_of_all_ranks[rank] must be the local box");
    k_peer_publish<<<1, 32, 0, (cudaStream_t)stream>>>(value_dev, reinterpret_cast<RnPeerBox*>(local_box), pp, rank, world);
    return rn_check_launch("rn_peer_publish");
}
