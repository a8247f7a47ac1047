// gt_peer.h — NVLink peer windows: device buffers that every member of a row / column group can write
// into directly (see gt_peer.cu).
#pragma once
#include "gt_internal.h"

namespace gt {

// One device allocation per group member, each mapped into every other member's address space
// (cudaIpc).  Layout of every member's allocation is the same: `data_bytes` of payload followed by one
// 32-bit arrival counter per group member, 16 bytes apart.
struct PeerWindow {
    CommGroup grp = COMM_WORLD;
    int size = 1, me = 0;
    size_t data_bytes = 0, total_bytes = 0;
    uint8_t* local = nullptr;
    std::vector<uint8_t*> remote;          // [size], remote[me] == local
    uint32_t* flag(int member, int from) const {              // counter on `member` that `from` advances
        return (uint32_t*) (remote[member] + data_bytes) + 4 * from;
    }
};

// Collective over the WORLD communicator (every rank calls it the same number of times, in the same order);
// the window itself spans `grp`.  Returns nullptr on every rank if any rank could not map its peers.
PeerWindow* peer_window_create(gt_ctx* ctx, CommGroup grp, size_t data_bytes);
void peer_window_destroy(gt_ctx* ctx, PeerWindow* w);

// Puts are issued between peer_put_begin (every put stream waits for `ready`, recorded by the producer of the bytes) and
// peer_put_end (done[stream] recorded, or nothing if done == nullptr; peer_puts_done makes a stream wait for all of them).
void peer_put_begin(gt_ctx* ctx, cudaEvent_t ready);
// dst_member's copy of the window <- local bytes by a copy engine over NVLink, followed by dst_member's counter for
// this rank <- value on the same stream, so the counter lands after the payload.  Every destination has its own
// stream (GT_PEER_LANES of them, default 4), so the puts to the 3 column-group peers at p = 8 are driven by different
// copy engines at the same time instead of queueing behind each other.
// `advance` = false: payload only, the counter stays (an extra piece that a later put on the same stream publishes)
void peer_put(gt_ctx* ctx, const PeerWindow* w, int dst_member, size_t dst_offset, const void* src, size_t bytes, uint32_t value, bool advance = true);
void peer_put_end(gt_ctx* ctx, cudaEvent_t* done);
void peer_puts_done(gt_ctx* ctx, cudaEvent_t* done, cudaStream_t s);
// blocks `s` (on the device) until every other member's counters in the local window have reached `value`
void peer_wait_all(gt_ctx* ctx, const PeerWindow* w, uint32_t value, cudaStream_t s);
// device-side error word: 0, or 1 + the member whose counter did not arrive within the timeout
uint32_t* peer_error_word(gt_ctx* ctx);

// Arrival counters count modulo kPeerSeqLen (the copy engines need the value in memory: a table of that length);
// a consumer is never more than a few epochs away from its producers, so "reached" is a modular comparison and the
// epoch numbers of a program may wrap as often as they like.
constexpr uint32_t kPeerSeqLen = 1u << 20;
__host__ __device__ __forceinline__ bool peer_reached(uint32_t counter, uint32_t value) {
    return ((counter - value) & (kPeerSeqLen - 1)) < kPeerSeqLen / 2;
}
// world-wide ordering point on stream `s` (a 1-element all-reduce): everything every rank enqueued before it has
// completed on the device before anything enqueued after it starts on any rank
void peer_fence_world(gt_ctx* ctx, cudaStream_t s);

}  // namespace gt
