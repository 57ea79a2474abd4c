// Shared between the streaming-scan search (flatip.cu) and the tensor-core batch search
// (flatip_batch.cu): id mapping, cross-GPU result delivery, the exact fp32 dot product both
// paths score with.
#pragma once
#include "common.cuh"

namespace cb {

constexpr int kD = 512;
constexpr int kMaxRanks = 16;           // shards that can deliver into one mailbox

// local row -> global id.  A shard filled by several add() calls of a multi-shard index holds
// several runs of consecutive global ids: segment j covers local rows [seg_local[j], seg_local[j+1])
// and starts at global id seg_global[j].  nseg == 0: global = id_base + local.
struct IdMap {
    const int64_t *seg_local = nullptr;
    const int64_t *seg_global = nullptr;
    int nseg = 0;
    int64_t id_base = 0;
};
__device__ __forceinline__ int64_t map_id(const IdMap &m, uint32_t local) {
    if (m.nseg == 0) return m.id_base + (int64_t)local;
    int lo = 0, hi = m.nseg - 1;            // last segment whose first local row is <= local
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(m.seg_local + mid) <= (int64_t)local) lo = mid; else hi = mid - 1;
    }
    return __ldg(m.seg_global + lo) + ((int64_t)local - __ldg(m.seg_local + lo));
}

// Cross-GPU delivery of a shard's sorted top-k (the one exchange step of the
// sharded search).  The block that writes a query's (D, I) list writes it STRAIGHT INTO the
// root GPU's mailbox over NVLink (peer stores; the pointers it is given already point there),
// then bumps the root's per-rank counter with a system-scope release.  The root's merge kernel
// acquires those counters and merges.  `consumed` (in the root's memory) is the back-pressure:
// a slot is rewritten only after the merge that read it has finished.  All counters are
// cumulative query counts, compared wrap-safe.  done == nullptr: plain local search.
struct PeerOut {
    uint32_t *done = nullptr;            // root mailbox: queries delivered by this rank
    const uint32_t *consumed = nullptr;  // root mailbox: queries merged by the root
    uint32_t need_consumed = 0;          // this slot is free once *consumed >= need_consumed
    uint32_t *error = nullptr;           // root mailbox: sticky, set when a wait timed out
};

constexpr uint64_t kPeerTimeoutNs = 20ull * 1000 * 1000 * 1000;   // a dead peer becomes an error, not a hang

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_sys_add(uint32_t *p, uint32_t v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// one thread spins until (*ctr - need) >= 0 (wrap-safe)
__device__ __forceinline__ void spin_until(const uint32_t *ctr, uint32_t need, uint32_t *error) {
    uint32_t spins = 0;
    uint64_t t0 = 0;
    while ((int32_t)(ld_acquire_sys(ctr) - need) < 0) {
        if ((++spins & 0x3ff) == 0) {
            const uint64_t now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kPeerTimeoutNs) { if (error) atomicExch(error, 1u); break; }
        }
    }
}
// whole block: wait until the destination slot may be overwritten (call before writing D / I).
// `already_free`: an earlier probe (peer_slot_probe) saw the slot free -- the counter only grows, so
// the NVLink round trip is skipped.
__device__ __forceinline__ void peer_wait_slot(const PeerOut &po, bool already_free = false) {
    if (po.done == nullptr) return;
    if (threadIdx.x == 0 && !already_free) spin_until(po.consumed, po.need_consumed, po.error);
    __syncthreads();
}
// one thread, non-blocking: is the destination slot free already?  Issued at kernel entry so that the
// remote read overlaps the pass over the shard instead of sitting on the critical path at its end.
__device__ __forceinline__ bool peer_slot_probe(const PeerOut &po) {
    if (po.done == nullptr) return true;
    return (int32_t)(ld_acquire_sys(po.consumed) - po.need_consumed) >= 0;
}
// whole block: publish `n` finished queries (call after the block's last D / I store)
__device__ __forceinline__ void peer_signal(const PeerOut &po, uint32_t n) {
    if (po.done == nullptr) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        red_release_sys_add(po.done, n);
    }
}

// ---- the exact score: fp32 <q, x> in ONE fixed summation order ---------------------------------
// A lane owns 16 of the 512 elements (fp16 rows: 16-byte chunks `lane` and `lane + 32`), multiplies
// them into a sequential fmaf chain per chunk, adds its chunks, and the 32 lane sums are combined by
// the xor butterfly 16, 8, 4, 2, 1.  The streaming scan and the batch path's re-scoring both use
// exactly this order, so a (query, row) pair has ONE score in the whole library, bit for bit.
__device__ __forceinline__ float dot8_h(const uint4 &u, const float *q) {
    const __half2 *h = reinterpret_cast<const __half2 *>(&u);
    float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
    float2 c = __half22float2(h[2]), d = __half22float2(h[3]);
    float s = a.x * q[0];
    s = fmaf(a.y, q[1], s);
    s = fmaf(b.x, q[2], s);
    s = fmaf(b.y, q[3], s);
    s = fmaf(c.x, q[4], s);
    s = fmaf(c.y, q[5], s);
    s = fmaf(d.x, q[6], s);
    s = fmaf(d.y, q[7], s);
    return s;
}
__device__ __forceinline__ float dot4_f(const uint4 &u, const float *q) {
    float s = __uint_as_float(u.x) * q[0];
    s = fmaf(__uint_as_float(u.y), q[1], s);
    s = fmaf(__uint_as_float(u.z), q[2], s);
    s = fmaf(__uint_as_float(u.w), q[3], s);
    return s;
}
// this lane's 16 query values for an fp16-stored row (chunks lane, lane + 32)
__device__ __forceinline__ void load_q_lane_f16(const float *q, int lane, float (&qr)[16]) {
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
        for (int e = 0; e < 8; e++) qr[c * 8 + e] = q[(lane + 32 * c) * 8 + e];
}
// full-warp exact score of one fp16 row (all lanes return the same value)
__device__ __forceinline__ float exact_score_f16(const uint4 *row, const float (&qr)[16], int lane) {
    const uint4 v0 = ld_stream_v4(row + lane), v1 = ld_stream_v4(row + lane + 32);
    float s = 0.f;
    s += dot8_h(v0, &qr[0]);
    s += dot8_h(v1, &qr[8]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s + 0.0f;
}

__device__ __forceinline__ uint64_t make_comp(uint32_t key, uint32_t id) {
    return ((uint64_t)key << 32) | (uint64_t)(0xffffffffu - id);
}

}  // namespace cb
