// sk_device.cuh -- device-side structures and small helpers shared by the kernels.
// sm_100a only.  See DESIGN.md for the data layout.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace sk {

#ifdef SK_PHASE_TIMING
__device__ unsigned long long g_walk_dbg[4];   // debug: walks, steps, threads that had to spin, spin iterations
#endif

constexpr int kMaxStreams = 3;

// Trimming parameters in the form the kernels want (built from sk_params on the host).
struct DevParams {
    int32_t qoff, qmin, qmax;  // quality_constants row, reference src/sickle.h:85-91
    int32_t qthr;              // -q
    int32_t lthr;              // -l
    int32_t no_fiveprime;      // -x
    int32_t trunc_n;           // -n
    int32_t mode;              // SK_MODE_*
    int32_t emu_threads;       // >= 1
    int32_t has_singles;
};

// One input buffer of a batch and its line index.
struct DevInput {
    const uint8_t *data;   // 16-byte aligned; 16 bytes of readable padding after nbytes
    uint32_t first;        // 0..15: the batch starts at data[first] (bytes before it are ignored)
    uint32_t nbytes;       // the batch ends at data[nbytes]
    uint32_t *line_end;    // line_end[l] = byte offset of the l-th '\n'
    uint32_t line_cap;     // capacity of line_end
};

// Per-record descriptor written by K2 and consumed by K3 (16 bytes).
//   dst_off : byte offset of the record inside its output stream
//   five    : 5' cut (first kept base)
//   nkeep   : number of kept bases (three - five)
//   route   : bits 0-1 output stream, bit 2 = emit, bit 3 = emit as "N record" (-M)
struct __align__(16) RecDesc {
    uint32_t dst_off;
    uint32_t five;
    uint32_t nkeep;
    uint32_t route;
};
constexpr uint32_t kRouteEmit = 4u;
constexpr uint32_t kRouteNRec = 8u;

// Device-resident control block of a slot; reset by k_finalize at the end of every batch.
struct Control {
    uint32_t tile_counter[4];     // dynamic tile tickets: K1 input 0, K1 input 1, K2, K3
    uint32_t nlines[2];           // written by K1's last tile
    uint32_t index_overflow;      // bit 0: a line index ran out of capacity; bit 1: an output buffer did
    uint32_t fast_fail;           // fused path met something it does not handle (bit 0; bit 1: too many records in a tile): re-run on the general path
    uint32_t fast_consumed;       // fused path: end of the last complete unit (max over tiles)
    uint32_t fast_records;        // fused path: complete records
    uint32_t fast_records2[2];    // fused path, two files: complete records per file (PASS 1)
    uint32_t fast_consumed2[2];   // fused path, two files: end of the last paired record per file (PASS 2)
    uint32_t fused_nunits;        // fused path, two files: pairs of the batch (kf2_between)
    uint32_t k2a_ticket;          // k2_trim_only: next pair of units
    unsigned long long err_key;   // min over offending (class, unit, mate, position); ~0 = none
    unsigned long long counters[8];  // kept, discard, kept_p, discard_p, kept_s1, kept_s2, discard_s1, discard_s2
    unsigned long long out_bytes[kMaxStreams];
};
constexpr unsigned long long kNoError = ~0ull;

// Summary produced by the finalize kernel and copied to the host (mirrors sk_result's scalars).
struct DevResult {
    unsigned long long out_bytes[kMaxStreams];
    unsigned long long consumed[2];
    unsigned long long records[2];
    long long counters[8];
    int32_t err_kind, err_file;
    long long err_record;
    int32_t err_position, err_byte;
    unsigned long long err_line_off[4], err_line_len[4];
    uint32_t index_overflow;
    uint32_t pad;
};

// ---- decoupled look-back tile status word ------------------------------------------------------
//   [63:62] flag (0 = not ready, 1 = tile aggregate, 2 = inclusive prefix)
//   [61:34] epoch of the batch that wrote it (a word from another epoch counts as "not ready",
//           so the status arrays are never cleared between batches)
//   [33:0]  value (line counts / output bytes of one batch: < 2^32)
constexpr unsigned long long kFlagAggregate = 1ull;
constexpr unsigned long long kFlagInclusive = 2ull;
constexpr int kEpochShift = 34;
constexpr unsigned long long kEpochMask = (1ull << 28) - 1;
constexpr unsigned long long kValueMask = (1ull << 34) - 1;

__device__ __forceinline__ unsigned long long pack_status(unsigned long long flag, uint32_t epoch,
                                                          unsigned long long value) {
    return (flag << 62) | (((unsigned long long)epoch & kEpochMask) << kEpochShift) | (value & kValueMask);
}
// flag of a status word as seen from `epoch` (0 if the word is stale)
__device__ __forceinline__ uint32_t status_flag(unsigned long long w, uint32_t epoch) {
    return (((w >> kEpochShift) & kEpochMask) == ((unsigned long long)epoch & kEpochMask)) ? (uint32_t)(w >> 62) : 0u;
}
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
    unsigned long long v;
#if defined(__CUDACC__)
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
#else   // host build of the kernels (tests/host_stub/simt): the same relaxed load
    v = __atomic_load_n(p, __ATOMIC_RELAXED);
#endif
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
#if defined(__CUDACC__)
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
    __atomic_store_n(p, v, __ATOMIC_RELEASE);
#endif
}

// Called by all 32 lanes of ONE warp.  Publishes this tile's aggregate, walks back over the
// predecessors' status words and returns the exclusive prefix of `tile` (same value in all lanes).
// Tiles are handed out by an atomic ticket, so every predecessor is running or done.
__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long *status, uint32_t tile,
                                                                 unsigned long long aggregate, uint32_t epoch,
                                                                 int lane) {
    if (tile == 0) {
        if (lane == 0) st_status(&status[0], pack_status(kFlagInclusive, epoch, aggregate));
        return 0ull;
    }
    if (lane == 0) st_status(&status[tile], pack_status(kFlagAggregate, epoch, aggregate));
    unsigned long long exclusive = 0;
    int64_t idx = (int64_t)tile - 1;
    while (true) {
        const int64_t my = idx - lane;
        uint32_t flag = 2;  // tiles before 0: inclusive prefix 0
        unsigned long long v = 0;
        if (my >= 0) {
            unsigned long long w;
            while (true) {
                w = ld_status(&status[my]);
                flag = status_flag(w, epoch);
                if (flag) break;
                __nanosleep(40);   // predecessor still working: do not burn issue slots
            }
            v = w & kValueMask;
        }
        const uint32_t incl = __ballot_sync(0xffffffffu, flag == 2);
        if (incl) {
            const int first = __ffs(incl) - 1;
            if (lane > first) v = 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        exclusive += v;
        if (incl) break;
        idx -= 32;
    }
    if (lane == 0) st_status(&status[tile], pack_status(kFlagInclusive, epoch, exclusive + aggregate));
    return exclusive;
}

// Wide-window look-back for heavy tiles, called by ALL threads of a 256-thread CTA (contains
// __syncthreads).  With hundreds of tiles in flight and tiles that take tens of microseconds, the
// nearest predecessor that already knows its inclusive prefix is hundreds of tiles back; a 32-wide
// window would need ~20 dependent L2 round trips to get there.  Here every thread inspects one
// predecessor per step: `nstreams` (1 or 2) independent prefixes are walked at once, each by
// 256/nstreams threads.  `scratch` needs 8 x 2 x 8 bytes of shared memory.
// aggregate[s] must be block-uniform.  Returns the exclusive prefix of stream s in excl[s].
//
// Status words of the wide look-back are padded to one 128-byte line each: hundreds of CTAs poll the
// same few hundred words at the same time, and 16 words per line would funnel all of that into a
// handful of L2 slices.
constexpr int kWideStatusStride = 16;   // in 8-byte words
//
// The two halves can be called apart: block_publish makes the tile's aggregate visible (successors
// can already add it), block_walk -- possibly much later, after other work -- finds the exclusive
// prefix and upgrades the status to "inclusive".  The later the walk, the less it waits.
__device__ __forceinline__ void block_publish(unsigned long long *const status[2], uint32_t tile,
                                              const unsigned long long aggregate[2], int nstreams, uint32_t epoch,
                                              int tid) {
    if (tid < nstreams)
        st_status(&status[tid][(size_t)tile * kWideStatusStride],
                  pack_status(tile == 0 ? kFlagInclusive : kFlagAggregate, epoch, aggregate[tid]));
}

// Barrier over a subset of the CTA's warps (named barrier `id` > 0, `nthreads` a multiple of 32);
// id 0 with all threads is __syncthreads().  Only ids 0 and 1 are used.
__device__ __forceinline__ void group_sync(int id, int nthreads) {
    // literal barrier numbers: with a register operand ptxas reserves all 16 barriers for the CTA
#if defined(__CUDACC__)
    if (id == 0) asm volatile("bar.sync 0, %0;" ::"r"(nthreads) : "memory");
    else asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
#else   // host build of the kernels: the test harness supplies the barrier (tests/host_stub/simt/simt_host.h)
    SK_HOST_BAR_SYNC(id, nthreads);
#endif
}

// Called by a group of `gwarps` (nstreams .. 8) whole warps, `gtid` = thread index
// inside the group, `bar` = the group's barrier (0: the whole 256-thread CTA).
// (De-inlining this and flush_previous_tile to shrink the kernel -- instruction fetch shows up as a
// stall -- was measured: 0.356 -> 0.391 ms; the by-pointer arguments go through local memory.)
__device__ __forceinline__ void block_walk(unsigned long long *const status[2], uint32_t tile,
                                           const unsigned long long aggregate[2], int nstreams, uint32_t epoch,
                                           int gtid, unsigned long long (*scratch)[2], unsigned long long excl[2],
                                           int gwarps = 8, int bar = 0) {
    const int lane = gtid & 31, wid = gtid >> 5;
    const int wps = gwarps / nstreams;            // warps per stream
    const int sid = wid / wps;                    // the stream this warp works on
    const int width = wps * 32;
    const int lt = gtid - sid * width;            // distance slot inside the stream's window
    unsigned long long *const my_status = sid ? status[1] : status[0];
    uint32_t ex0 = 0, ex1 = 0;                    // values of one batch fit 32 bits (see pack_status)
    excl[0] = excl[1] = 0;
    if (tile == 0) return;                        // published as inclusive already
    bool done0 = false, done1 = nstreams < 2;
    int64_t idx = (int64_t)tile - 1;
    uint2 *const sc = reinterpret_cast<uint2 *>(scratch);   // one (sum, has-inclusive) pair per warp
    uint2 *const res = sc + 8;                               // combined (sum, done) per stream
#ifdef SK_PHASE_TIMING
    bool spun = false;
    if (gtid == 0) atomicAdd(&g_walk_dbg[0], 1ull);
#endif
    while (!(done0 && done1)) {
        // ---- every thread: one predecessor of its stream (or nothing if that stream is finished)
        uint32_t flag = 2, v = 0;
        const int64_t my = idx - lt;
        if (sid < nstreams && !(sid ? done1 : done0) && my >= 0) {   // (a third warp of a 2-stream group idles)
            unsigned long long w;
            while (true) {
                w = ld_status(&my_status[(size_t)my * kWideStatusStride]);
                flag = status_flag(w, epoch);
                if (flag) break;
#ifdef SK_PHASE_TIMING
                atomicAdd(&g_walk_dbg[3], 1ull);
                if (!spun) { spun = true; atomicAdd(&g_walk_dbg[2], 1ull); }
#endif
                __nanosleep(40);
            }
            v = (uint32_t)w;
        }
#ifdef SK_PHASE_TIMING
        if (gtid == 0) atomicAdd(&g_walk_dbg[1], 1ull);
#endif
        const uint32_t incl = __ballot_sync(0xffffffffu, flag == 2);
        if (incl && lane > __ffs(incl) - 1) v = 0;            // nothing beyond the first inclusive prefix
        v = __reduce_add_sync(0xffffffffu, v);
        if (lane == 0) sc[wid] = make_uint2(v, incl ? 1u : 0u);
        group_sync(bar, gwarps * 32);
        // ---- warp 0 of the group combines the warps of each stream in distance order, up to the first
        // one that met an inclusive prefix; everybody else only reads the two results after the barrier
        if (wid == 0) {
            const uint2 e = lane < gwarps ? sc[lane] : make_uint2(0u, 0u);
            const uint32_t has = __ballot_sync(0xffffffffu, e.y != 0);
            const uint32_t m0 = (1u << wps) - 1u;                       // lanes of stream 0; stream 1 follows
            const uint32_t h0 = has & m0, h1 = (has >> wps) & m0;
            const uint32_t take0 = h0 ? ((2u << (__ffs(h0) - 1)) - 1u) : m0;
            const uint32_t take1 = h1 ? ((2u << (__ffs(h1) - 1)) - 1u) : m0;
            const uint32_t s0 = __reduce_add_sync(0xffffffffu, (lane < wps && ((take0 >> lane) & 1u) && !done0) ? e.x : 0u);
            const uint32_t s1 = __reduce_add_sync(0xffffffffu, (nstreams > 1 && lane >= wps && lane < 2 * wps &&
                                                                ((take1 >> (lane - wps)) & 1u) && !done1) ? e.x : 0u);
            if (lane == 0) {
                res[0] = make_uint2(s0, h0 ? 1u : 0u);
                res[1] = make_uint2(s1, h1 ? 1u : 0u);
            }
        }
        group_sync(bar, gwarps * 32);
        if (!done0) { const uint2 r0 = res[0]; ex0 += r0.x; done0 = r0.y != 0; }
        if (!done1) { const uint2 r1 = res[1]; ex1 += r1.x; done1 = r1.y != 0; }
        idx -= width;
    }
    excl[0] = ex0;
    excl[1] = ex1;
    if (gtid < nstreams)
        st_status(&status[gtid][(size_t)tile * kWideStatusStride],
                  pack_status(kFlagInclusive, epoch, (gtid ? ex1 : ex0) + aggregate[gtid]));
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
__device__ __forceinline__ int warp_incl_scan_i(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Byte offsets of a record's four lines inside its input buffer.
struct RecLines {
    uint32_t start[4];
    uint32_t len[4];
};
__device__ __forceinline__ RecLines record_lines(const DevInput &in, uint32_t rec) {
    RecLines r;
    const uint32_t *__restrict__ line_end = in.line_end;
    const uint4 e = *reinterpret_cast<const uint4 *>(line_end + 4ull * rec);
    const uint32_t prev = rec ? __ldg(line_end + 4ull * rec - 1) + 1u : in.first;
    r.start[0] = prev;      r.len[0] = e.x - prev;
    r.start[1] = e.x + 1u;  r.len[1] = e.y - e.x - 1u;
    r.start[2] = e.y + 1u;  r.len[2] = e.z - e.y - 1u;
    r.start[3] = e.z + 1u;  r.len[3] = e.w - e.z - 1u;
    return r;
}

}  // namespace sk
