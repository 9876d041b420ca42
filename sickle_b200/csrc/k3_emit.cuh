// k3_emit.cuh -- K3: stream compaction + FASTQ formatting, and the batch summary.
//
// Replaces the stringstream formatting of Trim_Single::output_single (src/trim_single.cpp:393-396)
// and Trim_Paired::get_read_string / output_paired (src/trim_paired.cpp:506-513,543-567): every kept
// record is written as  name '\n' seq[five:three] '\n' line3 '\n' qual[five:three] '\n'  (line 3 is
// echoed verbatim, as the reference does) at the byte offset K2's scan assigned to it, directly in
// the output stream's buffer.  -M "N records" follow README.md:116-120 (parity unpinned).
#pragma once

#include "k2_trim.cuh"
#include "sk_device.cuh"

namespace sk {

constexpr int kK3Threads = 256;

// Warp-cooperative copy of n bytes, arbitrary alignment on both sides.  Destination words are
// written 4-byte aligned; source words are fetched aligned and funnel-shifted.  May read up to 3
// bytes past src+n (inside the 16-byte padding every input buffer carries).
__device__ __forceinline__ void warp_copy(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, uint32_t n,
                                          int lane) {
    uint32_t head = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u;
    if (head > n) head = n;
    if ((uint32_t)lane < head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    const uint32_t nw = n >> 2;
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
    const uint32_t *__restrict__ s32 = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
    uint32_t *__restrict__ d32 = reinterpret_cast<uint32_t *>(dst);
    for (uint32_t k = lane; k < nw; k += 32) {
        const uint32_t lo = s32[k];
        const uint32_t hi = sh ? s32[k + 1] : 0u;
        d32[k] = __funnelshift_r(lo, hi, sh);
    }
    const uint32_t tail = n & 3u;
    if ((uint32_t)lane < tail) dst[4 * nw + lane] = src[4 * nw + lane];
}

// Whole-warp copy for long runs, 16 bytes per lane and two chunks in flight per lane (1 KB per round): destination
// chunks are written 16-byte aligned, source words fetched aligned and funnel-shifted.  May read up to 7 bytes past src+n.
__device__ __forceinline__ void warp_copy16(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, uint32_t n, int lane) {
    uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
    if (head > n) head = n;
    if ((uint32_t)lane < head) dst[lane] = src[lane];
    dst += head; src += head; n -= head;
    const uint32_t nq = n >> 4;
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
    const uint32_t *__restrict__ s32 = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
    uint4 *__restrict__ d128 = reinterpret_cast<uint4 *>(dst);
    for (uint32_t k0 = 0; k0 < nq; k0 += 64) {
        uint32_t w[2][5];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint32_t k = k0 + 32u * u + (uint32_t)lane;
#pragma unroll
            for (int i = 0; i < 5; ++i) w[u][i] = (k < nq && (i < 4 || sh)) ? s32[4 * k + i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint32_t k = k0 + 32u * u + (uint32_t)lane;
            if (k < nq)
                d128[k] = make_uint4(__funnelshift_r(w[u][0], w[u][1], sh), __funnelshift_r(w[u][1], w[u][2], sh),
                                     __funnelshift_r(w[u][2], w[u][3], sh), __funnelshift_r(w[u][3], w[u][4], sh));
        }
    }
    const uint32_t tail = n & 15u;
    if ((uint32_t)lane < tail) dst[16 * nq + lane] = src[16 * nq + lane];
}

// One kept record by a quarter-warp (8 lanes, `gl` = lane inside the group); the four quarters of a warp emit four
// records side by side.  The record is up to four source runs plus its final '\n' (run starts in output bytes: 0, b2,
// b3, b4, b5).  The work is cut by 16-byte aligned DESTINATION chunks: a chunk that lies inside one run is five aligned
// source words funnel-shifted into one 16-byte store; the chunks that hold a run boundary or one of the record's ragged
// ends (at most five: lanes 0..4 take one each) are gathered byte by byte.  Nothing here depends on an earlier store, so
// all the loads of a round (3 chunks per lane + the boundary chunk) are in flight together: the copy it replaces walked
// head / body / tail of run after run, eight dependent DRAM round trips per record, and was bound by exactly that
// (long-scoreboard stalls, 1.8 TB/s).  May read up to 3 bytes before and 7 bytes past a run (inside the buffer's padding).
struct EmitRuns {
    uint32_t d1, d2, d3, d4;   // source offset minus output offset, per run (d1 = source offset of run 1)
    uint32_t b2, b3, b4, b5;   // output offsets at which runs 2, 3, 4 and the final newline start
};

// source offset of output byte o (selects, no branches: the arms are registers)
__device__ __forceinline__ uint32_t emit_source(const EmitRuns &r, uint32_t o) {
    uint32_t d = r.d1;
    d = o >= r.b2 ? r.d2 : d;
    d = o >= r.b3 ? r.d3 : d;
    d = o >= r.b4 ? r.d4 : d;
    return o + d;
}

__device__ __forceinline__ void quarter_emit(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, const EmitRuns r,
                                             bool have, int gl) {
    const uint32_t total = have ? r.b5 + 1u : 0u;
    const uint32_t ph = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);
    uint8_t *__restrict__ gal = dst - ph;                     // 16-byte aligned
    const uint32_t end = ph + total;                          // in bytes from gal
    const uint32_t nch = (end + 15u) >> 4;
    // ---- boundary chunk of this lane: loads
    const uint32_t bnd = gl == 0 ? 0u : gl == 1 ? r.b2 : gl == 2 ? r.b3 : gl == 3 ? r.b4 : r.b5;
    const bool special = have && gl < 5;
    const uint32_t sc = (ph + bnd) >> 4;
    uint32_t sb[4] = {0u, 0u, 0u, 0u};
    uint32_t svalid = 0;
    if (special) {
        // sixteen unconditional byte loads (positions outside the record are clamped onto its last source byte), so that
        // they are all in flight together: a load inside an `if` is followed by its use, and an in-order warp then
        // pays one trip to memory per byte
        uint32_t vb[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t o = 16u * sc + (uint32_t)i - ph;   // wraps below the record's first byte: clamped as well
            vb[i] = (uint32_t)src[emit_source(r, min(o, r.b5 - 1u))];
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t pos = 16u * sc + (uint32_t)i;      // from gal
            const uint32_t v = pos - ph == r.b5 ? (uint32_t)'\n' : vb[i];
            sb[i >> 2] |= v << (8 * (i & 3));
            svalid |= (pos >= ph && pos < end) ? 1u << i : 0u;
        }
    }
    for (uint32_t c0 = 0; c0 < nch; c0 += 24u) {
        // ---- chunks inside one run: loads of three chunks per lane, then their stores
        uint32_t w[3][5];
        uint32_t sh[3];
        bool simple[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const uint32_t c = c0 + 8u * (uint32_t)u + (uint32_t)gl;
            const uint32_t o = 16u * c - ph;                  // meaningful when 16 c >= ph
            uint32_t lim = r.b2;                              // end of the run that holds o
            lim = o >= r.b2 ? r.b3 : lim;
            lim = o >= r.b3 ? r.b4 : lim;
            lim = o >= r.b4 ? r.b5 : lim;
            simple[u] = c < nch && 16u * c >= ph && o + 16u <= lim;
            const uint32_t a = emit_source(r, o);
            sh[u] = (a & 3u) * 8u;
            const uint32_t *__restrict__ s32 = reinterpret_cast<const uint32_t *>(src + (a & ~3u));
#pragma unroll
            for (int i = 0; i < 5; ++i) w[u][i] = (simple[u] && (i < 4 || sh[u])) ? s32[i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const uint32_t c = c0 + 8u * (uint32_t)u + (uint32_t)gl;
            if (simple[u])
                *reinterpret_cast<uint4 *>(gal + 16u * c) =
                    make_uint4(__funnelshift_r(w[u][0], w[u][1], sh[u]), __funnelshift_r(w[u][1], w[u][2], sh[u]),
                               __funnelshift_r(w[u][2], w[u][3], sh[u]), __funnelshift_r(w[u][3], w[u][4], sh[u]));
        }
    }
    // ---- boundary chunk: stores (a chunk shared with the neighbouring record is written byte by byte)
    if (special) {
        if (svalid == 0xffffu) {
            *reinterpret_cast<uint4 *>(gal + 16u * sc) = make_uint4(sb[0], sb[1], sb[2], sb[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (svalid & (1u << i)) gal[16u * sc + (uint32_t)i] = (uint8_t)(sb[i >> 2] >> (8 * (i & 3)));
        }
    }
}

__device__ __forceinline__ void emit_record(const DevInput &in, uint32_t rec, const RecDesc d, uint8_t *const *outs,
                                            const DevParams &P, int lane) {
    const RecLines r = record_lines(in, rec);
    const uint8_t *__restrict__ src = in.data;
    uint8_t *dst = outs[d.route & 3u] + d.dst_off;
    warp_copy(dst, src + r.start[0], r.len[0] + 1u, lane);           // name line incl. its '\n'
    dst += r.len[0] + 1u;
    if (d.route & kRouteNRec) {
        if (lane == 0) { dst[0] = 'N'; dst[1] = '\n'; }
        dst += 2;
        warp_copy(dst, src + r.start[2], r.len[2] + 1u, lane);       // line 3 verbatim incl. '\n'
        dst += r.len[2] + 1u;
        if (lane == 0) { dst[0] = (uint8_t)P.qmin; dst[1] = '\n'; }
        return;
    }
    warp_copy16(dst, src + r.start[1] + d.five, d.nkeep, lane);      // seq[five:three]
    dst += d.nkeep;
    if (lane == 0) *dst = '\n';
    dst += 1;
    warp_copy(dst, src + r.start[2], r.len[2] + 1u, lane);           // line 3 verbatim incl. '\n'
    dst += r.len[2] + 1u;
    warp_copy16(dst, src + r.start[3] + d.five, d.nkeep, lane);      // qual[five:three]
    dst += d.nkeep;
    if (lane == 0) *dst = '\n';
}

struct OutPtrs {
    uint8_t *p[kMaxStreams];
    unsigned long long cap[kMaxStreams];
};

// K3 for batches of long records (what capi.cu launches while the previous general-path batch averaged 1.5 KB or more per
// record): one record per warp, the whole warp on each of its runs.  The same loop is the `long_batch` branch of
// k3_emit; as a kernel of its own it keeps the registers (and CTAs per SM) it had before k3_emit's quarter-warp path
// grew (same-GPU: 0.074 -> 0.06x ms per 255 MB of 1-20 kb reads).  Correct for any batch, merely slow on short records.
__global__ void __launch_bounds__(256, 4)
k3_emit_long(DevInput in0, DevInput in1, DevParams P, const Control *__restrict__ ctl, const RecDesc *__restrict__ desc0,
             const RecDesc *__restrict__ desc1, OutPtrs outs) {
    const int lane = threadIdx.x & 31;
    const Geometry g = batch_geometry(ctl, P);
    if (ctl->err_key != kNoError || ctl->fast_fail) return;
#pragma unroll
    for (int s = 0; s < kMaxStreams; ++s)
        if (ctl->out_bytes[s] > outs.cap[s]) return;
    const uint32_t nw = gridDim.x * (256 / 32);
    for (uint32_t c = blockIdx.x * (256 / 32) + (threadIdx.x >> 5); c < g.nrec0 + g.nrec1; c += nw) {
        const bool second = c >= g.nrec0;
        const uint32_t rec = second ? c - g.nrec0 : c;
        const RecDesc d = (second ? desc1 : desc0)[rec];
        if (d.route & kRouteEmit) emit_record(second ? in1 : in0, rec, d, outs.p, P, lane);
    }
}

#ifndef SK_K3_MINCTAS
#define SK_K3_MINCTAS 3
#endif
__global__ void __launch_bounds__(kK3Threads, SK_K3_MINCTAS)
k3_emit(DevInput in0, DevInput in1, DevParams P, const Control *__restrict__ ctl, const RecDesc *__restrict__ desc0,
        const RecDesc *__restrict__ desc1, OutPtrs outs) {
    const int lane = threadIdx.x & 31;
    const Geometry g = batch_geometry(ctl, P);
    if (ctl->err_key != kNoError || ctl->fast_fail) return;  // outputs of a failing batch are never used (fast_fail: the index pass gave up)
#pragma unroll
    for (int s = 0; s < kMaxStreams; ++s)
        if (ctl->out_bytes[s] > outs.cap[s]) return;  // reported by finalize as a capacity error
    // A warp takes 32 consecutive records at a time.  First every lane fetches the descriptor and the
    // line ends of ITS record (coalesced, and the three dependent loads happen once per 32 records
    // instead of once per record); then the warp copies the records one after the other, the copy
    // parameters coming from the owning lane by shuffle.  Records cut at the 5' end and -M "N records"
    // (rare) take the four-piece path.
    // Batches of long records (>= 1.5 KB on average, K2's rule) hand a warp one record at a time instead of
    // 32: 20,000 reads of 1-20 kb would otherwise keep 625 of the grid's ~9,500 warps busy, each copying
    // ~125 KB, while the others exit at once.
    const uint32_t nwarps = gridDim.x * (kK3Threads / 32);
    const uint32_t nrec_all = g.nrec0 + g.nrec1;
    const bool long_batch = nrec_all > 0 && (in0.nbytes + in1.nbytes) / nrec_all >= kK2LongRecordBytes;
    if (long_batch) {   // one record per warp, the whole warp on each of its runs
        const uint32_t nw = gridDim.x * (kK3Threads / 32);
        for (uint32_t c = blockIdx.x * (kK3Threads / 32) + (threadIdx.x >> 5); c < g.nrec0 + g.nrec1; c += nw) {
            const bool second = c >= g.nrec0;
            const uint32_t rec = second ? c - g.nrec0 : c;
            const RecDesc d = (second ? desc1 : desc0)[rec];
            if (d.route & kRouteEmit) emit_record(second ? in1 : in0, rec, d, outs.p, P, lane);
        }
        return;
    }
    const uint32_t rpc = 32u;   // records per warp and round
    const uint32_t chunks0 = (g.nrec0 + rpc - 1u) / rpc, chunks1 = (g.nrec1 + rpc - 1u) / rpc;
    for (uint32_t c = blockIdx.x * (kK3Threads / 32) + (threadIdx.x >> 5); c < chunks0 + chunks1; c += nwarps) {
        const bool second = c >= chunks0;
        const DevInput &in = second ? in1 : in0;
        const uint32_t nrec = second ? g.nrec1 : g.nrec0;
        const uint32_t rec = (second ? c - chunks0 : c) * rpc + (uint32_t)lane;
        RecDesc d;
        d.route = 0; d.dst_off = 0; d.five = 0; d.nkeep = 0;
        if ((uint32_t)lane < rpc && rec < nrec) d = (second ? desc1 : desc0)[rec];
        const bool emit = (d.route & kRouteEmit) != 0;
        const bool slow = emit && (d.route & kRouteNRec);              // -M "N records": four-piece path
        // Runs of a kept record, in output order: name '\n' | seq[five:three] | '\n' line3 '\n' | qual[five:three]
        // | '\n'.  With nothing cut at the 5' end (most reads) the bases follow the name line in the input
        // and the qualities follow line 3, so runs 1+2 and 3+4 are one run each.
        uint32_t s1 = 0, l1 = 0, s2 = 0, l2 = 0, s3 = 0, l3 = 0, s4 = 0, l4 = 0;
        if (emit && !slow) {
            const RecLines r = record_lines(in, rec);
            const bool cut5 = d.five != 0;
            s1 = r.start[0]; l1 = r.len[0] + 1u + (cut5 ? 0u : d.nkeep);
            s2 = r.start[1] + d.five; l2 = cut5 ? d.nkeep : 0u;
            s3 = r.start[2] - 1u; l3 = r.len[2] + 2u + (cut5 ? 0u : d.nkeep);
            s4 = r.start[3] + d.five; l4 = cut5 ? d.nkeep : 0u;
        }
        // the four quarters of the warp copy four records at a time (a quarter without a record runs
        // the same code with empty runs: the shuffles and votes below are warp-wide)
        const int quarter = lane >> 3, gl = lane & 7;
        for (uint32_t todo = __ballot_sync(0xffffffffu, emit && !slow); todo;) {
            int k = -1;
            const int k0 = __ffs(todo) - 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (todo) {
                    if (q == quarter) k = __ffs(todo) - 1;
                    todo &= todo - 1;
                }
            }
            const int from = k < 0 ? k0 : k;
            const bool have = k >= 0;
            const uint32_t route = __shfl_sync(0xffffffffu, d.route, from);
            const uint32_t off = __shfl_sync(0xffffffffu, d.dst_off, from);
            const uint32_t a1 = __shfl_sync(0xffffffffu, s1, from), a3 = __shfl_sync(0xffffffffu, s3, from);
            uint32_t n1 = __shfl_sync(0xffffffffu, l1, from), n3 = __shfl_sync(0xffffffffu, l3, from);
            uint32_t n2 = __shfl_sync(0xffffffffu, l2, from), n4 = __shfl_sync(0xffffffffu, l4, from);
            if (!have) { n1 = 0; n2 = 0; n3 = 0; n4 = 0; }
            uint32_t a2 = 0, a4 = 0;
            if (__any_sync(0xffffffffu, (n2 | n4) != 0)) {            // some read of the four is cut at its 5' end
                a2 = __shfl_sync(0xffffffffu, s2, from);
                a4 = __shfl_sync(0xffffffffu, s4, from);
            }
            EmitRuns er;
            er.b2 = n1; er.b3 = n1 + n2; er.b4 = n1 + n2 + n3; er.b5 = n1 + n2 + n3 + n4;
            er.d1 = a1; er.d2 = a2 - er.b2; er.d3 = a3 - er.b3; er.d4 = a4 - er.b4;
            quarter_emit(outs.p[route & 3u] + off, in.data, er, have, gl);
        }
        for (uint32_t todo = __ballot_sync(0xffffffffu, slow); todo; todo &= todo - 1) {
            const int k = __ffs(todo) - 1;
            RecDesc dk;
            dk.route = __shfl_sync(0xffffffffu, d.route, k);
            dk.dst_off = __shfl_sync(0xffffffffu, d.dst_off, k);
            dk.five = __shfl_sync(0xffffffffu, d.five, k);
            dk.nkeep = __shfl_sync(0xffffffffu, d.nkeep, k);
            emit_record(in, rec - (uint32_t)lane + (uint32_t)k, dk, outs.p, P, lane);
        }
    }
}

// One thread: batch summary (sizes, consumed bytes, counters, first data error); resets the control block.
__global__ void k_finalize(DevInput in0, DevInput in1, DevParams P, Control *__restrict__ ctl, OutPtrs outs,
                           DevResult *__restrict__ res) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const Geometry g = batch_geometry(ctl, P);
    DevResult r;
    for (int s = 0; s < kMaxStreams; ++s) r.out_bytes[s] = ctl->out_bytes[s];
    r.records[0] = g.nrec0;
    r.records[1] = g.nrec1;
    r.consumed[0] = g.nrec0 ? (unsigned long long)in0.line_end[4ull * g.nrec0 - 1] + 1ull : 0ull;
    r.consumed[1] = g.nrec1 ? (unsigned long long)in1.line_end[4ull * g.nrec1 - 1] + 1ull : 0ull;
    for (int k = 0; k < 8; ++k) r.counters[k] = (long long)ctl->counters[k];
    // (fast_fail: the single-pass kernel's index + verdict pass in front of K2/K3 gave the batch up -- bits 2, 3 as in kf_finalize)
    r.index_overflow = ctl->index_overflow | (ctl->fast_fail ? 4u : 0u) | ((ctl->fast_fail & 2u) ? 8u : 0u);
    for (int s = 0; s < kMaxStreams; ++s)
        if (ctl->out_bytes[s] > outs.cap[s]) r.index_overflow |= 2u;
    r.pad = 0;
    r.err_kind = 0; r.err_file = 0; r.err_record = 0; r.err_position = 0; r.err_byte = 0;
    for (int k = 0; k < 4; ++k) { r.err_line_off[k] = 0; r.err_line_len[k] = 0; }
    const unsigned long long key = ctl->err_key;
    if (key != kNoError) {
        const bool quality = (key >> 63) != 0;
        const uint32_t unit = (uint32_t)((key >> 32) & 0x7fffffffu);
        const int mate = (int)((key >> 31) & 1u);
        const uint32_t pos = (uint32_t)(key & 0x7fffffffu);
        const bool inter = P.mode >= 2;
        const bool second = (P.mode == 1 && mate == 1);
        const DevInput &in = second ? in1 : in0;
        const uint32_t rec = inter ? 2 * unit + mate : unit;
        const RecLines l = record_lines(in, rec);
        r.err_file = second ? 1 : 0;
        r.err_record = rec;
        for (int k = 0; k < 4; ++k) { r.err_line_off[k] = l.start[k]; r.err_line_len[k] = l.len[k]; }
        if (quality) {
            r.err_kind = 6;
            r.err_position = (int)pos;
            r.err_byte = (int)(signed char)in.data[l.start[3] + pos];
        } else {
            r.err_kind = validate_record(in.data, l);
        }
    }
    *res = r;
    // leave the control block clean for the slot's next batch (no per-batch memset on the stream)
    Control z;
    memset(&z, 0, sizeof z);
    z.err_key = kNoError;
    *ctl = z;
}

}  // namespace sk
