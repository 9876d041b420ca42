// k2_trim.cuh -- K2: per-read sliding-window trimming + keep/discard/singles routing + output scan.
//
// Replaces Abstract_Trimmer::sliding_window / get_quality_num (src/trim.cpp:3-140), the record
// checks of FQEntry::validate (src/FQEntry.cpp:53-97), the keep flags of processing_thread
// (src/trim_single.cpp:357-372, src/trim_paired.cpp:483-504) and the routing decisions of
// output_single / output_paired (src/trim_single.cpp:382-404, src/trim_paired.cpp:531-567).
//
// Arithmetic is integer only.  The reference compares (double)total/(double)ws with the threshold;
// that is exactly  total >= qthr*ws  (IEEE division is correctly rounded and monotone; SURVEY.md
// section 8-a2), and (int)(0.1*L) == L/10 for every L < 5e6.
//
// This is the general ("any read length") path.  A warp handles 32 consecutive units (reads, or
// pairs) and lane k ends up holding unit k's result; routing, the 3-stream output-length scan (warp
// scan + decoupled look-back) and the per-record descriptors are then lane-parallel.  Every read is
// trimmed by its own lane (lane_sliding_window, shared with the single-pass kernel: a word-granular
// pass over the quality line, then the exact scan only where a window can be bad -- about 2.5
// instructions per base whatever the read length, 1 kb or 20 kb alike); whatever that declines --
// malformed records, quality bytes out of range, the last record before the end of the buffer -- is
// walked by the whole warp, 32 window positions per step, window totals from a warp-shuffle prefix
// scan of q[i+ws]-q[i]; that path also finds the position the reference reports for a bad quality byte.
//
// Tile = 256 units (8 warps x 32) for short records.  When the batch averages 1.5 KB or more per
// record a tile is 16 units, 2 per warp (lanes 0 and 1): a 255 MB batch of 1-20 kb reads is 20,000 units,
// i.e. 78 tiles of 256 -- fewer tiles than SMs, each warp walking up to 32 long reads one after the
// other, every step of the walk a trip to memory -- but 1,250 tiles of 16, 10,000 warps with at most two
// long reads each.  Reads above
// kThreadTrimMaxLen bases go to the warp-wide path, whose coarse pass is warp-cooperative (a single
// lane walking a 20 kb read alone is a 0.3 ms dependent chain).
#pragma once

#include "sk_device.cuh"
#include "trim_lane.cuh"

namespace sk {

constexpr int kK2Threads = 256;
constexpr int kK2UnitsPerTile = kK2Threads;  // 8 warps x 32 units (short records)
constexpr uint32_t kK2LongUnitsPerWarp = 2;   // batches of long records: 2 units per warp, 16 per tile
constexpr uint32_t kK2aLongUnitsPerTicket = 2; // k2_trim_only, long records: units per ticket (1 and 256-step rounds: no change, profiles/r2_call21.log)
constexpr int kCoarseU = 4, kNScanU = 4;       // warp_sliding_window: word loads per lane in flight per round (coarse pass: x2)
constexpr uint32_t kK2LongUnitsPerTile = kK2LongUnitsPerWarp * (kK2Threads / 32);
constexpr uint32_t kK2LongRecordBytes = 1500; // average record size from which a batch counts as "long"

struct Cut {
    int five, three;  // three < 0 => discard (src/trim_single.cpp:368)
};

// record / unit counts of the batch, derived from the line counts K1 produced
struct Geometry {
    uint32_t nrec0, nrec1, nunits;
};
__device__ __forceinline__ Geometry batch_geometry(const Control *ctl, const DevParams &P) {
    Geometry g;
    const bool ovf = ctl->index_overflow != 0;
    g.nrec0 = ovf ? 0u : ctl->nlines[0] >> 2;
    g.nrec1 = ovf ? 0u : ctl->nlines[1] >> 2;
    if (P.mode == 0) { g.nunits = g.nrec0; g.nrec1 = 0; }
    else if (P.mode == 1) { g.nunits = min(g.nrec0, g.nrec1); g.nrec0 = g.nrec1 = g.nunits; }
    else { g.nunits = g.nrec0 >> 1; g.nrec0 = g.nunits * 2; g.nrec1 = 0; }
    return g;
}

// Output position p (0-based, in the reference's write order) -> unit index.
// Reference: records are dealt round-robin to N queues and written queue by queue;
// se: record k -> queue (k+1)%N (src/trim_single.cpp:263,273-274); pe: pair k -> queue k%N
// (src/trim_paired.cpp:349,388,403).
__device__ __forceinline__ uint32_t position_to_unit(uint32_t p, uint32_t U, int N, bool paired) {
    if (N <= 1) return p;
    const uint32_t n = (uint32_t)N, a = U / n, b = U % n;
    if (!paired) {
        // queue 0 holds residue N-1 (a elements, since b <= N-1), queues 1.. hold residues 0..N-2
        if (p < a) return p * n + (n - 1);
        p -= a;
    }
    if (p < b * (a + 1)) return (p % (a + 1)) * n + p / (a + 1);
    p -= b * (a + 1);
    return (p % a) * n + b + p / a;
}

// FQEntry::validate, src/FQEntry.cpp:53-97.  Line 3 is not inspected by the reference.
__device__ __forceinline__ int validate_record(const uint8_t *__restrict__ d, const RecLines &r) {
    if (r.len[0] <= 1) return 1;
    if (d[r.start[0]] != '@') return 2;
    if (r.len[1] < 1) return 3;
    if (r.len[3] < 1) return 4;
    if (r.len[3] != r.len[1]) return 5;
    return 0;
}

// One warp, one read.  All lanes return the same Cut.  err_pos >= 0: a quality byte outside
// [qmin,qmax] was met at that position inside the prefix the reference's scalar loop visits.
__device__ __noinline__ Cut warp_sliding_window(const uint8_t *__restrict__ d, uint32_t seq_off, uint32_t L,
                                                uint32_t qual_off, const DevParams &P, int lane, int &err_pos) {
    const Cut discard = {-1, -1};
    err_pos = -1;
    if (L < (uint32_t)P.lthr) return discard;                    // trim.cpp:21-26 (nothing validated)
    const uint8_t *__restrict__ q = d + qual_off;
    uint32_t ws = L / 10u;                                       // trim.cpp:8
    if (ws == 0) ws = L;                                         // trim.cpp:30
    const long long thr_total = (long long)P.qthr * (long long)ws;
    const uint32_t nwin = L - ws + 1u;                           // trim.cpp:34

    // ---- coarse pass for long reads (the warp-wide twin of phase 1 of lane_sliding_window, same bound):
    // lane l takes word s0 + l of the quality line, 128 coalesced bytes per step; a warp scan of
    // (lead word sum - trail word sum) gives the k-word sliding sums.  Up to the first word whose bound
    // falls below the threshold every window is good, so the exact walk below starts at w_start instead of
    // window 0.  A quality byte outside [qmin, qmax] anywhere in the words looked at (it may lie beyond
    // what the reference visits) switches the shortcut off: the exact walk then runs from window 0 and
    // reports the error exactly as before.
    uint32_t w_start = 0;
    if (L >= 256u && thr_total <= 0x3fffffffLL && (L >> 24) == 0u) {
        RangeCheck rc;
        rc.init(P);
        const uint32_t *__restrict__ W = reinterpret_cast<const uint32_t *>(d);   // d is 16-byte aligned
        const uint32_t Q = qual_off, E = qual_off + L;
        const uint32_t A0 = (Q + 3u) >> 2, A1 = E >> 2;          // whole words [A0, A1)
        const uint32_t k = (ws - 3u) >> 2;                       // ws >= 25 here: k >= 5, and A1 - A0 >= k
        const uint32_t nfull = A1 - A0, nstep = nfull - k + 1u;
        const int slack = P.qmin < P.qoff ? (int)(ws - 4u * k) * (P.qmin - P.qoff) : 0;
        const int bias = (int)thr_total + (int)(4u * k) * P.qoff - slack;
        uint32_t scr = 0;
        if (lane == 0) {                                         // ragged head and tail of the line
            if (Q & 3u) { const uint32_t m = 0xffffffffu << (8u * (Q & 3u)); rc.screen3((W[A0 - 1u] & m) | (rc.kmin & ~m), scr); }
            if (E & 3u) { const uint32_t m = ~(0xffffffffu << (8u * (E & 3u))); rc.screen3((W[A1] & m) | (rc.kmin & ~m), scr); }
        }
        int part = 0;
        for (uint32_t j = (uint32_t)lane; j < k; j += 32) {
            const uint32_t v = W[A0 + j];
            rc.screen3(v, scr);
            part = (int)__dp4a(v, 0x01010101u, (uint32_t)part);
        }
        int R = warp_sum_i(part) - bias;                         // bound of the first whole word (step 0)
        uint32_t jf = nstep;                                     // first step whose bound is below the threshold
        // kCoarseU * 32 steps per round: the loads of a lane (kCoarseU lead and as many trail words, each
        // warp-coalesced) are issued together, so that a round costs one trip to memory rather than kCoarseU.
        // (256 steps per round and one read per ticket were measured against 128 and two: no change.)
        for (uint32_t s0 = 0; s0 < nstep && jf == nstep; s0 += 32u * kCoarseU) {
            uint32_t lw[kCoarseU], tw[kCoarseU];
#pragma unroll
            for (int u = 0; u < kCoarseU; ++u) {
                const uint32_t s = s0 + 32u * u + (uint32_t)lane;
                const bool in = s + 1u < nstep;
                lw[u] = in ? W[A0 + k + s] : rc.kmin;
                tw[u] = in ? W[A0 + s] : rc.kmin;
            }
            // what step s+1 has over step s (nothing past the last step: both words are the filler)
            int Dv[kCoarseU], neg = 0, tot = 0;
#pragma unroll
            for (int u = 0; u < kCoarseU; ++u) {
                rc.screen3(lw[u], scr);
                Dv[u] = dp4a_us(lw[u], 0x01010101, dp4a_us(tw[u], (int)0xFFFFFFFF, 0));
                neg += min(Dv[u], 0);
                tot += Dv[u];
            }
            // No prefix of the round's differences is below the sum of the negative ones: while the bound stays
            // that far above the threshold (all of a read's good stretch) a round is two warp reductions instead of
            // a scan, a ballot and a broadcast per 32 steps.
            if (R + (int)__reduce_add_sync(0xffffffffu, (uint32_t)neg) >= 0) {
                R += (int)__reduce_add_sync(0xffffffffu, (uint32_t)tot);
                continue;
            }
#pragma unroll
            for (int u = 0; u < kCoarseU; ++u) {
                const uint32_t s = s0 + 32u * u + (uint32_t)lane;
                const int D = Dv[u];
                const int incl = warp_incl_scan_i(D, lane);
                const uint32_t bm = __ballot_sync(0xffffffffu, s < nstep && R + incl - D < 0);
                if (bm) { jf = s0 + 32u * u + (uint32_t)__ffs(bm) - 1u; break; }
                R += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        if (!__any_sync(0xffffffffu, (scr & 0x80808080u) != 0)) {
            const int ws_i = jf == nstep ? (int)nwin : 4 * (int)(A0 + jf) - (int)Q - 3;
            w_start = ws_i < 0 ? 0u : min((uint32_t)ws_i, nwin);
        }
    }

    int part = 0;                                                // trim.cpp:31-33 (total of window w_start)
    if (w_start < nwin) {
        for (uint32_t j0 = 0; j0 < ws; j0 += 32) {
            const uint32_t j = j0 + lane;
            bool bad = false;
            if (j < ws) {
                const int b = q[w_start + j];
                bad = (b < P.qmin) | (b > P.qmax);
                part += b - P.qoff;
            }
            const uint32_t bm = __ballot_sync(0xffffffffu, bad);
            if (bm) { err_pos = (int)(w_start + j0 + __ffs(bm) - 1); return discard; }
        }
    }
    int carry = warp_sum_i(part);

    // w_start > 0: windows 0 .. w_start-1 are good, so window 0 is the first good one
    bool found = w_start > 0 && !P.no_fiveprime;
    int i5 = found ? 0 : -1, i3 = -1;
    for (uint32_t c = w_start; c < nwin; c += 32) {
        const uint32_t i = c + lane;
        const bool valid = i < nwin;
        int dlt = 0;
        bool bad = false;
        if (valid && i > w_start) {                              // trim.cpp:76-79 for window i-1
            const int lead = q[i - 1 + ws], trail = q[i - 1];
            bad = (lead < P.qmin) | (lead > P.qmax);
            dlt = lead - trail;
        }
        const int T = carry + warp_incl_scan_i(dlt, lane);
        carry = __shfl_sync(0xffffffffu, T, 31);
        const bool good = valid && (long long)T >= thr_total;    // trim.cpp:36,42,61
        const uint32_t vm = __ballot_sync(0xffffffffu, valid);
        const uint32_t gm = __ballot_sync(0xffffffffu, good);
        bool can3 = found || P.no_fiveprime;
        uint32_t from = 0;
        if (!can3 && gm) {                                       // first good window: trim.cpp:42
            from = __ffs(gm) - 1;
            i5 = (int)(c + from);
            found = true;
            can3 = true;
        }
        uint32_t brk = 32;
        if (can3) {                                              // first bad window after it: trim.cpp:61
            const uint32_t bm = vm & ~gm & (0xffffffffu << from);
            if (bm) { brk = __ffs(bm) - 1; i3 = (int)(c + brk); }
        }
        // a lead byte is range-checked iff its window is reached, i.e. lane <= brk
        const uint32_t reach = brk >= 31 ? 0xffffffffu : ((2u << brk) - 1u);
        const uint32_t badm = __ballot_sync(0xffffffffu, bad) & reach;
        if (badm) { err_pos = (int)(c + (__ffs(badm) - 1) - 1 + ws); return discard; }
        if (i3 >= 0) break;
    }

    int five = 0, three = (int)L;
    if (found) {                                                 // trim.cpp:46-51
        for (uint32_t j0 = (uint32_t)i5; j0 < (uint32_t)i5 + ws; j0 += 32) {
            const uint32_t j = j0 + lane;
            const bool hit = j < (uint32_t)i5 + ws && (int)q[j] - P.qoff >= P.qthr;
            const uint32_t hm = __ballot_sync(0xffffffffu, hit);
            if (hm) { five = (int)(j0 + __ffs(hm) - 1); break; }
        }
    }
    if (i3 >= 0) {                                               // trim.cpp:65-70
        for (uint32_t j0 = (uint32_t)i3; j0 < (uint32_t)i3 + ws; j0 += 32) {
            const uint32_t j = j0 + lane;
            const bool hit = j < (uint32_t)i3 + ws && (int)q[j] - P.qoff < P.qthr;
            const uint32_t hm = __ballot_sync(0xffffffffu, hit);
            if (hm) { three = (int)(j0 + __ffs(hm) - 1); break; }
        }
    }
    if (P.trunc_n) {                                             // trim.cpp:86-98 (bug kept: 'N' only => -2)
        // first lowercase n (it wins), else any uppercase N.  Aligned words, kNScanU per lane and round (1 KB of
        // coalesced bytes, the loads issued together): a byte-per-lane loop pays one trip to memory per 32 bases.
        const uint32_t *__restrict__ W = reinterpret_cast<const uint32_t *>(d);
        const uint32_t S0 = seq_off, SE = seq_off + L;
        const uint32_t wa = S0 >> 2, nwords = ((SE + 3u) >> 2) - wa;
        int pn = -1;
        bool anyN = false;
        // flag masks of the bytes that belong to the line, for its first and its last word
        const uint32_t head_ok = 0x80808080u << (8u * (S0 & 3u));
        const uint32_t tail_ok = (SE & 3u) ? 0x80808080u >> (8u * (4u - (SE & 3u))) : 0x80808080u;
        for (uint32_t w0 = 0; w0 < nwords && pn < 0; w0 += 32u * kNScanU) {
            uint32_t v[kNScanU];
#pragma unroll
            for (int u = 0; u < kNScanU; ++u) {
                const uint32_t idx = w0 + 32u * u + (uint32_t)lane;
                v[u] = idx < nwords ? W[wa + idx] : 0u;
            }
            // a round without any n / N (nearly all of them) is one vote
            uint32_t hn[kNScanU], hN[kNScanU], any = 0;
#pragma unroll
            for (int u = 0; u < kNScanU; ++u) {
                const uint32_t idx = w0 + 32u * u + (uint32_t)lane;
                uint32_t okf = 0x80808080u;
                okf = idx == 0u ? okf & head_ok : okf;
                okf = idx + 1u == nwords ? okf & tail_ok : okf;
                const uint32_t tn = ((v[u] ^ 0x6E6E6E6Eu) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
                const uint32_t tN = ((v[u] ^ 0x4E4E4E4Eu) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
                hn[u] = ~(tn | v[u]) & okf;                          // (a word past the line was loaded as 0: no match)
                hN[u] = ~(tN | v[u]) & okf;
                any |= hn[u] | hN[u];
            }
            if (!__any_sync(0xffffffffu, any != 0)) continue;
#pragma unroll
            for (int u = 0; u < kNScanU; ++u) {
                const uint32_t fn = flags_to_nibble(hn[u]);
                const uint32_t fN = flags_to_nibble(hN[u]);
                const uint32_t mn = __ballot_sync(0xffffffffu, fn != 0);
                anyN |= __ballot_sync(0xffffffffu, fN != 0) != 0;
                if (mn) {
                    const int first = __ffs(mn) - 1;
                    const uint32_t f1 = __shfl_sync(0xffffffffu, fn, first);
                    pn = (int)(4u * (wa + w0 + 32u * u + (uint32_t)first) + (uint32_t)__ffs(f1) - 1u - S0);
                    break;
                }
            }
        }
        if (pn >= 0) three = pn - 1;
        else if (anyN) three = -2;
    }
    if ((!found && !P.no_fiveprime) || (three - five < P.lthr)) return discard;  // trim.cpp:103
    return Cut{five, three};
}

__device__ __forceinline__ unsigned long long make_err_key(bool quality, uint32_t unit, int mate, uint32_t pos) {
    return ((unsigned long long)(quality ? 1u : 0u) << 63) | ((unsigned long long)(unit & 0x7fffffffu) << 32) |
           ((unsigned long long)(mate & 1) << 31) | (unsigned long long)(pos & 0x7fffffffu);
}

struct MateInfo {
    Cut cut;
    uint32_t fixed_len;  // name_len + plus_len + 4 newlines
};

// Trim one record with the whole warp; reports data errors through ctl->err_key.
__device__ __forceinline__ MateInfo trim_mate(const DevInput &in, uint32_t rec, uint32_t unit, int mate,
                                              const DevParams &P, Control *ctl, int lane) {
    const RecLines r = record_lines(in, rec);
    MateInfo m;
    m.fixed_len = r.len[0] + r.len[2] + 4u;
    const int bad = validate_record(in.data, r);
    if (bad) {
        if (lane == 0) atomicMin(&ctl->err_key, make_err_key(false, unit, mate, 0));
        m.cut = Cut{-1, -1};
        return m;
    }
    int err_pos;
    m.cut = warp_sliding_window(in.data, r.start[1], r.len[1], r.start[3], P, lane, err_pos);
    if (err_pos >= 0 && lane == 0) atomicMin(&ctl->err_key, make_err_key(true, unit, mate, (uint32_t)err_pos));
    return m;
}

// Short reads: one lane trims the record by itself, straight from global memory (the 32 lanes of a
// warp work on 32 different records; each lane walks its own cache lines, which stay in L1/L2 for the
// few hundred bytes of a read).  Far fewer warp instructions per read than the warp-wide scan above.
// Returns false when the warp-wide path has to take the record: malformed (it reports the error),
// a quality byte out of range (it finds the position), longer than kThreadTrimMaxLen (one lane walking
// a long read alone is a long dependent chain; the warp-wide path spreads it over 32 lanes), or so close
// to the end of the buffer that the lane's word-wise look-ahead (< 64 bytes) could leave it.
constexpr uint32_t kThreadTrimMaxLen = 1024;
__device__ __forceinline__ bool thread_trim_mate(const DevInput &in, uint32_t rec, const DevParams &P,
                                                 const RangeCheck &rc, int lane, MateInfo &m) {
    const RecLines r = record_lines(in, rec);
    m.fixed_len = r.len[0] + r.len[2] + 4u;
    m.cut = Cut{-1, -1};
    if (validate_record(in.data, r)) return false;
    if (r.len[1] > kThreadTrimMaxLen || (unsigned long long)r.start[3] + r.len[1] + 64ull > in.nbytes) return false;
    const TrimOut t = lane_sliding_window(in.data, r.start[1], r.len[1], r.start[3], P, rc);
    if (t.error) return false;
    m.cut = Cut{t.five, t.three};
    return true;
}

// Phase 1 of K2 for the units a warp holds (lane k < upw holds unit `my_unit`; has = it holds one): lane k trims
// unit k by itself (short reads); whatever that path declines is redone by the whole warp, one unit after
// the other.
__device__ __forceinline__ void trim_units(const DevInput &in0, const DevInput &in1, const DevParams &P, const RangeCheck &rc,
                                           Control *ctl, bool has, uint32_t my_unit, int lane, MateInfo &mine0, MateInfo &mine1) {
    const bool paired = P.mode != 0, inter = P.mode >= 2;
    bool redo = false;
    if (has) {
        redo = !thread_trim_mate(in0, inter ? 2 * my_unit : my_unit, P, rc, lane, mine0);
        if (paired && !redo)
            redo = !(inter ? thread_trim_mate(in0, 2 * my_unit + 1, P, rc, lane, mine1)
                           : thread_trim_mate(in1, my_unit, P, rc, lane, mine1));
    }
    for (uint32_t todo = __ballot_sync(0xffffffffu, redo); todo; todo &= todo - 1) {
        const int k = __ffs(todo) - 1;
        const uint32_t u = __shfl_sync(0xffffffffu, my_unit, k);
        const MateInfo a = trim_mate(in0, inter ? 2 * u : u, u, 0, P, ctl, lane);
        MateInfo b = {{-1, -1}, 0};
        if (paired) b = inter ? trim_mate(in0, 2 * u + 1, u, 1, P, ctl, lane) : trim_mate(in1, u, u, 1, P, ctl, lane);
        if (lane == k) { mine0 = a; mine1 = b; }
    }
}

// K2a -- batches of long records: trimming only, every warp on its own.  A warp draws two units at a time and
// leaves their verdicts {five, kept bases, keep} in the descriptor table; k2_trim_route<true> then does the routing
// and the scan from those.  In one kernel a tile's look-back has to wait for every earlier tile, and with reads of
// 1-20 kb the tiles' times differ by an order of magnitude: a quarter of K2's instructions were look-back polls and
// its warps spent more time at the tile barriers than working (ncu, profiles/r2_final_long_*).
// (five and six CTAs per SM -- 48 / 40 registers -- were measured against the four this compiles to: no change,
//  profiles/r2_call29.log; the kernel ends with its slowest warp, a 20 kb read walked by one warp)
__global__ void __launch_bounds__(kK2Threads)
k2_trim_only(DevInput in0, DevInput in1, DevParams P, Control *__restrict__ ctl, RecDesc *__restrict__ desc0,
             RecDesc *__restrict__ desc1, uint32_t upw /* units a warp draws at a time: 2 for long records, 32 otherwise */) {
    const int lane = threadIdx.x & 31;
    const Geometry g = batch_geometry(ctl, P);
    const bool paired = P.mode != 0, inter = P.mode >= 2;
    RangeCheck rc;
    rc.init(P);
    while (true) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(&ctl->k2a_ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        const uint32_t p0 = t * upw;
        if (p0 >= g.nunits) break;
        const bool has = (uint32_t)lane < upw && p0 + lane < g.nunits;
        const uint32_t my_unit = p0 + (uint32_t)lane;
        MateInfo mine0 = {{-1, -1}, 0}, mine1 = {{-1, -1}, 0};
        trim_units(in0, in1, P, rc, ctl, has, my_unit, lane, mine0, mine1);
        if (has) {
            RecDesc d;
            d.dst_off = 0;
            d.route = mine0.cut.three >= 0 ? 1u : 0u;
            d.five = d.route ? (uint32_t)mine0.cut.five : 0u;
            d.nkeep = d.route ? (uint32_t)(mine0.cut.three - mine0.cut.five) : 0u;
            desc0[inter ? 2 * my_unit : my_unit] = d;
            if (paired) {
                d.route = mine1.cut.three >= 0 ? 1u : 0u;
                d.five = d.route ? (uint32_t)mine1.cut.five : 0u;
                d.nkeep = d.route ? (uint32_t)(mine1.cut.three - mine1.cut.five) : 0u;
                (inter ? desc0 : desc1)[inter ? 2 * my_unit + 1 : my_unit] = d;
            }
        }
    }
}

// kPre: the verdicts are in the descriptor table already (k2_trim_only): routing, scan and descriptors only.
template <bool kPre>
__global__ void __launch_bounds__(kK2Threads, 4)
k2_trim_route(DevInput in0, DevInput in1, DevParams P, Control *__restrict__ ctl, RecDesc *__restrict__ desc0,
              RecDesc *__restrict__ desc1, unsigned long long *__restrict__ status /* [3][max_tiles] */,
              uint32_t status_stride, uint32_t epoch) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t warp_tot[kK2Threads / 32][kMaxStreams];
    __shared__ unsigned long long tile_prefix[kMaxStreams];
    __shared__ uint32_t s_cnt[kK2Threads / 32][6];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (kPre && ctl->fast_fail) return;   // the index + verdict pass (kf_fused<CH, 3>) gave the batch up: nothing here is valid
    const Geometry g = batch_geometry(ctl, P);
    // tile geometry (block-uniform, derived from what K1 counted)
    const uint32_t nrec_all = g.nrec0 + g.nrec1;
    const bool long_batch = !kPre && nrec_all > 0 && (in0.nbytes + in1.nbytes) / nrec_all >= kK2LongRecordBytes &&
                            (g.nunits + kK2LongUnitsPerTile - 1) / kK2LongUnitsPerTile < status_stride;
    const uint32_t upw = long_batch ? kK2LongUnitsPerWarp : 32u;                          // units per warp (lanes 0 .. upw-1)
    const uint32_t upt = upw * (kK2Threads / 32);                                         // units per tile
    const bool unit_lane = (uint32_t)lane < upw;
    const uint32_t num_tiles = (g.nunits + upt - 1) / upt;
    const bool paired = P.mode != 0;
    const bool inter = P.mode >= 2;
    const bool mmode = P.mode == 3;
    RangeCheck rc;
    rc.init(P);
    if (tid < (kK2Threads / 32) * 6) (&s_cnt[0][0])[tid] = 0u;

    while (true) {
        if (tid == 0) s_tile = atomicAdd(&ctl->tile_counter[2], 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= num_tiles) break;

        // ---- phase 1: lane k trims unit k by itself (short reads); whatever that path declines is
        // redone by the whole warp, one unit after the other ----
        const uint32_t p0 = tile * upt + wid * upw;
        MateInfo mine0 = {{-1, -1}, 0}, mine1 = {{-1, -1}, 0};
        uint32_t my_unit = 0;
        const bool has_unit = unit_lane && p0 + lane < g.nunits;
        if (has_unit) my_unit = position_to_unit(p0 + lane, g.nunits, P.emu_threads, paired);
        if (!kPre) {
            trim_units(in0, in1, P, rc, ctl, has_unit, my_unit, lane, mine0, mine1);
        } else if (has_unit) {   // verdicts from k2_trim_only; the lines' lengths from the index
            const uint32_t r0 = inter ? 2 * my_unit : my_unit;
            const RecDesc v0 = desc0[r0];
            const RecLines l0 = record_lines(in0, r0);
            mine0.fixed_len = l0.len[0] + l0.len[2] + 4u;
            if (v0.route) mine0.cut = Cut{(int)v0.five, (int)(v0.five + v0.nkeep)};
            if (paired) {
                const uint32_t r1 = inter ? 2 * my_unit + 1 : my_unit;
                const RecDesc v1 = (inter ? desc0 : desc1)[r1];
                const RecLines l1 = record_lines(inter ? in0 : in1, r1);
                mine1.fixed_len = l1.len[0] + l1.len[2] + 4u;
                if (v1.route) mine1.cut = Cut{(int)v1.five, (int)(v1.five + v1.nkeep)};
            }
        }

        // ---- phase 2: routing (lane = unit) ----
        const bool active = unit_lane && p0 + lane < g.nunits;
        const bool k1 = active && mine0.cut.three >= 0;
        const bool k2 = active && paired && mine1.cut.three >= 0;
        const uint32_t n1 = k1 ? (uint32_t)(mine0.cut.three - mine0.cut.five) : 0u;
        const uint32_t n2 = k2 ? (uint32_t)(mine1.cut.three - mine1.cut.five) : 0u;
        const uint32_t len1 = mine0.fixed_len + 2u * n1, len2 = mine1.fixed_len + 2u * n2;  // kept record bytes
        const uint32_t nlen1 = mine0.fixed_len + 2u, nlen2 = mine1.fixed_len + 2u;          // "N record" bytes
        uint32_t add[kMaxStreams] = {0, 0, 0};
        uint32_t route1 = 0, route2 = 0, rel2 = 0;  // rel2: offset of mate 2 after mate 1 in a shared stream
        if (active) {
            if (!paired) {
                if (k1) { add[0] = len1; route1 = kRouteEmit | 0u; }
            } else if (k1 && k2) {                                   // trim_paired.cpp:543-551
                route1 = kRouteEmit | 0u;
                if (inter) { add[0] = len1 + len2; route2 = kRouteEmit | 0u; rel2 = len1; }
                else { add[0] = len1; add[1] = len2; route2 = kRouteEmit | 1u; }
            } else if (k1 || k2) {                                   // trim_paired.cpp:552-563
                if (mmode) {
                    route1 = kRouteEmit | (k1 ? 0u : kRouteNRec);
                    route2 = kRouteEmit | (k2 ? 0u : kRouteNRec);
                    rel2 = k1 ? len1 : nlen1;
                    add[0] = rel2 + (k2 ? len2 : nlen2);
                } else if (P.has_singles) {                          // trim_paired.cpp:601,609
                    if (k1) { route1 = kRouteEmit | 2u; add[2] = len1; }
                    else { route2 = kRouteEmit | 2u; add[2] = len2; }
                }
            } else if (mmode) {                                      // both fail under -M: two N records
                route1 = route2 = kRouteEmit | kRouteNRec;
                rel2 = nlen1;
                add[0] = nlen1 + nlen2;
            }
        }

        // ---- phase 3: 3-stream exclusive scan over the tile, look-back across tiles ----
        uint32_t incl[kMaxStreams];
#pragma unroll
        for (int s = 0; s < kMaxStreams; ++s) {
            incl[s] = warp_incl_scan(add[s], lane);
            if (lane == 31) warp_tot[wid][s] = incl[s];
        }
        __syncthreads();
        if (wid < kMaxStreams) {
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < kK2Threads / 32; ++w) total += warp_tot[w][wid];
            const unsigned long long pre = lookback_exclusive(status + (size_t)wid * status_stride, tile, total, epoch, lane);
            if (lane == 0) {
                tile_prefix[wid] = pre;
                if (tile == num_tiles - 1) ctl->out_bytes[wid] = pre + total;
            }
        }
        __syncthreads();
        uint32_t off[kMaxStreams];
#pragma unroll
        for (int s = 0; s < kMaxStreams; ++s) {
            uint32_t wbase = 0;
            for (int w = 0; w < wid; ++w) wbase += warp_tot[w][s];
            off[s] = (uint32_t)tile_prefix[s] + wbase + incl[s] - add[s];
        }

        // ---- phase 4: descriptors + counters ----
        if (active) {
            RecDesc d1;
            d1.route = route1;
            d1.dst_off = off[route1 & 3u];
            d1.five = k1 ? (uint32_t)mine0.cut.five : 0u;
            d1.nkeep = n1;
            desc0[inter ? 2 * my_unit : my_unit] = d1;
            if (paired) {
                RecDesc d2;
                d2.route = route2;
                d2.dst_off = off[route2 & 3u] + rel2;
                d2.five = k2 ? (uint32_t)mine1.cut.five : 0u;
                d2.nkeep = n2;
                (inter ? desc0 : desc1)[inter ? 2 * my_unit + 1 : my_unit] = d2;
            }
        }
        // counters: trim_single.cpp:391,397; trim_paired.cpp:551,557-562,566
        const uint32_t m_both = __ballot_sync(0xffffffffu, active && paired && k1 && k2);
        const uint32_t m_only1 = __ballot_sync(0xffffffffu, active && paired && k1 && !k2);
        const uint32_t m_only2 = __ballot_sync(0xffffffffu, active && paired && !k1 && k2);
        const uint32_t m_none = __ballot_sync(0xffffffffu, active && paired && !k1 && !k2);
        const uint32_t m_sekeep = __ballot_sync(0xffffffffu, active && !paired && k1);
        const uint32_t m_sedrop = __ballot_sync(0xffffffffu, active && !paired && !k1);
        if (lane == 0) {   // per-warp running totals: they go to the Control block once, when the CTA runs out of tiles
            uint32_t *__restrict__ acc = s_cnt[wid];
            acc[0] += (uint32_t)__popc(m_sekeep); acc[1] += (uint32_t)__popc(m_sedrop);
            acc[2] += (uint32_t)__popc(m_both);   acc[3] += (uint32_t)__popc(m_none);
            acc[4] += (uint32_t)__popc(m_only1);  acc[5] += (uint32_t)__popc(m_only2);
        }
        __syncthreads();  // warp_tot / tile_prefix reused by the next tile
    }
    // counters: trim_single.cpp:391,397; trim_paired.cpp:551,557-562,566 -- one atomic per counter and CTA
    // (one per warp and tile put thousands of atomics on one cache line of the Control block)
    if (tid < 6) {
        uint32_t v = 0;
#pragma unroll
        for (int w = 0; w < kK2Threads / 32; ++w) v += s_cnt[w][tid];
        if (v) {
            switch (tid) {
                case 0: atomicAdd(&ctl->counters[0], (unsigned long long)v); break;
                case 1: atomicAdd(&ctl->counters[1], (unsigned long long)v); break;
                case 2: atomicAdd(&ctl->counters[2], 2ull * v); break;
                case 3: atomicAdd(&ctl->counters[3], 2ull * v); break;
                case 4: atomicAdd(&ctl->counters[4], (unsigned long long)v); atomicAdd(&ctl->counters[7], (unsigned long long)v); break;
                default: atomicAdd(&ctl->counters[5], (unsigned long long)v); atomicAdd(&ctl->counters[6], (unsigned long long)v); break;
            }
        }
    }
}

}  // namespace sk
