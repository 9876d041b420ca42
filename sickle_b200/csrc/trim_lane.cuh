// trim_lane.cuh -- the sliding-window trimmer run by one or two lanes per read (reference
// src/trim.cpp:3-140), shared by the single-pass kernel (reads staged in shared memory) and by K2's
// short-read path (reads straight from global memory).
#pragma once

#include "k1_index.cuh"
#include "sk_device.cuh"

namespace sk {

// four successive bytes at a time from an arbitrary byte offset (shared or global memory; reads whole
// aligned words, up to 7 bytes past the last byte asked for)
struct Stream4 {
    const uint32_t *w;
    uint32_t cur, nxt, sh;
    __device__ __forceinline__ void init(const uint8_t *sm, uint32_t off) {
        w = reinterpret_cast<const uint32_t *>(sm + (off & ~3u));
        sh = (off & 3u) * 8u;
        cur = w[0];
        nxt = w[1];
        w += 2;
    }
    __device__ __forceinline__ uint32_t next() {
        const uint32_t r = __funnelshift_r(cur, nxt, sh);
        cur = nxt;
        nxt = *w++;
        return r;
    }
};

struct RangeCheck {
    uint32_t kmin, kmax;   // qmin * 0x01010101, (qmax | 0x80) * 0x01010101
    uint32_t add_hi, add_lo;   // (0x7f - qmax) * 0x01010101, (0x80 - qmin) * 0x01010101
    __device__ __forceinline__ void init(const DevParams &P) {
        kmin = (uint32_t)P.qmin * 0x01010101u;
        kmax = ((uint32_t)P.qmax | 0x80u) * 0x01010101u;
        add_hi = (0x7fu - (uint32_t)P.qmax) * 0x01010101u;
        add_lo = (0x80u - (uint32_t)P.qmin) * 0x01010101u;
    }
    // Cheap screen for many words: fold every word into (hi, lo) with screen(), then suspicious() says
    // whether ANY byte seen may be out of range.  No false negatives (a carry out of a byte >= 0x80
    // can only disturb its neighbour, and that byte is flagged through `x` itself); callers re-check
    // the words exactly with bad4() when it fires.
    __device__ __forceinline__ void screen(uint32_t x, uint32_t &hi, uint32_t &lo) const {
        hi |= (x + add_hi) | x;      // bit 7: byte > qmax, or >= 0x80
        lo &= x + add_lo;            // bit 7 cleared: byte < qmin
    }
    __device__ __forceinline__ bool suspicious(uint32_t hi, uint32_t lo) const { return ((hi | ~lo) & 0x80808080u) != 0; }
    // 0x80 in every byte of x that is outside [qmin, qmax] (qmax <= 126)
    __device__ __forceinline__ uint32_t bad4(uint32_t x) const {
        const uint32_t lo = (x | 0x80808080u) - kmin;          // bit 7 set iff (b & 0x7f) >= qmin
        const uint32_t up = kmax - (x & 0x7F7F7F7Fu);          // bit 7 set iff (b & 0x7f) <= qmax
        return ~(lo & up & ~x) & 0x80808080u;
    }
};

struct TrimOut {
    int five, three;   // three < 0 => discard
    bool error;        // a quality byte outside the encoding's range inside the visited prefix
};

// dp4a with unsigned bytes in a and signed bytes in b
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {
#if defined(__CUDACC__)
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else   // host build of this header (tests/test_lane_logic.py): the same sum, spelled out
    for (int k = 0; k < 4; ++k) c += (int)((a >> (8 * k)) & 255u) * (int)(int8_t)((uint32_t)b >> (8 * k));
    return c;
#endif
}

// Sliding window over shared memory by ONE or TWO lanes per read.  Same decisions as
// warp_sliding_window.  Window i is "good" iff total(i) >= qthr*ws (trim.cpp:36,42,61).  Let i5 = first
// good window, i3 = first bad window after it (or the first bad window at all with -x).  Quality bytes
// are range-checked exactly where the reference's scalar loop touches them: the first window always,
// and the byte entering window w+1 iff the loop gets past window w (w < i3 and w+1 < nwin).
//
// With nsub == 2 the windows are cut into two contiguous halves (whole 32-window steps); lane `sub`
// scans its half without knowing what the other finds, and records what either outcome needs: its
// first good window, its first bad window, its first bad window after its first good one, and the
// first window whose entering byte is out of range.  The pair then exchanges these four numbers
// (shuffles inside the pair only: both lanes of a read take the same branches up to there) and
// both derive i5 / i3 / the error exactly as the sequential loop would have.  After that lane 0
// looks for the 5' cut inside window i5 while lane 1 looks for the 3' cut inside window i3.
struct HalfScan {
    int g, b_any, b_after, o_first;
};
__device__ __forceinline__ TrimOut lane_sliding_window(const uint8_t *__restrict__ sm, uint32_t seq_off, uint32_t L,
                                                       uint32_t qual_off, const DevParams &P, const RangeCheck &rc,
                                                       uint32_t sub, uint32_t nsub, int lane) {
    TrimOut o;
    o.five = -1; o.three = -1; o.error = false;
    if (L < (uint32_t)P.lthr) return o;                                  // trim.cpp:21-26 (same for both lanes)
    uint32_t ws = L / 10u;                                               // trim.cpp:8
    if (ws == 0) ws = L;                                                 // trim.cpp:30
    const long long thr_ll = (long long)P.qthr * (long long)ws;
    // window totals of a record that fits a tile are < 2^23; clamp so that T - thr cannot overflow
    const int thr = thr_ll > 0x3fffffffLL ? 0x3fffffff : (int)thr_ll;
    const uint32_t nwin = L - ws + 1u;                                   // trim.cpp:34
#ifdef SK_LANE_SPLIT4
    // Experimental (off by default; next round's A/B): the two lanes get equal shares, cut at a multiple
    // of 4 windows, and the last step of a share only runs the groups it needs -- instead of whole
    // 32-window steps (136 windows = 96 + 40 becomes 68 + 68: 17 groups per lane instead of 24).
    const uint32_t split = nsub == 2u ? min(nwin, (((nwin + 1u) >> 1) + 3u) & ~3u) : nwin;
    const uint32_t w_lo = sub ? split : 0u;                              // first window of this lane
    const uint32_t w_hi = sub ? nwin : split;                            // one past its last window
#else
    const uint32_t nseg = (nwin + 31u) >> 5;
    const uint32_t half = nsub == 2u ? (nseg + 1u) >> 1 : nseg;
    const uint32_t w_lo = sub ? half * 32u : 0u;                         // first window of this lane
    const uint32_t w_hi = sub ? nwin : min(nwin, half * 32u);            // one past its last window
#endif
    const bool x = P.no_fiveprime != 0;                                  // -x: as if the 5' end was already found

    // ---- total of this lane's first window: dp4a sums, 4 bytes at a time.  Window 0 is also range
    // checked here (trim.cpp:31-33); the bytes of lane 1's first window are entering bytes of lane 0's.
    int T = 0;
    uint32_t bad = 0;
    if (w_lo < w_hi) {
        Stream4 s;
        s.init(sm, qual_off + w_lo);
        uint32_t j = 0;
        for (; j + 4 <= ws; j += 4) {
            const uint32_t v = s.next();
            bad |= rc.bad4(v);
            T = (int)__dp4a(v, 0x01010101u, (uint32_t)T);
        }
        const uint8_t *__restrict__ q = sm + qual_off + w_lo;
        for (; j < ws; ++j) {
            const int b = q[j];
            bad |= (uint32_t)((b < P.qmin) | (b > P.qmax));
            T += b;
        }
        T -= (int)ws * P.qoff;
    }
    if (sub) bad = 0;
    int Tm = T - thr;                          // sign bit set <=> window is bad
    HalfScan me;
    me.g = -1; me.b_any = -1; me.b_after = -1; me.o_first = 0x7fffffff;

    Stream4 lead, trail;
    lead.init(sm, qual_off + w_lo + ws);
    trail.init(sm, qual_off + w_lo);
    for (uint32_t base = w_lo; base < w_hi && bad == 0 && (x ? me.b_any : me.b_after) < 0; base += 32) {
        // bit 31-k: window base+k is bad.  Each total's sign bit is shifted in from the right by one
        // funnel shift, so the step's first window ends up in the top bit (first = __clz).
        uint32_t negr = 0;
        uint32_t r_hi = 0, r_lo = 0xffffffffu;   // range screen of the 32 entering bytes
#ifdef SK_LANE_SPLIT4
        const uint32_t span = min(32u, w_hi - base);                     // windows of this step (>= 1)
        auto group = [&]() {
            const uint32_t lw = lead.next(), tw = trail.next();
            const int T1 = dp4a_us(lw, 0x00000001, dp4a_us(tw, 0x000000FF, Tm));
            const int T2 = dp4a_us(lw, 0x00000101, dp4a_us(tw, 0x0000FFFF, Tm));
            const int T3 = dp4a_us(lw, 0x00010101, dp4a_us(tw, 0x00FFFFFF, Tm));
            const int T4 = dp4a_us(lw, 0x01010101, dp4a_us(tw, (int)0xFFFFFFFF, Tm));
            negr = __funnelshift_l((uint32_t)Tm, negr, 1);
            negr = __funnelshift_l((uint32_t)T1, negr, 1);
            negr = __funnelshift_l((uint32_t)T2, negr, 1);
            negr = __funnelshift_l((uint32_t)T3, negr, 1);
            rc.screen(lw, r_hi, r_lo);
            Tm = T4;
        };
        if (span == 32u) {
#pragma unroll
            for (int g = 0; g < 8; ++g) group();
        } else {                                 // the share's last step: only the groups that hold its windows
            const uint32_t ng = (span + 3u) >> 2;
#pragma unroll 1
            for (uint32_t g = 0; g < ng; ++g) group();
            negr <<= 32u - 4u * ng;              // first window of the step back in the top bit
        }
        // ---- what the windows of this step contribute
        const uint32_t left = nwin - base;                               // windows of the READ from base on (>= 1)
        if ((left <= 32u || rc.suspicious(r_hi, r_lo)) && me.o_first == 0x7fffffff) {
            // the read's last step (its words run past the quality line) or, rarely, a suspect byte:
            // find the first window of this step whose entering byte is out of range, exactly
            // (bit k of oorw: the byte entering window base+k+1)
            Stream4 again;
            again.init(sm, qual_off + base + ws);
            uint32_t oorw = 0;
#pragma unroll 1
            for (int g = 0; g < 8; ++g) oorw += flags_to_nibble(rc.bad4(again.next())) << (4 * g);
            // only this step's windows count, and the last window of the read has no entering byte
            const uint32_t lim = min(span, left - 1u);
            const uint32_t oo = oorw & (lim >= 32u ? 0xffffffffu : ((1u << lim) - 1u));
            if (oo) me.o_first = (int)base + __ffs(oo) - 1;
        }
        const uint32_t vmask = span >= 32u ? 0xffffffffu : ~(0xffffffffu >> span);   // top `span` bits
#else
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const uint32_t lw = lead.next(), tw = trail.next();
            // totals of windows base+4g+1 .. +4: prefix sums of (lead - trail), independent of each other
            const int T1 = dp4a_us(lw, 0x00000001, dp4a_us(tw, 0x000000FF, Tm));
            const int T2 = dp4a_us(lw, 0x00000101, dp4a_us(tw, 0x0000FFFF, Tm));
            const int T3 = dp4a_us(lw, 0x00010101, dp4a_us(tw, 0x00FFFFFF, Tm));
            const int T4 = dp4a_us(lw, 0x01010101, dp4a_us(tw, (int)0xFFFFFFFF, Tm));
            negr = __funnelshift_l((uint32_t)Tm, negr, 1);
            negr = __funnelshift_l((uint32_t)T1, negr, 1);
            negr = __funnelshift_l((uint32_t)T2, negr, 1);
            negr = __funnelshift_l((uint32_t)T3, negr, 1);
            rc.screen(lw, r_hi, r_lo);
            Tm = T4;
        }
        // ---- what the 32 windows of this step contribute
        const uint32_t left = nwin - base;                               // windows from base on (>= 1)
        if ((left <= 32u || rc.suspicious(r_hi, r_lo)) && me.o_first == 0x7fffffff) {
            // the read's last step (its words run past the quality line) or, rarely, a suspect byte:
            // find the first window of this step whose entering byte is out of range, exactly
            // (bit k of oorw: the byte entering window base+k+1)
            Stream4 again;
            again.init(sm, qual_off + base + ws);
            uint32_t oorw = 0;
#pragma unroll 1
            for (int g = 0; g < 8; ++g) oorw += flags_to_nibble(rc.bad4(again.next())) << (4 * g);
            // the last window of the read has no entering byte
            const uint32_t oo = oorw & (left > 32 ? 0xffffffffu : ((1u << (left - 1u)) - 1u));
            if (oo) me.o_first = (int)base + __ffs(oo) - 1;
        }
        const uint32_t vmask = left >= 32 ? 0xffffffffu : ~(0xffffffffu >> left);   // top `left` bits
#endif
        const uint32_t goodw = ~negr & vmask, badw = negr & vmask;
        if (badw && me.b_any < 0) me.b_any = (int)base + __clz(badw);
        uint32_t after = 0xffffffffu;
        if (me.g < 0) {                                                  // first good window: trim.cpp:42
            if (goodw) {
                const int k = __clz(goodw);
                me.g = (int)base + k;
                after = 0xffffffffu >> k;
            } else after = 0;
        }
        const uint32_t cand = badw & after;                              // first bad window after it: trim.cpp:61
        if (cand && me.b_after < 0) me.b_after = (int)base + __clz(cand);
    }

    // ---- the two halves in order (A = windows from 0, B = the rest)
    HalfScan A = me, B;
    B.g = -1; B.b_any = -1; B.b_after = -1; B.o_first = 0x7fffffff;
    if (nsub == 2u) {
        const uint32_t pm = 3u << (lane & 30);
        HalfScan ot;
        ot.g = __shfl_xor_sync(pm, me.g, 1);
        ot.b_any = __shfl_xor_sync(pm, me.b_any, 1);
        ot.b_after = __shfl_xor_sync(pm, me.b_after, 1);
        ot.o_first = __shfl_xor_sync(pm, me.o_first, 1);
        bad |= __shfl_xor_sync(pm, bad, 1);
        if (sub) { A = ot; B = me; } else B = ot;
    }
    if (bad) { o.error = true; return o; }
    int i5 = x ? -1 : A.g;
    int i3 = x ? A.b_any : A.b_after;
    if (i3 < 0) {                                   // the loop runs on into the second half
        if (x || A.g >= 0) i3 = B.b_any;
        else { i5 = B.g; i3 = B.b_after; }
    }
    // entering bytes are fetched for the windows before the break (or before the last window)
    const int fetched = i3 >= 0 ? i3 : (int)nwin - 1;
    if (min(A.o_first, B.o_first) < fetched) { o.error = true; return o; }

    // The two in-window scans only touch bytes the loop above has range-checked (all < 128), so
    // "q - qoff >= qthr" is the SWAR test "(b | 0x80) - c has bit 7 set" with c = qthr + qoff.
    int five = 0, three = (int)L;
    const int cthr = P.qthr + P.qoff;
    const uint32_t c4 = (uint32_t)(cthr < 1 ? 0 : (cthr > 128 ? 128 : cthr)) * 0x01010101u;
    if (i5 >= 0 && (nsub == 1u || sub == 0u)) {                          // trim.cpp:46-51
        Stream4 s;
        s.init(sm, qual_off + (uint32_t)i5);
        for (uint32_t j = 0; j < ws; j += 4) {
            const uint32_t ge = ((s.next() | 0x80808080u) - c4) & 0x80808080u;
            uint32_t nib = flags_to_nibble(ge);
            if (ws - j < 4) nib &= (1u << (ws - j)) - 1u;
            if (nib) { five = i5 + (int)j + __ffs(nib) - 1; break; }
        }
    }
    if (i3 >= 0 && (nsub == 1u || sub == 1u)) {                          // trim.cpp:65-70
        Stream4 s;
        s.init(sm, qual_off + (uint32_t)i3);
        for (uint32_t j = 0; j < ws; j += 4) {
            const uint32_t lt = ~((s.next() | 0x80808080u) - c4) & 0x80808080u;
            uint32_t nib = flags_to_nibble(lt);
            if (ws - j < 4) nib &= (1u << (ws - j)) - 1u;
            if (nib) { three = i3 + (int)j + __ffs(nib) - 1; break; }
        }
    }
    int pn = -1;
    uint32_t anyN = 0;
    if (P.trunc_n) {                                                     // trim.cpp:86-98; each lane one half of the bases
        const uint32_t mid = nsub == 2u ? ((L / 2u + 3u) & ~3u) : L;
        const uint32_t j0 = sub ? mid : 0u, j1 = sub ? L : min(mid, L);
        Stream4 s;
        s.init(sm, seq_off + j0);
        for (uint32_t j = j0; j < j1; j += 4) {
            const uint32_t v = s.next();
            // exact zero-byte tests of v ^ 'n' and v ^ 'N' (same trick as newline_flags)
            const uint32_t tn = ((v ^ 0x6E6E6E6Eu) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
            const uint32_t tN = ((v ^ 0x4E4E4E4Eu) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
            uint32_t fn = flags_to_nibble(~(tn | v) & 0x80808080u);
            uint32_t fN = flags_to_nibble(~(tN | v) & 0x80808080u);
            if (j1 - j < 4) { const uint32_t m = (1u << (j1 - j)) - 1u; fn &= m; fN &= m; }
            if (fn) { pn = (int)j + __ffs(fn) - 1; break; }
            anyN |= fN;
        }
    }
    if (nsub == 2u) {
        const uint32_t pm = 3u << (lane & 30);
        const int o_cut = __shfl_xor_sync(pm, sub ? three : five, 1);    // lane 0 sends five, lane 1 sends three
        if (sub) five = o_cut; else three = o_cut;
        if (P.trunc_n) {
            const int o_pn = __shfl_xor_sync(pm, pn, 1);
            const uint32_t o_any = __shfl_xor_sync(pm, anyN, 1);
            // the scan stops at the first lowercase n: an uppercase N only counts if it comes before it
            const int pn_a = sub ? o_pn : pn, pn_b = sub ? pn : o_pn;
            const uint32_t any_a = sub ? o_any : anyN, any_b = sub ? anyN : o_any;
            pn = pn_a >= 0 ? pn_a : pn_b;
            anyN = any_a | (pn_a >= 0 ? 0u : any_b);
        }
    }
    if (P.trunc_n) {
        if (pn >= 0) three = pn - 1;
        else if (anyN) three = -2;
    }
    const bool have5 = (i5 >= 0) || x;
    if (!have5 || (three - five < P.lthr)) return o;                     // trim.cpp:103
    o.five = five;
    o.three = three;
    return o;
}

}  // namespace sk
