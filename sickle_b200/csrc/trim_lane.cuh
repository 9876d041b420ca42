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
    // Three-op form for words whose every byte is a quality byte: a word with a byte outside [qmin, qmax]
    // leaves bit 7 of some byte of `acc` set.  No false negatives: for b < 0x80 neither sum carries, bit 7
    // of (b + add_hi) says b > qmax and bit 7 of ~(b + add_lo) says b < qmin; a byte b >= 0x80 sets bit 7
    // of (b + add_hi) unless that sum wraps (b >= 0x81 + qmax), and then (b + add_lo) wraps to less than
    // 0x80 as well.  Only bytes >= 0x80 carry, so the lowest such byte of a word is always judged without
    // a carry coming in; what a carry does to the bytes above it no longer matters.
    __device__ __forceinline__ void screen3(uint32_t x, uint32_t &acc) const { acc |= (x + add_hi) | ~(x + add_lo); }
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

// Sliding window by ONE lane per read, coarse to fine.  Same decisions as warp_sliding_window.
// Window i is "good" iff total(i) >= qthr*ws (trim.cpp:36,42,61); i5 = first good window, i3 = first
// bad window after it (or the first bad window at all with -x).
//
// Phase 1 walks the ALIGNED 4-byte words of the quality line once.  It (a) screens every quality byte
// against [qmin, qmax] -- any byte outside it makes this function DECLINE (error = true): the callers
// hand such a read to the exact warp-wide path, which knows which bytes the reference's scalar loop
// really visits -- and (b) keeps a sliding sum R over k = (ws - 3) / 4 consecutive whole words.  Every
// window contains k consecutive whole words whatever its phase, all other bytes of it are >= qmin, so
//     total(i) >= R_A - 4k*qoff + (ws - 4k) * min(0, qmin - qoff)      for the windows i that cover words A..A+k-1
// and these windows are i in [4(A+k) - ws - Q, 4A - Q] (Q = byte offset of the quality line): the ranges
// of successive A overlap or abut and window 0 belongs to the first whole word.  So up to the first word
// A_f whose bound falls below qthr*ws every window is known to be good without having been summed:
// window 0 is the first good one and the exact scan may start at w_start = 4*A_f - Q - 3.
// Cost: 2 dp4a + 1 funnel shift + 3 screen ops per 4 quality bytes, against (8 dp4a + 4 shifts + screen)
// per 4 windows for the exact scan, which now only runs from w_start to the first bad window.
//
// Phase 2 is the exact scan: 32 windows per step, totals by dp4a straight from the packed quality
// words (prefix sums of lead - trail inside a word), the sign of every total shifted into a bit mask.
__device__ __forceinline__ TrimOut lane_sliding_window(const uint8_t *__restrict__ sm, uint32_t seq_off, uint32_t L,
                                                       uint32_t qual_off, const DevParams &P, const RangeCheck &rc) {
    TrimOut o;
    o.five = -1; o.three = -1; o.error = false;
    if (L < (uint32_t)P.lthr) return o;                                  // trim.cpp:21-26 (nothing is looked at)
    uint32_t ws = L / 10u;                                               // trim.cpp:8
    if (ws == 0) ws = L;                                                 // trim.cpp:30
    const long long thr_ll = (long long)P.qthr * (long long)ws;
    // window totals of a record that fits a tile are < 2^23; clamp so that T - thr cannot overflow
    const int thr = thr_ll > 0x3fffffffLL ? 0x3fffffff : (int)thr_ll;
    const uint32_t nwin = L - ws + 1u;                                   // trim.cpp:34
    const bool x = P.no_fiveprime != 0;                                  // -x: as if the 5' end was already found

    // ---- phase 1: screen + word-granular lower bounds
    uint32_t w_start = 0;                     // every window before it is good
    {
        const uint32_t *__restrict__ W = reinterpret_cast<const uint32_t *>(sm);
        const uint32_t Q = qual_off, E = qual_off + L;                   // quality bytes are [Q, E)
        const uint32_t A0 = (Q + 3u) >> 2;                               // first whole word
        const uint32_t A1 = E >> 2;                                      // one past the last whole word
        uint32_t scr = 0;                                                // bit 7 of a byte set: a byte outside [qmin, qmax]
        const uint32_t fill = rc.kmin;                                   // qmin in every byte: passes the screen
        if (A0 > A1) {                                                   // no whole word: the line lies inside one word
            const uint32_t m = (0xffffffffu << (8u * (Q & 3u))) & ~(0xffffffffu << (8u * (E & 3u)));
            rc.screen3((W[Q >> 2] & m) | (fill & ~m), scr);
        } else {
            if (Q & 3u) {                                                // ragged head: bytes [Q, 4*A0)
                const uint32_t m = 0xffffffffu << (8u * (Q & 3u));
                rc.screen3((W[A0 - 1u] & m) | (fill & ~m), scr);
            }
            if (E & 3u) {                                                // ragged tail: bytes [4*A1, E)
                const uint32_t m = ~(0xffffffffu << (8u * (E & 3u)));
                rc.screen3((W[A1] & m) | (fill & ~m), scr);
            }
            const uint32_t nfull = A1 - A0;
            const uint32_t k = ws >= 7u ? (ws - 3u) >> 2 : 0u;           // whole words inside every window
            if (k == 0u || nfull < k) {                                  // short windows: screen only
                for (uint32_t j = 0; j < nfull; ++j) rc.screen3(W[A0 + j], scr);
            } else {
                const int slack = P.qmin < P.qoff ? (int)(ws - 4u * k) * (P.qmin - P.qoff) : 0;
                // R = (sum of the k words ending at the lead word) - bias; the word in front of the first
                // whole word is added here and taken out again by the first trail subtraction
                int R = (int)__dp4a(W[A0 - 1u], 0x01010101u, 0u) - (thr + (int)(4u * k) * P.qoff - slack);
                for (uint32_t j = 0; j + 1u < k; ++j) {
                    const uint32_t v = W[A0 + j];
                    rc.screen3(v, scr);
                    R = (int)__dp4a(v, 0x01010101u, (uint32_t)R);
                }
                const uint32_t *__restrict__ lead = W + A0 + (k - 1u);
                const uint32_t *__restrict__ trail = W + A0 - 1u;
                const uint32_t nstep = nfull - k + 1u;                   // bounds R_A for A = A0 .. A0 + nstep - 1
                uint32_t jf = nstep;                                     // first A - A0 whose bound is below the threshold
                uint32_t j = 0;
                for (; j + 8u <= nstep; j += 8u) {
                    uint32_t negr = 0;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint32_t lw = lead[j + u], tw = trail[j + u];
                        rc.screen3(lw, scr);
                        R = dp4a_us(lw, 0x01010101, dp4a_us(tw, (int)0xFFFFFFFF, R));
                        negr = __funnelshift_l((uint32_t)R, negr, 1);
                    }
                    if (negr && jf == nstep) jf = j + (uint32_t)__clz(negr) - 24u;
                }
                for (; j < nstep; ++j) {
                    const uint32_t lw = lead[j], tw = trail[j];
                    rc.screen3(lw, scr);
                    R = dp4a_us(lw, 0x01010101, dp4a_us(tw, (int)0xFFFFFFFF, R));
                    if (R < 0 && jf == nstep) jf = j;
                }
                // windows up to 4*(A_f - 1) - Q are good; no failing word: every window is
                const int ws_i = jf == nstep ? (int)nwin : 4 * (int)(A0 + jf) - (int)Q - 3;
                w_start = ws_i < 0 ? 0u : min((uint32_t)ws_i, nwin);
            }
        }
        if (scr & 0x80808080u) { o.error = true; return o; }
    }

    // ---- phase 2: exact scan from w_start.  w_start > 0 means window 0 is good (it is the first good one).
    int g = w_start > 0 ? 0 : -1;              // first good window
    int b_stop = -1;                           // the window the loop breaks at: first bad one after g (-x: first bad one)
    if (w_start < nwin) {
        int T = 0;
        {
            Stream4 s;
            s.init(sm, qual_off + w_start);
            uint32_t j = 0;
            for (; j + 4 <= ws; j += 4) T = (int)__dp4a(s.next(), 0x01010101u, (uint32_t)T);
            const uint8_t *__restrict__ q = sm + qual_off + w_start;
            for (; j < ws; ++j) T += q[j];
            T -= (int)ws * P.qoff;
        }
        int Tm = T - thr;                      // sign bit set <=> window is bad
        Stream4 lead, trail;
        lead.init(sm, qual_off + w_start + ws);
        trail.init(sm, qual_off + w_start);
        for (uint32_t base = w_start; base < nwin && b_stop < 0; base += 32) {
            // bit 31-k: window base+k is bad.  Each total's sign bit is shifted in from the right by one
            // funnel shift, so the step's first window ends up in the top bit (first = __clz).
            uint32_t negr = 0;
#pragma unroll
            for (int gq = 0; gq < 8; ++gq) {
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
                Tm = T4;
            }
            const uint32_t left = nwin - base;                               // windows from base on (>= 1)
            const uint32_t vmask = left >= 32 ? 0xffffffffu : ~(0xffffffffu >> left);   // top `left` bits
            const uint32_t goodw = ~negr & vmask, badw = negr & vmask;
            uint32_t after = 0xffffffffu;
            if (!x && g < 0) {                                               // first good window: trim.cpp:42
                if (goodw) {
                    const int kk = __clz(goodw);
                    g = (int)base + kk;
                    after = 0xffffffffu >> kk;
                } else after = 0;
            }
            const uint32_t cand = badw & after;                              // first bad window after it: trim.cpp:61
            if (cand) b_stop = (int)base + __clz(cand);
        }
    }
    const int i5 = x ? -1 : g;
    const int i3 = b_stop;

    // The two in-window scans only touch bytes phase 1 has range-checked (all < 128), so
    // "q - qoff >= qthr" is the SWAR test "(b | 0x80) - c has bit 7 set" with c = qthr + qoff.
    int five = 0, three = (int)L;
    const int cthr = P.qthr + P.qoff;
    const uint32_t c4 = (uint32_t)(cthr < 1 ? 0 : (cthr > 128 ? 128 : cthr)) * 0x01010101u;
    if (i5 >= 0) {                                                           // trim.cpp:46-51
        Stream4 s;
        s.init(sm, qual_off + (uint32_t)i5);
        for (uint32_t j = 0; j < ws; j += 4) {
            const uint32_t ge = ((s.next() | 0x80808080u) - c4) & 0x80808080u;
            uint32_t nib = flags_to_nibble(ge);
            if (ws - j < 4) nib &= (1u << (ws - j)) - 1u;
            if (nib) { five = i5 + (int)j + __ffs(nib) - 1; break; }
        }
    }
    if (i3 >= 0) {                                                           // trim.cpp:65-70
        Stream4 s;
        s.init(sm, qual_off + (uint32_t)i3);
        for (uint32_t j = 0; j < ws; j += 4) {
            const uint32_t lt = ~((s.next() | 0x80808080u) - c4) & 0x80808080u;
            uint32_t nib = flags_to_nibble(lt);
            if (ws - j < 4) nib &= (1u << (ws - j)) - 1u;
            if (nib) { three = i3 + (int)j + __ffs(nib) - 1; break; }
        }
    }
    if (P.trunc_n) {                                                         // trim.cpp:86-98
        int pn = -1;
        uint32_t anyN = 0;
        Stream4 s;
        s.init(sm, seq_off);
        for (uint32_t j = 0; j < L; j += 4) {
            const uint32_t v = s.next();
            // exact zero-byte tests of v ^ 'n' and v ^ 'N' (same trick as newline_flags)
            const uint32_t tn = ((v ^ 0x6E6E6E6Eu) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
            const uint32_t tN = ((v ^ 0x4E4E4E4Eu) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
            uint32_t fn = flags_to_nibble(~(tn | v) & 0x80808080u);
            uint32_t fN = flags_to_nibble(~(tN | v) & 0x80808080u);
            if (L - j < 4) { const uint32_t m = (1u << (L - j)) - 1u; fn &= m; fN &= m; }
            if (fn) { pn = (int)j + __ffs(fn) - 1; break; }                 // a lowercase n wins over any N (trim.cpp:88-93)
            anyN |= fN;
        }
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
