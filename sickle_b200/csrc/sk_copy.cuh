// sk_copy.cuh -- the byte movers of the single-pass kernel: record pieces inside shared memory
// (S8a staging) and the staged tile out to global memory at an arbitrary 16-byte phase (S8b flush).
// Output bytes are the reference's formatted records (src/trim_single.cpp:393-396), so these two
// routines are what "byte-identical" finally rests on; they have no barriers or warp collectives and are
// also compiled for the host and checked against memcpy by tests/test_copy_logic.py.
#pragma once

#include "sk_device.cuh"

namespace sk {

// Thread-sequential copy inside shared memory, arbitrary alignment on both sides.  The bulk moves
// as 16-byte destination-aligned stores fed by funnel-shifted source words.
__device__ __forceinline__ void smem_copy(uint8_t *__restrict__ out, uint32_t dst, const uint8_t *__restrict__ in,
                                          uint32_t src, uint32_t len) {
    // Straight-line (predicated) ragged ends keep the lanes of a warp converged; only the 16-byte
    // loop has a data-dependent trip count, and that is nearly the same for reads of similar length.
    // head: up to 3 bytes to a 4-byte destination boundary
    const uint32_t h = min(len, (0u - dst) & 3u);
#pragma unroll
    for (uint32_t j = 0; j < 3; ++j)
        if (j < h) out[dst + j] = in[src + j];
    dst += h; src += h; len -= h;
    const uint32_t sh = (src & 3u) * 8u;
    const uint32_t *__restrict__ w = reinterpret_cast<const uint32_t *>(in + (src & ~3u));
    uint32_t *__restrict__ d = reinterpret_cast<uint32_t *>(out + dst);
    const uint32_t nw = len >> 2;                 // whole destination words
    uint32_t cur = w[0];
    // up to 3 words to a 16-byte destination boundary
    const uint32_t pre = min(nw, ((0u - dst) >> 2) & 3u);
#pragma unroll
    for (uint32_t j = 0; j < 3; ++j) {
        if (j < pre) {
            const uint32_t nxt = w[j + 1];
            d[j] = __funnelshift_r(cur, nxt, sh);
            cur = nxt;
        }
    }
    uint32_t k = pre;
    // (unrolling this loop was measured: no gain -- the copy is bound by shared-memory bank conflicts
    // of 32 lanes walking 32 unrelated records, ~3 wavefronts per access, not by the loop's latency)
    for (; k + 4 <= nw; k += 4) {
        const uint32_t a = w[k + 1], b = w[k + 2], c = w[k + 3], e = w[k + 4];
        uint4 v;
        v.x = __funnelshift_r(cur, a, sh);
        v.y = __funnelshift_r(a, b, sh);
        v.z = __funnelshift_r(b, c, sh);
        v.w = __funnelshift_r(c, e, sh);
        *reinterpret_cast<uint4 *>(d + k) = v;
        cur = e;
    }
    const uint32_t rem = nw - k;                  // 0..3 trailing words
#pragma unroll
    for (uint32_t j = 0; j < 3; ++j) {
        if (j < rem) {
            const uint32_t nxt = w[k + j + 1];
            d[k + j] = __funnelshift_r(cur, nxt, sh);
            cur = nxt;
        }
    }
    dst += nw * 4; src += nw * 4; len -= nw * 4;  // 0..3 trailing bytes
#pragma unroll
    for (uint32_t j = 0; j < 3; ++j)
        if (j < len) out[dst + j] = in[src + j];
}

// Flush `tot` staged bytes (shared memory, starting at the 16-byte aligned offset `sb`) to the global
// address `gdst`, which has an arbitrary 16-byte phase: destination-aligned 16-byte stores fed by a
// 128-bit funnel shift of two aligned shared-memory chunks; byte stores only on the two ragged ends.
// Q = word part of the shift (block-uniform, so the four variants never diverge).
template <int Q>
__device__ __forceinline__ void flush_full_chunks(uint4 *__restrict__ gal16, const uint4 *__restrict__ s16, uint32_t c_lo,
                                                  uint32_t c_hi, uint32_t sh, int tid, int nthreads) {
    for (uint32_t c = c_lo + tid; c < c_hi; c += nthreads) {
        const uint4 A = s16[c - 1], B = s16[c];
        const uint32_t w0 = Q == 0 ? A.x : Q == 1 ? A.y : Q == 2 ? A.z : A.w;
        const uint32_t w1 = Q == 0 ? A.y : Q == 1 ? A.z : Q == 2 ? A.w : B.x;
        const uint32_t w2 = Q == 0 ? A.z : Q == 1 ? A.w : Q == 2 ? B.x : B.y;
        const uint32_t w3 = Q == 0 ? A.w : Q == 1 ? B.x : Q == 2 ? B.y : B.z;
        const uint32_t w4 = Q == 0 ? B.x : Q == 1 ? B.y : Q == 2 ? B.z : B.w;
        uint4 v;
        v.x = __funnelshift_r(w0, w1, sh);
        v.y = __funnelshift_r(w1, w2, sh);
        v.z = __funnelshift_r(w2, w3, sh);
        v.w = __funnelshift_r(w3, w4, sh);
        __stcs(gal16 + c, v);
    }
}

__device__ __noinline__ void flush_realigned(uint8_t *__restrict__ gdst, const uint8_t *__restrict__ s_out,
                                                uint32_t sb, uint32_t tot, int tid, int nthreads) {
    if (tot == 0) return;
    const uint32_t ph = (uint32_t)(reinterpret_cast<uintptr_t>(gdst) & 15u);
    uint8_t *gal = gdst - ph;                                  // 16-byte aligned
    const uint32_t end = ph + tot;                             // in bytes from gal
    const uint32_t c_lo = ph ? 1u : 0u;                        // first chunk that is written in full
    const uint32_t c_hi = end >> 4;                            // one past the last full chunk
    const uint4 *__restrict__ s16 = reinterpret_cast<const uint4 *>(s_out + sb);
    if (ph == 0) {
        for (uint32_t c = tid; c < c_hi; c += nthreads) __stcs(reinterpret_cast<uint4 *>(gal) + c, s16[c]);
    } else {
        const uint32_t r = 16u - ph;                           // byte offset inside the older chunk
        const uint32_t sh = (r & 3u) * 8u;
        switch (r >> 2) {
            case 0: flush_full_chunks<0>(reinterpret_cast<uint4 *>(gal), s16, c_lo, c_hi, sh, tid, nthreads); break;
            case 1: flush_full_chunks<1>(reinterpret_cast<uint4 *>(gal), s16, c_lo, c_hi, sh, tid, nthreads); break;
            case 2: flush_full_chunks<2>(reinterpret_cast<uint4 *>(gal), s16, c_lo, c_hi, sh, tid, nthreads); break;
            default: flush_full_chunks<3>(reinterpret_cast<uint4 *>(gal), s16, c_lo, c_hi, sh, tid, nthreads); break;
        }
    }
    // ragged head (bytes [ph, 16) of chunk 0) and tail (bytes of the last, partial chunk): one byte per thread
    if (ph && tid < 16) {
        const uint32_t b = (uint32_t)tid;
        if (b >= ph && b < end) gal[b] = s_out[sb + b - ph];
    }
    const int tail0 = nthreads >= 48 ? 32 : 16;               // (a single warp flushing: lanes 16..31)
    if ((end & 15u) && c_hi >= c_lo && tid >= tail0 && tid < tail0 + 16) {
        const uint32_t b = 16u * c_hi + (uint32_t)(tid - tail0);
        if (b < end && (c_hi > 0 || ph == 0)) gal[b] = s_out[sb + b - ph];
    }
}

}  // namespace sk
