// k1_index.cuh -- K1: FASTQ batch -> line index, single pass (decoupled look-back).
//
// Replaces the serial per-line split of the reference: GZReader::read_lines (gzgets + new char[] +
// strncpy per line, src/GZReader.cpp:76-92), Batch::Batch (strlen per line, src/Batch.cpp:6-21) and
// the four next_line() calls of FQEntry::FQEntry (src/FQEntry.cpp:8-18).  After K1, record r of the
// batch is lines 4r..4r+3 and line l is bytes (line_end[l-1], line_end[l]) -- 4 bytes of index per
// line instead of a heap string + two string_views.
//
// Tile = 256 threads x 64 bytes = 16 KiB.  Loads are 16-byte, fully coalesced (a warp reads 512
// contiguous bytes per instruction), staged through XOR-swizzled shared memory so that each thread
// then owns 64 *contiguous* bytes (bank-conflict free both ways).  Newlines are found with a SWAR
// zero-byte test, counted with popc, ranked with a warp-shuffle + cross-warp prefix scan, and the tile
// prefix comes from a decoupled look-back over (padded) 8-byte status words.  Bytes are read once.
#pragma once

#include "sk_device.cuh"

namespace sk {

constexpr int kK1Threads = 256;
constexpr int kK1BytesPerThread = 64;
constexpr int kK1TileBytes = kK1Threads * kK1BytesPerThread;  // 16384

// 0x80 in every byte of w that equals '\n' (exact, no false positives).
// Three ALU ops per word: the low 7 bits of (w ^ 0x0A..) are zero iff +0x7F leaves bit 7 clear, and
// the byte's own bit 7 (unchanged by the xor) must be clear as well.
//
// A LOP3 takes one immediate, so "(w ^ 0x0A..) & 0x7F.." with two literals costs two instructions;
// SwarConsts keeps the two patterns in registers (made opaque to constant folding) so that the xor-and
// is a single three-register LOP3 (LUT 0x28 = (a ^ b) & c).
struct SwarConsts {
    uint32_t nl, low7;
    __device__ __forceinline__ void init() {
#if defined(__CUDACC__)
        asm volatile("mov.u32 %0, 0x0A0A0A0A;" : "=r"(nl));
        asm volatile("mov.u32 %0, 0x7F7F7F7F;" : "=r"(low7));
#else   // host build of the kernels (tests/host_stub/simt)
        nl = 0x0A0A0A0Au;
        low7 = 0x7F7F7F7Fu;
#endif
    }
};
__device__ __forceinline__ uint32_t newline_flags(uint32_t w, const SwarConsts &k) {
    uint32_t t;
#if defined(__CUDACC__)
    asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(t) : "r"(w), "r"(k.nl), "r"(k.low7));
#else
    t = (w ^ k.nl) & k.low7;
#endif
    t += 0x7F7F7F7Fu;
    return ~(t | w) & 0x80808080u;
}
__device__ __forceinline__ uint32_t newline_flags(uint32_t w) {
    const uint32_t t = ((w ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | w) & 0x80808080u;
}
// 4 flag bits (bit i = byte i flagged) from a word of 0x80 flags: bits 7/15/23/31 -> 28/29/30/31.
__device__ __forceinline__ uint32_t flags_to_nibble(uint32_t f) { return (f * 0x00204081u) >> 28; }

__device__ __forceinline__ uint32_t newline_mask16(const uint4 v) {
    return flags_to_nibble(newline_flags(v.x)) + (flags_to_nibble(newline_flags(v.y)) << 4) +
           (flags_to_nibble(newline_flags(v.z)) << 8) + (flags_to_nibble(newline_flags(v.w)) << 12);
}
__device__ __forceinline__ uint32_t newline_mask16(const uint4 v, const SwarConsts &k) {
    return flags_to_nibble(newline_flags(v.x, k)) + (flags_to_nibble(newline_flags(v.y, k)) << 4) +
           (flags_to_nibble(newline_flags(v.z, k)) << 8) + (flags_to_nibble(newline_flags(v.w, k)) << 12);
}

// swizzled position (in 16-byte chunks) of logical chunk c inside the staging tile
__device__ __forceinline__ uint32_t swz(uint32_t c) { return c ^ ((c >> 2) & 7u); }

__global__ void __launch_bounds__(kK1Threads)
k1_line_index(DevInput in, Control *__restrict__ ctl, int which, unsigned long long *__restrict__ tile_status,
              uint32_t num_tiles, uint32_t epoch) {
    __shared__ uint4 stage[kK1TileBytes / 16];
    __shared__ uint32_t warp_tot[kK1Threads / 32];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_lb[kK1Threads / 32][2];   // look-back scratch

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint4 *__restrict__ src = reinterpret_cast<const uint4 *>(in.data);
    const uint32_t nchunks = (in.nbytes + 15u) >> 4;  // padding bytes are masked below
    SwarConsts swar;
    swar.init();

    while (true) {
        if (tid == 0) s_tile = atomicAdd(&ctl->tile_counter[which], 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= num_tiles) break;
        const uint32_t chunk0 = tile * (kK1TileBytes / 16);

        // coalesced 16-byte loads -> swizzled shared memory
#pragma unroll
        for (int k = 0; k < kK1BytesPerThread / 16; ++k) {
            const uint32_t c = k * kK1Threads + tid;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (chunk0 + c < nchunks) v = __ldcs(src + chunk0 + c);  // streaming: read once
            stage[swz(c)] = v;
        }
        __syncthreads();

        // each thread: 64 contiguous bytes -> 64-bit newline mask
        const uint32_t byte0 = tile * kK1TileBytes + tid * kK1BytesPerThread;
        unsigned long long mask = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint4 v = stage[swz(4 * tid + k)];
            mask |= (unsigned long long)newline_mask16(v, swar) << (16 * k);
        }
        // mask off bytes at or beyond nbytes (zero padding cannot be '\n', but the caller's buffer
        // beyond nbytes inside the last 16-byte chunk may hold stale data)
        if (byte0 + 64u > in.nbytes) {
            const uint32_t valid = in.nbytes > byte0 ? in.nbytes - byte0 : 0u;
            mask &= valid >= 64 ? ~0ull : ((1ull << valid) - 1ull);
        }
        if (byte0 < in.first) mask &= ~0ull << (in.first - byte0);  // first <= 15: only thread 0 of tile 0
        const uint32_t cnt = __popcll(mask);
        const uint32_t incl = warp_incl_scan(cnt, lane);
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kK1Threads / 32; ++w) {
            const uint32_t t = warp_tot[w];
            if (w < wid) wbase += t;
            total += t;
        }
        // wide look-back (one predecessor per thread, 128-byte padded status words): with ~16 K tiles
        // per batch and a thousand CTAs in flight the 32-wide, unpadded walk was most of this kernel
        unsigned long long *const st[2] = {tile_status, nullptr};
        const unsigned long long agg[2] = {total, 0};
        unsigned long long ex[2];
        block_publish(st, tile, agg, 1, epoch, tid);
        block_walk(st, tile, agg, 1, epoch, tid, s_lb, ex);
        if (tid == 0 && tile == num_tiles - 1) ctl->nlines[which] = (uint32_t)(ex[0] + total);
        uint32_t rank = (uint32_t)ex[0] + wbase + incl - cnt;
        while (mask) {
            const int b = __ffsll((long long)mask) - 1;
            mask &= mask - 1;
            if (rank < in.line_cap) in.line_end[rank] = byte0 + (uint32_t)b;
            else ctl->index_overflow = 1u;
            ++rank;
        }
        // s_tile / warp_tot / stage are rewritten only after the next loop-top barrier pair
    }
}

}  // namespace sk
