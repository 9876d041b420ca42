// ref_batcher.h -- cuts a byte stream into the batches the reference's GZReader::read_lines would form
// (src/GZReader.cpp:59-132): lines are added until the sum of their lengths (without '\n') reaches
// batch_len -- at least one new line per batch --, a multiple of `minlines` lines is kept and the rest is
// carried into the next batch.  Only needed to reproduce the reference's `-a N` output order, which is
// defined per reference batch.
//
// The reference does this one gzgets() at a time.  Here the stream is read in large blocks straight
// into the destination (a pinned slot); with f(pos) = batch_len - (bytes - newlines) over the new
// lines, f never increases, so the batch ends at the first line end at or after the first position
// where f <= 0: newlines are counted a block at a time (8 bytes per step) and only the block in which
// f crosses zero is walked line by line.
#ifndef SICKLE_B200_HOST_REF_BATCHER_H
#define SICKLE_B200_HOST_REF_BATCHER_H

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "io.h"

namespace host {

inline unsigned long long count_newlines(const char *p, unsigned long long n) {
    unsigned long long c = 0, i = 0;
#if defined(__SSE2__)
    // 16 bytes per compare; per-byte counters (0xFF = -1 per hit) summed every 255 steps by psadbw
    const __m128i nl = _mm_set1_epi8('\n'), zero = _mm_setzero_si128();
    while (i + 64 <= n) {
        __m128i acc0 = zero, acc1 = zero;
        const unsigned long long stop = std::min<unsigned long long>(n - 63, i + 64ull * 127ull);
        for (; i < stop; i += 64) {
            const __m128i a = _mm_loadu_si128((const __m128i *)(p + i)), b = _mm_loadu_si128((const __m128i *)(p + i + 16));
            const __m128i d = _mm_loadu_si128((const __m128i *)(p + i + 32)), e = _mm_loadu_si128((const __m128i *)(p + i + 48));
            acc0 = _mm_sub_epi8(acc0, _mm_add_epi8(_mm_cmpeq_epi8(a, nl), _mm_cmpeq_epi8(b, nl)));
            acc1 = _mm_sub_epi8(acc1, _mm_add_epi8(_mm_cmpeq_epi8(d, nl), _mm_cmpeq_epi8(e, nl)));
        }
        const __m128i s = _mm_add_epi64(_mm_sad_epu8(acc0, zero), _mm_sad_epu8(acc1, zero));
        c += (unsigned long long)_mm_cvtsi128_si64(s) + (unsigned long long)_mm_cvtsi128_si64(_mm_unpackhi_epi64(s, s));
    }
#endif
    while (i + 8 <= n) {
        // per-byte counters in one 64-bit word, summed every 255 steps (no popcount instruction needed)
        uint64_t lanes = 0;
        const unsigned long long stop = std::min<unsigned long long>(n - 7, i + 8ull * 255ull);
        for (; i < stop; i += 8) {
            uint64_t w;
            memcpy(&w, p + i, 8);
            const uint64_t x = w ^ 0x0A0A0A0A0A0A0A0AULL;                 // zero byte <=> '\n'
            lanes += (~(((x & 0x7F7F7F7F7F7F7F7FULL) + 0x7F7F7F7F7F7F7F7FULL) | x) & 0x8080808080808080ULL) >> 7;
        }
        const uint64_t pairs = (lanes & 0x00FF00FF00FF00FFULL) + ((lanes >> 8) & 0x00FF00FF00FF00FFULL);   // 4 x 16 bit
        c += (pairs * 0x0001000100010001ULL) >> 48;
    }
    for (; i < n; ++i) c += p[i] == '\n';
    return c;
}

class RefBatcher {
public:
    RefBatcher(ByteSource *src, long long batch_len, int minlines) : src_(src), batch_len_(batch_len), minlines_(minlines) {}

    // Fills dst (capacity cap) with the next batch; returns its size, 0 at the end, -1 if it does not fit
    // (or on a read error).
    long long next(char *dst, unsigned long long cap) {
        if (done_) return 0;
        if (carry_.size() > cap) return -1;
        if (!carry_.empty()) memcpy(dst, carry_.data(), carry_.size());
        unsigned long long n = carry_.size();            // bytes in dst
        const unsigned long long c0 = carry_line_bytes_; // the carried complete lines come first ...
        const unsigned long long carried_lines = carry_lines_;
        carry_.clear();
        // f at the start of the new lines: batch_len minus the carried lines' lengths
        long long f = batch_len_ - (long long)(c0 - carried_lines);
        unsigned long long scanned = c0;                 // f is known up to here
        unsigned long long new_lines = 0;                // newlines in [c0, scanned)
        unsigned long long cross = 0;                    // end (exclusive) of the line that ends the batch
        bool found = false;
        while (!found) {
            // ---- a long unscanned stretch: count its newlines on several threads, a megabyte per piece,
            // and skip every piece that cannot hold the end of the batch
            if (n - scanned >= (8ull << 20)) {
                const unsigned long long piece = 1ull << 20;
                const unsigned long long npieces = (n - scanned) / piece;
                std::vector<unsigned long long> cnt((size_t)npieces);
                const int nt = std::max(1, std::min(io_threads(), (int)(npieces / 4)));
                std::vector<std::thread> th;
                for (int t = 0; t < nt; ++t)
                    th.emplace_back([&, t] {
                        for (unsigned long long k = (unsigned long long)t; k < npieces; k += (unsigned long long)nt)
                            cnt[(size_t)k] = count_newlines(dst + scanned + k * piece, piece);
                    });
                for (auto &x : th) x.join();
                for (unsigned long long k = 0; k < npieces; ++k) {
                    const long long f_end = f - (long long)(piece - cnt[(size_t)k]);
                    if (f_end <= 0 && cnt[(size_t)k] != 0) break;      // the batch may end in this piece: walk it below
                    f = f_end; scanned += piece; new_lines += cnt[(size_t)k];
                }
            }
            // ---- scan what is there, a block at a time
            while (scanned < n && !found) {
                const unsigned long long blk = std::min<unsigned long long>(n - scanned, 1u << 16);
                const unsigned long long nl = count_newlines(dst + scanned, blk);
                const long long f_end = f - (long long)(blk - nl);
                if (f_end > 0 || nl == 0) {              // no line end of this block can end the batch
                    f = f_end; scanned += blk; new_lines += nl;
                    continue;
                }
                // f reaches 0 inside this block (or was there already): walk its lines
                unsigned long long p = scanned;
                const unsigned long long e = scanned + blk;
                while (p < e) {
                    const char *q = (const char *)memchr(dst + p, '\n', (size_t)(e - p));
                    if (!q) { f -= (long long)(e - p); p = e; break; }
                    const unsigned long long le = (unsigned long long)(q - dst) + 1;
                    f -= (long long)(le - p - 1);
                    p = le;
                    ++new_lines;
                    if (f <= 0) { found = true; cross = le; break; }
                }
                scanned = p;
            }
            if (found) break;
            // ---- need more bytes
            if (src_done_) break;
            if (n >= cap) return -1;
            unsigned long long want = (unsigned long long)std::max<long long>(f, 0);
            want += want / 16 + (1u << 16);
            want = std::min(want, cap - n);
            const long long r = src_->read(dst + n, want);
            if (r < 0) return -1;
            if ((unsigned long long)r < want) {
                src_done_ = true;
                // an unterminated last line loses its final character, as in the reference (GZReader.cpp:81-88)
                if (n + (unsigned long long)r > 0 && dst[n + (unsigned long long)r - 1] != '\n') dst[n + (unsigned long long)r - 1] = '\n';
            }
            n += (unsigned long long)r;
        }
        if (!found) {            // the stream ended first: every line read belongs to this (last) batch
            done_ = true;
            cross = n;           // n ends with '\n' (patched above) or is 0
        }
        const unsigned long long lines = carried_lines + new_lines;
        const unsigned long long extra = lines % (unsigned long long)minlines_;
        if (lines == extra) { done_ = true; return 0; }
        // drop the last `extra` lines: walk back over their newlines
        unsigned long long end = cross;
        for (unsigned long long k = 0; k < extra; ++k) {
            const void *q = end >= 2 ? memrchr(dst, '\n', (size_t)(end - 1)) : nullptr;
            end = q ? (unsigned long long)((const char *)q - dst) + 1 : 0;
        }
        carry_.assign(dst + end, dst + n);
        carry_line_bytes_ = cross - end;
        carry_lines_ = extra;
        return (long long)end;
    }

private:
    ByteSource *src_;
    long long batch_len_;
    int minlines_;
    bool done_ = false, src_done_ = false;
    std::vector<char> carry_;                  // bytes read beyond the previous batch: carried lines, then unscanned bytes
    unsigned long long carry_line_bytes_ = 0;  // size of the carried complete lines at the front of carry_
    unsigned long long carry_lines_ = 0;
};

}  // namespace host

#endif
