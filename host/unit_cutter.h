// unit_cutter.h -- where a byte buffer can be cut so that every piece holds whole FASTQ records (pairs).
//
// Batches that run on different GPUs at the same time must be independent: each has to start on the
// first line of a record (4 lines; 8 for an interleaved pair).  Which of the four lines a given line
// is cannot be told from its content ('@' and '+' are valid quality characters), so it is counted
// (SURVEY.md 8-e): the newlines of the buffer are counted a megabyte at a time on several threads, the
// per-piece counts are kept, and "the byte after newline number m" is then a prefix walk over the
// pieces plus a memchr walk inside one of them.  The reference gets the same grouping from reading
// line by line (src/GZReader.cpp:59-132, 4 / 8 lines per entry: src/Batch.cpp:6-21).
#ifndef SICKLE_B200_HOST_UNIT_CUTTER_H
#define SICKLE_B200_HOST_UNIT_CUTTER_H

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "io.h"
#include "ref_batcher.h"   // count_newlines

namespace host {

class LineMap {
public:
    static constexpr unsigned long long kPiece = 1ull << 20;

    // Count the newlines of buf[0, n).
    void build(const char *buf, unsigned long long n) {
        buf_ = buf;
        n_ = n;
        const unsigned long long np = (n + kPiece - 1) / kPiece;
        cnt_.assign((size_t)np, 0);
        auto one = [&](unsigned long long k) {
            const unsigned long long lo = k * kPiece;
            cnt_[(size_t)k] = count_newlines(buf + lo, std::min(kPiece, n - lo));
        };
        const int nt = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((unsigned long long)io_threads(), np / 4));
        if (nt <= 1) {
            for (unsigned long long k = 0; k < np; ++k) one(k);
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t)
                th.emplace_back([&, t] {
                    for (unsigned long long k = (unsigned long long)t; k < np; k += (unsigned long long)nt) one(k);
                });
            for (auto &x : th) x.join();
        }
        total_ = 0;
        for (unsigned long long c : cnt_) total_ += c;
    }

    unsigned long long lines() const { return total_; }

    // Offset of the byte after newline number m (1-based; m == 0 gives 0).  m <= lines().
    unsigned long long after_line(unsigned long long m) const {
        if (m == 0) return 0;
        unsigned long long seen = 0, k = 0;
        while (seen + cnt_[(size_t)k] < m) seen += cnt_[(size_t)k++];
        unsigned long long p = k * kPiece;
        const unsigned long long e = std::min(n_, p + kPiece);
        // skip 64 KiB blocks of this piece that end before the wanted newline, then walk lines
        while (e - p > (1ull << 16)) {
            const unsigned long long c = count_newlines(buf_ + p, 1ull << 16);
            if (seen + c >= m) break;
            seen += c;
            p += 1ull << 16;
        }
        while (true) {
            const char *q = (const char *)memchr(buf_ + p, '\n', (size_t)(e - p));
            p = (unsigned long long)(q - buf_) + 1;   // q is never null: the piece holds the newline
            if (++seen == m) return p;
        }
    }

private:
    const char *buf_ = nullptr;
    unsigned long long n_ = 0, total_ = 0;
    std::vector<unsigned long long> cnt_;
};

// One input stream cut into whole-unit batches.  fill() moves the carried bytes of the previous
// batch to the front of `dst`, tops the buffer up from the source and counts its lines; after the
// caller has decided how many units the batch holds, cut() fixes the end of the batch and what is
// carried over.  The carried bytes stay in the previous buffer until the next fill().
class UnitStream {
public:
    UnitStream(ByteSource *src, int lines_per_unit) : src_(src), lpu_((unsigned long long)lines_per_unit) {}

    // false on a read error
    bool fill(char *dst, unsigned long long cap) {
        if (carry_len_ > cap) return false;
        if (carry_len_) memmove(dst, carry_ptr_, (size_t)carry_len_);
        n_ = carry_len_;
        if (!eof_ && n_ < cap) {
            const long long r = src_->read(dst + n_, cap - n_);
            if (r < 0) return false;
            if ((unsigned long long)r < cap - n_) eof_ = true;
            n_ += (unsigned long long)r;
        }
        // an unterminated last line loses its final character, as in the reference (src/GZReader.cpp:81-88)
        if (eof_ && n_ > 0 && dst[n_ - 1] != '\n') dst[n_ - 1] = '\n';
        buf_ = dst;
        full_ = n_ == cap;
        map_.build(dst, n_);
        return true;
    }
    unsigned long long units() const { return map_.lines() / lpu_; }
    unsigned long long bytes() const { return n_; }
    bool eof() const { return eof_; }
    bool full() const { return full_; }

    // The batch is the first `units` units of the buffer; returns its size in bytes.
    unsigned long long cut(unsigned long long units) {
        const unsigned long long end = map_.after_line(units * lpu_);
        carry_ptr_ = buf_ + end;
        carry_len_ = n_ - end;
        return end;
    }

private:
    ByteSource *src_;
    unsigned long long lpu_;
    LineMap map_;
    const char *buf_ = nullptr, *carry_ptr_ = nullptr;
    unsigned long long n_ = 0, carry_len_ = 0;
    bool eof_ = false, full_ = false;
};

}  // namespace host

#endif
