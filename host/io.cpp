// io.cpp -- see io.h.
#include "io.h"

#include <fcntl.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace host {

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int io_threads() {
    static const int n = [] {
        const char *e = getenv("SICKLE_B200_IO_THREADS");
        int v = e && *e ? atoi(e) : 0;
        if (v <= 0) v = (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
        return std::min(v, 64);
    }();
    return n;
}

// (De)compression is CPU work, not memory copying: every hardware thread, SICKLE_B200_ZIP_THREADS.
int zip_threads() {
    static const int n = [] {
        const char *e = getenv("SICKLE_B200_ZIP_THREADS");
        int v = e && *e ? atoi(e) : 0;
        if (v <= 0) v = (int)std::max(1u, std::thread::hardware_concurrency());
        return std::min(v, 256);
    }();
    return n;
}

// zlib level of -g output: SICKLE_B200_GZIP_LEVEL (1..9), default 4 -- about twice the speed of zlib's
// default 6 for ~5 % more bytes on FASTQ (the reference's -g never produced a usable file, so there is
// no byte compatibility to keep).
int gzip_level() {
    static const int n = [] {
        const char *e = getenv("SICKLE_B200_GZIP_LEVEL");
        const int v = e && *e ? atoi(e) : 4;
        return v < 1 ? 1 : (v > 9 ? 9 : v);
    }();
    return n;
}

namespace {

// Run f(offset, length) over [0, n) on up to io_threads() threads, each piece at least min_piece.
template <class F>
bool parallel_ranges(unsigned long long n, unsigned long long min_piece, F f) {
    int nt = (int)std::min<unsigned long long>((unsigned long long)io_threads(), std::max<unsigned long long>(1, n / min_piece));
    if (nt <= 1) return f(0ull, n);
    const unsigned long long piece = ((n + (unsigned long long)nt - 1) / (unsigned long long)nt + 4095ull) & ~4095ull;
    std::vector<std::thread> th;
    std::vector<char> ok((size_t)nt, 1);
    for (int t = 0; t < nt; ++t) {
        const unsigned long long lo = std::min(n, piece * (unsigned long long)t), hi = std::min(n, piece * (unsigned long long)(t + 1));
        if (hi <= lo) continue;
        th.emplace_back([&ok, t, lo, hi, &f] { ok[(size_t)t] = f(lo, hi - lo) ? 1 : 0; });
    }
    for (auto &x : th) x.join();
    return std::all_of(ok.begin(), ok.end(), [](char c) { return c != 0; });
}

unsigned get_le16(const unsigned char *p) { return (unsigned)p[0] | ((unsigned)p[1] << 8); }
unsigned long get_le32(const unsigned char *p) { return (unsigned long)get_le16(p) | ((unsigned long)get_le16(p + 2) << 16); }

// If p[0, avail) starts with a complete BGZF block header, return the block's total size, else 0.
// (RFC 1952 member with FEXTRA holding a 'B','C' subfield of length 2 = block size - 1.)
unsigned bgzf_block_size(const unsigned char *p, unsigned long long avail) {
    if (avail < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
    const unsigned xlen = get_le16(p + 10);
    if (avail < 12ull + xlen) return 0;
    for (unsigned o = 0; o + 4 <= xlen;) {
        const unsigned slen = get_le16(p + 12 + o + 2);
        if (p[12 + o] == 'B' && p[12 + o + 1] == 'C' && slen == 2 && o + 6 <= xlen) return get_le16(p + 12 + o + 4) + 1u;
        o += 4 + slen;
    }
    return 0;
}

bool write_all(int fd, const char *src, unsigned long long n) {
    unsigned long long done = 0;
    while (done < n) {
        const ssize_t r = ::write(fd, src + done, (size_t)std::min<unsigned long long>(n - done, 1ull << 30));
        if (r < 0) return false;
        done += (unsigned long long)r;
    }
    return true;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
ByteSource::~ByteSource() {
    if (gz_) gzclose((gzFile)gz_);
    else if (fd_ >= 0) ::close(fd_);
}

bool ByteSource::open(const char *path) {
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) == 0) {
        size_ = (unsigned long long)st.st_size;
        seekable_ = S_ISREG(st.st_mode);
    }
    unsigned char magic[64];
    memset(magic, 0, sizeof magic);
    const ssize_t got = seekable_ ? ::pread(fd, magic, sizeof magic, 0) : 0;
    fd_ = fd;
    if (got >= 18 && bgzf_block_size(magic, (unsigned long long)got) != 0 && !getenv("SICKLE_B200_NO_BGZF")) {
        bgzf_ = true;
    } else if (got >= 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
        gzFile g = gzdopen(fd, "rb");
        if (!g) { ::close(fd); fd_ = -1; return false; }
        gzbuffer(g, 1u << 20);
        gz_ = g;
    } else {
#ifdef POSIX_FADV_SEQUENTIAL
        if (seekable_) posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
    }
    return true;
}

long long ByteSource::read(char *dst, unsigned long long n) {
    const double t0 = now_s();
    unsigned long long done = 0;
    if (bgzf_) {
        const long long r = read_bgzf(dst, n);
        read_s_ += now_s() - t0;
        return r;
    }
    if (gz_) {
        while (done < n) {
            const unsigned want = (unsigned)std::min<unsigned long long>(n - done, 1u << 30);
            const int r = gzread((gzFile)gz_, dst + done, want);
            if (r < 0) return -1;
            if (r == 0) break;
            done += (unsigned long long)r;
        }
    } else if (seekable_) {
        // the file may grow or shrink under us: trust pread's return values, not size_
        const unsigned long long want = std::min(n, size_ > pos_ ? size_ - pos_ : 0ull);
        const int fd = fd_;
        const unsigned long long base = pos_;
        std::atomic<bool> short_read{false};
        const bool ok = parallel_ranges(want, 4ull << 20, [&](unsigned long long off, unsigned long long len) {
            unsigned long long d = 0;
            while (d < len) {
                const ssize_t r = ::pread(fd, dst + off + d, (size_t)std::min<unsigned long long>(len - d, 1ull << 30), (off_t)(base + off + d));
                if (r < 0) return false;
                if (r == 0) { short_read = true; return false; }
                d += (unsigned long long)r;
            }
            return true;
        });
        if (!ok && !short_read) return -1;
        if (short_read) {   // truncated while reading: fall back to what a sequential reader would have seen
            done = 0;
            while (done < want) {
                const ssize_t r = ::pread(fd, dst + done, (size_t)(want - done), (off_t)(base + done));
                if (r < 0) return -1;
                if (r == 0) break;
                done += (unsigned long long)r;
            }
        } else {
            done = want;
        }
        // bytes appended after open(): keep reading sequentially
        while (done < n && done == want) {
            const ssize_t r = ::pread(fd, dst + done, (size_t)std::min<unsigned long long>(n - done, 1ull << 30), (off_t)(base + done));
            if (r <= 0) break;
            done += (unsigned long long)r;
        }
        pos_ += done;
    } else {
        while (done < n) {
            const ssize_t r = ::read(fd_, dst + done, (size_t)std::min<unsigned long long>(n - done, 1ull << 30));
            if (r < 0) return -1;
            if (r == 0) break;
            done += (unsigned long long)r;
        }
    }
    read_s_ += now_s() - t0;
    return (long long)done;
}


// One raw-deflate block -> exactly `isize` bytes at out; checks the CRC like gzread does.
static bool bgzf_inflate_block(z_stream &z, const unsigned char *blk, unsigned bsize, unsigned char *out, unsigned long isize) {
    const unsigned xlen = get_le16(blk + 10);
    const unsigned hdr = 12 + xlen;
    if (bsize < hdr + 8) return false;
    if (inflateReset(&z) != Z_OK) return false;
    z.next_in = (Bytef *)(blk + hdr);
    z.avail_in = bsize - hdr - 8;
    z.next_out = out;
    z.avail_out = (uInt)isize;
    const int rc = inflate(&z, Z_FINISH);
    if (rc != Z_STREAM_END || z.avail_out != 0) return false;
    return crc32(crc32(0L, Z_NULL, 0), out, (uInt)isize) == get_le32(blk + bsize - 8);
}

long long ByteSource::read_bgzf(char *dst, unsigned long long n) {
    unsigned long long done = 0;
    while (done < n) {
        if (spill_pos_ < spill_.size()) {   // rest of the block that straddled the previous call
            const size_t k = (size_t)std::min<unsigned long long>(n - done, spill_.size() - spill_pos_);
            memcpy(dst + done, spill_.data() + spill_pos_, k);
            spill_pos_ += k;
            done += k;
            continue;
        }
        if (cpos_ >= size_) break;
        // a window of compressed bytes (FASTQ deflates ~3-4x; whatever does not fit waits for the next round)
        const unsigned long long room = n - done;
        const unsigned long long wlen = std::min(size_ - cpos_, std::max<unsigned long long>(room / 2, 1ull << 17));
        std::vector<unsigned char> win((size_t)wlen);
        {
            const int fd = fd_;
            const unsigned long long base = cpos_;
            unsigned char *w = win.data();
            const bool ok = parallel_ranges(wlen, 4ull << 20, [&](unsigned long long off, unsigned long long len) {
                unsigned long long d = 0;
                while (d < len) {
                    const ssize_t r = ::pread(fd, w + off + d, (size_t)(len - d), (off_t)(base + off + d));
                    if (r <= 0) return false;
                    d += (unsigned long long)r;
                }
                return true;
            });
            if (!ok) return -1;
        }
        // block table of the window: (offset, size, uncompressed size, destination offset)
        struct Blk { unsigned long long off; unsigned size; unsigned long isize; unsigned long long dst; };
        std::vector<Blk> blks;
        unsigned long long o = 0, total = 0;
        bool straddle = false;
        while (o < wlen) {
            const unsigned bs = bgzf_block_size(win.data() + o, wlen - o);
            if (bs == 0) {
                if (wlen - o >= 18 || cpos_ + wlen >= size_) return -1;   // not a BGZF block / truncated file
                break;                                                    // header cut by the window
            }
            if (o + bs > wlen) {
                if (cpos_ + wlen >= size_) return -1;                     // truncated file
                break;
            }
            const unsigned long isize = get_le32(win.data() + o + bs - 4);
            if (isize > 0x10000) return -1;
            if (total + isize > room) { straddle = true; break; }
            blks.push_back(Blk{o, bs, isize, total});
            total += isize;
            o += bs;
        }
        if (!blks.empty()) {
            std::atomic<bool> bad{false};
            const size_t nb = blks.size();
            const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)zip_threads(), nb / 8));
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t) {
                th.emplace_back([&, t] {
                    z_stream z;
                    memset(&z, 0, sizeof z);
                    if (inflateInit2(&z, -15) != Z_OK) { bad = true; return; }
                    for (size_t i = nb * (size_t)t / (size_t)nt; i < nb * (size_t)(t + 1) / (size_t)nt && !bad; ++i)
                        if (blks[i].isize && !bgzf_inflate_block(z, win.data() + blks[i].off, blks[i].size,
                                                                 (unsigned char *)dst + done + blks[i].dst, blks[i].isize))
                            bad = true;
                    inflateEnd(&z);
                });
            }
            for (auto &x : th) x.join();
            if (bad) return -1;
            done += total;
            cpos_ += o;
        }
        if (straddle) {   // the next block is larger than what is left of dst: inflate it aside
            const unsigned bs = bgzf_block_size(win.data() + o, wlen - o);
            const unsigned long isize = get_le32(win.data() + o + bs - 4);
            spill_.resize(isize);
            spill_pos_ = 0;
            z_stream z;
            memset(&z, 0, sizeof z);
            if (inflateInit2(&z, -15) != Z_OK) return -1;
            const bool ok = bgzf_inflate_block(z, win.data() + o, bs, spill_.data(), isize);
            inflateEnd(&z);
            if (!ok) return -1;
            cpos_ += bs;
        } else if (blks.empty()) {
            return -1;   // no progress possible (cannot happen with room >= 64 KiB windows; defensive)
        }
    }
    return (long long)done;
}

// ---------------------------------------------------------------------------------------------
ByteSink::~ByteSink() { close(); }

static void on_sigbus(int) {
    static const char msg[] = "****Error: write failed (output file system full?)\n\n";
    if (::write(2, msg, sizeof msg - 1) < 0) {}
    _exit(EXIT_FAILURE);
}

bool ByteSink::open(const char *path, bool gzip) {
    fd_ = ::open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (fd_ < 0) fd_ = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);   // write-only targets
    if (fd_ < 0) return false;
    gzip_ = gzip;
    pos_ = 0;
    // SICKLE_B200_MMAP_OUT: plain regular files can be extended and filled through a shared mapping by
    // io_threads() threads (page-cache writes through write(2) serialise on the inode lock: 3.4 GB/s
    // into tmpfs on the B200 host against 6.9 GB/s this way).
    //   1 = map after ftruncate: a full disk then surfaces as SIGBUS (reported and turned into exit 1 here);
    //   2 (default) = reserve the range with fallocate first: a full disk is an ordinary error (the writer
    //       falls back to write(2), which reports it) before any byte is copied.  Measured on the B200 host,
    //       24 M reads to tmpfs: write(2) 3.9 s wall, mapped 2.6 s, fallocate + mapped 1.9 s;
    //   0 = write(2).
    struct stat st;
    const char *e = getenv("SICKLE_B200_MMAP_OUT");
    map_mode_ = e ? atoi(e) : 2;
    mmap_ = !gzip && map_mode_ != 0 && fstat(fd_, &st) == 0 && S_ISREG(st.st_mode) && (fcntl(fd_, F_GETFL) & O_ACCMODE) == O_RDWR;
    if (mmap_ && map_mode_ == 1) signal(SIGBUS, on_sigbus);
    stop_ = failed_ = false;
    submitted_ = completed_ = 0;
    bytes_in_ = 0;
    worker_ = std::thread([this] { run(); });
    return true;
}

unsigned long long ByteSink::write_async(const char *src, unsigned long long n) {
    std::lock_guard<std::mutex> lk(mu_);
    jobs_.push_back(Job{src, n});
    ++submitted_;
    cv_job_.notify_one();
    return submitted_;
}

bool ByteSink::wait(unsigned long long ticket) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return completed_ >= ticket; });
    return !failed_;
}

void ByteSink::drain() {
    if (fd_ < 0) return;
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return completed_ >= submitted_; });
}

void ByteSink::run() {
    while (true) {
        Job j;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_job_.wait(lk, [&] { return stop_ || !jobs_.empty(); });
            if (jobs_.empty()) return;
            j = jobs_.front();
            jobs_.pop_front();
        }
        const double t0 = now_s();
        const bool ok = failed_ ? false : (gzip_ ? put_gzip(j.src, j.n) : put(j.src, j.n));
        busy_s_ += now_s() - t0;
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (!ok) failed_ = true;
            ++completed_;
        }
        cv_done_.notify_all();
    }
}

bool ByteSink::put(const char *src, unsigned long long n) {
    bytes_in_ += n;
    if (!mmap_) return write_all(fd_, src, n);
    if (n >= (8ull << 20) && (map_mode_ == 2 ? posix_fallocate(fd_, (off_t)pos_, (off_t)n) == 0 : ftruncate(fd_, (off_t)(pos_ + n)) == 0)) {
        const unsigned long long map_off = pos_ & ~4095ull, lead = pos_ - map_off;
        void *m = mmap(nullptr, (size_t)(lead + n), PROT_READ | PROT_WRITE, MAP_SHARED, fd_, (off_t)map_off);
        if (m != MAP_FAILED) {
            char *d = (char *)m + lead;
            parallel_ranges(n, 4ull << 20, [&](unsigned long long off, unsigned long long len) {
                memcpy(d + off, src + off, (size_t)len);
                return true;
            });
            munmap(m, (size_t)(lead + n));
            pos_ += n;
            return true;
        }
    }
    unsigned long long done = 0;
    while (done < n) {
        const ssize_t r = ::pwrite(fd_, src + done, (size_t)std::min<unsigned long long>(n - done, 1ull << 30), (off_t)(pos_ + done));
        if (r < 0) return false;
        done += (unsigned long long)r;
    }
    pos_ += n;
    return true;
}

// BGZF (the blocked gzip of htslib/bgzip, also what Illumina's converters write): every <= 65280
// input bytes become one gzip member that records its own compressed size, so members are deflated
// in parallel here and can be inflated in parallel by ByteSource (or anything else; to gunzip it is
// an ordinary multi-member file).
namespace {
constexpr unsigned kBgzfIn = 0xff00;      // uncompressed bytes per block
constexpr unsigned kBgzfHeader = 18, kBgzfFooter = 8;
const unsigned char kBgzfEof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};

void put_le16(unsigned char *p, unsigned v) { p[0] = (unsigned char)v; p[1] = (unsigned char)(v >> 8); }
void put_le32(unsigned char *p, unsigned long v) { put_le16(p, (unsigned)(v & 0xffff)); put_le16(p + 2, (unsigned)(v >> 16)); }
// Append the BGZF blocks of src[0, n) to out.
bool bgzf_deflate(const char *src, unsigned long long n, std::vector<unsigned char> &out) {
    z_stream z;
    memset(&z, 0, sizeof z);
    if (deflateInit2(&z, gzip_level(), Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    bool ok = true;
    for (unsigned long long off = 0; off < n && ok; off += kBgzfIn) {
        const unsigned len = (unsigned)std::min<unsigned long long>(kBgzfIn, n - off);
        const size_t at = out.size();
        out.resize(at + 0x10000);
        unsigned char *b = out.data() + at;
        static const unsigned char head[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
        memcpy(b, head, 16);
        deflateReset(&z);
        z.next_in = (Bytef *)(src + off);
        z.avail_in = len;
        z.next_out = b + kBgzfHeader;
        z.avail_out = 0x10000 - kBgzfHeader - kBgzfFooter;
        if (deflate(&z, Z_FINISH) != Z_STREAM_END) { ok = false; break; }
        const unsigned clen = (unsigned)(0x10000 - kBgzfHeader - kBgzfFooter - z.avail_out);
        const unsigned total = kBgzfHeader + clen + kBgzfFooter;
        put_le16(b + 16, total - 1);
        put_le32(b + kBgzfHeader + clen, crc32(crc32(0L, Z_NULL, 0), (const Bytef *)(src + off), len));
        put_le32(b + kBgzfHeader + clen + 4, len);
        out.resize(at + total);
    }
    deflateEnd(&z);
    return ok;
}
}  // namespace

bool ByteSink::put_gzip(const char *src, unsigned long long n) {
    bytes_in_ += n;
    const unsigned long long super = 1024ull * kBgzfIn;   // ~64 MB of input per round
    for (unsigned long long s0 = 0; s0 < n; s0 += super) {
        const unsigned long long sn = std::min(super, n - s0);
        const unsigned long long nblocks = (sn + kBgzfIn - 1) / kBgzfIn;
        const int nt = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((unsigned long long)zip_threads(), nblocks / 4));
        const unsigned long long per = (nblocks + (unsigned long long)nt - 1) / (unsigned long long)nt * kBgzfIn;
        std::vector<std::vector<unsigned char>> out((size_t)nt);
        std::vector<char> ok((size_t)nt, 1);
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) {
            const unsigned long long lo = std::min(sn, per * (unsigned long long)t), hi = std::min(sn, per * (unsigned long long)(t + 1));
            if (hi <= lo) continue;
            th.emplace_back([&, t, lo, hi] { ok[(size_t)t] = bgzf_deflate(src + s0 + lo, hi - lo, out[(size_t)t]) ? 1 : 0; });
        }
        for (auto &x : th) x.join();
        for (int t = 0; t < nt; ++t) {
            if (!ok[(size_t)t]) return false;
            if (!out[(size_t)t].empty() && !write_all(fd_, (const char *)out[(size_t)t].data(), out[(size_t)t].size())) return false;
        }
    }
    return true;
}

bool ByteSink::close() {
    if (fd_ < 0) return !failed_;
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_job_.notify_all();
    if (worker_.joinable()) worker_.join();
    bool ok = !failed_;
    if (gzip_ && ok) ok = write_all(fd_, (const char *)kBgzfEof, sizeof kBgzfEof);   // BGZF end marker (an empty member)
    if (::close(fd_) != 0) ok = false;
    fd_ = -1;
    return ok;
}

}  // namespace host
