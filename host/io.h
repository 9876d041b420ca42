// io.h -- host I/O stages of the `sickle` command line.
//
// The reference reads every line through zlib (gzgets, src/GZReader.cpp:77) and formats + writes
// every kept record through an ofstream / gzprintf (src/trim_single.cpp:415-419).  Here whole byte
// ranges move between files and the pinned slots of the C ABI:
//   * ByteSource: plain files are read with parallel pread() straight into the pinned slot (one
//     thread tops out near 5 GB/s out of the page cache, eight reach 35 GB/s on the B200 host);
//     gzip files (detected by magic, as gzopen does) go through one zlib inflate stream, unless they
//     are BGZF (bgzip, Illumina BCL converters, this program's own -g): those are blocked, each block
//     records its compressed size, and the blocks of a batch are inflated in parallel.
//   * ByteSink: an ordered, asynchronous writer thread per output file, so that writing batch k
//     overlaps the GPU work of batch k+1 and the read of batch k+2.  `-g` output is BGZF, deflated
//     in parallel (to any gunzip an ordinary multi-member file); the reference's `-g` is unusable
//     (payload passed as a printf format), so only validity matters.
#ifndef SICKLE_B200_HOST_IO_H
#define SICKLE_B200_HOST_IO_H

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace host {

// Worker threads per I/O call: SICKLE_B200_IO_THREADS, default min(8, hardware threads).
int io_threads();
// Threads for inflate / deflate: SICKLE_B200_ZIP_THREADS, default all hardware threads.
int zip_threads();
double now_s();

class ByteSource {
public:
    ~ByteSource();
    bool open(const char *path);
    // Read up to n bytes; returns the count (less than n only at end of file), -1 on error.
    long long read(char *dst, unsigned long long n);
    bool gzip() const { return gz_ != nullptr || bgzf_; }
    bool bgzf() const { return bgzf_; }
    unsigned long long file_size() const { return size_; }
    double read_seconds() const { return read_s_; }

private:
    int fd_ = -1;
    void *gz_ = nullptr;
    unsigned long long size_ = 0, pos_ = 0;
    bool seekable_ = false;
    double read_s_ = 0;
    // BGZF input: blocks are inflated in parallel straight into the caller's buffer
    long long read_bgzf(char *dst, unsigned long long n);
    bool bgzf_ = false;
    unsigned long long cpos_ = 0;            // file offset of the next block
    std::vector<unsigned char> spill_;       // tail of a block that did not fit the previous read()
    size_t spill_pos_ = 0;
};

class ByteSink {
public:
    ~ByteSink();
    bool open(const char *path, bool gzip);
    bool is_open() const { return fd_ >= 0; }
    // Queue n bytes for writing (in call order).  The bytes must stay untouched until wait() on the
    // returned ticket has returned.
    unsigned long long write_async(const char *src, unsigned long long n);
    // Block until every write up to `ticket` is on its way to the file; false if any write failed.
    bool wait(unsigned long long ticket);
    bool write(const char *src, unsigned long long n) { return wait(write_async(src, n)); }
    // Block until every queued write has been carried out (or has failed): after this the writer thread
    // holds no pointer into the caller's buffers.  Needed before those buffers are freed on an error path.
    void drain();
    // Drain and close; false if any write failed.
    bool close();
    double busy_seconds() const { return busy_s_; }

private:
    struct Job { const char *src; unsigned long long n; };
    void run();
    bool put(const char *src, unsigned long long n);
    bool put_gzip(const char *src, unsigned long long n);
    int fd_ = -1;
    bool gzip_ = false, mmap_ = false;
    int map_mode_ = 0;
    unsigned long long bytes_in_ = 0, pos_ = 0;
    std::thread worker_;
    std::mutex mu_;
    std::condition_variable cv_job_, cv_done_;
    std::deque<Job> jobs_;
    unsigned long long submitted_ = 0, completed_ = 0;
    bool stop_ = false, failed_ = false;
    double busy_s_ = 0;
};

}  // namespace host

#endif
